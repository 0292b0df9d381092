/*
 * eosvr.h -- C ABI of the B200-native test-time episodic hot path
 *            (video-segment augmentation + one-shot episode scoring).
 *
 * The reference (lovelyqian/Embodied-One-Shot-Video-Recognition) is pure Python and has
 * no FFI; this boundary is new.  Each entry point replaces the arithmetic of the
 * reference lines cited beside it (paths relative to the reference checkout).  The
 * reference-side binding is a ctypes stub; see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative EOSVR_E* code; the message of
 *     the last failure on the calling thread is eosvr_last_error();
 *   - all d_* pointers are DEVICE pointers owned by the caller (e.g. tensor.data_ptr());
 *     the library allocates device memory only inside the opaque handles;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); no call
 *     synchronises the host except where stated; outputs are valid once the stream has
 *     reached the end of the call's work;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef EOSVR_H_
#define EOSVR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EOSVR_VERSION 100

/* error codes */
#define EOSVR_OK          0
#define EOSVR_EINVAL     -1   /* bad argument (shape, alignment, enum) */
#define EOSVR_ECUDA      -2   /* CUDA runtime / driver error           */
#define EOSVR_ENOMEM     -3   /* device allocation failed              */
#define EOSVR_EUNSUPPORTED -4 /* device is not sm_100                  */
#define EOSVR_ESHAPE     -5   /* valid arguments, but a shape this entry point has no kernel for */

/* feature storage types */
#define EOSVR_F32  0          /* float32 rows (the reference's dtype, network_test.py:187-189) */
#define EOSVR_BF16 1          /* bfloat16 rows: half the HBM and PCIe bytes.  All arithmetic runs on the exactly
                               * upcast values, so results are bit-equal to the reference evaluated on the same
                               * (bfloat16-rounded) inputs.  With EOSVR_SCREEN_BF16 and D % 8 == 0 the tensor-core
                               * pass reads the caller's rows in place: no second copy of the gallery in HBM. */

/* 16-bit format of the tensor-core screening copy */
#define EOSVR_SCREEN_F16   0  /* default: fp16, 11-bit significand -> tight error bound */
#define EOSVR_SCREEN_BF16  1  /* bf16; exact when the features are bf16-representable  */

/* metric / selection rule of the segment matching */
#define EOSVR_METRIC_EUCLID_TEMPORAL 0 /* cdist euclidean + [lam1,lam2,lam1] taps + arg-min (network_test.py:208-212) */
#define EOSVR_METRIC_COSINE          1 /* L2-normalise both sides, cosine similarity, arg-max: the reference's other
                                        * metric (classifier.py:117-120 idiom: argsort(-cosine_similarity)[:,0]), no
                                        * temporal taps (lam1, lam2, rows_per_episode are ignored).  The similarity
                                        * is evaluated in float64 on the ORIGINAL rows (dot / (|a| |b|), 0 for a zero
                                        * row) and rounded to float32; lowest index on exact ties.  d_out_score holds
                                        * the cosine; the packed word holds -cosine (so that the shard merge stays an
                                        * unsigned minimum; eosvr_merge_top1 then reports -cosine as its score). */

/* "original clip feature" row of the augmented support set */
#define EOSVR_ORIG_REF_QUIRK 0 /* flat segment row i, as written at network_test.py:229 */
#define EOSVR_ORIG_CLIP_MEAN 1 /* mean of clip i's segment rows, the commented intent :227-228 */

typedef struct eosvr_gallery   eosvr_gallery_t;    /* gallery feature cache (replaces the arrays of network_test.py:184-189) */
typedef struct eosvr_workspace eosvr_workspace_t;  /* per-stream scratch of the matcher */

int         eosvr_version(void);
const char *eosvr_last_error(void);
/* 0 if a usable sm_100 device is current, EOSVR_EUNSUPPORTED / EOSVR_ECUDA otherwise. */
int         eosvr_device_check(void);

/* ---- gallery feature cache -------------------------------------------------------
 * Wraps the gallery segment features gallery_seg_features[G,D] (network_test.py:187-189)
 * resident in HBM.  d_feats stays owned by the caller and must outlive the handle (the
 * exact re-rank reads it).  Builds, on `stream`, the 16-bit screening copy [G, Dpad]
 * (K-major, 128-byte swizzle friendly), the squared-norm side array and the error-bound
 * scalars.  global_offset is the index of row 0 in the un-sharded gallery (multi-GPU
 * sharding by segment); returned indices are global. */
int eosvr_gallery_create(const void *d_feats, int64_t G, int32_t D, int32_t dtype,
                         int64_t global_offset, int32_t screen_fmt, void *stream,
                         eosvr_gallery_t **out);
int eosvr_gallery_destroy(eosvr_gallery_t *g);
int eosvr_gallery_rows(const eosvr_gallery_t *g, int64_t *G, int32_t *D, int64_t *global_offset);
/* Storage type (EOSVR_F32 / EOSVR_BF16) of the handle's rows and whether the library holds its own 16-bit copy. */
int eosvr_gallery_info(const eosvr_gallery_t *g, int32_t *dtype, int32_t *owns_screen_copy);

/* bfloat16 -> float32 (exact) for probe / query batches that travelled as bfloat16: d_in [n] bf16, d_out [n] f32. */
int eosvr_upcast_bf16(const void *d_in, int64_t n, float *d_out, void *stream);

/* ---- workspace --------------------------------------------------------------------
 * Scratch for up to max_probe_rows probe segments of dimension D per call.
 * cand_capacity = number of (probe, gallery) near-minimum candidates the screening pass
 * may hand to the exact re-rank per call (0 = default: 128 per probe row; the value is divided
 * by max_probe_rows and clamped to 32..4096 per row).  Rows whose list fills up spill to a shared
 * buffer and, beyond that, are resolved exhaustively -- results never depend on the capacity. */
int eosvr_workspace_create(int64_t max_probe_rows, int32_t D, int64_t cand_capacity,
                           eosvr_workspace_t **out);
int eosvr_workspace_destroy(eosvr_workspace_t *ws);

/* ---- segment matching -------------------------------------------------------------
 * Replaces, for a batch of episodes, network_test.py:208-212:
 *     distance = cdist(support_seg_features, gallery_seg_features, 'euclidean')   # fp64
 *     distance = temporal_convolution_flating_layer(distance)                     # fp32 taps
 *     gallery_pool_ids = np.argsort(distance, axis=1)[:, :1]
 * d_probes [P,D] float32 = the support segment features of P/rows_per_episode episodes
 * (rows_per_episode = n_way*k_shot*num_segs consecutive rows are smoothed together with
 * zero padding at both ends, exactly as the reference pads one episode).
 * The [P,G] matrix is never written to memory.
 * Outputs (any may be NULL except d_out_packed):
 *   d_out_packed[P] uint64 : (order-preserving bits of the float32 smoothed distance) << 32
 *                            | global gallery index -- the shard-merge format;
 *   d_out_score[P]  float32: smoothed distance of the winner (bit-equal to the reference's);
 *   d_out_idx[P]    int64  : global index of the winning gallery segment, lowest index on
 *                            exact ties. */
/* Hardware assumption (checked, not trusted): the screening kernel lets three threads issue tcgen05.mma
 * into one TMEM accumulator without ordering them against each other.  PTX orders the MMAs of one thread
 * only; that several issuers add up exactly is an observed property of B200.  The first eosvr_match on a
 * device therefore runs a small synthetic match and compares every tensor-core screening value with a
 * CUDA-core evaluation (one host synchronisation, once per device and process); on a mismatch the library
 * switches to a single issuer, and if that fails too eosvr_match returns EOSVR_ECUDA.  Environment:
 * EOSVR_SELFCHECK=0 skips the check, EOSVR_ISSUERS=1 forces the single-issuer ordering. */
int eosvr_match(const eosvr_gallery_t *g, eosvr_workspace_t *ws, const float *d_probes,
                int64_t P, int32_t rows_per_episode, int32_t metric, float lam1, float lam2,
                uint64_t *d_out_packed, float *d_out_score, int64_t *d_out_idx, void *stream);

/* Same contract, evaluated entirely with the exact float64 CUDA-core kernel (no tensor
 * cores, no screening).  Used by the matcher for rows whose candidate list overflowed and
 * exposed for validation at small sizes. */
int eosvr_match_exact(const eosvr_gallery_t *g, eosvr_workspace_t *ws, const float *d_probes,
                      int64_t P, int32_t rows_per_episode, int32_t metric, float lam1, float lam2,
                      uint64_t *d_out_packed, float *d_out_score, int64_t *d_out_idx, void *stream);

/* Counters of the last eosvr_match on this workspace (synchronises `stream`):
 * out[0] candidates appended, out[1] candidates evaluated exactly, out[2] rows sent to the
 * exact fallback, out[3] candidate capacity, out[4] screening tiles, out[5] N of the MMA,
 * out[6] unsafe (cancellation-guard) candidates, out[7] candidates spilled from full row lists to
 * the shared buffer. */
int eosvr_match_stats(eosvr_workspace_t *ws, void *stream, int64_t out[8]);
/* The same counters, the first n of: out[0..7] as above, out[8] candidates the re-rank evaluated in float32 (one
 * gallery row read each: the re-rank's real traffic), out[9] exact evaluations whose float32 rounding the fast float64
 * sum could not decide, so that the reference's sequential summation was evaluated (about 1e-6 of them; later
 * counters read 0). */
int eosvr_match_stats_ex(eosvr_workspace_t *ws, void *stream, int64_t *out, int32_t n);

/* ---- multi-GPU shard merge (new) --------------------------------
 * d_gathered [nshards, P] = the d_out_packed arrays of all gallery shards (e.g. after one
 * ncclAllGather).  Element-wise minimum = global winner with the lowest-global-index tie
 * rule. */
int eosvr_merge_top1(const uint64_t *d_gathered, int32_t nshards, int64_t P,
                     uint64_t *d_out_packed, float *d_out_score, int64_t *d_out_idx, void *stream);

/* Rows of the winners: d_out_rows[p,:] = gallery[idx[p] - global_offset,:] when this shard
 * owns idx[p], else zeros (so a sum-reduce over shards delivers every winner row exactly). */
int eosvr_gather_rows(const eosvr_gallery_t *g, const int64_t *d_idx, int64_t P,
                      float *d_out_rows, void *stream);

/* ---- augmented-clip assembly -------------------------------------------------------
 * Replaces network_test.py:220-250 (video_segment_augmentation :119-129 + backbone
 * re-encode) in feature space:  for episode e, clip i: row 0 = "original" (orig_mode),
 * rows 1+s = float32 mean over the clip's S segment rows with row s replaced by the
 * winner row.  d_probes [E*n, S, D]; d_winner_rows [E*n*S, D]; d_out [E, n*(1+S), D].
 * Summation order = numpy's (sequential rows, one true division): bit-equal results. */
int eosvr_splice(const float *d_probes, const float *d_winner_rows, int64_t E, int32_t n,
                 int32_t S, int32_t D, int32_t orig_mode, float *d_out, void *stream);

/* ---- ProtoNet episode scoring ------------------------------------------------------
 * Replaces Classifier('protonet').predict, classifier.py:9-90, for E episodes:
 * d_support [E, R, D], d_support_y [E, R] float32 labels, d_query [E, Q, D].
 * Prototypes in first-appearance label order (float32 sequential mean), float64 distance
 * cast to float32, softmax(-d), first arg-max.  max_proto bounds the classes per episode.
 * Outputs (nullable): d_dist [E,Q,max_proto] float32 distances (logits = -d; unused slots
 * = +inf), d_prob [E,Q,max_proto], d_pred [E,Q] int64 prototype POSITION (classifier.py:85),
 * d_nproto [E] int32. */
int eosvr_proto_score(const float *d_support, const float *d_support_y, const float *d_query,
                      int64_t E, int32_t R, int32_t Q, int32_t D, int32_t max_proto,
                      float *d_dist, float *d_prob, int64_t *d_pred, int32_t *d_nproto,
                      void *stream);

/* ---- fused augmented-clip assembly + ProtoNet scoring -------------------------------
 * eosvr_splice followed by eosvr_proto_score for E episodes of n support clips x S segments,
 * without writing the augmented support set (network_test.py:220-259, classifier.py:9-90):
 * every clip contributes its "original" row and its S spliced rows, all labelled
 * d_support_y[e, i] (network_test.py:225,:232,:248).  Winner rows come from d_winner_rows
 * [E*n*S, D] when given (multi-GPU, after the shard exchange), else straight from gallery g
 * through the global indices d_idx [E*n*S] (rows g does not own read as zero).
 * Bit-equal to the two-call path.  Outputs as in eosvr_proto_score. */
int eosvr_episode_score(const float *d_probes, const float *d_winner_rows, const eosvr_gallery_t *g,
                        const int64_t *d_idx, const float *d_support_y, const float *d_query,
                        int64_t E, int32_t n, int32_t S, int32_t Q, int32_t D, int32_t orig_mode,
                        int32_t max_proto, float *d_dist, float *d_prob, int64_t *d_pred,
                        int32_t *d_nproto, void *stream);

/* ---- the whole path in ONE call (`eosvr_episode_batch`) ------------------
 * eosvr_match followed by eosvr_episode_score on the local gallery g for E episodes of n support clips x S
 * segments: the loop body of TestNetwork.test_network_aug_segment (network_test.py:195-259) from the cached
 * embeddings to the predictions.  d_probes [E*n*S, D], d_support_y [E, n], d_query [E, Q, D].
 * Outputs are caller-provided device buffers: d_out_packed [E*n*S] is required (it is also the shard-merge
 * payload), d_out_score / d_out_idx [E*n*S] and d_dist / d_prob [E,Q,max_proto], d_pred [E,Q], d_nproto [E] may
 * be NULL.  No allocation, no host synchronisation; 5 kernel launches (probe prep, seed pass, screening,
 * re-rank, finish) + 1 (fused splice + ProtoNet). */
int eosvr_episode_batch(const eosvr_gallery_t *g, eosvr_workspace_t *ws, const float *d_probes,
                        const float *d_support_y, const float *d_query, int64_t E, int32_t n, int32_t S,
                        int32_t Q, int32_t metric, float lam1, float lam2, int32_t orig_mode, int32_t max_proto,
                        uint64_t *d_out_packed, float *d_out_score, int64_t *d_out_idx, float *d_dist,
                        float *d_prob, int64_t *d_pred, int32_t *d_nproto, void *stream);

/* Same, with the gallery SHARDED by segment over the GPUs of one box (new): shard s holds the
 * global rows [d_shard_begin[s], d_shard_begin[s+1]) at d_shard_bases[s], a float32 / bfloat16 (shard_dtype) [rows, D] array in the memory
 * of the GPU that owns it, mapped into this process (peer / symmetric memory).  Winner rows are read IN PLACE over
 * NVLink by the scoring kernel -- no dense row exchange.  d_shard_bases [nshards] and d_shard_begin [nshards+1] are
 * device arrays.  Needs D % 4 == 0, D >= 256 and S in {2,4,8}. */
int eosvr_episode_score_sharded(const float *d_probes, const void *const *d_shard_bases, int32_t shard_dtype,
                                const int64_t *d_shard_begin, int32_t nshards, const int64_t *d_idx,
                                const float *d_support_y, const float *d_query, int64_t E, int32_t n, int32_t S,
                                int32_t Q, int32_t D, int32_t orig_mode, int32_t max_proto, float *d_dist,
                                float *d_prob, int64_t *d_pred, int32_t *d_nproto, void *stream);

/* ---- reference-shaped helpers ------------------------------------------------------------
 * eosvr_temporal_smooth: TestNetwork.temporal_convolution_flating_layer (network_test.py:103-117,
 * models.py:42-56) on an explicit float64 distance matrix d_dist64 [P,G] -> float32 [P,G]; blocks of
 * rows_per_episode rows are zero-padded separately (the reference passes one episode: rows_per_episode = P).
 * eosvr_cosine_predict: Classifier('cosine').predict (classifier.py:117-120): cosine similarity of each
 * query [E,Q,D] to each support row [E,R,D]; d_best[E,Q] = index of the best SUPPORT ROW (lowest index on
 * ties), optional d_sim [E,Q,R]. */
int eosvr_temporal_smooth(const double *d_dist64, int64_t P, int64_t G, int32_t rows_per_episode,
                          float lam1, float lam2, float *d_out, void *stream);
int eosvr_cosine_predict(const float *d_support, const float *d_query, int64_t E, int32_t R, int32_t Q,
                         int32_t D, float *d_sim, int64_t *d_best, void *stream);

/* ---- gallery / probe cache builder -------------------------------------------------
 * Replaces np.resize + np.mean of network_test.py:188-189 / :204-205 (+ per-frame
 * F.normalize of :79-80 when l2 != 0): d_frames [N*seg_len, D] -> d_out [N, D]. */
int eosvr_segment_features(const float *d_frames, int64_t N, int32_t seg_len, int32_t D,
                           int32_t l2, float *d_out, void *stream);

/* ---- callers either side of the path -----------------------------------
 * eosvr_clip_features: TestNetwork.generate_epoch_features (network_test.py:49-68) for a batch of clips:
 * d_frames [N, F, D] per-frame embeddings -> d_out [N, D] = mean over the first d_nframes[i] frames (all F
 * when d_nframes is NULL) of the frame features, per-frame L2-normalised first when l2 != 0 (:63).  The
 * baseline test truncates every support clip to its real frame count (:54-55, :145, :151).
 * eosvr_take_rows: d_out[i, :] = d_src[d_idx[i], :] for rows of row_elems floats -- index-only episode
 * assembly from a device-resident embedding cache (episode_novel_dataloader.py:19-80 without pixels). */
int eosvr_clip_features(const float *d_frames, int64_t N, int32_t F, int32_t D, const int32_t *d_nframes,
                        int32_t l2, float *d_out, void *stream);
int eosvr_take_rows(const float *d_src, int64_t n_src, int64_t row_elems, const int64_t *d_idx, int64_t n,
                    float *d_out, void *stream);

/* ---- measurement hooks -----------------------------------------------------------------
 * set_timing(on): bracket every kernel the following calls on this workspace launch with CUDA
 * events on the call's stream (per kernel class a ring of the last 128 launches, restarted by
 * this call).  kernel_ms: sum of the recorded durations of one kernel class and their count
 * (synchronises the events); screen_ms = kernel_ms(EOSVR_KERNEL_SCREEN).
 * launch_count: kernels this library has launched in the process (host-side count). */
#define EOSVR_KERNEL_PROBE_PREP 0  /* k_probe_prep: 16-bit probe plan, norms, error bounds, row-state reset */
#define EOSVR_KERNEL_SEED       1  /* k_match_screen over the strided seed sample                           */
#define EOSVR_KERNEL_SCREEN     2  /* k_match_screen, main pass (the dominant kernel)                       */
#define EOSVR_KERNEL_RERANK     3  /* k_rerank_warp / k_rerank_rows: exact re-rank of the candidates           */
#define EOSVR_KERNEL_FINISH     4  /* k_finish: unpack winners (+ exhaustive fallback)                      */
#define EOSVR_KERNEL_EPISODE    5  /* k_episode_partial: fused splice + ProtoNet (eosvr_episode_batch only) */
#define EOSVR_KERNEL_COUNT      6
int      eosvr_workspace_set_timing(eosvr_workspace_t *ws, int32_t on);
int      eosvr_workspace_kernel_ms(eosvr_workspace_t *ws, int32_t kernel, double *sum_ms, int64_t *calls);
int      eosvr_workspace_screen_ms(eosvr_workspace_t *ws, double *sum_ms, int64_t *calls);
uint64_t eosvr_launch_count(void);
/* Host-only: the probe tiling eosvr_match uses for (P, rows_per_episode):
 * out = {probe rows emitted per tile, halo columns per side, MMA N, number of probe tiles}. */
int      eosvr_plan(int64_t P, int32_t rows_per_episode, int64_t out[4]);

/* ---- test hook (not part of the reference-facing surface) ---------------------------
 * Dump the screening values t~[P,G] (float32, un-emitted entries untouched) of subsequent
 * eosvr_match calls into a caller-owned device buffer of `elems` floats; NULL disables. */
int eosvr_workspace_set_debug(eosvr_workspace_t *ws, float *d_dump, int64_t elems);

/* Cycle accounting of the last screening kernel when the environment variable EOSVR_EXP has bit 16 set
 * (measurement only, results unaffected; the timing modes that skip work exist only in -DEOSVR_EXPERIMENTS
 * builds; sums over CTAs): out = {epilogue busy, epilogue waiting for accumulators, MMA issuer
 * waiting for operands, MMA issuer waiting for a free accumulator, TMA producers waiting for a free stage,
 * kernel cycles summed over pair leaders, epilogue busy time before / inside the chunk loop of a tile}. */
int eosvr_workspace_debug_cycles(eosvr_workspace_t *ws, void *stream, int64_t out[8]);

#ifdef __cplusplus
}
#endif
#endif /* EOSVR_H_ */
