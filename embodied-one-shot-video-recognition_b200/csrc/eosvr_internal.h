// eosvr_internal.h -- handle layouts and launch prototypes shared by the translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/eosvr.h"

namespace eosvr {

// ---- tiling constants of the screening kernel ----------------------------------------
constexpr int kBM = 128;          // gallery rows per CTA and tile (one TMEM lane per row)
constexpr int kPairM = 256;       // gallery rows per CTA pair and tile (UMMA M with cta_group::2)
constexpr int kBK = 64;           // K elements per pipeline stage (128-byte rows, SWIZZLE_128B)
constexpr int kMaxBN = 256;       // probe columns per tile (UMMA N), multiple of 16
constexpr int kSub = 1;           // K blocks per pipeline stage (4*kSub MMAs per tcgen05.commit)
constexpr int kStages = 6;        // TMA -> MMA smem ring (kSub x (16 KiB A + 16 KiB B-half) per stage and CTA)
constexpr int kIssuers = 3;       // MMA issuer warps 1..3 of the pair's leader; stage s belongs to issuer s % kIssuers.
                                  // A tcgen05.commit blocks its thread for ~600 cycles and a UTCHMMA for ~70
                                  // (tools/bench_micro/mma_rate.cu): one thread sustains only ~1/3 of the MMA rate.
constexpr int kMaxAccStages = 3;  // TMEM accumulators: 2 x 256 columns, or 3 x 160 when the probe tile has <= 160 columns
constexpr bool kZeroAcc = true;   // every MMA accumulates; the epilogue zeroes the accumulator behind itself.  false: the
                                  // tile's first stage overwrites and the other issuers wait for its completion (a
                                  // ~600-cycle bubble per tile)
constexpr int kTmemCols = 512;
constexpr int kBN3 = 160;         // widest probe tile that leaves room for three accumulator stages (3 x 160 <= 512 columns)
constexpr int kAlign = 20;        // chunk width of the episode-aligned epilogue: 5-way x 4 segments (the metric's episode)
// Epilogue warps (template parameter EW of k_match_screen): warps 4 .. 4+EW-1; warp % 4 = TMEM lane quadrant, (warp - 4) / 4 =
// column group.  8 warps keep up with the tensor pipe when D >= 1024; for shorter rows the epilogue is the critical path and 16
// warps run it (tools/bench_micro/epi_rate.cu: 3538 -> 2793 cycles per 128 x 240 tile against 3840 cycles of MMA at D = 512).
constexpr int kMaxEpiWarps = 16;
// TMA producers: warp 0 loads the gallery operand; with 8 epilogue warps a second producer warp (4+EW) loads the probe
// operand (a UTMALDG occupies its issuing thread for ~150 cycles; one thread for both measured 3-5 % slower at D = 2048,
// where a stage lasts 448 cycles, and no different at D = 512).  With 16 epilogue warps warp 0 loads both: 20 warps = 5 per
// scheduler can start at 96 registers, which is what lets the epilogue warps grow to 104 (setmaxnreg moves registers
// inside the CTA's own allocation only).
__host__ __device__ constexpr bool two_producers(int ew) { return ew <= 8; }
__host__ __device__ constexpr int screen_threads(int ew) { return 128 + 32 * ew + (two_producers(ew) ? 32 : 0); }
static_assert(kStages % kIssuers == 0, "every stage barrier must have a single consumer warp");
constexpr int kChunk = 16;        // TMEM columns per tcgen05.ld
constexpr int kMaxSeedTiles = 4;   // strided gallery tiles screened first to seed the per-probe thresholds
constexpr float kPadNorm = 1.0e30f;
constexpr int kTimingRing = 128;  // event pairs kept per kernel class for eosvr_workspace_kernel_ms (the last 128 launches)

void count_launches(unsigned n);        // kernels launched by this library (host-side count, atomic)
#define EOSVR_COUNT_LAUNCH(n) (::eosvr::count_launches(n))

void set_error(const char *fmt, ...);

#ifdef __CUDACC__
// Gallery feature rows are float32 or bfloat16 (eosvr_gallery_create dtype); bfloat16 -> float32 is exact (a shift).
// i4 / i index groups of 4 elements / single elements from the start of the array.
__device__ __forceinline__ float4 ld_feat4(const void *base, int dtype, int64_t i4)
{
    if (dtype == EOSVR_F32) return reinterpret_cast<const float4 *>(base)[i4];
    const uint2 u = reinterpret_cast<const uint2 *>(base)[i4];
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                       __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}
__device__ __forceinline__ float ld_feat(const void *base, int dtype, int64_t i)
{
    if (dtype == EOSVR_F32) return reinterpret_cast<const float *>(base)[i];
    return __uint_as_float(static_cast<uint32_t>(reinterpret_cast<const unsigned short *>(base)[i]) << 16);
}
#endif

#define EOSVR_CUDA(call)                                                                     \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            ::eosvr::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),      \
                               __FILE__, __LINE__);                                          \
            return EOSVR_ECUDA;                                                              \
        }                                                                                    \
    } while (0)

// near-minimum candidate handed from the tensor-core screening pass to the exact re-rank;
// candidates are kept in per-probe-row lists of fixed capacity
struct Cand {
    int32_t g;        // gallery row, local to the shard
    uint32_t tbits;   // screening value (float bits); kCandUnsafe = inside the cancellation guard
};
constexpr uint32_t kCandUnsafe = 0xFFFFFFFFu;
// spill-over entry for candidates that did not fit their row's list (shared buffer)
struct OvfCand {
    int32_t p, g;
    uint32_t tbits, pad;
};

struct Counters {
    unsigned long long cand_count;   // appended (may exceed capacity)
    unsigned long long n_exact;      // evaluated exactly by the re-rank
    unsigned long long n_unsafe;
    unsigned int overflow;           // appends dropped
    unsigned int n_flag_rows;        // rows sent to the exact fallback
    unsigned int xfloor_bits;        // max over columns of the cancellation guard (x domain)
    unsigned int ovf_count;          // entries appended to the shared spill-over buffer
    unsigned int done_blocks;        // k_finish: blocks that have finished the exhaustive evaluation (last one unpacks)
    unsigned int pad0;
    // cycle accounting of the screening kernel (EOSVR_EXP bit 16; measurement only), summed over CTAs
    unsigned long long cyc_epi_busy, cyc_epi_wait, cyc_mma_wait_full, cyc_mma_wait_acc, cyc_prod_wait, cyc_total;
    unsigned long long cyc_epi_pre, cyc_epi_loop;   // epilogue busy time split: before / inside the chunk loop of a tile
    unsigned long long n_eval32;     // candidates the re-rank evaluated in float32 (one gallery row read each)
    unsigned long long n_seq;        // exact evaluations the fast float64 sum could not decide (sequential chain taken)
};

}  // namespace eosvr

// 16-bit screening copy of the gallery + side arrays + tensor maps (one per metric)
struct eosvr_screen_copy {
    void *h16;               // [Gpad, Dp] fp16/bf16
    float *gnorm;            // [Gpad] squared norm of the screened row (kPadNorm beyond G)
    float *scalars;          // [4] error-bound scalars
    CUtensorMap tmapA, tmapSeed;
    cudaEvent_t ready;       // recorded after the build kernel (lazily built copies)
};

struct eosvr_gallery {
    const void *feats;       // [G,D] float32 or bfloat16 (dtype), caller-owned
    int32_t dtype;           // EOSVR_F32 / EOSVR_BF16 storage of feats
    int32_t owns_h16;        // 0: the screening copy IS the caller's bfloat16 rows (no second copy in HBM)
    int64_t G;
    int32_t seed_tiles;      // gallery tiles of the strided seed pass
    int64_t seed_stride;     // row stride of the seed pass
    CUtensorMap tmapSeed;
    int32_t D, Dp;           // Dp = D rounded up to kBK
    int64_t offset;          // global index of row 0
    int32_t screen_fmt;
    void *h16;               // [Gpad, Dp] fp16/bf16 screening copy (Gpad = G rounded up to kBM)
    float *gnorm;            // [Gpad] ||b||^2 (kPadNorm beyond G)
    float *scalars;          // [4]: max ||b||^2, max ||b_lo||^2, max ||b_hi||^2 (float bits, >= 0)
    CUtensorMap tmapA;
    int device;
    eosvr_screen_copy *cos;  // L2-normalised rows for the cosine metric; built by the first cosine match
    void *cos_mutex;         // std::mutex guarding the lazy build
};

struct eosvr_workspace {
    int64_t maxP;
    int32_t D, Dp;
    int64_t cap_rows;        // plan rows capacity
    int64_t cand_cap;        // candidates per probe row
    void *slab;              // single device allocation
    // carved views
    void *q16;               // [cap_rows, Dp] packed probe plan (16-bit)
    float *na, *wl, *wr, *epsd;   // [cap_rows]
    int32_t *rowmap;         // [cap_rows] emitted probe row or -1
    unsigned int *gthr;      // [maxP] float bits of the running threshold
    unsigned long long *best;  // [maxP] packed winners
    int32_t *rowflag;        // [maxP]
    unsigned int *rowcnt;    // [maxP] candidates appended per probe row
    eosvr::Cand *cand;       // [maxP, cand_cap]
    eosvr::OvfCand *ovf;     // [ovf_cap] shared spill-over of full row lists
    int64_t ovf_cap;
    eosvr::Counters *counters;
    float *dbg;              // optional [P,G] dump of screening values (tests)
    int64_t dbg_elems;
    // last-call info for eosvr_match_stats
    int64_t last_tiles;
    int32_t last_bn;
    int device;
    // optional CUDA-event timing of every kernel class of the path (bench.py rooflines); EOSVR_KERNEL_* ids
    int timing_on;
    int64_t timing_calls[EOSVR_KERNEL_COUNT];
    cudaEvent_t ev0[EOSVR_KERNEL_COUNT][eosvr::kTimingRing], ev1[EOSVR_KERNEL_COUNT][eosvr::kTimingRing];
    int64_t *idx_scratch;    // [maxP] winner indices for eosvr_episode_batch when the caller does not want them
};

namespace eosvr {

struct MatchPlan {
    int64_t P;
    int32_t rpe;       // rows per episode
    int32_t R;         // emitted probe rows per tile
    int32_t halo;      // 0/1 halo column on each side
    int32_t BN;        // UMMA N (multiple of 16)
    int64_t NT;        // probe tiles
};

MatchPlan make_plan(int64_t P, int32_t rpe);

// Per-device launch state (the library may drive several GPUs from one process); guarded by one mutex.
struct DeviceState {
    int num_sms = 0;
    int max_clusters[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};   // [plain / DIAG / aligned / aligned DIAG][EW == 16]: co-resident CTA pairs of k_match_screen
    bool rr_attr = false;
    int issuers = 0;                              // MMA issuer warps in use; 0 = not decided yet (self-check pending)
};
// First screening call on a device: run a small match through the tensor-core path and compare every screening
// value with a CUDA-core evaluation (eosvr_selfcheck.cu).  Decides ds->issuers (kIssuers, or 1 if several threads
// issuing tcgen05.mma into one accumulator do not add up on this part) or fails.
int screening_selfcheck(DeviceState *ds, cudaStream_t st);

int launch_gallery_prep(eosvr_gallery *g, cudaStream_t st);
int build_cosine_copy(eosvr_gallery *g, cudaStream_t st);      // eosvr_api.cu
int launch_gallery_prep_cos(const eosvr_gallery *g, eosvr_screen_copy *c, cudaStream_t st);
int launch_match(eosvr_gallery *g, eosvr_workspace *ws, const float *probes, int64_t P,
                 int32_t rpe, int32_t metric, float lam1, float lam2, bool exact_only, uint64_t *out_packed,
                 float *out_score, int64_t *out_idx, cudaStream_t st);
int launch_merge(const uint64_t *gathered, int32_t nshards, int64_t P, uint64_t *out_packed,
                 float *out_score, int64_t *out_idx, cudaStream_t st);
int launch_gather_rows(const eosvr_gallery *g, const int64_t *idx, int64_t P, float *out, cudaStream_t st);
int launch_upcast_bf16(const void *in, int64_t n, float *out, cudaStream_t st);
int launch_clip_features(const float *frames, int64_t N, int32_t F, int32_t D, const int32_t *nframes, int32_t l2,
                         float *out, cudaStream_t st);
int launch_take_rows(const float *src, int64_t n_src, int64_t row_elems, const int64_t *idx, int64_t n, float *out, cudaStream_t st);
int launch_splice(const float *probes, const float *wrows, int64_t E, int32_t n, int32_t S, int32_t D,
                  int32_t orig_mode, float *out, cudaStream_t st);
int launch_episode_score(const float *probes, const float *wrows, const void *gal, int32_t gal_dtype, int64_t G, int64_t goff,
                         const void *const *shard_bases, const int64_t *shard_begin, int32_t nshards,
                         const int64_t *idx, const float *sup_y, const float *query, int64_t E, int32_t n,
                         int32_t S, int32_t Q, int32_t D, int32_t orig_mode, int32_t max_proto, float *dist,
                         float *prob, int64_t *pred, int32_t *nproto, cudaStream_t st, eosvr_workspace *timing_ws = nullptr);
// CUDA-event brackets around a kernel class when ws->timing_on (no-ops otherwise); eosvr_api.cu
int timing_begin(eosvr_workspace *ws, int kernel, cudaStream_t st);
int timing_end(eosvr_workspace *ws, int kernel, cudaStream_t st);
int launch_proto_score(const float *sup, const float *sup_y, const float *query, int64_t E, int32_t R,
                       int32_t Q, int32_t D, int32_t max_proto, float *dist, float *prob, int64_t *pred,
                       int32_t *nproto, cudaStream_t st);
int launch_temporal_smooth(const double *d64, int64_t P, int64_t G, int32_t rpe, float lam1, float lam2,
                           float *out, cudaStream_t st);
int launch_cosine_predict(const float *sup, const float *query, int64_t E, int32_t R, int32_t Q, int32_t D,
                          float *sim, int64_t *best, cudaStream_t st);
int launch_segment_features(const float *frames, int64_t N, int32_t seg_len, int32_t D, int32_t l2,
                            float *out, cudaStream_t st);
int encode_tmap_2d(CUtensorMap *m, const void *base, int fmt, uint64_t rows, uint64_t cols,
                   uint32_t box_rows, uint32_t box_cols, uint64_t row_stride = 1);

}  // namespace eosvr
