// eosvr_episode.cu -- augmented-clip assembly and ProtoNet scoring (bandwidth kernels).
//
//   k_splice          replaces network_test.py:220-250 (+ video_segment_augmentation :119-129 and
//                     the backbone re-encode :239-245) in feature space;
//   k_proto_score     replaces classifier.py:9-90 (prototypes, cdist, softmax, arg-max);
//   k_segment_features replaces network_test.py:188-189 / :204-205 (+ per-frame L2, :79-80).
//
// The float32 evaluation orders are numpy's (sequential row accumulation, one true division), so
// given equal winners the results are bit-equal to the reference's, not merely close
// ("Bit-level evaluation orders").
#include <math.h>

#include <map>
#include <mutex>
#include <utility>

#include "eosvr_internal.h"

namespace eosvr {

// Grow-only scratch per (device, stream): calls on one stream are ordered, calls on different streams get
// different buffers, so the library stays re-entrant without a workspace argument and without stream-ordered
// allocations on the hot path (cudaMallocAsync / cudaFreeAsync cost milliseconds per call when other streams
// are busy).  Freed at process exit.
static void *stream_scratch(cudaStream_t st, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, std::pair<void *, size_t>> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    auto &slot = cache[std::make_pair(dev, st)];
    if (slot.second < bytes) {
        if (slot.first) cudaFree(slot.first);            // synchronises; only when the scratch grows
        slot.first = nullptr; slot.second = 0;
        const size_t want = bytes + bytes / 2;
        if (cudaMalloc(&slot.first, want) != cudaSuccess) { cudaGetLastError(); slot.first = nullptr; return nullptr; }
        slot.second = want;
    }
    return slot.first;
}


// classifier.py:63 is scipy's cdist(query, prototypes) on float32 rows: promoted to float64 and, per pair, ONE
// sequential pass  s = 0; for k: e = q[k] - p[k]; s = s + e*e  (product rounded, then the sum; no FMA), sqrt(s),
// then the float32 cast of :66.  A float64 sum in another order flips the float32 rounding of about one distance
// in 10^6, and the sequential chain cannot be parallelised, so the distance is FILTERED: the kernels sum in any
// order (threads stride k, FMA, tree reduction); that sum S and scipy's are both within gamma_(D+1) of the exact
// sum of the same squared differences, |S_scipy - S| <= (2 D + 8) 2^-53 S; if float32(sqrt(.)) is the same at both
// ends of that interval (directed roundings; sqrt and the cast are monotone and correctly rounded) it IS scipy's
// float32.  Otherwise the pair is parked and a warp evaluates scipy's chain itself from the episode's prototypes in
// a scratch array (the lanes form 32 consecutive products in parallel, the running sum takes them in order).
__device__ __forceinline__ bool dist_f32_decided(double S, int D, float &out)
{
    const double delta = (2.0 * D + 8.0) * 1.1102230246251565e-16;        // (2 D + 8) 2^-53
    const float lo = static_cast<float>(sqrt(__dmul_rd(S, 1.0 - delta)));
    const float hi = static_cast<float>(sqrt(__dmul_ru(S, 1.0 + delta)));
    out = lo;
    return lo == hi;
}
// proto is read with ld.global.cg (with a split feature axis other blocks wrote it); all lanes return the value
__device__ __forceinline__ float warp_seq_dist(const float *__restrict__ q, const float *proto, int D, int lane)
{
    double s = 0.0;
    for (int k0 = 0; k0 < D; k0 += 32) {
        const int k = k0 + lane;
        double pr = 0.0;                                   // a lane beyond D adds +0.0: the sum is unchanged
        if (k < D) {
            const double e = __dsub_rn(static_cast<double>(q[k]), static_cast<double>(__ldcg(proto + k)));
            pr = __dmul_rn(e, e);
        }
#pragma unroll 4
        for (int i = 0; i < 32; ++i) s = __dadd_rn(s, __shfl_sync(0xffffffffu, pr, i));
    }
    return static_cast<float>(sqrt(s));
}
// Parked pairs of one episode (block-wide): decide-or-park, then resolve.  s_amb holds q * 256 + c.
constexpr int kMaxAmb = 32;
__device__ __forceinline__ void dist_decide(double S, int D, int q, int c, float (*s_d)[64], int *s_namb, int *s_amb)
{
    float d;
    if (dist_f32_decided(S, D, d)) { s_d[q][c] = d; return; }
    const int i = atomicAdd(s_namb, 1);
    if (i < kMaxAmb) s_amb[i] = q * 256 + c;
    s_d[q][c] = __int_as_float(0x7fc00000);                // resolved below (or by the caller if the list overflowed)
}
// after a __syncthreads(): the block's warps resolve the parked pairs; a list overflow (never seen: it takes > 32
// boundary hits in one episode) resolves every pair of the episode.  Ends with a __syncthreads().
__device__ __forceinline__ void dist_resolve(const float *__restrict__ Qp, const float *PR, int D, int Q, int np,
                                             float (*s_d)[64], const int *s_namb, const int *s_amb, int nthreads)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = nthreads >> 5;
    const int n = *s_namb;
    if (n > kMaxAmb) {
        for (int t = warp; t < Q * np; t += nw) {
            const int q = t / np, c = t - q * np;
            const float d = warp_seq_dist(Qp + static_cast<int64_t>(q) * D, PR + static_cast<int64_t>(c) * D, D, lane);
            if (lane == 0) s_d[q][c] = d;
        }
    } else {
        for (int i = warp; i < n; i += nw) {
            const int q = s_amb[i] >> 8, c = s_amb[i] & 255;
            const float d = warp_seq_dist(Qp + static_cast<int64_t>(q) * D, PR + static_cast<int64_t>(c) * D, D, lane);
            if (lane == 0) s_d[q][c] = d;
        }
    }
    __syncthreads();
}

// grid: (E*n, 1+S); block: 128 threads striding D.
// out[e, i*(1+S) + j, :]:  j = 0 "original" row, j = 1+s the clip with segment s replaced.
__global__ void k_splice(const float *__restrict__ probes, const float *__restrict__ wrows,
                         int n, int S, int D, int orig_mode, float *__restrict__ out)
{
    const int64_t clip = blockIdx.x;              // e*n + i
    const int j = blockIdx.y;
    const int64_t e = clip / n;
    const int i = static_cast<int>(clip % n);
    const float *pc = probes + clip * S * D;      // the clip's S segment rows
    float *o = out + (clip * (1 + S) + j) * D;
    if (j == 0 && orig_mode == EOSVR_ORIG_REF_QUIRK) {
        // network_test.py:229: support_seg_features[i] = flat segment row i of the episode
        const float *src = probes + (e * n * S + i) * D;
        for (int k = threadIdx.x; k < D; k += blockDim.x) o[k] = src[k];
        return;
    }
    const int s_rep = j - 1;                       // -1: no replacement (clip mean)
    const float *wr = s_rep >= 0 ? wrows + (clip * S + s_rep) * D : nullptr;
    const float fS = static_cast<float>(S);
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
        float acc = (s_rep == 0) ? wr[k] : pc[k];
        for (int s = 1; s < S; ++s) {
            const float v = (s == s_rep) ? wr[k] : pc[static_cast<int64_t>(s) * D + k];
            acc = __fadd_rn(acc, v);
        }
        o[k] = __fdiv_rn(acc, fS);
    }
}

int launch_splice(const float *probes, const float *wrows, int64_t E, int32_t n, int32_t S, int32_t D,
                  int32_t orig_mode, float *out, cudaStream_t st)
{
    if (E == 0) return EOSVR_OK;
    dim3 grid(static_cast<unsigned>(E * n), static_cast<unsigned>(1 + S));
    k_splice<<<grid, 128, 0, st>>>(probes, wrows, n, S, D, orig_mode, out);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

// One block per episode.
constexpr int kProtoThreads = 256;
constexpr int kMaxRows = 1024;     // support rows per episode
constexpr int kMaxProto = 64;
constexpr int kMaxQ = 8;

__global__ void __launch_bounds__(kProtoThreads)
k_proto_score(const float *__restrict__ sup, const float *__restrict__ sup_y, const float *__restrict__ query,
              int R, int Q, int D, int max_proto, float *protos, int pstride, float *__restrict__ dist,
              float *__restrict__ prob, int64_t *__restrict__ pred, int32_t *__restrict__ nproto_out)
{
    __shared__ int16_t s_cls[kMaxRows];
    __shared__ float s_pid[kMaxProto];
    __shared__ int s_np, s_namb, s_amb[kMaxAmb];
    __shared__ double s_red[kProtoThreads / 32][kMaxQ];
    __shared__ float s_d[kMaxQ][kMaxProto];

    const int64_t e = blockIdx.x;
    const float *S0 = sup + e * R * D;
    const float *Y = sup_y + e * R;
    const float *Qp = query + e * Q * D;
    float *PR = protos + e * pstride * D;              // this episode's prototypes [np, D] (read by the rare sequential pass)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        s_namb = 0;
        // classifier.py:21-29: classes keyed by label value, first-appearance order
        int np = 0;
        for (int r = 0; r < R; ++r) {
            const float y = Y[r];
            int c = -1;
            for (int j = 0; j < np; ++j) if (s_pid[j] == y) { c = j; break; }
            if (c < 0) { if (np < max_proto) { c = np; s_pid[np++] = y; } else c = -1; }
            s_cls[r] = static_cast<int16_t>(c);
        }
        s_np = np;
    }
    __syncthreads();
    const int np = s_np;

    for (int c = 0; c < np; ++c) {
        double part[kMaxQ];
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) part[q] = 0.0;
        for (int k = tid; k < D; k += kProtoThreads) {
            // classifier.py:34-35: float32 mean over the class rows, sequential, true division
            float acc = 0.f; int cnt = 0;
            for (int r = 0; r < R; ++r) {
                if (s_cls[r] != c) continue;
                const float v = S0[static_cast<int64_t>(r) * D + k];
                acc = cnt ? __fadd_rn(acc, v) : v;
                ++cnt;
            }
            const float pk = __fdiv_rn(acc, static_cast<float>(cnt));
            PR[static_cast<int64_t>(c) * D + k] = pk;
            // classifier.py:63: cdist in double, direct differences (summation order: see dist_f32_decided)
#pragma unroll
            for (int q = 0; q < kMaxQ; ++q) {
                if (q < Q) {
                    const double df = static_cast<double>(Qp[static_cast<int64_t>(q) * D + k]) - static_cast<double>(pk);
                    part[q] += df * df;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
            double v = part[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[warp][q] = v;
        }
        __syncthreads();
        if (tid < Q) {
            double v = 0.0;
            for (int w = 0; w < kProtoThreads / 32; ++w) v += s_red[w][tid];
            dist_decide(v, D, tid, c, s_d, &s_namb, s_amb);  // classifier.py:66 float32 cast
        }
        __syncthreads();
    }
    dist_resolve(Qp, PR, D, Q, np, s_d, &s_namb, s_amb, kProtoThreads);

    if (tid < Q) {
        const int q = tid;
        // classifier.py:67 softmax(-d); :85 first arg-max == first arg-min of the float32 distance
        float mx = -s_d[q][0]; int best = 0;
        for (int c = 1; c < np; ++c) { if (-s_d[q][c] > mx) mx = -s_d[q][c]; if (s_d[q][c] < s_d[q][best]) best = c; }
        float sum = 0.f;
        for (int c = 0; c < np; ++c) sum += expf(-s_d[q][c] - mx);
        for (int c = 0; c < max_proto; ++c) {
            const int64_t o = (e * Q + q) * max_proto + c;
            if (dist) dist[o] = c < np ? s_d[q][c] : INFINITY;
            if (prob) prob[o] = c < np ? expf(-s_d[q][c] - mx) / sum : 0.f;
        }
        if (pred) pred[e * Q + q] = best;
    }
    if (tid == 0 && nproto_out) nproto_out[e] = np;
}

int launch_proto_score(const float *sup, const float *sup_y, const float *query, int64_t E, int32_t R,
                       int32_t Q, int32_t D, int32_t max_proto, float *dist, float *prob, int64_t *pred,
                       int32_t *nproto, cudaStream_t st)
{
    if (E == 0) return EOSVR_OK;
    if (R < 1 || R > kMaxRows || Q < 1 || Q > kMaxQ || max_proto < 1 || max_proto > kMaxProto) {
        set_error("proto_score: need 1<=R<=%d, 1<=Q<=%d, 1<=max_proto<=%d (got R=%d Q=%d max_proto=%d)",
                  kMaxRows, kMaxQ, kMaxProto, R, Q, max_proto);
        return EOSVR_EINVAL;
    }
    const int pstride = R < max_proto ? R : max_proto;      // prototypes per episode, at most
    float *protos = static_cast<float *>(stream_scratch(st, static_cast<size_t>(E) * pstride * D * sizeof(float)));
    if (!protos) { set_error("proto_score: scratch allocation failed"); return EOSVR_ENOMEM; }
    k_proto_score<<<static_cast<unsigned>(E), kProtoThreads, 0, st>>>(sup, sup_y, query, R, Q, D, max_proto, protos, pstride,
                                                                     dist, prob, pred, nproto);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

// -------------------------------------------------------------------------------------------
// Fused augmented-clip assembly + ProtoNet scoring (network_test.py:220-259 + classifier.py:9-90) for a
// batch of episodes: the augmented support features [E, n(1+S), D] are formed in registers and folded
// straight into the class prototypes; nothing but the distances / predictions is written.
// Bit-equal to k_splice followed by k_proto_score (same float32 evaluation order).
// One block per episode; threads stride D.  Winner rows come either from d_wrows [E*n*S, D] (multi-GPU:
// after the shard exchange) or directly from the local gallery through idx.
// -------------------------------------------------------------------------------------------
constexpr int kMaxClips = 256;

template <int S_T>
__global__ void __launch_bounds__(kProtoThreads)
k_episode_score(const float *__restrict__ probes, const float *__restrict__ wrows, const void *__restrict__ gal,
                int gdt, int64_t G, int64_t goff, const int64_t *__restrict__ idx, const float *__restrict__ sup_y,
                const float *__restrict__ query, int n, int S_rt, int Q, int D, int orig_mode, int max_proto,
                float *protos, int pstride, float *__restrict__ dist, float *__restrict__ prob, int64_t *__restrict__ pred,
                int32_t *__restrict__ nproto_out)
{
    __shared__ int16_t s_cls[kMaxClips];
    __shared__ int16_t s_order[kMaxClips];      // clips grouped by class, original order inside a class
    __shared__ int16_t s_start[kMaxProto + 1];
    __shared__ float s_pid[kMaxProto];
    __shared__ int s_np, s_namb, s_amb[kMaxAmb];
    __shared__ double s_red[kProtoThreads / 32][kMaxQ];
    __shared__ float s_d[kMaxQ][kMaxProto];

    const int S = S_T > 0 ? S_T : S_rt;
    const int64_t e = blockIdx.x;
    const float *Pe = probes + e * n * S * D;             // the episode's n*S segment rows
    const float *Y = sup_y + e * n;
    const float *Qp = query + e * Q * D;
    float *PR = protos + e * pstride * D;                 // this episode's prototypes [np, D]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        s_namb = 0;
        int np = 0;
        for (int i = 0; i < n; ++i) {                     // classifier.py:21-29 (every row of clip i carries Y[i])
            const float y = Y[i];
            int c = -1;
            for (int j = 0; j < np; ++j) if (s_pid[j] == y) { c = j; break; }
            if (c < 0) { if (np < max_proto) { c = np; s_pid[np++] = y; } else c = -1; }
            s_cls[i] = static_cast<int16_t>(c);
        }
        int pos = 0;
        for (int c = 0; c < np; ++c) {
            s_start[c] = static_cast<int16_t>(pos);
            for (int i = 0; i < n; ++i) if (s_cls[i] == c) s_order[pos++] = static_cast<int16_t>(i);
        }
        s_start[np] = static_cast<int16_t>(pos);
        s_np = np;
    }
    __syncthreads();
    const int np = s_np;
    const float fS = static_cast<float>(S);

    for (int c = 0; c < np; ++c) {
        double part[kMaxQ];
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) part[q] = 0.0;
        const int i0 = s_start[c], i1 = s_start[c + 1];
        for (int k = tid; k < D; k += kProtoThreads) {
            float acc = 0.f;
            int cnt = 0;
            for (int ii = i0; ii < i1; ++ii) {
                const int i = s_order[ii];
                const float *pc = Pe + static_cast<int64_t>(i) * S * D + k;
                const int64_t wbase = (e * n + i) * S;
                if (S_T > 0) {
                    float pr[S_T > 0 ? S_T : 1], w[S_T > 0 ? S_T : 1];
#pragma unroll
                    for (int s = 0; s < S_T; ++s) pr[s] = pc[static_cast<int64_t>(s) * D];
#pragma unroll
                    for (int s = 0; s < S_T; ++s) {
                        if (wrows) w[s] = wrows[(wbase + s) * D + k];
                        else {
                            const int64_t g = idx[wbase + s] - goff;
                            w[s] = (g >= 0 && g < G) ? ld_feat(gal, gdt, g * D + k) : 0.f;
                        }
                    }
                    float o;
                    if (orig_mode == EOSVR_ORIG_REF_QUIRK) o = Pe[static_cast<int64_t>(i) * D + k];   // network_test.py:229
                    else {
                        o = pr[0];
#pragma unroll
                        for (int s = 1; s < S_T; ++s) o = __fadd_rn(o, pr[s]);
                        o = __fdiv_rn(o, fS);
                    }
                    acc = cnt ? __fadd_rn(acc, o) : o; ++cnt;
#pragma unroll
                    for (int s = 0; s < S_T; ++s) {
                        float a = (s == 0) ? w[0] : pr[0];
#pragma unroll
                        for (int s2 = 1; s2 < S_T; ++s2) a = __fadd_rn(a, (s2 == s) ? w[s2] : pr[s2]);
                        acc = __fadd_rn(acc, __fdiv_rn(a, fS)); ++cnt;
                    }
                } else {
                    float o;
                    if (orig_mode == EOSVR_ORIG_REF_QUIRK) o = Pe[static_cast<int64_t>(i) * D + k];
                    else {
                        o = pc[0];
                        for (int s = 1; s < S; ++s) o = __fadd_rn(o, pc[static_cast<int64_t>(s) * D]);
                        o = __fdiv_rn(o, fS);
                    }
                    acc = cnt ? __fadd_rn(acc, o) : o; ++cnt;
                    for (int s = 0; s < S; ++s) {
                        float wv;
                        if (wrows) wv = wrows[(wbase + s) * D + k];
                        else {
                            const int64_t g = idx[wbase + s] - goff;
                            wv = (g >= 0 && g < G) ? ld_feat(gal, gdt, g * D + k) : 0.f;
                        }
                        float a = (s == 0) ? wv : pc[0];
                        for (int s2 = 1; s2 < S; ++s2) a = __fadd_rn(a, (s2 == s) ? wv : pc[static_cast<int64_t>(s2) * D]);
                        acc = __fadd_rn(acc, __fdiv_rn(a, fS)); ++cnt;
                    }
                }
            }
            const float pk = __fdiv_rn(acc, static_cast<float>(cnt));
            PR[static_cast<int64_t>(c) * D + k] = pk;
#pragma unroll
            for (int q = 0; q < kMaxQ; ++q) {
                if (q < Q) {
                    const double df = static_cast<double>(Qp[static_cast<int64_t>(q) * D + k]) - static_cast<double>(pk);
                    part[q] += df * df;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
            double v = part[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[warp][q] = v;
        }
        __syncthreads();
        if (tid < Q) {
            double v = 0.0;
            for (int w = 0; w < kProtoThreads / 32; ++w) v += s_red[w][tid];
            dist_decide(v, D, tid, c, s_d, &s_namb, s_amb);
        }
        __syncthreads();
    }
    dist_resolve(Qp, PR, D, Q, np, s_d, &s_namb, s_amb, kProtoThreads);

    if (tid < Q) {
        const int q = tid;
        float mx = -s_d[q][0]; int best = 0;
        for (int c = 1; c < np; ++c) { if (-s_d[q][c] > mx) mx = -s_d[q][c]; if (s_d[q][c] < s_d[q][best]) best = c; }
        float sum = 0.f;
        for (int c = 0; c < np; ++c) sum += expf(-s_d[q][c] - mx);
        for (int c = 0; c < max_proto; ++c) {
            const int64_t o = (e * Q + q) * max_proto + c;
            if (dist) dist[o] = c < np ? s_d[q][c] : INFINITY;
            if (prob) prob[o] = c < np ? expf(-s_d[q][c] - mx) / sum : 0.f;
        }
        if (pred) pred[e * Q + q] = best;
    }
    if (tid == 0 && nproto_out) nproto_out[e] = np;
}

// -------------------------------------------------------------------------------------------
// Split variant of the fused kernel for D % 4 == 0 and S in {2,4,8}: grid (E, nsplit), every block
// folds one slice of the feature axis with float4 loads and writes its float64 partial sums of squared
// query-prototype differences (and its slice of the float32 prototypes, for the rare sequential pass); the last
// block of the episode adds the slices in a fixed order (deterministic), decides the float32 distances
// (dist_f32_decided) and finishes softmax / arg-max.  Same float32 evaluation order per feature as
// k_episode_score, so everything is bit-equal.  One block per episode left most of the 148 SMs idle at E = 256.
// -------------------------------------------------------------------------------------------
constexpr int kEpThreads = 128;

template <int S_T>
__global__ void __launch_bounds__(kEpThreads)
k_episode_partial(const float *__restrict__ probes, const float *__restrict__ wrows, const void *__restrict__ gal,
                  int gdt, int64_t G, int64_t goff, const void *const *__restrict__ shard_bases,
                  const int64_t *__restrict__ shard_begin, int nshards, const int64_t *__restrict__ idx,
                  const float *__restrict__ sup_y, const float *__restrict__ query, int n, int Q, int D, int orig_mode,
                  int max_proto, int nsplit, double *partial, float *protos, int pstride, unsigned int *done,
                  float *__restrict__ dist, float *__restrict__ prob, int64_t *__restrict__ pred, int32_t *__restrict__ nproto_out)
{
    // where each winner row of the episode lives: the exchanged rows, the local gallery, or -- gallery sharded
    // over the GPUs of the box -- the owning GPU's memory, read in place over NVLink (peer loads)
    // (winner rows given explicitly are float32; gallery / shard rows are float32 or bfloat16, gdt)
    __shared__ const void *s_wbase[kMaxClips * S_T];
    __shared__ int64_t s_wrow4[kMaxClips * S_T];      // index of the row's first float4 group in its array
    __shared__ int16_t s_cls[kMaxClips];
    __shared__ int16_t s_order[kMaxClips];
    __shared__ int16_t s_start[kMaxProto + 1];
    __shared__ float s_pid[kMaxProto];
    __shared__ int s_np, s_namb, s_amb[kMaxAmb];
    __shared__ double s_red[kEpThreads / 32][kMaxQ];
    __shared__ float s_d[kMaxQ][kMaxProto];
    __shared__ int s_last;

    const int64_t e = blockIdx.x;
    const int sp = blockIdx.y;
    const int D4 = D >> 2;
    const float4 *Pe = reinterpret_cast<const float4 *>(probes + e * n * S_T * D);
    const float *Y = sup_y + e * n;
    const float4 *Qp = reinterpret_cast<const float4 *>(query + e * Q * D);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int t = tid; t < n * S_T; t += kEpThreads) {
        const int64_t slot = e * n * S_T + t;
        const void *base = nullptr;
        int64_t row4 = 0;
        if (wrows) { base = wrows; row4 = slot * D4; }
        else {
            const int64_t g = idx[slot];
            if (nshards > 0) {
                for (int sh = 0; sh < nshards; ++sh)
                    if (g >= shard_begin[sh] && g < shard_begin[sh + 1]) { base = shard_bases[sh]; row4 = (g - shard_begin[sh]) * D4; }
            } else if (g - goff >= 0 && g - goff < G) { base = gal; row4 = (g - goff) * D4; }
        }
        s_wbase[t] = base;
        s_wrow4[t] = row4;
    }
    const int wdt = wrows ? EOSVR_F32 : gdt;
    float *PR = protos + e * pstride * D;                 // this episode's prototypes [np, D] (read by the rare sequential pass)
    if (tid == 0) {
        s_namb = 0;
        int np = 0;
        for (int i = 0; i < n; ++i) {                     // classifier.py:21-29 (every row of clip i carries Y[i])
            const float y = Y[i];
            int c = -1;
            for (int j = 0; j < np; ++j) if (s_pid[j] == y) { c = j; break; }
            if (c < 0) { if (np < max_proto) { c = np; s_pid[np++] = y; } else c = -1; }
            s_cls[i] = static_cast<int16_t>(c);
        }
        int pos = 0;
        for (int c = 0; c < np; ++c) {
            s_start[c] = static_cast<int16_t>(pos);
            for (int i = 0; i < n; ++i) if (s_cls[i] == c) s_order[pos++] = static_cast<int16_t>(i);
        }
        s_start[np] = static_cast<int16_t>(pos);
        s_np = np;
    }
    __syncthreads();
    const int np = s_np;
    const float fS = static_cast<float>(S_T);
    const int per = (D4 + nsplit - 1) / nsplit;
    const int k0 = sp * per, k1 = min(D4, k0 + per);

    for (int c = 0; c < np; ++c) {
        double part[kMaxQ];
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) part[q] = 0.0;
        const int i0 = s_start[c], i1 = s_start[c + 1];
        for (int k = k0 + tid; k < k1; k += kEpThreads) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            int cnt = 0;
            for (int ii = i0; ii < i1; ++ii) {
                const int i = s_order[ii];
                const float4 *pc = Pe + static_cast<int64_t>(i) * S_T * D4 + k;
                float pr[S_T][4], w[S_T][4];
#pragma unroll
                for (int s = 0; s < S_T; ++s) {
                    const float4 v = pc[static_cast<int64_t>(s) * D4];
                    pr[s][0] = v.x; pr[s][1] = v.y; pr[s][2] = v.z; pr[s][3] = v.w;
                }
#pragma unroll
                for (int s = 0; s < S_T; ++s) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    const void *wb = s_wbase[i * S_T + s];
                    if (wb) v = ld_feat4(wb, wdt, s_wrow4[i * S_T + s] + k);
                    w[s][0] = v.x; w[s][1] = v.y; w[s][2] = v.z; w[s][3] = v.w;
                }
                float o[4];
                if (orig_mode == EOSVR_ORIG_REF_QUIRK) {                       // network_test.py:229
                    const float4 v = Pe[static_cast<int64_t>(i) * D4 + k];
                    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
                } else {
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        float a = pr[0][x];
#pragma unroll
                        for (int s = 1; s < S_T; ++s) a = __fadd_rn(a, pr[s][x]);
                        o[x] = __fdiv_rn(a, fS);
                    }
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) acc[x] = cnt ? __fadd_rn(acc[x], o[x]) : o[x];
                ++cnt;
#pragma unroll
                for (int s = 0; s < S_T; ++s) {
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        float a = (s == 0) ? w[0][x] : pr[0][x];
#pragma unroll
                        for (int s2 = 1; s2 < S_T; ++s2) a = __fadd_rn(a, (s2 == s) ? w[s2][x] : pr[s2][x]);
                        acc[x] = __fadd_rn(acc[x], __fdiv_rn(a, fS));
                    }
                    ++cnt;
                }
            }
            const float fc = static_cast<float>(cnt);
            const float pk[4] = {__fdiv_rn(acc[0], fc), __fdiv_rn(acc[1], fc), __fdiv_rn(acc[2], fc), __fdiv_rn(acc[3], fc)};
            reinterpret_cast<float4 *>(PR + static_cast<int64_t>(c) * D)[k] = make_float4(pk[0], pk[1], pk[2], pk[3]);
#pragma unroll
            for (int q = 0; q < kMaxQ; ++q) {
                if (q < Q) {
                    const float4 qv = Qp[static_cast<int64_t>(q) * D4 + k];
                    const float qq[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const double df = static_cast<double>(qq[x]) - static_cast<double>(pk[x]);
                        part[q] += df * df;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
            double v = part[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[warp][q] = v;
        }
        __syncthreads();
        if (tid < Q) {
            double v = 0.0;
            for (int w = 0; w < kEpThreads / 32; ++w) v += s_red[w][tid];
            partial[((e * nsplit + sp) * max_proto + c) * kMaxQ + tid] = v;
        }
        __syncthreads();
    }
    // The last block of the episode to get here adds the slices in a FIXED order (deterministic), decides the float32
    // distances (dist_f32_decided; the rare undecided pair is evaluated in scipy's order from the prototype scratch) and
    // finishes softmax / arg-max (classifier.py:63-67,:85); its ticket orders it after the other blocks' partial sums
    // and prototype slices.
    if (nsplit > 1) {
        __threadfence();                                  // every thread: its prototype stores before ...
        __syncthreads();                                  // ... thread 0 takes the block's ticket
        if (tid == 0) {
            const unsigned int ticket = atomicAdd(done + e, 1u);
            s_last = ticket == static_cast<unsigned int>(nsplit - 1);
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
    }
    if (tid == 0 && nproto_out) nproto_out[e] = np;
    for (int t = tid; t < Q * np; t += kEpThreads) {
        const int q = t / np, c = t - q * np;
        double v = 0.0;
        for (int s2 = 0; s2 < nsplit; ++s2) v += __ldcg(partial + ((e * nsplit + s2) * max_proto + c) * kMaxQ + q);
        dist_decide(v, D, q, c, s_d, &s_namb, s_amb);         // classifier.py:66 float32 cast
    }
    __syncthreads();
    dist_resolve(query + e * Q * D, PR, D, Q, np, s_d, &s_namb, s_amb, kEpThreads);
    if (tid >= Q) return;
    const int q = tid;
    float mx = -s_d[q][0]; int best = 0;
    for (int c = 1; c < np; ++c) { if (-s_d[q][c] > mx) mx = -s_d[q][c]; if (s_d[q][c] < s_d[q][best]) best = c; }
    float sum = 0.f;
    for (int c = 0; c < np; ++c) sum += expf(-s_d[q][c] - mx);
    for (int c = 0; c < max_proto; ++c) {
        const int64_t o = (e * Q + q) * max_proto + c;
        if (dist) dist[o] = c < np ? s_d[q][c] : INFINITY;
        if (prob) prob[o] = c < np ? expf(-s_d[q][c] - mx) / sum : 0.f;
    }
    if (pred) pred[e * Q + q] = best;
}

int launch_episode_score(const float *probes, const float *wrows, const void *gal, int32_t gal_dtype, int64_t G, int64_t goff,
                         const void *const *shard_bases, const int64_t *shard_begin, int32_t nshards,
                         const int64_t *idx, const float *sup_y, const float *query, int64_t E, int32_t n,
                         int32_t S, int32_t Q, int32_t D, int32_t orig_mode, int32_t max_proto, float *dist,
                         float *prob, int64_t *pred, int32_t *nproto, cudaStream_t st, eosvr_workspace *timing_ws)
{
    if (E == 0) return EOSVR_OK;
    if (nshards > 0 && !((D & 3) == 0 && (S == 2 || S == 4 || S == 8) && D >= 256)) {
        set_error("episode_score_sharded: needs D %% 4 == 0, D >= 256 and S in {2,4,8}");
        return EOSVR_ESHAPE;
    }
    if (n < 1 || n > kMaxClips || S < 1 || Q < 1 || Q > kMaxQ || max_proto < 1 || max_proto > kMaxProto) {
        set_error("episode_score: need 1<=n<=%d, S>=1, 1<=Q<=%d, 1<=max_proto<=%d", kMaxClips, kMaxQ, kMaxProto);
        return EOSVR_EINVAL;
    }
    const int pstride = n < max_proto ? n : max_proto;      // prototypes per episode, at most
    if ((D & 3) == 0 && (S == 2 || S == 4 || S == 8) && D >= 256) {
        int nsplit = D / 512;                 // >= 128 float4 columns per block
        if (nsplit < 1) nsplit = 1;
        if (nsplit > 8) nsplit = 8;
        // scratch: [ticket counters E x u32] [partial sums] [float32 prototypes E x pstride x D]
        const size_t pbytes = (static_cast<size_t>(E) * nsplit * max_proto * kMaxQ * sizeof(double) + 255) / 256 * 256;
        const size_t nbytes = (static_cast<size_t>(E) * sizeof(unsigned int) + 255) / 256 * 256;
        const size_t rbytes = static_cast<size_t>(E) * pstride * D * sizeof(float);
        char *scratch = static_cast<char *>(stream_scratch(st, nbytes + pbytes + rbytes));
        if (!scratch) { set_error("episode_score: scratch allocation of %zu bytes failed", nbytes + pbytes + rbytes); return EOSVR_ENOMEM; }
        float *protos = reinterpret_cast<float *>(scratch + nbytes + pbytes);
        unsigned int *done = reinterpret_cast<unsigned int *>(scratch);
        if (nsplit > 1) EOSVR_CUDA(cudaMemsetAsync(done, 0, static_cast<size_t>(E) * sizeof(unsigned int), st));   // ticket counters
        double *partial = reinterpret_cast<double *>(scratch + nbytes);
        dim3 pgrid(static_cast<unsigned>(E), static_cast<unsigned>(nsplit));
#define EOSVR_EPP_LAUNCH(ST)                                                                                      \
        k_episode_partial<ST><<<pgrid, kEpThreads, 0, st>>>(probes, wrows, gal, gal_dtype, G, goff, shard_bases, shard_begin,     \
                                                           nshards, idx, sup_y, query, n, Q, D, orig_mode, max_proto, \
                                                           nsplit, partial, protos, pstride, done, dist, prob, pred, nproto)
        { int trc = timing_begin(timing_ws, EOSVR_KERNEL_EPISODE, st); if (trc) return trc; }
        if (S == 2) EOSVR_EPP_LAUNCH(2); else if (S == 4) EOSVR_EPP_LAUNCH(4); else EOSVR_EPP_LAUNCH(8);
#undef EOSVR_EPP_LAUNCH
        EOSVR_CUDA(cudaGetLastError());
        { int trc = timing_end(timing_ws, EOSVR_KERNEL_EPISODE, st); if (trc) return trc; }
        EOSVR_COUNT_LAUNCH(1);
        return EOSVR_OK;
    }
    const unsigned grid = static_cast<unsigned>(E);
    float *protos = static_cast<float *>(stream_scratch(st, static_cast<size_t>(E) * pstride * D * sizeof(float)));
    if (!protos) { set_error("episode_score: scratch allocation failed"); return EOSVR_ENOMEM; }
#define EOSVR_EP_LAUNCH(ST)                                                                                    \
    k_episode_score<ST><<<grid, kProtoThreads, 0, st>>>(probes, wrows, gal, gal_dtype, G, goff, idx, sup_y, query, n, S, Q, D, \
                                                        orig_mode, max_proto, protos, pstride, dist, prob, pred, nproto)
    switch (S) {
        case 2: EOSVR_EP_LAUNCH(2); break;
        case 4: EOSVR_EP_LAUNCH(4); break;
        case 8: EOSVR_EP_LAUNCH(8); break;
        case 16: EOSVR_EP_LAUNCH(16); break;
        default: EOSVR_EP_LAUNCH(0); break;
    }
#undef EOSVR_EP_LAUNCH
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

// -------------------------------------------------------------------------------------------
// Materialised temporal smoothing: TestNetwork.temporal_convolution_flating_layer
// (network_test.py:103-117, TemporalLayer models.py:42-56) on an explicit [P,G] float64 distance matrix
// (the reference-shaped call; the matcher never materialises this matrix).  Same float32 FMA chain as the
// exact re-rank: float32 cast, acc = lam1*d[p-1]; acc = fma(lam2, d[p], acc); acc = fma(lam1, d[p+1], acc),
// zero padding at the ends of each block of rows_per_episode rows.
// -------------------------------------------------------------------------------------------
__global__ void k_temporal_smooth(const double *__restrict__ d64, int64_t P, int64_t G, int rpe, float lam1,
                                  float lam2, float *__restrict__ out)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= P * G) return;
    const int64_t p = i / G;
    const int r = static_cast<int>(p % rpe);
    const bool hl = r > 0, hr = (r + 1 < rpe) && (p + 1 < P);
    const float dl = hl ? static_cast<float>(d64[i - G]) : 0.f;
    const float dc = static_cast<float>(d64[i]);
    const float dr = hr ? static_cast<float>(d64[i + G]) : 0.f;
    float acc = __fmul_rn(lam1, dl);
    acc = __fmaf_rn(lam2, dc, acc);
    acc = __fmaf_rn(lam1, dr, acc);
    out[i] = acc;
}

int launch_temporal_smooth(const double *d64, int64_t P, int64_t G, int32_t rpe, float lam1, float lam2,
                           float *out, cudaStream_t st)
{
    if (P * G == 0) return EOSVR_OK;
    const int threads = 256;
    k_temporal_smooth<<<static_cast<unsigned>((P * G + threads - 1) / threads), threads, 0, st>>>(d64, P, G, rpe,
                                                                                               lam1, lam2, out);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

// -------------------------------------------------------------------------------------------
// Cosine scoring: Classifier('cosine').predict, classifier.py:117-120 -- cosine similarity of every query
// to every SUPPORT ROW and the index of the best row (not its label); lowest index on
// ties.  One block per (episode, query); float64 accumulation, float32 result.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_cosine_predict(const float *__restrict__ sup, const float *__restrict__ query, int R, int Q, int D,
                 float *__restrict__ sim, int64_t *__restrict__ best)
{
    __shared__ float s_best[8];
    __shared__ int s_idx[8];
    const int64_t eq = blockIdx.x;                 // e*Q + q
    const int64_t e = eq / Q;
    const float *qv = query + eq * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double qq = 0.0;
    for (int k = lane; k < D; k += 32) qq += static_cast<double>(qv[k]) * static_cast<double>(qv[k]);
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    float bestv = -INFINITY; int besti = 0x7fffffff;
    for (int r = warp; r < R; r += 8) {
        const float *sv = sup + (e * R + r) * D;
        double dot = 0.0, ss = 0.0;
        for (int k = lane; k < D; k += 32) {
            const double a = sv[k], b = qv[k];
            dot += a * b; ss += a * a;
        }
        for (int o = 16; o > 0; o >>= 1) { dot += __shfl_xor_sync(0xffffffffu, dot, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
        const double den = sqrt(qq) * sqrt(ss);
        const float c = den > 0.0 ? static_cast<float>(dot / den) : 0.f;
        if (sim && lane == 0) sim[eq * R + r] = c;
        if (c > bestv) { bestv = c; besti = r; }   // rows ascend within a warp: first maximum kept
    }
    if (lane == 0) { s_best[warp] = bestv; s_idx[warp] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float bv = s_best[0]; int bi = s_idx[0];
        for (int w = 1; w < 8; ++w)
            if (s_best[w] > bv || (s_best[w] == bv && s_idx[w] < bi)) { bv = s_best[w]; bi = s_idx[w]; }
        best[eq] = bi;
    }
}

int launch_cosine_predict(const float *sup, const float *query, int64_t E, int32_t R, int32_t Q, int32_t D,
                          float *sim, int64_t *best, cudaStream_t st)
{
    if (E * Q == 0) return EOSVR_OK;
    k_cosine_predict<<<static_cast<unsigned>(E * Q), 256, 0, st>>>(sup, query, R, Q, D, sim, best);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

// One warp per output segment row.
__global__ void k_segment_features(const float *__restrict__ frames, int64_t N, int seg_len, int D, int l2,
                                   float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    float inv[16];
    const int sl = seg_len < 16 ? seg_len : 16;
    for (int f = 0; f < sl; ++f) {
        float nrm = 1.f;
        if (l2) {
            const float *x = frames + (row * seg_len + f) * D;
            double s = 0.0;
            for (int k = lane; k < D; k += 32) s += static_cast<double>(x[k]) * static_cast<double>(x[k]);
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            nrm = fmaxf(static_cast<float>(sqrt(s)), 1e-12f);      // F.normalize eps
        }
        inv[f] = nrm;
    }
    for (int k = lane; k < D; k += 32) {
        float acc = 0.f;
        for (int f = 0; f < sl; ++f) {
            float v = frames[(row * seg_len + f) * D + k];
            if (l2) v = __fdiv_rn(v, inv[f]);
            acc = f ? __fadd_rn(acc, v) : v;
        }
        out[row * D + k] = __fdiv_rn(acc, static_cast<float>(seg_len));
    }
}

int launch_segment_features(const float *frames, int64_t N, int32_t seg_len, int32_t D, int32_t l2,
                            float *out, cudaStream_t st)
{
    if (N == 0) return EOSVR_OK;
    if (seg_len < 1 || seg_len > 16) { set_error("segment_features: 1 <= seg_len <= 16"); return EOSVR_EINVAL; }
    const int threads = 256;
    k_segment_features<<<static_cast<unsigned>((N * 32 + threads - 1) / threads), threads, 0, st>>>(
        frames, N, seg_len, D, l2, out);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

// -------------------------------------------------------------------------------------------
// Clip features with per-clip frame counts: TestNetwork.generate_epoch_features (network_test.py:49-68) for a
// batch of clips -- the mean over the first nframes[i] frames of the (optionally per-frame L2-normalised, :63)
// frame features; the baseline test truncates every support clip to its real frame count (:54-55, :145).
// Sequential float32 accumulation and one true division (numpy's np.mean order).  One warp per clip.
// -------------------------------------------------------------------------------------------
__global__ void k_clip_features(const float *__restrict__ frames, int64_t N, int F, int D, const int32_t *__restrict__ nframes,
                                int l2, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t clip = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (clip >= N) return;
    int nf = nframes ? nframes[clip] : F;
    nf = nf < 1 ? 1 : (nf > F ? F : nf);
    const float *base = frames + clip * F * D;
    for (int k0 = 0; k0 < D; k0 += 32 * 4) {                  // 4 feature columns per lane and sweep
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int f = 0; f < nf; ++f) {
            const float *x = base + static_cast<int64_t>(f) * D;
            float nrm = 1.f;
            if (l2) {
                double s = 0.0;
                for (int k = lane; k < D; k += 32) s += static_cast<double>(x[k]) * static_cast<double>(x[k]);
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                nrm = fmaxf(static_cast<float>(sqrt(s)), 1e-12f);      // F.normalize eps
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + j * 32 + lane;
                if (k < D) {
                    float v = x[k];
                    if (l2) v = __fdiv_rn(v, nrm);
                    acc[j] = f ? __fadd_rn(acc[j], v) : v;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j * 32 + lane;
            if (k < D) out[clip * D + k] = __fdiv_rn(acc[j], static_cast<float>(nf));
        }
    }
}

int launch_clip_features(const float *frames, int64_t N, int32_t F, int32_t D, const int32_t *nframes, int32_t l2,
                         float *out, cudaStream_t st)
{
    if (N == 0) return EOSVR_OK;
    const int threads = 256;
    k_clip_features<<<static_cast<unsigned>((N * 32 + threads - 1) / threads), threads, 0, st>>>(frames, N, F, D, nframes, l2, out);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

// -------------------------------------------------------------------------------------------
// Index-only episode assembly from a device-resident embedding cache (episode_novel_dataloader.py:19-80 without
// touching pixels or the host): out[i, :] = src[idx[i], :], rows of row_elems floats.  One block per output row.
// -------------------------------------------------------------------------------------------
__global__ void k_take_rows(const float *__restrict__ src, int64_t n_src, int64_t row_elems, const int64_t *__restrict__ idx,
                            float *__restrict__ out)
{
    const int64_t i = blockIdx.x;
    int64_t r = idx[i];
    r = r < 0 ? 0 : (r >= n_src ? n_src - 1 : r);             // (validated on the host; never read out of bounds)
    const float *s = src + r * row_elems;
    float *d = out + i * row_elems;
    if ((row_elems & 3) == 0) {
        const float4 *s4 = reinterpret_cast<const float4 *>(s);
        float4 *d4 = reinterpret_cast<float4 *>(d);
        for (int64_t k = threadIdx.x; k < row_elems / 4; k += blockDim.x) d4[k] = s4[k];
    } else {
        for (int64_t k = threadIdx.x; k < row_elems; k += blockDim.x) d[k] = s[k];
    }
}

int launch_take_rows(const float *src, int64_t n_src, int64_t row_elems, const int64_t *idx, int64_t n, float *out, cudaStream_t st)
{
    if (n == 0) return EOSVR_OK;
    k_take_rows<<<static_cast<unsigned>(n), 256, 0, st>>>(src, n_src, row_elems, idx, out);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

}  // namespace eosvr
