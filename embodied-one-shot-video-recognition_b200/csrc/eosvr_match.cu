// eosvr_match.cu -- segment matching (replaces network_test.py:208-212 of the reference):
//   cdist(probe, gallery, 'euclidean') [fp64] -> float32 -> [lam1,lam2,lam1] taps along the
//   probe axis (zero padded per episode) -> arg-min per probe row, lowest index on ties.
//
// Two cooperating parts, neither of which ever writes the [P,G] matrix:
//   1. SCREENING (k_match_screen): a persistent, warp-specialised tcgen05 kernel.  Gallery
//      rows ride the UMMA M axis (one TMEM lane = one gallery row), probe segments ride N, so
//      every epilogue thread owns one gallery row and sees consecutive probe segments in
//      consecutive registers: the temporal 3-tap is register-local.  The epilogue forms
//      d = sqrt(|a|^2 + |b|^2 - 2 a.b), the taps, and compares against a per-probe running
//      threshold (min so far + rigorous error margin); the few elements below it are
//      appended to a candidate list.  13 warps per CTA: 2 TMA producers, 3 MMA issuers (one
//      issuing thread cannot keep the tensor pipe fed), 8 epilogue warps.
//   2. EXACT RE-RANK (k_rerank_rows): candidates are first evaluated with float32 direct
//      differences (an error bound orders of magnitude tighter than the 16-bit screening) and
//      the float32 near-ties -- normally one per row -- exactly as the reference does (float64
//      direct differences, float32 cast, float32 FMA chain), merged with a packed 64-bit
//      atomicMin whose low word is the gallery index: the lowest-index tie rule for free.
// The error margins guarantee the true winner (and every exact tie) is evaluated exactly, so
// indices and scores are bit-equal to the reference, not merely close.  The cosine metric
// (EOSVR_METRIC_COSINE) runs the same kernels on L2-normalised screening copies.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "eosvr_internal.h"
#include "eosvr_ptx.cuh"

namespace eosvr {

using namespace ptx;

constexpr float kBig = 1.0e30f;
constexpr float kSlopMul = 1.0f + 1.0f / 65536.0f;   // covers fp32 rounding of the screening taps / sqrt.approx

__device__ __forceinline__ unsigned long long pack_score_idx(float t, uint32_t gidx)
{
    uint32_t b = __float_as_uint(t);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);          // order-preserving float -> uint
    return (static_cast<unsigned long long>(b) << 32) | gidx;
}
__device__ __forceinline__ float unpack_score(unsigned long long v)
{
    uint32_t b = static_cast<uint32_t>(v >> 32);
    b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
    return __uint_as_float(b);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// -------------------------------------------------------------------------------------------
// Gallery cache build (off the timed path): 16-bit screening copy, squared norms, bound scalars.
// One warp per gallery row.
// -------------------------------------------------------------------------------------------
template <typename T16>
__device__ __forceinline__ T16 to16(float x);
template <>
__device__ __forceinline__ __half to16<__half>(float x) { return __float2half_rn(x); }
template <>
__device__ __forceinline__ __nv_bfloat16 to16<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
__device__ __forceinline__ float from16(__half x) { return __half2float(x); }
__device__ __forceinline__ float from16(__nv_bfloat16 x) { return __bfloat162float(x); }

__device__ __forceinline__ void atomic_max_posf(float *addr, float v)
{
    atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// NORM: screen the L2-normalised row (cosine metric).  The bound scalars are taken against the exactly
// normalised row (float64), the stored copy is the 16-bit rounding of the float32 product x * (1/|x|); a zero
// row stays zero and gets squared norm 1 so that |a' - 0|^2 = 2 - 2*0, consistent with cosine 0.
template <typename T16, bool NORM>
__global__ void k_gallery_prep(const float *__restrict__ feats, int64_t G, int64_t Gpad, int D, int Dp,
                               T16 *__restrict__ h16, float *__restrict__ gnorm, float *__restrict__ scalars)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= Gpad) return;
    T16 *dst = h16 + row * Dp;
    if (row >= G) {
        for (int k = lane; k < Dp; k += 32) dst[k] = to16<T16>(0.f);
        if (lane == 0) gnorm[row] = kPadNorm;
        return;
    }
    const float *src = feats + row * D;
    double rn = 1.0; float rnf = 1.f;
    if (NORM) {
        double q = 0.0;
        for (int k = lane; k < D; k += 32) q += static_cast<double>(src[k]) * static_cast<double>(src[k]);
        q = warp_sum(q);
        rn = q > 0.0 ? 1.0 / sqrt(q) : 0.0;
        rnf = static_cast<float>(rn);
    }
    double s = 0.0, sh = 0.0, sl = 0.0;
    for (int k = lane; k < Dp; k += 32) {
        float x = k < D ? src[k] : 0.f;
        T16 h = to16<T16>(NORM ? __fmul_rn(x, rnf) : x);
        dst[k] = h;
        double xd = NORM ? static_cast<double>(x) * rn : static_cast<double>(x), hd = from16(h);
        s += xd * xd;
        sh += hd * hd;
        sl += (xd - hd) * (xd - hd);
    }
    s = warp_sum(s); sh = warp_sum(sh); sl = warp_sum(sl);
    if (lane == 0) {
        if (NORM && rn == 0.0) s = 1.0;
        gnorm[row] = static_cast<float>(s);
        atomic_max_posf(scalars + 0, __double2float_ru(s));
        atomic_max_posf(scalars + 1, __double2float_ru(sl));
        atomic_max_posf(scalars + 2, __double2float_ru(sh));
    }
}

template <bool NORM>
static int launch_gallery_prep_t(const eosvr_gallery *g, void *h16, float *gnorm, float *scalars, cudaStream_t st)
{
    const int64_t Gpad = (g->G + kPairM - 1) / kPairM * kPairM;
    EOSVR_CUDA(cudaMemsetAsync(scalars, 0, 4 * sizeof(float), st));
    const int threads = 256;
    const int64_t blocks = (Gpad * 32 + threads - 1) / threads;
    if (g->screen_fmt == EOSVR_SCREEN_F16)
        k_gallery_prep<__half, NORM><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
            g->feats, g->G, Gpad, g->D, g->Dp, static_cast<__half *>(h16), gnorm, scalars);
    else
        k_gallery_prep<__nv_bfloat16, NORM><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
            g->feats, g->G, Gpad, g->D, g->Dp, static_cast<__nv_bfloat16 *>(h16), gnorm, scalars);
    EOSVR_CUDA(cudaGetLastError());
    return EOSVR_OK;
}

int launch_gallery_prep(eosvr_gallery *g, cudaStream_t st)
{
    return launch_gallery_prep_t<false>(g, g->h16, g->gnorm, g->scalars, st);
}

int launch_gallery_prep_cos(const eosvr_gallery *g, eosvr_screen_copy *c, cudaStream_t st)
{
    return launch_gallery_prep_t<true>(g, c->h16, c->gnorm, c->scalars, st);
}

// -------------------------------------------------------------------------------------------
// Probe plan.  Probe rows are laid out as NT tiles of BN columns; tile t emits probe rows
// [t*R, (t+1)*R).  When an episode fits a tile (rpe <= 256) R is a whole number of episodes and
// no halo is needed; otherwise each tile carries one halo column on each side so the taps of its
// first/last emitted row see their neighbours.
// -------------------------------------------------------------------------------------------
MatchPlan make_plan(int64_t P, int32_t rpe)
{
    MatchPlan pl;
    pl.P = P;
    pl.rpe = rpe;
    if (rpe <= kMaxBN) {
        int64_t k = kMaxBN / rpe;
        pl.R = static_cast<int32_t>(k * rpe);
        pl.halo = 0;
    } else {
        pl.R = kMaxBN - 2;
        pl.halo = 1;
    }
    if (pl.R > P) pl.R = static_cast<int32_t>(P);
    pl.BN = (pl.R + 2 * pl.halo + 15) / 16 * 16;
    pl.NT = (P + pl.R - 1) / pl.R;
    return pl;
}

struct PlanDev {
    int64_t P;
    int32_t rpe, R, halo, BN;
    int64_t NT;
};

__device__ __forceinline__ int64_t plan_row_of(const PlanDev &pl, int64_t col, bool &valid, bool &emit)
{
    const int64_t t = col / pl.BN;
    const int c = static_cast<int>(col % pl.BN);
    const int jj = c - pl.halo;
    const int64_t p = t * pl.R + jj;
    valid = (c < pl.R + 2 * pl.halo) && p >= 0 && p < pl.P;
    emit = valid && jj >= 0 && jj < pl.R;
    return p;
}

// One warp per plan column: convert the probe row to the 16-bit screening format, compute the
// squared norm and the per-column error bound E2 (see DESIGN.md "Error bound").
template <typename T16, bool NORM>
__global__ void k_probe_prep(const float *__restrict__ probes, PlanDev pl, int D, int Dp,
                             const float *__restrict__ gscal, T16 *__restrict__ q16,
                             float *__restrict__ na, float *__restrict__ epsd, int32_t *__restrict__ rowmap,
                             Counters *ctr)
{
    const int lane = threadIdx.x & 31;
    const int64_t col = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (col >= pl.NT * pl.BN) return;
    bool valid, emit;
    const int64_t p = plan_row_of(pl, col, valid, emit);
    T16 *dst = q16 + col * Dp;
    if (!valid) {
        for (int k = lane; k < Dp; k += 32) dst[k] = to16<T16>(0.f);
        if (lane == 0) { na[col] = kPadNorm; epsd[col] = 0.f; rowmap[col] = -1; }
        return;
    }
    const float *src = probes + p * D;
    double rn = 1.0; float rnf = 1.f;
    if (NORM) {                                          // cosine metric: screen the L2-normalised row
        double q = 0.0;
        for (int k = lane; k < D; k += 32) q += static_cast<double>(src[k]) * static_cast<double>(src[k]);
        q = warp_sum(q);
        rn = q > 0.0 ? 1.0 / sqrt(q) : 0.0;
        rnf = static_cast<float>(rn);
    }
    double s = 0.0, sh = 0.0, sl = 0.0;
    for (int k = lane; k < Dp; k += 32) {
        float x = k < D ? src[k] : 0.f;
        T16 h = to16<T16>(NORM ? __fmul_rn(x, rnf) : x);
        dst[k] = h;
        double xd = NORM ? static_cast<double>(x) * rn : static_cast<double>(x), hd = from16(h);
        s += xd * xd;
        sh += hd * hd;
        sl += (xd - hd) * (xd - hd);
    }
    s = warp_sum(s); sh = warp_sum(sh); sl = warp_sum(sl);
    if (NORM && rn == 0.0) s = 1.0;                      // zero row: |0 - b'|^2 + ... = 2 - 2*0
    if (lane == 0) {
        const double B2 = gscal[0], Bl2 = gscal[1], Bh2 = gscal[2];
        const double ulp = 1.0 / 4194304.0;   // 2^-22
        // |x~ - x| <= 2(|a_lo||b| + |a_hi||b_lo|) + tensor-core accumulation + fp32 rounding
        double e2 = 2.0 * (sqrt(sl * B2) + sqrt(sh * Bl2))
                  + 2.0 * ulp * (Dp / 16 + 1) * sqrt(sh * Bh2)
                  + 2.0 * ulp * (s + B2);
        e2 *= 1.01;
        if (e2 < 1e-30) e2 = 1e-30;
        na[col] = static_cast<float>(s);
        epsd[col] = __double2float_ru(sqrt(e2) / 16.0);           // E2 / (2 sqrt(64 E2))
        rowmap[col] = emit ? static_cast<int32_t>(p) : -1;
        atomicMax(&ctr->xfloor_bits, __float_as_uint(__double2float_ru(65.0 * e2)));
    }
}

// Thread per plan column: tap weights and threshold margin.
__global__ void k_column_plan(PlanDev pl, float w, const float *__restrict__ epsd,
                              float *__restrict__ wl, float *__restrict__ wr, float *__restrict__ margin)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t ncol = pl.NT * pl.BN;
    if (i >= ncol) return;
    bool valid, emit;
    const int64_t p = plan_row_of(pl, i, valid, emit);
    const int c = static_cast<int>(i % pl.BN);
    float l = 0.f, r = 0.f, el = 0.f, er = 0.f;
    if (valid) {
        if (c > 0 && p > 0 && (p % pl.rpe) != 0) { l = w; el = epsd[i - 1]; }
        if (c + 1 < pl.BN && p + 1 < pl.P && ((p + 1) % pl.rpe) != 0) {
            bool v2, e2;
            plan_row_of(pl, i + 1, v2, e2);
            if (v2) { r = w; er = epsd[i + 1]; }
        }
    }
    wl[i] = l; wr[i] = r;
    margin[i] = valid ? 2.02f * (epsd[i] + l * el + r * er) + 1e-7f : 0.f;
}

__global__ void k_reset(Counters *ctr, unsigned long long *best, int32_t *rowflag, unsigned int *rowcnt,
                        unsigned int *gthr, int64_t P, int flag_all)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i == 0) {
        ctr->cand_count = 0; ctr->n_exact = 0; ctr->n_unsafe = 0;
        ctr->overflow = flag_all ? 1u : 0u; ctr->n_flag_rows = 0; ctr->xfloor_bits = 0; ctr->ovf_count = 0;
        ctr->cyc_epi_busy = ctr->cyc_epi_wait = ctr->cyc_mma_wait_full = ctr->cyc_mma_wait_acc = 0;
        ctr->cyc_prod_wait = ctr->cyc_total = 0;
    }
    if (i < P) { best[i] = ~0ull; rowflag[i] = flag_all; rowcnt[i] = 0; gthr[i] = 0x7f800000u; }
}

// -------------------------------------------------------------------------------------------
// Screening kernel
// -------------------------------------------------------------------------------------------
struct ScreenParams {
    const float *gnorm;
    int64_t G;
    int32_t KB;             // K blocks of 64 elements
    int32_t BN;
    int64_t NT, GT;
    int32_t TPU;            // gallery tiles per work unit
    int64_t n_chunks;       // gallery chunks (of TPU tiles)
    int64_t NTG;            // probe tile groups: NT / (CTA pairs per cluster)
    int64_t n_units;        // n_chunks * NTG, chunk-major: concurrent clusters share a gallery chunk
    int32_t order;          // unit order: 0 chunk-major, 1 probe-tile-major, 2 diagonal (rotated chunks)
    int64_t g_stride;       // gallery row stride (1; > 1 in the seed pass)
    int32_t seed_mode;      // 1: only tighten the thresholds, append nothing
    const float *na, *wl, *wr, *margin;
    const int32_t *rowmap;
    unsigned int *gthr;
    Cand *cand;
    unsigned int *rowcnt;
    int32_t cand_cap;       // per probe row
    OvfCand *ovf;           // shared spill-over buffer
    int32_t ovf_cap;
    Counters *ctr;
    int32_t *rowflag;
    uint32_t idesc;
    float *dbg;             // optional [P,G] dump of the screening values
    int32_t exp_mode;       // experiments (EOSVR_EXP): 1 = epilogue releases the accumulator untouched,
                            // 2 = producer skips the TMA loads, 4 = MMA issuer skips the MMAs
};

struct StagedCand {
    int32_t rm;         // probe row
    int32_t g;          // gallery row (local to the shard)
    uint32_t tbits;
};
constexpr int kWarpStage = 64;

struct __align__(16) ScreenSmemTail {
    float na[kMaxBN];
    float wl[kMaxBN];
    float wr[kMaxBN];
    float mg[kMaxBN];
    unsigned int thr[kMaxBN];
    int32_t row[kMaxBN];
    uint64_t full[kStages];           // stage s is always consumed by issuer s % kIssuers: a parity wait is only safe
                                      // on a barrier whose every phase the waiter observes
    uint64_t empty[kStages];
    uint64_t tfull[kAccStages];
    uint64_t tempty[kAccStages];
    uint64_t tfirst[kAccStages];      // the overwriting first stage of a tile has completed
    uint32_t tmem_base;
    // candidate staging: every epilogue warp parks its candidates in shared memory and hands them to the
    // per-row lists 32 at a time (one global atomicAdd per lane, all in flight together) instead of stalling on
    // one atomic round trip per column
    StagedCand stage[kEpiWarps][kWarpStage];
};

constexpr int kABytes = kBM * kBK * 2;              // 16 KiB: this CTA's 128 gallery rows
constexpr int kBBytes = (kMaxBN / 2) * kBK * 2;     // 16 KiB: this CTA's half of the probe tile
constexpr int kAStage = kSub * kABytes;             // one pipeline stage = kSub K blocks of each operand
constexpr int kBStage = kSub * kBBytes;
constexpr size_t kScreenSmem = 1024 + static_cast<size_t>(kStages) * (kAStage + kBStage) + sizeof(ScreenSmemTail);

struct UnitIter {
    int64_t u, jt, gt0, gt1;    // jt: probe tile GROUP of the unit (pair q of the cluster takes tile jt*NP + q)
};

// Work units (gallery chunk x probe tile), three orders (EOSVR_ORDER):
//   1 (default) probe-tile-major: unit u = probe_tile * n_chunks + chunk.  The pairs that run at the same time work
//     on the same few probe tiles against different gallery chunks, so a probe tile's K blocks are fetched from HBM
//     once (L2 hits for the other pairs) and the whole 16-bit gallery stays L2-resident: DRAM traffic close to the
//     algorithmic bytes.  ~10 % more candidates than chunk-major (several pairs screen the same probe rows at once
//     and see each other's thresholds one tile late) and still 1-2 % faster;
//   0 chunk-major: unit u = chunk * NT + probe_tile.  Tightest thresholds, but every probe tile is re-read from HBM
//     once per gallery chunk (measured 4.7x the algorithmic bytes at the bench size);
//   2 diagonal (rotated chunks).
__device__ __forceinline__ bool decode_unit(const ScreenParams &p, int64_t u, UnitIter &it)
{
    int64_t chunk;
    it.u = u;
    if (p.order == 1) {
        it.jt = u / p.n_chunks;
        chunk = u % p.n_chunks;
    } else {
        chunk = u / p.NTG;
        it.jt = u % p.NTG;
        if (p.order == 2) chunk = (chunk + it.jt) % p.n_chunks;   // every (tile, chunk) still visited once
    }
    it.gt0 = chunk * p.TPU;
    it.gt1 = min(it.gt0 + p.TPU, p.GT);
    return it.gt0 < it.gt1;
}

// NP = CTA pairs per cluster (cluster size 2*NP).  With NP = 2 the two pairs of a cluster screen the SAME
// gallery tile against two different probe tiles: every CTA fetches half of its 128-row gallery slab and
// multicasts it to the CTA of equal parity in the other pair, so the gallery operand crosses L2 -> SM once
// per cluster instead of once per pair (the kernel is bound by operand delivery).
// t[j] for a warp-uniform j without indexing the register array dynamically.
__device__ __forceinline__ float pick16(const float (&t)[kChunk], int j)
{
    switch (j) {
        case 0: return t[0];   case 1: return t[1];   case 2: return t[2];   case 3: return t[3];
        case 4: return t[4];   case 5: return t[5];   case 6: return t[6];   case 7: return t[7];
        case 8: return t[8];   case 9: return t[9];   case 10: return t[10]; case 11: return t[11];
        case 12: return t[12]; case 13: return t[13]; case 14: return t[14]; default: return t[15];
    }
}

// Store candidate (g, tbits) of probe row rm at list position pos; full lists spill to the shared buffer and only
// if that is full too is the row handed to the exhaustive exact kernel.
__device__ __forceinline__ void put_candidate(const ScreenParams &p, int32_t rm, unsigned pos, int32_t g, uint32_t tbits)
{
    if (pos < static_cast<unsigned>(p.cand_cap)) {
        Cand cd;
        cd.g = g; cd.tbits = tbits;
        p.cand[static_cast<int64_t>(rm) * p.cand_cap + pos] = cd;
    } else {
        const unsigned op = atomicAdd(&p.ctr->ovf_count, 1u);
        if (op < static_cast<unsigned>(p.ovf_cap)) {
            OvfCand oc;
            oc.p = rm; oc.g = g; oc.tbits = tbits; oc.pad = 0;
            p.ovf[op] = oc;
        } else {
            p.rowflag[rm] = 1;
            p.ctr->overflow = 1u;
        }
    }
}

// Hand the n (< 64) candidates parked by a warp to their rows' lists: one atomicAdd per lane, all in flight together.
__device__ __forceinline__ void flush_staged(const ScreenParams &p, const StagedCand *st, int n, int lane)
{
    __syncwarp();
    for (int e = lane; e < n; e += 32) {
        const StagedCand sc = st[e];
        const unsigned pos = atomicAdd(p.rowcnt + sc.rm, 1u);
        put_candidate(p, sc.rm, pos, sc.g, sc.tbits);
    }
    __syncwarp();
}

template <int NP>
__global__ void __launch_bounds__(kThreads, 1)
k_match_screen(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ScreenParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem;
    uint8_t *sB = smem + kStages * kAStage;
    ScreenSmemTail *tl = reinterpret_cast<ScreenSmemTail *>(smem + kStages * (kAStage + kBStage));
    const int KS = (p.KB + kSub - 1) / kSub;            // pipeline stages per gallery tile

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = crank & 1u;                   // 0 = leader of its pair (issues the MMAs), 1 = peer
    const uint32_t pq = crank >> 1;                     // pair within the cluster
    const uint32_t leader = crank & ~1u;                // cluster rank of this pair's leader
    const int64_t pair = blockIdx.x / (2 * NP), npairs = gridDim.x / (2 * NP);   // cluster index / count

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        // NP > 1: every CTA collects its own bytes on its own full barrier and the peer relays its completion
        // to the leader (second arrival); the stage is free once the MMAs of ALL pairs have read it.
        const uint32_t full_count = (NP > 1 && rank == 0) ? 2u : 1u;
        for (int s = 0; s < kStages; ++s) { mbar_init(&tl->full[s], full_count); mbar_init(&tl->empty[s], NP); }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&tl->tfull[s], kIssuers);          // one commit per MMA issuer warp
            mbar_init(&tl->tempty[s], 2 * kEpiWarps);
            mbar_init(&tl->tfirst[s], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_2sm(&tl->tmem_base, kTmemCols);
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = tl->tmem_base;
    const bool prof = (p.exp_mode & 16) != 0;

    // Producer and issuer warps run their loops with the whole warp; lane 0 issues.
    if (warp == 0 || (kTwoProducers && warp == kProdBWarp)) {
        // ===== TMA producers (every CTA): warp 0 loads the CTA's gallery rows (A), warp 12 its half of the pair's
        //       probe tile (B) -- a UTMALDG occupies its issuing thread for ~150 cycles, so one thread per operand.
        //       Warp 0 also posts the byte count of the whole stage. =====
        const bool isA = warp == 0, isB = kTwoProducers ? warp == kProdBWarp : true;   // one warp may play both roles
        int stage = 0; uint32_t phase = 0;
        const int32_t bhalf = p.BN >> 1;
        const uint32_t tx_cta = kABytes + static_cast<uint32_t>(bhalf) * kBK * 2;
        uint16_t mc_mask = 0;
        for (int k = 0; k < NP; ++k) mc_mask |= static_cast<uint16_t>(1u << (2 * k + rank));
        unsigned long long w_prod = 0;
        const long long t_begin = clock64();
        const uint32_t lbar0 = mapa(smem_u32(&tl->full[0]), 0);
        for (int64_t u = pair; u < p.n_units; u += npairs) {
            UnitIter it;
            if (!decode_unit(p, u, it)) continue;
            const int32_t brow = static_cast<int32_t>((it.jt * NP + pq) * p.BN + rank * bhalf);
            for (int64_t gt = it.gt0; gt < it.gt1; ++gt) {
                const int32_t arow = static_cast<int32_t>(gt * kPairM + rank * kBM + (NP > 1 ? pq * (kBM / NP) : 0));
                for (int ks = 0; ks < KS; ++ks) {
                    const int nsub = min(kSub, p.KB - ks * kSub);
                    if (prof) { const long long t0 = clock64(); mbar_wait(&tl->empty[stage], phase ^ 1); w_prod += clock64() - t0; }
                    else mbar_wait(&tl->empty[stage], phase ^ 1);
                    if (lane == 0) {
                        if (p.exp_mode & 2) {
                            if (isA && (NP > 1 || rank == 0)) mbar_arrive(&tl->full[stage]);
                        } else if (NP == 1) {
                            // the bytes of both CTAs complete on the LEADER's full barrier: only the leader
                            // arrives; the peer cannot run ahead of the phase because its stage is freed by
                            // the leader's MMA commit
                            const uint32_t lbar = lbar0 + stage * static_cast<uint32_t>(sizeof(uint64_t));
                            if (isA && rank == 0) mbar_arrive_expect_tx(&tl->full[stage], 2u * tx_cta * nsub);
                            for (int sb = 0; sb < nsub; ++sb) {
                                const int32_t kc = (ks * kSub + sb) * kBK;
                                if (isA) tma_load_2d_2sm(sA + stage * kAStage + sb * kABytes, &tmA, lbar, kc, arow);
                                if (isB) tma_load_2d_2sm(sB + stage * kBStage + sb * kBBytes, &tmB, lbar, kc, brow);
                            }
                        } else {
                            if (isA) mbar_arrive_expect_tx(&tl->full[stage], tx_cta * nsub);
                            for (int sb = 0; sb < nsub; ++sb) {
                                const int32_t kc = (ks * kSub + sb) * kBK;
                                if (isA) tma_load_2d_mc(sA + stage * kAStage + sb * kABytes + pq * (kABytes / NP), &tmA,
                                                        &tl->full[stage], kc, arow, mc_mask);
                                if (isB) tma_load_2d(sB + stage * kBStage + sb * kBBytes, &tmB, &tl->full[stage], kc, brow);
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        if (prof && lane == 0 && isA) {
            atomicAdd(&p.ctr->cyc_prod_wait, w_prod);
            if (rank == 0) atomicAdd(&p.ctr->cyc_total, static_cast<unsigned long long>(clock64() - t_begin));
        }
    } else if (warp >= 1 && warp <= kIssuers && rank == 0) {
        // ===== MMA issuers: kIssuers warps of the pair's leader drive the tensor cores of both SMs, taking the
        //       pipeline stages round-robin (stage s -> issuer s % kIssuers) into the same accumulator.  kZeroAcc:
        //       every MMA accumulates onto an accumulator the epilogue left zeroed.  Otherwise the first stage of a
        //       tile overwrites the accumulator: its owner commits to tfirst[acc] and the other warps wait for that
        //       before accumulating on top.  Every issuer observes every phase of tempty / tfirst. =====
        const int w = warp - 1;
        int64_t seq = 0;                                  // running stage number of this pair
        int acc = 0; uint32_t accphase = 0;
        const uint16_t all_mask = static_cast<uint16_t>((1u << (2 * NP)) - 1u);
        const uint16_t pair_mask = static_cast<uint16_t>(3u << leader);
        const uint16_t self_mask = static_cast<uint16_t>(1u << leader);
        unsigned long long w_full = 0, w_acc = 0;
        for (int64_t u = pair; u < p.n_units; u += npairs) {
            UnitIter it;
            if (!decode_unit(p, u, it)) continue;
            for (int64_t gt = it.gt0; gt < it.gt1; ++gt) {
                const bool first_owner = static_cast<int>(seq % kIssuers) == w;   // issues the overwriting stage
                // kZeroAcc: phase 0 of tempty is the epilogue's initial zeroing, phase k+1 the release after the k-th use
                const uint32_t epar = kZeroAcc ? accphase : (accphase ^ 1);
                if (prof) { const long long t0 = clock64(); mbar_wait(&tl->tempty[acc], epar); w_acc += clock64() - t0; }
                else mbar_wait(&tl->tempty[acc], epar);
                if (!kZeroAcc && !first_owner) mbar_wait(&tl->tfirst[acc], accphase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * kMaxBN;
                for (int ks = 0; ks < KS; ++ks, ++seq) {
                    if (static_cast<int>(seq % kIssuers) != w) continue;
                    const int stage = static_cast<int>(seq % kStages);
                    const uint32_t phase = static_cast<uint32_t>(seq / kStages) & 1u;
                    const int nsub = min(kSub, p.KB - ks * kSub);
                    if (prof) { const long long t0 = clock64(); mbar_wait(&tl->full[stage], phase); w_full += clock64() - t0; }
                    else mbar_wait(&tl->full[stage], phase);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + stage * kAStage);
                    const uint32_t b0 = smem_u32(sB + stage * kBStage);
                    if (lane == 0) {
                        if (!(p.exp_mode & 4)) {
                            for (int sb = 0; sb < nsub; ++sb) {
#pragma unroll
                                for (int k = 0; k < kBK / 16; ++k)
                                    mma_f16_ss_2sm(d_tmem, umma_desc_sw128(a0 + sb * kABytes, k * 32),
                                                   umma_desc_sw128(b0 + sb * kBBytes, k * 32), p.idesc,
                                                   (kZeroAcc || (ks | sb | k) != 0) ? 1u : 0u);
                            }
                        }
                        mma_commit_2sm(&tl->empty[stage], all_mask);    // one of the NP arrivals that free the stage
                        if (!kZeroAcc && ks == 0) mma_commit_2sm(&tl->tfirst[acc], self_mask);
                    }
                    __syncwarp();
                }
                if (!kZeroAcc && first_owner) mbar_wait(&tl->tfirst[acc], accphase);   // long complete; keeps the phase observed
                if (lane == 0) mma_commit_2sm(&tl->tfull[acc], pair_mask);   // this warp's share of the tile is done
                __syncwarp();
                if (++acc == kAccStages) { acc = 0; accphase ^= 1; }
            }
        }
        if (prof && lane == 0 && w == 0) atomicAdd(&p.ctr->cyc_mma_wait_acc, w_acc);
        if (prof && lane == 0) atomicAdd(&p.ctr->cyc_mma_wait_full, w_full);
    } else if (NP > 1 && warp == 1 && rank == 1) {
        // ===== relay (peer CTA): tell the leader when this CTA's stage has landed =====
        int stage = 0; uint32_t phase = 0;
        const uint32_t lfull0 = mapa(smem_u32(&tl->full[0]), leader);
        for (int64_t u = pair; u < p.n_units; u += npairs) {
            UnitIter it;
            if (!decode_unit(p, u, it)) continue;
            for (int64_t n = (it.gt1 - it.gt0) * KS; n > 0; --n) {
                mbar_wait(&tl->full[stage], phase);
                if (lane == 0) mbar_arrive_cluster(lfull0 + stage * static_cast<uint32_t>(sizeof(uint64_t)));
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 4 + kEpiWarps) {
        // ===== epilogue: warp%4 selects the TMEM lane quadrant, (warp-4)/4 the group of column chunks =====
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int te = threadIdx.x - 128;
        const int BN = p.BN;
        const int nchunks = BN / kChunk;
        const int cbeg = nchunks * half / kEpiGroups;
        const int cend = nchunks * (half + 1) / kEpiGroups;
        const float xfloor = __uint_as_float(p.ctr->xfloor_bits);
        const float dfloor = sqrtf(xfloor) * 1.000001f;
        int acc = 0; uint32_t accphase = 0;
        const uint32_t tempty_leader0 = mapa(smem_u32(&tl->tempty[0]), leader);
        const uint32_t tempty_leader1 = mapa(smem_u32(&tl->tempty[1]), leader);
        unsigned long long e_busy = 0, e_wait = 0;
        if (kZeroAcc) {
            // column group `half` of this lane quadrant zeroes accumulator `half`; every warp then reports both
            // accumulators (a barrier completes only when all 2*kEpiWarps warps of the pair have arrived)
            static_assert(!kZeroAcc || kEpiGroups == kAccStages, "initial zeroing: one column group per accumulator");
            const uint32_t tz = tmem_base + static_cast<uint32_t>(half) * kMaxBN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < kMaxBN; c += kChunk) tmem_st_zero_x16(tz + c);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive_cluster(tempty_leader0); mbar_arrive_cluster(tempty_leader1); }
        }
        StagedCand *wstage = tl->stage[warp - 4];
        int wn = 0;                                       // candidates parked by this warp (warp-uniform)
        for (int64_t u = pair; u < p.n_units; u += npairs) {
            UnitIter it;
            if (!decode_unit(p, u, it)) continue;
            // per-unit column arrays -> smem
            named_bar_sync(1, 32 * kEpiWarps);
            if (te < BN) {
                const int64_t c = (it.jt * NP + pq) * BN + te;
                tl->na[te] = p.na[c];
                tl->wl[te] = p.wl[c];
                tl->wr[te] = p.wr[c];
                tl->mg[te] = p.margin[c];
                const int32_t rm = p.rowmap[c];
                tl->row[te] = rm;
                tl->thr[te] = rm >= 0 ? *reinterpret_cast<volatile unsigned int *>(p.gthr + rm)
                                      : __float_as_uint(-1.0f);
            }
            named_bar_sync(1, 32 * kEpiWarps);

            for (int64_t gt = it.gt0; gt < it.gt1; ++gt) {
                // pick up thresholds tightened by other CTAs since the last tile (benign race)
                if (te < BN) {
                    const int32_t rm = tl->row[te];
                    if (rm >= 0) atomicMin(&tl->thr[te], *reinterpret_cast<volatile unsigned int *>(p.gthr + rm));
                }
                // this lane's gallery row and its squared norm: requested BEFORE waiting for the accumulator so the
                // global-load latency hides behind the wait
                const int64_t g = (gt * kPairM + rank * kBM + q * 32 + lane) * p.g_stride;
                const float nb = p.gnorm[g];
                long long t_e0 = 0;
                if (prof) { const long long t0 = clock64(); mbar_wait(&tl->tfull[acc], accphase); t_e0 = clock64(); e_wait += t_e0 - t0; }
                else mbar_wait(&tl->tfull[acc], accphase);
                tc_fence_after();
                const uint32_t trow = tmem_base + static_cast<uint32_t>(acc) * kMaxBN + (static_cast<uint32_t>(q * 32) << 16);

                float dprev = kBig;
                if (p.exp_mode & 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
                    if (++acc == kAccStages) { acc = 0; accphase ^= 1; }
                    continue;
                }
                uint32_t va[kChunk], vb[kChunk];
                uint32_t vleft = 0, vright = 0;          // the neighbouring column groups' boundary columns
                {   // first chunk and the boundary columns of the neighbouring column groups: one wait for all
                    if (cbeg > 0) tmem_ld_x1(trow + cbeg * kChunk - 1, vleft);
                    if (cend * kChunk < BN) tmem_ld_x1(trow + cend * kChunk, vright);
                    tmem_ld_x16(trow + cbeg * kChunk, va);
                    tmem_ld_wait();
                    if (cbeg > 0) {
                        const float x = fmaf(-2.f, __uint_as_float(vleft), nb) + tl->na[cbeg * kChunk - 1];
                        dprev = sqrt_approx(fabsf(x));
                    }
                    // kZeroAcc: nobody zeroes a column before every warp of the lane quadrant holds its boundary reads
                    if (kZeroAcc) named_bar_sync(2 + q, 32 * kEpiGroups);
                }
                // One chunk of 16 accumulator columns.  `v` holds the chunk (already loaded); the NEXT chunk (or just its
                // first column, the right-hand neighbour of column 15) is requested before the arithmetic on `v` starts
                // and waited for only when the taps need it, so the TMEM read latency overlaps the sqrt chain.
                auto do_chunk = [&](const int ch, uint32_t (&v)[kChunk], uint32_t (&vnx)[kChunk]) {
                    const int c0 = ch * kChunk;
                    const bool hasn = (c0 + kChunk) < BN;
                    if (ch + 1 < cend) tmem_ld_x16(trow + c0 + kChunk, vnx);
                    else vnx[0] = vright;

                    float d[kChunk];
                    float minx = kBig;
                    const float4 *na4 = reinterpret_cast<const float4 *>(tl->na + c0);
#pragma unroll
                    for (int j4 = 0; j4 < kChunk / 4; ++j4) {
                        const float4 a = na4[j4];
                        const float aa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = j4 * 4 + jj;
                            const float x = fmaf(-2.f, __uint_as_float(v[j]), nb) + aa[jj];
                            minx = fminf(minx, x);
                            d[j] = sqrt_approx(fabsf(x));
                        }
                    }
                    tmem_ld_wait_x16(vnx);
                    float dn = kBig;
                    if (hasn) {
                        const float x = fmaf(-2.f, __uint_as_float(vnx[0]), nb) + tl->na[c0 + kChunk];
                        dn = sqrt_approx(fabsf(x));
                    }
                    const float dprev_in = dprev;
                    dprev = d[kChunk - 1];

                    bool any = false;
                    float t[kChunk];
                    const float4 *wl4 = reinterpret_cast<const float4 *>(tl->wl + c0);
                    const float4 *wr4 = reinterpret_cast<const float4 *>(tl->wr + c0);
                    const float4 *th4 = reinterpret_cast<const float4 *>(tl->thr + c0);
#pragma unroll
                    for (int j4 = 0; j4 < kChunk / 4; ++j4) {
                        const float4 l = wl4[j4], r = wr4[j4], th = th4[j4];
                        const float ll[4] = {l.x, l.y, l.z, l.w};
                        const float rr[4] = {r.x, r.y, r.z, r.w};
                        const float tt[4] = {th.x, th.y, th.z, th.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = j4 * 4 + jj;
                            const float dl = j ? d[j - 1] : dprev_in;
                            const float dr = (j < kChunk - 1) ? d[j + 1] : dn;
                            t[j] = fmaf(ll[jj], dl, fmaf(rr[jj], dr, d[j]));
                            any |= (t[j] <= tt[jj]);
                        }
                    }
                    const bool guard = (minx < xfloor) || (fminf(dprev_in, dn) < dfloor);
                    if (__any_sync(0xffffffffu, any || guard) && !(p.exp_mode & 32)) {
                        // ---- rare path (whole warp, kept small: one loop body, no unrolling).  Only the columns in
                        //      which some lane is below its threshold (or, if the cancellation guard fired, all
                        //      columns) are visited.  The column and its two neighbours are read again from TMEM
                        //      (same arithmetic, same bits as above) so no register array is indexed dynamically.
                        //      First tighten the threshold with the warp minimum, then park what is still below it
                        //      in the warp's staging buffer. ----
                        unsigned cm = 0;
                        {
                            const float4 *th4b = reinterpret_cast<const float4 *>(tl->thr + c0);
#pragma unroll
                            for (int j4 = 0; j4 < kChunk / 4; ++j4) {
                                const float4 th = th4b[j4];
                                cm |= (t[j4 * 4 + 0] <= th.x ? 1u : 0u) << (j4 * 4 + 0);
                                cm |= (t[j4 * 4 + 1] <= th.y ? 1u : 0u) << (j4 * 4 + 1);
                                cm |= (t[j4 * 4 + 2] <= th.z ? 1u : 0u) << (j4 * 4 + 2);
                                cm |= (t[j4 * 4 + 3] <= th.w ? 1u : 0u) << (j4 * 4 + 3);
                            }
                        }
                        if (guard) cm = 0xFFFFu;
                        cm = __reduce_or_sync(0xffffffffu, cm);
                        const bool rowok = g < p.G;
                        const bool any_guard = __any_sync(0xffffffffu, guard);
#pragma unroll 1
                        while (cm) {
                            const int c = c0 + __ffs(cm) - 1;              // warp-uniform
                            cm &= cm - 1;
                            const int32_t rm = tl->row[c];
                            if (rm < 0) continue;                          // warp-uniform
                            float tj;
                            bool uns = false;
                            if (!any_guard) {
                                tj = pick16(t, c - c0);                    // warp-uniform index: a jump, no local memory
                            } else {
                                // some lane is inside the cancellation guard: the column and its neighbours again,
                                // straight from TMEM (same arithmetic, same bits as the fast path)
                                uint32_t vc, vl = 0, vr = 0;
                                const bool hl = c > 0, hr = c + 1 < BN;
                                tmem_ld_x1(trow + c, vc);
                                if (hl && c != cbeg * kChunk) tmem_ld_x1(trow + c - 1, vl);
                                if (hr && c + 1 != cend * kChunk) tmem_ld_x1(trow + c + 1, vr);
                                tmem_ld_wait();
                                if (c == cbeg * kChunk) vl = vleft;                 // other column groups' columns may
                                if (c + 1 == cend * kChunk) vr = vright;            // already be zeroed: use the early reads
                                const float dj = sqrt_approx(fabsf(fmaf(-2.f, __uint_as_float(vc), nb) + tl->na[c]));
                                const float dl = hl ? sqrt_approx(fabsf(fmaf(-2.f, __uint_as_float(vl), nb) + tl->na[c - 1])) : kBig;
                                const float dr = hr ? sqrt_approx(fabsf(fmaf(-2.f, __uint_as_float(vr), nb) + tl->na[c + 1])) : kBig;
                                const float wlc = tl->wl[c], wrc = tl->wr[c];
                                tj = fmaf(wlc, dl, fmaf(wrc, dr, dj));
                                const float m3 = fminf(dj, fminf(wlc > 0.f ? dl : kBig, wrc > 0.f ? dr : kBig));
                                uns = rowok && (m3 < dfloor);
                            }
                            float thr = __uint_as_float(*reinterpret_cast<volatile unsigned int *>(&tl->thr[c]));
                            bool pass = rowok && !uns && (tj <= thr);
                            if (__ballot_sync(0xffffffffu, pass)) {
                                // warp minimum in one instruction: t >= 0, so float order == unsigned order of the bits
                                const float mn = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(pass ? tj : kBig)));
                                const float nt = fmaf(mn, kSlopMul, tl->mg[c]);
                                if (nt < thr) {
                                    if (lane == 0) {
                                        atomicMin(&tl->thr[c], __float_as_uint(nt));
                                        atomicMin(p.gthr + rm, __float_as_uint(nt));
                                    }
                                    thr = nt;
                                }
                                pass = pass && (tj <= thr);
                            }
                            if (p.seed_mode) continue;
                            const bool app = pass || uns;
                            const unsigned ma = __ballot_sync(0xffffffffu, app);
                            if (ma) {
                                if (app) {
                                    StagedCand sc;
                                    sc.rm = rm; sc.g = static_cast<int32_t>(g);
                                    sc.tbits = uns ? kCandUnsafe : __float_as_uint(tj);
                                    wstage[wn + __popc(ma & ((1u << lane) - 1u))] = sc;
                                }
                                wn += __popc(ma);
                                if (wn >= 32) { flush_staged(p, wstage, wn, lane); wn = 0; }
                            }
                        }
                    }
                    if (p.dbg != nullptr && g < p.G && !p.seed_mode) {
#pragma unroll
                        for (int j = 0; j < kChunk; ++j) {
                            const int32_t rm = tl->row[c0 + j];
                            if (rm >= 0) p.dbg[static_cast<int64_t>(rm) * p.G + g] = t[j];
                        }
                    }
                };
                for (int ch = cbeg; ch < cend; ch += 2) {
                    do_chunk(ch, va, vb);
                    // zero one chunk behind: the rare path of chunk ch may still read the last column of chunk ch - 1
                    if (kZeroAcc && ch > cbeg) tmem_st_zero_x16(trow + (ch - 1) * kChunk);
                    if (ch + 1 < cend) {
                        do_chunk(ch + 1, vb, va);
                        if (kZeroAcc) tmem_st_zero_x16(trow + ch * kChunk);
                    }
                }
                if (kZeroAcc && cend > cbeg) { tmem_st_zero_x16(trow + (cend - 1) * kChunk); tmem_st_wait(); }   // (a column
                                                                          // group is empty when the tile has one chunk)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
                if (++acc == kAccStages) { acc = 0; accphase ^= 1; }
                if (prof) e_busy += clock64() - t_e0;
            }
        }
        if (wn) flush_staged(p, wstage, wn, lane);
        if (prof && lane == 0) { atomicAdd(&p.ctr->cyc_epi_busy, e_busy); atomicAdd(&p.ctr->cyc_epi_wait, e_wait); }
    }

    tc_fence_before();
    cluster_sync();          // neither CTA may exit (or free TMEM) while its peer can still reach it
    if (warp == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
}

// -------------------------------------------------------------------------------------------
// Exact evaluation (the reference's arithmetic): network_test.py:208 (scipy cdist, float64 direct
// differences), :109 (float32 cast), models.py:42-56 taps as the float32 FMA chain
//   acc = lam1*d[p-1];  acc = fma(lam2, d[p], acc);  acc = fma(lam1, d[p+1], acc)
// with zero padding at episode ends.  Warp-cooperative; all lanes return the value.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float exact_t(const float *__restrict__ probes, int64_t P, int D, int rpe,
                                         int64_t p, const float *__restrict__ b, float lam1, float lam2, int lane)
{
    const int r = static_cast<int>(p % rpe);
    const bool hl = r > 0, hr = (r + 1 < rpe) && (p + 1 < P);
    const float *a1 = probes + p * D;
    const float *a0 = hl ? a1 - D : a1;
    const float *a2 = hr ? a1 + D : a1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int k = lane; k < D; k += 32) {
        const double bv = b[k];
        const double e0 = static_cast<double>(a0[k]) - bv;
        const double e1 = static_cast<double>(a1[k]) - bv;
        const double e2 = static_cast<double>(a2[k]) - bv;
        s0 += e0 * e0; s1 += e1 * e1; s2 += e2 * e2;
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    const float d0 = hl ? static_cast<float>(sqrt(s0)) : 0.f;
    const float d1 = static_cast<float>(sqrt(s1));
    const float d2 = hr ? static_cast<float>(sqrt(s2)) : 0.f;
    float acc = __fmul_rn(lam1, d0);
    acc = __fmaf_rn(lam2, d1, acc);
    acc = __fmaf_rn(lam1, d2, acc);
    return acc;
}

struct RerankParams {
    const float *probes;
    const float *gal;
    int64_t P, G, offset;
    int32_t D, rpe;
    int32_t metric;          // EOSVR_METRIC_*; the cosine metric works in the score domain s = -cosine (a minimum)
    float lam1, lam2;
    const Cand *cand;
    const unsigned int *rowcnt;
    int32_t cand_cap;
    Counters *ctr;
    const unsigned int *gthr;
    unsigned long long *best;
    int32_t *rowflag;
    int32_t *flaglist;
    const OvfCand *ovf;
    int32_t ovf_cap;
    const float *margin;     // per plan column
    int32_t planR, planBN, planHalo;
    int32_t prof;            // EOSVR_EXP bit 64: phase timing of k_rerank_rows into the cycle counters
    int32_t rows_per_block;  // consecutive probe rows per k_rerank_rows block (1..kRrRowsPerBlock)
};

// Cosine metric, exactly: float64 dot / (|a| |b|) on the original rows (0 for a zero row), rounded to
// float32; returned negated (score domain).  Warp-cooperative; all lanes return the value.
__device__ __forceinline__ float exact_negcos(const float *__restrict__ a, const float *__restrict__ b, int D, int lane)
{
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (int k = lane; k < D; k += 32) {
        const double x = a[k], y = b[k];
        dot += x * y; na += x * x; nb += y * y;
    }
    dot = warp_sum(dot); na = warp_sum(na); nb = warp_sum(nb);
    const double den = sqrt(na) * sqrt(nb);
    return 0.f - (den > 0.0 ? static_cast<float>(dot / den) : 0.f);   // 0 - c: never -0 (packed order)
}

__device__ __forceinline__ float exact_score(const RerankParams &p, int64_t row, int64_t g, int lane)
{
    if (p.metric == EOSVR_METRIC_COSINE) return exact_negcos(p.probes + row * p.D, p.gal + g * p.D, p.D, lane);
    return exact_t(p.probes, p.P, p.D, p.rpe, row, p.gal + g * p.D, p.lam1, p.lam2, lane);
}

// Spill-over candidates (row lists that filled up): one warp per entry.  Exits at once when empty.
__global__ void k_rerank_ovf(const RerankParams p)
{
    const unsigned cnt = p.ctr->ovf_count;
    if (cnt == 0) return;
    const int64_t n = cnt < static_cast<unsigned>(p.ovf_cap) ? cnt : p.ovf_cap;
    const int lane = threadIdx.x & 31;
    const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    unsigned long long done = 0;
    for (int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n; w += nw) {
        const OvfCand c = p.ovf[w];
        if (c.tbits != kCandUnsafe && __uint_as_float(c.tbits) > __uint_as_float(p.gthr[c.p])) continue;
        const float t = exact_score(p, c.p, c.g, lane);
        if (lane == 0) { atomicMin(p.best + c.p, pack_score_idx(t, static_cast<uint32_t>(p.offset + c.g))); ++done; }
    }
    if (lane == 0 && done) atomicAdd(&p.ctr->n_exact, done);
}

// One warp per probe row: keep the candidates still below the row's FINAL threshold, evaluate them
// exactly, and keep the smallest packed (score, index).  The three probe rows stay in L1 across the
// row's candidates.
__global__ void k_rerank(const RerankParams p)
{
    const int lane = threadIdx.x & 31;
    const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    unsigned long long appended = 0, done = 0, uns = 0;
    for (int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < p.P; row += nw) {
        const unsigned cnt = p.rowcnt[row];
        const int n = cnt < static_cast<unsigned>(p.cand_cap) ? static_cast<int>(cnt) : p.cand_cap;
        appended += cnt;
        const float thr = __uint_as_float(p.gthr[row]);
        const Cand *list = p.cand + row * p.cand_cap;
        unsigned long long loc = ~0ull;
        for (int b0 = 0; b0 < n; b0 += 32) {
            Cand c;
            c.g = 0; c.tbits = 0x7f800000u;
            if (b0 + lane < n) c = list[b0 + lane];
            const bool keep = (b0 + lane < n) && (c.tbits == kCandUnsafe || __uint_as_float(c.tbits) <= thr);
            unsigned m = __ballot_sync(0xffffffffu, keep);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int32_t g = __shfl_sync(0xffffffffu, c.g, src);
                const uint32_t tb = __shfl_sync(0xffffffffu, c.tbits, src);
                const float t = exact_score(p, row, g, lane);
                const unsigned long long v = pack_score_idx(t, static_cast<uint32_t>(p.offset + g));
                loc = v < loc ? v : loc;
                ++done; uns += (tb == kCandUnsafe);
            }
        }
        if (lane == 0 && loc != ~0ull) atomicMin(p.best + row, loc);
    }
    if (lane == 0) {
        if (appended) atomicAdd(&p.ctr->cand_count, appended);
        if (done) atomicAdd(&p.ctr->n_exact, done);
        if (uns) atomicAdd(&p.ctr->n_unsafe, uns);
    }
}

// Block-per-row re-rank, warp-per-candidate.  The three probe rows the taps need are staged in shared memory
// once per row; then, in two steps:
//   (1) float32 pre-filter: the candidates still below the row's threshold are sorted by screening value and the
//       block's warps pull them in ascending order, each warp evaluating one candidate with float32 direct
//       differences (relative error of the smoothed value < kF32Rel, orders of magnitude tighter than the 16-bit
//       screening) and tightening the shared bound; a warp stops as soon as the next screening value cannot win;
//   (2) the candidates within 2*kF32Rel of the best float32 value -- the winner and its exact ties, normally ONE
//       candidate -- are evaluated exactly as the reference does (float64 direct differences -> float32 ->
//       float32 FMA chain) and merged by packed 64-bit atomicMin.
// No block-wide barrier sits between candidates, the gallery-row loads of the warps overlap, and float64 work
// (conversions run at 16/clk/SM) is spent only where it decides the answer.
constexpr int kRrThreads = 128;
constexpr float kF32Rel = 32.0f / 16777216.0f;   // 32 ulp: |t32 - t_reference| <= kF32Rel * t

constexpr float kF32AbsCos = 64.0f / 16777216.0f;   // |cos32 - cos_reference| <= 64 ulp(1) absolute

__device__ __forceinline__ unsigned int f2o(float x)    // order-preserving float -> unsigned
{
    const unsigned int b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float o2f(unsigned int b)
{
    return __uint_as_float((b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b);
}

// COS = cosine metric: one probe row is staged, the score is s = -cosine (so both metrics minimise), the
// screening value of a candidate is ~ sqrt(2 + 2 s) and the float32 error is absolute (kF32AbsCos).
constexpr int kRrRowsPerBlock = 8;   // consecutive probe rows per block: neighbours are staged once (sliding window)

template <bool COS>
__global__ void __launch_bounds__(kRrThreads)
k_rerank_rows(const RerankParams p)
{
    extern __shared__ float4 s_probe4[];          // ring of 4 probe rows [4][D/4]: row q lives in slot q % 4  (COS: 2 rows)
    __shared__ int32_t s_g[kRrThreads], s_g2[kRrThreads];
    __shared__ float s_t[kRrThreads], s_t2[kRrThreads];
    __shared__ int s_warpcnt[kRrThreads / 32];
    __shared__ unsigned int s_bound, s_best32;    // s_bound: float bits of a positive value; s_best32: f2o() order
    __shared__ int s_next, s_n32;
    __shared__ unsigned int s_cnt[kRrRowsPerBlock];
    __shared__ float s_thr[kRrRowsPerBlock], s_eps[kRrRowsPerBlock];
    __shared__ double s_part64[kRrThreads / 32][3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D, D4 = p.D >> 2;
    unsigned long long appended = 0, done = 0, unsafe_n = 0;
    const float inv_lam2 = 1.0f / p.lam2;
    const float4 *probes4 = reinterpret_cast<const float4 *>(p.probes);
    const int rpb = p.rows_per_block;
    const int64_t nblk = (p.P + rpb - 1) / rpb;
    unsigned long long c_setup = 0, c_sort = 0, c_p1 = 0, c_p2 = 0;   // tid 0, p.prof only
    long long tk = p.prof ? clock64() : 0;
#define RR_MARK(acc) do { if (p.prof && tid == 0) { const long long now = clock64(); acc += now - tk; tk = now; } } while (0)

    // probe rows live in a shared-memory ring filled with cp.async one or two rows ahead of their use, so a row's
    // global-memory latency hides behind the previous row's work
    constexpr int RING = COS ? 2 : 4, PF = COS ? 1 : 2;
    auto stage_row = [&](int64_t q) {                 // all threads: request probe row q into its ring slot
        const uint32_t dst = smem_u32(s_probe4 + static_cast<int>(q % RING) * D4);
        const float4 *src = probes4 + q * D4;
        for (int k = tid; k < D4; k += kRrThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * 16), "l"(src + k) : "memory");
    };

    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const int64_t r0 = blk * rpb, r1 = min(p.P, r0 + rpb);
      const int64_t last_needed = COS ? r1 - 1 : min(r1, p.P - 1);      // last probe row this block reads
      __syncthreads();                                  // the previous block's rows are done with the ring
      if (tid < r1 - r0) {
          const int64_t row = r0 + tid;
          s_cnt[tid] = p.rowcnt[row];
          s_thr[tid] = __uint_as_float(p.gthr[row]);
          // one-sided error bound of the row's screening values (half of the two-sided threshold margin)
          s_eps[tid] = 0.5f * p.margin[(row / p.planR) * p.planBN + p.planHalo + (row % p.planR)];
      }
      if (!COS && r0 > 0) stage_row(r0 - 1);
      for (int64_t q = r0; q < r0 + PF && q <= last_needed; ++q) stage_row(q);
      asm volatile("cp.async.commit_group;" ::: "memory");
      for (int64_t row = r0; row < r1; ++row) {
        __syncthreads();                                // the previous row is done with the slot about to be refilled
        if (row + PF <= last_needed) stage_row(row + PF);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");           // everything but the newest request has landed
        __syncthreads();
        const unsigned cnt = s_cnt[row - r0];
        const int n = cnt < static_cast<unsigned>(p.cand_cap) ? static_cast<int>(cnt) : p.cand_cap;
        if (tid == 0) appended += cnt;
        if (n == 0) continue;                                           // block-uniform
        const int r = static_cast<int>(row % p.rpe);
        const bool hl = r > 0, hr = (r + 1 < p.rpe) && (row + 1 < p.P);
        const float4 *sp1 = s_probe4 + static_cast<int>(row % RING) * D4;
        const float4 *sp0 = s_probe4 + static_cast<int>((row + RING - 1) % RING) * D4;        // row - 1
        const float4 *sp2 = s_probe4 + static_cast<int>((row + 1) % RING) * D4;
        const float thr = s_thr[row - r0];
        const float eps1 = s_eps[row - r0];
        const Cand *list = p.cand + row * p.cand_cap;
        unsigned long long loc = ~0ull;
        if (tid == 0) { s_bound = __float_as_uint(thr); s_best32 = f2o(INFINITY); }
        __syncthreads();
        RR_MARK(c_setup);

        for (int b0 = 0; b0 < n; b0 += kRrThreads) {
            // ---- keep what can still win; compact; sort ascending by screening value ----
            Cand c;
            c.g = 0; c.tbits = 0x7f800000u;
            const bool in = b0 + tid < n;
            if (in) c = list[b0 + tid];
            const bool isuns = in && c.tbits == kCandUnsafe;
            const float tv = isuns ? -INFINITY : __uint_as_float(c.tbits);
            const bool keep = in && (isuns || tv <= __uint_as_float(s_bound));
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_warpcnt[warp] = __popc(m);
            __syncthreads();
            int base = 0, ns = 0;
#pragma unroll
            for (int w = 0; w < kRrThreads / 32; ++w) { if (w < warp) base += s_warpcnt[w]; ns += s_warpcnt[w]; }
            if (keep) {
                const int pos = base + __popc(m & ((1u << lane) - 1u));
                s_g[pos] = c.g; s_t[pos] = tv;
                if (isuns) ++unsafe_n;
            }
            __syncthreads();
            if (tid < ns) {
                const float mine = s_t[tid];
                int rank = 0;
                for (int j = 0; j < ns; ++j) {
                    const float o = s_t[j];
                    rank += (o < mine) || (o == mine && j < tid);
                }
                s_g2[rank] = s_g[tid]; s_t2[rank] = mine;
            }
            if (tid == 0) { s_next = 0; s_n32 = 0; }
            __syncthreads();
            RR_MARK(c_sort);
            // ---- (1) float32 evaluation, one candidate per warp, ascending screening value.  Every candidate that
            //      could tie or beat the best exact value t* has t~ <= t*/lam2 + eps1, and t* <= best32*(1+kF32Rel):
            //      once a warp's next candidate is above the shared bound, so are all later ones. ----
            for (;;) {
                int j = 0;
                if (lane == 0) j = atomicAdd(&s_next, 1);
                j = __shfl_sync(0xffffffffu, j, 0);
                if (j >= ns) break;
                if (s_t2[j] > __uint_as_float(*reinterpret_cast<volatile unsigned int *>(&s_bound))) break;
                const int32_t g = s_g2[j];
                const float4 *gp = reinterpret_cast<const float4 *>(p.gal + static_cast<int64_t>(g) * D);
                float x0 = 0.f, x1 = 0.f, x2 = 0.f;
                if (COS) {
#pragma unroll 4
                    for (int k = lane; k < D4; k += 32) {
                        const float4 b = gp[k];
                        const float4 q = sp1[k];
                        x0 = fmaf(q.x, b.x, x0); x0 = fmaf(q.y, b.y, x0); x0 = fmaf(q.z, b.z, x0); x0 = fmaf(q.w, b.w, x0);
                        x1 = fmaf(b.x, b.x, x1); x1 = fmaf(b.y, b.y, x1); x1 = fmaf(b.z, b.z, x1); x1 = fmaf(b.w, b.w, x1);
                        x2 = fmaf(q.x, q.x, x2); x2 = fmaf(q.y, q.y, x2); x2 = fmaf(q.z, q.z, x2); x2 = fmaf(q.w, q.w, x2);
                    }
                } else {
#pragma unroll 4
                    for (int k = lane; k < D4; k += 32) {
                        const float4 b = gp[k];
                        const float4 q0 = sp0[k], q1 = sp1[k], q2 = sp2[k];
                        float e;
                        e = q0.x - b.x; x0 = fmaf(e, e, x0); e = q0.y - b.y; x0 = fmaf(e, e, x0);
                        e = q0.z - b.z; x0 = fmaf(e, e, x0); e = q0.w - b.w; x0 = fmaf(e, e, x0);
                        e = q1.x - b.x; x1 = fmaf(e, e, x1); e = q1.y - b.y; x1 = fmaf(e, e, x1);
                        e = q1.z - b.z; x1 = fmaf(e, e, x1); e = q1.w - b.w; x1 = fmaf(e, e, x1);
                        e = q2.x - b.x; x2 = fmaf(e, e, x2); e = q2.y - b.y; x2 = fmaf(e, e, x2);
                        e = q2.z - b.z; x2 = fmaf(e, e, x2); e = q2.w - b.w; x2 = fmaf(e, e, x2);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    x0 += __shfl_xor_sync(0xffffffffu, x0, o);
                    x1 += __shfl_xor_sync(0xffffffffu, x1, o);
                    x2 += __shfl_xor_sync(0xffffffffu, x2, o);
                }
                if (lane == 0) {
                    float acc, nb;
                    if (COS) {
                        const float den = sqrtf(x2) * sqrtf(x1);
                        acc = 0.f - (den > 0.f ? x0 / den : 0.f);
                        // a candidate that ties or beats the best cosine c* >= -(acc + kF32AbsCos) has normalised
                        // distance <= sqrt(2 - 2 c*)
                        nb = fmaf(sqrtf(fmaxf(2.0f + 2.0f * (acc + 2.0f * kF32AbsCos), 0.f)), 1.000001f, eps1);
                    } else {
                        const float d0 = hl ? sqrtf(x0) : 0.f;
                        const float d1 = sqrtf(x1);
                        const float d2 = hr ? sqrtf(x2) : 0.f;
                        acc = __fmul_rn(p.lam1, d0);
                        acc = __fmaf_rn(p.lam2, d1, acc);
                        acc = __fmaf_rn(p.lam1, d2, acc);
                        nb = fmaf(acc * (1.0f + kF32Rel) * inv_lam2, 1.00002f, eps1);
                    }
                    const int i32 = atomicAdd(&s_n32, 1);
                    s_g[i32] = g; s_t[i32] = acc;
                    atomicMin(&s_best32, f2o(acc));
                    atomicMin(&s_bound, __float_as_uint(nb));
                }
            }
            __syncthreads();
            RR_MARK(c_p1);
            // ---- (2) exact evaluation of everything within the float32 error of the best float32 value ----
            const int n32 = s_n32;
            const float best32 = o2f(s_best32);
            const float cut = COS ? best32 + 2.0f * kF32AbsCos : best32 * (1.0f + 2.0f * kF32Rel);
            for (int j = 0; j < n32; ++j) {                              // normally ONE candidate: the whole block on it
                if (!(s_t[j] <= cut)) continue;                          // block-uniform
                const int32_t g = s_g[j];
                const float4 *gp = reinterpret_cast<const float4 *>(p.gal + static_cast<int64_t>(g) * D);
                double y0 = 0.0, y1 = 0.0, y2 = 0.0;
                if (COS) {
                    for (int k = tid; k < D4; k += kRrThreads) {
                        const float4 b = gp[k];
                        const float4 q = sp1[k];
                        const double bb[4] = {b.x, b.y, b.z, b.w}, qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) { y0 += qq[e] * bb[e]; y1 += bb[e] * bb[e]; y2 += qq[e] * qq[e]; }
                    }
                } else {
                    for (int k = tid; k < D4; k += kRrThreads) {
                        const float4 b = gp[k];
                        const float4 q0 = sp0[k], q1 = sp1[k], q2 = sp2[k];
                        const double bb[4] = {b.x, b.y, b.z, b.w};
                        const float qq0[4] = {q0.x, q0.y, q0.z, q0.w}, qq1[4] = {q1.x, q1.y, q1.z, q1.w},
                                    qq2[4] = {q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const double e0 = static_cast<double>(qq0[e]) - bb[e];
                            const double e1 = static_cast<double>(qq1[e]) - bb[e];
                            const double e2 = static_cast<double>(qq2[e]) - bb[e];
                            y0 += e0 * e0; y1 += e1 * e1; y2 += e2 * e2;
                        }
                    }
                }
                y0 = warp_sum(y0); y1 = warp_sum(y1); y2 = warp_sum(y2);
                if (lane == 0) { s_part64[warp][0] = y0; s_part64[warp][1] = y1; s_part64[warp][2] = y2; }
                __syncthreads();
                if (tid == 0) {
                    y0 = y1 = y2 = 0.0;
#pragma unroll
                    for (int w = 0; w < kRrThreads / 32; ++w) { y0 += s_part64[w][0]; y1 += s_part64[w][1]; y2 += s_part64[w][2]; }
                    float acc;
                    if (COS) {
                        const double den = sqrt(y2) * sqrt(y1);
                        acc = 0.f - (den > 0.0 ? static_cast<float>(y0 / den) : 0.f);
                    } else {
                        const float d0 = hl ? static_cast<float>(sqrt(y0)) : 0.f;
                        const float d1 = static_cast<float>(sqrt(y1));
                        const float d2 = hr ? static_cast<float>(sqrt(y2)) : 0.f;
                        acc = __fmul_rn(p.lam1, d0);
                        acc = __fmaf_rn(p.lam2, d1, acc);
                        acc = __fmaf_rn(p.lam1, d2, acc);
                    }
                    const unsigned long long v = pack_score_idx(acc, static_cast<uint32_t>(p.offset + g));
                    loc = v < loc ? v : loc;
                    ++done;
                }
                __syncthreads();
            }
            __syncthreads();
            RR_MARK(c_p2);
        }
        if (tid == 0 && loc != ~0ull) atomicMin(p.best + row, loc);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (p.prof && tid == 0) {
        atomicAdd(&p.ctr->cyc_epi_busy, c_setup); atomicAdd(&p.ctr->cyc_epi_wait, c_sort);
        atomicAdd(&p.ctr->cyc_mma_wait_full, c_p1); atomicAdd(&p.ctr->cyc_mma_wait_acc, c_p2);
    }
#undef RR_MARK
    if (tid == 0) {
        if (appended) atomicAdd(&p.ctr->cand_count, appended);
        if (done) atomicAdd(&p.ctr->n_exact, done);
    }
    unsafe_n = static_cast<unsigned long long>(warp_sum(static_cast<double>(unsafe_n)));
    if (lane == 0 && unsafe_n) atomicAdd(&p.ctr->n_unsafe, unsafe_n);
}

// Rows whose candidates overflowed the list (or all rows, for eosvr_match_exact) are resolved by
// exhaustive exact evaluation.  Early exit when nothing overflowed.
__global__ void k_compact_flags(const RerankParams p)
{
    if (p.ctr->overflow == 0) return;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.P;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (p.rowflag[i]) {
            const unsigned int pos = atomicAdd(&p.ctr->n_flag_rows, 1u);
            p.flaglist[pos] = static_cast<int32_t>(i);
        }
    }
}

constexpr int kStrip = 8;    // gallery rows per warp work item in the exhaustive kernel

__global__ void k_exact_fallback(const RerankParams p)
{
    if (p.ctr->overflow == 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nrows = p.ctr->n_flag_rows;
    const int64_t nstrips = (p.G + kStrip - 1) / kStrip;
    for (; w < nrows * nstrips; w += nw) {
        const int64_t row = p.flaglist[w / nstrips];
        const int64_t g0 = (w % nstrips) * kStrip, g1 = min(g0 + kStrip, p.G);
        unsigned long long loc = ~0ull;
        for (int64_t g = g0; g < g1; ++g) {
            const float t = exact_score(p, row, g, lane);
            const unsigned long long v = pack_score_idx(t, static_cast<uint32_t>(p.offset + g));
            loc = v < loc ? v : loc;
        }
        if (lane == 0) atomicMin(p.best + row, loc);
    }
}

__global__ void k_finalize(const unsigned long long *__restrict__ best, int64_t P, int negate,
                           uint64_t *out_packed, float *out_score, int64_t *out_idx)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const unsigned long long v = best[i];
    if (out_packed) out_packed[i] = v;
    if (out_score) out_score[i] = (v == ~0ull) ? __int_as_float(0x7fc00000) : (negate ? -unpack_score(v) : unpack_score(v));
    if (out_idx) out_idx[i] = (v == ~0ull) ? -1 : static_cast<int64_t>(v & 0xFFFFFFFFull);
}

__global__ void k_merge_top1(const unsigned long long *__restrict__ gathered, int nshards, int64_t P,
                             uint64_t *out_packed, float *out_score, int64_t *out_idx)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= P) return;
    unsigned long long v = ~0ull;
    for (int s = 0; s < nshards; ++s) {
        const unsigned long long x = gathered[static_cast<int64_t>(s) * P + i];
        v = x < v ? x : v;
    }
    if (out_packed) out_packed[i] = v;
    if (out_score) out_score[i] = (v == ~0ull) ? __int_as_float(0x7fc00000) : unpack_score(v);
    if (out_idx) out_idx[i] = (v == ~0ull) ? -1 : static_cast<int64_t>(v & 0xFFFFFFFFull);
}

int launch_merge(const uint64_t *gathered, int32_t nshards, int64_t P, uint64_t *out_packed,
                 float *out_score, int64_t *out_idx, cudaStream_t st)
{
    if (P == 0) return EOSVR_OK;
    const int threads = 256;
    k_merge_top1<<<static_cast<unsigned>((P + threads - 1) / threads), threads, 0, st>>>(
        reinterpret_cast<const unsigned long long *>(gathered), nshards, P, out_packed, out_score, out_idx);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

// -------------------------------------------------------------------------------------------
// Host orchestration of one eosvr_match call (all asynchronous on `st`).
// -------------------------------------------------------------------------------------------
static int g_num_sms = 0;

static int g_max_clusters[3] = {0, 0, 0};   // co-resident clusters of k_match_screen<NP>, NP = 1, 2

template <int NP>
static int launch_screen_np(const CUtensorMap &tmA, const CUtensorMap &tmB, const ScreenParams &sp, cudaStream_t st)
{
    auto kern = k_match_screen<NP>;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * NP; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kScreenSmem;
    cfg.stream = st;
    if (g_max_clusters[NP] == 0) {
        EOSVR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kScreenSmem)));
        cfg.gridDim = dim3(static_cast<unsigned>(g_num_sms / (2 * NP) * (2 * NP)), 1, 1);
        int n = 0;
        EOSVR_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
        if (n < 1) { set_error("k_match_screen<%d>: no cluster of %d CTAs can be resident", NP, 2 * NP); return EOSVR_ECUDA; }
        g_max_clusters[NP] = n;
    }
    const int64_t ncl = sp.n_units < g_max_clusters[NP] ? sp.n_units : g_max_clusters[NP];
    cfg.gridDim = dim3(static_cast<unsigned>(ncl * 2 * NP), 1, 1);
    EOSVR_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, sp));
    return EOSVR_OK;
}

// pairs per cluster of the next screening launches.  Two pairs sharing (multicasting) the gallery slab halve its
// L2 reads but measured 3-6 % SLOWER on B200 (the kernel is not L2-bound), so the default is
// one pair; EOSVR_NP=2 selects the 2-pair cluster for experiments when there are at least two probe tiles.
static int choose_np(int64_t NT)
{
    static int np_env = -1;
    if (np_env < 0) { const char *e = getenv("EOSVR_NP"); np_env = e ? atoi(e) : 0; }
    return (np_env == 2 && NT >= 2) ? 2 : 1;
}

struct ScreenView {          // the screening copy of the gallery a launch reads (one per metric)
    const float *gnorm;
    const CUtensorMap *tmapA, *tmapSeed, *tmapAH, *tmapSeedH;
};

static int launch_screen(const eosvr_gallery *g, const ScreenView &sv, eosvr_workspace *ws, const MatchPlan &pl, int np,
                         bool seed, const CUtensorMap &tmB, int64_t gallery_tiles, int64_t g_stride, int64_t P,
                         bool timed, cudaStream_t st)
{
    ScreenParams sp;
    sp.gnorm = sv.gnorm; sp.G = g->G; sp.KB = g->Dp / kBK; sp.BN = pl.BN; sp.NT = pl.NT;
    sp.NTG = pl.NT / np;
    sp.GT = gallery_tiles;
    const int seed_mode = seed ? 1 : 0;
    const int64_t total_tiles = sp.NT * sp.GT;
    int64_t tpu = total_tiles / (static_cast<int64_t>(g_num_sms / 2) * 6);
    if (tpu < 1) tpu = 1;
    if (tpu > 16) tpu = 16;
    if (tpu > sp.GT) tpu = sp.GT;
    sp.TPU = static_cast<int32_t>(tpu);
    sp.n_chunks = (sp.GT + tpu - 1) / tpu;
    sp.n_units = sp.n_chunks * sp.NTG;
    sp.g_stride = g_stride; sp.seed_mode = seed_mode;
    {
        static int order_env = -1, tpu_env = -1;
        if (order_env < 0) { const char *e = getenv("EOSVR_ORDER"); order_env = e ? atoi(e) : 1; }
        if (tpu_env < 0) { const char *e = getenv("EOSVR_TPU"); tpu_env = e ? atoi(e) : 0; }
        sp.order = order_env;
        if (tpu_env > 0 && !seed_mode) {
            sp.TPU = static_cast<int32_t>(tpu_env < sp.GT ? tpu_env : sp.GT);
            sp.n_chunks = (sp.GT + sp.TPU - 1) / sp.TPU;
            sp.n_units = sp.n_chunks * sp.NTG;
        }
    }
    sp.na = ws->na; sp.wl = ws->wl; sp.wr = ws->wr; sp.margin = ws->margin; sp.rowmap = ws->rowmap;
    sp.gthr = ws->gthr; sp.cand = ws->cand; sp.rowcnt = ws->rowcnt; sp.cand_cap = static_cast<int32_t>(ws->cand_cap);
    sp.ovf = ws->ovf; sp.ovf_cap = static_cast<int32_t>(ws->ovf_cap);
    sp.ctr = ws->counters; sp.rowflag = ws->rowflag;
    sp.idesc = umma_idesc_f16(g->screen_fmt == EOSVR_SCREEN_F16 ? 0 : 1, kPairM, pl.BN);
    sp.dbg = (!seed_mode && ws->dbg && ws->dbg_elems >= P * g->G) ? ws->dbg : nullptr;
    {
        static int exp_env = -1;
        if (exp_env < 0) { const char *e = getenv("EOSVR_EXP"); exp_env = e ? atoi(e) : 0; }
        sp.exp_mode = exp_env;
    }
    const bool rec = timed && ws->timing_on && ws->timing_calls < kTimingRing;
    if (rec) EOSVR_CUDA(cudaEventRecord(ws->ev0[ws->timing_calls], st));
    int rc;
    if (np == 2) rc = launch_screen_np<2>(seed ? *sv.tmapSeedH : *sv.tmapAH, tmB, sp, st);
    else rc = launch_screen_np<1>(seed ? *sv.tmapSeed : *sv.tmapA, tmB, sp, st);
    if (rc) return rc;
    if (rec) { EOSVR_CUDA(cudaEventRecord(ws->ev1[ws->timing_calls], st)); ++ws->timing_calls; }
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

int launch_match(eosvr_gallery *g, eosvr_workspace *ws, const float *probes, int64_t P,
                 int32_t rpe, int32_t metric, float lam1, float lam2, bool exact_only, uint64_t *out_packed,
                 float *out_score, int64_t *out_idx, cudaStream_t st)
{
    if (P == 0) return EOSVR_OK;
    const bool cosm = metric == EOSVR_METRIC_COSINE;
    if (cosm) { lam1 = 0.f; lam2 = 1.f; rpe = 1; }          // no temporal taps: every probe row stands alone
    ScreenView sv{g->gnorm, &g->tmapA, &g->tmapSeed, &g->tmapAH, &g->tmapSeedH};
    const float *scalars = g->scalars;
    if (cosm && !exact_only) {
        int rc = build_cosine_copy(g, st);                  // first cosine call builds the normalised copy
        if (rc) return rc;
        EOSVR_CUDA(cudaStreamWaitEvent(st, g->cos->ready, 0));
        sv = ScreenView{g->cos->gnorm, &g->cos->tmapA, &g->cos->tmapSeed, &g->cos->tmapAH, &g->cos->tmapSeedH};
        scalars = g->cos->scalars;
    }
    if (g_num_sms == 0) {
        int dev = 0;
        EOSVR_CUDA(cudaGetDevice(&dev));
        EOSVR_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    MatchPlan pl = make_plan(P, rpe);
    const int np = exact_only ? 1 : choose_np(pl.NT);
    pl.NT = (pl.NT + np - 1) / np * np;           // pad with empty probe tiles: every pair of a cluster has one
    const int64_t ncol = pl.NT * pl.BN;
    if (ncol > ws->cap_rows) {
        set_error("workspace too small: plan needs %lld rows, capacity %lld", (long long)ncol, (long long)ws->cap_rows);
        return EOSVR_EINVAL;
    }
    PlanDev pd{pl.P, pl.rpe, pl.R, pl.halo, pl.BN, pl.NT};
    const int threads = 256;

    k_reset<<<static_cast<unsigned>((P + threads - 1) / threads), threads, 0, st>>>(
        ws->counters, ws->best, ws->rowflag, ws->rowcnt, ws->gthr, P, exact_only ? 1 : 0);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);

    RerankParams rp;
    rp.probes = probes; rp.gal = g->feats; rp.P = P; rp.G = g->G; rp.offset = g->offset;
    rp.D = g->D; rp.rpe = rpe; rp.metric = metric; rp.lam1 = lam1; rp.lam2 = lam2;
    rp.cand = ws->cand; rp.rowcnt = ws->rowcnt; rp.cand_cap = static_cast<int32_t>(ws->cand_cap);
    rp.ctr = ws->counters; rp.gthr = ws->gthr;
    rp.best = ws->best; rp.rowflag = ws->rowflag; rp.flaglist = ws->flaglist;
    rp.ovf = ws->ovf; rp.ovf_cap = static_cast<int32_t>(ws->ovf_cap);
    rp.margin = ws->margin; rp.planR = pl.R; rp.planBN = pl.BN; rp.planHalo = pl.halo;
    { const char *e = getenv("EOSVR_EXP"); rp.prof = (e && (atoi(e) & 64)) ? 1 : 0; }
    rp.rows_per_block = 1;

    ws->last_tiles = 0;
    ws->last_bn = pl.BN;
    if (!exact_only) {
        const unsigned pblocks = static_cast<unsigned>((ncol * 32 + threads - 1) / threads);
#define EOSVR_PROBE_PREP(T16, NORM)                                                                              \
        k_probe_prep<T16, NORM><<<pblocks, threads, 0, st>>>(probes, pd, g->D, g->Dp, scalars,                     \
            static_cast<T16 *>(ws->q16), ws->na, ws->epsd, ws->rowmap, ws->counters)
        if (g->screen_fmt == EOSVR_SCREEN_F16) { if (cosm) EOSVR_PROBE_PREP(__half, true); else EOSVR_PROBE_PREP(__half, false); }
        else { if (cosm) EOSVR_PROBE_PREP(__nv_bfloat16, true); else EOSVR_PROBE_PREP(__nv_bfloat16, false); }
#undef EOSVR_PROBE_PREP
        EOSVR_CUDA(cudaGetLastError());
        k_column_plan<<<static_cast<unsigned>((ncol + threads - 1) / threads), threads, 0, st>>>(
            pd, lam1 / lam2, ws->epsd, ws->wl, ws->wr, ws->margin);
        EOSVR_CUDA(cudaGetLastError());
        EOSVR_COUNT_LAUNCH(2);

        CUtensorMap tmB;
        int rc = encode_tmap_2d(&tmB, ws->q16, g->screen_fmt, static_cast<uint64_t>(ncol),
                                static_cast<uint64_t>(g->Dp), static_cast<uint32_t>(pl.BN / 2), kBK);
        if (rc) return rc;
        const int64_t GT = (g->G + kPairM - 1) / kPairM;   // 256-row tiles of the CTA pair
        // seed pass over a strided sample of the gallery: tightens every probe row's threshold before the
        // full pass so that concurrent CTAs do not flood the candidate lists
        static int seed_env = -1;                           // EOSVR_SEED=0 skips the seed pass (experiments)
        if (seed_env < 0) { const char *e = getenv("EOSVR_SEED"); seed_env = e ? atoi(e) : 1; }
        if (seed_env && g->seed_tiles > 0 && GT > g->seed_tiles) {
            rc = launch_screen(g, sv, ws, pl, np, true, tmB, seed_env == 1 ? g->seed_tiles : 1, g->seed_stride, P, false, st);
            if (rc) return rc;
        }
        rc = launch_screen(g, sv, ws, pl, np, false, tmB, GT, 1, P, true, st);
        if (rc) return rc;
        ws->last_tiles = pl.NT * GT;

        // few probe rows (a single episode): one row per block keeps the call's latency low; large batches take 8
        // consecutive rows per block so that neighbouring probe rows are staged once
        int64_t rpb = P / (static_cast<int64_t>(g_num_sms) * 8);
        rpb = rpb < 1 ? 1 : (rpb > kRrRowsPerBlock ? kRrRowsPerBlock : rpb);
        rp.rows_per_block = static_cast<int32_t>(rpb);
        const int64_t rr_blocks = (P + rpb - 1) / rpb;
        const unsigned rr_grid = static_cast<unsigned>(rr_blocks < static_cast<int64_t>(g_num_sms) * 32 ? rr_blocks : g_num_sms * 32);
        const size_t rr_smem = static_cast<size_t>(cosm ? 2 : 4) * g->D * sizeof(float);
        if ((g->D & 3) == 0 && rr_smem <= 96 * 1024) {
            static bool rr_attr = false;
            if (!rr_attr) {
                EOSVR_CUDA(cudaFuncSetAttribute(k_rerank_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                EOSVR_CUDA(cudaFuncSetAttribute(k_rerank_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                rr_attr = true;
            }
            if (cosm) k_rerank_rows<true><<<rr_grid, kRrThreads, rr_smem, st>>>(rp);
            else k_rerank_rows<false><<<rr_grid, kRrThreads, rr_smem, st>>>(rp);
        }
        else k_rerank<<<g_num_sms * 8, 256, 0, st>>>(rp);
        EOSVR_CUDA(cudaGetLastError());
        k_rerank_ovf<<<g_num_sms * 4, 256, 0, st>>>(rp);
        EOSVR_CUDA(cudaGetLastError());
        EOSVR_COUNT_LAUNCH(2);
    }
    k_compact_flags<<<64, 256, 0, st>>>(rp);
    EOSVR_CUDA(cudaGetLastError());
    k_exact_fallback<<<g_num_sms * 8, 256, 0, st>>>(rp);
    EOSVR_CUDA(cudaGetLastError());
    k_finalize<<<static_cast<unsigned>((P + threads - 1) / threads), threads, 0, st>>>(
        ws->best, P, cosm ? 1 : 0, out_packed, out_score, out_idx);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(3);
    return EOSVR_OK;
}

// -------------------------------------------------------------------------------------------
// Winner rows (multi-GPU exchange helper and input of the splice kernel).
// -------------------------------------------------------------------------------------------
__global__ void k_gather_rows(const float *__restrict__ gal, int64_t G, int64_t offset, int D,
                              const int64_t *__restrict__ idx, float *__restrict__ out)
{
    const int64_t p = blockIdx.x;
    const int64_t g = idx[p] - offset;
    const bool own = g >= 0 && g < G;
    const float *src = gal + (own ? g : 0) * D;
    float *dst = out + p * D;
    if ((D & 3) == 0) {
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int k = threadIdx.x; k < D / 4; k += blockDim.x) d4[k] = own ? s4[k] : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int k = threadIdx.x; k < D; k += blockDim.x) dst[k] = own ? src[k] : 0.f;
    }
}

int launch_gather_rows(const eosvr_gallery *g, const int64_t *idx, int64_t P, float *out, cudaStream_t st)
{
    if (P == 0) return EOSVR_OK;
    k_gather_rows<<<static_cast<unsigned>(P), 128, 0, st>>>(g->feats, g->G, g->offset, g->D, idx, out);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

}  // namespace eosvr
