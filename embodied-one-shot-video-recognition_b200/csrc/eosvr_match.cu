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

#include <mutex>

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
__global__ void k_gallery_prep(const void *__restrict__ feats, int src_dtype, int64_t G, int64_t Gpad, int D, int Dp,
                               T16 *__restrict__ h16, int write_h16, float *__restrict__ gnorm, float *__restrict__ scalars)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= Gpad) return;
    T16 *dst = h16 + row * Dp;
    if (row >= G) {
        if (write_h16) for (int k = lane; k < Dp; k += 32) dst[k] = to16<T16>(0.f);
        if (lane == 0) gnorm[row] = kPadNorm;
        return;
    }
    const int64_t src0 = row * D;
    double rn = 1.0; float rnf = 1.f;
    if (NORM) {
        double q = 0.0;
        for (int k = lane; k < D; k += 32) { const double x = ld_feat(feats, src_dtype, src0 + k); q += x * x; }
        q = warp_sum(q);
        rn = q > 0.0 ? 1.0 / sqrt(q) : 0.0;
        rnf = static_cast<float>(rn);
    }
    double s = 0.0, sh = 0.0, sl = 0.0;
    for (int k = lane; k < Dp; k += 32) {
        float x = k < D ? ld_feat(feats, src_dtype, src0 + k) : 0.f;
        T16 h = to16<T16>(NORM ? __fmul_rn(x, rnf) : x);
        if (write_h16) dst[k] = h;
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
static int launch_gallery_prep_t(const eosvr_gallery *g, void *h16, int write_h16, float *gnorm, float *scalars, cudaStream_t st)
{
    const int64_t Gpad = (g->G + kPairM - 1) / kPairM * kPairM;
    EOSVR_CUDA(cudaMemsetAsync(scalars, 0, 4 * sizeof(float), st));
    const int threads = 256;
    const int64_t blocks = (Gpad * 32 + threads - 1) / threads;
    if (g->screen_fmt == EOSVR_SCREEN_F16)
        k_gallery_prep<__half, NORM><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
            g->feats, g->dtype, g->G, Gpad, g->D, g->Dp, static_cast<__half *>(h16), write_h16, gnorm, scalars);
    else
        k_gallery_prep<__nv_bfloat16, NORM><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
            g->feats, g->dtype, g->G, Gpad, g->D, g->Dp, static_cast<__nv_bfloat16 *>(h16), write_h16, gnorm, scalars);
    EOSVR_CUDA(cudaGetLastError());
    return EOSVR_OK;
}

int launch_gallery_prep(eosvr_gallery *g, cudaStream_t st)
{
    return launch_gallery_prep_t<false>(g, g->h16, g->owns_h16, g->gnorm, g->scalars, st);
}

int launch_gallery_prep_cos(const eosvr_gallery *g, eosvr_screen_copy *c, cudaStream_t st)
{
    return launch_gallery_prep_t<true>(g, c->h16, 1, c->gnorm, c->scalars, st);
}

// -------------------------------------------------------------------------------------------
// Probe plan.  Probe rows are laid out as NT tiles of BN columns; tile t emits probe rows
// [t*R, (t+1)*R).  When an episode fits a tile (rpe <= 256) R is a whole number of episodes and
// no halo is needed; otherwise each tile carries one halo column on each side so the taps of its
// first/last emitted row see their neighbours.
// -------------------------------------------------------------------------------------------
static bool plan_bn3_enabled()
{
    static const bool on = [] { const char *e = getenv("EOSVR_BN3"); return e ? atoi(e) != 0 : false; }();   // experiments: 1 = 160-column tiles, three accumulators
    return on;
}

MatchPlan make_plan(int64_t P, int32_t rpe)
{
    MatchPlan pl;
    pl.P = P;
    pl.rpe = rpe;
    if (rpe <= kMaxBN) {
        // Experiment (EOSVR_BN3=1, off by default): 20-row episodes in tiles of 8 episodes = 160 columns, so that THREE
        // accumulators fit the 512 TMEM columns.  Measured 10-20 % SLOWER than 240-column tiles with two accumulators
        // (profiles/r02_ab_three_accumulators_rejected.txt): the single TMA producer thread and the per-tile fixed cost
        // of the epilogue outweigh the deeper accumulator ring.
        const int cap = (rpe == kAlign && plan_bn3_enabled()) ? kBN3 : kMaxBN;
        int64_t k = cap / rpe;
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

// Threshold margin of plan column i: twice the one-sided error bound of its smoothed screening value
// (DESIGN.md "Error bound"); epsd / wl / wr are the per-column arrays k_probe_prep writes.
__device__ __forceinline__ float column_margin(const float *__restrict__ epsd, const float *__restrict__ wl,
                                               const float *__restrict__ wr, int64_t i)
{
    const float l = wl[i], r = wr[i];
    const float el = l > 0.f ? epsd[i - 1] : 0.f, er = r > 0.f ? epsd[i + 1] : 0.f;
    return 2.02f * (epsd[i] + l * el + r * er) + 1e-7f;
}

// per-probe-row state of one eosvr_match call
struct RowState {
    unsigned long long *best;
    int32_t *rowflag;
    unsigned int *rowcnt;
    unsigned int *gthr;
};

// One warp per plan column: convert the probe row to the 16-bit screening format, compute the squared norm, the
// per-column error bound E2 (see DESIGN.md "Error bound") and the tap weights (w = lam1/lam2 towards a neighbour of
// the same episode, 0 at episode ends), and reset the state of the probe row the column emits.
template <typename T16, bool NORM>
__global__ void k_probe_prep(const float *__restrict__ probes, PlanDev pl, int D, int Dp, float w,
                             const float *__restrict__ gscal, T16 *__restrict__ q16,
                             float *__restrict__ na, float *__restrict__ epsd, float *__restrict__ wl,
                             float *__restrict__ wr, int32_t *__restrict__ rowmap, RowState rs, Counters *ctr)
{
    const int lane = threadIdx.x & 31;
    const int64_t col = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (col >= pl.NT * pl.BN) return;
    bool valid, emit;
    const int64_t p = plan_row_of(pl, col, valid, emit);
    T16 *dst = q16 + col * Dp;
    if (!valid) {
        for (int k = lane; k < Dp; k += 32) dst[k] = to16<T16>(0.f);
        if (lane == 0) { na[col] = kPadNorm; epsd[col] = 0.f; wl[col] = 0.f; wr[col] = 0.f; rowmap[col] = -1; }
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
        // |x~ - x| <= 2(|a_lo||b| + |a_hi||b_lo|) + tensor-core accumulation + fp32 rounding.  The accumulator starts
        // at -|b|^2/2 (the epilogue's fill value), so the partial sums are bounded by |a_hi||b_hi| + |b|^2/2.
        double e2 = 2.0 * (sqrt(sl * B2) + sqrt(sh * Bl2))
                  + 2.0 * ulp * (Dp / 16 + 1) * (sqrt(sh * Bh2) + 0.5 * B2)
                  + 2.0 * ulp * (s + B2);
        e2 *= 1.01;
        if (e2 < 1e-30) e2 = 1e-30;
        na[col] = static_cast<float>(s);
        epsd[col] = __double2float_ru(sqrt(e2) / 16.0);           // E2 / (2 sqrt(64 E2))
        rowmap[col] = emit ? static_cast<int32_t>(p) : -1;
        atomicMax(&ctr->xfloor_bits, __float_as_uint(__double2float_ru(65.0 * e2)));
        const int c = static_cast<int>(col % pl.BN);
        float l = 0.f, r = 0.f;
        if (c > 0 && p > 0 && (p % pl.rpe) != 0) l = w;
        if (c + 1 < pl.BN && p + 1 < pl.P && ((p + 1) % pl.rpe) != 0) {
            bool v2, e2n;
            plan_row_of(pl, col + 1, v2, e2n);
            if (v2) r = w;
        }
        wl[col] = l; wr[col] = r;
        if (emit) { rs.best[p] = ~0ull; rs.rowflag[p] = 0; rs.rowcnt[p] = 0; rs.gthr[p] = 0x7f800000u; }
    }
}

// eosvr_match_exact only (no screening pass): every row goes to the exhaustive kernel.
__global__ void k_reset_exact(Counters *ctr, RowState rs, int64_t P)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i == 0) ctr->overflow = 2u;                      // 2: all rows
    if (i < P) { rs.best[i] = ~0ull; rs.rowflag[i] = 1; rs.rowcnt[i] = 0; rs.gthr[i] = 0x7f800000u; }
}

// -------------------------------------------------------------------------------------------
// Screening kernel
// -------------------------------------------------------------------------------------------
struct ScreenParams {
    const float *gnorm;
    int64_t G;
    int32_t KB;             // K blocks of 64 elements
    int32_t BN;
    int32_t NT, GT;         // probe tiles, gallery tiles (32-bit: unit arithmetic stays cheap in the kernel)
    int32_t TPU;            // gallery tiles per work unit
    int32_t n_chunks;       // gallery chunks (of TPU tiles)
    int32_t n_units;        // n_chunks * NT
    int32_t order;          // unit order: 0 chunk-major, 1 probe-tile-major, 2 diagonal (rotated chunks)
    int64_t g_stride;       // gallery row stride (1; > 1 in the seed pass)
    int32_t seed_mode;      // 1: only tighten the thresholds, append nothing
    int32_t issuers;        // MMA issuer warps in use: kIssuers, or 1 (the documented single-issuer ordering)
    int32_t acc_stages;     // TMEM accumulator stages: 3 when the tile has <= 160 columns, else 2
    int32_t acc_stride;     // TMEM columns between accumulator stages
    int32_t rpe;            // probe rows per episode (the episode-aligned epilogue needs rpe == its chunk width)
    float w;                // tap weight lam1 / lam2 towards a neighbour of the same episode
    const float *na, *wl, *wr, *epsd;
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
    int32_t exp_mode;       // measurement modes (EOSVR_EXP): 16 = cycle accounting.  Only in builds with
                            // -DEOSVR_EXPERIMENTS (timing only, WRONG results): 1 = the epilogue releases the accumulator
                            // untouched, 2 = producers skip the TMA loads, 4 = issuers skip the MMAs, 32 = no rare path,
                            // 128 = no accumulator refill, 256 = no lane-quadrant barrier, 512 = no square roots
};

#ifdef EOSVR_EXPERIMENTS
#define EOSVR_EXP_ON(p, bit) (((p).exp_mode & (bit)) != 0)
#else
#define EOSVR_EXP_ON(p, bit) false
#endif

struct StagedCand {
    int32_t rm;         // probe row
    int32_t g;          // gallery row (local to the shard)
    uint32_t tbits;
};
constexpr int kWarpStage = 64;
constexpr int kStashCols = 6;      // parked columns per epilogue warp (a full stash is worked off on the spot)

template <int EW>
struct __align__(16) ScreenSmemTail {
    float na[kMaxBN];
    float wl[kMaxBN];
    float wr[kMaxBN];
    float mg[kMaxBN];
    unsigned int thr[kMaxBN];
    int32_t row[kMaxBN];
    uint64_t full[kStages];           // stage s is always consumed by issuer s % issuers: a parity wait is only safe
                                      // on a barrier whose every phase the waiter observes
    uint64_t empty[kStages];
    uint64_t tfull[kMaxAccStages];
    uint64_t tempty[kMaxAccStages];
    uint32_t tmem_base;
    uint32_t zeros[8];                // zeros the compiler cannot see through (refill-store source registers)
    // candidate staging: every epilogue warp parks its candidates in shared memory and hands them to the
    // per-row lists 32 at a time (one global atomicAdd per lane, all in flight together) instead of stalling on
    // one atomic round trip per column
    StagedCand stage[EW][kWarpStage];
    // episode-aligned epilogue: columns in which some lane fell below its threshold are parked here (the 32 lanes'
    // values of the column) and worked off AFTER the accumulator has been handed back to the MMA issuers
    uint32_t sval[EW][kStashCols][32];
    int32_t scol[EW][kStashCols];       // column within the tile
    int32_t sgt[EW][kStashCols];        // gallery tile of the entry
};

constexpr int kABytes = kBM * kBK * 2;              // 16 KiB: this CTA's 128 gallery rows
constexpr int kBBytes = (kMaxBN / 2) * kBK * 2;     // 16 KiB: this CTA's half of the probe tile
constexpr int kAStage = kSub * kABytes;             // one pipeline stage = kSub K blocks of each operand
constexpr int kBStage = kSub * kBBytes;
template <int EW>
constexpr size_t screen_smem() { return static_cast<size_t>(kStages) * (kAStage + kBStage) + sizeof(ScreenSmemTail<EW>); }
constexpr int kEpiRegs = 104;     // registers of an epilogue thread when 16 epilogue warps run (setmaxnreg) ...
constexpr int kLightRegs = 64;    // ... and of the producer / issuer threads: 64 + 4 x 104 = 5 x 96 per scheduler

struct UnitIter {
    int32_t u, jt, gt0, gt1;    // jt: probe tile of the unit
};

// Work units (gallery chunk x probe tile), three orders (EOSVR_ORDER):
//   1 probe-tile-major: unit u = probe_tile * n_chunks + chunk.  The pairs that run at the same time work
//     on the same few probe tiles against different gallery chunks, so a probe tile's K blocks are fetched from HBM
//     once (L2 hits for the other pairs) and the whole 16-bit gallery stays L2-resident: DRAM traffic close to the
//     algorithmic bytes.  ~10 % more candidates than chunk-major (several pairs screen the same probe rows at once
//     and see each other's thresholds one tile late) and still 1-2 % faster;
//   0 chunk-major: unit u = chunk * NT + probe_tile.  Tightest thresholds, but every probe tile is re-read from HBM
//     once per gallery chunk (measured 4.7x the algorithmic bytes at the bench size);
//   2 diagonal (rotated chunks).
__device__ __forceinline__ bool decode_unit(const ScreenParams &p, int32_t u, UnitIter &it)
{
    int32_t chunk;
    it.u = u;
    if (p.order == 1) {
        it.jt = u / p.n_chunks;
        chunk = u % p.n_chunks;
    } else {
        chunk = u / p.NT;
        it.jt = u % p.NT;
        if (p.order == 2) chunk = (chunk + it.jt) % p.n_chunks;   // every (tile, chunk) still visited once
    }
    it.gt0 = chunk * p.TPU;
    it.gt1 = min(it.gt0 + p.TPU, p.GT);
    return it.gt0 < it.gt1;
}

// The sequence of (unit, gallery tile) a CTA pair works through; the epilogue keeps a second cursor two tiles
// ahead (the next user of the accumulator it is reading).
struct TileCursor {
    int32_t u;
    UnitIter it;
    int32_t gt;
    bool valid;
};
__device__ __forceinline__ void cursor_seek(const ScreenParams &p, TileCursor &c, int32_t stride)
{
    c.valid = false;
    for (; c.u < p.n_units; c.u += stride)
        if (decode_unit(p, c.u, c.it)) { c.gt = c.it.gt0; c.valid = true; return; }
}
__device__ __forceinline__ void cursor_next(const ScreenParams &p, TileCursor &c, int32_t stride)
{
    if (!c.valid) return;
    if (++c.gt < c.it.gt1) return;
    c.u += stride;
    cursor_seek(p, c, stride);
}

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

__device__ __forceinline__ float pick20(const float (&t)[kAlign], int j)
{
    switch (j) {
        case 0: return t[0];   case 1: return t[1];   case 2: return t[2];   case 3: return t[3];
        case 4: return t[4];   case 5: return t[5];   case 6: return t[6];   case 7: return t[7];
        case 8: return t[8];   case 9: return t[9];   case 10: return t[10]; case 11: return t[11];
        case 12: return t[12]; case 13: return t[13]; case 14: return t[14]; case 15: return t[15];
        case 16: return t[16]; case 17: return t[17]; case 18: return t[18]; default: return t[19];
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

// EW = epilogue warps (8 or 16).  DIAG = diagnostic build of the kernel: cycle accounting (EOSVR_EXP bit 16) and the
// screening-value dump (eosvr_workspace_set_debug, used by the tests and by the per-device self-check); the
// production instantiation carries neither (they cost ~25 predicated-off instructions per chunk).
//
// The accumulator never starts from zero: the epilogue leaves -|b|^2/2 of the gallery row that will occupy the
// lane NEXT in every column it has finished with (tcgen05.st, one chunk behind its reads), and every MMA
// accumulates, so the tensor core delivers a.b - |b|^2/2 and the squared distance is one FFMA per element:
// x = fma(acc, -2, |a|^2).  The three issuers therefore never order themselves against each other either.
//
// AL > 0 selects the EPISODE-ALIGNED epilogue (requires rows-per-episode == AL, no halo, DIAG off): a chunk is the AL
// columns of one whole episode, so no tap ever crosses a chunk -- no boundary reads between column groups, no lane-
// quadrant barrier, no per-column tap weights (w inside the episode, nothing at its ends), 240 columns split evenly
// as 3 episodes per column group.  Columns in which some lane falls below its threshold are only PARKED in shared
// memory inside the chunk loop (the column's 32 values); tightening the threshold and appending the candidates
// happens after the accumulator has gone back to the MMA issuers, in time the warp would otherwise spend waiting for
// the next accumulator -- the tile's release no longer waits for the slowest warp's candidate handling.
template <int EW, bool DIAG, int AL>
__global__ void __launch_bounds__(screen_threads(EW), 1)
k_match_screen(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ScreenParams p)
{
    // 1024-byte alignment (SWIZZLE_128B operand tiles) comes from the declaration, so that every pointer below stays
    // in the shared address space for the compiler (an integer round trip turns the epilogue's loads into generic
    // LD.E: +20 % epilogue time, tools/bench_micro/epi_rate.cu); checked once below
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;
    uint8_t *sB = smem + kStages * kAStage;
    using Tail = ScreenSmemTail<EW>;
    Tail *tl = reinterpret_cast<Tail *>(smem + kStages * (kAStage + kBStage));
    const int KS = (p.KB + kSub - 1) / kSub;            // pipeline stages per gallery tile
    constexpr bool kTwoProducers = two_producers(EW);
    constexpr int kProdBWarp = 4 + EW;                  // second TMA producer (probe operand); warp 0 loads the gallery
    constexpr int EG = EW / 4;                          // column groups of the epilogue

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = crank & 1u;                   // 0 = leader of its pair (issues the MMAs), 1 = peer
    const int32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;   // cluster (CTA pair) index / count

    if (warp == 0 && lane == 0) {
        if ((smem_u32(smem) & 1023u) != 0u) __trap();
        for (int i = 0; i < 8; ++i) tl->zeros[i] = 0u;
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < kStages; ++s) { mbar_init(&tl->full[s], 1); mbar_init(&tl->empty[s], 1); }
        for (int s = 0; s < kMaxAccStages; ++s) {
            mbar_init(&tl->tfull[s], p.issuers);         // one commit per MMA issuer warp
            mbar_init(&tl->tempty[s], 2 * EW);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_2sm(&tl->tmem_base, kTmemCols);
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = tl->tmem_base;
    const bool prof = DIAG && (p.exp_mode & 16) != 0;
    const bool epi_warp = warp >= 4 && warp < 4 + EW;
    // EW = 16: 20 warps share 64 K registers; every role branch below starts by handing back (setmaxnreg.dec) or
    // claiming (setmaxnreg.inc, epilogue) registers, and the branches only meet again at the final cluster barrier
    constexpr bool kRepartition = EW > 8;

    // Producer and issuer warps run their loops with the whole warp; lane 0 issues.
    if (warp == 0 || (kTwoProducers && warp == kProdBWarp)) {
        // ===== TMA producers (every CTA): warp 0 loads the CTA's gallery rows (A), warp 4+EW its half of the pair's
        //       probe tile (B) -- a UTMALDG occupies its issuing thread for ~150 cycles, so one thread per operand.
        //       Warp 0 also posts the byte count of the whole stage. =====
        if (kRepartition) reg_release<kLightRegs>();
        const bool isA = warp == 0, isB = kTwoProducers ? warp == kProdBWarp : true;   // one warp may play both roles
        int stage = 0; uint32_t phase = 0;
        const int32_t bhalf = p.BN >> 1;
        const uint32_t tx_cta = kABytes + static_cast<uint32_t>(bhalf) * kBK * 2;
        unsigned long long w_prod = 0;
        const long long t_begin = clock64();
        const uint32_t lbar0 = mapa(smem_u32(&tl->full[0]), 0);
        for (int32_t u = pair; u < p.n_units; u += npairs) {
            UnitIter it;
            if (!decode_unit(p, u, it)) continue;
            const int32_t brow = static_cast<int32_t>(it.jt * p.BN + rank * bhalf);
            for (int32_t gt = it.gt0; gt < it.gt1; ++gt) {
                const int32_t arow = static_cast<int32_t>(gt * kPairM + rank * kBM);
                for (int ks = 0; ks < KS; ++ks) {
                    const int nsub = min(kSub, p.KB - ks * kSub);
                    if (prof) { const long long t0 = clock64(); mbar_wait(&tl->empty[stage], phase ^ 1); w_prod += clock64() - t0; }
                    else mbar_wait(&tl->empty[stage], phase ^ 1);
                    if (lane == 0) {
                        if (EOSVR_EXP_ON(p, 2)) {
                            if (isA && rank == 0) mbar_arrive(&tl->full[stage]);
                        } else {
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
    } else if (warp >= 1 && warp <= p.issuers && rank == 0) {
        // ===== MMA issuers: p.issuers warps of the pair's leader drive the tensor cores of both SMs, taking the
        //       pipeline stages round-robin (stage s -> issuer s % issuers) into the same accumulator.  (One commit per
        //       PAIR of stages -- fewer of the ~900-cycle commits per issuing thread -- measured 1-2 % slower on every
        //       shape, profiles/r02_ab_three_accumulators_rejected.txt.)  Every MMA
        //       accumulates onto the values the epilogue left behind, so the issuers need no ordering among
        //       themselves.  Every issuer observes every phase of tempty. =====
        if (kRepartition) reg_release<kLightRegs>();
        const int w = warp - 1;
        const int nis = p.issuers;
        int64_t seq = 0;                                  // running stage number of this pair
        int acc = 0; uint32_t accphase = 0;
        const uint16_t pair_mask = 3u;
        unsigned long long w_full = 0, w_acc = 0;
        for (int32_t u = pair; u < p.n_units; u += npairs) {
            UnitIter it;
            if (!decode_unit(p, u, it)) continue;
            for (int32_t gt = it.gt0; gt < it.gt1; ++gt) {
                // phase 0 of tempty is the epilogue's initial fill, phase k+1 the release after the k-th use
                if (prof) { const long long t0 = clock64(); mbar_wait(&tl->tempty[acc], accphase); w_acc += clock64() - t0; }
                else mbar_wait(&tl->tempty[acc], accphase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride);
                for (int ks = 0; ks < KS; ++ks, ++seq) {
                    if (static_cast<int>(seq % nis) != w) continue;
                    const int stage = static_cast<int>(seq % kStages);
                    const uint32_t phase = static_cast<uint32_t>(seq / kStages) & 1u;
                    const int nsub = min(kSub, p.KB - ks * kSub);
                    if (prof) { const long long t0 = clock64(); mbar_wait(&tl->full[stage], phase); w_full += clock64() - t0; }
                    else mbar_wait(&tl->full[stage], phase);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + stage * kAStage);
                    const uint32_t b0 = smem_u32(sB + stage * kBStage);
                    if (lane == 0) {
                        if (!EOSVR_EXP_ON(p, 4)) {
                            for (int sb = 0; sb < nsub; ++sb) {
#pragma unroll
                                for (int k = 0; k < kBK / 16; ++k)
                                    mma_f16_ss_2sm(d_tmem, umma_desc_sw128(a0 + sb * kABytes, k * 32),
                                                   umma_desc_sw128(b0 + sb * kBBytes, k * 32), p.idesc, 1u);
                            }
                        }
                        mma_commit_2sm(&tl->empty[stage], pair_mask);   // frees the stage in both CTAs
                    }
                    __syncwarp();
                }
                if (lane == 0) mma_commit_2sm(&tl->tfull[acc], pair_mask);   // this warp's share of the tile is done
                __syncwarp();
                if (++acc == p.acc_stages) { acc = 0; accphase ^= 1; }
            }
        }
        if (prof && lane == 0 && w == 0) atomicAdd(&p.ctr->cyc_mma_wait_acc, w_acc);
        if (prof && lane == 0) atomicAdd(&p.ctr->cyc_mma_wait_full, w_full);
    } else if (epi_warp) {
        // ===== epilogue: warp%4 selects the TMEM lane quadrant, (warp-4)/4 the group of column chunks =====
        if (kRepartition) reg_acquire<kEpiRegs>();
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int te = threadIdx.x - 128;
        const int BN = p.BN;
        const int nchunks = BN / kChunk;
        const int cbeg = nchunks * grp / EG;
        const int cend = nchunks * (grp + 1) / EG;
        const float xfloor = __uint_as_float(p.ctr->xfloor_bits);
        const float dfloor = sqrtf(xfloor) * 1.000001f;
        const float dfloor_hard = dfloor * 0.24806947f;      // sqrt(4/65): x~ < 4 E2
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        const int64_t row_in_tile = static_cast<int64_t>(rank) * kBM + q * 32 + lane;
        int acc = 0; uint32_t accphase = 0;
        const uint32_t tempty_leader0 = mapa(smem_u32(&tl->tempty[0]), 0);
        unsigned long long e_busy = 0, e_wait = 0, e_pre = 0, e_loop = 0;

        // accumulator fill value for tile c: -|b|^2/2 of this lane's gallery row (0 when there is no such tile)
        auto fill_bits = [&](const TileCursor &c) -> uint32_t {
            if (!c.valid) return 0u;
            return __float_as_uint(-0.5f * p.gnorm[(static_cast<int64_t>(c.gt) * kPairM + row_in_tile) * p.g_stride]);
        };
        TileCursor cur;
        cur.u = pair;
        cursor_seek(p, cur, npairs);
        TileCursor ahead = cur;
        {
            // initial fill: accumulator s for the pair's tile s; every column group takes its share of the 512 columns
            // and every warp then reports all accumulators (a barrier completes only when all 2*EW warps of the pair
            // have arrived)
            const uint32_t f0 = fill_bits(ahead);
            cursor_next(p, ahead, npairs);
            const uint32_t f1 = fill_bits(ahead);
            cursor_next(p, ahead, npairs);
            uint32_t f2 = 0u;
            if (p.acc_stages > 2) { f2 = fill_bits(ahead); cursor_next(p, ahead, npairs); }
            constexpr int kShare = kTmemCols / EG;
#pragma unroll 1
            for (int c = grp * kShare; c < (grp + 1) * kShare; c += kChunk) {
                const int sidx = c / p.acc_stride;
                tmem_st_fill_x16(tmem_base + lane_off + c, sidx == 0 ? f0 : (sidx == 1 ? f1 : f2));
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                for (int sidx = 0; sidx < p.acc_stages; ++sidx) mbar_arrive_cluster(tempty_leader0 + sidx * 8u);
        }
        // from here on `ahead` is the next tile to use the accumulator `cur` is in (acc_stages tiles later)
        StagedCand *wstage = tl->stage[warp - 4];
        int wn = 0;                                       // candidates parked by this warp (warp-uniform)
        int ns = 0;                                       // columns parked by the aligned epilogue (warp-uniform)
        // Work off the parked columns: per column tighten the threshold with the warp minimum, then stage what is still
        // below it for the row's candidate list (the generic epilogue's per-column step, off the tile's critical path).
        auto drain_stash = [&]() {
            __syncwarp();
            for (int k = 0; k < ns; ++k) {
                const int c = tl->scol[warp - 4][k];
                const int32_t rm = tl->row[c];
                if (rm < 0) continue;                             // warp-uniform
                const uint32_t tb = tl->sval[warp - 4][k][lane];
                const bool uns = tb == kCandUnsafe;
                const float tj = __uint_as_float(tb);
                float thr = __uint_as_float(*reinterpret_cast<volatile unsigned int *>(&tl->thr[c]));
                bool pass = !uns && (tj <= thr);                  // (lanes that were not below the threshold hold +inf)
                if (__ballot_sync(0xffffffffu, pass)) {
                    const float mn = __uint_as_float(__reduce_min_sync(0xffffffffu, pass ? tb : 0x7f800000u));
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
                const bool app = pass || uns;
                const unsigned ma = __ballot_sync(0xffffffffu, app);
                if (ma) {
                    if (app) {
                        StagedCand sc;
                        sc.rm = rm;
                        sc.g = static_cast<int32_t>((static_cast<int64_t>(tl->sgt[warp - 4][k]) * kPairM + row_in_tile) * p.g_stride);
                        sc.tbits = tb;
                        wstage[wn + __popc(ma & ((1u << lane) - 1u))] = sc;
                    }
                    wn += __popc(ma);
                    if (wn >= 32) { flush_staged(p, wstage, wn, lane); wn = 0; }
                }
            }
            __syncwarp();
        };
        int32_t smem_unit = -1;
        while (cur.valid) {
            if (cur.it.u != smem_unit) {
                // per-unit column arrays -> smem
                smem_unit = cur.it.u;
                named_bar_sync(1, 32 * EW);
                if (te < BN) {
                    const int64_t c = static_cast<int64_t>(cur.it.jt) * BN + te;
                    tl->na[te] = p.na[c];
                    tl->wl[te] = p.wl[c];
                    tl->wr[te] = p.wr[c];
                    tl->mg[te] = column_margin(p.epsd, p.wl, p.wr, c);
                    const int32_t rm = p.rowmap[c];
                    tl->row[te] = rm;
                    tl->thr[te] = rm >= 0 ? ld_volatile_u32(p.gthr + rm) : __float_as_uint(-1.0f);
                }
                named_bar_sync(1, 32 * EW);
            }
            // thresholds tightened by other CTAs since the last tile (benign race): requested now, folded in after
            // this tile's chunks, so the load's latency hides behind the tile instead of stalling the warp
            unsigned int pend_thr = 0xFFFFFFFFu;
            if (te < BN) {
                const int32_t rm = tl->row[te];
                if (rm >= 0) pend_thr = ld_volatile_u32(p.gthr + rm);
            }
            // this lane's gallery row; the fill value of the accumulator's next tile is requested BEFORE waiting for
            // the accumulator so the global-load latency hides behind the wait
            const int64_t g = (static_cast<int64_t>(cur.gt) * kPairM + row_in_tile) * p.g_stride;
            const uint32_t fill = fill_bits(ahead);
            long long t_e0 = 0;
            if (prof) { const long long t0 = clock64(); mbar_wait(&tl->tfull[acc], accphase); t_e0 = clock64(); e_wait += t_e0 - t0; }
            else mbar_wait(&tl->tfull[acc], accphase);
            tc_fence_after();
            const uint32_t trow = tmem_base + static_cast<uint32_t>(acc * p.acc_stride) + lane_off;
            bool released = false;                        // aligned epilogue: the accumulator was handed back inside the tile

            if constexpr (AL > 0) {
                // ===== episode-aligned tile =====
                const int nep = BN / AL;                              // whole episodes in the tile (padding columns beyond)
                const int ebeg = nep * grp / EG, eend = nep * (grp + 1) / EG;
                if (!EOSVR_EXP_ON(p, 1) && eend > ebeg) {
                    const bool rowok = g < p.G;
                    const float w = p.w;
                    uint32_t fv[8];
                    {
                        const volatile unsigned int *zp = tl->zeros;
#pragma unroll
                        for (int i = 0; i < 8; ++i) fv[i] = fill | zp[i];
                    }
                    uint32_t v[AL];
                    tmem_ld_x20(trow + ebeg * AL, v);
                    tmem_ld_wait_x20(v);
                    long long t_e1 = 0;
                    if (prof) { t_e1 = clock64(); e_pre += t_e1 - t_e0; }
#pragma unroll 1
                    for (int e = ebeg; e < eend; ++e) {
                        const int c0 = e * AL;
                        float d[AL];
                        float minx = kBig;
                        {
                            const float4 *na4 = reinterpret_cast<const float4 *>(tl->na + c0);
#pragma unroll
                            for (int j4 = 0; j4 < AL / 4; ++j4) {
                                const float4 a = na4[j4];
                                const float aa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const int j = j4 * 4 + jj;
                                    const float x = fmaf(__uint_as_float(v[j]), -2.f, aa[jj]);
                                    minx = fminf(minx, x);
                                    d[j] = sqrt_approx(fabsf(x));
                                }
                            }
                        }
                        // the chunk's accumulator columns are in registers: hand them back (fill value of the lane's next
                        // gallery row) and request the next episode
                        tmem_st_fill8_x20(trow + c0, fv);
                        const bool more = e + 1 < eend;
                        if (more) {
                            tmem_ld_x20(trow + c0 + AL, v);
                        } else {
                            // The warp's last TMEM read of the tile is in registers and its columns are refilled: the
                            // accumulator goes back to the MMA issuers NOW; the taps, the compare and the candidate
                            // handling of this last episode run while the tensor pipe is already on the tile after next.
                            tmem_st_wait();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(tempty_leader0 + static_cast<uint32_t>(acc) * 8u);
                            released = true;
                        }
                        const float4 *th4 = reinterpret_cast<const float4 *>(tl->thr + c0);
                        bool any0 = false, any1 = false, any2 = false, any3 = false;
#pragma unroll
                        for (int j4 = 0; j4 < AL / 4; ++j4) {
                            const float4 th = th4[j4];
                            const int j = j4 * 4;
                            const float t0 = j ? fmaf(w, d[j - 1], fmaf(w, d[j + 1], d[j])) : fmaf(w, d[1], d[0]);
                            const float t1 = fmaf(w, d[j], fmaf(w, d[j + 2], d[j + 1]));
                            const float t2 = fmaf(w, d[j + 1], fmaf(w, d[j + 3], d[j + 2]));
                            const float t3 = (j + 4 < AL) ? fmaf(w, d[j + 2], fmaf(w, d[j + 4], d[j + 3])) : fmaf(w, d[AL - 2], d[AL - 1]);
                            any0 |= (t0 <= th.x); any1 |= (t1 <= th.y);
                            any2 |= (t2 <= th.z); any3 |= (t3 <= th.w);
                        }
                        const bool guard = minx < xfloor;
                        const bool hit = (any0 || any1) || (any2 || any3) || guard;
                        if (__any_sync(0xffffffffu, hit) && !EOSVR_EXP_ON(p, 32)) {
                            // ---- some lane is below a threshold (or inside the cancellation guard): the taps once more
                            //      (same arithmetic, same bits), now with a per-lane mask of the columns concerned ----
                            float t[AL];
                            unsigned m = 0;
#pragma unroll
                            for (int j4 = 0; j4 < AL / 4; ++j4) {
                                const float4 th = th4[j4];
                                const float tt[4] = {th.x, th.y, th.z, th.w};
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const int j = j4 * 4 + jj;
                                    float tj = d[j];
                                    if (j + 1 < AL) tj = fmaf(w, d[j + 1], tj);
                                    if (j > 0) tj = fmaf(w, d[j - 1], tj);
                                    t[j] = tj;
                                    m |= (tj <= tt[jj] ? 1u : 0u) << j;
                                }
                            }
                            if (p.seed_mode) {
                                // seed pass: thresholds only (see the generic epilogue)
                                const bool okl = rowok && !guard;
                                unsigned mine = 0x7f800000u;
#pragma unroll
                                for (int j = 0; j < AL; ++j) {
                                    const unsigned mn = __reduce_min_sync(0xffffffffu, okl ? __float_as_uint(t[j]) : 0x7f800000u);
                                    if (lane == j) mine = mn;
                                }
                                if (lane < AL && mine != 0x7f800000u) {
                                    const int c = c0 + lane;
                                    const int32_t rm = tl->row[c];
                                    if (rm >= 0) {
                                        const float nt = fmaf(__uint_as_float(mine), kSlopMul, tl->mg[c]);
                                        if (nt < __uint_as_float(*reinterpret_cast<volatile unsigned int *>(&tl->thr[c]))) {
                                            atomicMin(&tl->thr[c], __float_as_uint(nt));
                                            atomicMin(p.gthr + rm, __float_as_uint(nt));
                                        }
                                    }
                                }
                            } else {
                                // Inside the cancellation guard (a tap with x~ < 65 E2: the margin in mg[] assumes larger
                                // taps) nothing is screened: the column of the small tap and its two neighbours are handed
                                // to the exact re-rank unconditionally ("unsafe").  Only near-duplicates of a probe row
                                // get here.
                                unsigned um = 0;
                                if (guard) {
                                    unsigned sm = 0;
#pragma unroll
                                    for (int j = 0; j < AL; ++j) sm |= (d[j] < dfloor ? 1u : 0u) << j;
                                    um = (sm | (sm << 1) | (sm >> 1)) & ((1u << AL) - 1u);
                                }
                                if (!rowok) { m = 0; um = 0; }
                                m &= ~um;
                                unsigned cm = __reduce_or_sync(0xffffffffu, m | um);
#pragma unroll 1
                                while (cm) {
                                    const int jc = __ffs(cm) - 1;                  // warp-uniform
                                    cm &= cm - 1;
                                    if (ns == kStashCols) { drain_stash(); ns = 0; }
                                    const float tj = pick20(t, jc);
                                    const uint32_t val = ((um >> jc) & 1u) ? kCandUnsafe
                                                       : (((m >> jc) & 1u) ? __float_as_uint(tj) : 0x7f800000u);
                                    tl->sval[warp - 4][ns][lane] = val;
                                    if (lane == 0) { tl->scol[warp - 4][ns] = c0 + jc; tl->sgt[warp - 4][ns] = cur.gt; }
                                    ++ns;
                                }
                            }
                        }
                        if (more) tmem_ld_wait_x20(v);
                    }
                    if (prof) e_loop += clock64() - t_e1;
                }
            } else
            if (!EOSVR_EXP_ON(p, 1) && cend > cbeg) {
                const bool rowok = g < p.G;
                // eight registers holding the fill value: the source of the refill stores (a tcgen05.st wants consecutive
                // registers; or-ing in zeros the compiler cannot see through keeps it from re-copying one register into
                // fifteen others in front of every store)
                uint32_t fv[8];
                {
                    const volatile unsigned int *zp = tl->zeros;
#pragma unroll
                    for (int i = 0; i < 8; ++i) fv[i] = fill | zp[i];
                }
                uint32_t v[kChunk];                      // TMEM landing registers
                uint32_t vleft = 0, vright = 0;          // the neighbouring column groups' boundary columns
                float dprev = kBig, dright = kBig;
                // first chunk and the boundary columns of the neighbouring column groups: one wait for all
                if (cbeg > 0) tmem_ld_x1(trow + cbeg * kChunk - 1, vleft);
                if (cend * kChunk < BN) tmem_ld_x1(trow + cend * kChunk, vright);
                tmem_ld_x16(trow + cbeg * kChunk, v);
                tmem_ld_wait();
                if (cbeg > 0) dprev = sqrt_approx(fabsf(fmaf(__uint_as_float(vleft), -2.f, tl->na[cbeg * kChunk - 1])));
                if (cend * kChunk < BN) dright = sqrt_approx(fabsf(fmaf(__uint_as_float(vright), -2.f, tl->na[cend * kChunk])));
                // nobody refills a column before every warp of the lane quadrant holds its boundary reads
                if (!EOSVR_EXP_ON(p, 256)) named_bar_sync(2 + q, 32 * EG);
                long long t_e1 = 0;
                if (prof) { t_e1 = clock64(); e_pre += t_e1 - t_e0; }

                // One chunk of 16 accumulator columns per iteration.  x = |a|^2 + |b|^2 - 2 a.b and d = sqrt(x) consume the
                // landing registers, which are then immediately re-used for the NEXT chunk's TMEM read; that read is
                // waited for only when the last column's right-hand neighbour is needed, so its latency overlaps the taps.
#pragma unroll 1
                for (int ch = cbeg; ch < cend; ++ch) {
                    const int c0 = ch * kChunk;
                    float d[kChunk];
                    float minx = kBig;
                    {
                        const float4 *na4 = reinterpret_cast<const float4 *>(tl->na + c0);
#pragma unroll
                        for (int j4 = 0; j4 < kChunk / 4; ++j4) {
                            const float4 a = na4[j4];
                            const float aa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = j4 * 4 + jj;
                                const float x = fmaf(__uint_as_float(v[j]), -2.f, aa[jj]);
                                minx = fminf(minx, x);
                                d[j] = EOSVR_EXP_ON(p, 512) ? fabsf(x) : sqrt_approx(fabsf(x));
                            }
                        }
                    }
                    const bool more = ch + 1 < cend;
                    if (more) tmem_ld_x16(trow + c0 + kChunk, v);
                    bool any0 = false, any1 = false, any2 = false, any3 = false;      // four short predicate chains
                    const float dprev_in = dprev;
                    const float4 *wl4 = reinterpret_cast<const float4 *>(tl->wl + c0);
                    const float4 *wr4 = reinterpret_cast<const float4 *>(tl->wr + c0);
                    const float4 *th4 = reinterpret_cast<const float4 *>(tl->thr + c0);
                    float4 l3, r3, th3;
#pragma unroll
                    for (int j4 = 0; j4 < kChunk / 4; ++j4) {
                        const float4 l = wl4[j4], r = wr4[j4], th = th4[j4];
                        const float dl0 = j4 ? d[j4 * 4 - 1] : dprev_in;
                        const float t0 = fmaf(l.x, dl0, fmaf(r.x, d[j4 * 4 + 1], d[j4 * 4 + 0]));
                        const float t1 = fmaf(l.y, d[j4 * 4 + 0], fmaf(r.y, d[j4 * 4 + 2], d[j4 * 4 + 1]));
                        const float t2 = fmaf(l.z, d[j4 * 4 + 1], fmaf(r.z, d[j4 * 4 + 3], d[j4 * 4 + 2]));
                        any0 |= (t0 <= th.x); any1 |= (t1 <= th.y); any2 |= (t2 <= th.z);
                        if (j4 < kChunk / 4 - 1) {
                            const float t3 = fmaf(l.w, d[j4 * 4 + 2], fmaf(r.w, d[j4 * 4 + 4], d[j4 * 4 + 3]));
                            any3 |= (t3 <= th.w);
                        } else { l3 = l; r3 = r; th3 = th; }
                    }
                    // last column: its right-hand neighbour is the first column of the next chunk (or of the next group)
                    float dn = dright;
                    if (more) {
                        tmem_ld_wait_x16(v);
                        dn = sqrt_approx(fabsf(fmaf(__uint_as_float(v[0]), -2.f, tl->na[c0 + kChunk])));
                    }
                    {
                        const float t15 = fmaf(l3.w, d[kChunk - 2], fmaf(r3.w, dn, d[kChunk - 1]));
                        any3 |= (t15 <= th3.w);
                    }
                    dprev = d[kChunk - 1];
                    const bool guard = (minx < xfloor) || (fminf(dprev_in, dn) < dfloor);
                    const bool hit = (any0 || any1) || (any2 || any3) || guard;
                    if ((DIAG && p.dbg != nullptr) || (__any_sync(0xffffffffu, hit) && !EOSVR_EXP_ON(p, 32))) {
                        // ---- slow path (whole warp; ~10 % of the chunks): the taps once more from the distances still in
                        //      registers (same arithmetic, same bits), now keeping every column's value and a mask of
                        //      the columns in which some lane is below its threshold (all columns if the cancellation
                        //      guard fired).  Per such column: tighten the threshold with the warp minimum, then park
                        //      what is still below it in the warp's staging buffer.  The fast path keeps only d[] alive
                        //      across the vote, so this block costs it no registers. ----
                        float t[kChunk];
                        unsigned cm = 0;
#pragma unroll
                        for (int j4 = 0; j4 < kChunk / 4; ++j4) {
                            const float4 l = wl4[j4], r = wr4[j4], th = th4[j4];
                            const float ll[4] = {l.x, l.y, l.z, l.w}, rr[4] = {r.x, r.y, r.z, r.w}, tt[4] = {th.x, th.y, th.z, th.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = j4 * 4 + jj;
                                const float dl = j ? d[j - 1] : dprev_in;
                                const float dr = (j < kChunk - 1) ? d[j + 1] : dn;
                                t[j] = fmaf(ll[jj], dl, fmaf(rr[jj], dr, d[j]));
                                cm |= (t[j] <= tt[jj] ? 1u : 0u) << j;
                            }
                        }
                        if (DIAG && p.dbg != nullptr && rowok && !p.seed_mode) {
#pragma unroll
                            for (int j = 0; j < kChunk; ++j) {
                                const int32_t rm = tl->row[c0 + j];
                                if (rm >= 0) p.dbg[static_cast<int64_t>(rm) * p.G + g] = t[j];
                            }
                        }
                        if (p.seed_mode) {
                            // seed pass: nothing is appended, thresholds only.  Warp minimum of every column in one REDUX
                            // each (lanes inside the cancellation guard stay out: their values are not trusted), then
                            // lane j publishes column j -- sixteen columns at once instead of a serial loop with ballots
                            // (the seed pass starts from infinite thresholds, so EVERY chunk comes through here).
                            const bool okl = rowok && !guard;
                            unsigned mine = 0x7f800000u;
#pragma unroll
                            for (int j = 0; j < kChunk; ++j) {
                                const unsigned m = __reduce_min_sync(0xffffffffu, okl ? __float_as_uint(t[j]) : 0x7f800000u);
                                if (lane == j) mine = m;
                            }
                            if (lane < kChunk && mine != 0x7f800000u) {
                                const int c = c0 + lane;
                                const int32_t rm = tl->row[c];
                                if (rm >= 0) {
                                    const float nt = fmaf(__uint_as_float(mine), kSlopMul, tl->mg[c]);
                                    if (nt < __uint_as_float(*reinterpret_cast<volatile unsigned int *>(&tl->thr[c]))) {
                                        atomicMin(&tl->thr[c], __float_as_uint(nt));
                                        atomicMin(p.gthr + rm, __float_as_uint(nt));
                                    }
                                }
                            }
                            cm = 0;
                        }
                        if (guard) cm = p.seed_mode ? 0u : 0xFFFFu;
                        cm = __reduce_or_sync(0xffffffffu, cm);
                        const bool any_guard = __any_sync(0xffffffffu, guard);
#pragma unroll 1
                        while (cm) {
                            const int jc = __ffs(cm) - 1;                  // warp-uniform
                            cm &= cm - 1;
                            const int c = c0 + jc;
                            const int32_t rm = tl->row[c];
                            if (rm < 0) continue;                          // warp-uniform
                            const float tj = pick16(t, jc);                // warp-uniform index: a jump, no local memory
                            bool uns = false;
                            float extra = 0.f;
                            if (any_guard) {
                                // Small squared distances amplify the screening error of the distance: |d~ - d| <=
                                // E2 / (d~ + d).  Three bands of the smallest tap (DESIGN.md "Error bound"):
                                //   x~ >= 65 E2 : the margin in mg[] holds (|d~ - d| <= sqrt(E2)/16);
                                //   4 E2 <= x~ < 65 E2 : |d~ - d| <= 0.268 sqrt(E2) = 4.29 x that -- still screened, with
                                //     3.5 margins of slack on either side (candidate test, threshold update, stored value);
                                //   x~ < 4 E2 : not trusted at all, always a candidate ("unsafe").
                                const float dj = pick16(d, jc);
                                const float dl = jc ? pick16(d, jc - 1) : dprev_in;
                                const float dr = (jc < kChunk - 1) ? pick16(d, jc + 1) : dn;
                                const float m3 = fminf(dj, fminf(tl->wl[c] > 0.f ? dl : kBig, tl->wr[c] > 0.f ? dr : kBig));
                                if (m3 < dfloor_hard) uns = rowok;
                                else if (m3 < dfloor) extra = 3.5f * tl->mg[c];
                            }
                            float thr = __uint_as_float(*reinterpret_cast<volatile unsigned int *>(&tl->thr[c]));
                            bool pass = rowok && !uns && (tj <= thr + extra);
                            if (__ballot_sync(0xffffffffu, pass)) {
                                // warp minimum in one instruction: t >= 0, so float order == unsigned order of the bits
                                const float mn = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(pass ? tj + extra : kBig)));
                                const float nt = fmaf(mn, kSlopMul, tl->mg[c]);
                                if (nt < thr) {
                                    if (lane == 0) {
                                        atomicMin(&tl->thr[c], __float_as_uint(nt));
                                        atomicMin(p.gthr + rm, __float_as_uint(nt));
                                    }
                                    thr = nt;
                                }
                                pass = pass && (tj <= thr + extra);
                            }
                            const bool app = pass || uns;
                            const unsigned ma = __ballot_sync(0xffffffffu, app);
                            if (ma) {
                                if (app) {
                                    StagedCand sc;
                                    sc.rm = rm; sc.g = static_cast<int32_t>(g);
                                    // (wide-band values are stored 3.5 margins low: the re-rank may rely on
                                    //  stored value - margin/2 <= true value for every candidate)
                                    sc.tbits = uns ? kCandUnsafe : __float_as_uint(fmaxf(tj - extra, 0.f));
                                    wstage[wn + __popc(ma & ((1u << lane) - 1u))] = sc;
                                }
                                wn += __popc(ma);
                                if (wn >= 32) { flush_staged(p, wstage, wn, lane); wn = 0; }
                            }
                        }
                    }
                    // refill one chunk behind (the next group's boundary reads are protected by the quadrant barrier)
                    if (ch > cbeg && !EOSVR_EXP_ON(p, 128)) tmem_st_fill8_x16(trow + c0 - kChunk, fv);
                }
                if (prof) e_loop += clock64() - t_e1;
                if (!EOSVR_EXP_ON(p, 128)) tmem_st_fill8_x16(trow + (cend - 1) * kChunk, fv);
                tmem_st_wait();
            } else if (!EOSVR_EXP_ON(p, 256)) {
                named_bar_sync(2 + q, 32 * EG);           // (a column group is empty when the tile has fewer chunks than groups)
            }
            if (!released) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader0 + static_cast<uint32_t>(acc) * 8u);
            }
            if (++acc == p.acc_stages) { acc = 0; accphase ^= 1; }
            if (pend_thr != 0xFFFFFFFFu) atomicMin(&tl->thr[te], pend_thr);
            if constexpr (AL > 0) {
                if (ns) { drain_stash(); ns = 0; }       // before the unit (and with it the per-column arrays) can change
            }
            if (prof) e_busy += clock64() - t_e0;
            cursor_next(p, cur, npairs);
            cursor_next(p, ahead, npairs);
        }
        if (wn) flush_staged(p, wstage, wn, lane);
        if (prof && lane == 0) {
            atomicAdd(&p.ctr->cyc_epi_busy, e_busy); atomicAdd(&p.ctr->cyc_epi_wait, e_wait);
            atomicAdd(&p.ctr->cyc_epi_pre, e_pre); atomicAdd(&p.ctr->cyc_epi_loop, e_loop);
        }
    } else {
        if (kRepartition) reg_release<kLightRegs>();             // idle warps (issuer slots of the peer CTA)
    }

    tc_fence_before();
    cluster_sync();          // neither CTA may exit (or free TMEM) while its peer can still reach it
    if (warp == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
}

// -------------------------------------------------------------------------------------------
// Exact evaluation (the reference's arithmetic, with the reference's ROUNDING): network_test.py:208 is scipy's
// cdist(..., 'euclidean') on float32 rows -- promoted to float64, then, per pair, ONE sequential pass
//   s = 0;  for k = 0 .. D-1:  e = a[k] - b[k];  s = s + e*e        (product rounded, then the sum: no FMA)
// and sqrt(s) (scipy/spatial/src/distance_metrics.h, transform_reduce_2d_: the 4-way unrolling is over ROWS, the
// reduction along a row is sequential; checked bit for bit in float64 by the CPU tests).  Then :109 (float32 cast)
// and models.py:42-56 as the float32 FMA chain
//   acc = lam1*d[p-1];  acc = fma(lam2, d[p], acc);  acc = fma(lam1, d[p+1], acc)
// with zero padding at episode ends.
//
// A float64 sum in another order differs from scipy's in its last bits, and about one distance in 10^6 then rounds
// to the neighbouring float32.  The sequential chain cannot be parallelised, so the evaluation is FILTERED:
//   (1) fast sum S in any order (lanes stride k, FMA, tree reduction).  Both S and scipy's sum are within
//       gamma_(D+1) of the exact sum of the (identical) squared differences, so |S_scipy - S| <= delta * S with
//       delta = (2 D + 8) 2^-53;
//   (2) if float32(sqrt(S (1 - delta))) == float32(sqrt(S (1 + delta))) (directed roundings; sqrt and the cast are
//       monotone and correctly rounded), scipy's float32 is that value -- decided;
//   (3) otherwise (a rounding boundary of float32 lies inside the interval: ~1e-6 of the distances) the warp
//       evaluates scipy's sequential chain itself: the lanes form the products of 32 consecutive k in parallel and
//       the running sums take them in order through shuffles (warp_seq3).
// tests/golden/golden_order_sensitive.npz holds pairs constructed to fall into (3).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ double seq_delta(int D) { return (2.0 * D + 8.0) * 1.1102230246251565e-16; }   // (2D+8) 2^-53

// float32(sqrt(s_scipy)) if it is determined by S and delta (see above); false: ambiguous
__device__ __forceinline__ bool sqrt_f32_decided(double S, double delta, float &out)
{
    const float lo = static_cast<float>(sqrt(__dmul_rd(S, 1.0 - delta)));
    const float hi = static_cast<float>(sqrt(__dmul_ru(S, 1.0 + delta)));
    out = lo;
    return lo == hi;
}

// scipy's sequential float64 sums of squared differences of gallery row b0 against up to three probe rows,
// warp-cooperative: lane l forms the products of element k0 + l, the running sums add them in the order of k
// (a lane beyond D contributes +0.0, which leaves a non-negative sum unchanged).  All lanes return the sums.
__device__ __forceinline__ void warp_seq3(const float *a0, const float *a1, const float *a2, const void *__restrict__ gal,
                                          int gdt, int64_t b0, int D, int lane, double &s0, double &s1, double &s2)
{
    s0 = s1 = s2 = 0.0;
    for (int k0 = 0; k0 < D; k0 += 32) {
        const int k = k0 + lane;
        double p0 = 0.0, p1 = 0.0, p2 = 0.0;
        if (k < D) {
            const double bv = ld_feat(gal, gdt, b0 + k);
            const double e0 = __dsub_rn(static_cast<double>(a0[k]), bv);
            const double e1 = __dsub_rn(static_cast<double>(a1[k]), bv);
            const double e2 = __dsub_rn(static_cast<double>(a2[k]), bv);
            p0 = __dmul_rn(e0, e0); p1 = __dmul_rn(e1, e1); p2 = __dmul_rn(e2, e2);
        }
#pragma unroll 2
        for (int i = 0; i < 32; ++i) {
            s0 = __dadd_rn(s0, __shfl_sync(0xffffffffu, p0, i));
            s1 = __dadd_rn(s1, __shfl_sync(0xffffffffu, p1, i));
            s2 = __dadd_rn(s2, __shfl_sync(0xffffffffu, p2, i));
        }
    }
}

// The three float32 distances of the taps from fast sums y0..y2 (any order), falling back to scipy's order where
// the float32 rounding is not decided by them.  Warp-cooperative (all lanes hold the same y); a0/a1/a2 may live in
// shared or global memory.
__device__ __forceinline__ void taps_f32(double y0, double y1, double y2, bool hl, bool hr, const float *a0, const float *a1,
                                         const float *a2, const void *__restrict__ gal, int gdt, int64_t b0, int D, int lane,
                                         float &d0, float &d1, float &d2, unsigned long long *n_seq)
{
    const double delta = seq_delta(D);
    bool ok = sqrt_f32_decided(y1, delta, d1);
    d0 = 0.f; d2 = 0.f;
    if (hl) ok = sqrt_f32_decided(y0, delta, d0) && ok;
    if (hr) ok = sqrt_f32_decided(y2, delta, d2) && ok;
    if (!ok) {                                               // warp-uniform
        if (lane == 0 && n_seq) atomicAdd(n_seq, 1ull);      // (statistics: the tests check that the constructed pairs get here)
        double s0, s1, s2;
        warp_seq3(a0, a1, a2, gal, gdt, b0, D, lane, s0, s1, s2);
        d0 = hl ? static_cast<float>(sqrt(s0)) : 0.f;
        d1 = static_cast<float>(sqrt(s1));
        d2 = hr ? static_cast<float>(sqrt(s2)) : 0.f;
    }
}

__device__ __forceinline__ float exact_t(const float *__restrict__ probes, int64_t P, int D, int rpe,
                                         int64_t p, const void *__restrict__ gal, int gdt, int64_t b0, float lam1, float lam2, int lane,
                                         unsigned long long *n_seq = nullptr)
{
    const int r = static_cast<int>(p % rpe);
    const bool hl = r > 0, hr = (r + 1 < rpe) && (p + 1 < P);
    const float *a1 = probes + p * D;
    const float *a0 = hl ? a1 - D : a1;
    const float *a2 = hr ? a1 + D : a1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if ((D & 3) == 0) {                                  // 16-byte loads (every row starts 16-byte aligned then)
        const float4 *a04 = reinterpret_cast<const float4 *>(a0), *a14 = reinterpret_cast<const float4 *>(a1),
                     *a24 = reinterpret_cast<const float4 *>(a2);
        const int64_t b04 = b0 >> 2;
        for (int k = lane; k < (D >> 2); k += 32) {
            const float4 b = ld_feat4(gal, gdt, b04 + k);
            const float4 q0 = a04[k], q1 = a14[k], q2 = a24[k];
            const double bb[4] = {b.x, b.y, b.z, b.w};
            const float qq0[4] = {q0.x, q0.y, q0.z, q0.w}, qq1[4] = {q1.x, q1.y, q1.z, q1.w}, qq2[4] = {q2.x, q2.y, q2.z, q2.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double e0 = static_cast<double>(qq0[e]) - bb[e];
                const double e1 = static_cast<double>(qq1[e]) - bb[e];
                const double e2 = static_cast<double>(qq2[e]) - bb[e];
                s0 += e0 * e0; s1 += e1 * e1; s2 += e2 * e2;
            }
        }
    } else {
        for (int k = lane; k < D; k += 32) {
            const double bv = ld_feat(gal, gdt, b0 + k);
            const double e0 = static_cast<double>(a0[k]) - bv;
            const double e1 = static_cast<double>(a1[k]) - bv;
            const double e2 = static_cast<double>(a2[k]) - bv;
            s0 += e0 * e0; s1 += e1 * e1; s2 += e2 * e2;
        }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    float d0, d1, d2;
    taps_f32(s0, s1, s2, hl, hr, a0, a1, a2, gal, gdt, b0, D, lane, d0, d1, d2, n_seq);
    float acc = __fmul_rn(lam1, d0);
    acc = __fmaf_rn(lam2, d1, acc);
    acc = __fmaf_rn(lam1, d2, acc);
    return acc;
}

struct RerankParams {
    const float *probes;
    const void *gal;         // gallery rows, float32 or bfloat16 (gal_dtype)
    int32_t gal_dtype;
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
    const OvfCand *ovf;
    int32_t ovf_cap;
    const float *epsd, *wl, *wr;   // per plan column (column_margin)
    int32_t planR, planBN, planHalo;
    int32_t prof;            // EOSVR_EXP bit 64: phase timing of k_rerank_rows into the cycle counters
    int32_t rows_per_block;  // consecutive probe rows per k_rerank_rows block (1..kRrRowsPerBlock)
};

// The definition's sequential chains of the cosine (rare path of exact_negcos), warp-cooperative like warp_seq3.
__device__ __forceinline__ float warp_seq_negcos(const float *a, const void *__restrict__ gal, int gdt, int64_t b0, int D, int lane)
{
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (int k0 = 0; k0 < D; k0 += 32) {
        const int k = k0 + lane;
        double p0 = 0.0, p1 = 0.0, p2 = 0.0;
        if (k < D) {
            const double x = a[k], y = ld_feat(gal, gdt, b0 + k);
            p0 = __dmul_rn(x, y); p1 = __dmul_rn(x, x); p2 = __dmul_rn(y, y);
        }
#pragma unroll 2
        for (int i = 0; i < 32; ++i) {
            dot = __dadd_rn(dot, __shfl_sync(0xffffffffu, p0, i));
            na = __dadd_rn(na, __shfl_sync(0xffffffffu, p1, i));
            nb = __dadd_rn(nb, __shfl_sync(0xffffffffu, p2, i));
        }
    }
    return 0.f - static_cast<float>(__ddiv_rn(dot, __dmul_rn(sqrt(na), sqrt(nb))));
}

// Cosine metric, exactly: float64 dot / (|a| |b|) on the original rows (0 for a zero row), rounded to
// float32; returned negated (score domain).  The definition sums dot, |a|^2 and |b|^2 each in ONE sequential
// float64 pass (products rounded before they are added); filtered like the distances: both the fast value c and the
// sequential one are within (2 D + 7) 2^-53 of the true cosine (Cauchy-Schwarz on the dot product's rounding
// errors, |c| <= 1), so a float32 rounding that is the same at c -+ (4 D + 32) 2^-53 is the answer, and otherwise the
// warp evaluates the sequential chains.  Warp-cooperative; all lanes return the value.
__device__ __forceinline__ float exact_negcos(const float *__restrict__ a, const void *__restrict__ gal, int gdt, int64_t b0,
                                              int D, int lane, unsigned long long *n_seq = nullptr)
{
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (int k = lane; k < D; k += 32) {
        const double x = a[k], y = ld_feat(gal, gdt, b0 + k);
        dot += x * y; na += x * x; nb += y * y;
    }
    dot = warp_sum(dot); na = warp_sum(na); nb = warp_sum(nb);
    const double den = sqrt(na) * sqrt(nb);
    if (!(den > 0.0)) return 0.f;                            // a zero row (the same in any summation order)
    const double c = dot / den, dc = (4.0 * D + 32.0) * 1.1102230246251565e-16;
    const float lo = static_cast<float>(__dadd_rd(c, -dc)), hi = static_cast<float>(__dadd_ru(c, dc));
    if (lo == hi) return 0.f - lo;                           // 0 - c: never -0 (packed order)
    if (lane == 0 && n_seq) atomicAdd(n_seq, 1ull);
    return warp_seq_negcos(a, gal, gdt, b0, D, lane);
}

__device__ __forceinline__ float exact_score(const RerankParams &p, int64_t row, int64_t g, int lane)
{
    if (p.metric == EOSVR_METRIC_COSINE) return exact_negcos(p.probes + row * p.D, p.gal, p.gal_dtype, g * p.D, p.D, lane, &p.ctr->n_seq);
    return exact_t(p.probes, p.P, p.D, p.rpe, row, p.gal, p.gal_dtype, g * p.D, p.lam1, p.lam2, lane, &p.ctr->n_seq);
}

// Spill-over candidates (row lists that filled up): one warp per entry, grid-strided over the warps of the
// calling kernel.  Nothing to do when the buffer is empty (the normal case).
__device__ __forceinline__ void rerank_spilled(const RerankParams &p, int64_t warp_id, int64_t n_warps, int lane)
{
    const unsigned cnt = p.ctr->ovf_count;
    if (cnt == 0) return;
    const int64_t n = cnt < static_cast<unsigned>(p.ovf_cap) ? cnt : p.ovf_cap;
    unsigned long long done = 0;
    for (int64_t w = warp_id; w < n; w += n_warps) {
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
    rerank_spilled(p, (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nw, lane);
}

// Block-per-row re-rank, warp-per-candidate.  The three probe rows the taps need are staged in shared memory
// once per row; then, in two steps:
//   (1) float32 pre-filter: the candidates still below the row's threshold are sorted by screening value and the
//       block's warps pull them in ascending order, each warp evaluating one candidate with float32 direct
//       differences (relative error of the smoothed value < kF32Rel, orders of magnitude tighter than the 16-bit
//       screening) and tightening the shared bound; a warp stops as soon as the next screening value cannot win;
//   (2) the candidates within 2*kF32Rel of the best float32 value -- the winner and its exact ties, normally ONE
//       candidate -- are evaluated exactly as the reference does (float64 direct differences, scipy's rounding by
//       the filter of taps_f32 -> float32 -> float32 FMA chain) and merged by packed 64-bit atomicMin.
// No block-wide barrier sits between candidates, the gallery-row loads of the warps overlap, and float64 work
// (conversions run at 16/clk/SM) is spent only where it decides the answer.
constexpr int kRrThreads = 128;
#ifndef EOSVR_RR_MINBLOCKS
#define EOSVR_RR_MINBLOCKS 6
#endif
// resident blocks per SM the compiler must allow: the kernel is bound by the latency of a row's dependent steps, so
// occupancy is what it needs (cfg-3 / cfg-2: 128 registers = 4 blocks 0.198 / 0.444 ms; 80 registers = 6 blocks 0.165 / 0.394 ms;
// 64 registers = 8 blocks spills and loses: 0.186 ms on cfg-3 -- profiles/r02_ab_rerank_occupancy.txt)
constexpr int kRrMinBlocks = EOSVR_RR_MINBLOCKS;
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
__global__ void __launch_bounds__(kRrThreads, kRrMinBlocks)
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
    const int D4 = p.D >> 2;
    unsigned long long appended = 0, done = 0, unsafe_n = 0, ev32 = 0;   // ev32: per lane-0 thread
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
          s_eps[tid] = 0.5f * column_margin(p.epsd, p.wl, p.wr, (row / p.planR) * p.planBN + p.planHalo + (row % p.planR));
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
                const int64_t gp = static_cast<int64_t>(g) * D4;          // float4 index of the gallery row
                const int gdt = p.gal_dtype;
                float x0 = 0.f, x1 = 0.f, x2 = 0.f;
                if (COS) {
#pragma unroll 4
                    for (int k = lane; k < D4; k += 32) {
                        const float4 b = ld_feat4(p.gal, gdt, gp + k);
                        const float4 q = sp1[k];
                        x0 = fmaf(q.x, b.x, x0); x0 = fmaf(q.y, b.y, x0); x0 = fmaf(q.z, b.z, x0); x0 = fmaf(q.w, b.w, x0);
                        x1 = fmaf(b.x, b.x, x1); x1 = fmaf(b.y, b.y, x1); x1 = fmaf(b.z, b.z, x1); x1 = fmaf(b.w, b.w, x1);
                        x2 = fmaf(q.x, q.x, x2); x2 = fmaf(q.y, q.y, x2); x2 = fmaf(q.z, q.z, x2); x2 = fmaf(q.w, q.w, x2);
                    }
                } else {
#pragma unroll 4
                    for (int k = lane; k < D4; k += 32) {
                        const float4 b = ld_feat4(p.gal, gdt, gp + k);
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
                    ++ev32;
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
                const int64_t gp = static_cast<int64_t>(g) * D4;
                const int gdt = p.gal_dtype;
                double y0 = 0.0, y1 = 0.0, y2 = 0.0;
                if (COS) {
                    for (int k = tid; k < D4; k += kRrThreads) {
                        const float4 b = ld_feat4(p.gal, gdt, gp + k);
                        const float4 q = sp1[k];
                        const double bb[4] = {b.x, b.y, b.z, b.w}, qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) { y0 += qq[e] * bb[e]; y1 += bb[e] * bb[e]; y2 += qq[e] * qq[e]; }
                    }
                } else {
                    for (int k = tid; k < D4; k += kRrThreads) {
                        const float4 b = ld_feat4(p.gal, gdt, gp + k);
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
                if (warp == 0) {                                         // the whole warp: the fallback is warp-cooperative
                    y0 = y1 = y2 = 0.0;
#pragma unroll
                    for (int w = 0; w < kRrThreads / 32; ++w) { y0 += s_part64[w][0]; y1 += s_part64[w][1]; y2 += s_part64[w][2]; }
                    float acc;
                    if (COS) {
                        // the float32 rounding of the fast value is the definition's unless a rounding boundary lies
                        // within its error: then the warp evaluates the sequential chains (exact_negcos)
                        const double den = sqrt(y2) * sqrt(y1);
                        const double c = den > 0.0 ? y0 / den : 0.0, dc = (4.0 * p.D + 32.0) * 1.1102230246251565e-16;
                        const float lo = static_cast<float>(__dadd_rd(c, -dc)), hi = static_cast<float>(__dadd_ru(c, dc));
                        acc = 0.f - lo;
                        if (den > 0.0 && lo != hi)
                            acc = exact_negcos(reinterpret_cast<const float *>(sp1), p.gal, gdt, static_cast<int64_t>(g) * p.D, p.D, lane, &p.ctr->n_seq);
                    } else {
                        float d0, d1, d2;
                        taps_f32(y0, y1, y2, hl, hr, reinterpret_cast<const float *>(sp0), reinterpret_cast<const float *>(sp1),
                                 reinterpret_cast<const float *>(sp2), p.gal, gdt, static_cast<int64_t>(g) * p.D, p.D, lane, d0, d1, d2, &p.ctr->n_seq);
                        acc = __fmul_rn(p.lam1, d0);
                        acc = __fmaf_rn(p.lam2, d1, acc);
                        acc = __fmaf_rn(p.lam1, d2, acc);
                    }
                    if (lane == 0) {
                        const unsigned long long v = pack_score_idx(acc, static_cast<uint32_t>(p.offset + g));
                        loc = v < loc ? v : loc;
                        ++done;
                    }
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
    if (lane == 0 && ev32) atomicAdd(&p.ctr->n_eval32, ev32);
    rerank_spilled(p, static_cast<int64_t>(blockIdx.x) * (kRrThreads / 32) + warp,
                   static_cast<int64_t>(gridDim.x) * (kRrThreads / 32), lane);
}

// -------------------------------------------------------------------------------------------
// Warp-per-row re-rank (the default for 512-element rows; EOSVR_RR=0 / 1 forces one kernel or the other): the same
// two steps with the same arithmetic as k_rerank_rows, but a probe row belongs to ONE warp, so nothing in a row's chain
// of dependent steps waits at a block barrier; one resident wave of warps, each taking several rows.
//   * candidates are read 32 at a time; the lanes keep what is still below the row's final threshold and the warp
//     takes them in ascending order of their screening value (one REDUX per pick);
//   * float32 evaluation of a candidate: the lanes stride the row exactly like k_rerank_rows (same per-lane FMA
//     order, same butterfly: the kF32Rel bound is the same); the gallery row of the NEXT candidate in order is
//     requested before the current one is reduced (the pick order does not depend on the results, only the stop does);
//   * the probe rows come from global memory (L1 hits after the first candidate; neighbouring rows belong to the
//     neighbouring warps of the block);
//   * candidates within 2 kF32Rel of the best float32 value are evaluated exactly (exact_score: filtered float64).
// -------------------------------------------------------------------------------------------
constexpr int kRwWarps = 4;
#ifndef EOSVR_RW_MINBLOCKS
#define EOSVR_RW_MINBLOCKS 5
#endif
// BF: bfloat16 gallery rows; D512: the rows have exactly 512 elements (the metric's shape: every loop is static, the
// loads are immediate offsets from one row pointer -- integer bookkeeping was half of the generic kernel's instructions)
template <bool BF>
__device__ __forceinline__ float4 ld_row4(const void *rowp, int k)
{
    if (!BF) return reinterpret_cast<const float4 *>(rowp)[k];
    const uint2 u = reinterpret_cast<const uint2 *>(rowp)[k];
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                       __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}

template <bool COS, bool BF, bool D512>
__global__ void __launch_bounds__(32 * kRwWarps, EOSVR_RW_MINBLOCKS)
k_rerank_warp(const RerankParams p)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kRwWarps + (threadIdx.x >> 5);
    const int64_t nw = static_cast<int64_t>(gridDim.x) * kRwWarps;
    const int D4 = D512 ? 128 : (p.D >> 2);
    const size_t row_bytes = static_cast<size_t>(D4) * (BF ? 8 : 16);
    const char *galb = static_cast<const char *>(p.gal);
    const float inv_lam2 = 1.0f / p.lam2;
    const float4 *probes4 = reinterpret_cast<const float4 *>(p.probes);
    unsigned long long appended = 0, done = 0, unsafe_n = 0, ev32 = 0;

    // the row's counters and its first 32 candidates are requested one row ahead (the three loads are independent: the
    // candidate slot is read whether or not it is filled), so a row does not start with two dependent round trips
    unsigned cnt_n = 0, thr_n = 0;
    Cand cand_n;
    cand_n.g = 0; cand_n.tbits = 0x7f800000u;
    if (warp0 < p.P) { cnt_n = p.rowcnt[warp0]; thr_n = p.gthr[warp0]; cand_n = p.cand[warp0 * p.cand_cap + lane]; }
    for (int64_t row = warp0; row < p.P; row += nw) {
        const unsigned cnt = cnt_n;
        const unsigned thr_bits = thr_n;
        const Cand first = cand_n;
        if (row + nw < p.P) { cnt_n = p.rowcnt[row + nw]; thr_n = p.gthr[row + nw]; cand_n = p.cand[(row + nw) * p.cand_cap + lane]; }
        const int n = cnt < static_cast<unsigned>(p.cand_cap) ? static_cast<int>(cnt) : p.cand_cap;
        appended += cnt;
        if (n == 0) continue;
        const int r = static_cast<int>(row % p.rpe);
        const bool hl = !COS && r > 0, hr = !COS && (r + 1 < p.rpe) && (row + 1 < p.P);
        const float4 *a1 = probes4 + row * D4;
        const float4 *a0 = hl ? a1 - D4 : a1;
        const float4 *a2 = hr ? a1 + D4 : a1;
        const float eps1 = 0.5f * column_margin(p.epsd, p.wl, p.wr, (row / p.planR) * p.planBN + p.planHalo + (row % p.planR));
        float bound = __uint_as_float(thr_bits);                    // candidates above it cannot win (tightens)
        float best32 = INFINITY;
        unsigned long long loc = ~0ull;
        const Cand *list = p.cand + row * p.cand_cap;

        for (int b0 = 0; b0 < n; b0 += 32) {
            Cand c;
            c.g = 0; c.tbits = 0x7f800000u;
            const bool in = b0 + lane < n;
            if (in) c = b0 == 0 ? first : list[b0 + lane];
            const bool isuns = in && c.tbits == kCandUnsafe;
            // order key: unsafe candidates first (0), then the screening value's bits (t~ >= 0: float order == uint order)
            unsigned key = isuns ? 0u : c.tbits;
            bool keep = in && (isuns || __uint_as_float(c.tbits) <= bound);
            unsafe_n += __popc(__ballot_sync(0xffffffffu, isuns));
            // this lane's slot of the evaluated list of the batch
            int32_t eg = 0; float et = INFINITY;
            int nev = 0;
            // pick the first candidate and request its row
            unsigned kmin = __reduce_min_sync(0xffffffffu, keep ? key : 0xFFFFFFFFu);
            int src = __ffs(__ballot_sync(0xffffffffu, keep && key == kmin)) - 1;
            float4 gb[4];                                            // this lane's float4 groups k = lane + 32 i, i < 4, of the row
            int32_t g = 0;
            auto request = [&](int32_t gg) {
                const void *rp = galb + static_cast<size_t>(gg) * row_bytes;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (D512 || lane + 32 * i < D4) gb[i] = ld_row4<BF>(rp, lane + 32 * i);
            };
            if (src >= 0) { g = __shfl_sync(0xffffffffu, c.g, src); request(g); }
            while (src >= 0) {
                // the picked candidate is lane src's; it leaves the set
                const float tv = __uint_as_float(kmin);
                if (lane == src) keep = false;
                if (kmin != 0u && tv > bound) break;                 // ascending order: nothing later can win either
                // ---- float32 evaluation (k_rerank_rows's arithmetic) ----
                const int32_t gcur = g;
                float x0 = 0.f, x1 = 0.f, x2 = 0.f;
                float4 cb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) cb[i] = gb[i];
                // next pick, its row requested before this one is reduced
                kmin = __reduce_min_sync(0xffffffffu, keep ? key : 0xFFFFFFFFu);
                src = __ffs(__ballot_sync(0xffffffffu, keep && key == kmin)) - 1;
                if (src >= 0) { g = __shfl_sync(0xffffffffu, c.g, src); request(g); }
                auto accum = [&](const float4 b, int k) {
                    if (COS) {
                        const float4 q = a1[k];
                        x0 = fmaf(q.x, b.x, x0); x0 = fmaf(q.y, b.y, x0); x0 = fmaf(q.z, b.z, x0); x0 = fmaf(q.w, b.w, x0);
                        x1 = fmaf(b.x, b.x, x1); x1 = fmaf(b.y, b.y, x1); x1 = fmaf(b.z, b.z, x1); x1 = fmaf(b.w, b.w, x1);
                        x2 = fmaf(q.x, q.x, x2); x2 = fmaf(q.y, q.y, x2); x2 = fmaf(q.z, q.z, x2); x2 = fmaf(q.w, q.w, x2);
                    } else {
                        const float4 q0 = a0[k], q1 = a1[k], q2 = a2[k];
                        float e;
                        e = q0.x - b.x; x0 = fmaf(e, e, x0); e = q0.y - b.y; x0 = fmaf(e, e, x0);
                        e = q0.z - b.z; x0 = fmaf(e, e, x0); e = q0.w - b.w; x0 = fmaf(e, e, x0);
                        e = q1.x - b.x; x1 = fmaf(e, e, x1); e = q1.y - b.y; x1 = fmaf(e, e, x1);
                        e = q1.z - b.z; x1 = fmaf(e, e, x1); e = q1.w - b.w; x1 = fmaf(e, e, x1);
                        e = q2.x - b.x; x2 = fmaf(e, e, x2); e = q2.y - b.y; x2 = fmaf(e, e, x2);
                        e = q2.z - b.z; x2 = fmaf(e, e, x2); e = q2.w - b.w; x2 = fmaf(e, e, x2);
                    }
                };
                // (per lane the groups are taken in ascending k, as in k_rerank_rows: first the four requested ahead)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (D512 || lane + 32 * i < D4) accum(cb[i], lane + 32 * i);
                if (!D512) {
                    const void *rp = galb + static_cast<size_t>(gcur) * row_bytes;
                    for (int k = lane + 128; k < D4; k += 32) accum(ld_row4<BF>(rp, k), k);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    x0 += __shfl_xor_sync(0xffffffffu, x0, o);
                    x1 += __shfl_xor_sync(0xffffffffu, x1, o);
                    x2 += __shfl_xor_sync(0xffffffffu, x2, o);
                }
                float acc, nb;
                if (COS) {
                    const float den = sqrtf(x2) * sqrtf(x1);
                    acc = 0.f - (den > 0.f ? x0 / den : 0.f);
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
                if (lane == nev) { eg = gcur; et = acc; }
                ++nev; ++ev32;
                best32 = fminf(best32, acc);
                bound = fminf(bound, nb);
            }
            // ---- exact evaluation of what is within the float32 error of the best float32 value ----
            const float cut = COS ? best32 + 2.0f * kF32AbsCos : best32 * (1.0f + 2.0f * kF32Rel);
            unsigned m2 = __ballot_sync(0xffffffffu, lane < nev && et <= cut);
            while (m2) {
                const int j = __ffs(m2) - 1;
                m2 &= m2 - 1;
                const int32_t gj = __shfl_sync(0xffffffffu, eg, j);
                const float t = exact_score(p, row, gj, lane);
                const unsigned long long v = pack_score_idx(t, static_cast<uint32_t>(p.offset + gj));
                loc = v < loc ? v : loc;
                ++done;
            }
        }
        if (lane == 0 && loc != ~0ull) atomicMin(p.best + row, loc);
    }
    if (lane == 0) {
        if (appended) atomicAdd(&p.ctr->cand_count, appended);
        if (done) atomicAdd(&p.ctr->n_exact, done);
        if (unsafe_n) atomicAdd(&p.ctr->n_unsafe, unsafe_n);
        if (ev32) atomicAdd(&p.ctr->n_eval32, ev32);
    }
    rerank_spilled(p, warp0, nw, lane);
}

__device__ __forceinline__ void finalize_row(const unsigned long long *best, int64_t i, int negate,
                                             uint64_t *out_packed, float *out_score, int64_t *out_idx)
{
    const unsigned long long v = __ldcg(best + i);       // L2: other blocks' atomicMin results
    if (out_packed) out_packed[i] = v;
    if (out_score) out_score[i] = (v == ~0ull) ? __int_as_float(0x7fc00000) : (negate ? -unpack_score(v) : unpack_score(v));
    if (out_idx) out_idx[i] = (v == ~0ull) ? -1 : static_cast<int64_t>(v & 0xFFFFFFFFull);
}

// Last kernel of a match call.  Normal case (no row list overflowed beyond the shared spill buffer): unpack the
// winners, grid-strided.  Otherwise the flagged rows (ctr->overflow == 1), or all rows (== 2, eosvr_match_exact),
// are first resolved by exhaustive exact evaluation: every block derives the same ordered list of flagged rows,
// the grid shares the (row, strip of gallery rows) work items, and the LAST block to finish unpacks the winners
// (its atomic ticket orders it after every other block's atomicMin).
constexpr int kStrip = 8;          // gallery rows per warp work item in the exhaustive evaluation
constexpr int kFinThreads = 256;
constexpr int kMaxFlagList = 2048;

__global__ void __launch_bounds__(kFinThreads)
k_finish(const RerankParams p, int negate, uint64_t *out_packed, float *out_score, int64_t *out_idx)
{
    const unsigned mode = p.ctr->overflow;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (mode == 0) {
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * kFinThreads + tid; i < p.P; i += static_cast<int64_t>(gridDim.x) * kFinThreads)
            finalize_row(p.best, i, negate, out_packed, out_score, out_idx);
        return;
    }
    __shared__ int32_t s_list[kMaxFlagList];
    __shared__ int s_wcnt[kFinThreads / 32];
    __shared__ int s_n;
    __shared__ unsigned int s_ticket;
    int64_t nrows = p.P;                                 // mode 2, or too many flagged rows: every row is a candidate row
    bool listed = false;
    if (mode == 1) {
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (int64_t i0 = 0; i0 < p.P; i0 += kFinThreads) {          // ordered compaction, identical in every block
            const int64_t i = i0 + tid;
            const bool f = i < p.P && p.rowflag[i] != 0;
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_wcnt[warp] = __popc(m);
            __syncthreads();
            int base = s_n;
            for (int w = 0; w < warp; ++w) base += s_wcnt[w];
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (f && pos < kMaxFlagList) s_list[pos] = static_cast<int32_t>(i);
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < kFinThreads / 32; ++w) t += s_wcnt[w]; s_n += t; }
            __syncthreads();
        }
        if (s_n <= kMaxFlagList) { nrows = s_n; listed = true; }
        if (blockIdx.x == 0 && tid == 0) p.ctr->n_flag_rows = static_cast<unsigned>(s_n);
    } else if (blockIdx.x == 0 && tid == 0) p.ctr->n_flag_rows = static_cast<unsigned>(p.P);
    const int64_t nstrips = (p.G + kStrip - 1) / kStrip;
    const int64_t nw = static_cast<int64_t>(gridDim.x) * (kFinThreads / 32);
    for (int64_t w = static_cast<int64_t>(blockIdx.x) * (kFinThreads / 32) + warp; w < nrows * nstrips; w += nw) {
        const int64_t row = listed ? s_list[w / nstrips] : w / nstrips;
        if (!listed && mode == 1 && p.rowflag[row] == 0) continue;
        const int64_t g0 = (w % nstrips) * kStrip, g1 = min(g0 + kStrip, p.G);
        unsigned long long loc = ~0ull;
        for (int64_t g = g0; g < g1; ++g) {
            const float t = exact_score(p, row, g, lane);
            const unsigned long long v = pack_score_idx(t, static_cast<uint32_t>(p.offset + g));
            loc = v < loc ? v : loc;
        }
        if (lane == 0) atomicMin(p.best + row, loc);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&p.ctr->done_blocks, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    for (int64_t i = tid; i < p.P; i += kFinThreads)
        finalize_row(p.best, i, negate, out_packed, out_score, out_idx);
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
// Experiment knobs, read from the environment ONCE per process.  None of them changes results: unit order, gallery
// tiles per unit, seed pass, epilogue warps, issuer warps, and the EOSVR_EXP measurement bits
// 16 (cycle accounting of the screening kernel) and 64 (phase timing of the re-rank).  The result-destroying
// timing modes (EOSVR_EXP bits 1, 2, 4, 32) exist only in builds with -DEOSVR_EXPERIMENTS (tools/exp_perf.sh).
struct Tunables {
    int order, tpu, seed, ew, issuers, exp, aligned, rr;
};
static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
const Tunables &tunables()
{
    static const Tunables t = [] {
        Tunables v;
        v.order = env_int("EOSVR_ORDER", -1);          // -1: chosen per call (launch_screen)
        v.tpu = env_int("EOSVR_TPU", 0);
        v.seed = env_int("EOSVR_SEED", 1);
        v.ew = env_int("EOSVR_EW", 0);
        v.issuers = env_int("EOSVR_ISSUERS", 0);
        v.exp = env_int("EOSVR_EXP", 0);
        v.aligned = env_int("EOSVR_ALIGNED", 1);       // 0: never use the episode-aligned epilogue
        v.rr = env_int("EOSVR_RR", -1);                // re-rank kernel: 0 block-per-rows, 1 warp-per-row, -1 chosen per call
#ifndef EOSVR_EXPERIMENTS
        v.exp &= (16 | 64);
#endif
        return v;
    }();
    return t;
}

static std::mutex g_dev_mu;
static DeviceState g_dev[64];

static int device_state(DeviceState **out)
{
    int dev = 0;
    EOSVR_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return EOSVR_EINVAL; }
    std::lock_guard<std::mutex> lock(g_dev_mu);
    DeviceState &d = g_dev[dev];
    if (d.num_sms == 0) EOSVR_CUDA(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
    *out = &d;
    return EOSVR_OK;
}

template <int EW, bool DIAG, int AL>
static int launch_screen_t(DeviceState *ds, const CUtensorMap &tmA, const CUtensorMap &tmB, const ScreenParams &sp, cudaStream_t st)
{
    auto kern = k_match_screen<EW, DIAG, AL>;
    constexpr size_t smem = screen_smem<EW>();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(screen_threads(EW), 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    int &maxcl = ds->max_clusters[(AL > 0 ? 2 : 0) + (DIAG ? 1 : 0)][EW == 16 ? 1 : 0];
    {
        std::lock_guard<std::mutex> lock(g_dev_mu);
        if (maxcl == 0) {
            EOSVR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            cfg.gridDim = dim3(static_cast<unsigned>(ds->num_sms / 2 * 2), 1, 1);
            int n = 0;
            EOSVR_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
            if (n < 1) { set_error("k_match_screen<%d>: no CTA pair can be resident", EW); return EOSVR_ECUDA; }
            maxcl = n;
        }
    }
    const int64_t ncl = sp.n_units < maxcl ? sp.n_units : maxcl;
    cfg.gridDim = dim3(static_cast<unsigned>(ncl * 2), 1, 1);
    EOSVR_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, sp));
    return EOSVR_OK;
}

// Epilogue warps: 16 when the rows are short (the epilogue, not the tensor pipe, bounds a tile below ~1024
// dimensions), 8 otherwise.
static int choose_ew(int32_t Dp)
{
    const int e = tunables().ew;
    if (e == 8 || e == 16) return e;
    return Dp <= 1024 ? 16 : 8;
}

struct ScreenView {          // the screening copy of the gallery a launch reads (one per metric)
    const float *gnorm;
    const CUtensorMap *tmapA, *tmapSeed;
};

static int launch_screen(DeviceState *ds, const eosvr_gallery *g, const ScreenView &sv, eosvr_workspace *ws, const MatchPlan &pl,
                         bool seed, const CUtensorMap &tmB, int64_t gallery_tiles, int64_t g_stride, int64_t P, float w,
                         cudaStream_t st)
{
    const Tunables &tn = tunables();
    ScreenParams sp;
    sp.gnorm = sv.gnorm; sp.G = g->G; sp.KB = g->Dp / kBK; sp.BN = pl.BN; sp.NT = static_cast<int32_t>(pl.NT);
    sp.GT = static_cast<int32_t>(gallery_tiles);
    const int seed_mode = seed ? 1 : 0;
    const int64_t total_tiles = static_cast<int64_t>(sp.NT) * sp.GT;
    int64_t tpu = total_tiles / (static_cast<int64_t>(ds->num_sms / 2) * 6);
    if (tpu < 1) tpu = 1;
    if (tpu > 16) tpu = 16;
    if (tpu > sp.GT) tpu = sp.GT;
    if (tn.tpu > 0 && !seed_mode) tpu = tn.tpu < sp.GT ? tn.tpu : sp.GT;
    sp.TPU = static_cast<int32_t>(tpu);
    sp.n_chunks = static_cast<int32_t>((sp.GT + tpu - 1) / tpu);
    if (static_cast<int64_t>(sp.n_chunks) * sp.NT > 0x7FFFFFFFll) { set_error("match: too many work units"); return EOSVR_EINVAL; }
    sp.n_units = sp.n_chunks * sp.NT;
    // Unit order: the operand that is swept repeatedly should be the one that stays in L2.  Chunk-major (0) streams the
    // gallery once and re-reads the probe plan per gallery chunk; probe-tile-major (1) does the opposite.  Measured on
    // B200 (profiles/r02_unit_order.txt): cfg-3 (21 MB of probes, 102 MB gallery) 1.91 -> 1.68 ms and a 1.25 M-row
    // shard of cfg-4 28.9 -> 21.9 ms with chunk-major; cfg-2 (117 MB of probes, 46 MB gallery) is 2 % better the other
    // way round.  Chunk-major also gives every probe row to ONE pair at a time, so thresholds are never stale: half
    // the candidates.
    {
        const int64_t probe_bytes = static_cast<int64_t>(pl.NT) * pl.BN * g->Dp * 2;
        const int64_t gallery_bytes = gallery_tiles * kPairM * static_cast<int64_t>(g->Dp) * 2;
        sp.order = tn.order >= 0 ? tn.order : (probe_bytes <= gallery_bytes ? 0 : 1);
    }
    sp.g_stride = g_stride; sp.seed_mode = seed_mode;
    sp.issuers = ds->issuers > 0 ? ds->issuers : kIssuers;
    sp.rpe = pl.rpe; sp.w = w;
    sp.acc_stages = pl.BN <= kBN3 ? 3 : 2;
    sp.acc_stride = pl.BN <= kBN3 ? kBN3 : kMaxBN;
    sp.na = ws->na; sp.wl = ws->wl; sp.wr = ws->wr; sp.epsd = ws->epsd; sp.rowmap = ws->rowmap;
    sp.gthr = ws->gthr; sp.cand = ws->cand; sp.rowcnt = ws->rowcnt; sp.cand_cap = static_cast<int32_t>(ws->cand_cap);
    sp.ovf = ws->ovf; sp.ovf_cap = static_cast<int32_t>(ws->ovf_cap);
    sp.ctr = ws->counters; sp.rowflag = ws->rowflag;
    sp.idesc = umma_idesc_f16(g->screen_fmt == EOSVR_SCREEN_F16 ? 0 : 1, kPairM, pl.BN);
    sp.dbg = (!seed_mode && ws->dbg && ws->dbg_elems >= P * g->G) ? ws->dbg : nullptr;
    sp.exp_mode = tn.exp;
    const int kid = seed ? EOSVR_KERNEL_SEED : EOSVR_KERNEL_SCREEN;
    { int trc = timing_begin(ws, kid, st); if (trc) return trc; }
    const int ew = choose_ew(g->Dp);
    const bool diag = sp.dbg != nullptr || (tn.exp & 16) != 0;
    const CUtensorMap &tmA = seed ? *sv.tmapSeed : *sv.tmapA;
    int rc;
    // episode-aligned epilogue: whole episodes of kAlign rows per tile, no halo, tap weight uniform inside an episode
    // (the screening-value dump of the tests runs the generic epilogue; cycle accounting exists for both)
    const bool aligned = tn.aligned != 0 && ew == 16 && sp.dbg == nullptr && pl.rpe == kAlign && pl.halo == 0 && pl.R % kAlign == 0 &&
                         P % kAlign == 0;
    if (aligned) rc = diag ? launch_screen_t<16, true, kAlign>(ds, tmA, tmB, sp, st) : launch_screen_t<16, false, kAlign>(ds, tmA, tmB, sp, st);
    else if (ew == 16) rc = diag ? launch_screen_t<16, true, 0>(ds, tmA, tmB, sp, st) : launch_screen_t<16, false, 0>(ds, tmA, tmB, sp, st);
    else rc = diag ? launch_screen_t<8, true, 0>(ds, tmA, tmB, sp, st) : launch_screen_t<8, false, 0>(ds, tmA, tmB, sp, st);
    if (rc) return rc;
    { int trc = timing_end(ws, kid, st); if (trc) return trc; }
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
    DeviceState *ds = nullptr;
    int rc = device_state(&ds);
    if (rc) return rc;
    if (!exact_only && ds->issuers == 0) {                   // first screening call on this device: self-check
        rc = screening_selfcheck(ds, st);
        if (rc) return rc;
    }
    ScreenView sv{g->gnorm, &g->tmapA, &g->tmapSeed};
    const float *scalars = g->scalars;
    if (cosm && !exact_only) {
        rc = build_cosine_copy(g, st);                      // first cosine call builds the normalised copy
        if (rc) return rc;
        EOSVR_CUDA(cudaStreamWaitEvent(st, g->cos->ready, 0));
        sv = ScreenView{g->cos->gnorm, &g->cos->tmapA, &g->cos->tmapSeed};
        scalars = g->cos->scalars;
    }
    const Tunables &tn = tunables();
    MatchPlan pl = make_plan(P, rpe);
    const int64_t ncol = pl.NT * pl.BN;
    if (ncol > ws->cap_rows) {
        set_error("workspace too small: plan needs %lld rows, capacity %lld", (long long)ncol, (long long)ws->cap_rows);
        return EOSVR_EINVAL;
    }
    PlanDev pd{pl.P, pl.rpe, pl.R, pl.halo, pl.BN, pl.NT};
    const int threads = 256;
    const int num_sms = ds->num_sms;
    RowState rs{ws->best, ws->rowflag, ws->rowcnt, ws->gthr};

    RerankParams rp;
    rp.probes = probes; rp.gal = g->feats; rp.gal_dtype = g->dtype; rp.P = P; rp.G = g->G; rp.offset = g->offset;
    rp.D = g->D; rp.rpe = rpe; rp.metric = metric; rp.lam1 = lam1; rp.lam2 = lam2;
    rp.cand = ws->cand; rp.rowcnt = ws->rowcnt; rp.cand_cap = static_cast<int32_t>(ws->cand_cap);
    rp.ctr = ws->counters; rp.gthr = ws->gthr;
    rp.best = ws->best; rp.rowflag = ws->rowflag;
    rp.ovf = ws->ovf; rp.ovf_cap = static_cast<int32_t>(ws->ovf_cap);
    rp.epsd = ws->epsd; rp.wl = ws->wl; rp.wr = ws->wr; rp.planR = pl.R; rp.planBN = pl.BN; rp.planHalo = pl.halo;
    rp.prof = (tn.exp & 64) ? 1 : 0;
    rp.rows_per_block = 1;

    ws->last_tiles = 0;
    ws->last_bn = pl.BN;
    EOSVR_CUDA(cudaMemsetAsync(ws->counters, 0, sizeof(Counters), st));
    if (exact_only) {
        k_reset_exact<<<static_cast<unsigned>((P + threads - 1) / threads), threads, 0, st>>>(ws->counters, rs, P);
        EOSVR_CUDA(cudaGetLastError());
        EOSVR_COUNT_LAUNCH(1);
    } else {
        const unsigned pblocks = static_cast<unsigned>((ncol * 32 + threads - 1) / threads);
        rc = timing_begin(ws, EOSVR_KERNEL_PROBE_PREP, st);
        if (rc) return rc;
#define EOSVR_PROBE_PREP(T16, NORM)                                                                              \
        k_probe_prep<T16, NORM><<<pblocks, threads, 0, st>>>(probes, pd, g->D, g->Dp, lam1 / lam2, scalars,         \
            static_cast<T16 *>(ws->q16), ws->na, ws->epsd, ws->wl, ws->wr, ws->rowmap, rs, ws->counters)
        if (g->screen_fmt == EOSVR_SCREEN_F16) { if (cosm) EOSVR_PROBE_PREP(__half, true); else EOSVR_PROBE_PREP(__half, false); }
        else { if (cosm) EOSVR_PROBE_PREP(__nv_bfloat16, true); else EOSVR_PROBE_PREP(__nv_bfloat16, false); }
#undef EOSVR_PROBE_PREP
        EOSVR_CUDA(cudaGetLastError());
        rc = timing_end(ws, EOSVR_KERNEL_PROBE_PREP, st);
        if (rc) return rc;
        EOSVR_COUNT_LAUNCH(1);

        CUtensorMap tmB;
        rc = encode_tmap_2d(&tmB, ws->q16, g->screen_fmt, static_cast<uint64_t>(ncol),
                            static_cast<uint64_t>(g->Dp), static_cast<uint32_t>(pl.BN / 2), kBK);
        if (rc) return rc;
        const int64_t GT = (g->G + kPairM - 1) / kPairM;   // 256-row tiles of the CTA pair
        // seed pass over a strided sample of the gallery: tightens every probe row's threshold before the
        // full pass so that concurrent CTAs do not flood the candidate lists (EOSVR_SEED=0 skips it: experiments)
        if (tn.seed && g->seed_tiles > 0 && GT > g->seed_tiles) {
            rc = launch_screen(ds, g, sv, ws, pl, true, tmB, tn.seed == 1 ? g->seed_tiles : 1, g->seed_stride, P, lam1 / lam2, st);
            if (rc) return rc;
        }
        rc = launch_screen(ds, g, sv, ws, pl, false, tmB, GT, 1, P, lam1 / lam2, st);
        if (rc) return rc;
        ws->last_tiles = pl.NT * GT;

        // few probe rows (a single episode): one row per block keeps the call's latency low; large batches take 8
        // consecutive rows per block so that neighbouring probe rows are staged once
        int64_t rpb = P / (static_cast<int64_t>(num_sms) * 8);
        rpb = rpb < 1 ? 1 : (rpb > kRrRowsPerBlock ? kRrRowsPerBlock : rpb);
        rp.rows_per_block = static_cast<int32_t>(rpb);
        const int64_t rr_blocks = (P + rpb - 1) / rpb;
        const unsigned rr_grid = static_cast<unsigned>(rr_blocks < static_cast<int64_t>(num_sms) * 32 ? rr_blocks : num_sms * 32);
        const size_t rr_smem = static_cast<size_t>(cosm ? 2 : 4) * g->D * sizeof(float);
        rc = timing_begin(ws, EOSVR_KERNEL_RERANK, st);
        if (rc) return rc;
        if ((g->D & 3) == 0 && rr_smem <= 96 * 1024) {
            {
                std::lock_guard<std::mutex> lock(g_dev_mu);
                if (!ds->rr_attr) {
                    EOSVR_CUDA(cudaFuncSetAttribute(k_rerank_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                    EOSVR_CUDA(cudaFuncSetAttribute(k_rerank_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                    ds->rr_attr = true;
                }
            }
            // warp-per-row where its static specialisation applies (512-element rows: cfg-3 / cfg-4 0.165 -> 0.098 ms);
            // the block-per-rows kernel elsewhere (at D = 2048 the warp kernel is 1.6x slower: four warps share a
            // candidate's 8 KiB row there) -- profiles/r02_ab_rerank_warp.txt
            if (tn.rr == 1 || (tn.rr < 0 && g->D == 512)) {
                // one resident wave: a warp takes several rows and requests a row's counters and candidates a row ahead
                const int64_t wb = (P + kRwWarps - 1) / kRwWarps;
                const unsigned wgrid = static_cast<unsigned>(wb < static_cast<int64_t>(num_sms) * EOSVR_RW_MINBLOCKS ? wb : num_sms * EOSVR_RW_MINBLOCKS);
                const bool bf = g->dtype == EOSVR_BF16, d512 = g->D == 512;
#define EOSVR_RW(C, B, X) k_rerank_warp<C, B, X><<<wgrid, 32 * kRwWarps, 0, st>>>(rp)
                if (cosm) { if (bf) { if (d512) EOSVR_RW(true, true, true); else EOSVR_RW(true, true, false); }
                            else { if (d512) EOSVR_RW(true, false, true); else EOSVR_RW(true, false, false); } }
                else { if (bf) { if (d512) EOSVR_RW(false, true, true); else EOSVR_RW(false, true, false); }
                       else { if (d512) EOSVR_RW(false, false, true); else EOSVR_RW(false, false, false); } }
#undef EOSVR_RW
            }
            else if (cosm) k_rerank_rows<true><<<rr_grid, kRrThreads, rr_smem, st>>>(rp);
            else k_rerank_rows<false><<<rr_grid, kRrThreads, rr_smem, st>>>(rp);
        }
        else k_rerank<<<num_sms * 8, 256, 0, st>>>(rp);
        EOSVR_CUDA(cudaGetLastError());
        rc = timing_end(ws, EOSVR_KERNEL_RERANK, st);
        if (rc) return rc;
        EOSVR_COUNT_LAUNCH(1);
    }
    rc = timing_begin(ws, EOSVR_KERNEL_FINISH, st);
    if (rc) return rc;
    k_finish<<<num_sms * 4, kFinThreads, 0, st>>>(rp, cosm ? 1 : 0, out_packed, out_score, out_idx);
    EOSVR_CUDA(cudaGetLastError());
    rc = timing_end(ws, EOSVR_KERNEL_FINISH, st);
    if (rc) return rc;
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

// -------------------------------------------------------------------------------------------
// Winner rows (multi-GPU exchange helper and input of the splice kernel).
// -------------------------------------------------------------------------------------------
__global__ void k_gather_rows(const void *__restrict__ gal, int gdt, int64_t G, int64_t offset, int D,
                              const int64_t *__restrict__ idx, float *__restrict__ out)
{
    const int64_t p = blockIdx.x;
    const int64_t g = idx[p] - offset;
    const bool own = g >= 0 && g < G;
    const int64_t src = (own ? g : 0) * D;
    float *dst = out + p * D;
    if ((D & 3) == 0) {
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int k = threadIdx.x; k < D / 4; k += blockDim.x) d4[k] = own ? ld_feat4(gal, gdt, src / 4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int k = threadIdx.x; k < D; k += blockDim.x) dst[k] = own ? ld_feat(gal, gdt, src + k) : 0.f;
    }
}

__global__ void k_upcast_bf16(const unsigned short *__restrict__ in, int64_t n, float *__restrict__ out)
{
    const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const uint2 u = *reinterpret_cast<const uint2 *>(in + i);
        *reinterpret_cast<float4 *>(out + i) = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                                                           __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
    } else {
        for (int64_t j = i; j < n; ++j) out[j] = __uint_as_float(static_cast<uint32_t>(in[j]) << 16);
    }
}

int launch_upcast_bf16(const void *in, int64_t n, float *out, cudaStream_t st)
{
    if (n == 0) return EOSVR_OK;
    const int threads = 256;
    const int64_t groups = (n + 3) / 4;
    k_upcast_bf16<<<static_cast<unsigned>((groups + threads - 1) / threads), threads, 0, st>>>(
        static_cast<const unsigned short *>(in), n, out);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

int launch_gather_rows(const eosvr_gallery *g, const int64_t *idx, int64_t P, float *out, cudaStream_t st)
{
    if (P == 0) return EOSVR_OK;
    k_gather_rows<<<static_cast<unsigned>(P), 128, 0, st>>>(g->feats, g->dtype, g->G, g->offset, g->D, idx, out);
    EOSVR_CUDA(cudaGetLastError());
    EOSVR_COUNT_LAUNCH(1);
    return EOSVR_OK;
}

}  // namespace eosvr
