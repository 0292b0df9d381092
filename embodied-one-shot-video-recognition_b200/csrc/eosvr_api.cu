// eosvr_api.cu -- the extern "C" surface declared in include/eosvr.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>

#include <nvtx3/nvToolsExt.h>

#include "eosvr_internal.h"

namespace eosvr {

// NVTX range around every compute entry point (header-only NVTX v3: a no-op unless a profiler injects itself).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define EOSVR_RANGE(name) ::eosvr::NvtxRange nvtx_range__(name)

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void count_launches(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D row-major [rows, cols] 16-bit tensor, box = box_rows x box_cols, 128-byte swizzle.
int encode_tmap_2d(CUtensorMap *m, const void *base, int fmt, uint64_t rows, uint64_t cols,
                   uint32_t box_rows, uint32_t box_cols, uint64_t row_stride)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return EOSVR_ECUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * 2 * row_stride};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, fmt == EOSVR_SCREEN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                     2, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return EOSVR_ECUDA; }
    return EOSVR_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int timing_begin(eosvr_workspace *ws, int kernel, cudaStream_t st)
{
    if (!ws || !ws->timing_on) return EOSVR_OK;
    EOSVR_CUDA(cudaEventRecord(ws->ev0[kernel][ws->timing_calls[kernel] % kTimingRing], st));
    return EOSVR_OK;
}
int timing_end(eosvr_workspace *ws, int kernel, cudaStream_t st)
{
    if (!ws || !ws->timing_on) return EOSVR_OK;
    EOSVR_CUDA(cudaEventRecord(ws->ev1[kernel][ws->timing_calls[kernel] % kTimingRing], st));
    ++ws->timing_calls[kernel];
    return EOSVR_OK;
}

static void free_screen_copy(eosvr_screen_copy *c)
{
    if (!c) return;
    if (c->h16) cudaFree(c->h16);
    if (c->gnorm) cudaFree(c->gnorm);
    if (c->scalars) cudaFree(c->scalars);
    if (c->ready) cudaEventDestroy(c->ready);
    delete c;
}

// Build (once, under the handle's mutex) the L2-normalised screening copy the cosine metric reads.  The build
// kernel runs on `st`; later calls on other streams wait for c->ready.
int build_cosine_copy(eosvr_gallery *g, cudaStream_t st)
{
    std::lock_guard<std::mutex> lock(*static_cast<std::mutex *>(g->cos_mutex));
    if (g->cos) return EOSVR_OK;
    eosvr_screen_copy *c = new (std::nothrow) eosvr_screen_copy();
    if (!c) { set_error("out of host memory"); return EOSVR_ENOMEM; }
    memset(c, 0, sizeof(*c));
    const int64_t Gpad = (g->G + kPairM - 1) / kPairM * kPairM;
    if (cudaMalloc(&c->h16, static_cast<size_t>(Gpad) * g->Dp * 2) != cudaSuccess ||
        cudaMalloc(&c->gnorm, static_cast<size_t>(Gpad) * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&c->scalars, 4 * sizeof(float)) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        set_error("cosine screening copy: device allocation failed (%lld rows x %d)", (long long)Gpad, g->Dp);
        free_screen_copy(c);
        return EOSVR_ENOMEM;
    }
    int rc = launch_gallery_prep_cos(g, c, st);
    const uint64_t rows = static_cast<uint64_t>(Gpad), cols = static_cast<uint64_t>(g->Dp);
    const uint64_t srows = static_cast<uint64_t>(g->seed_tiles) * kPairM, sstr = static_cast<uint64_t>(g->seed_stride);
    if (!rc) rc = encode_tmap_2d(&c->tmapA, c->h16, g->screen_fmt, rows, cols, kBM, kBK, 1);
    if (!rc) rc = encode_tmap_2d(&c->tmapSeed, c->h16, g->screen_fmt, srows, cols, kBM, kBK, sstr);
    if (!rc && cudaEventRecord(c->ready, st) != cudaSuccess) { set_error("cudaEventRecord failed"); rc = EOSVR_ECUDA; }
    if (rc) { free_screen_copy(c); return rc; }
    g->cos = c;
    return EOSVR_OK;
}

}  // namespace eosvr

using namespace eosvr;

extern "C" {

int eosvr_version(void) { return EOSVR_VERSION; }

const char *eosvr_last_error(void) { return g_err; }

int eosvr_device_check(void)
{
    int dev = 0;
    EOSVR_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    EOSVR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    EOSVR_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
        return EOSVR_EUNSUPPORTED;
    }
    return EOSVR_OK;
}

int eosvr_gallery_create(const void *d_feats, int64_t G, int32_t D, int32_t dtype, int64_t global_offset,
                         int32_t screen_fmt, void *stream, eosvr_gallery_t **out)
{
    EOSVR_RANGE("eosvr_gallery_create");
    if (!out) { set_error("gallery_create: out is NULL"); return EOSVR_EINVAL; }
    *out = nullptr;
    if (!d_feats || G < 1 || D < 1) { set_error("gallery_create: need d_feats != NULL, G >= 1, D >= 1"); return EOSVR_EINVAL; }
    if (dtype != EOSVR_F32 && dtype != EOSVR_BF16) { set_error("gallery_create: unsupported dtype %d", dtype); return EOSVR_EINVAL; }
    if (screen_fmt != EOSVR_SCREEN_F16 && screen_fmt != EOSVR_SCREEN_BF16) { set_error("gallery_create: bad screen_fmt %d", screen_fmt); return EOSVR_EINVAL; }
    if (global_offset < 0 || global_offset + G > 0xFFFFFFFFll) { set_error("gallery_create: global indices must fit 32 bits"); return EOSVR_EINVAL; }
    if (G > 0x7FFFFF00ll) { set_error("gallery_create: shard too large"); return EOSVR_EINVAL; }
    int rc = eosvr_device_check();
    if (rc) return rc;
    eosvr_gallery *g = new (std::nothrow) eosvr_gallery();
    if (!g) { set_error("out of host memory"); return EOSVR_ENOMEM; }
    memset(g, 0, sizeof(*g));
    g->feats = d_feats;
    g->dtype = dtype;
    g->G = G; g->D = D; g->Dp = (D + kBK - 1) / kBK * kBK;
    // bfloat16 rows screened as bfloat16: the TMA maps read the caller's rows in place (out-of-bounds rows / columns
    // of a box are zero-filled by the hardware), no second copy of the gallery.  Needs 16-byte row pitch.
    g->owns_h16 = !(dtype == EOSVR_BF16 && screen_fmt == EOSVR_SCREEN_BF16 && (D % 8) == 0 &&
                    (reinterpret_cast<uintptr_t>(d_feats) & 15) == 0);
    g->offset = global_offset; g->screen_fmt = screen_fmt;
    cudaGetDevice(&g->device);
    g->cos_mutex = new (std::nothrow) std::mutex();
    if (!g->cos_mutex) { delete g; set_error("out of host memory"); return EOSVR_ENOMEM; }
    const int64_t Gpad = (G + kPairM - 1) / kPairM * kPairM;
    if (!g->owns_h16) g->h16 = const_cast<void *>(d_feats);
    if ((g->owns_h16 && cudaMalloc(&g->h16, static_cast<size_t>(Gpad) * g->Dp * 2) != cudaSuccess) ||
        cudaMalloc(&g->gnorm, static_cast<size_t>(Gpad) * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&g->scalars, 4 * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        set_error("gallery_create: device allocation failed (%lld rows x %d)", (long long)Gpad, g->Dp);
        eosvr_gallery_destroy(g);
        return EOSVR_ENOMEM;
    }
    rc = launch_gallery_prep(g, static_cast<cudaStream_t>(stream));
    // (in place: the map covers exactly the caller's [G, D] rows; boxes reaching beyond read zeros)
    const uint64_t map_rows = g->owns_h16 ? static_cast<uint64_t>(Gpad) : static_cast<uint64_t>(G);
    const uint64_t map_cols = g->owns_h16 ? static_cast<uint64_t>(g->Dp) : static_cast<uint64_t>(D);
    if (!rc) rc = encode_tmap_2d(&g->tmapA, g->h16, screen_fmt, map_rows, map_cols, kBM, kBK, 1);
    // strided seed sample: seed_tiles tiles of rows {0, stride, 2*stride, ...}
    // (an ODD stride so that periodic class layouts of the gallery cannot alias with the sample)
    const int64_t GT = Gpad / kPairM;
    int64_t st_tiles = GT / 24;
    if (st_tiles < 1) st_tiles = 1;
    if (st_tiles > kMaxSeedTiles) st_tiles = kMaxSeedTiles;
    { const char *e = getenv("EOSVR_SEED_TILES"); if (e && atoi(e) > 0 && atoi(e) <= GT) st_tiles = atoi(e); }   // experiments
    g->seed_tiles = static_cast<int32_t>(st_tiles);
    g->seed_stride = Gpad / (st_tiles * kPairM);
    if (g->seed_stride > 1 && (g->seed_stride & 1) == 0) g->seed_stride -= 1;
    {
        // strided view: row r of the map = gallery row r * seed_stride (rows beyond G must not be addressed: clamp)
        uint64_t srows = static_cast<uint64_t>(st_tiles * kPairM);
        if (!g->owns_h16) { const uint64_t fit = (static_cast<uint64_t>(G) - 1) / static_cast<uint64_t>(g->seed_stride) + 1; if (srows > fit) srows = fit; }
        if (!rc) rc = encode_tmap_2d(&g->tmapSeed, g->h16, screen_fmt, srows, map_cols, kBM, kBK, static_cast<uint64_t>(g->seed_stride));
    }
    if (rc) { eosvr_gallery_destroy(g); return rc; }
    *out = g;
    return EOSVR_OK;
}

int eosvr_gallery_destroy(eosvr_gallery_t *g)
{
    if (!g) return EOSVR_OK;
    if (g->h16 && g->owns_h16) cudaFree(g->h16);
    if (g->gnorm) cudaFree(g->gnorm);
    if (g->scalars) cudaFree(g->scalars);
    free_screen_copy(g->cos);
    delete static_cast<std::mutex *>(g->cos_mutex);
    delete g;
    return EOSVR_OK;
}

int eosvr_gallery_rows(const eosvr_gallery_t *g, int64_t *G, int32_t *D, int64_t *global_offset)
{
    if (!g) { set_error("gallery_rows: NULL handle"); return EOSVR_EINVAL; }
    if (G) *G = g->G;
    if (D) *D = g->D;
    if (global_offset) *global_offset = g->offset;
    return EOSVR_OK;
}

int eosvr_gallery_info(const eosvr_gallery_t *g, int32_t *dtype, int32_t *owns_screen_copy)
{
    if (!g) { set_error("gallery_info: NULL handle"); return EOSVR_EINVAL; }
    if (dtype) *dtype = g->dtype;
    if (owns_screen_copy) *owns_screen_copy = g->owns_h16;
    return EOSVR_OK;
}

int eosvr_upcast_bf16(const void *d_in, int64_t n, float *d_out, void *stream)
{
    EOSVR_RANGE("eosvr_upcast_bf16");
    if (n < 0 || (n > 0 && (!d_in || !d_out))) { set_error("upcast_bf16: bad arguments"); return EOSVR_EINVAL; }
    if ((reinterpret_cast<uintptr_t>(d_in) & 7) || (reinterpret_cast<uintptr_t>(d_out) & 15)) { set_error("upcast_bf16: d_in must be 8-byte and d_out 16-byte aligned"); return EOSVR_EINVAL; }
    return launch_upcast_bf16(d_in, n, d_out, static_cast<cudaStream_t>(stream));
}

int eosvr_workspace_create(int64_t max_probe_rows, int32_t D, int64_t cand_capacity, eosvr_workspace_t **out)
{
    if (!out) { set_error("workspace_create: out is NULL"); return EOSVR_EINVAL; }
    *out = nullptr;
    if (max_probe_rows < 1 || D < 1 || cand_capacity < 0) { set_error("workspace_create: bad sizes"); return EOSVR_EINVAL; }
    if (max_probe_rows > 0x7FFFFFF0ll) { set_error("workspace_create: max_probe_rows too large"); return EOSVR_EINVAL; }
    int rc = eosvr_device_check();
    if (rc) return rc;
    eosvr_workspace *ws = new (std::nothrow) eosvr_workspace();
    if (!ws) { set_error("out of host memory"); return EOSVR_ENOMEM; }
    memset(ws, 0, sizeof(*ws));
    ws->maxP = max_probe_rows; ws->D = D; ws->Dp = (D + kBK - 1) / kBK * kBK;
    ws->cap_rows = max_probe_rows + max_probe_rows / 4 + 4 * kMaxBN;
    ws->cand_cap = cand_capacity ? (cand_capacity + max_probe_rows - 1) / max_probe_rows : 128;   // per probe row
    if (ws->cand_cap < 32) ws->cand_cap = 32;
    if (ws->cand_cap > 4096) ws->cand_cap = 4096;
    cudaGetDevice(&ws->device);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_q = carve(static_cast<size_t>(ws->cap_rows) * ws->Dp * 2);
    const size_t o_na = carve(ws->cap_rows * 4), o_wl = carve(ws->cap_rows * 4), o_wr = carve(ws->cap_rows * 4);
    const size_t o_ep = carve(ws->cap_rows * 4), o_rm = carve(ws->cap_rows * 4);
    const size_t o_thr = carve(ws->maxP * 4), o_best = carve(ws->maxP * 8), o_rf = carve(ws->maxP * 4);
    const size_t o_rc = carve(ws->maxP * 4), o_ix = carve(ws->maxP * 8);
    const size_t o_cd = carve(static_cast<size_t>(ws->maxP) * ws->cand_cap * sizeof(Cand)), o_ct = carve(sizeof(Counters));
    // shared spill-over buffer of full row lists (16 B per entry): large, because running out of it sends rows to the
    // exhaustive evaluation, which costs G x D float64 operations per row (53 s per step on a 10 M-row gallery)
    // (256 entries per probe row or half the row lists' total, whichever is larger; 1 Mi .. 16 Mi entries)
    ws->ovf_cap = ws->maxP * 256 > ws->maxP * ws->cand_cap / 2 ? ws->maxP * 256 : ws->maxP * ws->cand_cap / 2;
    if (ws->ovf_cap < (1ll << 20)) ws->ovf_cap = 1ll << 20;
    if (ws->ovf_cap > (1ll << 24)) ws->ovf_cap = 1ll << 24;
    const size_t o_ov = carve(static_cast<size_t>(ws->ovf_cap) * sizeof(OvfCand));
    if (cudaMalloc(&ws->slab, off) != cudaSuccess) {
        cudaGetLastError();
        set_error("workspace_create: device allocation of %zu bytes failed", off);
        delete ws;
        return EOSVR_ENOMEM;
    }
    cudaMemset(ws->slab, 0, off);
    char *b = static_cast<char *>(ws->slab);
    ws->q16 = b + o_q;
    ws->na = reinterpret_cast<float *>(b + o_na); ws->wl = reinterpret_cast<float *>(b + o_wl);
    ws->wr = reinterpret_cast<float *>(b + o_wr);
    ws->epsd = reinterpret_cast<float *>(b + o_ep); ws->rowmap = reinterpret_cast<int32_t *>(b + o_rm);
    ws->gthr = reinterpret_cast<unsigned int *>(b + o_thr); ws->best = reinterpret_cast<unsigned long long *>(b + o_best);
    ws->rowflag = reinterpret_cast<int32_t *>(b + o_rf);
    ws->rowcnt = reinterpret_cast<unsigned int *>(b + o_rc); ws->cand = reinterpret_cast<Cand *>(b + o_cd);
    ws->counters = reinterpret_cast<Counters *>(b + o_ct);
    ws->idx_scratch = reinterpret_cast<int64_t *>(b + o_ix);
    ws->ovf = reinterpret_cast<OvfCand *>(b + o_ov);
    *out = ws;
    return EOSVR_OK;
}

int eosvr_workspace_destroy(eosvr_workspace_t *ws)
{
    if (!ws) return EOSVR_OK;
    for (int k = 0; k < EOSVR_KERNEL_COUNT; ++k)
        for (int i = 0; i < kTimingRing; ++i) {
            if (ws->ev0[k][i]) cudaEventDestroy(ws->ev0[k][i]);
            if (ws->ev1[k][i]) cudaEventDestroy(ws->ev1[k][i]);
        }
    if (ws->slab) cudaFree(ws->slab);
    delete ws;
    return EOSVR_OK;
}

// Test hook (not part of the reference-facing surface): dump the screening values t~[P,G] of the
// next eosvr_match calls into a caller-owned device buffer (NULL disables).
int eosvr_workspace_set_debug(eosvr_workspace_t *ws, float *d_dump, int64_t elems)
{
    if (!ws) { set_error("set_debug: NULL workspace"); return EOSVR_EINVAL; }
    ws->dbg = d_dump; ws->dbg_elems = d_dump ? elems : 0;
    return EOSVR_OK;
}

int eosvr_workspace_set_timing(eosvr_workspace_t *ws, int32_t on)
{
    if (!ws) { set_error("set_timing: NULL workspace"); return EOSVR_EINVAL; }
    if (on && !ws->ev0[0][0]) {
        for (int k = 0; k < EOSVR_KERNEL_COUNT; ++k)
            for (int i = 0; i < kTimingRing; ++i) {
                EOSVR_CUDA(cudaEventCreate(&ws->ev0[k][i]));
                EOSVR_CUDA(cudaEventCreate(&ws->ev1[k][i]));
            }
    }
    ws->timing_on = on ? 1 : 0;
    for (int k = 0; k < EOSVR_KERNEL_COUNT; ++k) ws->timing_calls[k] = 0;
    return EOSVR_OK;
}

int eosvr_workspace_kernel_ms(eosvr_workspace_t *ws, int32_t kernel, double *sum_ms, int64_t *calls)
{
    if (!ws || !sum_ms || !calls || kernel < 0 || kernel >= EOSVR_KERNEL_COUNT) { set_error("kernel_ms: bad argument"); return EOSVR_EINVAL; }
    const int64_t n = ws->timing_calls[kernel] < kTimingRing ? ws->timing_calls[kernel] : kTimingRing;
    double tot = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        float ms = 0.f;
        EOSVR_CUDA(cudaEventSynchronize(ws->ev1[kernel][i]));
        EOSVR_CUDA(cudaEventElapsedTime(&ms, ws->ev0[kernel][i], ws->ev1[kernel][i]));
        tot += ms;
    }
    *sum_ms = tot; *calls = n;
    return EOSVR_OK;
}

int eosvr_workspace_screen_ms(eosvr_workspace_t *ws, double *sum_ms, int64_t *calls)
{
    return eosvr_workspace_kernel_ms(ws, EOSVR_KERNEL_SCREEN, sum_ms, calls);
}

uint64_t eosvr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int eosvr_plan(int64_t P, int32_t rows_per_episode, int64_t out[4])
{
    if (P < 1 || rows_per_episode < 1 || !out) { set_error("plan: bad arguments"); return EOSVR_EINVAL; }
    const MatchPlan pl = make_plan(P, rows_per_episode);
    out[0] = pl.R; out[1] = pl.halo; out[2] = pl.BN; out[3] = pl.NT;
    return EOSVR_OK;
}

static int check_match_args(const eosvr_gallery_t *g, eosvr_workspace_t *ws, const float *d_probes, int64_t P,
                            int32_t rpe, int32_t metric, float lam1, float lam2, uint64_t *d_out_packed)
{
    if (!g || !ws) { set_error("match: NULL handle"); return EOSVR_EINVAL; }
    if (P < 0 || (P > 0 && !d_probes)) { set_error("match: bad probes"); return EOSVR_EINVAL; }
    if (P > ws->maxP) { set_error("match: P=%lld exceeds workspace max_probe_rows=%lld", (long long)P, (long long)ws->maxP); return EOSVR_EINVAL; }
    if (ws->D != g->D) { set_error("match: workspace D=%d != gallery D=%d", ws->D, g->D); return EOSVR_EINVAL; }
    if (rpe < 1) { set_error("match: rows_per_episode must be >= 1"); return EOSVR_EINVAL; }
    if (metric != EOSVR_METRIC_EUCLID_TEMPORAL && metric != EOSVR_METRIC_COSINE) { set_error("match: unsupported metric %d", metric); return EOSVR_EINVAL; }
    if (metric == EOSVR_METRIC_EUCLID_TEMPORAL && (!(lam2 > 0.f) || !(lam1 >= 0.f))) { set_error("match: need lam2 > 0 and lam1 >= 0"); return EOSVR_EINVAL; }
    if (P > 0 && !d_out_packed) { set_error("match: d_out_packed is required"); return EOSVR_EINVAL; }
    int dev = -1;
    EOSVR_CUDA(cudaGetDevice(&dev));
    if (dev != g->device || dev != ws->device) {
        set_error("match: current device %d, gallery on device %d, workspace on device %d -- make the handles' device current",
                  dev, g->device, ws->device);
        return EOSVR_EINVAL;
    }
    return EOSVR_OK;
}

int eosvr_match(const eosvr_gallery_t *g, eosvr_workspace_t *ws, const float *d_probes, int64_t P,
                int32_t rows_per_episode, int32_t metric, float lam1, float lam2, uint64_t *d_out_packed,
                float *d_out_score, int64_t *d_out_idx, void *stream)
{
    EOSVR_RANGE("eosvr_match");
    int rc = check_match_args(g, ws, d_probes, P, rows_per_episode, metric, lam1, lam2, d_out_packed);
    if (rc) return rc;
    // (the handle is logically const: the cosine screening copy is a lazily built cache behind a mutex)
    return launch_match(const_cast<eosvr_gallery_t *>(g), ws, d_probes, P, rows_per_episode, metric, lam1, lam2, false,
                        d_out_packed, d_out_score, d_out_idx, static_cast<cudaStream_t>(stream));
}

int eosvr_match_exact(const eosvr_gallery_t *g, eosvr_workspace_t *ws, const float *d_probes, int64_t P,
                      int32_t rows_per_episode, int32_t metric, float lam1, float lam2, uint64_t *d_out_packed,
                      float *d_out_score, int64_t *d_out_idx, void *stream)
{
    EOSVR_RANGE("eosvr_match_exact");
    int rc = check_match_args(g, ws, d_probes, P, rows_per_episode, metric, lam1, lam2, d_out_packed);
    if (rc) return rc;
    return launch_match(const_cast<eosvr_gallery_t *>(g), ws, d_probes, P, rows_per_episode, metric, lam1, lam2, true,
                        d_out_packed, d_out_score, d_out_idx, static_cast<cudaStream_t>(stream));
}

int eosvr_match_stats(eosvr_workspace_t *ws, void *stream, int64_t out[8])
{
    if (!ws || !out) { set_error("match_stats: NULL argument"); return EOSVR_EINVAL; }
    Counters c;
    EOSVR_CUDA(cudaMemcpyAsync(&c, ws->counters, sizeof(c), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    EOSVR_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    out[0] = static_cast<int64_t>(c.cand_count);
    out[1] = static_cast<int64_t>(c.n_exact);
    out[2] = c.n_flag_rows;
    out[3] = ws->cand_cap * ws->maxP;
    out[4] = ws->last_tiles;
    out[5] = ws->last_bn;
    out[6] = static_cast<int64_t>(c.n_unsafe);
    out[7] = c.ovf_count;
    return EOSVR_OK;
}

int eosvr_match_stats_ex(eosvr_workspace_t *ws, void *stream, int64_t *out, int32_t n)
{
    if (!ws || !out || n < 0) { set_error("match_stats_ex: bad argument"); return EOSVR_EINVAL; }
    Counters c;
    EOSVR_CUDA(cudaMemcpyAsync(&c, ws->counters, sizeof(c), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    EOSVR_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    const int64_t v[10] = {static_cast<int64_t>(c.cand_count), static_cast<int64_t>(c.n_exact), c.n_flag_rows,
                           ws->cand_cap * ws->maxP, ws->last_tiles, ws->last_bn, static_cast<int64_t>(c.n_unsafe),
                           c.ovf_count, static_cast<int64_t>(c.n_eval32), static_cast<int64_t>(c.n_seq)};
    for (int i = 0; i < n; ++i) out[i] = i < 10 ? v[i] : 0;
    return EOSVR_OK;
}

int eosvr_workspace_debug_cycles(eosvr_workspace_t *ws, void *stream, int64_t out[8])
{
    if (!ws || !out) { set_error("debug_cycles: NULL argument"); return EOSVR_EINVAL; }
    Counters c;
    EOSVR_CUDA(cudaMemcpyAsync(&c, ws->counters, sizeof(c), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    EOSVR_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    out[0] = static_cast<int64_t>(c.cyc_epi_busy); out[1] = static_cast<int64_t>(c.cyc_epi_wait);
    out[2] = static_cast<int64_t>(c.cyc_mma_wait_full); out[3] = static_cast<int64_t>(c.cyc_mma_wait_acc);
    out[4] = static_cast<int64_t>(c.cyc_prod_wait); out[5] = static_cast<int64_t>(c.cyc_total);
    out[6] = static_cast<int64_t>(c.cyc_epi_pre); out[7] = static_cast<int64_t>(c.cyc_epi_loop);
    return EOSVR_OK;
}

int eosvr_merge_top1(const uint64_t *d_gathered, int32_t nshards, int64_t P, uint64_t *d_out_packed,
                     float *d_out_score, int64_t *d_out_idx, void *stream)
{
    EOSVR_RANGE("eosvr_merge_top1");
    if (nshards < 1 || P < 0 || (P > 0 && !d_gathered)) { set_error("merge_top1: bad arguments"); return EOSVR_EINVAL; }
    return launch_merge(d_gathered, nshards, P, d_out_packed, d_out_score, d_out_idx, static_cast<cudaStream_t>(stream));
}

int eosvr_gather_rows(const eosvr_gallery_t *g, const int64_t *d_idx, int64_t P, float *d_out_rows, void *stream)
{
    EOSVR_RANGE("eosvr_gather_rows");
    if (!g || P < 0 || (P > 0 && (!d_idx || !d_out_rows))) { set_error("gather_rows: bad arguments"); return EOSVR_EINVAL; }
    return launch_gather_rows(g, d_idx, P, d_out_rows, static_cast<cudaStream_t>(stream));
}

int eosvr_splice(const float *d_probes, const float *d_winner_rows, int64_t E, int32_t n, int32_t S, int32_t D,
                 int32_t orig_mode, float *d_out, void *stream)
{
    EOSVR_RANGE("eosvr_splice");
    if (E < 0 || n < 1 || S < 1 || D < 1 || (E > 0 && (!d_probes || !d_winner_rows || !d_out))) { set_error("splice: bad arguments"); return EOSVR_EINVAL; }
    if (orig_mode != EOSVR_ORIG_REF_QUIRK && orig_mode != EOSVR_ORIG_CLIP_MEAN) { set_error("splice: bad orig_mode %d", orig_mode); return EOSVR_EINVAL; }
    if (orig_mode == EOSVR_ORIG_REF_QUIRK && n > n * S) { set_error("splice: internal"); return EOSVR_EINVAL; }
    if (1 + S > 65535) { set_error("splice: S too large"); return EOSVR_EINVAL; }
    return launch_splice(d_probes, d_winner_rows, E, n, S, D, orig_mode, d_out, static_cast<cudaStream_t>(stream));
}

int eosvr_proto_score(const float *d_support, const float *d_support_y, const float *d_query, int64_t E, int32_t R,
                      int32_t Q, int32_t D, int32_t max_proto, float *d_dist, float *d_prob, int64_t *d_pred,
                      int32_t *d_nproto, void *stream)
{
    EOSVR_RANGE("eosvr_proto_score");
    if (E < 0 || D < 1 || (E > 0 && (!d_support || !d_support_y || !d_query))) { set_error("proto_score: bad arguments"); return EOSVR_EINVAL; }
    return launch_proto_score(d_support, d_support_y, d_query, E, R, Q, D, max_proto, d_dist, d_prob, d_pred, d_nproto,
                              static_cast<cudaStream_t>(stream));
}

int eosvr_episode_score(const float *d_probes, const float *d_winner_rows, const eosvr_gallery_t *g,
                        const int64_t *d_idx, const float *d_support_y, const float *d_query, int64_t E, int32_t n,
                        int32_t S, int32_t Q, int32_t D, int32_t orig_mode, int32_t max_proto, float *d_dist,
                        float *d_prob, int64_t *d_pred, int32_t *d_nproto, void *stream)
{
    EOSVR_RANGE("eosvr_episode_score");
    if (E < 0 || D < 1 || (E > 0 && (!d_probes || !d_support_y || !d_query))) { set_error("episode_score: bad arguments"); return EOSVR_EINVAL; }
    if (E == 0) return EOSVR_OK;
    if (!d_winner_rows && !(g && d_idx)) { set_error("episode_score: need d_winner_rows, or a gallery handle and d_idx"); return EOSVR_EINVAL; }
    if (!d_winner_rows && g && g->D != D) { set_error("episode_score: gallery D=%d != D=%d", g->D, D); return EOSVR_EINVAL; }
    if (orig_mode != EOSVR_ORIG_REF_QUIRK && orig_mode != EOSVR_ORIG_CLIP_MEAN) { set_error("episode_score: bad orig_mode %d", orig_mode); return EOSVR_EINVAL; }
    return launch_episode_score(d_probes, d_winner_rows, d_winner_rows ? nullptr : g->feats, d_winner_rows ? 0 : g->dtype, d_winner_rows ? 0 : g->G,
                                d_winner_rows ? 0 : g->offset, nullptr, nullptr, 0, d_idx, d_support_y, d_query, E, n, S,
                                Q, D, orig_mode, max_proto, d_dist, d_prob, d_pred, d_nproto,
                                static_cast<cudaStream_t>(stream));
}

int eosvr_episode_batch(const eosvr_gallery_t *g, eosvr_workspace_t *ws, const float *d_probes,
                        const float *d_support_y, const float *d_query, int64_t E, int32_t n, int32_t S,
                        int32_t Q, int32_t metric, float lam1, float lam2, int32_t orig_mode, int32_t max_proto,
                        uint64_t *d_out_packed, float *d_out_score, int64_t *d_out_idx, float *d_dist,
                        float *d_prob, int64_t *d_pred, int32_t *d_nproto, void *stream)
{
    EOSVR_RANGE("eosvr_episode_batch");
    if (E < 0 || n < 1 || S < 1) { set_error("episode_batch: bad shape"); return EOSVR_EINVAL; }
    const int64_t P = E * n * S;
    int rc = check_match_args(g, ws, d_probes, P, n * S, metric, lam1, lam2, d_out_packed);
    if (rc) return rc;
    if (E == 0) return EOSVR_OK;
    if (!d_support_y || !d_query) { set_error("episode_batch: d_support_y and d_query are required"); return EOSVR_EINVAL; }
    if (orig_mode != EOSVR_ORIG_REF_QUIRK && orig_mode != EOSVR_ORIG_CLIP_MEAN) { set_error("episode_batch: bad orig_mode %d", orig_mode); return EOSVR_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t *idx = d_out_idx ? d_out_idx : ws->idx_scratch;
    rc = launch_match(const_cast<eosvr_gallery_t *>(g), ws, d_probes, P, n * S, metric, lam1, lam2, false, d_out_packed,
                      d_out_score, idx, st);
    if (rc) return rc;
    return launch_episode_score(d_probes, nullptr, g->feats, g->dtype, g->G, g->offset, nullptr, nullptr, 0, idx, d_support_y, d_query, E,
                                n, S, Q, g->D, orig_mode, max_proto, d_dist, d_prob, d_pred, d_nproto, st, ws);
}

int eosvr_episode_score_sharded(const float *d_probes, const void *const *d_shard_bases, int32_t shard_dtype,
                                const int64_t *d_shard_begin, int32_t nshards, const int64_t *d_idx, const float *d_support_y, const float *d_query,
                                int64_t E, int32_t n, int32_t S, int32_t Q, int32_t D, int32_t orig_mode,
                                int32_t max_proto, float *d_dist, float *d_prob, int64_t *d_pred, int32_t *d_nproto,
                                void *stream)
{
    EOSVR_RANGE("eosvr_episode_score_sharded");
    if (E < 0 || D < 1 || nshards < 1 || nshards > 64 || !d_shard_bases || !d_shard_begin ||
        (E > 0 && (!d_probes || !d_support_y || !d_query || !d_idx))) { set_error("episode_score_sharded: bad arguments"); return EOSVR_EINVAL; }
    if (orig_mode != EOSVR_ORIG_REF_QUIRK && orig_mode != EOSVR_ORIG_CLIP_MEAN) { set_error("episode_score_sharded: bad orig_mode %d", orig_mode); return EOSVR_EINVAL; }
    int rc = eosvr_device_check();
    if (rc) return rc;
    if (shard_dtype != EOSVR_F32 && shard_dtype != EOSVR_BF16) { set_error("episode_score_sharded: bad shard_dtype %d", shard_dtype); return EOSVR_EINVAL; }
    return launch_episode_score(d_probes, nullptr, nullptr, shard_dtype, 0, 0, d_shard_bases, d_shard_begin, nshards, d_idx,
                                d_support_y, d_query, E, n, S, Q, D, orig_mode, max_proto, d_dist, d_prob, d_pred,
                                d_nproto, static_cast<cudaStream_t>(stream));
}

int eosvr_temporal_smooth(const double *d_dist64, int64_t P, int64_t G, int32_t rows_per_episode, float lam1,
                          float lam2, float *d_out, void *stream)
{
    EOSVR_RANGE("eosvr_temporal_smooth");
    if (P < 0 || G < 0 || rows_per_episode < 1 || (P * G > 0 && (!d_dist64 || !d_out))) { set_error("temporal_smooth: bad arguments"); return EOSVR_EINVAL; }
    int rc = eosvr_device_check();
    if (rc) return rc;
    return launch_temporal_smooth(d_dist64, P, G, rows_per_episode, lam1, lam2, d_out, static_cast<cudaStream_t>(stream));
}

int eosvr_cosine_predict(const float *d_support, const float *d_query, int64_t E, int32_t R, int32_t Q, int32_t D,
                         float *d_sim, int64_t *d_best, void *stream)
{
    EOSVR_RANGE("eosvr_cosine_predict");
    if (E < 0 || R < 1 || Q < 1 || D < 1 || (E > 0 && (!d_support || !d_query || !d_best))) { set_error("cosine_predict: bad arguments"); return EOSVR_EINVAL; }
    int rc = eosvr_device_check();
    if (rc) return rc;
    return launch_cosine_predict(d_support, d_query, E, R, Q, D, d_sim, d_best, static_cast<cudaStream_t>(stream));
}

int eosvr_segment_features(const float *d_frames, int64_t N, int32_t seg_len, int32_t D, int32_t l2, float *d_out,
                           void *stream)
{
    EOSVR_RANGE("eosvr_segment_features");
    if (N < 0 || D < 1 || (N > 0 && (!d_frames || !d_out))) { set_error("segment_features: bad arguments"); return EOSVR_EINVAL; }
    return launch_segment_features(d_frames, N, seg_len, D, l2, d_out, static_cast<cudaStream_t>(stream));
}

int eosvr_clip_features(const float *d_frames, int64_t N, int32_t F, int32_t D, const int32_t *d_nframes, int32_t l2,
                        float *d_out, void *stream)
{
    EOSVR_RANGE("eosvr_clip_features");
    if (N < 0 || F < 1 || D < 1 || (N > 0 && (!d_frames || !d_out))) { set_error("clip_features: bad arguments"); return EOSVR_EINVAL; }
    int rc = eosvr_device_check();
    if (rc) return rc;
    return launch_clip_features(d_frames, N, F, D, d_nframes, l2, d_out, static_cast<cudaStream_t>(stream));
}

int eosvr_take_rows(const float *d_src, int64_t n_src, int64_t row_elems, const int64_t *d_idx, int64_t n, float *d_out,
                    void *stream)
{
    EOSVR_RANGE("eosvr_take_rows");
    if (n < 0 || n_src < 1 || row_elems < 1 || (n > 0 && (!d_src || !d_idx || !d_out))) { set_error("take_rows: bad arguments"); return EOSVR_EINVAL; }
    int rc = eosvr_device_check();
    if (rc) return rc;
    return launch_take_rows(d_src, n_src, row_elems, d_idx, n, d_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
