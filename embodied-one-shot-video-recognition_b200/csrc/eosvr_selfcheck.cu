// eosvr_selfcheck.cu -- one-time check, per device, of the hardware behaviour the screening kernel relies on.
//
// k_match_screen lets THREE threads issue tcgen05.mma into the same TMEM accumulator without ordering them against
// each other (one thread cannot keep the tensor pipe fed).  PTX orders the MMAs of one thread
// only; that concurrent accumulation from several issuing threads adds up exactly is an observed property of this
// part (tools/bench_micro/mma_rate.cu), not a documented one.  A lost update would silently drop true winners: the
// exact re-rank only sees the candidates the screening pass hands it.  So the first screening call on a device
// runs a small synthetic match here and compares EVERY screening value t~[P,G] with a CUDA-core float32 evaluation
// of the same quantity.  On a mismatch the library falls back to the single-issuer ordering (slower, documented
// semantics) and checks again; if that fails too, eosvr_match returns an error instead of wrong answers.
// EOSVR_SELFCHECK=0 in the environment skips the check (and keeps the three issuers).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "eosvr_internal.h"

namespace eosvr {

namespace {

constexpr int kScG = 1024, kScP = 96, kScD = 320, kScRpe = 12;   // D = 320: 5 K blocks -> every issuer takes stages of a tile

__device__ __forceinline__ uint32_t sc_hash(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// rows of norm ~0.7 around a few shared directions, so that near and far pairs both occur
__global__ void k_sc_fill(float *x, int rows, int D, uint32_t seed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * D) return;
    const int r = i / D, k = i % D;
    const float c = static_cast<float>(sc_hash(seed * 7919u + (r % 5) * 977u + k) & 0xFFFF) / 65536.f - 0.5f;
    const float n = static_cast<float>(sc_hash(seed * 104729u + i) & 0xFFFF) / 65536.f - 0.5f;
    x[i] = (c + 0.35f * n) * (1.4f / sqrtf(static_cast<float>(D)) * 2.45f);
}

// float32 direct-difference evaluation of the smoothed distance in the screening domain: d + w (d_left + d_right)
__global__ void k_sc_ref(const float *A, const float *B, int P, int G, int D, int rpe, float w, float *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * G) return;
    const int p = i / G, g = i % G;
    const int r = p % rpe;
    float d[3] = {0.f, 0.f, 0.f};
    for (int o = -1; o <= 1; ++o) {
        const int q = p + o;
        if ((o < 0 && r == 0) || (o > 0 && (r + 1 == rpe || q >= P))) continue;
        float s = 0.f;
        for (int k = 0; k < D; ++k) { const float e = A[q * D + k] - B[g * D + k]; s = fmaf(e, e, s); }
        d[o + 1] = sqrtf(s);
    }
    out[i] = d[1] + w * (d[0] + d[2]);
}

struct Buffers {
    float *gal = nullptr, *probes = nullptr, *dump = nullptr, *ref = nullptr;
    uint64_t *packed = nullptr;
    eosvr_gallery_t *g = nullptr;
    eosvr_workspace_t *ws = nullptr;
    ~Buffers()
    {
        if (ws) eosvr_workspace_destroy(ws);
        if (g) eosvr_gallery_destroy(g);
        if (gal) cudaFree(gal);
        if (probes) cudaFree(probes);
        if (dump) cudaFree(dump);
        if (ref) cudaFree(ref);
        if (packed) cudaFree(packed);
    }
};

}  // namespace

int screening_selfcheck(DeviceState *ds, cudaStream_t st)
{
    const char *e = getenv("EOSVR_SELFCHECK");
    const char *ei = getenv("EOSVR_ISSUERS");
    const int forced = ei ? atoi(ei) : 0;
    if (e && atoi(e) == 0) { ds->issuers = (forced == 1) ? 1 : kIssuers; return EOSVR_OK; }
    Buffers b;
    EOSVR_CUDA(cudaMalloc(&b.gal, sizeof(float) * kScG * kScD));
    EOSVR_CUDA(cudaMalloc(&b.probes, sizeof(float) * kScP * kScD));
    EOSVR_CUDA(cudaMalloc(&b.dump, sizeof(float) * kScP * kScG));
    EOSVR_CUDA(cudaMalloc(&b.ref, sizeof(float) * kScP * kScG));
    EOSVR_CUDA(cudaMalloc(&b.packed, sizeof(uint64_t) * kScP));
    k_sc_fill<<<(kScG * kScD + 255) / 256, 256, 0, st>>>(b.gal, kScG, kScD, 11u);
    k_sc_fill<<<(kScP * kScD + 255) / 256, 256, 0, st>>>(b.probes, kScP, kScD, 23u);
    const float lam1 = 0.1f, lam2 = 1.0f;
    k_sc_ref<<<(kScP * kScG + 255) / 256, 256, 0, st>>>(b.probes, b.gal, kScP, kScG, kScD, kScRpe, lam1 / lam2, b.ref);
    EOSVR_CUDA(cudaGetLastError());
    int rc = eosvr_gallery_create(b.gal, kScG, kScD, EOSVR_F32, 0, EOSVR_SCREEN_F16, st, &b.g);
    if (rc) return rc;
    rc = eosvr_workspace_create(kScP, kScD, 0, &b.ws);
    if (rc) return rc;
    std::vector<float> h_ref(static_cast<size_t>(kScP) * kScG), h_dump(h_ref.size());
    EOSVR_CUDA(cudaMemcpyAsync(h_ref.data(), b.ref, sizeof(float) * h_ref.size(), cudaMemcpyDeviceToHost, st));
    const int tries[2] = {(forced == 1) ? 1 : kIssuers, 1};
    double worst = 0.0;
    for (int t = 0; t < 2; ++t) {
        if (t == 1 && tries[0] == 1) break;
        ds->issuers = tries[t];                              // (non-zero: the nested call below does not re-enter)
        EOSVR_CUDA(cudaMemsetAsync(b.dump, 0xFF, sizeof(float) * h_dump.size(), st));      // NaN pattern
        eosvr_workspace_set_debug(b.ws, b.dump, static_cast<int64_t>(h_dump.size()));
        rc = launch_match(b.g, b.ws, b.probes, kScP, kScRpe, EOSVR_METRIC_EUCLID_TEMPORAL, lam1, lam2, false, b.packed,
                          nullptr, nullptr, st);
        if (rc) { ds->issuers = 0; return rc; }
        EOSVR_CUDA(cudaMemcpyAsync(h_dump.data(), b.dump, sizeof(float) * h_dump.size(), cudaMemcpyDeviceToHost, st));
        EOSVR_CUDA(cudaStreamSynchronize(st));
        worst = 0.0;
        for (size_t i = 0; i < h_ref.size(); ++i) {
            const double err = fabs(static_cast<double>(h_dump[i]) - static_cast<double>(h_ref[i]));
            if (!(err <= worst)) worst = (err == err) ? err : 1e30;       // NaN = element never written
        }
        // fp16 operands: |t~ - t| is a few 1e-3 at these norms; one lost K block moves a value by ~0.1 or more
        if (worst < 2e-2) return EOSVR_OK;
    }
    ds->issuers = 0;
    set_error("screening self-check failed on this device: tensor-core values differ from the CUDA-core evaluation by %.3g "
              "(multi-issuer and single-issuer orderings both tried)", worst);
    return EOSVR_ECUDA;
}

}  // namespace eosvr
