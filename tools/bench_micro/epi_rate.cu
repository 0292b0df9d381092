// epi_rate.cu -- microbenchmark of the screening epilogue ALONE (no TMA, no MMA): how many SM cycles one
// 128-row x BN-column accumulator tile costs for several formulations of the per-element arithmetic and for
// 8 / 12 / 16 epilogue warps.  TMEM holds arbitrary finite values; no element ever passes the threshold, so
// this is the steady-state fast path of k_match_screen.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o epi_rate epi_rate.cu && ./epi_rate
// Variants:
//   0  round-1 code shape: four per-column float arrays behind GENERIC pointers (LD.E), x = fma(-2,acc,nb)+na,
//      guard min, sqrt, two 3-register FFMA taps, FSETP chain
//   1  same arithmetic, arrays addressed as shared memory (LDS)
//   2  nb pre-folded into the accumulator: x = fma(acc,-2,na)
//   3  variant 2 with packed f32x2 arithmetic on interleaved streams (column pairs (2k,2k+1) are two independent
//      probe streams, so temporal neighbours are whole register pairs)
//   4  norms folded into the MMA (x = -2*acc): sqrt, taps, FSETP, guard only
//   5  variant 4 with packed f32x2 taps
//   6  TMEM reads only (LDTM.x16 + one use per register pair)
//   7  MUFU.SQRT only on the loaded values
//   8  variant 3 with one 16-byte record per column pair {na2, wl2|wr2 ...} -> two LDS.128 per pair ... (see code)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../embodied-one-shot-video-recognition_b200/csrc/eosvr_ptx.cuh"
using namespace eosvr::ptx;

struct Args { int variant, warps, iters, BN; };

constexpr float kBig = 1.0e30f;

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

struct Smem {
    float na[256], wl[256], wr[256], thr[256];
    uint32_t tbase;
};

// ---- one 16-column chunk, scalar formulations -----------------------------------------------------------------
// MODE 0/1: x = fma(-2,acc,nb) + na;  MODE 2: x = fma(acc,-2,na);  MODE 4: x = -2*acc (folded; sqrt(|acc|) stands in)
template <int MODE>
__device__ __forceinline__ bool chunk_scalar(const uint32_t (&v)[16], const float vnext0, const bool hasn, float &dprev,
                                             const float nb, const float *na, const float *wl, const float *wr,
                                             const float *thr, const int c0, const float xfloor, const float dfloor)
{
    float d[16];
    float minx = kBig;
    const float4 *na4 = reinterpret_cast<const float4 *>(na + c0);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
        float aa[4] = {0.f, 0.f, 0.f, 0.f};
        if (MODE != 4) { const float4 a = na4[j4]; aa[0] = a.x; aa[1] = a.y; aa[2] = a.z; aa[3] = a.w; }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = j4 * 4 + jj;
            float x;
            if (MODE <= 1) x = fmaf(-2.f, __uint_as_float(v[j]), nb) + aa[jj];
            else if (MODE == 2) x = fmaf(__uint_as_float(v[j]), -2.f, aa[jj]);
            else x = __uint_as_float(v[j]);
            minx = fminf(minx, x);
            d[j] = sqrt_approx(fabsf(x));
        }
    }
    float dn = kBig;
    if (hasn) {
        float x;
        if (MODE <= 1) x = fmaf(-2.f, vnext0, nb) + na[c0 + 16];
        else if (MODE == 2) x = fmaf(vnext0, -2.f, na[c0 + 16]);
        else x = vnext0;
        dn = sqrt_approx(fabsf(x));
    }
    const float dprev_in = dprev;
    dprev = d[15];
    bool any = false;
    const float4 *wl4 = reinterpret_cast<const float4 *>(wl + c0);
    const float4 *wr4 = reinterpret_cast<const float4 *>(wr + c0);
    const float4 *th4 = reinterpret_cast<const float4 *>(thr + c0);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
        const float4 l = wl4[j4], r = wr4[j4], th = th4[j4];
        const float ll[4] = {l.x, l.y, l.z, l.w}, rr[4] = {r.x, r.y, r.z, r.w}, tt[4] = {th.x, th.y, th.z, th.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = j4 * 4 + jj;
            const float dl = j ? d[j - 1] : dprev_in;
            const float dr = (j < 15) ? d[j + 1] : dn;
            const float t = fmaf(ll[jj], dl, fmaf(rr[jj], dr, d[j]));
            any |= (t <= tt[jj]);
        }
    }
    const bool guard = (minx < xfloor) || (fminf(dprev_in, dn) < dfloor);
    return any || guard;
}

// ---- packed formulation: the chunk's 16 TMEM columns are 8 temporal positions of two interleaved streams --------
// per-column arrays are stored in TMEM-column order, so a float4 = two consecutive pairs.
// MODE 2: x = fma2(acc, -2, na);  MODE 4: x = acc (norms folded into the MMA)
template <int MODE>
__device__ __forceinline__ bool chunk_packed(const uint32_t (&v)[16], const uint32_t vn0, const uint32_t vn1, const bool hasn,
                                             uint64_t &Dprev, const float *na, const float *wl, const float *wr,
                                             const float *thr, const int c0, const float xfloor, const float dfloor)
{
    uint64_t D[8];
    float minx = kBig;
    const uint64_t m2 = pack2(-2.f, -2.f);
    const float4 *na4 = reinterpret_cast<const float4 *>(na + c0);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        float x0, x1, x2, x3;
        if (MODE == 2) {
            const float4 a = na4[k2];
            const uint64_t X0 = fma2(pack2(__uint_as_float(v[4 * k2]), __uint_as_float(v[4 * k2 + 1])), m2, pack2(a.x, a.y));
            const uint64_t X1 = fma2(pack2(__uint_as_float(v[4 * k2 + 2]), __uint_as_float(v[4 * k2 + 3])), m2, pack2(a.z, a.w));
            unpack2(X0, x0, x1); unpack2(X1, x2, x3);
        } else {
            x0 = __uint_as_float(v[4 * k2]); x1 = __uint_as_float(v[4 * k2 + 1]);
            x2 = __uint_as_float(v[4 * k2 + 2]); x3 = __uint_as_float(v[4 * k2 + 3]);
        }
        minx = fminf(fminf(minx, x0), x1);
        minx = fminf(fminf(minx, x2), x3);
        D[2 * k2] = pack2(sqrt_approx(fabsf(x0)), sqrt_approx(fabsf(x1)));
        D[2 * k2 + 1] = pack2(sqrt_approx(fabsf(x2)), sqrt_approx(fabsf(x3)));
    }
    uint64_t Dn = pack2(kBig, kBig);
    if (hasn) {
        float x0, x1;
        if (MODE == 2) {
            const float2 a = *reinterpret_cast<const float2 *>(na + c0 + 16);
            unpack2(fma2(pack2(__uint_as_float(vn0), __uint_as_float(vn1)), m2, pack2(a.x, a.y)), x0, x1);
        } else { x0 = __uint_as_float(vn0); x1 = __uint_as_float(vn1); }
        Dn = pack2(sqrt_approx(fabsf(x0)), sqrt_approx(fabsf(x1)));
    }
    const uint64_t Dprev_in = Dprev;
    Dprev = D[7];
    bool any = false;
    const float4 *wl4 = reinterpret_cast<const float4 *>(wl + c0);
    const float4 *wr4 = reinterpret_cast<const float4 *>(wr + c0);
    const float4 *th4 = reinterpret_cast<const float4 *>(thr + c0);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        const float4 l = wl4[k2], r = wr4[k2], th = th4[k2];
        const int k = 2 * k2;
        const uint64_t Dl0 = k ? D[k - 1] : Dprev_in, Dr0 = D[k + 1];
        const uint64_t Dl1 = D[k], Dr1 = (k + 2 < 8) ? D[k + 2] : Dn;
        const uint64_t T0 = fma2(pack2(l.x, l.y), Dl0, fma2(pack2(r.x, r.y), Dr0, D[k]));
        const uint64_t T1 = fma2(pack2(l.z, l.w), Dl1, fma2(pack2(r.z, r.w), Dr1, D[k + 1]));
        float t0, t1, t2, t3;
        unpack2(T0, t0, t1); unpack2(T1, t2, t3);
        any |= (t0 <= th.x); any |= (t1 <= th.y); any |= (t2 <= th.z); any |= (t3 <= th.w);
    }
    float dp0, dp1, dn0, dn1;
    unpack2(Dprev_in, dp0, dp1); unpack2(Dn, dn0, dn1);
    const bool guard = (minx < xfloor) || (fminf(fminf(dp0, dp1), fminf(dn0, dn1)) < dfloor);
    return any || guard;
}

template <int V>
__global__ void __launch_bounds__(544, 1) k_epi(Args a, unsigned long long *out_cycles, unsigned *sink)
{
    extern __shared__ uint8_t raw[];
    // the round-1 kernel reaches its per-column arrays through an integer round trip, which the compiler can no
    // longer prove to be shared memory (generic LD.E); variant 0 keeps that, the others use the typed pointer
    Smem *sg = reinterpret_cast<Smem *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ Smem ss;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        ss.na[i] = sg->na[i] = 0.5f + 0.001f * (i & 7);
        ss.wl[i] = sg->wl[i] = (i % 20) ? 0.1f : 0.f;
        ss.wr[i] = sg->wr[i] = (i % 20 != 19) ? 0.1f : 0.f;
        ss.thr[i] = sg->thr[i] = 0.01f;
    }
    if (warp == 16) tmem_alloc(&ss.tbase, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = ss.tbase;
    if (warp < 4) {      // fill both accumulator stages with finite values (lane quadrant = warp)
        uint32_t v[16];
        for (int c = 0; c < 512; c += 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(0.05f + 0.001f * ((c + j) & 15) + 1e-4f * lane);
            tmem_st_x16(tbase + c + (static_cast<uint32_t>(warp * 32) << 16), v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    unsigned hits = 0;
    unsigned long long cyc = 0;
    if (warp < a.warps) {
        const int q = warp & 3, grp = warp >> 2, ngrp = a.warps / 4;
        const int nchunks = a.BN / 16;
        const int cbeg = nchunks * grp / ngrp, cend = nchunks * (grp + 1) / ngrp;
        const float nb = 0.25f + 0.001f * lane;
        const float xfloor = 1e-3f, dfloor = 0.0316f;
        const Smem *S = (V == 0) ? sg : &ss;
        const long long t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
            const uint32_t trow = tbase + static_cast<uint32_t>(it & 1) * 256 + (static_cast<uint32_t>(q * 32) << 16);
            uint32_t va[16], vb[16];
            tmem_ld_x16(trow + cbeg * 16, va);
            tmem_ld_wait();
            float dprev = kBig;
            uint64_t Dprev = pack2(kBig, kBig);
            bool hit = false;
            auto body = [&](int ch, uint32_t (&v)[16], uint32_t (&vnx)[16]) {
                const int c0 = ch * 16;
                const bool hasn = (c0 + 16) < a.BN;
                if (ch + 1 < cend) tmem_ld_x16(trow + c0 + 16, vnx);
                else { vnx[0] = __float_as_uint(0.07f); vnx[1] = __float_as_uint(0.07f); }
                if constexpr (V == 6) {
                    unsigned acc = 0;
#pragma unroll
                    for (int j = 0; j < 16; j += 2) acc += v[j] ^ v[j + 1];
                    hit |= (acc == 0x12345u);
                    tmem_ld_wait_x16(vnx);
                    return;
                }
                if constexpr (V == 7) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc += sqrt_approx(fabsf(__uint_as_float(v[j])));
                    hit |= (acc == 12345.f);
                    tmem_ld_wait_x16(vnx);
                    return;
                }
                // (the real kernel waits for the next chunk between the sqrt block and the taps; here the wait sits
                //  in front of the chunk function -- one wait per chunk either way)
                if constexpr (V < 6) tmem_ld_wait_x16(vnx);
                if constexpr (V == 0) hit |= chunk_scalar<0>(v, __uint_as_float(vnx[0]), hasn, dprev, nb, S->na, S->wl, S->wr, S->thr, c0, xfloor, dfloor);
                if constexpr (V == 1) hit |= chunk_scalar<1>(v, __uint_as_float(vnx[0]), hasn, dprev, nb, ss.na, ss.wl, ss.wr, ss.thr, c0, xfloor, dfloor);
                if constexpr (V == 2) hit |= chunk_scalar<2>(v, __uint_as_float(vnx[0]), hasn, dprev, nb, ss.na, ss.wl, ss.wr, ss.thr, c0, xfloor, dfloor);
                if constexpr (V == 3) hit |= chunk_packed<2>(v, vnx[0], vnx[1], hasn, Dprev, ss.na, ss.wl, ss.wr, ss.thr, c0, xfloor, dfloor);
                if constexpr (V == 4) hit |= chunk_scalar<4>(v, __uint_as_float(vnx[0]), hasn, dprev, nb, ss.na, ss.wl, ss.wr, ss.thr, c0, xfloor, dfloor);
                if constexpr (V == 5) hit |= chunk_packed<4>(v, vnx[0], vnx[1], hasn, Dprev, ss.na, ss.wl, ss.wr, ss.thr, c0, xfloor, dfloor);
            };
            for (int ch = cbeg; ch < cend; ch += 2) {
                body(ch, va, vb);
                if (ch + 1 < cend) body(ch + 1, vb, va);
            }
            if (__any_sync(0xffffffffu, hit)) ++hits;
        }
        cyc = static_cast<unsigned long long>(clock64() - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tbase, 512);
    if (lane == 0 && warp < a.warps) atomicMax(out_cycles + blockIdx.x, cyc);
    if (hits) atomicAdd(sink, hits);
}

int main(int argc, char **argv)
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *d_cyc; unsigned *d_sink;
    cudaMalloc(&d_cyc, nsm * sizeof(unsigned long long));
    cudaMalloc(&d_sink, sizeof(unsigned));
    const size_t smem = 200 * 1024;     // one CTA per SM
    typedef void (*kern_t)(Args, unsigned long long *, unsigned *);
    kern_t kerns[8] = {k_epi<0>, k_epi<1>, k_epi<2>, k_epi<3>, k_epi<4>, k_epi<5>, k_epi<6>, k_epi<7>};
    for (int i = 0; i < 8; ++i) cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const int iters = 2000;
    const int bns[2] = {240, 224};
    const char *names[8] = {"0 round-1 shape, generic loads", "1 scalar, LDS", "2 scalar, nb folded", "3 packed f32x2, nb folded",
                            "4 scalar, norms in MMA", "5 packed, norms in MMA", "6 LDTM only", "7 MUFU only"};
    printf("cycles per 128 x BN accumulator tile (tensor pipe needs 16*BN/... : D=512 -> 32 MMAs x BN/2 cycles)\n");
    for (int b = 0; b < 2; ++b) {
        const int BN = bns[b];
        printf("BN = %d (MMA time at D = 512: %d cycles, at D = 2048: %d)\n", BN, 32 * BN / 2, 128 * BN / 2);
        for (int variant = 0; variant < 8; ++variant) {
            for (int warps = 8; warps <= 16; warps += 4) {
                Args a{variant, warps, iters, BN};
                cudaMemset(d_cyc, 0, nsm * sizeof(unsigned long long));
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                kerns[variant]<<<nsm, 544, smem>>>(a, d_cyc, d_sink);      // warm-up
                cudaMemset(d_cyc, 0, nsm * sizeof(unsigned long long));
                cudaEventRecord(e0);
                kerns[variant]<<<nsm, 544, smem>>>(a, d_cyc, d_sink);
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                unsigned long long h[256];
                cudaMemcpy(h, d_cyc, nsm * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                double mean = 0.0;
                for (int i = 0; i < nsm; ++i) mean += static_cast<double>(h[i]);
                mean /= nsm;
                printf("  variant %-34s warps %2d: %7.0f cycles/tile  (%.3f ms total, %.2f GHz eff)\n", names[variant], warps,
                       mean / iters, ms, mean / (ms * 1e6));
            }
        }
    }
    return 0;
}
