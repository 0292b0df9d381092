// mma_rate.cu -- microbenchmark: peak tcgen05.mma.cta_group::2 kind::f16 issue rate on this GPU as a function
// of N, MMAs per commit and commit/wait depth.  No loads, no epilogue (operands are whatever is in smem).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../embodied-one-shot-video-recognition_b200/csrc/eosvr_ptx.cuh"
using namespace eosvr::ptx;

struct Args { int N, nmma, depth, iters, split, flags, issuers; };   // flags: 1 no fence, 2 no waits in the loop, 4 plain (non-multicast) commit

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_rate(Args a, unsigned long long *cycles)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem, *sB = smem + 6 * 16384;
    __shared__ uint64_t bars[32];
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 12 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 32; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    fence_proxy_async();
    if (warp == 1) tmem_alloc_2sm(&tbase, 512);
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    const uint32_t tb = tbase;
    if ((a.flags & 8) && warp == 0 && rank == 0) {          // D = 16 everywhere (one overwriting MMA of ones)
        if (lane == 0) {
            mma_f16_ss_2sm(tb, umma_desc_sw128(smem_u32(sA), 0), umma_desc_sw128(smem_u32(sB), 0), umma_idesc_f16(0, 256, a.N), 0u);
            mma_commit_2sm(&bars[31], 1);
        }
        __syncwarp();
        mbar_wait(&bars[31], 0);
        tc_fence_after();
    }
    __syncthreads();
    if ((warp == 0 || (warp == 2 && a.issuers == 2)) && rank == 0) {      // whole warp runs the loop (uniform operands); one lane issues
        uint64_t *bars_w = bars + (warp == 2 ? 16 : 0);
        const uint32_t idesc = umma_idesc_f16(0, 256, a.N);
        unsigned long long t0 = clock64();
        uint32_t phase_bits = 0;
        for (int it = 0; it < a.iters; ++it) {
            const int st = it % 6;
            const uint32_t a0 = smem_u32(sA + st * 16384), b0 = smem_u32(sB + st * 16384);
            const uint32_t d = (a.flags & 8) ? tb : tb + (a.issuers == 2 ? (warp == 2 ? 256 : 0) : ((it / a.split) & 1) * 256);
            const int slot = it % a.depth;
            if (lane == 0) {
                for (int j = 0; j < a.nmma; ++j)
                    mma_f16_ss_2sm(d, umma_desc_sw128(a0, (j & 3) * 32), umma_desc_sw128(b0, (j & 3) * 32), idesc, 1u);
                if (a.flags & 4)
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars_w[slot])) : "memory");
                else
                    mma_commit_2sm(&bars_w[slot], 1);
            }
            __syncwarp();
            if (!(a.flags & 2) && it >= a.depth - 1) {
                const int ws = (it - (a.depth - 1)) % a.depth;
                mbar_wait(&bars_w[ws], (phase_bits >> ws) & 1u);
                phase_bits ^= 1u << ws;
                if (!(a.flags & 1)) tc_fence_after();
            }
        }
        if (a.flags & 2) {       // commits complete in order: drain through one last commit on a fresh barrier
            if (lane == 0) mma_commit_2sm(&bars_w[15], 1);
            __syncwarp();
            mbar_wait(&bars_w[15], 0);
        } else
        for (int k = a.iters - (a.depth - 1); k < a.iters; ++k) {
            if (k < 0) continue;
            const int ws = k % a.depth;
            mbar_wait(&bars_w[ws], (phase_bits >> ws) & 1u);
            phase_bits ^= 1u << ws;
        }
        if (lane == 0 && warp == 0) cycles[blockIdx.x / 2] = clock64() - t0;
    }
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    if (a.flags & 8) {                                      // every element must equal 16 * (1 + MMAs issued)
        const float expect = 16.0f * (1.0f + (float)a.nmma * (float)a.iters * (float)a.issuers);
        unsigned bad = 0;
        for (int c0 = 0; c0 < a.N; c0 += 16) {
            uint32_t v[16];
            tmem_ld_x16(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) bad += (__uint_as_float(v[j]) != expect);
        }
        if (bad) atomicAdd(reinterpret_cast<unsigned *>(cycles + 512), bad);
        if (warp == 0 && blockIdx.x == 0) {
            uint32_t v[16];
            tmem_ld_x16(tb, v); tmem_ld_wait();
            if (lane == 0) {
                reinterpret_cast<float *>(cycles + 513)[0] = __uint_as_float(v[0]);
                reinterpret_cast<float *>(cycles + 513)[1] = expect;
            }
        }
        tc_fence_before();
        cluster_sync();
    }
    if (warp == 1) tmem_dealloc_2sm(tb, 512);
}

int main(int argc, char **argv)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *cyc;
    cudaMalloc(&cyc, 1024 * 8); cudaMemset(cyc, 0, 1024 * 8);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    int Ns[] = {224};
    int nmmas[] = {2, 8};
    int flagset[] = {8};
    int issuers[] = {1, 2};
    printf("sms=%d\n", sms); fflush(stdout);
    for (int N : Ns) for (int nm : nmmas) for (int fl : flagset) for (int is : issuers) {
        const int dp = 4;
        Args a{N, nm, dp, 2048 * 4 / nm, 32 * 4 / nm > 0 ? 32 * 4 / nm : 1, fl, is};
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_rate<<<sms / 2 * 2, 128, 200 * 1024>>>(a, cyc);   // warm-up
        cudaEventRecord(e0);
        k_rate<<<sms / 2 * 2, 128, 200 * 1024>>>(a, cyc);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long c0 = 0;
        cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
        if (fl & 8) {
            unsigned bad = 0; float two[2];
            cudaMemcpy(&bad, cyc + 512, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(two, cyc + 513, 8, cudaMemcpyDeviceToHost);
            printf("   same-accumulator check: wrong elements=%u (over 2 launches) sample=%.1f expect=%.1f\n", bad, two[0], two[1]);
            cudaMemset(cyc + 512, 0, 8);
        }
        const double flops = 2.0 * 256 * N * 16 * nm * a.iters * (sms / 2) * is;
        printf("N=%3d mma/commit=%2d issuers=%d : %.3f ms  %.1f TFLOP/s  cycles/mma=%.1f (ideal %d)  clk~%.0f MHz\n", N, nm, is, ms,
               flops / ms / 1e9, (double)c0 / (nm * (double)a.iters * is), N / 2, c0 / (ms * 1e3));
    }
    return 0;
}
