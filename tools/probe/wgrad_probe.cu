// GPU-box probe (not product code): what paces the stacked wgrad kernel's MMA stream?
// Each CTA issues `iters` rounds of 8 MMAs (M=128, N=192, two accumulators alternating) with MN-major or K-major
// operands, followed by `ncommit` tcgen05.commit to distinct mbarriers, and reports cycles per round.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I4dflownet_b200/csrc tools/probe/wgrad_probe.cu -o tools/probe/wgrad_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_ptx.cuh"

__device__ __forceinline__ uint64_t dsc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// mode 0: MN-major A (two 8 KB atoms, LBO 8192, SBO 1024) and B (three atoms 128 B apart, SBO 1280) -- the wgrad operands
// mode 1: K-major A and B (the conv kernel's operands)
__global__ void probe(int mode, int ncommit, int nacc, int iters, int ring, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 176 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 16);
    const int tid = threadIdx.x;
    for (int e = tid; e < 176 * 1024 / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
    if (tid == 0) { for (int i = 0; i < 16; ++i) mbar_init(bars + i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (tid == 0) {
        const uint32_t tr = mode == 0 ? ((1u << 15) | (1u << 16)) : 0u;
        const uint32_t id = (1u << 4) | tr | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t ya = smem_u32(smem), xa = smem_u32(smem) + 72 * 1024;      // 9 x 8 KB dY slots, 10 x 10 KB X slots
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t ys = ya + (uint32_t)(it % ring) * 8192;
            const uint32_t x1 = xa + (uint32_t)(it % ring) * 10240, x2 = xa + (uint32_t)((it + 2) % ring) * 10240;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint64_t ad, b1, b2;
                if (mode == 0) {
                    ad = dsc(ys + j * 2048, 8192, 1024); b1 = dsc(x1 + j * 2560, 128, 1280); b2 = dsc(x2 + j * 2560, 128, 1280);
                } else {
                    ad = dsc(ys + j * 32, 16, 1024); b1 = dsc(x1 + j * 32, 16, 1280); b2 = dsc(x2 + j * 32, 16, 1280);
                }
                tc_mma_f16(tm, ad, b1, id, 1);
                tc_mma_f16(tm + (nacc == 2 ? 192 : 0), ad, b2, id, 1);
            }
            for (int c = 0; c < ncommit; ++c) tc_commit(bars + ((it * 4 + c) & 7));
        }
        tc_commit(bars + 15);
        mbar_wait(bars + 15, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
    long long* dcyc;
    CHECK(cudaMalloc(&dcyc, 148 * 8));
    const int smem = 176 * 1024 + 256 + 1024;
    CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<long long> hc(148);
    const int iters = 400;
    for (int mode : {0, 1})
        for (int nacc : {2, 1})
            for (int ncommit : {0, 1, 2, 3, 4}) {
                probe<<<148, 128, smem>>>(mode, ncommit, nacc, iters, 8, dcyc);
                CHECK(cudaDeviceSynchronize());
                CHECK(cudaMemcpy(hc.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (auto v : hc) if (v > mx) mx = v;
                printf("%s operands, %d accumulator(s), %d commits/round: %7.1f cycles per round of 8 MMAs = %6.1f per MMA (floor 96)\n",
                       mode == 0 ? "MN-major" : "K-major ", nacc, ncommit, (double)mx / iters, (double)mx / iters / 8);
            }
    return 0;
}
