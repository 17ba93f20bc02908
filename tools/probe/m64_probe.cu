// GPU-box probe (not product code): (1) which TMEM lanes does a cta_group::1 tcgen05.mma with M=64 write?
// (2) issue rate of M=64 vs M=128 MMAs at the conv kernel's shapes, and of small-N MMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I4dflownet_b200/csrc tools/probe/m64_probe.cu -o tools/probe/m64_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_ptx.cuh"

__device__ __forceinline__ uint64_t dsc(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// A1: 128 rows, value 1000+i ; A2: 128 rows, value i+1 ; B: 16 rows one-hot (k == n)
__global__ void layout_probe(int m2, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* A1 = smem; uint8_t* A2 = smem + 16384; uint8_t* B = smem + 32768;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 2048);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int e = tid; e < 128 * 64; e += blockDim.x) {
        int i = e >> 6, k = e & 63;
        uint32_t byte = i * 128 + k * 2;
        uint32_t sw = byte ^ (((byte >> 7) & 7) << 4);
        *reinterpret_cast<__half*>(A1 + sw) = __float2half((float)(1000 + i));
        *reinterpret_cast<__half*>(A2 + sw) = __float2half((float)(i + 1));
    }
    for (int e = tid; e < 16 * 64; e += blockDim.x) {
        int n = e >> 6, k = e & 63;
        uint32_t byte = n * 128 + k * 2;
        uint32_t sw = byte ^ (((byte >> 7) & 7) << 4);
        *reinterpret_cast<__half*>(B + sw) = __float2half(k == n ? 1.f : 0.f);
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (tid == 0) {
        const uint32_t id128 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idm2 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(m2 >> 4) << 24);
        tc_mma_f16(tm, dsc(smem_u32(A1), 1024), dsc(smem_u32(B), 1024), id128, 0);      // every lane = 1000 + lane
        tc_mma_f16(tm, dsc(smem_u32(A2), 1024), dsc(smem_u32(B), 1024), idm2, 1);        // += row+1 where M=m2 writes
        tc_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (tid < 128) {
        float v[16];
        tc_ld16(tm + ((uint32_t)(tid & ~31) << 16), v);
        tc_ld_wait();
        out[tid] = v[0];
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u) : "memory");
}

// every CTA issues iters x { k1 MMAs (M=m1,N=n1) ; k2 MMAs (M=m2,N=n2) } with precomputed descriptors
__global__ void rate_probe(int m1, int n1, int k1, int m2, int n2, int k2, int iters, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int e = tid; e < 160 * 1024 / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
        const uint32_t id1 = (1u << 4) | ((uint32_t)(n1 >> 3) << 17) | ((uint32_t)(m1 >> 4) << 24);
        const uint32_t id2 = (1u << 4) | ((uint32_t)(n2 >> 3) << 17) | ((uint32_t)(m2 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
        uint64_t ad[4], bd[4], bd2[4];
        for (int k = 0; k < 4; ++k) { ad[k] = dsc(a0 + k * 32, 1024); bd[k] = dsc(b0 + k * 32, 1280); bd2[k] = dsc(b0 + 40960 + k * 32, 1280); }
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint64_t aoff = (uint64_t)(((it & 3) * 16384) >> 4);
            for (int k = 0; k < k1; ++k) tc_mma_f16(tm, ad[k & 3] + aoff, bd[k & 3], id1, 1);
            for (int k = 0; k < k2; ++k) tc_mma_f16(tm, ad[k & 3] + aoff + (8192 >> 4), bd2[k & 3], id2, 1);
        }
        tc_commit(bar);
        mbar_wait(bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

// smaller footprint variant: `issuers` warps of the CTA each issue their own MMA stream into their own accumulator;
// launched with 1 or 2 CTAs per SM.  Tells whether the ~96-cycle floor of small-N MMAs is an issue-rate limit of
// one thread or an execution limit of the tensor pipe.
__global__ void rate_probe2(int n, int k, int iters, int issuers, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    const int tid = threadIdx.x;
    for (int e = tid; e < 96 * 1024 / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    const int w = tid >> 5;
    if ((tid & 31) == 0 && w < issuers) {
        const uint32_t id = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem) + w * 16384, b0 = smem_u32(smem) + 48 * 1024 + w * 16384;
        uint64_t ad[4], bd[4];
        for (int q = 0; q < 4; ++q) { ad[q] = dsc(a0 + q * 32, 1024); bd[q] = dsc(b0 + q * 32, 1024); }
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
            for (int q = 0; q < k; ++q) tc_mma_f16(tm + w * 128, ad[q & 3], bd[q & 3], id, 1);
        tc_commit(bar + w);
        mbar_wait(bar + w, 0);
        cycles[blockIdx.x * 2 + w] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
    float* dout;
    CHECK(cudaMalloc(&dout, 128 * 4));
    std::vector<float> h(128);
    const int smem_l = 32768 + 2048 + 64 + 1024;
    CHECK(cudaFuncSetAttribute(layout_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_l));
    for (int m2 : {128, 64}) {
        layout_probe<<<1, 128, smem_l>>>(m2, dout);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(h.data(), dout, 128 * 4, cudaMemcpyDeviceToHost));
        printf("== M=%d: TMEM lane -> (value - 1000 - lane) = row+1 written by the second MMA (0 = untouched) ==\n", m2);
        for (int l = 0; l < 128; ++l) printf("%3.0f%s", h[l] - 1000 - l, (l % 32 == 31) ? "\n" : " ");
    }
    printf("== rates (cycles per round, all SMs busy) ==\n");
    long long* dcyc;
    CHECK(cudaMalloc(&dcyc, 148 * 8));
    const int smem_r = 160 * 1024 + 64 + 1024;
    CHECK(cudaFuncSetAttribute(rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r));
    struct Cfg { int m1, n1, k1, m2, n2, k2; const char* name; };
    const Cfg cfgs[] = {
        {128, 192, 8, 128, 192, 0, "M128 N192 x8"},
        {128, 192, 4, 128, 192, 4, "M128 N192 x4 + M128 N192 x4 (conv v3 pattern)"},
        {128, 192, 4, 64, 192, 4, "M128 N192 x4 + M64 N192 x4"},
        {64, 192, 8, 64, 192, 0, "M64 N192 x8"},
        {128, 208, 4, 64, 208, 4, "M128 N208 x4 + M64 N208 x4"},
        {128, 256, 8, 128, 256, 0, "M128 N256 x8"},
        {128, 128, 8, 128, 128, 0, "M128 N128 x8"},
        {128, 64, 8, 128, 64, 0, "M128 N64 x8"},
        {128, 32, 8, 128, 32, 0, "M128 N32 x8"},
        {128, 16, 8, 128, 16, 0, "M128 N16 x8"},
        {128, 128, 4, 128, 64, 4, "M128 N128 x4 + M128 N64 x4"},
    };
    std::vector<long long> hc(148);
    for (const Cfg& c : cfgs) {
        const int iters = 400;
        rate_probe<<<148, 128, smem_r>>>(c.m1, c.n1, c.k1, c.m2, c.n2, c.k2, iters, dcyc);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(hc.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : hc) if (v > mx) mx = v;
        const double per_round = (double)mx / iters;
        const double floor_cyc = c.k1 * (128.0 * c.n1 / 256) + c.k2 * (128.0 * c.n2 / 256);
        printf("%-48s: %8.1f cyc/round = %6.1f per MMA, M=128 floor %6.1f, ratio %.2f\n", c.name, per_round,
               per_round / (c.k1 + c.k2), floor_cyc, per_round / floor_cyc);
    }
    printf("== small-N floor: issue rate or execution? (cycles per MMA per issuer) ==\n");
    const int smem2 = 96 * 1024 + 64 + 1024;
    CHECK(cudaFuncSetAttribute(rate_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    long long* dc2;
    CHECK(cudaMalloc(&dc2, 296 * 2 * 8));
    std::vector<long long> h2(296 * 2);
    for (int n : {64, 128, 192})
        for (int ctas : {148, 296})
            for (int issuers : {1, 2}) {
                if (n > 128 && issuers > 1) continue;      // two 192-column accumulators do not fit the 256 allocated columns
                CHECK(cudaMemset(dc2, 0, 296 * 2 * 8));
                rate_probe2<<<ctas, 64, smem2>>>(n, 8, 400, issuers, dc2);
                CHECK(cudaDeviceSynchronize());
                CHECK(cudaMemcpy(h2.data(), dc2, 296 * 2 * 8, cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (auto v : h2) if (v > mx) mx = v;
                printf("N=%3d  %d CTA/SM  %d issuer warp(s): %6.1f cycles per MMA per issuer -> %6.1f cycles per MMA per SM\n", n,
                       ctas / 148, issuers, (double)mx / (400 * 8), (double)mx / (400 * 8) / (issuers * (ctas / 148)));
            }
    return 0;
}
