// GPU-box probe (not product code): sustained tcgen05.mma throughput under the power cap for the instruction mixes of
// the 64->64 forward, to see what the fourth (masked) row-product of the split-fp16 scheme costs in ENERGY:
//   * M=128 + M=128 with lanes 64..127 masked (the conv kernel today; A rows 64..127 = the next weights in shared memory)
//   * the same with zeros in A rows 64..127 of the second instruction (data-dependent power?)
//   * M=128 + M=64 (half the rows physically absent)
// Every config runs ~3 s with all 148 SMs busy, operands = pseudo-random fp16 in shared memory (no memory traffic);
// the host samples nvidia-smi (power, SM clock) 1.5 s into the run.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I4dflownet_b200/csrc tools/probe/power_probe.cu -o tools/probe/power_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
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

// rounds of { k1 x MMA(m1, n) ; k2 x MMA(m2, n) [second: optional lane mask 64..127, optional zero A rows 64..127] }
__global__ void power_probe(int m1, int m2, int n, int k1, int k2, int mask2, int zero2, int iters, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    // pseudo-random halves in [-1, 1): A images at 0..64 KB (four 16 KB images), B tiles at 64 KB.. (two 40 KB planes)
    for (int e = tid; e < 160 * 1024 / 2; e += blockDim.x) {
        uint32_t x = (uint32_t)e * 2654435761u + blockIdx.x * 40503u;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        reinterpret_cast<__half*>(smem)[e] = __float2half(((float)(x & 0xffff) - 32768.f) / 32768.f);
    }
    __syncthreads();
    if (zero2)   // the second instruction reads A at image + 8 KB: zero its rows 64..127 (= first half of the next image)
        for (int img = 0; img < 4; ++img)
            for (int e = tid; e < 8192 / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem + ((img + 1) & 3) * 16384)[e] = 0u;
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
        const uint32_t id1 = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m1 >> 4) << 24);
        const uint32_t id2 = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m2 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
        uint64_t ad[4], bd[4], bd2[4];
        for (int k = 0; k < 4; ++k) { ad[k] = dsc(a0 + k * 32, 1024); bd[k] = dsc(b0 + k * 32, 1280); bd2[k] = dsc(b0 + 40960 + k * 32, 1280); }
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            // with zero2 only images 0 and 2 are used as first operands (1 and 3 have a zeroed first half)
            const int img = zero2 ? (it & 1) * 2 : (it & 3);
            const uint64_t aoff = (uint64_t)((img * 16384) >> 4);
            const uint32_t acc = it != 0;
            for (int k = 0; k < k1; ++k) tc_mma_f16(tm, ad[k & 3] + aoff, bd[k & 3], id1, acc);
            if (mask2)
                for (int k = 0; k < k2; ++k)
                    tc_mma_f16_masked(tm + 256, ad[k & 3] + aoff + (8192 >> 4), bd2[k & 3], id2, acc, 0u, 0u, 0xffffffffu, 0xffffffffu);
            else
                for (int k = 0; k < k2; ++k) tc_mma_f16(tm + 256, ad[k & 3] + aoff + (8192 >> 4), bd2[k & 3], id2, acc);
        }
        tc_commit(bar);
        mbar_wait(bar, 0);
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
    const int smem_r = 160 * 1024 + 64 + 1024;
    CHECK(cudaFuncSetAttribute(power_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r));
    struct Cfg { int m1, m2, n, k1, k2, mask2, zero2; const char* name; };
    const Cfg cfgs[] = {
        {128, 128, 192, 4, 4, 1, 0, "M128 + M128 masked 64..127 (conv forward today)"},
        {128, 128, 192, 4, 4, 1, 1, "M128 + M128 masked, A rows 64..127 zero"},
        {128, 64, 192, 4, 4, 0, 0, "M128 + M64"},
        {128, 128, 192, 4, 0, 0, 0, "M128 only (2 useful products per MAC: dgrad)"},
        {128, 128, 192, 4, 4, 0, 0, "M128 + M128 unmasked"},
        {64, 64, 192, 4, 4, 0, 0, "M64 + M64"},
        {128, 64, 128, 4, 4, 0, 0, "M128 + M64, N=128"},
        {128, 128, 128, 4, 4, 1, 0, "M128 + M128 masked, N=128"},
    };
    std::vector<long long> hc(148);
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    printf("%-52s %10s %10s %9s %s\n", "config", "ns/round", "cyc/round", "eff. MHz", "nvidia-smi (power W, SM MHz) mid-run");
    for (const Cfg& c : cfgs) {
        const int iters = 200000;              // ~0.1 s per launch
        power_probe<<<148, 128, smem_r>>>(c.m1, c.m2, c.n, c.k1, c.k2, c.mask2, c.zero2, 20000, dcyc);   // warm-up
        CHECK(cudaDeviceSynchronize());
        char smi[256] = "";
        std::thread sampler([&] {
            std::this_thread::sleep_for(std::chrono::milliseconds(1500));
            FILE* f = popen("nvidia-smi --query-gpu=power.draw,clocks.sm --format=csv,noheader,nounits", "r");
            if (f) { if (!fgets(smi, sizeof smi, f)) smi[0] = 0; pclose(f); }
        });
        const int launches = 30;
        CHECK(cudaEventRecord(e0));
        for (int l = 0; l < launches; ++l) power_probe<<<148, 128, smem_r>>>(c.m1, c.m2, c.n, c.k1, c.k2, c.mask2, c.zero2, iters, dcyc);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaDeviceSynchronize());
        sampler.join();
        float ms = 0;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        CHECK(cudaMemcpy(hc.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : hc) if (v > mx) mx = v;
        const double ns_round = (double)ms * 1e6 / ((double)launches * iters);
        const double cyc_round = (double)mx / iters;
        for (char* p = smi; *p; ++p) if (*p == '\n') *p = 0;
        printf("%-52s %10.1f %10.1f %9.0f %s\n", c.name, ns_round, cyc_round, cyc_round / ns_round * 1e3, smi);
        fflush(stdout);
    }
    return 0;
}
