// GPU-box probe (not product code): (1) does a K-major SWIZZLE_128B tcgen05 shared-memory
// descriptor tolerate a start address that is shifted by whole 128-byte rows and a stride
// between 8-row groups that is not a multiple of 1024 B?  (2) MMA issue rates for the operand
// shapes the conv kernel could use.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -I4dflownet_b200/csrc tools/probe/mma_probe.cu -o tools/probe/mma_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_ptx.cuh"

__device__ __forceinline__ uint64_t desc_ex(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

// ---------------- functional probe ----------------
// A rows live in a dense array of 128-byte rows (row i at byte i*128 from a 1024-aligned base),
// stored with the absolute-address 128B swizzle (what TMA SWIZZLE_128B produces).  A[i][k] = (i%32)*64+k
// (test 0) or i (test 1).  B is one-hot: B[n][k] = (k == 4n+1), so D[m][n] = A[row(m)][4n+1].
__global__ void func_probe(int shift_rows, int sbo_bytes, int base_mode, int test, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __half* A = reinterpret_cast<__half*>(smem);                 // 512 rows x 64
    __half* B = reinterpret_cast<__half*>(smem + 512 * 128);     // 16 rows x 64
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 512 * 128 + 16 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int e = tid; e < 512 * 64; e += blockDim.x) {
        int i = e >> 6, k = e & 63;
        float v = test == 0 ? (float)((i % 32) * 64 + k) : (float)i;
        uint32_t byte = i * 128 + k * 2;
        uint32_t sw = byte ^ (((byte >> 7) & 7) << 4);
        *reinterpret_cast<__half*>(smem + sw) = __float2half(v);
    }
    for (int e = tid; e < 16 * 64; e += blockDim.x) {
        int n = e >> 6, k = e & 63;
        uint32_t byte = n * 128 + k * 2;
        uint32_t sw = byte ^ (((byte >> 7) & 7) << 4);
        *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(B) + sw) = __float2half(k == 4 * n + 1 ? 1.f : 0.f);
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
        const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(A) + shift_rows * 128;
        const uint32_t b0 = smem_u32(B);
        for (int k = 0; k < 4; ++k) {
            uint32_t sa = a0 + k * 32;
            uint32_t bo = base_mode ? ((sa >> 7) & 7) : 0;
            tc_mma_f16(tm, desc_ex(sa, sbo_bytes, bo), desc_ex(b0 + k * 32, 1024, 0), idesc, k != 0);
        }
        tc_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (tid < 128) {
        float v[16];
        tc_ld16(tm + ((uint32_t)(tid & ~31) << 16), v);
        tc_ld_wait();
        for (int n = 0; n < 16; ++n) out[tid * 16 + n] = v[n];
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u) : "memory");
}


// ---------------- MN-major functional probe (the wgrad operand form) ----------------
// Operands are [k rows][64 elements] 128-byte rows (k = voxel index), MN-major: M (or N) = 128 = two
// 64-wide atoms LBO bytes apart ("hi" and "lo" arrays); K = 16 = two 8-row groups SBO bytes apart.
// which = 0: probe A addressing (B one-hot), which = 1: probe B addressing (A one-hot).
__global__ void func_probe_mn(int shift_rows, int sbo_bytes, int which, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    // probed operand: hi rows at 0, lo rows at 32 KB (256 rows each); one-hot operand at 64 KB (hi) / 66 KB (lo)
    const uint32_t LBO_P = 32768, OH = 65536, LBO_OH = 2048;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 72 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int e = tid; e < 2 * 256 * 64; e += blockDim.x) {
        int arr = e / (256 * 64), i = (e >> 6) & 255, m = e & 63;
        float v = (float)((((i + (arr ? 7 : 0)) % 32) * 64) + m);
        uint32_t byte = arr * LBO_P + i * 128 + m * 2;
        uint32_t sw = byte ^ (((byte >> 7) & 7) << 4);
        *reinterpret_cast<__half*>(smem + sw) = __float2half(v);
    }
    // one-hot operand: element (j, k) = (k == j % 16), j = 0..127 (atom = j / 64), rows k = 0..15
    for (int e = tid; e < 2 * 16 * 64; e += blockDim.x) {
        int arr = e / (16 * 64), k = (e >> 6) & 15, jj = e & 63;
        int j = arr * 64 + jj;
        uint32_t byte = OH + arr * LBO_OH + k * 128 + jj * 2;
        uint32_t sw = byte ^ (((byte >> 7) & 7) << 4);
        *reinterpret_cast<__half*>(smem + sw) = __float2half((j % 16) == k ? 1.f : 0.f);
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (tid == 0) {
        // D=f32, A=B=f16, both MN-major, M=128, N=128
        const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        auto desc_mn = [](uint32_t saddr, uint32_t lbo, uint32_t sbo) {
            uint64_t d = 0;
            d |= (uint64_t)((saddr >> 4) & 0x3FFF);
            d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
            d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
            d |= (uint64_t)1 << 46;
            d |= (uint64_t)2 << 61;
            return d;
        };
        const uint64_t dp = desc_mn(smem_u32(smem) + shift_rows * 128, LBO_P, sbo_bytes);
        const uint64_t doh = desc_mn(smem_u32(smem) + OH, LBO_OH, 1024);
        if (which == 0) tc_mma_f16(tm, dp, doh, idesc, 0);
        else tc_mma_f16(tm, doh, dp, idesc, 0);
        tc_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (tid < 128) {
        for (int c = 0; c < 128; c += 16) {
            float v[16];
            tc_ld16(tm + ((uint32_t)(tid & ~31) << 16) + c, v);
            tc_ld_wait();
            for (int n = 0; n < 16; ++n) out[tid * 128 + c + n] = v[n];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u) : "memory");
}

// ---------------- rate probe ----------------
// every CTA issues `iters` rounds of {MMA(M=128,N=n1) x k1 ; MMA(M=128,N=n2) x k2} on resident smem
__global__ void rate_probe(int n1, int k1, int n2, int k2, int iters, int a_rows_stride, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
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
        const uint32_t id1 = (1u << 4) | ((uint32_t)(n1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t id2 = (1u << 4) | ((uint32_t)(n2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            // walk A through a 64 KB window so consecutive MMAs read different shared memory
            const uint32_t aoff = (it & 3) * 16384;
            for (int k = 0; k < k1; ++k)
                tc_mma_f16(tm, desc_ex(a0 + aoff + (k & 3) * 32, a_rows_stride, 0), desc_ex(b0 + (k & 3) * 32, 1024, 0), id1, 1);
            for (int k = 0; k < k2; ++k)
                tc_mma_f16(tm + 256, desc_ex(a0 + aoff + (k & 3) * 32, a_rows_stride, 0), desc_ex(b0 + 32768 + (k & 3) * 32, 1024, 0), id2, 1);
        }
        tc_commit(bar);
        mbar_wait(bar, 0);
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
    float* dout;
    CHECK(cudaMalloc(&dout, 128 * 16 * 4));
    std::vector<float> h(128 * 16);
    const int smem_f = 512 * 128 + 16 * 128 + 64 + 1024;
    CHECK(cudaFuncSetAttribute(func_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
    printf("== functional: shifted start / non-1024 SBO ==\n");
    const int shifts[] = {0, 1, 3, 8, 10, 11};
    const int sbos[] = {1024, 1280, 1152};
    for (int sbo : sbos)
        for (int sh : shifts)
            for (int bm = 0; bm < 2; ++bm) {
                int bad_rows = 0, bad_k = 0;
                int first_obs_row = -1, first_obs_k = -1, first_m = -1;
                for (int test = 0; test < 2; ++test) {
                    func_probe<<<1, 128, smem_f>>>(sh, sbo, bm, test, dout);
                    CHECK(cudaDeviceSynchronize());
                    CHECK(cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost));
                    for (int m = 0; m < 128; ++m) {
                        int row = sh + (m / 8) * (sbo / 128) + (m % 8);
                        for (int n = 0; n < 16; ++n) {
                            float v = h[m * 16 + n];
                            if (test == 1) {
                                if (v != (float)row) { ++bad_rows; if (first_m < 0) { first_m = m; first_obs_row = (int)v; } }
                            } else {
                                float exp = (float)((row % 32) * 64 + 4 * n + 1);
                                if (v != exp) { ++bad_k; if (first_obs_k < 0) first_obs_k = ((int)v) % 64 * 1000 + 4 * n + 1; }
                            }
                        }
                    }
                }
                printf("sbo=%4d shift=%2d base_off=%s : row mismatches %4d, k mismatches %4d", sbo, sh, bm ? "auto" : "0   ", bad_rows, bad_k);
                if (first_m >= 0) printf("  (first: m=%d expected row %d got %d)", first_m, sh + (first_m / 8) * (sbo / 128) + first_m % 8, first_obs_row);
                if (first_obs_k >= 0) printf("  (first k: got*1000+exp = %d)", first_obs_k);
                printf("\n");
            }


    {
        printf("== functional MN-major: shifted start / non-1024 SBO ==\n");
        float* dmn;
        CHECK(cudaMalloc(&dmn, 128 * 128 * 4));
        std::vector<float> hm(128 * 128);
        const int smem_mn = 72 * 1024 + 64 + 1024;
        CHECK(cudaFuncSetAttribute(func_probe_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_mn));
        for (int sbo : {1024, 1280})
            for (int sh : {0, 1, 3, 10, 21})
                for (int which = 0; which < 2; ++which) {
                    func_probe_mn<<<1, 128, smem_mn>>>(sh, sbo, which, dmn);
                    CHECK(cudaDeviceSynchronize());
                    CHECK(cudaMemcpy(hm.data(), dmn, hm.size() * 4, cudaMemcpyDeviceToHost));
                    int bad = 0, fm = -1, fn = -1; float fv = 0, fe = 0;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < 128; ++n) {
                            // which=0: D[m][n] = P(m, k = n%16); which=1: D[m][n] = P(n, k = m%16)
                            int j = which == 0 ? m : n, k = which == 0 ? n % 16 : m % 16;
                            int row = sh + (k / 8) * (sbo / 128) + (k % 8);
                            int arr = j / 64;
                            float exp = (float)((((row + (arr ? 7 : 0)) % 32) * 64) + (j % 64));
                            float v = hm[m * 128 + n];
                            if (v != exp) { if (!bad) { fm = m; fn = n; fv = v; fe = exp; } ++bad; }
                        }
                    printf("MN sbo=%4d shift=%2d probe %c: mismatches %5d", sbo, sh, which ? 'B' : 'A', bad);
                    if (bad) printf("  (first m=%d n=%d got %.0f expected %.0f)", fm, fn, fv, fe);
                    printf("\n");
                }
    }
    printf("== rates (cycles per MMA, all SMs busy) ==\n");
    long long* dcyc;
    CHECK(cudaMalloc(&dcyc, 148 * 8));
    const int smem_r = 160 * 1024 + 64 + 1024;
    CHECK(cudaFuncSetAttribute(rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r));
    struct Cfg { int n1, k1, n2, k2, stride; const char* name; };
    const Cfg cfgs[] = {
        {256, 8, 256, 0, 1024, "N=256"}, {192, 8, 192, 0, 1024, "N=192"}, {208, 8, 208, 0, 1024, "N=208"},
        {128, 8, 128, 0, 1024, "N=128"}, {64, 8, 64, 0, 1024, "N=64"}, {96, 8, 96, 0, 1024, "N=96"},
        {128, 4, 64, 4, 1024, "N=128 x4 + N=64 x4"}, {128, 4, 64, 4, 1280, "N=128 x4 + N=64 x4, SBO 1280"},
        {128, 8, 128, 0, 1280, "N=128, SBO 1280"}, {256, 4, 128, 4, 1024, "N=256 x4 + N=128 x4"},
    };
    std::vector<long long> hc(148);
    for (const Cfg& c : cfgs) {
        const int iters = 400;
        rate_probe<<<148, 128, smem_r>>>(c.n1, c.k1, c.n2, c.k2, iters, c.stride, dcyc);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(hc.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0, mn = 1LL << 60;
        for (auto v : hc) { if (v > mx) mx = v; if (v < mn) mn = v; }
        double per_round = (double)mx / iters;
        double floor_cyc = c.k1 * (128.0 * c.n1 / 256) + c.k2 * (128.0 * c.n2 / 256);
        printf("%-34s: %8.1f cyc/round (min SM %8.1f), floor %6.1f, ratio %.2f\n", c.name, per_round, (double)mn / iters, floor_cyc, per_round / floor_cyc);
    }
    return 0;
}
