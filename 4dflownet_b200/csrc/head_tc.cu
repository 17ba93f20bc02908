// The three 64->1 head convolutions (SR4DFlowNet.py:40,43,46,49) on the tensor cores, forward.
//
//     out[v][c] = b_c + sum_t sum_ci w_c[t][ci] * hpad_c[v + t][ci]
//
// is split into a GEMM over voxel ROWS and a 27-point re-indexed sum:
//   (1) head_tapdot_kernel:  P_c[row][t] = sum_ci h_c[row][ci] * w_c[t][ci]   for every row of the padded Act tensor
//       (B (H+2)^3 rows of 64 channels -- the replicate halo is materialised, so halo rows simply repeat interior rows).
//       tcgen05 with the VOXELS on M: A = 128 consecutive rows of the hi (then lo) plane, K-major SWIZZLE_128B straight
//       from a TMA box; B = the head's weight image [Whi (27 taps, padded to 32 rows) ; Wlo (32 rows)] x 64 ci.
//       Split-fp16 like the 64->64 layers: TMEM columns 0..31 = Xhi Whi, columns 32..63 = Xhi Wlo + Xlo Whi;
//       P = H + L / 2048.  The kernel is HBM-bound by construction (reads 256 B, writes 128 B per row; ~350 cycles of
//       MMAs per 128 rows), so a plain persistent pipeline suffices: warp 0 TMA producer (4 stages of 32 KB), warp 1
//       MMA issuer, warps 2..5 epilogue (one accumulator row per thread, 8 accumulator slots of 64 TMEM columns).
//   (2) head_sum_kernel:  out[v][c] = b_c + sum_t P_c[pad(v) + t][t], a CTA marching along x with three planes of P rows
//       in shared memory.
// Replaces head_out_kernel (fp32 FMA register-tiled GEMM, 0.59 ms per step at B = 8) on the tensor-core path.
#include <cuda.h>

#include <cstdio>

#include "kernels.h"
#include "tc_host.h"
#include "tc_ptx.cuh"

namespace {

constexpr int ROWS_T = 128;                      // voxel rows per tile
constexpr int PLANE_BYTES = ROWS_T * 128;        // 16 KB: 128 rows x 64 fp16
constexpr int NST = 4;                           // TMA stages (hi + lo tile each)
constexpr int WIMG_BYTES = 64 * 128;             // 8 KB per head: [Whi 32 rows ; Wlo 32 rows] x 64 ci, swizzled
constexpr int NACC = 8;                          // accumulator slots of 64 TMEM columns
constexpr int HT_SMEM = 1024 + NST * 2 * PLANE_BYTES + 3 * WIMG_BYTES + 256;
constexpr int HT_THREADS = 64 + 128;

struct HeadTcParams {
    const __half* wimg;      // [3][64 rows][64] fp16 swizzled images
    float* P[3];             // [nrows][32] fp32 per head
    long long nrows;         // B (H+2)^3
    int ntiles;              // ceil(nrows / 128)
};

__device__ __forceinline__ uint64_t kdesc(uint32_t saddr) {      // K-major SWIZZLE_128B, 8-row groups 1024 B apart
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(HT_THREADS, 1)
head_tapdot_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                   const __grid_constant__ CUtensorMap m2, HeadTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* xs = smem;                                   // NST x [hi tile | lo tile]
    uint8_t* wsm = smem + NST * 2 * PLANE_BYTES;          // 3 weight images
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + 3 * WIMG_BYTES);
    uint64_t* full = bars;                  // [NST]
    uint64_t* empty = full + NST;           // [NST]
    uint64_t* acc_full = empty + NST;       // [NACC]
    uint64_t* acc_empty = acc_full + NACC;  // [NACC]
    uint64_t* wbar = acc_empty + NACC;      // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&m0); prefetch_tmap(&m1); prefetch_tmap(&m2);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total = 3 * p.ntiles;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wbar, 3 * WIMG_BYTES);
            bulk_load(wsm, p.wimg, 3 * WIMG_BYTES, wbar);
            uint32_t it = 0;
            for (int id = blockIdx.x; id < total; id += gridDim.x, ++it) {
                const int c = id / p.ntiles, t = id - c * p.ntiles;
                const CUtensorMap* m = c == 0 ? &m0 : (c == 1 ? &m1 : &m2);
                const uint32_t s = it % NST, ph = (it / NST) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], 2 * PLANE_BYTES);
                tma_load_3d(xs + s * 2 * PLANE_BYTES, m, &full[s], 0, t * ROWS_T, 0);                    // hi rows
                tma_load_3d(xs + s * 2 * PLANE_BYTES + PLANE_BYTES, m, &full[s], 0, t * ROWS_T, 1);      // lo rows
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D=f32, A=B=f16, both K-major, M=128; N = 64 (Xhi x [Whi;Wlo]) and N = 32 (Xlo x Whi)
            const uint32_t id64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t id32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            mbar_wait(wbar, 0);
            uint32_t it = 0;
            for (int id = blockIdx.x; id < total; id += gridDim.x, ++it) {
                const int c = id / p.ntiles;
                const uint32_t s = it % NST, ph = (it / NST) & 1;
                const uint32_t a = it % NACC, aph = (it / NACC) & 1;
                mbar_wait(&acc_empty[a], aph ^ 1);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t xhi = smem_u32(xs + s * 2 * PLANE_BYTES), xlo = xhi + PLANE_BYTES;
                const uint32_t wb = smem_u32(wsm + c * WIMG_BYTES);
                const uint32_t d = tmem_base + a * 64;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc_mma_f16(d, kdesc(xhi + k * 32), kdesc(wb + k * 32), id64, k != 0);        // cols 0..31 H, 32..63 L
                    tc_mma_f16(d + 32, kdesc(xlo + k * 32), kdesc(wb + k * 32), id32, 1u);       // L += Xlo Whi
                }
                tc_commit(&empty[s]);
                tc_commit(&acc_full[a]);
            }
        }
    } else {
        // epilogue: warps 2..5, TMEM lane quarter = warp % 4, one row per thread
        const int q = warp & 3;
        const int row_in_tile = 32 * q + lane;
        uint32_t it = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x, ++it) {
            const int c = id / p.ntiles, t = id - c * p.ntiles;
            const uint32_t a = it % NACC, aph = (it / NACC) & 1;
            mbar_wait(&acc_full[a], aph);
            tc_fence_after();
            float hv[32], lv[32];
            const uint32_t taddr = tmem_base + a * 64 + ((uint32_t)(32 * q) << 16);
            tc_ld32(taddr, hv);
            tc_ld32(taddr + 32, lv);
            tc_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
            const long long row = (long long)t * ROWS_T + row_in_tile;
            if (row < p.nrows) {
                float* Pc = c == 0 ? p.P[0] : (c == 1 ? p.P[1] : p.P[2]);
                float4* o = reinterpret_cast<float4*>(Pc + row * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    o[j] = make_float4(fmaf(lv[4 * j], SR4D_LO_INV, hv[4 * j]), fmaf(lv[4 * j + 1], SR4D_LO_INV, hv[4 * j + 1]),
                                       fmaf(lv[4 * j + 2], SR4D_LO_INV, hv[4 * j + 2]), fmaf(lv[4 * j + 3], SR4D_LO_INV, hv[4 * j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// weight images of the three heads: rows 0..31 = Whi[t] (t >= 27: zero), rows 32..63 = Wlo[t]; K = ci; SWIZZLE_128B
__global__ void head_wimg_kernel(const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2,
                                 __half* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * 64 * 64) return;
    const int k = i & 63, row = (i >> 6) & 63, c = i >> 12;
    const float* w = c == 0 ? w0 : (c == 1 ? w1 : w2);
    const int t = row & 31;
    const float v = t < 27 ? w[t * 64 + k] : 0.f;          // Keras kernel (3,3,3,64,1) = [27][64]
    __half h, lo;
    split_f16(v, h, lo);
    const int grp = row >> 3, rr = row & 7;
    img[(size_t)c * 4096 + grp * 512 + rr * 64 + (((k >> 3) ^ rr) << 3) + (k & 7)] = row < 32 ? h : lo;
}

// ---- second half: out[v][c] = bias_c + sum_t P_c[padded(v) + t][t] -------------------------------------------------------
// A CTA owns a (YT x ZT) tile of one head and sample and marches along x with three planes of P rows (halo included) in
// shared memory, row pitch 33 floats (consecutive rows -> consecutive banks).
constexpr int HS_ZT = 48, HS_PITCH = 33;
template <int YT>
__global__ void __launch_bounds__(256) head_sum_kernel(const float* __restrict__ P0, const float* __restrict__ P1,
                                                       const float* __restrict__ P2, const float* __restrict__ b0,
                                                       const float* __restrict__ b1, const float* __restrict__ b2,
                                                       float* __restrict__ out, int B, int D, int xseg) {
    extern __shared__ float hs_sm[];
    constexpr int ROWS = (YT + 2) * (HS_ZT + 2);
    const int Dp = D + 2;
    const int nyt = (D + YT - 1) / YT, nzt = (D + HS_ZT - 1) / HS_ZT;
    int bi = blockIdx.x;
    const int zt = bi % nzt; bi /= nzt;
    const int yt = bi % nyt; bi /= nyt;
    const int c = bi % 3;
    const int b = bi / 3;
    const float* P = c == 0 ? P0 : (c == 1 ? P1 : P2);
    const float bias = *(c == 0 ? b0 : (c == 1 ? b1 : b2));
    const int y0 = yt * YT, z0 = zt * HS_ZT;
    const int ny = min(YT, D - y0), nz = min(HS_ZT, D - z0);
    // A plane of P rows travels global -> registers -> shared memory: the loads of plane x+3 are issued before the sums
    // of output plane x are formed and stored to shared memory after them, so their latency hides behind the arithmetic
    // (one load at a time -- the first version -- cost 4 us per plane).
    constexpr int NLD = (ROWS * 7 + 255) / 256;
    float4 v[NLD];
    // this thread's items are the same (row, 16-byte piece) pairs in every plane: offsets computed once, so the fetch is a
    // run of independent loads (with the index arithmetic inside it the compiler serialised them: 14 exposed latencies)
    int goff[NLD], soff[NLD];
    unsigned w4 = 0;                                   // bit k: item k carries a fourth tap (piece q < 6)
    const int nitem = (ny + 2) * (nz + 2) * 7;
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
        const int i = threadIdx.x + k * 256;
        const int row = i / 7, q = i - row * 7;
        const int ly = row / (nz + 2), lz = row - ly * (nz + 2);
        goff[k] = i < nitem ? (((y0 + ly) * Dp + z0 + lz) * 32 + q * 4) : -1;
        soff[k] = (ly * (HS_ZT + 2) + lz) * HS_PITCH + q * 4;
        if (q < 6) w4 |= 1u << k;
    }
    const float* Pb = P + (size_t)b * Dp * Dp * Dp * 32;
    auto fetch = [&](int xp) {                         // xp: padded plane index 0..D+1
        const float* Pp = Pb + (size_t)xp * Dp * Dp * 32;
#pragma unroll
        for (int k = 0; k < NLD; ++k)
            if (goff[k] >= 0) v[k] = *reinterpret_cast<const float4*>(Pp + goff[k]);
    };
    auto stash = [&](int xp) {
        float* dst = hs_sm + (xp % 3) * ROWS * HS_PITCH;
#pragma unroll
        for (int k = 0; k < NLD; ++k)
            if (goff[k] >= 0) {
                float* d = dst + soff[k];
                d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z;
                if ((w4 >> k) & 1u) d[3] = v[k].w;     // taps 0..26
            }
    };
    // blockIdx.y: segment of xseg output planes (small batches: more, shorter marches instead of 48 latency-bound steps)
    const int xs0 = blockIdx.y * xseg, xs1 = min(D, xs0 + xseg);
    fetch(xs0); stash(xs0);
    fetch(xs0 + 1); stash(xs0 + 1);
    fetch(xs0 + 2);
    for (int x = xs0; x < xs1; ++x) {
        stash(x + 2);                                  // slot (x+2)%3 held plane x-1: its readers passed the barrier below
        __syncthreads();
        if (x + 3 <= xs1 + 1) fetch(x + 3);
        const float* pl[3] = {hs_sm + (x % 3) * ROWS * HS_PITCH, hs_sm + ((x + 1) % 3) * ROWS * HS_PITCH,
                              hs_sm + ((x + 2) % 3) * ROWS * HS_PITCH};
        for (int i = threadIdx.x; i < ny * nz; i += 256) {
            const int ly = i / nz, lz = i - ly * nz;
            float s = bias;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz)
                        s += pl[dx][((ly + dy) * (HS_ZT + 2) + lz + dz) * HS_PITCH + (dx * 3 + dy) * 3 + dz];
            out[((((size_t)b * D + x) * D + y0 + ly) * D + z0 + lz) * 3 + c] = s;
        }
        __syncthreads();
    }
}

}  // namespace

size_t head_tc_wimg_halves() { return 3 * 64 * 64; }

cudaError_t launch_head_out_tc(ActView h0, ActView h1, ActView h2, const float* w0, const float* w1, const float* w2,
                               const float* b0, const float* b1, const float* b2, __half* wimg, float* P0, float* P1,
                               float* P2, float* out, cudaStream_t s) {
    const int B = h0.B, D = h0.D;
    const long long nrows = (long long)B * (D + 2) * (D + 2) * (D + 2);
    const long long plane = (long long)act_plane_elems(B, D);
    if (h0.lo != h0.hi + plane || h1.lo != h1.hi + plane || h2.lo != h2.hi + plane) return cudaErrorInvalidValue;
    head_wimg_kernel<<<(3 * 64 * 64 + 255) / 256, 256, 0, s>>>(w0, w1, w2, wimg);
    CUtensorMap m0, m1, m2;
    if (!tc_make_rows_map(&m0, h0.hi, nrows, plane, ROWS_T) || !tc_make_rows_map(&m1, h1.hi, nrows, plane, ROWS_T) ||
        !tc_make_rows_map(&m2, h2.hi, nrows, plane, ROWS_T))
        return cudaErrorUnknown;
    cudaError_t e = tc_func_smem(reinterpret_cast<const void*>(head_tapdot_kernel), HT_SMEM);
    if (e != cudaSuccess) return e;
    HeadTcParams p;
    p.wimg = wimg; p.P[0] = P0; p.P[1] = P1; p.P[2] = P2; p.nrows = nrows;
    p.ntiles = (int)((nrows + ROWS_T - 1) / ROWS_T);
    const int sms = tc_num_sms();
    const int grid = 3 * p.ntiles < sms ? 3 * p.ntiles : sms;
    head_tapdot_kernel<<<grid, HT_THREADS, HT_SMEM, s>>>(m0, m1, m2, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // tile height of the sum kernel: the largest of 8 / 4 / 2 / 1 lines that still gives (90 % of) the SMs a CTA
    // and, before the tiles get flat, the march along x is cut into up to 8 segments of >= 6 planes (batch 1: 8 lines x 8
    // segments instead of 1 line x 48 planes: a step of the march is one exposed memory latency)
    const int nzt = (D + HS_ZT - 1) / HS_ZT;
    int yt = 8, nseg = 1;
    while (nseg < 8 && D / (nseg * 2) >= 6 && (long)B * 3 * ((D + yt - 1) / yt) * nzt * nseg * 10 < (long)sms * 9) nseg *= 2;
    while (yt > 1 && (long)B * 3 * ((D + yt - 1) / yt) * nzt * nseg * 10 < (long)sms * 9) yt >>= 1;
    const int xseg = (D + nseg - 1) / nseg;
    const dim3 sgrid((unsigned)(B * 3 * ((D + yt - 1) / yt) * nzt), (unsigned)((D + xseg - 1) / xseg));
#define HS_LAUNCH(YT)                                                                                                 \
    e = tc_func_smem(reinterpret_cast<const void*>(head_sum_kernel<YT>), 3 * (YT + 2) * (HS_ZT + 2) * HS_PITCH * 4);  \
    if (e != cudaSuccess) return e;                                                                                   \
    head_sum_kernel<YT><<<sgrid, 256, 3 * (YT + 2) * (HS_ZT + 2) * HS_PITCH * 4, s>>>(P0, P1, P2, b0, b1, b2, out, B, D, xseg)
    switch (yt) {
        case 8: HS_LAUNCH(8); break;
        case 4: HS_LAUNCH(4); break;
        case 2: HS_LAUNCH(2); break;
        default: HS_LAUNCH(1); break;
    }
#undef HS_LAUNCH
    return cudaGetLastError();
}
