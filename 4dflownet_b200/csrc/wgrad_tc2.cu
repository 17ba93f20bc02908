// tcgen05 weight gradient of the 64->64 3x3x3 convolution, single-fp16-operand version (SR4D_OPT_WGRAD_SINGLE, the
// default since round 2): Conv3DBackpropFilter of the layers conv3d() builds (Network/SR4DFlowNet.py:93-108; generated
// by tape.gradient at TrainerController.py:223).
//
//     dW[dx,dy,dz][ci][co] = sum_{b,v} Xpad[b, v + (dx,dy,dz)][ci] * dY[b, v][co]                
//
// Both operands are the hi planes of the split-fp16 tensors (their rounding errors are independent per voxel and
// average out over the 10^5..10^6-voxel sum: tools/gradient_precision_emulation.py, profiles/r02_grad_parity*.txt).
//
// GEMM view (K = voxels, both operands MN-major: a shared-memory row is one voxel's 64 channels = 128 B, straight
// from TMA).  With a single gradient plane the M = 128 rows of the tensor-core instruction would be half empty, so
// two x-planes of dY are stacked on M -- a shifted gradient against the same activations is a neighbouring tap:
//     A (M=128) = [ dY[x-1] tile (co 0..63) ; dY[x] tile (co 0..63) ]      two 8 KB tiles, adjacent ring slots
//     B (N=192) = Xpad[plane, y0+dy.., z0..] three 64-wide atoms one voxel row apart = the three dz taps
//     MMA-1: B = Xpad plane x     rows 0..63  -> sum dY[x-1] Xpad[(x-1)+1] = tap dx=1,  rows 64..127 -> tap dx=0
//     MMA-2: B = Xpad plane x+2   rows 64..127 -> tap dx=2   (rows 0..63 would be "tap 3": never read)
// i.e. 2 instructions per K-step give the 9 (dx,dz) taps of one dy, 24 instructions per 64-voxel tile for all 27 taps
// (the two-plane kernel in wgrad_tc.cu needs 72, its hi-only mode 36).  A CTA of kind dy walks (column, x) units,
// x = 0..D (the extra iteration x = D pairs dY[D-1] with the zero halo plane dY[D]), so every loaded tile is used by
// the following iterations out of shared memory: per iteration ONE new dY tile (8 KB) and ONE new Xpad tile (10 KB)
// arrive for 8 instructions -- 23 B/clk per SM against the ~43 B/clk the L2 delivers chip-wide (the previous kernel
// sat at that limit).
//
// Accumulator chains: the tcgen05 fp32 accumulator truncates after every instruction (measured: ~6e-8 relative loss
// per accumulation), so a chain is cut after FLUSH_ITERS iterations (4 accumulations each): the epilogue warps drain
// TMEM, descale by the gradient's 2^-e and add into the CTA's private fp32 partial in global memory (plain stores the
// first time, RED.ADD afterwards: round-to-nearest fp32, one owner thread per address), while the MMA warp waits --
// ~3 % of the chain time.  A deterministic second stage (reduce_rows) sums the nslab partials.
#include <cuda.h>

#include <cstdio>

#include "conv_tc.h"
#include "tc_host.h"
#include "tc_ptx.cuh"

namespace {

constexpr int TY = 8, TZ = 8, ZP = TZ + 2;
constexpr int DY_SLOT = TY * TZ * 128;              // 8 KB: 64 voxel rows
constexpr int X_SLOT = TY * ZP * 128;               // 10 KB: 8 y-lines x 10 z rows
constexpr int RY = 8;                               // dY ring slots (+ 1 mirror of slot 0 so (prev, cur) are always adjacent)
constexpr int RX = 10;                              // Xpad ring slots
constexpr int SMEM_BYTES = 1024 + (RY + 1) * DY_SLOT + RX * X_SLOT + 512;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NTHREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int FLUSH_ITERS = 96;                     // 384 accumulations per chain
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct W2Params {
    float* partial;          // [nslab][27][64 ci][64 co]
    const int* exp;          // device: exponent of the scaled split gradient (NULL = 0)
    int B, D, nyt, nzt;
    int total;               // (column, x) iterations: B * nyt * nzt * (D + 1)
    int nslab;
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// A CTA's iteration range [t0, t1) over t = column * (D+1) + x is a list of segments (one per column touched).  Both
// the producer and the MMA warp walk it with this helper so their element counters agree.
struct Seg {
    int col, xa, len;        // iterations x = xa .. xa+len-1 of column col
    int nx;                  // Xpad planes xa .. xa+nx-1 are loaded (up to plane D+1)
};
__device__ __forceinline__ Seg seg_at(int t, int t1, int D) {
    Seg s;
    s.col = t / (D + 1);
    s.xa = t - s.col * (D + 1);
    const int xend = min(D + 1, s.xa + (t1 - t));
    s.len = xend - s.xa;
    // planes read: x (MMA-1) and x+2 (MMA-2, while x+2 <= D+1) for x = xa..xend-1
    const int last = s.xa <= D - 1 ? min(xend + 1, D + 1) : D;
    s.nx = last - s.xa + 1;
    return s;
}

__global__ void __launch_bounds__(NTHREADS, 1)
wgrad64_tc2_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap gmap, W2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ysm = smem;                                   // (RY + 1) x 8 KB
    uint8_t* xsm = smem + (RY + 1) * DY_SLOT;              // RX x 10 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(xsm + RX * X_SLOT);
    uint64_t* y_full = bars;                 // [RY]
    uint64_t* y_empty = y_full + RY;         // [RY]
    uint64_t* x_full = y_empty + RY;         // [RX]
    uint64_t* x_empty = x_full + RX;         // [RX]
    uint64_t* acc_full = x_empty + RX;       // [1]
    uint64_t* acc_empty = acc_full + 1;      // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dy = blockIdx.x;                   // CTA kind
    const int slab = blockIdx.y;
    const int t0 = (int)((long long)p.total * slab / p.nslab);
    const int t1 = (int)((long long)p.total * (slab + 1) / p.nslab);
    const int D = p.D;
    const int cols_per_b = p.nyt * p.nzt;

    if (threadIdx.x == 0) {
        for (int i = 0; i < RY; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); }
        for (int i = 0; i < RX; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, NUM_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&xmap);
        prefetch_tmap(&gmap);
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

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t eY = 0, eX = 0;                  // elements issued so far
            auto load_y = [&](int b, int y0, int z0, int plane) {
                const uint32_t s = eY % RY, ph = (eY / RY) & 1;
                mbar_wait(&y_empty[s], ph ^ 1);
                mbar_expect_tx(&y_full[s], s == 0 ? 2 * DY_SLOT : DY_SLOT);
                // interior voxel (plane, y0.., z0..) sits at +2 in the zero-haloed gradient tensor (plane in -1..D)
                tma_load_5d(ysm + s * DY_SLOT, &gmap, &y_full[s], 0, z0 + 2, y0 + 2, plane + 2, b);
                if (s == 0) tma_load_5d(ysm + RY * DY_SLOT, &gmap, &y_full[s], 0, z0 + 2, y0 + 2, plane + 2, b);
                ++eY;
            };
            auto load_x = [&](int b, int y0, int z0, int plane) {
                const uint32_t s = eX % RX, ph = (eX / RX) & 1;
                mbar_wait(&x_empty[s], ph ^ 1);
                mbar_expect_tx(&x_full[s], X_SLOT);
                // padded coordinates: plane in 0..D+1, lines y0+dy.., rows z0..z0+9
                tma_load_5d(xsm + s * X_SLOT, &xmap, &x_full[s], 0, z0, y0 + dy, plane, b);
                ++eX;
            };
            for (int t = t0; t < t1;) {
                const Seg sg = seg_at(t, t1, D);
                const int b = sg.col / cols_per_b;
                const int rem = sg.col % cols_per_b;
                const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
                load_y(b, y0, z0, sg.xa - 1);
                load_x(b, y0, z0, sg.xa);
                if (sg.nx > 1) load_x(b, y0, z0, sg.xa + 1);     // nx == 1 only for a segment that is the lone iteration x == D
                for (int i = 0; i < sg.len; ++i) {
                    load_y(b, y0, z0, sg.xa + i);
                    if (i + 2 < sg.nx) load_x(b, y0, z0, sg.xa + i + 2);
                }
                t += sg.len;
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // D=f32, A=B=f16, both MN-major, M=128, N=192
            const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t d1 = tmem_base, d2 = tmem_base + 192;
            uint32_t bY = 0, bX = 0;                  // element index of the current segment's element 0
            uint32_t wY = 0, wX = 0;                  // elements whose full barrier has been waited for
            int chain_it = 0;                         // iterations issued into the current accumulator chain
            uint32_t nchain = 0;                      // chains completed
            auto need_y = [&](uint32_t upto) {        // wait until dY elements [0, upto] have landed
                while (wY <= upto) { mbar_wait(&y_full[wY % RY], (wY / RY) & 1); ++wY; }
            };
            auto need_x = [&](uint32_t upto) {
                while (wX <= upto) { mbar_wait(&x_full[wX % RX], (wX / RX) & 1); ++wX; }
            };
            for (int t = t0; t < t1;) {
                const Seg sg = seg_at(t, t1, D);
                for (int i = 0; i < sg.len; ++i) {
                    const uint32_t gcur = bY + i + 1;                       // dY[x]; dY[x-1] is gcur - 1
                    const int k2 = (i + 2 < sg.nx) ? i + 2 : i;            // x == D: MMA-2 re-uses plane x (rows 64..127 = zero plane)
                    need_y(gcur);
                    need_x(bX + (uint32_t)max(i, k2));
                    if (chain_it == 0 && nchain > 0) mbar_wait(acc_empty, (nchain - 1) & 1);
                    tc_fence_after();
                    const uint32_t sc = gcur % RY;
                    const uint32_t a0 = smem_u32(ysm + (sc == 0 ? RY - 1 : sc - 1) * DY_SLOT);   // (prev, cur) adjacent; slot 0's mirror sits behind slot RY-1
                    const uint32_t x1 = smem_u32(xsm + ((bX + i) % RX) * X_SLOT);
                    const uint32_t x2 = smem_u32(xsm + ((bX + k2) % RX) * X_SLOT);
#pragma unroll
                    for (int j = 0; j < TY / 2; ++j) {
                        const uint64_t ad = desc_mn(a0 + j * 2 * TZ * 128, DY_SLOT, TZ * 128);
                        const uint32_t acc = (chain_it | j) != 0;
                        tc_mma_f16(d1, ad, desc_mn(x1 + j * 2 * ZP * 128, 128, ZP * 128), idesc, acc);
                        tc_mma_f16(d2, ad, desc_mn(x2 + j * 2 * ZP * 128, 128, ZP * 128), idesc, acc);
                    }
                    // releases: dY element i of the segment (last read as "previous" here), Xpad element i (MMA-1), and the
                    // elements only MMA-2 reads at the end of a segment
                    tc_commit(&y_empty[(bY + i) % RY]);
                    if (i == sg.len - 1) tc_commit(&y_empty[(bY + i + 1) % RY]);
                    tc_commit(&x_empty[(bX + i) % RX]);
                    if (i == sg.len - 1)
                        for (int k = sg.len; k < sg.nx; ++k) tc_commit(&x_empty[(bX + k) % RX]);
                    ++chain_it;
                    const bool last = (t + i + 1 == t1);
                    if (chain_it == FLUSH_ITERS || last) {
                        tc_commit(acc_full);
                        chain_it = 0;
                        ++nchain;
                    }
                }
                bY += sg.len + 1;
                bX += sg.nx;
                t += sg.len;
            }
        }
    } else {
        // ================= epilogue / chain flush (warps 2..9) =================
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int hsel = (warp - 2) >> 2;                // which 96 of the 192 columns
        const int m = 32 * q + lane;                     // accumulator row
        const int co = m & 63;
        const bool upper = m >= 64;                      // rows 64..127: dY[x] -> taps dx=0 (acc 1) and dx=2 (acc 2); rows 0..63: dx=1
        const float descale = p.exp ? exp2f(-(float)*p.exp) : 1.f;
        float* out = p.partial + (size_t)slab * 27 * 4096 + co;
        const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16);
        const int nit = t1 - t0;
        const int nchains = (nit + FLUSH_ITERS - 1) / FLUSH_ITERS;
        auto flush = [&](uint32_t col0, int dx, bool first) {
            // columns n = dz*64 + ci of this accumulator -> partial[(dx*3+dy)*3+dz][ci][co]
#pragma unroll 1
            for (int c0 = 96 * hsel; c0 < 96 * hsel + 96; c0 += 16) {
                float a[16];
                tc_ld16(trow + col0 + c0, a);
                tc_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = c0 + j;
                    float* o = out + (size_t)((dx * 3 + dy) * 3 + (n >> 6)) * 4096 + (n & 63) * 64;
                    const float v = a[j] * descale;
                    if (first) *o = v;
                    else atomicAdd(o, v);
                }
            }
        };
        if (nchains == 0) {
            // empty range (more slabs than work): the second stage still sums this partial
            for (int dxs = 0; dxs < 3; ++dxs) {
                if ((dxs == 1) == upper) continue;
                for (int n = 96 * hsel; n < 96 * hsel + 96; ++n)
                    out[(size_t)((dxs * 3 + dy) * 3 + (n >> 6)) * 4096 + (n & 63) * 64] = 0.f;
            }
        }
        for (int c = 0; c < nchains; ++c) {
            mbar_wait(acc_full, c & 1);
            tc_fence_after();
            flush(0, upper ? 0 : 1, c == 0);
            if (upper) flush(192, 2, c == 0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

int tc_wgrad2_slabs(int B, int D) {
    const int nyt = (D + TY - 1) / TY, nzt = (D + TZ - 1) / TZ;
    const long total = (long)B * nyt * nzt * (D + 1);
    long n = tc_num_sms() / 3;              // three CTA kinds (dy) per slab, one CTA per SM
    if (n > total) n = total;
    if (n < 1) n = 1;
    return (int)n;
}

cudaError_t tc_wgrad64_single(ActView x, const __half* dy_split, const int* dy_exp, float* partial, cudaStream_t s) {
    const int B = x.B, D = x.D;
    cudaError_t e = tc_func_smem(reinterpret_cast<const void*>(wgrad64_tc2_kernel), SMEM_BYTES);
    if (e != cudaSuccess) return e;
    CUtensorMap xmap, gmap;
    // hi planes only: the maps cover the first B "samples" of the packed [2B] plane arrays
    if (!tc_make_act_map(&xmap, x.hi, B, D + 2, TY, ZP)) return cudaErrorUnknown;
    if (!tc_make_act_map(&gmap, dy_split, B, D + 4, TY, TZ)) return cudaErrorUnknown;
    W2Params p;
    p.partial = partial; p.exp = dy_exp; p.B = B; p.D = D;
    p.nyt = (D + TY - 1) / TY; p.nzt = (D + TZ - 1) / TZ;
    p.total = B * p.nyt * p.nzt * (D + 1);
    p.nslab = tc_wgrad2_slabs(B, D);
    dim3 grid(3, p.nslab);
    wgrad64_tc2_kernel<<<grid, NTHREADS, SMEM_BYTES, s>>>(xmap, gmap, p);
    return cudaGetLastError();
}
