// tcgen05 weight gradient of the 64->64 3x3x3 convolution, single-fp16-operand version (SR4D_OPT_WGRAD_SINGLE, the
// default since round 2): Conv3DBackpropFilter of the layers conv3d() builds (Network/SR4DFlowNet.py:93-108; generated
// by tape.gradient at TrainerController.py:223).
//
//     dW[dx,dy,dz][ci][co] = sum_{b,v} Xpad[b, v + (dx,dy,dz)][ci] * dY[b, v][co]                
//
// Both operands are the hi planes of the split-fp16 tensors (their rounding errors are independent per voxel and
// average out over the 10^5..10^6-voxel sum: tools/gradient_precision_emulation.py, profiles/r02_grad_parity.txt).
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
// x = 0..D (the extra iteration x = D pairs dY[D-1] with the zero halo plane dY[D]); every loaded plane tile is used by
// the following iterations out of shared memory, so per iteration ONE new dY tile (8 KB) and ONE new Xpad tile (10 KB)
// are needed for 8 instructions.
//
// Data movement.  The tiles arrive in GROUPS of four consecutive x-planes, one TMA box per tensor and group (32 KB of
// dY, 40 KB of Xpad): the first version issued one 8 / 10 KB box per plane and ran at half speed -- the MMA warp spent
// 35 % of its cycles waiting for tiles that had been requested five iterations earlier, because the TMA unit works
// through small boxes nearly one memory latency at a time (SR4D_TC_DEBUG stamps, profiles/r02_wgrad2_notes.txt).  Three
// groups form the ring (12 plane slots per tensor, + a copy of slot 0 behind slot 11 so the (previous, current) pair of
// dY tiles is always adjacent for the A descriptor); a group is handed back with ONE tcgen05.commit after the
// iteration that read its last plane (a commit costs the tensor pipe ~100 cycles: tools/probe/wgrad_probe.cu).
//
// Accumulator chains: the tcgen05 fp32 accumulator truncates after every instruction (measured: ~6e-8 relative loss
// per accumulation), so a chain is cut after FLUSH_ITERS iterations (4 accumulations each): the epilogue warps drain
// TMEM, descale by the gradient's 2^-e and add into the CTA's private fp32 partial in global memory (plain stores the
// first time, RED.ADD afterwards: round-to-nearest fp32, one owner thread per address), while the MMA warp waits --
// ~3 % of the chain time.  A deterministic second stage (reduce_rows) sums the nslab partials.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv_tc.h"
#include "tc_host.h"
#include "tc_ptx.cuh"

namespace {

constexpr int TY = 8, TZ = 8, ZP = TZ + 2;
constexpr int DY_SLOT = TY * TZ * 128;              // 8 KB: 64 voxel rows
constexpr int X_SLOT = TY * ZP * 128;               // 10 KB: 8 y-lines x 10 z rows
constexpr int GP = 4;                               // planes per group (one TMA box per tensor)
constexpr int NGR = 3;                              // groups in the ring
constexpr int NSLOT = NGR * GP;                     // plane slots per tensor
constexpr int SMEM_BYTES = 1024 + (NSLOT + 1) * DY_SLOT + NSLOT * X_SLOT + 512;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NTHREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int FLUSH_ITERS = 96;                     // 384 accumulations per chain
constexpr int GROUP_BYTES = GP * (DY_SLOT + X_SLOT);
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct W2Params {
    float* partial;          // [nslab][27][64 ci][64 co]
    const int* exp;          // device: exponent of the scaled split gradient (NULL = 0)
    int B, D, nyt, nzt;
    int total;               // (column, x) iterations: B * nyt * nzt * (D + 1)
    int nslab;
    int flush_iters;         // accumulator chain length in iterations (FLUSH_ITERS; SR4D_WGRAD_FLUSH overrides it for experiments)
    long long* dbg;          // SR4D_TC_DEBUG=1: per-CTA cycles of the MMA warp {tile waits, acc_empty waits, issue (MMAs + commits), total, iterations}
};

// Batched launch (BATCH instantiation): the weight gradients of `nlayers` layers on the same grid in one launch.  The
// layers are independent (every dY and X tensor exists by the end of the backward pass), so the roles simply walk the
// layers with their ring / chain counters running on: what is saved per layer is the launch, the prologue, the pipeline
// fill and the drain -- 25 of the 47 us of a 24^3 layer at batch 8.
struct W2Layer {
    CUtensorMap xmap, gmap, gmap1;
    float* partial;
    const int* exp;
    long long pad[6];
};
static_assert(sizeof(W2Layer) % 64 == 0, "tensor maps in global memory must stay 64-byte aligned");

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// A CTA's iteration range [t0, t1) over t = column * (D+1) + x is a list of segments (one per column touched):
// iterations x = xa .. xa+len-1 of one column.  Element k of a segment is dY plane xa-1+k and Xpad plane xa+k; iteration i
// reads dY elements i (previous) and i+1 (current) and Xpad elements i (MMA-1) and i+2 (MMA-2; at x == D, where plane D+2
// does not exist and the current dY plane is the zero halo, MMA-2 re-reads element i).  Elements 0 .. len+1 are loaded,
// four per group; every segment starts on a fresh group.
__device__ __forceinline__ int seg_len(int xa, int remaining, int D) { return min(D + 1 - xa, remaining); }
__device__ __forceinline__ int seg_groups(int len) { return (len + 2 + GP - 1) / GP; }

template <bool BATCH>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad64_tc2_kernel(const __grid_constant__ CUtensorMap xmap0, const __grid_constant__ CUtensorMap gmap0,
                   const __grid_constant__ CUtensorMap gmap10, W2Params p, const W2Layer* __restrict__ layers, int nlayers) {
    const int nl = BATCH ? nlayers : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ysm = smem;                                   // (NSLOT + 1) x 8 KB; slot NSLOT mirrors slot 0
    uint8_t* xsm = smem + (NSLOT + 1) * DY_SLOT;           // NSLOT x 10 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(xsm + NSLOT * X_SLOT);
    uint64_t* full = bars;                   // [NGR]
    uint64_t* empty = full + NGR;            // [NGR]
    uint64_t* acc_full = empty + NGR;        // [1]
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
        for (int i = 0; i < NGR; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, NUM_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!BATCH) {
            prefetch_tmap(&xmap0);
            prefetch_tmap(&gmap0);
            prefetch_tmap(&gmap10);
        }
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
            uint32_t gq = 0;                          // groups issued so far
            for (int L = 0; L < nl; ++L) {
            const CUtensorMap* xmapp = BATCH ? &layers[L].xmap : &xmap0;
            const CUtensorMap* gmapp = BATCH ? &layers[L].gmap : &gmap0;
            const CUtensorMap* gmap1p = BATCH ? &layers[L].gmap1 : &gmap10;
            int col = t0 / (D + 1);
            int xa = t0 - col * (D + 1);
            for (int t = t0; t < t1;) {
                const int len = seg_len(xa, t1 - t, D);
                const int ngrp = seg_groups(len);
                const int b = col / cols_per_b;
                const int rem = col - b * cols_per_b;
                const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
                for (int g = 0; g < ngrp; ++g, ++gq) {
                    const uint32_t r = gq % NGR, ph = (gq / NGR) & 1;
                    mbar_wait(&empty[r], ph ^ 1);
                    mbar_expect_tx(&full[r], GROUP_BYTES + (r == 0 ? DY_SLOT : 0));
                    // dY planes xa-1+4g ..+3: interior voxel (plane, y0.., z0..) sits at +2 in the zero-haloed gradient tensor;
                    // Xpad planes xa+4g ..+3 (padded coordinates), lines y0+dy.., rows z0..z0+9.  Planes past the tensor are
                    // zero-filled by the TMA unit and never read.
                    tma_load_5d(ysm + r * GP * DY_SLOT, gmapp, &full[r], 0, z0 + 2, y0 + 2, xa - 1 + GP * g + 2, b);
                    if (r == 0) tma_load_5d(ysm + NSLOT * DY_SLOT, gmap1p, &full[r], 0, z0 + 2, y0 + 2, xa - 1 + GP * g + 2, b);
                    tma_load_5d(xsm + r * GP * X_SLOT, xmapp, &full[r], 0, z0, y0 + dy, xa + GP * g, b);
                }
                t += len;
                xa = 0;                                // every further segment starts a new column
                ++col;
            }
            }   // layers
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp runs the loop converged and one ELECTED lane issues: with `if (lane == 0)` around the loop ptxas wraps
        // every UTCHMMA in an ELECT / BRA.U.ANY loop with R2UR moves (~15 instructions per MMA; eight MMAs per iteration
        // made the issuing thread the bottleneck: 1114 cycles per iteration against 768 of tensor-pipe time).
        {
            const bool leader = elect_one();
            // D=f32, A=B=f16, both MN-major, M=128, N=192
            const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t d1 = tmem_base, d2 = tmem_base + 192;
            const uint64_t adesc0 = desc_mn(0, DY_SLOT, TZ * 128);            // address field filled in per instruction
            const uint64_t bdesc0 = desc_mn(0, 128, ZP * 128);
            const uint32_t ybase = smem_u32(ysm) >> 4, xbase = smem_u32(xsm) >> 4;
            uint32_t gq = 0;                          // group counter at the start of the current segment
            uint32_t wq = 0;                          // groups whose full barrier has been waited for
            int chain_it = 0;                         // iterations issued into the current accumulator chain
            uint32_t nchain = 0;                      // chains completed
            long long dwf = 0, dwa = 0, dis = 0;
            const long long tbeg = p.dbg ? clock64() : 0;
            for (int L = 0; L < nl; ++L) {
            int xa = t0 % (D + 1);
            for (int t = t0; t < t1;) {
                const int len = seg_len(xa, t1 - t, D);
                const int ngrp = seg_groups(len);
                const bool has_x2_last = xa + len - 1 + 2 <= D + 1;         // false only when the segment ends at x == D
                uint32_t sprev = (gq % NGR) * GP;                           // ring slot of element i (starts at element 0)
                for (int i = 0; i < len; ++i) {
                    const int k2 = (i < len - 1 || has_x2_last) ? i + 2 : i;
                    long long c0 = p.dbg ? clock64() : 0, c1;
                    const uint32_t gneed = gq + (uint32_t)(k2 > i + 1 ? k2 : i + 1) / GP;
                    while (wq <= gneed) { mbar_wait(&full[wq % NGR], (wq / NGR) & 1); ++wq; }
                    if (p.dbg) { c1 = clock64(); dwf += c1 - c0; c0 = c1; }
                    if (chain_it == 0 && nchain > 0) mbar_wait(acc_empty, (nchain - 1) & 1);
                    if (p.dbg) { c1 = clock64(); dwa += c1 - c0; c0 = c1; }
                    tc_fence_after();
                    uint32_t s2 = sprev + (uint32_t)(k2 - i);               // slot of Xpad element k2
                    if (s2 >= NSLOT) s2 -= NSLOT;
                    // (previous, current) dY tiles are adjacent slots; when the current tile sits in slot 0 (sprev == NSLOT-1)
                    // the A descriptor's second atom lands on the copy of slot 0 kept behind slot NSLOT-1
                    const uint32_t a0 = ybase + sprev * (DY_SLOT >> 4);
                    const uint32_t x1 = xbase + sprev * (X_SLOT >> 4);
                    const uint32_t x2 = xbase + s2 * (X_SLOT >> 4);
#pragma unroll
                    for (int j = 0; j < TY / 2; ++j) {
                        const uint64_t ad = adesc0 | (uint64_t)(a0 + j * (2 * TZ * 128 >> 4));
                        const uint32_t acc = (chain_it | j) != 0;
                        if (leader) tc_mma_f16(d1, ad, bdesc0 | (uint64_t)(x1 + j * (2 * ZP * 128 >> 4)), idesc, acc);
                        // rows 0..63 of this product ("tap 3") are never read: their output lanes are disabled, which the
                        // power probe prices at -23 % of the instruction's energy (profiles/r02_power_probe.txt)
                        if (leader) tc_mma_f16_masked(d2, ad, bdesc0 | (uint64_t)(x2 + j * (2 * ZP * 128 >> 4)), idesc, acc, ~0u, ~0u, 0u, 0u);
                    }
                    // hand a group back after the iteration that read its last plane (dY as "previous", Xpad in MMA-1), the
                    // rest of the segment's groups after its last iteration
                    if ((i & (GP - 1)) == GP - 1 && leader) tc_commit(&empty[(gq + i / GP) % NGR]);
                    if (i == len - 1) {
                        // (a trailing group may hold only planes nobody read: see its load land before giving it back)
                        while (wq < gq + (uint32_t)ngrp) { mbar_wait(&full[wq % NGR], (wq / NGR) & 1); ++wq; }
                        for (int g = len / GP; g < ngrp; ++g)
                            if (leader) tc_commit(&empty[(gq + g) % NGR]);
                    }
                    if (p.dbg) dis += clock64() - c0;
                    ++chain_it;
                    if (chain_it == p.flush_iters || t + i + 1 == t1) {
                        if (leader) tc_commit(acc_full);
                        chain_it = 0;
                        ++nchain;
                    }
                    __syncwarp();
                    if (++sprev == NSLOT) sprev = 0;
                }
                gq += ngrp;
                t += len;
                xa = 0;
            }
            }   // layers (every layer ends its accumulator chain: t + i + 1 == t1 above)
            if (p.dbg && leader) {
                long long* d = p.dbg + (size_t)(blockIdx.y * 3 + blockIdx.x) * 8;
                d[0] = dwf; d[1] = dwa; d[2] = dis; d[3] = clock64() - tbeg; d[4] = t1 - t0;
            }
        }
    } else {
        // ================= epilogue / chain flush (warps 2..9) =================
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int hsel = (warp - 2) >> 2;                // which 96 of the 192 columns
        const int m = 32 * q + lane;                     // accumulator row
        const int co = m & 63;
        const bool upper = m >= 64;                      // rows 64..127: dY[x] -> taps dx=0 (acc 1) and dx=2 (acc 2); rows 0..63: dx=1
        const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16);
        const int nit = t1 - t0;
        const int nchains = (nit + p.flush_iters - 1) / p.flush_iters;
        uint32_t cbase = 0;                              // chains drained so far (parity of acc_full runs across layers)
        for (int L = 0; L < nl; ++L) {
        const int* expp = BATCH ? layers[L].exp : p.exp;
        const float descale = expp ? exp2f(-(float)*expp) : 1.f;
        float* out = (BATCH ? layers[L].partial : p.partial) + (size_t)slab * 27 * 4096 + co;
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
            mbar_wait(acc_full, (cbase + c) & 1);
            tc_fence_after();
            flush(0, upper ? 0 : 1, c == 0);
            if (upper) flush(192, 2, c == 0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
        cbase += (uint32_t)nchains;
        }   // layers
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
    cudaError_t e = tc_func_smem(reinterpret_cast<const void*>(wgrad64_tc2_kernel<false>), SMEM_BYTES);
    if (e != cudaSuccess) return e;
    CUtensorMap xmap, gmap, gmap1;
    // hi planes only: the maps cover the packed [2B] plane arrays, the kernel addresses samples b < B
    if (!tc_make_act_map(&xmap, x.hi, B, D + 2, TY, ZP, GP)) return cudaErrorUnknown;
    if (!tc_make_act_map(&gmap, dy_split, B, D + 4, TY, TZ, GP)) return cudaErrorUnknown;
    if (!tc_make_act_map(&gmap1, dy_split, B, D + 4, TY, TZ, 1)) return cudaErrorUnknown;
    W2Params p;
    p.partial = partial; p.exp = dy_exp; p.B = B; p.D = D;
    p.nyt = (D + TY - 1) / TY; p.nzt = (D + TZ - 1) / TZ;
    p.total = B * p.nyt * p.nzt * (D + 1);
    p.nslab = tc_wgrad2_slabs(B, D);
    static const int flush_env = getenv("SR4D_WGRAD_FLUSH") ? atoi(getenv("SR4D_WGRAD_FLUSH")) : 0;
    p.flush_iters = flush_env > 0 ? flush_env : FLUSH_ITERS;
    dim3 grid(3, p.nslab);
    static const bool debug = getenv("SR4D_TC_DEBUG") != nullptr;
    static long long* dbg_buf = nullptr;
    p.dbg = nullptr;
    if (debug) {
        if (!dbg_buf) cudaMalloc((void**)&dbg_buf, 3 * 64 * 8 * sizeof(long long));
        cudaMemsetAsync(dbg_buf, 0, 3 * 64 * 8 * sizeof(long long), s);
        p.dbg = dbg_buf;
    }
    wgrad64_tc2_kernel<false><<<grid, NTHREADS, SMEM_BYTES, s>>>(xmap, gmap, gmap1, p, nullptr, 0);
    cudaError_t e2 = cudaGetLastError();
    if (debug && e2 == cudaSuccess) {
        long long hb[3 * 64 * 8];
        cudaStreamSynchronize(s);
        cudaMemcpy(hb, dbg_buf, sizeof hb, cudaMemcpyDeviceToHost);
        double a[5] = {0, 0, 0, 0, 0};
        const int n = 3 * p.nslab;
        for (int i = 0; i < n; ++i) for (int k = 0; k < 5; ++k) a[k] += (double)hb[i * 8 + k] / n;
        fprintf(stderr, "[wgrad2 dbg] D=%d B=%d: MMA-warp cycles avg/CTA: tile waits %.0f  acc_empty %.0f  issue %.0f  of total %.0f for %.0f iterations (%.0f per iteration)\n",
                D, B, a[0], a[1], a[2], a[3], a[4], a[3] / (a[4] > 0 ? a[4] : 1));
    }
    return e2;
}

// ---- batched launch -----------------------------------------------------------------------------------------------------
struct TcWgradBatch {
    int n = 0, B = 0, D = 0;
    W2Layer* dlayers = nullptr;
};

cudaError_t tc_wgrad_batch_build(const TcWgradItem* items, int n, TcWgradBatch** out) {
    if (n < 1) return cudaErrorInvalidValue;
    const int B = items[0].x.B, D = items[0].x.D;
    std::vector<W2Layer> hl(n);
    for (int i = 0; i < n; ++i) {
        if (items[i].x.B != B || items[i].x.D != D) return cudaErrorInvalidValue;
        memset(&hl[i], 0, sizeof(W2Layer));
        if (!tc_make_act_map(&hl[i].xmap, items[i].x.hi, B, D + 2, TY, ZP, GP)) return cudaErrorUnknown;
        if (!tc_make_act_map(&hl[i].gmap, items[i].dy_split, B, D + 4, TY, TZ, GP)) return cudaErrorUnknown;
        if (!tc_make_act_map(&hl[i].gmap1, items[i].dy_split, B, D + 4, TY, TZ, 1)) return cudaErrorUnknown;
        hl[i].partial = items[i].partial;
        hl[i].exp = items[i].dy_exp;
    }
    TcWgradBatch* b = new TcWgradBatch();
    b->n = n; b->B = B; b->D = D;
    cudaError_t e = cudaMalloc((void**)&b->dlayers, n * sizeof(W2Layer));
    if (e == cudaSuccess) e = cudaMemcpy(b->dlayers, hl.data(), n * sizeof(W2Layer), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { tc_wgrad_batch_free(b); return e; }
    *out = b;
    return cudaSuccess;
}
void tc_wgrad_batch_free(TcWgradBatch* b) {
    if (!b) return;
    cudaFree(b->dlayers);
    delete b;
}
cudaError_t tc_wgrad_batch_launch(TcWgradBatch* b, cudaStream_t s) {
    cudaError_t e = tc_func_smem(reinterpret_cast<const void*>(wgrad64_tc2_kernel<true>), SMEM_BYTES);
    if (e != cudaSuccess) return e;
    const int B = b->B, D = b->D;
    W2Params p;
    p.partial = nullptr; p.exp = nullptr; p.B = B; p.D = D;
    p.nyt = (D + TY - 1) / TY; p.nzt = (D + TZ - 1) / TZ;
    p.total = B * p.nyt * p.nzt * (D + 1);
    p.nslab = tc_wgrad2_slabs(B, D);
    static const int flush_env = getenv("SR4D_WGRAD_FLUSH") ? atoi(getenv("SR4D_WGRAD_FLUSH")) : 0;
    p.flush_iters = flush_env > 0 ? flush_env : FLUSH_ITERS;
    p.dbg = nullptr;
    CUtensorMap dummy;
    memset(&dummy, 0, sizeof dummy);
    wgrad64_tc2_kernel<true><<<dim3(3, p.nslab), NTHREADS, SMEM_BYTES, s>>>(dummy, dummy, dummy, p, b->dlayers, b->n);
    return cudaGetLastError();
}
