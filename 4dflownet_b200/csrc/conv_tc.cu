// tcgen05 (5th-gen tensor core) 64->64 3x3x3 convolution for sm_100a, split-fp16 "fp32-accurate".
//
// Reference op: conv3d() of Network/SR4DFlowNet.py:93-108 / resnet_block :111-120 (30 of the
// 36 layers, 99.5 % of the FLOPs); with the dgrad weight image the same kernel evaluates
// Conv3DBackpropInput on the padded grid.
//
// Formulation (implicit GEMM, weights as the M operand so the voxel dimension is the flexible N):
//     D[128 x N] += A_tap[128 x 64] * B_tap[N x 64]^T        for the 27 taps, K = 64 channels
//   Operands are split in fp16: W = Whi + Wlo/2048, X = Xhi + Xlo/2048.  ONE accumulator per tile:
//     TMEM lanes 0..63   ("L", lane = co)      accumulate  Wlo*Xhi + Whi*Xlo   (both scaled by 2048)
//     TMEM lanes 64..127 ("H", lane = 64 + co) accumulate  Whi*Xhi
//   through two MMAs per K-step that share one 16 KB weight image [Wlo rows ; Whi rows]:
//     MMA_a: A = image            = [Wlo ; Whi],  B = the activation's hi plane
//     MMA_b: A = image + 8 KB     = [Whi ; the 8 KB that follow], B = the activation's lo plane, issued with the
//            disable-output-lane vector masking lanes 64..127, so whatever the upper 64 rows read is never
//            accumulated (no zero block behind the image: four 16 KB weight stages fit where three 24 KB did)
//   out[co][v] = D[H] + D[L]/2048 (the Wlo*Xlo term, 2^-22 relative, is dropped).  With a single
//   accumulator of N <= 208 columns the 512 TMEM columns hold TWO tiles, so the epilogue of tile i
//   overlaps the MMAs of tile i+1 (the previous two-accumulator version stalled the tensor pipe for
//   ~30 % of every tile while the epilogue drained TMEM).
//   B_tap rows = the N = 8*TY voxels of a (TY y-lines x 8 z) tile at a fixed x, 64 channels each.
//
// Data movement: ONE TMA box per dx loads the (TY+2) x 10 voxel halo plane (hi and lo) into
// shared memory as dense 128-byte rows (SWIZZLE_128B).  All nine (dy,dz) taps of that plane are
// then plain descriptor start-address shifts of (dy*10+dz) rows with an 8-row-group stride of
// 1280 B: the tensor core applies the 128B swizzle on absolute shared-memory address bits, so
// shifted starts and a non-1024 group stride read the TMA-written rows correctly (verified on
// B200 by tools/probe/mma_probe.cu).  Every activation byte is fetched 3x from L2 per layer
// (plus y/z halo), weights stream per tap as pre-swizzled 16 KB images (cp.async.bulk).
// Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue:
// TMEM -> registers -> fp32 shared-memory staging (H part + bias, L part / 2048) -> coalesced
// 16-byte residual loads / activation / fp16 split / stores with the replicate halo.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "conv_tc.h"
#include "tc_host.h"
#include "tc_ptx.cuh"

namespace {

constexpr int W_TAP_BYTES = 128 * 64 * 2;            // 16 KB: [Wlo ; Whi] x 64 ci, fp16, swizzled
constexpr int W_HI_OFFSET = 64 * 64 * 2;             // the Whi rows start 8 KB into the image
constexpr int W_STAGE_BYTES = W_TAP_BYTES;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI = NUM_EPI_WARPS * 32;   // epilogue threads
constexpr int NUM_THREADS = 64 + NUM_EPI;
constexpr int TZ = 8;                       // voxels per 8-row group (one z run)
constexpr int ZP = TZ + 2;                  // plane row pitch in voxels
constexpr int TMEM_COLS = 512;              // two accumulators of up to 256 columns
constexpr int ACC_STRIDE = 256;             // column offset of the second accumulator

template <int TY>
struct Cfg {
    static constexpr int N = TY * TZ;                                  // voxels per tile = MMA N
    static constexpr int ROWS = (TY + 2) * ZP;                         // rows of one staged plane part
    static constexpr int PART_BYTES = (ROWS * 128 + 1023) / 1024 * 1024;
    static constexpr int XSTAGE_BYTES = 2 * PART_BYTES;                // hi part | lo part
    static constexpr int NXS = 2;
    static constexpr int NWS = 4;
    static constexpr int LPC = (TY % 4 == 0) ? 4 : 3;                  // y-lines per epilogue chunk
    static constexpr int CV = LPC * TZ;                                // voxels (TMEM columns) per chunk
    static constexpr int NCHUNK = (TY + LPC - 1) / LPC;
    static constexpr int STAGE_FLOATS = 2 * CV * 64;                   // [H | L][voxel][channel]
    static constexpr int SMEM_BYTES = 1024 + NXS * XSTAGE_BYTES + NWS * W_STAGE_BYTES + STAGE_FLOATS * 4 + 256;
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N constraint for M=128");
    static_assert(CV == 32 || CV == 24, "epilogue chunk shapes");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct KParams {
    const __half* w_img;     // [27][128][64] fp16 pre-swizzled, tap order = Keras (dx*3+dy)*3+dz
    __half* out_hi;
    __half* out_lo;
    const __half* res_hi;
    const __half* res_lo;
    const float* bias;
    float* out_raw;          // optional fp32 [B][Do^3][64] output instead of Act
    unsigned int* absmax;    // optional: atomicMax of |out_raw| bit patterns
    float slope;
    int B, Do, halo;
    int nyt, nzt, ntiles;
    int nx;                  // x planes that get tiles (Do, or the interior edge in fused-dgrad mode)
    // fused dgrad: the padded-grid result is folded onto the interior voxels inside the kernel
    // (x by re-indexing the boundary planes' (input plane, weight slice) pairs, y/z in the epilogue staging), then
    // out = (fold * 2^-e + add_pre) * act'(saved) + add_post goes to the interior of an fp32 G4 tensor
    int fused, Dint;
    const int* dy_exp;
    const float* add_pre;
    const float* add_post;
    const __half* sav_hi;
    const __half* sav_lo;
    float* out_g4;
    // optional split-fp16 copy of out_g4 written by the same epilogue, scaled by 2^e with e derived from a
    // rigorous bound on |out|: 8 (fold multiplicity) * gain (max_ci sum_{t,co} |W|) * max|dy| + max|add_pre|
    __half* split_hi;
    __half* split_lo;
    int* split_exp;
    const float* gain;
    const unsigned int* dy_amax;
    const unsigned int* add_amax;
    int single_b;            // 1: the B operand has no lo part (dgrad with a single-fp16 gradient): no second MMA, no lo load
    int xsplit;              // producer order: 1 = next activation plane requested mid-pass (default), 0 = at the pass boundary (SR4D_TC_XSPLIT=0)
    int exp_skip;            // TIMING EXPERIMENT ONLY (SR4D_TC_EXP_SKIP, wrong results): bit 0 = re-use stale weight slots
                             // after the first fill, bit 1 = re-use stale activation stages (how much do the L2 streams cost?)
    long long* dbg;          // SR4D_TC_DEBUG=1: per-CTA cycles the MMA warp waited on {t_empty, x_full, w_full} and its total
};

// element offset of channel 0 of interior voxel (x,y,z) of an fp32 G4 tensor [B][D+4]^3[64]
__device__ __forceinline__ size_t g4_off(int D, int b, int x, int y, int z) {
    const int dp = D + 4;
    return ((((size_t)b * dp + (x + 2)) * dp + (y + 2)) * dp + (z + 2)) * 64;
}

__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int TY>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv64_tc_kernel(const __grid_constant__ CUtensorMap xmap, KParams p) {
    using C = Cfg<TY>;
    extern __shared__ uint8_t smem_raw[];
    // align by OFFSET (not by pointer cast) so the compiler keeps the shared address space (STS/LDS, not generic)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* xs = smem;                                   // NXS x [hi part | lo part]
    uint8_t* wsm = smem + C::NXS * C::XSTAGE_BYTES;       // NWS x 16 KB image
    float* stage = reinterpret_cast<float*>(wsm + C::NWS * W_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stage) + C::STAGE_FLOATS * 4);
    uint64_t* x_full = bars;                 // [NXS]
    uint64_t* x_empty = bars + C::NXS;       // [NXS]
    uint64_t* w_full = bars + 2 * C::NXS;    // [NWS]
    uint64_t* w_empty = w_full + C::NWS;     // [NWS]
    uint64_t* t_full = w_empty + C::NWS;     // [2]
    uint64_t* t_empty = t_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_entry = p.dbg ? clock64() : 0;
    if (p.dbg && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        atomicMin(reinterpret_cast<unsigned long long*>(p.dbg + 148 * 8), gt);          // first CTA entry (ns)
    }

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NXS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < C::NWS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&xmap);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 4] = clock64() - t_entry;     // prologue

    const int tiles_per_x = p.nyt * p.nzt;
    const int tiles_per_b = p.nx * tiles_per_x;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            // Issue order.  Weight taps run NWS stages ahead of the MMA warp.  The activation plane of the NEXT
            // (tile, dx) pass is requested in the middle of the current pass (hi part before tap NWS, lo part two taps
            // later): its stage was released when the previous pass finished, which is exactly when the slot of tap NWS
            // frees, so the wait never blocks the weight stream, the 2 x 33 KB transfers get a whole pass to land
            // and they no longer queue three weight images behind one 66 KB burst on the SM's L2 port.
            uint32_t wi = 0;           // running weight-stage counter
            const int npass = p.ntiles > (int)blockIdx.x ? ((p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) * 3 : 0;
            auto x_load = [&](int q, int part) {
                const int t = blockIdx.x + (q / 3) * gridDim.x, dx = q % 3;
                const int b = t / tiles_per_b;
                int rem = t % tiles_per_b;
                const int x = rem / tiles_per_x;
                rem %= tiles_per_x;
                const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
                const uint32_t s = (uint32_t)q % C::NXS, ph = ((uint32_t)q / C::NXS) & 1;
                int plane = x + dx;
                if (p.fused) {
                    // interior plane x of dX: storage planes x+1..x+3 of the zero-haloed dY; at the two boundary
                    // planes the all-zero halo pass is replaced by the pass that the folded halo plane would add
                    plane = x + 1 + dx;
                    if (dx == 0 && x == 0) plane = 2;
                    if (dx == 2 && x == p.Dint - 1) plane = p.Dint + 1;
                }
                uint8_t* dst = xs + s * C::XSTAGE_BYTES;
                if ((p.exp_skip & 2) && q >= C::NXS) {
                    if (part == 0) { mbar_wait(&x_empty[s], ph ^ 1); mbar_expect_tx(&x_full[s], 0); }
                } else if (part == 0) {
                    mbar_wait(&x_empty[s], ph ^ 1);
                    mbar_expect_tx(&x_full[s], (p.single_b ? 1 : 2) * C::ROWS * 128);
                    tma_load_5d(dst, &xmap, &x_full[s], 0, z0, y0, plane, b);
                } else if (!p.single_b) {
                    tma_load_5d(dst + C::PART_BYTES, &xmap, &x_full[s], 0, z0, y0, plane, p.B + b);
                }
            };
            if (npass) { x_load(0, 0); x_load(0, 1); }
            for (int q = 0; q < npass; ++q) {
                const int t = blockIdx.x + (q / 3) * gridDim.x, dx = q % 3;
                int wsel = dx;
                if (p.fused) {
                    const int x = (t % tiles_per_b) / tiles_per_x;
                    if (dx == 0 && x == 0) wsel = 2;
                    if (dx == 2 && x == p.Dint - 1) wsel = 0;
                }
                for (int tp = 0; tp < 9; ++tp) {
                    if (q + 1 < npass && p.xsplit) {
                        if (tp == C::NWS) x_load(q + 1, 0);
                        if (tp == C::NWS + 2) x_load(q + 1, 1);
                    }
                    const uint32_t ws = wi % C::NWS, wph = (wi / C::NWS) & 1;
                    mbar_wait(&w_empty[ws], wph ^ 1);
                    if ((p.exp_skip & 1) && wi >= (uint32_t)C::NWS) {
                        mbar_expect_tx(&w_full[ws], 0);
                    } else {
                        mbar_expect_tx(&w_full[ws], W_TAP_BYTES);
                        bulk_load(wsm + ws * W_STAGE_BYTES,
                                  reinterpret_cast<const uint8_t*>(p.w_img) + (size_t)(wsel * 9 + tp) * W_TAP_BYTES,
                                  W_TAP_BYTES, &w_full[ws]);
                    }
                    ++wi;
                }
                if (q + 1 < npass && !p.xsplit) { x_load(q + 1, 0); x_load(q + 1, 1); }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N
            const uint32_t idesc = (1u << 4) | ((uint32_t)(C::N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t xi = 0, wi = 0, ti = 0;
            long long wt = 0, wx = 0, ww = 0, c0 = 0;
            const long long tbeg = p.dbg ? clock64() : 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
                const uint32_t buf = ti & 1;
                const uint32_t dacc = tmem_base + buf * ACC_STRIDE;
                if (p.dbg) c0 = clock64();
                mbar_wait(&t_empty[buf], ((ti >> 1) & 1) ^ 1);
                if (p.dbg) wt += clock64() - c0;
                tc_fence_after();
                for (int dx = 0; dx < 3; ++dx) {
                    const uint32_t s = xi % C::NXS, ph = (xi / C::NXS) & 1;
                    if (p.dbg) c0 = clock64();
                    mbar_wait(&x_full[s], ph);
                    if (p.dbg) wx += clock64() - c0;
                    tc_fence_after();
                    const uint32_t xhi = smem_u32(xs + s * C::XSTAGE_BYTES);
                    const uint32_t xlo = xhi + C::PART_BYTES;
#pragma unroll 1
                    for (int tp = 0; tp < 9; ++tp) {
                        const uint32_t ws = wi % C::NWS, wph = (wi / C::NWS) & 1;
                        if (p.dbg) c0 = clock64();
                        mbar_wait(&w_full[ws], wph);
                        if (p.dbg) ww += clock64() - c0;
                        tc_fence_after();
                        const uint32_t wa = smem_u32(wsm + ws * W_STAGE_BYTES);
                        const uint32_t boff = ((tp / 3) * ZP + (tp % 3)) * 128;     // (dy, dz) row shift
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t acc = (dx | tp | k) != 0;
                            // [Wlo ; Whi] x Xhi, then [Whi ; (next 8 KB, lanes 64-127 masked off)] x Xlo into the same accumulator
                            tc_mma_f16(dacc, make_desc_sbo(wa + k * 32, 1024),
                                       make_desc_sbo(xhi + boff + k * 32, ZP * 128), idesc, acc);
                            if (!p.single_b)
                                tc_mma_f16_masked(dacc, make_desc_sbo(wa + W_HI_OFFSET + k * 32, 1024),
                                                  make_desc_sbo(xlo + boff + k * 32, ZP * 128), idesc, 1u, 0u, 0u, ~0u, ~0u);
                        }
                        tc_commit(&w_empty[ws]);
                        ++wi;
                    }
                    tc_commit(&x_empty[s]);
                    ++xi;
                }
                tc_commit(&t_full[buf]);
            }
            if (p.dbg) {
                p.dbg[blockIdx.x * 8 + 0] = wt; p.dbg[blockIdx.x * 8 + 1] = wx;
                p.dbg[blockIdx.x * 8 + 2] = ww; p.dbg[blockIdx.x * 8 + 3] = clock64() - tbeg;
                p.dbg[blockIdx.x * 8 + 5] = clock64();          // MMA issue loop end (absolute)
            }
        }
    } else {
        // ================= epilogue (warps 2..9, 256 threads) =================
        // warp w may only read TMEM lanes 32*(w%4)..+31: quadrants 0,1 hold the L rows (co = 32e + lane),
        // quadrants 2,3 the H rows; the two warps of a quadrant split a chunk's columns
        const int e = warp & 3;
        const int half = (warp - 2) >> 2;
        const int et = threadIdx.x - 64;             // 0..255
        const bool is_h = e >= 2;
        const int co = 32 * (e & 1) + lane;
        const float bias = (is_h && p.bias) ? p.bias[co] : 0.f;
        const float s1 = is_h ? 1.f : SR4D_LO_INV;
        const int Do = p.Do;
        const int g8 = et & 7;                       // 8-channel group handled in the store phase
        const int vq = et >> 3;                      // voxel of the chunk handled in the store phase
        float* my_stage = stage + (is_h ? 0 : C::CV * 64) + co;
        float amax = 0.f;
        float ksplit = 0.f;
        if (p.fused && p.split_hi) {
            float bound = 8.f * *p.gain * __uint_as_float(*p.dy_amax);
            if (p.add_amax) bound += __uint_as_float(*p.add_amax);
            int e = 0;
            if (bound > 0.f && bound < 3.0e38f) e = 14 - ilogbf(bound);     // bound * 2^e < 2^15
            e = max(-120, min(120, e));
            ksplit = exp2f((float)e);
            if (blockIdx.x == 0 && et == 0) *p.split_exp = e;
        }
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
            const int b = t / tiles_per_b;
            int rem = t % tiles_per_b;
            const int x = rem / tiles_per_x;
            rem %= tiles_per_x;
            const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
            const bool x_edge = p.halo && (x == 0 || x == Do - 1);
            const uint32_t buf = ti & 1;
            mbar_wait(&t_full[buf], (ti >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + buf * ACC_STRIDE + ((uint32_t)(32 * e) << 16);
#pragma unroll 1
            for (int ch = 0; ch < C::NCHUNK; ++ch) {
                const int line = ch * C::LPC + (vq >> 3);
                const int y = y0 + line, z = z0 + (vq & 7);
                bool active = vq < C::CV && line < TY && y < Do && z < Do;
                // fused dgrad: (y, z) are padded-grid coordinates; only interior voxels produce output
                const int Di = p.Dint;
                if (p.fused) active = active && y >= 1 && y <= Di && z >= 1 && z <= Di;
                float4 ap0, ap1, aq0, aq1;
                // residual prefetch for the (voxel, channel-group) item this thread stores
                uint4 rh, rl;
                if (p.fused) {
                    if (active) {
                        const size_t go = g4_off(Di, b, x, y - 1, z - 1) + g8 * 8;
                        if (p.add_pre) {
                            ap0 = *reinterpret_cast<const float4*>(p.add_pre + go);
                            ap1 = *reinterpret_cast<const float4*>(p.add_pre + go + 4);
                        }
                        if (p.add_post) {
                            aq0 = *reinterpret_cast<const float4*>(p.add_post + go);
                            aq1 = *reinterpret_cast<const float4*>(p.add_post + go + 4);
                        }
                        if (p.sav_hi) {
                            const size_t o = act_off(Di, b, x, y - 1, z - 1) + g8 * 8;
                            rh = *reinterpret_cast<const uint4*>(p.sav_hi + o);
                            rl = *reinterpret_cast<const uint4*>(p.sav_lo + o);
                        }
                    }
                } else if (p.res_hi && active) {
                    const size_t o = act_off(Do, b, x, y, z) + g8 * 8;
                    rh = *reinterpret_cast<const uint4*>(p.res_hi + o);
                    rl = *reinterpret_cast<const uint4*>(p.res_lo + o);
                }
                // ---- phase A: TMEM -> registers -> fp32 staging [part][voxel][channel] ----
                const int c0 = ch * C::CV;
                const int ncols = (C::N - c0) < C::CV ? (C::N - c0) : C::CV;
                if (half == 0) {
                    float a[16];
                    tc_ld16(trow + c0, a);
                    tc_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) my_stage[j * 64] = fmaf(a[j], s1, bias);
                } else if (ncols > 16) {
                    if (C::CV == 32) {
                        float a[16];
                        tc_ld16(trow + c0 + 16, a);
                        tc_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) my_stage[(16 + j) * 64] = fmaf(a[j], s1, bias);
                    } else {
                        float a[8];
                        tc_ld8(trow + c0 + 16, a);
                        tc_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) my_stage[(16 + j) * 64] = fmaf(a[j], s1, bias);
                    }
                }
                if (ch == C::NCHUNK - 1) {
                    // all TMEM reads of this tile are done: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[buf]);
                }
                named_bar(1, NUM_EPI);
                // ---- phase B: coalesced residual / activation / split / store ----
                if (active) {
                    const float* sp = stage + vq * 64 + g8 * 8;
                    const float4 h0 = *reinterpret_cast<const float4*>(sp);
                    const float4 h1 = *reinterpret_cast<const float4*>(sp + 4);
                    const float4 l0 = *reinterpret_cast<const float4*>(sp + C::CV * 64);
                    const float4 l1 = *reinterpret_cast<const float4*>(sp + C::CV * 64 + 4);
                    float val[8] = {h0.x + l0.x, h0.y + l0.y, h0.z + l0.z, h0.w + l0.w,
                                    h1.x + l1.x, h1.y + l1.y, h1.z + l1.z, h1.w + l1.w};
                    if (p.fused) {
                        // MirrorPadGrad inside the chunk: halo lines / columns fold onto their clamped neighbour
                        auto fold = [&](int dv) {
                            const float* q = sp + dv * 64;
                            const float4 a0 = *reinterpret_cast<const float4*>(q);
                            const float4 a1 = *reinterpret_cast<const float4*>(q + 4);
                            const float4 b0 = *reinterpret_cast<const float4*>(q + C::CV * 64);
                            const float4 b1 = *reinterpret_cast<const float4*>(q + C::CV * 64 + 4);
                            val[0] += a0.x + b0.x; val[1] += a0.y + b0.y; val[2] += a0.z + b0.z; val[3] += a0.w + b0.w;
                            val[4] += a1.x + b1.x; val[5] += a1.y + b1.y; val[6] += a1.z + b1.z; val[7] += a1.w + b1.w;
                        };
                        const bool ylo = y == 1, yhi = y == Di, zlo = z == 1, zhi = z == Di;
                        if (ylo) fold(-8);
                        if (yhi) fold(8);
                        if (zlo) { fold(-1); if (ylo) fold(-9); if (yhi) fold(7); }
                        if (zhi) { fold(1); if (ylo) fold(-7); if (yhi) fold(9); }
                        const float ks = p.dy_exp ? exp2f((float)(-*p.dy_exp)) : 1.f;
#pragma unroll
                        for (int k = 0; k < 8; ++k) val[k] *= ks;
                        if (p.add_pre) {
                            val[0] += ap0.x; val[1] += ap0.y; val[2] += ap0.z; val[3] += ap0.w;
                            val[4] += ap1.x; val[5] += ap1.y; val[6] += ap1.z; val[7] += ap1.w;
                        }
                        if (p.sav_hi) {
                            const __half2* hh = reinterpret_cast<const __half2*>(&rh);
                            const __half2* ll = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 ha = __half22float2(hh[k]), la = __half22float2(ll[k]);
                                val[2 * k] *= act_grad_from_out(fmaf(la.x, SR4D_LO_INV, ha.x), p.slope);
                                val[2 * k + 1] *= act_grad_from_out(fmaf(la.y, SR4D_LO_INV, ha.y), p.slope);
                            }
                        }
                        if (p.add_post) {
                            val[0] += aq0.x; val[1] += aq0.y; val[2] += aq0.z; val[3] += aq0.w;
                            val[4] += aq1.x; val[5] += aq1.y; val[6] += aq1.z; val[7] += aq1.w;
                        }
                        const size_t go = g4_off(Di, b, x, y - 1, z - 1) + g8 * 8;
                        float* o = p.out_g4 + go;
                        *reinterpret_cast<float4*>(o) = make_float4(val[0], val[1], val[2], val[3]);
                        *reinterpret_cast<float4*>(o + 4) = make_float4(val[4], val[5], val[6], val[7]);
                        if (p.split_hi) {
                            __align__(16) __half hv[8];
                            __align__(16) __half lv[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) split_f16(val[k] * ksplit, hv[k], lv[k]);
                            *reinterpret_cast<uint4*>(p.split_hi + go) = *reinterpret_cast<const uint4*>(hv);
                            *reinterpret_cast<uint4*>(p.split_lo + go) = *reinterpret_cast<const uint4*>(lv);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(val[k]));
                    } else if (p.out_raw) {
                        float* o = p.out_raw + ((((size_t)b * Do + x) * Do + y) * Do + z) * 64 + g8 * 8;
                        *reinterpret_cast<float4*>(o) = make_float4(val[0], val[1], val[2], val[3]);
                        *reinterpret_cast<float4*>(o + 4) = make_float4(val[4], val[5], val[6], val[7]);
#pragma unroll
                        for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(val[k]));
                    } else {
                        if (p.res_hi) {
                            const __half2* hh = reinterpret_cast<const __half2*>(&rh);
                            const __half2* ll = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                float2 ha = __half22float2(hh[k]), la = __half22float2(ll[k]);
                                val[2 * k] += fmaf(la.x, SR4D_LO_INV, ha.x);
                                val[2 * k + 1] += fmaf(la.y, SR4D_LO_INV, ha.y);
                            }
                        }
                        __align__(16) __half hv[8];
                        __align__(16) __half lv[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) split_f16(act_fn(val[k], p.slope), hv[k], lv[k]);
                        const uint4 H = *reinterpret_cast<const uint4*>(hv), L = *reinterpret_cast<const uint4*>(lv);
                        const size_t o = act_off(Do, b, x, y, z) + g8 * 8;
                        *reinterpret_cast<uint4*>(p.out_hi + o) = H;
                        *reinterpret_cast<uint4*>(p.out_lo + o) = L;
                        const bool edge = p.halo && (x_edge || y == 0 || y == Do - 1 || z == 0 || z == Do - 1);
                        if (edge) {
                            // replicate into the halo positions this voxel is the clamp image of
                            for (int ddx = -1; ddx <= 1; ++ddx) {
                                if ((ddx == -1 && x != 0) || (ddx == 1 && x != Do - 1)) continue;
                                for (int ddy = -1; ddy <= 1; ++ddy) {
                                    if ((ddy == -1 && y != 0) || (ddy == 1 && y != Do - 1)) continue;
                                    for (int ddz = -1; ddz <= 1; ++ddz) {
                                        if ((ddz == -1 && z != 0) || (ddz == 1 && z != Do - 1)) continue;
                                        if ((ddx | ddy | ddz) == 0) continue;
                                        const size_t oo = act_off(Do, b, x + ddx, y + ddy, z + ddz) + g8 * 8;
                                        *reinterpret_cast<uint4*>(p.out_hi + oo) = H;
                                        *reinterpret_cast<uint4*>(p.out_lo + oo) = L;
                                    }
                                }
                            }
                        }
                    }
                }
                named_bar(1, NUM_EPI);
            }
        }
        if (p.absmax) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            if (lane == 0) atomicMax(p.absmax, __float_as_uint(amax));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.dbg && threadIdx.x == 64) {
        const long long now = clock64();
        p.dbg[blockIdx.x * 8 + 6] = now - t_entry;                                  // whole CTA lifetime
        p.dbg[blockIdx.x * 8 + 7] = now - p.dbg[blockIdx.x * 8 + 5];                // tail after the last MMA was issued
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        atomicMax(reinterpret_cast<unsigned long long*>(p.dbg + 148 * 8 + 1), gt);      // last CTA exit (ns)
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// weight image: fp32 Keras [27][ci][co] -> fp16 split, rows [Wlo(co 0..63) ; Whi(co 0..63)], swizzled
// ------------------------------------------------------------------------------------------
struct PrepList {
    int n;
    int layer[64];           // image slot of entry i
    long long off[64];       // float offset of its Keras kernel in the flat parameter buffer
};
// grid (27*128*64/256, n entries, 2 images): one thread per (tap, row, k) element
__global__ void prep_weights_kernel(const float* __restrict__ params, PrepList l, __half* __restrict__ images) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 27 * 128 * 64) return;
    const float* w = params + l.off[blockIdx.y];
    const int dgrad = blockIdx.z;
    __half* img = images + ((size_t)l.layer[blockIdx.y] * 2 + dgrad) * (27 * 128 * 64);
    const int k = i & 63, row = (i >> 6) & 127, tap = i >> 13;
    const int n = row & 63;
    const bool is_lo = row < 64;
    // forward: A[n=co][k=ci] = W[tap][ci][co];  dgrad: A[n=ci][k=co] = W[26-tap][ci][co]
    const float v = dgrad ? w[((size_t)(26 - tap) * 64 + n) * 64 + k] : w[((size_t)tap * 64 + k) * 64 + n];
    __half h, lo;
    split_f16(v, h, lo);
    const int grp = row >> 3, rr = row & 7;
    const size_t off = (size_t)tap * (128 * 64) + grp * 512 + rr * 64 + (((k >> 3) ^ rr) << 3) + (k & 7);
    img[off] = is_lo ? lo : h;
}

// dgrad gain of each listed layer: max over ci of sum_{tap,co} |W[tap][ci][co]| (bounds |dX| <= gain * max|dY|)
__global__ void __launch_bounds__(256) weight_gain_kernel(const float* __restrict__ params, PrepList l, float* __restrict__ gain) {
    const float* w = params + l.off[blockIdx.x];
    const int co = threadIdx.x & 63, ph = threadIdx.x >> 6;     // coalesced over co; 4 phases over (tap, ci)
    __shared__ float colsum[64][65];
    __shared__ float red[64];
    // thread (ph, co) accumulates |W[t][ci][co]| into per-ci sums: walk (t, ci) pairs phase-strided
    for (int i = threadIdx.x; i < 64 * 65; i += 256) (&colsum[0][0])[i] = 0.f;
    __syncthreads();
    for (int ci = ph; ci < 64; ci += 4) {
        float s = 0.f;
        for (int t = 0; t < 27; ++t) s += fabsf(w[((size_t)t * 64 + ci) * 64 + co]);
        colsum[ci][co] = s;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f;
        for (int c2 = 0; c2 < 64; ++c2) s += colsum[threadIdx.x][c2];
        red[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int i = 0; i < 64; ++i) m = fmaxf(m, red[i]);
        gain[l.layer[blockIdx.x]] = m;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int TY>
cudaError_t launch_cfg(const CUtensorMap& map, KParams p, cudaStream_t s) {
    using C = Cfg<TY>;
    cudaError_t ea = tc_func_smem(reinterpret_cast<const void*>(conv64_tc_kernel<TY>), C::SMEM_BYTES);
    if (ea != cudaSuccess) return ea;
    p.nyt = (p.Do + TY - 1) / TY;
    p.nzt = (p.Do + TZ - 1) / TZ;
    p.ntiles = p.B * p.nx * p.nyt * p.nzt;
    const int sms = tc_num_sms();
    int grid = p.ntiles < sms ? p.ntiles : sms;
    conv64_tc_kernel<TY><<<grid, NUM_THREADS, C::SMEM_BYTES, s>>>(map, p);
    return cudaGetLastError();
}

}  // namespace

struct TcWeights {
    int nlayers = 0;
    __half* img = nullptr;   // [nlayers][2 (fwd, dgrad)][27*128*64]
    float* gain = nullptr;   // [nlayers] dgrad gain bound (weight_gain_kernel)
};

cudaError_t tc_alloc_weights(TcWeights** w, int nlayers) {
    TcWeights* t = new TcWeights();
    t->nlayers = nlayers;
    cudaError_t e = cudaMalloc((void**)&t->img, (size_t)nlayers * 2 * 27 * 128 * 64 * sizeof(__half));
    if (e == cudaSuccess) e = cudaMalloc((void**)&t->gain, (size_t)nlayers * sizeof(float));
    if (e != cudaSuccess) { cudaFree(t->img); delete t; return e; }
    *w = t;
    return cudaSuccess;
}
void tc_free_weights(TcWeights* w) {
    if (!w) return;
    cudaFree(w->img);
    cudaFree(w->gain);
    delete w;
}
cudaError_t tc_prepare_weights(TcWeights* w, const float* params, const int* layers, const long long* offsets, int n,
                               cudaStream_t s) {
    for (int i0 = 0; i0 < n; i0 += 64) {
        PrepList l;
        l.n = n - i0 < 64 ? n - i0 : 64;
        for (int i = 0; i < l.n; ++i) { l.layer[i] = layers[i0 + i]; l.off[i] = offsets[i0 + i]; }
        dim3 grid((27 * 128 * 64 + 255) / 256, l.n, 2);
        prep_weights_kernel<<<grid, 256, 0, s>>>(params, l, w->img);
        weight_gain_kernel<<<l.n, 256, 0, s>>>(params, l, w->gain);
    }
    return cudaGetLastError();
}

namespace {
int pick_ty(int Do) {
    // y-tile height: the candidate that wastes the fewest MMA columns on this grid (26 serves the
    // padded dgrad grids 26^3 / 50^3, 24 the forward grids 24^3 / 48^3)
    int ty = 8;
    if (Do > 8) {
        const int cand[3] = {16, 24, 26};
        long best = -1;
        for (int c : cand) {
            const long cost = (long)((Do + c - 1) / c) * c;
            if (best < 0 || cost < best || (cost == best && c > ty)) { best = cost; ty = c; }
        }
    }
    return ty;
}
}  // namespace

bool tc_dgrad_fusable(int D) {
    if (D < 2) return false;
    const int ty = pick_ty(D + 2);
    const int lpc = ty % 4 == 0 ? 4 : 3;
    auto chunk = [&](int line) { return (line / ty) * 1000 + (line % ty) / lpc; };
    return chunk(0) == chunk(1) && chunk(D) == chunk(D + 1) && D / TZ == (D + 1) / TZ;
}

cudaError_t tc_conv64(TcWeights* w, const TcConvArgs& a, cudaStream_t s) {
    const int Do = a.in.D, B = a.in.B, Dp = a.in.D + 2;
    if (!a.out_raw && !a.fused && a.out.D != Do) return cudaErrorInvalidValue;
    if (a.fused && (!a.dgrad || !a.out_g4 || !tc_dgrad_fusable(Do - 2))) return cudaErrorInvalidValue;
    KParams p;
    p.w_img = w->img + ((size_t)a.layer * 2 + (a.dgrad ? 1 : 0)) * 27 * 128 * 64;
    p.out_hi = a.out.hi; p.out_lo = a.out.lo;
    p.res_hi = a.res_hi; p.res_lo = a.res_lo;
    p.bias = a.bias; p.out_raw = a.out_raw; p.absmax = a.absmax;
    p.slope = a.slope; p.B = B; p.Do = Do; p.halo = a.halo;
    p.nx = a.fused ? Do - 2 : Do;
    p.fused = a.fused; p.Dint = Do - 2;
    p.single_b = (a.dgrad && a.single_b) ? 1 : 0;
    p.dy_exp = a.dy_exp; p.add_pre = a.add_pre; p.add_post = a.add_post;
    p.sav_hi = a.sav_hi; p.sav_lo = a.sav_lo; p.out_g4 = a.out_g4;
    p.split_hi = a.split_out;
    p.split_lo = a.split_out ? a.split_out + act_plane_elems(B, Do) : nullptr;   // [2B][D+4]^3[64]: hi planes then lo planes
    p.split_exp = a.split_exp; p.gain = w->gain + a.layer; p.dy_amax = a.dy_amax; p.add_amax = a.add_amax;
    if (a.split_out && (!a.split_exp || !a.dy_amax || !a.fused)) return cudaErrorInvalidValue;
    static const bool xsplit = !(getenv("SR4D_TC_XSPLIT") && atoi(getenv("SR4D_TC_XSPLIT")) == 0);
    p.xsplit = xsplit;
    static const int exp_skip = getenv("SR4D_TC_EXP_SKIP") ? atoi(getenv("SR4D_TC_EXP_SKIP")) : 0;
    p.exp_skip = exp_skip;
    p.dbg = nullptr;
    static const bool debug = getenv("SR4D_TC_DEBUG") != nullptr;
    static long long* dbg_buf = nullptr;
    if (debug) {
        if (!dbg_buf) cudaMalloc((void**)&dbg_buf, (148 * 8 + 2) * sizeof(long long));
        cudaMemsetAsync(dbg_buf, 0, (148 * 8 + 2) * sizeof(long long), s);
        cudaMemsetAsync(dbg_buf + 148 * 8, 0x7f, sizeof(long long), s);      // min slot starts high
        p.dbg = dbg_buf;
    }
    const int ty = pick_ty(Do);
    CUtensorMap map;
    if (!tc_make_act_map(&map, a.in.hi, B, Dp, ty + 2, ZP)) return cudaErrorUnknown;
    if (a.in.lo != a.in.hi + act_plane_elems(B, a.in.D)) return cudaErrorInvalidValue;   // planes must be packed
    cudaError_t e;
    switch (ty) {
        case 8: e = launch_cfg<8>(map, p, s); break;
        case 16: e = launch_cfg<16>(map, p, s); break;
        case 24: e = launch_cfg<24>(map, p, s); break;
        default: e = launch_cfg<26>(map, p, s); break;
    }
    if (debug && e == cudaSuccess) {
        long long hb[148 * 8 + 2];
        cudaStreamSynchronize(s);
        cudaMemcpy(hb, dbg_buf, sizeof hb, cudaMemcpyDeviceToHost);
        double a4[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 148; ++i) for (int k = 0; k < 8; ++k) a4[k] += (double)hb[i * 8 + k] / 148;
        fprintf(stderr, "[tc dbg] Do=%d B=%d ty=%d dgrad=%d: MMA-warp wait cycles avg/CTA: t_empty %.0f  x_full %.0f  w_full %.0f  of total %.0f | prologue %.0f  tail after last MMA issue %.0f  CTA lifetime %.0f | first entry -> last exit %.1f us\n",
                Do, B, ty, a.dgrad, a4[0], a4[1], a4[2], a4[3], a4[4], a4[7], a4[6], (double)(hb[148 * 8 + 1] - hb[148 * 8]) * 1e-3);
    }
    return e;
}
