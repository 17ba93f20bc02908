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
#include <cstring>
#include <vector>

#include "conv_tc.h"
#include "tc_host.h"
#include "tc_ptx.cuh"

namespace {

constexpr int W_TAP_BYTES = 128 * 64 * 2;            // 16 KB: [Wlo half | Whi half], each MN-major: 64 K rows (ci) x 64 co, fp16, swizzled
constexpr int W_HI_OFFSET = 64 * 64 * 2;             // the Whi half starts 8 KB into the image
constexpr int W_ZERO_BYTES = 2048;                   // 16 K rows of zeros: the masked rows of the second instruction read them
constexpr int W_STAGE_BYTES = W_TAP_BYTES;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI = NUM_EPI_WARPS * 32;   // epilogue threads
constexpr int NUM_THREADS = 64 + NUM_EPI;
constexpr int TZ = 8;                       // voxels per 8-row group (one z run)
constexpr int ZP = TZ + 2;                  // plane row pitch in voxels
constexpr int TMEM_COLS = 512;              // two accumulators of up to 256 columns
constexpr int ACC_STRIDE = 256;             // column offset of the second accumulator

template <int TY, bool SINGLE>
struct Cfg {
    static constexpr int N = TY * TZ;                                  // voxels per tile = MMA N
    static constexpr int ROWS = (TY + 2) * ZP;                         // rows of one staged plane part
    static constexpr int PART_BYTES = (ROWS * 128 + 1023) / 1024 * 1024;
    static constexpr int XSTAGE_BYTES = (SINGLE ? 1 : 2) * PART_BYTES; // hi part | lo part (SINGLE: the B operand has no lo plane)
    static constexpr int NXS = 2;
    // Weight taps stream through NWS 16 KB stages that are handed back to the producer in GROUPS of WG taps: every
    // tcgen05.commit costs the tensor pipe ~100 cycles (tools/probe/wgrad_probe.cu on B200: 8 MMAs + 1 commit take 870
    // cycles instead of 770, + 2 commits 1008), so one commit per tap was a 13 % (two-plane: 8 instructions per tap) to
    // 24 % (single-plane dgrad: 4 per tap) tax.  Two groups are in flight (NWS = 2 WG); the activation stages need no
    // commit of their own: by the time the producer may refill the weight stage of a tap it has seen the group commit
    // that covers the whole previous-but-one pass (checked exhaustively when the schedule was written).
    static constexpr int WG = SINGLE ? 3 : 2;
    static constexpr int NGW = 2;
    static constexpr int NWS = NGW * WG;
    // y-lines per epilogue chunk (measured: sending a whole TY = 12 tile through the staging buffer as ONE chunk of three
    // items per thread is slower -- 357 vs 333 us for the 17 chained 24^3 layers at batch 1, profiles/r02_chain_cluster_ab.txt)
    static constexpr int LPC = 4;
    static constexpr int CV = LPC * TZ;                                // voxels (TMEM columns) per chunk
    static constexpr int NCHUNK = (TY + LPC - 1) / LPC;
    // Epilogue organisation.  Two-plane kernels (forward, two-plane dgrad) spend >= 22k cycles of MMAs per tile and
    // run ONE group of 8 warps chunk by chunk.  The single-plane dgrad has half the MMA time per tile (11k cycles),
    // which the latency-bound chunk loop (global loads of the skip gradient / saved activation -> fold -> stores) of one
    // group does not fit into: there TWO groups of 4 warps work on alternate chunks, each with its own staging buffer
    // and named barrier, and every thread carries two (voxel, 8-channel) items -- four times the loads in flight.
    static constexpr int NG = SINGLE ? 2 : 1;
    static constexpr int GT = NUM_EPI / NG;                            // threads per group
    static constexpr int IT = CV * 8 / GT;                             // items per thread per chunk (1 or 2)
    static constexpr int CP = NUM_EPI_WARPS / NG / 4;                  // warps sharing one TMEM lane quadrant (2 or 1)
    static constexpr int COLS = CV / CP;                               // TMEM columns a warp reads per chunk (16 or 32)
    static constexpr int GSTAGE_FLOATS = 2 * CV * 64;                  // per group: [H | L][voxel][channel]
    static constexpr int STAGE_FLOATS = NG * GSTAGE_FLOATS;
    static constexpr int SMEM_BYTES = 1024 + NXS * XSTAGE_BYTES + NWS * W_STAGE_BYTES + W_ZERO_BYTES + STAGE_FLOATS * 4 + 256;
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N constraint for M=128");
    static_assert(IT * GT == CV * 8 && COLS * CP == CV, "epilogue work split");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// what an epilogue thread brings along for its (voxel, 8-channel group) items of one chunk: coordinates and the global-memory
// operands of the chunk's second phase (skip gradients / saved activation of the fused dgrad, residual of the forward)
template <int IT>
struct EpiItems {
    bool active[IT];
    int vy[IT], vz[IT];
    float4 ap0[IT], ap1[IT], aq0[IT], aq1[IT];
    uint4 rh[IT], rl[IT];
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
    unsigned int* ovf;       // optional: set to 1 when an output activation had to be clamped to the fp16 range
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
    long long* dbg;          // SR4D_TC_DEBUG=1: per-CTA cycles the MMA warp waited on {t_empty, x_full, w_full} and its total
};

// A chain of forward layers on one grid in ONE launch (n > 0): layer L reads params[L] / maps[L].  At batch 1 a 24^3 layer
// is ~7 us of MMAs behind ~12 us of launch, prologue and tail; the chain keeps barriers, TMEM and the pipeline state alive
// across the 17 low-resolution (11 high-resolution) layers of a forward.  Layers are ordered by PLANE-LEVEL dependencies
// instead of a grid-wide barrier: done[(L * B + b) * nx + x] counts the finished tiles of x-plane x of layer L, and the
// activation load of a pass of layer L+1 waits until the plane it reads is complete in layer L -- so the first planes of
// a layer start while the last planes of the previous one are still in flight (tiles run in plane order on every CTA, and
// every dependency points to a lower layer: no cycles).  exit_count: CTAs that left the kernel; the last one re-arms `done`.
struct ChainArgs {
    const KParams* params;
    const CUtensorMap* maps;
    int n;
    unsigned int* done;          // [(n - 1) * B * nx] tiles done per plane, [n - 1] tiles done per layer, then the exit counter
    int nplane;                  // (n - 1) * B * nx
    int ndone;                   // nplane + n - 1
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

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

// MN-major SWIZZLE_128B operand (a shared-memory row = one K index, 64 M elements = 128 B): lbo = distance between the two
// 64-row atoms of an M = 128 operand, sbo = distance between 8-row K groups
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// CL: the kernel runs as CTA pairs (cluster of 2, consecutive tiles of one x-plane): every weight tap is fetched from L2 by
// ONE CTA of the pair (taps alternate) and multicast into both CTAs' stages, a stage is handed back when both MMA warps
// have consumed it (multicast commit, barrier count 2) -- half the L2 -> SM weight traffic, which is 4/5 of the kernel's.
template <int TY, bool SINGLE, bool CHAIN, bool CL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv64_tc_kernel(const __grid_constant__ CUtensorMap xmap, const KParams p0, const ChainArgs chain) {
    using C = Cfg<TY, SINGLE>;
    const uint32_t crank = CL ? cluster_rank() : 0u;
    const int nl = CHAIN ? chain.n : 1;
    const KParams& p = p0;                   // geometry, debug buffer: identical for every layer of a chain
    extern __shared__ uint8_t smem_raw[];
    // align by OFFSET (not by pointer cast) so the compiler keeps the shared address space (STS/LDS, not generic)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* xs = smem;                                   // NXS x [hi part | lo part]
    uint8_t* wsm = smem + C::NXS * C::XSTAGE_BYTES;       // NWS x 16 KB image
    uint8_t* zsm = wsm + C::NWS * W_STAGE_BYTES;          // zeros (behind every weight stage: the LBO of a descriptor is positive)
    float* stage = reinterpret_cast<float*>(zsm + W_ZERO_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stage) + C::STAGE_FLOATS * 4);
    uint64_t* x_full = bars;                 // [NXS]
    uint64_t* w_full = bars + C::NXS;        // [NWS]
    uint64_t* w_empty = w_full + C::NWS;     // [NGW] one per group of WG stages
    uint64_t* t_full = w_empty + C::NGW;     // [2]
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
        for (int i = 0; i < C::NXS; ++i) mbar_init(&x_full[i], 1);
        for (int i = 0; i < C::NWS; ++i) mbar_init(&w_full[i], 1);
        for (int i = 0; i < C::NGW; ++i) mbar_init(&w_empty[i], CL ? 2 : 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&xmap);
    }
    for (int i = threadIdx.x; i < W_ZERO_BYTES / 16; i += NUM_THREADS) reinterpret_cast<uint4*>(zsm)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL) cluster_sync_all();              // the peer's barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 4] = clock64() - t_entry;     // prologue

    const int tiles_per_x = p.nyt * p.nzt;
    const int tiles_per_b = p.nx * tiles_per_x;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            // Issue order.  Weight taps run up to NWS stages ahead of the MMA warp, released WG at a time.  The activation
            // plane of the NEXT (tile, dx) pass is requested in the middle of the current pass (hi part at tap NWS, lo part
            // two taps later): its stage was last read by the previous pass, whose completion the producer has seen through
            // the weight-group barriers by then, the transfers get half a pass to land and they do not queue the
            // weight images behind one 66 KB burst on the SM's L2 port.
            uint32_t wi = 0;           // running weight-tap counter
            uint32_t xq = 0;           // running activation-pass counter (stage = xq % NXS)
            const int npass = p.ntiles > (int)blockIdx.x ? ((p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) * 3 : 0;
            for (int L = 0; L < nl; ++L) {
            KParams pl;                              // chained layers: a private copy of the layer's parameter block
            if (CHAIN) pl = chain.params[L];
            const KParams& p = CHAIN ? pl : p0;
            const CUtensorMap* xm = CHAIN ? chain.maps + L : &xmap;
            // Activation loads of layers after the first wait for the plane they read to be complete in the previous layer
            // (release on the writers' side: __threadfence + CTA barrier + atomic per tile).
            auto plane_of = [&](int q, int& b) {
                const int t = blockIdx.x + (q / 3) * gridDim.x;
                b = t / tiles_per_b;
                const int x = (t % tiles_per_b) / tiles_per_x;
                return min(max(x + q % 3 - 1, 0), p.nx - 1);
            };
            // (an acquire load is an L2 round trip in the middle of the weight stream: once the layer counter says that the
            // whole previous layer is complete -- after the first round on large grids -- no more flags are read)
            bool prev_complete = false;
            auto ready = [&](int q) {
                if (prev_complete) return true;
                int b;
                const int xp = plane_of(q, b);
                if (ld_acquire_gpu(chain.done + ((size_t)(L - 1) * p.B + b) * p.nx + xp) < (unsigned int)tiles_per_x) return false;
                prev_complete = ld_acquire_gpu(chain.done + chain.nplane + (L - 1)) >= (unsigned int)p.ntiles;
                return true;
            };
            const uint32_t xq0 = xq;
            auto x_load = [&](int q, int part) {
                const int t = blockIdx.x + (q / 3) * gridDim.x, dx = q % 3;
                const int b = t / tiles_per_b;
                int rem = t % tiles_per_b;
                const int x = rem / tiles_per_x;
                rem %= tiles_per_x;
                const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
                const uint32_t s = (xq0 + (uint32_t)q) % C::NXS;
                int plane = x + dx;
                if (p.fused) {
                    // interior plane x of dX: storage planes x+1..x+3 of the zero-haloed dY; at the two boundary
                    // planes the all-zero halo pass is replaced by the pass that the folded halo plane would add
                    plane = x + 1 + dx;
                    if (dx == 0 && x == 0) plane = 2;
                    if (dx == 2 && x == p.Dint - 1) plane = p.Dint + 1;
                }
                uint8_t* dst = xs + s * C::XSTAGE_BYTES;
                if (part == 0) {
                    mbar_expect_tx(&x_full[s], (SINGLE ? 1 : 2) * C::ROWS * 128);
                    tma_load_5d(dst, xm, &x_full[s], 0, z0, y0, plane, b);
                } else if (!SINGLE) {
                    tma_load_5d(dst + C::PART_BYTES, xm, &x_full[s], 0, z0, y0, plane, p.B + b);
                }
            };
            const bool dep = CHAIN && L > 0;
            // `pend`: the activation load of the current pass has not been issued yet because its plane was not complete at the
            // prefetch point.  It is issued (after a blocking wait) before tap NWS - WG + 1 of the pass: a group-start wait at tap
            // counter wi needs taps <= wi - NWS + WG - 1 consumed, which from that tap on includes taps of this pass.
            bool pend = dep;
            if (npass && !dep) { x_load(0, 0); x_load(0, 1); }
            for (int q = 0; q < npass; ++q) {
                const int t = blockIdx.x + (q / 3) * gridDim.x, dx = q % 3;
                int wsel = dx;
                if (p.fused) {
                    const int x = (t % tiles_per_b) / tiles_per_x;
                    if (dx == 0 && x == 0) wsel = 2;
                    if (dx == 2 && x == p.Dint - 1) wsel = 0;
                }
                for (int tp = 0; tp < 9; ++tp) {
                    const uint32_t ws = wi % C::NWS;
                    if (pend && tp == C::NWS - C::WG + 1) {
                        while (!ready(q)) { }
                        asm volatile("fence.proxy.async;" ::: "memory");
                        x_load(q, 0); x_load(q, 1);
                        pend = false;
                    }
                    if (wi % C::WG == 0)                 // first stage of a group: wait until the group's previous round was consumed
                        mbar_wait(&w_empty[(wi / C::WG) % C::NGW], ((wi / C::NWS) & 1) ^ 1);
                    if (q + 1 < npass && p.xsplit) {
                        if (tp == 4) {
                            if (dep && !ready(q + 1)) pend = true;           // not yet: issued inside the next pass
                            else { if (dep) asm volatile("fence.proxy.async;" ::: "memory"); x_load(q + 1, 0); }
                        }
                        if (tp == 6 && !pend) x_load(q + 1, 1);
                    }
                    mbar_expect_tx(&w_full[ws], W_TAP_BYTES);
                    if (!CL)
                        bulk_load(wsm + ws * W_STAGE_BYTES,
                                  reinterpret_cast<const uint8_t*>(p.w_img) + (size_t)(wsel * 9 + tp) * W_TAP_BYTES,
                                  W_TAP_BYTES, &w_full[ws]);
                    else if ((wi & 1u) == crank)       // this CTA's turn: one fetch from L2 lands in both CTAs' stages
                        bulk_load_mc(wsm + ws * W_STAGE_BYTES,
                                     reinterpret_cast<const uint8_t*>(p.w_img) + (size_t)(wsel * 9 + tp) * W_TAP_BYTES,
                                     W_TAP_BYTES, &w_full[ws], (uint16_t)3);
                    ++wi;
                }
                if (q + 1 < npass && !p.xsplit) {
                    if (dep && !ready(q + 1)) pend = true;
                    else { if (dep) asm volatile("fence.proxy.async;" ::: "memory"); x_load(q + 1, 0); x_load(q + 1, 1); }
                }
            }
            xq += (uint32_t)npass;
            }   // layers
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp runs the loop converged and one ELECTED lane issues (elect.sync): with `if (lane == 0)` around the loop
        // ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY loop with R2UR moves (~15 instructions per MMA).
        {
            const bool leader = elect_one();
            // instruction descriptor: D=f32, A=B=f16, A (weights) MN-major, B (voxels) K-major, M=128, N
            const uint32_t idesc = (1u << 4) | (1u << 15) | ((uint32_t)(C::N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            // Everything the loop needs per tap is a running counter or a compile-time constant: the nine taps of a pass are
            // unrolled (row shift and accumulate flag fold into immediates), the weight stage / phase / group indices advance
            // by increments instead of divisions, and a descriptor is its constant part OR-ed with (address >> 4) -- the
            // bases are masked to the 14-bit field once, the offsets added to them stay inside it.  (The first version spent ~100 uniform-datapath
            // instructions per tap on this arithmetic: as much issue time as the four MMAs of a single-plane tap execute.)
            const uint64_t a_const = make_desc_mn(0, W_HI_OFFSET, 1024);           // [Wlo ; Whi] halves 8 KB apart
            const uint64_t a2_const = make_desc_mn(0, 0, 1024);                    // second instruction: LBO filled in per stage
            const uint64_t b_const = make_desc_sbo(0, ZP * 128);
            // (masked to the CTA-local 256 KB window: in a cluster launch the shared address of a CTA carries its rank in the upper
            // bits, which must not leak into the descriptor fields next to the 14-bit address)
            const uint32_t wsm4 = (smem_u32(wsm) >> 4) & 0x3FFFu, zaddr4 = (smem_u32(zsm) >> 4) & 0x3FFFu, xs4 = (smem_u32(xs) >> 4) & 0x3FFFu;
            uint32_t ti = 0, wi = 0;
            uint32_t xs_i = 0, xph = 0;                  // activation stage and its phase
            uint32_t ws = 0, wph = 0;                    // weight stage and its phase
            uint32_t gfill = 0, gidx = 0;                // taps consumed of the current stage group, the group's barrier
            long long wt = 0, wx = 0, ww = 0, c0 = 0;
            const long long tbeg = p.dbg ? clock64() : 0;
            for (int L = 0; L < nl; ++L)
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
                const uint32_t buf = ti & 1;
                const uint32_t dacc = tmem_base + buf * ACC_STRIDE;
                if (p.dbg) c0 = clock64();
                mbar_wait(&t_empty[buf], ((ti >> 1) & 1) ^ 1);
                if (p.dbg) wt += clock64() - c0;
                tc_fence_after();
#pragma unroll 1
                for (int dx = 0; dx < 3; ++dx) {
                    if (p.dbg) c0 = clock64();
                    mbar_wait(&x_full[xs_i], xph);
                    if (p.dbg) wx += clock64() - c0;
                    tc_fence_after();
                    const uint32_t xhi4 = xs4 + xs_i * (C::XSTAGE_BYTES >> 4);
                    const uint32_t xlo4 = xhi4 + (C::PART_BYTES >> 4);
                    const uint32_t acc0 = dx != 0;                                  // the tile's first MMA overwrites the accumulator
                    if constexpr (SINGLE) {
                        // The single-plane dgrad keeps the rolled loop with its index arithmetic: with the lean loop below its HR
                        // class went from 4.25 to 4.95 ms (same-box A/B, profiles/r02_issue_loop_ab.txt) -- the kernel is bound by
                        // its epilogue and by weight-stage latency, and an MMA warp that reaches its waits early only spins there,
                        // on the scheduler it shares with two epilogue warps.
                        const uint32_t xhi = smem_u32(xs + xs_i * C::XSTAGE_BYTES);
#pragma unroll 1
                        for (int tp = 0; tp < 9; ++tp) {
                            const uint32_t wst = wi % C::NWS, wphase = (wi / C::NWS) & 1;
                            if (p.dbg) c0 = clock64();
                            mbar_wait(&w_full[wst], wphase);
                            if (p.dbg) ww += clock64() - c0;
                            tc_fence_after();
                            const uint32_t wa = smem_u32(wsm + wst * W_STAGE_BYTES);
                            const uint32_t boff = ((tp / 3) * ZP + (tp % 3)) * 128;     // (dy, dz) row shift
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t acc = (dx | tp | k) != 0;
                                if (leader)
                                    tc_mma_f16(dacc, make_desc_mn(wa + k * 2048, W_HI_OFFSET, 1024),
                                               make_desc_sbo(xhi + boff + k * 32, ZP * 128), idesc, acc);
                            }
                            if (wi % C::WG == C::WG - 1 && leader) {
                                if (CL) tc_commit_mc(&w_empty[(wi / C::WG) % C::NGW], (uint16_t)3);
                                else tc_commit(&w_empty[(wi / C::WG) % C::NGW]);
                            }
                            __syncwarp();
                            ++wi;
                        }
                    } else {
#pragma unroll
                    for (int tp = 0; tp < 9; ++tp) {
                        if (p.dbg) c0 = clock64();
                        mbar_wait(&w_full[ws], wph);
                        if (p.dbg) ww += clock64() - c0;
                        tc_fence_after();
                        const uint32_t wk4 = wsm4 + ws * (W_STAGE_BYTES >> 4);
                        const uint32_t boff4 = ((tp / 3) * ZP + (tp % 3)) * 8;      // (dy, dz) row shift, in 16-byte units (an immediate: tp is unrolled)
                        // [Wlo ; Whi] x Xhi, then [Whi ; zeros] x Xlo (lanes 64-127 masked off) into the same accumulator.  The
                        // weights are MN-major, so the two 64-row halves of the A operand are independent atoms (LBO apart):
                        // the second instruction pairs the Whi half with the zero block instead of whatever follows it in
                        // shared memory -- the masked rows are still multiplied, zeros cost less (profiles/r02_power_probe.txt)
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t acc = (tp | k) != 0 ? 1u : acc0;
                                tc_mma_f16(dacc, a_const | (uint64_t)(wk4 + k * 128), b_const | (uint64_t)(xhi4 + boff4 + k * 2), idesc, acc);
                                if (!SINGLE) {
                                    const uint32_t a2 = wk4 + (W_HI_OFFSET >> 4) + k * 128;
                                    tc_mma_f16_masked(dacc, a2_const | (uint64_t)a2 | ((uint64_t)(zaddr4 - a2) << 16),
                                                      b_const | (uint64_t)(xlo4 + boff4 + k * 2), idesc, 1u, 0u, 0u, ~0u, ~0u);
                                }
                            }
                        }
                        if (++gfill == C::WG) {                                    // hand the group of stages back (to both producers of a pair)
                            if (leader) {
                                if (CL) tc_commit_mc(&w_empty[gidx], (uint16_t)3);
                                else tc_commit(&w_empty[gidx]);
                            }
                            gfill = 0;
                            gidx = gidx + 1 == C::NGW ? 0 : gidx + 1;
                        }
                        __syncwarp();
                        if (++ws == C::NWS) { ws = 0; wph ^= 1; }
                    }
                    }
                    if (++xs_i == C::NXS) { xs_i = 0; xph ^= 1; }
                }
                if (leader) tc_commit(&t_full[buf]);
            }
            if (p.dbg && leader) {
                p.dbg[blockIdx.x * 8 + 0] = wt; p.dbg[blockIdx.x * 8 + 1] = wx;
                p.dbg[blockIdx.x * 8 + 2] = ww; p.dbg[blockIdx.x * 8 + 3] = clock64() - tbeg;
                p.dbg[blockIdx.x * 8 + 5] = clock64();          // MMA issue loop end (absolute)
            }
        }
    } else {
        // ================= epilogue (warps 2..9): C::NG groups working on alternate chunks =================
        // warp w may only read TMEM lanes 32*(w%4)..+31: quadrants 0,1 hold the L rows (co = 32e + lane),
        // quadrants 2,3 the H rows; the C::CP warps of a quadrant (within a group) split a chunk's columns
        const int ew = warp - 2;                                   // 0..7
        const int grp = C::NG == 2 ? (ew >> 2) : 0;                // which group
        const int cpart = C::NG == 2 ? 0 : (ew >> 2);              // which part of the chunk's columns this warp reads
        const int e = warp & 3;
        const int gt = threadIdx.x - 64 - grp * C::GT;             // thread within the group
        const bool is_h = e >= 2;
        const int co = 32 * (e & 1) + lane;
        uint32_t ti = 0;
        float amax = 0.f;
        for (int L = 0; L < nl; ++L) {
        KParams pl;
        if (CHAIN) pl = chain.params[L];
        const KParams& p = CHAIN ? pl : p0;
        const float bias = (is_h && p.bias) ? p.bias[co] : 0.f;
        const float s1 = is_h ? 1.f : SR4D_LO_INV;
        const int Do = p.Do;
        const int Di = p.Dint;
        float* gstage = stage + grp * C::GSTAGE_FLOATS;
        float* my_stage = gstage + (is_h ? 0 : C::CV * 64) + co;
        float ksplit = 0.f;
        if (p.fused && p.split_hi) {
            float bound = 8.f * *p.gain * __uint_as_float(*p.dy_amax);
            if (p.add_amax) bound += __uint_as_float(*p.add_amax);
            int ex = 0;
            if (bound > 0.f && bound < 3.0e38f) ex = 14 - ilogbf(bound);     // bound * 2^e < 2^15
            ex = max(-120, min(120, ex));
            ksplit = exp2f((float)ex);
            if (blockIdx.x == 0 && threadIdx.x == 64) *p.split_exp = ex;
        }
        const float ks = (p.fused && p.dy_exp) ? exp2f((float)(-*p.dy_exp)) : 1.f;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
            const int b = t / tiles_per_b;
            int rem = t % tiles_per_b;
            const int x = rem / tiles_per_x;
            rem %= tiles_per_x;
            const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
            const bool x_edge = p.halo && (x == 0 || x == Do - 1);
            const uint32_t buf = ti & 1;
            // The loads of a chunk's second phase are issued ONE CHUNK AHEAD (and, outside the chained launches, those of a
            // tile's first chunk before the wait for its accumulator): the single-plane dgrad has ~11k cycles of MMAs per tile
            // and its MMA warp spent 15 % of its time waiting for the epilogue, whose chunks were each one exposed memory
            // latency long (SR4D_TC_DEBUG stamps).
            auto load_items = [&](int ch, EpiItems<C::IT>& it) {
#pragma unroll
                for (int k = 0; k < C::IT; ++k) {
                    const int item = gt + k * C::GT;
                    const int vq = item >> 3, g8 = item & 7;
                    const int line = ch * C::LPC + (vq >> 3);
                    const int y = y0 + line, z = z0 + (vq & 7);
                    it.vy[k] = y; it.vz[k] = z;
                    bool act = line < TY && y < Do && z < Do;
                    // fused dgrad: (y, z) are padded-grid coordinates; only interior voxels produce output
                    if (p.fused) act = act && y >= 1 && y <= Di && z >= 1 && z <= Di;
                    it.active[k] = act;
                    if (p.fused) {
                        if (act) {
                            const size_t go = g4_off(Di, b, x, y - 1, z - 1) + g8 * 8;
                            if (p.add_pre) {
                                it.ap0[k] = *reinterpret_cast<const float4*>(p.add_pre + go);
                                it.ap1[k] = *reinterpret_cast<const float4*>(p.add_pre + go + 4);
                            }
                            if (p.add_post) {
                                it.aq0[k] = *reinterpret_cast<const float4*>(p.add_post + go);
                                it.aq1[k] = *reinterpret_cast<const float4*>(p.add_post + go + 4);
                            }
                            if (p.sav_hi) {
                                const size_t o = act_off(Di, b, x, y - 1, z - 1) + g8 * 8;
                                it.rh[k] = *reinterpret_cast<const uint4*>(p.sav_hi + o);
                                it.rl[k] = *reinterpret_cast<const uint4*>(p.sav_lo + o);
                            }
                        }
                    } else if (p.res_hi && act) {
                        const size_t o = act_off(Do, b, x, y, z) + g8 * 8;
                        if (CHAIN) {
                            // the residual was written earlier in THIS launch, possibly into a buffer this SM has read before
                            // (inference reuses activation slots): bypass the non-coherent L1
                            it.rh[k] = __ldcg(reinterpret_cast<const uint4*>(p.res_hi + o));
                            it.rl[k] = __ldcg(reinterpret_cast<const uint4*>(p.res_lo + o));
                        } else {
                            it.rh[k] = *reinterpret_cast<const uint4*>(p.res_hi + o);
                            it.rl[k] = *reinterpret_cast<const uint4*>(p.res_lo + o);
                        }
                    }
                }
            };
            // Only the single-plane dgrad does this: the two-plane kernels have twice the MMA time per tile, their epilogue is
            // never waited for, and the forward measured 1-2 % slower with the prefetch (profiles/r02_epilogue_prefetch_ab.txt).
            constexpr bool PREFETCH = SINGLE && !CHAIN;
            EpiItems<C::IT> nxt;
            if (PREFETCH && grp < C::NCHUNK) load_items(grp, nxt);
            mbar_wait(&t_full[buf], (ti >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + buf * ACC_STRIDE + ((uint32_t)(32 * e) << 16);
            if (grp >= C::NCHUNK) {
                // this group has no chunk of the tile: hand the accumulator back right away
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[buf]);
            }
#pragma unroll 1
            for (int ch = grp; ch < C::NCHUNK; ch += C::NG) {
                EpiItems<C::IT> cur;
                if (PREFETCH) {
                    cur = nxt;
                    if (ch + C::NG < C::NCHUNK) load_items(ch + C::NG, nxt);
                } else {
                    load_items(ch, cur);
                }
                // ---- phase A: TMEM -> registers -> fp32 staging [part][voxel][channel] ----
                const int c0 = ch * C::CV + cpart * C::COLS;           // first accumulator column this warp reads
#pragma unroll
                for (int cc = 0; cc < C::COLS; cc += 16) {
                    if (c0 + cc < C::N) {                              // N is a multiple of 16
                        float a[16];
                        tc_ld16(trow + c0 + cc, a);
                        tc_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) my_stage[(cpart * C::COLS + cc + j) * 64] = fmaf(a[j], s1, bias);
                    }
                }
                if (ch + C::NG >= C::NCHUNK) {
                    // all TMEM reads of this tile by this warp are done: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[buf]);
                }
                named_bar(1 + grp, C::GT);
                // ---- phase B: coalesced residual / activation / split / store ----
#pragma unroll
                for (int k = 0; k < C::IT; ++k) {
                    if (!cur.active[k]) continue;
                    const int item = gt + k * C::GT;
                    const int vq = item >> 3, g8 = item & 7;
                    const int y = cur.vy[k], z = cur.vz[k];
                    const float* sp = gstage + vq * 64 + g8 * 8;
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
#pragma unroll
                        for (int c = 0; c < 8; ++c) val[c] *= ks;
                        if (p.add_pre) {
                            val[0] += cur.ap0[k].x; val[1] += cur.ap0[k].y; val[2] += cur.ap0[k].z; val[3] += cur.ap0[k].w;
                            val[4] += cur.ap1[k].x; val[5] += cur.ap1[k].y; val[6] += cur.ap1[k].z; val[7] += cur.ap1[k].w;
                        }
                        if (p.sav_hi) {
                            const __half2* hh = reinterpret_cast<const __half2*>(&cur.rh[k]);
                            const __half2* ll = reinterpret_cast<const __half2*>(&cur.rl[k]);
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float2 ha = __half22float2(hh[c]), la = __half22float2(ll[c]);
                                val[2 * c] *= act_grad_from_out(fmaf(la.x, SR4D_LO_INV, ha.x), p.slope);
                                val[2 * c + 1] *= act_grad_from_out(fmaf(la.y, SR4D_LO_INV, ha.y), p.slope);
                            }
                        }
                        if (p.add_post) {
                            val[0] += cur.aq0[k].x; val[1] += cur.aq0[k].y; val[2] += cur.aq0[k].z; val[3] += cur.aq0[k].w;
                            val[4] += cur.aq1[k].x; val[5] += cur.aq1[k].y; val[6] += cur.aq1[k].z; val[7] += cur.aq1[k].w;
                        }
                        const size_t go = g4_off(Di, b, x, y - 1, z - 1) + g8 * 8;
                        if (p.out_g4) {                  // NULL when every consumer reads the split copy (mid-block gradients)
                            float* o = p.out_g4 + go;
                            *reinterpret_cast<float4*>(o) = make_float4(val[0], val[1], val[2], val[3]);
                            *reinterpret_cast<float4*>(o + 4) = make_float4(val[4], val[5], val[6], val[7]);
                        }
                        if (p.split_hi) {
                            __align__(16) __half hv[8];
                            __align__(16) __half lv[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) split_f16(val[c] * ksplit, hv[c], lv[c]);
                            *reinterpret_cast<uint4*>(p.split_hi + go) = *reinterpret_cast<const uint4*>(hv);
                            if (p.split_lo) *reinterpret_cast<uint4*>(p.split_lo + go) = *reinterpret_cast<const uint4*>(lv);
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c) amax = fmaxf(amax, fabsf(val[c]));
                    } else if (p.out_raw) {
                        float* o = p.out_raw + ((((size_t)b * Do + x) * Do + y) * Do + z) * 64 + g8 * 8;
                        *reinterpret_cast<float4*>(o) = make_float4(val[0], val[1], val[2], val[3]);
                        *reinterpret_cast<float4*>(o + 4) = make_float4(val[4], val[5], val[6], val[7]);
#pragma unroll
                        for (int c = 0; c < 8; ++c) amax = fmaxf(amax, fabsf(val[c]));
                    } else {
                        if (p.res_hi) {
                            const __half2* hh = reinterpret_cast<const __half2*>(&cur.rh[k]);
                            const __half2* ll = reinterpret_cast<const __half2*>(&cur.rl[k]);
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                float2 ha = __half22float2(hh[c]), la = __half22float2(ll[c]);
                                val[2 * c] += fmaf(la.x, SR4D_LO_INV, ha.x);
                                val[2 * c + 1] += fmaf(la.y, SR4D_LO_INV, ha.y);
                            }
                        }
                        __align__(16) __half hv[8];
                        __align__(16) __half lv[8];
                        bool over = false;
#pragma unroll
                        for (int c = 0; c < 8; ++c) over |= split_f16_chk(act_fn(val[c], p.slope), hv[c], lv[c]);
                        if (over && p.ovf) *p.ovf = 1u;
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
                named_bar(1 + grp, C::GT);
            }
            if (CHAIN && L + 1 < nl) {
                // tile done: its stores (halo replicas included) are ordered before the CTA barrier, the signalling thread's
                // device-scope RELEASE after it is cumulative over them (the grid.sync pattern, without its full fence);
                // the tile then counts for its x-plane.
                named_bar(3, NUM_EPI);
                if (threadIdx.x == 64) {
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(chain.done + ((size_t)L * p.B + b) * p.nx + x) : "memory");
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(chain.done + chain.nplane + L) : "memory");
                }
            }
        }
        }   // layers
        if (p.absmax) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            if (lane == 0) atomicMax(p.absmax, __float_as_uint(amax));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL) cluster_sync_all();              // the peer may still multicast commits / weight taps into this CTA
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
    if (CHAIN) {
        // the last CTA to leave re-arms the counters for the next chained launch (nobody polls them any more)
        __shared__ unsigned int is_last;
        if (threadIdx.x == 0) is_last = atomicAdd(chain.done + chain.ndone, 1u) == gridDim.x - 1 ? 1u : 0u;
        __syncthreads();
        if (is_last) {
            for (int i = threadIdx.x; i <= chain.ndone; i += NUM_THREADS) chain.done[i] = 0u;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------
// weight image: fp32 Keras [27][ci][co] -> fp16 split, per tap [Wlo half | Whi half], each MN-major (64 K rows x 64 M), swizzled
// ------------------------------------------------------------------------------------------
struct PrepList {
    int n;
    int layer[64];           // image slot of entry i
    long long off[64];       // float offset of its Keras kernel in the flat parameter buffer
};
// grid (27*128*64/256, n entries, 2 images): one thread per (tap, row, k) element
__global__ void prep_weights_kernel(const float* __restrict__ params, PrepList l, __half* __restrict__ images) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // = offset inside the layer's image: coalesced stores
    if (i >= 27 * 128 * 64) return;
    const float* w = params + l.off[blockIdx.y];
    const int dgrad = blockIdx.z;
    __half* img = images + ((size_t)l.layer[blockIdx.y] * 2 + dgrad) * (27 * 128 * 64);
    // MN-major SWIZZLE_128B: per tap [lo half | hi half]; inside a half the shared-memory row is the K index k (8-row groups
    // of 1024 B), a row holds the 64 M elements n in 16-byte pieces stored at piece ^ (k & 7)
    const int tap = i >> 13, is_lo = ((i >> 12) & 1) == 0;
    const int grp = (i >> 9) & 7, rr = (i >> 6) & 7, piece = (i >> 3) & 7, e = i & 7;
    const int k = grp * 8 + rr, n = ((piece ^ rr) << 3) + e;
    // forward: A[n=co][k=ci] = W[tap][ci][co];  dgrad: A[n=ci][k=co] = W[26-tap][ci][co]
    const float v = dgrad ? w[((size_t)(26 - tap) * 64 + n) * 64 + k] : w[((size_t)tap * 64 + k) * 64 + n];
    __half h, lo;
    split_f16(v, h, lo);
    img[i] = is_lo ? lo : h;
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
template <int TY, bool SINGLE>
cudaError_t launch_cfg(const CUtensorMap& map, KParams p, cudaStream_t s) {
    using C = Cfg<TY, SINGLE>;
    cudaError_t ea = tc_func_smem(reinterpret_cast<const void*>(conv64_tc_kernel<TY, SINGLE, false, false>), C::SMEM_BYTES);
    if (ea != cudaSuccess) return ea;
    p.nyt = (p.Do + TY - 1) / TY;
    p.nzt = (p.Do + TZ - 1) / TZ;
    p.ntiles = p.B * p.nx * p.nyt * p.nzt;
    const int sms = tc_num_sms();
    int grid = p.ntiles < sms ? p.ntiles : sms;
    ChainArgs none;
    none.params = nullptr; none.maps = nullptr; none.n = 0; none.done = nullptr; none.nplane = 0; none.ndone = 0;
    // CTA pairs: both CTAs of a pair must run the same tap sequence (consecutive tiles of one x-plane: even tiles per plane)
    // and the same number of tiles (even tile count on an even grid)
    static const bool cluster_env = !(getenv("SR4D_TC_CLUSTER") != nullptr && atoi(getenv("SR4D_TC_CLUSTER")) == 0);   // default on
    static bool cluster_ok = true;          // cleared when a cluster launch is refused (e.g. a partition that cannot co-schedule pairs)
    if (cluster_env && cluster_ok && p.ntiles % 2 == 0 && (p.nyt * p.nzt) % 2 == 0 && grid >= 2 && !p.dbg) {
        ea = tc_func_smem(reinterpret_cast<const void*>(conv64_tc_kernel<TY, SINGLE, false, true>), C::SMEM_BYTES);
        if (ea != cudaSuccess) return ea;
        grid &= ~1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t ec = cudaLaunchKernelEx(&cfg, conv64_tc_kernel<TY, SINGLE, false, true>, map, p, none);
        if (ec == cudaSuccess) return ec;
        (void)cudaGetLastError();           // launch-configuration error (not sticky): run the single-CTA kernel from now on
        cluster_ok = false;
        grid = p.ntiles < sms ? p.ntiles : sms;
    }
    conv64_tc_kernel<TY, SINGLE, false, false><<<grid, NUM_THREADS, C::SMEM_BYTES, s>>>(map, p, none);
    return cudaGetLastError();
}

// chained forward layers: p0 carries the (shared) geometry, the per-layer blocks live in device memory.  The persistent
// CTAs wait for each other between layers, so all of them must be co-resident: cooperative launch.
template <int TY>
cudaError_t launch_chain_cfg(KParams p0, ChainArgs ch, cudaStream_t s) {
    using C = Cfg<TY, false>;
    const int sms = tc_num_sms();
    int grid = p0.ntiles < sms ? p0.ntiles : sms;
    CUtensorMap dummy;
    memset(&dummy, 0, sizeof dummy);
    // Optional CTA pairs with multicast weight taps (see launch_cfg).  Both CTAs of a pair must hold the same number of tiles
    // of every layer: even tile count on an even grid.  Cooperative (co-residency: the CTAs wait for each other between
    // layers) + cluster launch; a refusal falls back to single CTAs for the rest of the process.
    // Measured (profiles/r02_chain_cluster_ab.txt): batch-1 forward 752.7 vs 748.7 patches/s, batch-8 forward 7.42 vs 7.63 ms --
    // the chained grids are power-limited like the rest, not L2-limited: opt-in only (SR4D_CHAIN_CLUSTER=1).
    static const bool cluster_env = getenv("SR4D_CHAIN_CLUSTER") != nullptr && atoi(getenv("SR4D_CHAIN_CLUSTER")) != 0;
    static bool cluster_ok = true;
    if (cluster_env && cluster_ok && p0.ntiles % 2 == 0 && grid >= 2) {
        cudaError_t ea = tc_func_smem(reinterpret_cast<const void*>(conv64_tc_kernel<TY, false, true, true>), C::SMEM_BYTES);
        if (ea != cudaSuccess) return ea;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid & ~1); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 2;
        cudaError_t ec = cudaLaunchKernelEx(&cfg, conv64_tc_kernel<TY, false, true, true>, dummy, p0, ch);
        if (ec == cudaSuccess) return ec;
        (void)cudaGetLastError();
        cluster_ok = false;
    }
    cudaError_t ea = tc_func_smem(reinterpret_cast<const void*>(conv64_tc_kernel<TY, false, true, false>), C::SMEM_BYTES);
    if (ea != cudaSuccess) return ea;
    void* args[3] = {&dummy, &p0, &ch};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(conv64_tc_kernel<TY, false, true, false>), dim3(grid), dim3(NUM_THREADS),
                                       args, C::SMEM_BYTES, s);
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

// forward grids: the tile height that finishes first on this machine.  A tile of N = 8*TY voxels costs
// 27 taps x 4 K-steps x 2 instructions of max(N/2, 32 + N/4) cycles (tensor pipe vs shared-memory operand fetch,
// profiles/r01_tcgen05_probe.txt); the persistent grid runs ceil(tiles / SMs) rounds.  At batch 8 the tall tiles win
// (TY = 24 on the 24^3 / 48^3 grids); at batch 1 the 24^3 grid has only 72 such tiles for 148 SMs and TY = 12 (144
// tiles, one round of cheaper tiles) is ~40 % faster (configs[0], predictor.py at batch 1).
int pick_ty_fwd(int Do, int B) {
    if (Do <= 8) return 8;
    const int sms = tc_num_sms();
    const int cand[5] = {8, 12, 16, 24, 26};
    int best = 24;
    long best_cost = -1;
    for (int c : cand) {
        const long tiles = (long)B * Do * ((Do + c - 1) / c) * ((Do + TZ - 1) / TZ);
        const long rounds = (tiles + sms - 1) / sms;
        const int n = c * TZ;
        const long per_tile = 216L * (n / 2 > 32 + n / 4 ? n / 2 : 32 + n / 4) + 3000;    // + epilogue / pipeline fill
        const long cost = rounds * per_tile;
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && c > best)) { best_cost = cost; best = c; }
    }
    return best;
}

long tc_fwd_tiles(int Do, int B) {
    const int ty = pick_ty_fwd(Do, B);
    return (long)B * Do * ((Do + ty - 1) / ty) * ((Do + TZ - 1) / TZ);
}

bool tc_dgrad_fusable(int D) {
    if (D < 2) return false;
    const int ty = pick_ty(D + 2);
    const int lpc = 4;
    auto chunk = [&](int line) { return (line / ty) * 1000 + (line % ty) / lpc; };
    return chunk(0) == chunk(1) && chunk(D) == chunk(D + 1) && D / TZ == (D + 1) / TZ;
}

namespace {
// the kernel's parameter block of one layer call (everything but the tile geometry and the debug buffer)
cudaError_t fill_params(TcWeights* w, const TcConvArgs& a, KParams& p) {
    const int Do = a.in.D, B = a.in.B;
    if (!a.out_raw && !a.fused && a.out.D != Do) return cudaErrorInvalidValue;
    if (a.fused && (!a.dgrad || (!a.out_g4 && !a.split_out) || !tc_dgrad_fusable(Do - 2))) return cudaErrorInvalidValue;
    p.w_img = w->img + ((size_t)a.layer * 2 + (a.dgrad ? 1 : 0)) * 27 * 128 * 64;
    p.out_hi = a.out.hi; p.out_lo = a.out.lo;
    p.res_hi = a.res_hi; p.res_lo = a.res_lo;
    p.bias = a.bias; p.out_raw = a.out_raw; p.absmax = a.absmax; p.ovf = a.out.ovf;
    p.slope = a.slope; p.B = B; p.Do = Do; p.halo = a.halo;
    p.nx = a.fused ? Do - 2 : Do;
    p.fused = a.fused; p.Dint = Do - 2;
    p.single_b = (a.dgrad && a.single_b) ? 1 : 0;
    p.dy_exp = a.dy_exp; p.add_pre = a.add_pre; p.add_post = a.add_post;
    p.sav_hi = a.sav_hi; p.sav_lo = a.sav_lo; p.out_g4 = a.out_g4;
    p.split_hi = a.split_out;
    p.split_lo = (a.split_out && !a.split_hi_only) ? a.split_out + act_plane_elems(B, Do) : nullptr;   // [2B][D+4]^3[64]: hi planes then lo planes
    p.split_exp = a.split_exp; p.gain = w->gain + a.layer; p.dy_amax = a.dy_amax; p.add_amax = a.add_amax;
    if (a.split_out && (!a.split_exp || !a.dy_amax || !a.fused)) return cudaErrorInvalidValue;
    static const bool xsplit = !(getenv("SR4D_TC_XSPLIT") && atoi(getenv("SR4D_TC_XSPLIT")) == 0);
    p.xsplit = xsplit;
    p.dbg = nullptr;
    p.nyt = p.nzt = p.ntiles = 0;
    if (a.in.lo != a.in.hi + act_plane_elems(B, a.in.D)) return cudaErrorInvalidValue;   // planes must be packed
    return cudaSuccess;
}
}  // namespace

struct TcChain {
    int n = 0, ty = 0;
    KParams p0;
    KParams* dparams = nullptr;
    CUtensorMap* dmaps = nullptr;
    unsigned int* done = nullptr;
    int nplane = 0, ndone = 0;
};

cudaError_t tc_chain_build(TcWeights* w, const TcConvArgs* a, int n, TcChain** out) {
    if (n < 1 || n > 64) return cudaErrorInvalidValue;
    const int Do = a[0].in.D, B = a[0].in.B, Dp = Do + 2;
    const int ty = pick_ty_fwd(Do, B);
    std::vector<KParams> hp(n);
    std::vector<CUtensorMap> hm(n);
    for (int i = 0; i < n; ++i) {
        if (a[i].dgrad || a[i].fused || a[i].out_raw || a[i].in.D != Do || a[i].in.B != B) return cudaErrorInvalidValue;
        cudaError_t e = fill_params(w, a[i], hp[i]);
        if (e != cudaSuccess) return e;
        hp[i].nyt = (Do + ty - 1) / ty;
        hp[i].nzt = (Do + TZ - 1) / TZ;
        hp[i].ntiles = B * hp[i].nx * hp[i].nyt * hp[i].nzt;
        if (!tc_make_act_map(&hm[i], a[i].in.hi, B, Dp, ty + 2, ZP)) return cudaErrorUnknown;
    }
    TcChain* c = new TcChain();
    c->n = n; c->ty = ty; c->p0 = hp[0];
    cudaError_t e = cudaMalloc((void**)&c->dparams, n * sizeof(KParams));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->dmaps, n * sizeof(CUtensorMap));
    c->nplane = (n - 1) * B * hp[0].nx;
    c->ndone = c->nplane + n - 1;
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->done, (size_t)(c->ndone + 1) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemcpy(c->dparams, hp.data(), n * sizeof(KParams), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(c->dmaps, hm.data(), n * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(c->done, 0, (size_t)(c->ndone + 1) * sizeof(unsigned int));
    if (e != cudaSuccess) { tc_chain_free(c); return e; }
    *out = c;
    return cudaSuccess;
}
void tc_chain_free(TcChain* c) {
    if (!c) return;
    cudaFree(c->dparams); cudaFree(c->dmaps); cudaFree(c->done);
    delete c;
}
int tc_chain_layers(const TcChain* c) { return c ? c->n : 0; }
cudaError_t tc_chain_launch(TcChain* c, cudaStream_t s) {
    ChainArgs ch;
    ch.params = c->dparams; ch.maps = c->dmaps; ch.n = c->n; ch.done = c->done; ch.nplane = c->nplane; ch.ndone = c->ndone;
    switch (c->ty) {
        case 8: return launch_chain_cfg<8>(c->p0, ch, s);
        case 12: return launch_chain_cfg<12>(c->p0, ch, s);
        case 16: return launch_chain_cfg<16>(c->p0, ch, s);
        case 24: return launch_chain_cfg<24>(c->p0, ch, s);
        default: return launch_chain_cfg<26>(c->p0, ch, s);
    }
}

cudaError_t tc_conv64(TcWeights* w, const TcConvArgs& a, cudaStream_t s) {
    const int Do = a.in.D, B = a.in.B, Dp = a.in.D + 2;
    KParams p;
    cudaError_t ep = fill_params(w, a, p);
    if (ep != cudaSuccess) return ep;
    p.dbg = nullptr;
    static const bool debug = getenv("SR4D_TC_DEBUG") != nullptr;
    static long long* dbg_buf = nullptr;
    if (debug) {
        if (!dbg_buf) cudaMalloc((void**)&dbg_buf, (148 * 8 + 2) * sizeof(long long));
        cudaMemsetAsync(dbg_buf, 0, (148 * 8 + 2) * sizeof(long long), s);
        cudaMemsetAsync(dbg_buf + 148 * 8, 0x7f, sizeof(long long), s);      // min slot starts high
        p.dbg = dbg_buf;
    }
    const int ty = a.dgrad ? pick_ty(Do) : pick_ty_fwd(Do, B);
    CUtensorMap map;
    if (!tc_make_act_map(&map, a.in.hi, B, Dp, ty + 2, ZP)) return cudaErrorUnknown;
    cudaError_t e;
    if (p.single_b) {
        switch (ty) {
            case 8: e = launch_cfg<8, true>(map, p, s); break;
            case 16: e = launch_cfg<16, true>(map, p, s); break;
            case 24: e = launch_cfg<24, true>(map, p, s); break;
            default: e = launch_cfg<26, true>(map, p, s); break;
        }
    } else {
        switch (ty) {
            case 8: e = launch_cfg<8, false>(map, p, s); break;
            case 12: e = launch_cfg<12, false>(map, p, s); break;
            case 16: e = launch_cfg<16, false>(map, p, s); break;
            case 24: e = launch_cfg<24, false>(map, p, s); break;
            default: e = launch_cfg<26, false>(map, p, s); break;
        }
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
