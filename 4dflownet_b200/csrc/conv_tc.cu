// tcgen05 (5th-gen tensor core) 64->64 3x3x3 convolution for sm_100a, split-fp16 "fp32-accurate".
//
// Reference op: conv3d() of Network/SR4DFlowNet.py:93-108 / resnet_block :111-120 (30 of the
// 36 layers, 99.5 % of the FLOPs); with the dgrad weight image the same kernel evaluates
// Conv3DBackpropInput on the padded grid.
//
// Formulation (implicit GEMM, weights as the M operand so the voxel dimension is the flexible N):
//     D[128 x N] += A_tap[128 x 64] * B_tap[N x 64]^T        for the 27 taps, K = 64 channels
//   A_tap rows  = the layer's weights for that tap, split in fp16: W = Whi + Wlo/2048.
//                 Row 32q+l holds Whi[co=16q+l] for l<16 and Wlo[co=16q+l-16] for l>=16, so
//                 the hi and lo partial sums of one output channel sit in lanes l and l^16 of
//                 the same epilogue warp (one shuffle combines them).
//   B_tap rows  = the N = 8*TY voxels of a (TY y-lines x 8 z) tile at a fixed x: their 64 input
//                 channels, from the activation's fp16 hi plane (accumulator D1) and lo plane (D2).
//   out[co][v]  = D1[hi] + (D1[lo] + D2[hi]) / 2048 + D2[lo] / 2048^2     (fp32 in TMEM)
//
// Data movement: ONE TMA box per dx loads the (TY+2) x 10 voxel halo plane (hi and lo) into
// shared memory as dense 128-byte rows (SWIZZLE_128B).  All nine (dy,dz) taps of that plane are
// then plain descriptor start-address shifts of (dy*10+dz) rows with an 8-row-group stride of
// 1280 B: the tensor core applies the 128B swizzle on absolute shared-memory address bits, so
// shifted starts and a non-1024 group stride read the TMA-written rows correctly (verified on
// B200 by tools/probe/mma_probe.cu).  Every activation byte is fetched 3x from L2 per layer
// (plus y/z halo), weights stream per tap as pre-swizzled 16 KB images (cp.async.bulk).
// Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue:
// TMEM -> registers -> (hi/lo row combine, bias) -> fp32 shared-memory transpose -> coalesced
// 16-byte residual loads / activation / fp16 split / stores with the replicate halo.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "conv_tc.h"
#include "tc_ptx.cuh"

namespace {

constexpr int W_TAP_BYTES = 128 * 64 * 2;   // 16 KB: [Whi;Wlo] x 64 ci, fp16, swizzled
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI = NUM_EPI_WARPS * 32;   // epilogue threads
constexpr int NUM_THREADS = 64 + NUM_EPI;
constexpr int TZ = 8;                       // voxels per 8-row group (one z run)
constexpr int ZP = TZ + 2;                  // plane row pitch in voxels
constexpr int STAGE_FLOATS = 64 * 64;       // epilogue transpose buffer: 64 voxels x 64 channels

template <int TY>
struct Cfg {
    static constexpr int N = TY * TZ;                                  // voxels per tile = MMA N
    static constexpr int ROWS = (TY + 2) * ZP;                         // rows of one staged plane part
    static constexpr int PART_BYTES = (ROWS * 128 + 1023) / 1024 * 1024;
    static constexpr int XSTAGE_BYTES = 2 * PART_BYTES;                // hi part | lo part
    static constexpr int NXS = 2;
    static constexpr int NWS = 4;
    static constexpr int SMEM_BYTES = 1024 + NXS * XSTAGE_BYTES + NWS * W_TAP_BYTES + STAGE_FLOATS * 4 + 256;
    static constexpr int TMEM_COLS = 2 * N <= 32 ? 32 : 2 * N <= 64 ? 64 : 2 * N <= 128 ? 128 : 2 * N <= 256 ? 256 : 512;
    static constexpr int NCHUNK = TY / 8;                              // epilogue chunks of 64 voxels
    static_assert(TY % 8 == 0 && N % 16 == 0 && N >= 16 && N <= 256, "UMMA N constraint for M=128");
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
    long long* dbg;          // SR4D_TC_DEBUG=1: per-CTA cycles the MMA warp spent waiting {t_empty, x_full, w_full, total}
};

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
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* xs = smem;                                   // NXS x [hi part | lo part]
    uint8_t* wsm = smem + C::NXS * C::XSTAGE_BYTES;       // NWS x 16 KB
    float* stage = reinterpret_cast<float*>(wsm + C::NWS * W_TAP_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stage) + STAGE_FLOATS * 4);
    uint64_t* x_full = bars;                 // [NXS]
    uint64_t* x_empty = bars + C::NXS;       // [NXS]
    uint64_t* w_full = bars + 2 * C::NXS;    // [NWS]
    uint64_t* w_empty = w_full + C::NWS;     // [NWS]
    uint64_t* t_full = w_empty + C::NWS;     // [1]
    uint64_t* t_empty = t_full + 1;          // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NXS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < C::NWS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        mbar_init(t_full, 1);
        mbar_init(t_empty, NUM_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&xmap);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_x = p.nyt * p.nzt;
    const int tiles_per_b = p.Do * tiles_per_x;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t xi = 0, wi = 0;   // running stage counters
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                const int b = t / tiles_per_b;
                int rem = t % tiles_per_b;
                const int x = rem / tiles_per_x;
                rem %= tiles_per_x;
                const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
                for (int dx = 0; dx < 3; ++dx) {
                    const uint32_t s = xi % C::NXS, ph = (xi / C::NXS) & 1;
                    mbar_wait(&x_empty[s], ph ^ 1);
                    mbar_expect_tx(&x_full[s], 2 * C::ROWS * 128);
                    uint8_t* dst = xs + s * C::XSTAGE_BYTES;
                    tma_load_5d(dst, &xmap, &x_full[s], 0, z0, y0, x + dx, b);
                    tma_load_5d(dst + C::PART_BYTES, &xmap, &x_full[s], 0, z0, y0, x + dx, p.B + b);
                    ++xi;
                    for (int tp = 0; tp < 9; ++tp) {
                        const uint32_t ws = wi % C::NWS, wph = (wi / C::NWS) & 1;
                        mbar_wait(&w_empty[ws], wph ^ 1);
                        mbar_expect_tx(&w_full[ws], W_TAP_BYTES);
                        bulk_load(wsm + ws * W_TAP_BYTES,
                                  reinterpret_cast<const uint8_t*>(p.w_img) + (size_t)(dx * 9 + tp) * W_TAP_BYTES,
                                  W_TAP_BYTES, &w_full[ws]);
                        ++wi;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N
            const uint32_t idesc = (1u << 4) | ((uint32_t)(C::N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t d1 = tmem_base, d2 = tmem_base + C::N;
            uint32_t xi = 0, wi = 0, ti = 0;
            long long wt = 0, wx = 0, ww = 0, c0 = 0, tbeg = p.dbg ? clock64() : 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
                if (p.dbg) c0 = clock64();
                mbar_wait(t_empty, (ti & 1) ^ 1);
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
                        const uint32_t wa = smem_u32(wsm + ws * W_TAP_BYTES);
                        const uint32_t boff = ((tp / 3) * ZP + (tp % 3)) * 128;     // (dy, dz) row shift
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ad = make_desc_sbo(wa + k * 32, 1024);
                            const uint32_t acc = (dx | tp | k) != 0;
                            tc_mma_f16(d1, ad, make_desc_sbo(xhi + boff + k * 32, ZP * 128), idesc, acc);
                            tc_mma_f16(d2, ad, make_desc_sbo(xlo + boff + k * 32, ZP * 128), idesc, acc);
                        }
                        tc_commit(&w_empty[ws]);
                        ++wi;
                    }
                    tc_commit(&x_empty[s]);
                    ++xi;
                }
                tc_commit(t_full);
            }
            if (p.dbg) {
                p.dbg[blockIdx.x * 4 + 0] = wt; p.dbg[blockIdx.x * 4 + 1] = wx;
                p.dbg[blockIdx.x * 4 + 2] = ww; p.dbg[blockIdx.x * 4 + 3] = clock64() - tbeg;
            }
        }
    } else {
        // ================= epilogue (warps 2..9, 256 threads) =================
        // warp w may only read TMEM lanes 32*(w%4)..+31; the two warps of a lane quarter split the columns
        const int e = warp & 3;
        const int half = (warp - 2) >> 2;            // which 32 of a chunk's 64 columns this warp converts
        const int et = threadIdx.x - 64;             // 0..255
        const int co = 16 * e + (lane & 15);
        const bool is_lo = lane >= 16;
        const float bias = p.bias ? p.bias[co] : 0.f;
        const float s1 = is_lo ? SR4D_LO_INV : 1.f;
        const int Do = p.Do;
        const int g8 = et & 7;                       // 8-channel group handled in the store phase
        float amax = 0.f;
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
            const int b = t / tiles_per_b;
            int rem = t % tiles_per_b;
            const int x = rem / tiles_per_x;
            rem %= tiles_per_x;
            const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
            const bool x_edge = p.halo && (x == 0 || x == Do - 1);
            mbar_wait(t_full, ti & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(32 * e) << 16);
#pragma unroll 1
            for (int ch = 0; ch < C::NCHUNK; ++ch) {
                // residual prefetch for the 2 (voxel, channel-group) items this thread stores
                uint4 rh[2], rl[2];
                if (p.res_hi) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int v = (et >> 3) + 32 * r;
                        const int y = y0 + ch * 8 + (v >> 3), z = z0 + (v & 7);
                        if (y < Do && z < Do) {
                            const size_t o = act_off(Do, b, x, y, z) + g8 * 8;
                            rh[r] = *reinterpret_cast<const uint4*>(p.res_hi + o);
                            rl[r] = *reinterpret_cast<const uint4*>(p.res_lo + o);
                        }
                    }
                }
                // ---- phase A: TMEM -> registers -> fp32 transpose buffer [voxel][channel] ----
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    const int q = half * 2 + qq;
                    const int c0 = ch * 64 + q * 16;
                    float a[16], d[16];
                    tc_ld16(trow + c0, a);
                    tc_ld16(trow + C::N + c0, d);
                    tc_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float v = fmaf(d[j], SR4D_LO_INV, a[j]) * s1;
                        a[j] = v + __shfl_xor_sync(0xffffffffu, v, 16);
                    }
                    // lanes 0..15 keep columns 0..7 (y-line 2q), lanes 16..31 columns 8..15 (y-line 2q+1)
                    const int vb = q * 16 + (is_lo ? 8 : 0);
                    const int sw = co ^ (is_lo ? 16 : 0);      // bank swizzle: ((v >> 3) & 1) << 4
#pragma unroll
                    for (int j = 0; j < 8; ++j) stage[(vb + j) * 64 + sw] = (is_lo ? a[8 + j] : a[j]) + bias;
                }
                if (ch == C::NCHUNK - 1) {
                    // all TMEM reads of this tile are done: let the MMA warp start the next tile
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty);
                }
                named_bar(1, NUM_EPI);
                // ---- phase B: coalesced residual / activation / split / store ----
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int v = (et >> 3) + 32 * r;
                    const int y = y0 + ch * 8 + (v >> 3), z = z0 + (v & 7);
                    if (y >= Do || z >= Do) continue;
                    const float* sp = stage + v * 64 + ((g8 * 8) ^ (((v >> 3) & 1) << 4));
                    float4 f0 = *reinterpret_cast<const float4*>(sp);
                    float4 f1 = *reinterpret_cast<const float4*>(sp + 4);
                    float val[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                    if (p.out_raw) {
                        float* o = p.out_raw + ((((size_t)b * Do + x) * Do + y) * Do + z) * 64 + g8 * 8;
                        *reinterpret_cast<float4*>(o) = f0;
                        *reinterpret_cast<float4*>(o + 4) = f1;
#pragma unroll
                        for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(val[k]));
                        continue;
                    }
                    if (p.res_hi) {
                        const __half2* hh = reinterpret_cast<const __half2*>(&rh[r]);
                        const __half2* ll = reinterpret_cast<const __half2*>(&rl[r]);
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
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// weight image: fp32 Keras [27][ci][co] -> fp16 split, row-permuted, swizzled
// ------------------------------------------------------------------------------------------
__global__ void prep_weights_kernel(const float* __restrict__ w, __half* __restrict__ img, int dgrad) {
    // one thread per (tap, row, k)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 27 * 128 * 64) return;
    const int k = i & 63, row = (i >> 6) & 127, tap = i >> 13;
    const int q = row >> 5, l = row & 31;
    const int n = 16 * q + (l & 15);
    const bool is_lo = l >= 16;
    // forward: A[n=co][k=ci] = W[tap][ci][co];  dgrad: A[n=ci][k=co] = W[26-tap][ci][co]
    const float v = dgrad ? w[((size_t)(26 - tap) * 64 + n) * 64 + k] : w[((size_t)tap * 64 + k) * 64 + n];
    __half h, lo;
    split_f16(v, h, lo);
    const int grp = row >> 3, rr = row & 7;
    const size_t off = (size_t)tap * (128 * 64) + grp * 512 + rr * 64 + (((k >> 3) ^ rr) << 3) + (k & 7);
    img[off] = is_lo ? lo : h;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 5-D map over the packed [2B][Dp][Dp][Dp][64] fp16 planes of an Act (hi planes then lo planes)
bool make_xmap(CUtensorMap* map, const __half* base, int B, int Dp, int ty2, int tz2) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[5] = {64, (cuuint64_t)Dp, (cuuint64_t)Dp, (cuuint64_t)Dp, (cuuint64_t)(2 * B)};
    cuuint64_t strides[4] = {128, (cuuint64_t)128 * Dp, (cuuint64_t)128 * Dp * Dp, (cuuint64_t)128 * Dp * Dp * Dp};
    cuuint32_t box[5] = {64, (cuuint32_t)tz2, (cuuint32_t)ty2, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <int TY>
cudaError_t launch_cfg(const CUtensorMap& map, KParams p, cudaStream_t s) {
    using C = Cfg<TY>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(conv64_tc_kernel<TY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    p.nyt = (p.Do + TY - 1) / TY;
    p.nzt = (p.Do + TZ - 1) / TZ;
    p.ntiles = p.B * p.Do * p.nyt * p.nzt;
    int grid = p.ntiles < num_sms() ? p.ntiles : num_sms();
    conv64_tc_kernel<TY><<<grid, NUM_THREADS, C::SMEM_BYTES, s>>>(map, p);
    return cudaGetLastError();
}

}  // namespace

struct TcWeights {
    int nlayers = 0;
    __half* img = nullptr;   // [nlayers][2 (fwd, dgrad)][27*128*64]
};

bool tc_available() { return true; }

cudaError_t tc_alloc_weights(TcWeights** w, int nlayers) {
    TcWeights* t = new TcWeights();
    t->nlayers = nlayers;
    cudaError_t e = cudaMalloc((void**)&t->img, (size_t)nlayers * 2 * 27 * 128 * 64 * sizeof(__half));
    if (e != cudaSuccess) { delete t; return e; }
    *w = t;
    return cudaSuccess;
}
void tc_free_weights(TcWeights* w) {
    if (!w) return;
    cudaFree(w->img);
    delete w;
}
cudaError_t tc_prepare_weights(TcWeights* w, int layer, const float* kernel, cudaStream_t s) {
    const int n = 27 * 128 * 64;
    __half* base = w->img + (size_t)layer * 2 * n;
    prep_weights_kernel<<<(n + 255) / 256, 256, 0, s>>>(kernel, base, 0);
    prep_weights_kernel<<<(n + 255) / 256, 256, 0, s>>>(kernel, base + n, 1);
    return cudaGetLastError();
}

cudaError_t tc_conv64(TcWeights* w, const TcConvArgs& a, cudaStream_t s) {
    const int Do = a.in.D, B = a.in.B, Dp = a.in.D + 2;
    if (!a.out_raw && a.out.D != Do) return cudaErrorInvalidValue;
    KParams p;
    p.w_img = w->img + ((size_t)a.layer * 2 + (a.dgrad ? 1 : 0)) * 27 * 128 * 64;
    p.out_hi = a.out.hi; p.out_lo = a.out.lo;
    p.res_hi = a.res_hi; p.res_lo = a.res_lo;
    p.bias = a.bias; p.out_raw = a.out_raw; p.absmax = a.absmax;
    p.slope = a.slope; p.B = B; p.Do = Do; p.halo = a.halo;
    p.dbg = nullptr;
    static const bool debug = getenv("SR4D_TC_DEBUG") != nullptr;
    static long long* dbg_buf = nullptr;
    if (debug) {
        if (!dbg_buf) cudaMalloc((void**)&dbg_buf, 148 * 4 * sizeof(long long));
        cudaMemsetAsync(dbg_buf, 0, 148 * 4 * sizeof(long long), s);
        p.dbg = dbg_buf;
    }
    const int ty = Do <= 8 ? 8 : Do <= 16 ? 16 : 24;
    CUtensorMap map;
    if (!make_xmap(&map, a.in.hi, B, Dp, ty + 2, ZP)) return cudaErrorUnknown;
    if (a.in.lo != a.in.hi + act_plane_elems(B, a.in.D)) return cudaErrorInvalidValue;   // planes must be packed
    cudaError_t e;
    switch (ty) {
        case 8: e = launch_cfg<8>(map, p, s); break;
        case 16: e = launch_cfg<16>(map, p, s); break;
        default: e = launch_cfg<24>(map, p, s); break;
    }
    if (debug && e == cudaSuccess) {
        long long h[148 * 4];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, dbg_buf, sizeof h, cudaMemcpyDeviceToHost);
        double a4[4] = {0, 0, 0, 0};
        for (int i = 0; i < 148; ++i) for (int k = 0; k < 4; ++k) a4[k] += (double)h[i * 4 + k] / 148;
        fprintf(stderr, "[tc dbg] Do=%d B=%d tiles=%d: MMA-warp wait cycles avg/CTA: t_empty %.0f  x_full %.0f  w_full %.0f  of total %.0f\n",
                Do, B, B * Do * ((Do + ty - 1) / ty) * ((Do + TZ - 1) / TZ), a4[0], a4[1], a4[2], a4[3]);
    }
    return e;
}
