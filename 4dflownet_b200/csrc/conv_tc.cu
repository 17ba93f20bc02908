// tcgen05 (5th-gen tensor core) 64->64 3x3x3 convolution for sm_100a, split-fp16 "fp32-accurate".
//
// Reference op: conv3d() of Network/SR4DFlowNet.py:93-108 / resnet_block :111-120 (30 of the
// 36 layers, 99.5 % of the FLOPs).
//
// Formulation (implicit GEMM, transposed so the voxel dimension is the flexible MMA N):
//     D[128 x N] += A_tap[128 x 64] * B_tap[N x 64]^T        for the 27 taps
//   A_tap rows  = the layer's weights for that tap, split in fp16: W = Whi + Wlo/2048.
//                 Row 32q+l holds Whi[co=16q+l] for l<16 and Wlo[co=16q+l-16] for l>=16, so
//                 the hi and lo partial sums of one output channel sit in lanes l and l^16 of
//                 the same epilogue warp (one shuffle combines them).
//   B_tap rows  = N = TY*TZ voxels of one (y,z) tile at a fixed x: their 64 input channels,
//                 taken from the activation's fp16 hi plane (accumulator D1) and lo plane (D2).
//   out[co][v]  = D1[hi] + (D1[lo] + D2[hi]) / 2048 + D2[lo] / 2048^2     (fp32 in TMEM)
//
// Data movement: one TMA box per (dx,dz) loads the (TY+2) x TZ x 64ch plane (hi and lo) into
// shared memory in the canonical K-major SWIZZLE_128B layout; the three dy taps are then
// 1024B-aligned row offsets into that plane, so every activation byte is fetched 9x (not 27x)
// from L2.  Weights stream per tap as pre-swizzled 16 KB images (cp.async.bulk).
// Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
// (TMEM -> registers -> bias / residual / activation / fp16 split -> global, replicate halo).
#include <cuda.h>

#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "conv_tc.h"
#include "tc_ptx.cuh"

namespace {

constexpr int W_TAP_BYTES = 128 * 64 * 2;   // 16 KB: [Whi;Wlo] x 64 ci, fp16, swizzled
constexpr int NUM_THREADS = 192;

template <int TY, int TZ>
struct Cfg {
    static constexpr int N = TY * TZ;                       // voxels per tile = MMA N
    static constexpr int ROWS = (TY + 2) * TZ;              // rows of one staged plane
    static constexpr int PLANE_BYTES = ROWS * 128;          // one of hi / lo
    static constexpr int XSTAGE_BYTES = 2 * PLANE_BYTES;
    static constexpr int NXS = 2;
    static constexpr int NWS = (200 * 1024 - NXS * XSTAGE_BYTES) / W_TAP_BYTES >= 6
                                   ? 6 : (200 * 1024 - NXS * XSTAGE_BYTES) / W_TAP_BYTES;
    static constexpr int SMEM_BYTES = 1024 + NXS * XSTAGE_BYTES + NWS * W_TAP_BYTES + 256;
    static constexpr int TMEM_COLS = 2 * N <= 32 ? 32 : 2 * N <= 64 ? 64 : 2 * N <= 128 ? 128 : 2 * N <= 256 ? 256 : 512;
    static_assert(TZ % 8 == 0, "plane rows must stay 1024B aligned under dy shifts");
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N constraint for M=128");
    static_assert(NWS >= 3, "need at least one plane worth of weight taps in flight");
    static_assert(TY + 2 <= 256 && TZ <= 256, "TMA box limits");
};

struct KParams {
    const __half* w_img;     // [27][128][64] fp16 pre-swizzled, tap order = (dx*3+dz)*3+dy
    __half* out_hi;
    __half* out_lo;
    const __half* res_hi;
    const __half* res_lo;
    const float* bias;
    float* out_raw;          // optional fp32 [B][Do^3][64] output instead of Act
    float slope;
    int B, Do, halo;
    int nyt, nzt, ntiles;
};

template <int TY, int TZ>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv64_tc_kernel(const __grid_constant__ CUtensorMap xmap, KParams p) {
    using C = Cfg<TY, TZ>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* xs = smem;                                   // NXS x [hi plane | lo plane]
    uint8_t* wsm = smem + C::NXS * C::XSTAGE_BYTES;       // NWS x 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + C::NWS * W_TAP_BYTES);
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
        mbar_init(t_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
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

    const int tiles_per_b = p.Do * p.nyt * p.nzt;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t xi = 0, wi = 0;   // running stage counters
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                const int b = t / tiles_per_b;
                int rem = t % tiles_per_b;
                const int x = rem / (p.nyt * p.nzt);
                rem %= p.nyt * p.nzt;
                const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
                for (int pl = 0; pl < 9; ++pl) {
                    const int dx = pl / 3, dz = pl % 3;
                    const uint32_t s = xi % C::NXS, ph = (xi / C::NXS) & 1;
                    mbar_wait(&x_empty[s], ph ^ 1);
                    mbar_expect_tx(&x_full[s], C::XSTAGE_BYTES);
                    uint8_t* dst = xs + s * C::XSTAGE_BYTES;
                    tma_load_5d(dst, &xmap, &x_full[s], 0, z0 + dz, y0, x + dx, b);
                    tma_load_5d(dst + C::PLANE_BYTES, &xmap, &x_full[s], 0, z0 + dz, y0, x + dx, p.B + b);
                    ++xi;
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t ws = wi % C::NWS, wph = (wi / C::NWS) & 1;
                        mbar_wait(&w_empty[ws], wph ^ 1);
                        mbar_expect_tx(&w_full[ws], W_TAP_BYTES);
                        bulk_load(wsm + ws * W_TAP_BYTES,
                                  reinterpret_cast<const uint8_t*>(p.w_img) + (size_t)(pl * 3 + dy) * W_TAP_BYTES,
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
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
                mbar_wait(t_empty, (ti & 1) ^ 1);
                tc_fence_after();
                for (int pl = 0; pl < 9; ++pl) {
                    const uint32_t s = xi % C::NXS, ph = (xi / C::NXS) & 1;
                    mbar_wait(&x_full[s], ph);
                    tc_fence_after();
                    const uint32_t xhi = smem_u32(xs + s * C::XSTAGE_BYTES);
                    const uint32_t xlo = xhi + C::PLANE_BYTES;
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t ws = wi % C::NWS, wph = (wi / C::NWS) & 1;
                        mbar_wait(&w_full[ws], wph);
                        tc_fence_after();
                        const uint32_t wa = smem_u32(wsm + ws * W_TAP_BYTES);
                        const uint32_t boff = dy * TZ * 128;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ad = make_desc(wa + k * 32);
                            const uint32_t acc = (pl | dy | k) != 0;
                            tc_mma_f16(d1, ad, make_desc(xhi + boff + k * 32), idesc, acc);
                            tc_mma_f16(d2, ad, make_desc(xlo + boff + k * 32), idesc, acc);
                        }
                        tc_commit(&w_empty[ws]);
                        ++wi;
                    }
                    tc_commit(&x_empty[s]);
                    ++xi;
                }
                tc_commit(t_full);
            }
        }
    } else {
        // ================= epilogue (warps 2..5) =================
        const int e = warp & 3;                      // TMEM lane quarter this warp may access
        const int co = 16 * e + (lane & 15);
        const bool is_lo = lane >= 16;
        const float bias = p.bias ? p.bias[co] : 0.f;
        const int Do = p.Do;
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
            const int b = t / tiles_per_b;
            int rem = t % tiles_per_b;
            const int x = rem / (p.nyt * p.nzt);
            rem %= p.nyt * p.nzt;
            const int y0 = (rem / p.nzt) * TY, z0 = (rem % p.nzt) * TZ;
            mbar_wait(t_full, ti & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(32 * e) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < C::N; c0 += 16) {
                float a[16], d[16];
                tc_ld16(trow + c0, a);
                tc_ld16(trow + C::N + c0, d);
                tc_ld_wait();
                const float s1 = is_lo ? SR4D_LO_INV : 1.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float v = fmaf(d[j], SR4D_LO_INV, a[j]) * s1;
                    a[j] = v + __shfl_xor_sync(0xffffffffu, v, 16);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v0 = is_lo ? a[8 + j] : a[j];
                    const int n = c0 + (is_lo ? 8 : 0) + j;
                    const int y = y0 + n / TZ, z = z0 + n % TZ;
                    if (y >= Do || z >= Do) continue;
                    float v = v0 + bias;
                    if (p.out_raw) {
                        p.out_raw[((((size_t)b * Do + x) * Do + y) * Do + z) * 64 + co] = v;
                        continue;
                    }
                    const size_t o = act_off(Do, b, x, y, z) + co;
                    if (p.res_hi) v += join_f16(p.res_hi[o], p.res_lo[o]);
                    v = act_fn(v, p.slope);
                    __half h, l;
                    split_f16(v, h, l);
                    if (!p.halo) {
                        p.out_hi[o] = h;
                        p.out_lo[o] = l;
                    } else {
                        for (int ddx = -1; ddx <= 1; ++ddx) {
                            if ((ddx == -1 && x != 0) || (ddx == 1 && x != Do - 1)) continue;
                            for (int ddy = -1; ddy <= 1; ++ddy) {
                                if ((ddy == -1 && y != 0) || (ddy == 1 && y != Do - 1)) continue;
                                for (int ddz = -1; ddz <= 1; ++ddz) {
                                    if ((ddz == -1 && z != 0) || (ddz == 1 && z != Do - 1)) continue;
                                    const size_t oo = act_off(Do, b, x + ddx, y + ddy, z + ddz) + co;
                                    p.out_hi[oo] = h;
                                    p.out_lo[oo] = l;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty);
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
// weight image: fp32 Keras [27][ci][co] -> fp16 split, row-permuted, swizzled, tap-reordered
// ------------------------------------------------------------------------------------------
__global__ void prep_weights_kernel(const float* __restrict__ w, __half* __restrict__ img, int dgrad) {
    // one thread per (tap_img, row, k)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 27 * 128 * 64) return;
    const int k = i & 63, row = (i >> 6) & 127, ti = i >> 13;
    const int pl = ti / 3, dy = ti % 3, dx = pl / 3, dz = pl % 3;
    const int tap = (dx * 3 + dy) * 3 + dz;
    const int q = row >> 5, l = row & 31;
    const int n = 16 * q + (l & 15);
    const bool is_lo = l >= 16;
    // forward: A[n=co][k=ci] = W[tap][ci][co];  dgrad: A[n=ci][k=co] = W[26-tap][ci][co]
    const float v = dgrad ? w[((size_t)(26 - tap) * 64 + n) * 64 + k] : w[((size_t)tap * 64 + k) * 64 + n];
    __half h, lo;
    split_f16(v, h, lo);
    const int grp = row >> 3, rr = row & 7;
    const size_t off = (size_t)ti * (128 * 64) + grp * 512 + rr * 64 + (((k >> 3) ^ rr) << 3) + (k & 7);
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
bool make_xmap(CUtensorMap* map, const __half* base, int B, int Dp, int ty2, int tz) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[5] = {64, (cuuint64_t)Dp, (cuuint64_t)Dp, (cuuint64_t)Dp, (cuuint64_t)(2 * B)};
    cuuint64_t strides[4] = {128, (cuuint64_t)128 * Dp, (cuuint64_t)128 * Dp * Dp, (cuuint64_t)128 * Dp * Dp * Dp};
    cuuint32_t box[5] = {64, (cuuint32_t)tz, (cuuint32_t)ty2, 1, 1};
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

template <int TY, int TZ>
cudaError_t launch_cfg(const CUtensorMap& map, KParams p, cudaStream_t s) {
    using C = Cfg<TY, TZ>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(conv64_tc_kernel<TY, TZ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    p.nyt = (p.Do + TY - 1) / TY;
    p.nzt = (p.Do + TZ - 1) / TZ;
    p.ntiles = p.B * p.Do * p.nyt * p.nzt;
    int grid = p.ntiles < num_sms() ? p.ntiles : num_sms();
    conv64_tc_kernel<TY, TZ><<<grid, NUM_THREADS, C::SMEM_BYTES, s>>>(map, p);
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
    const int Do = a.out.D, B = a.in.B, Dp = a.in.D + 2;
    if (a.in.D != Do) return cudaErrorInvalidValue;
    KParams p;
    p.w_img = w->img + ((size_t)a.layer * 2 + (a.dgrad ? 1 : 0)) * 27 * 128 * 64;
    p.out_hi = a.out.hi; p.out_lo = a.out.lo;
    p.res_hi = a.res_hi; p.res_lo = a.res_lo;
    p.bias = a.bias; p.out_raw = nullptr;
    p.slope = a.slope; p.B = B; p.Do = Do; p.halo = a.halo;
    // tile shape by grid edge: full-z lines, N = TY*TZ <= 256
    int ty, tz;
    if (Do <= 8) { ty = 8; tz = 8; }
    else if (Do <= 16) { ty = 8; tz = 16; }
    else if (Do <= 24) { ty = 8; tz = 24; }
    else if (Do <= 32) { ty = 6; tz = 32; }
    else { ty = 4; tz = 48; }            // larger grids are tiled along z as well
    CUtensorMap map;
    if (!make_xmap(&map, a.in.hi, B, Dp, ty + 2, tz)) return cudaErrorUnknown;
    if (a.in.lo != a.in.hi + act_plane_elems(B, a.in.D)) return cudaErrorInvalidValue;   // planes must be packed
    switch (tz) {
        case 8: return launch_cfg<8, 8>(map, p, s);
        case 16: return launch_cfg<8, 16>(map, p, s);
        case 24: return launch_cfg<8, 24>(map, p, s);
        case 32: return launch_cfg<6, 32>(map, p, s);
        default: return launch_cfg<4, 48>(map, p, s);
    }
}
