// fp32 CUDA-core 64->64 3x3x3 convolution: the correctness anchor for the tensor-core
// kernel and the fallback-free path for shapes it does not cover.
//
// Reference op: conv3d() of Network/SR4DFlowNet.py:93-108 (clamp pad + VALID Conv3D
// [+bias][+ReLU]) and the resnet_block epilogue of :111-120; with dgrad=1 the same
// kernel evaluates Conv3DBackpropInput on the padded grid.
//
// Tiling: a CTA computes a 2x8x8 voxel brick x 64 output channels; the 4x10x10 input
// brick and the 27x8x64 weight slab of one 8-channel K chunk are staged in shared
// memory; each of the 256 threads owns 4 z-consecutive voxels x 8 output channels.
#include "kernels.h"
#include "tc_host.h"

namespace {
constexpr int TX = 2, TY = 8, TZ = 8, CK = 8;
constexpr int BX = TX + 2, BY = TY + 2, BZ = TZ + 2;
constexpr int BRICK = BX * BY * BZ;          // 400
constexpr int XS_STRIDE = BRICK + 8;         // 408: keeps rows 32B aligned, breaks bank period
constexpr int WS_ELEMS = 27 * CK * 64;
constexpr int SMEM_BYTES = (CK * XS_STRIDE + WS_ELEMS) * 4;

template <bool IN_ACT, bool OUT_ACT>
__global__ void __launch_bounds__(256, 2) conv64_simt_kernel(Conv64Args a) {
    extern __shared__ float smem[];
    float* xs = smem;
    float* ws = smem + CK * XS_STRIDE;

    const int tid = threadIdx.x;
    const int cg = tid & 7;
    const int vg = tid >> 3;
    const int vz0 = (vg & 1) * 4, vy = (vg >> 1) & 7, vx = vg >> 4;

    const int Do = a.Do, Din = Do + 2;
    const int nxt = (Do + TX - 1) / TX;
    const int z0 = blockIdx.x * TZ, y0 = blockIdx.y * TY;
    const int b = blockIdx.z / nxt, x0 = (blockIdx.z % nxt) * TX;

    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[j][n] = 0.f;

    const size_t in_b = (size_t)b * Din * Din * Din;

    for (int c = 0; c < 64 / CK; ++c) {
        __syncthreads();
        for (int i = tid; i < BRICK; i += 256) {
            int bx = i / (BY * BZ), by = (i / BZ) % BY, bz = i % BZ;
            int gx = x0 + bx, gy = y0 + by, gz = z0 + bz;
            float v[8];
            if (gx < Din && gy < Din && gz < Din) {
                size_t off = (in_b + ((size_t)gx * Din + gy) * Din + gz) * 64 + c * CK;
                if (IN_ACT) {
                    act_load8(a.in_hi, a.in_lo, off, v);
                } else {
                    float4 p = *reinterpret_cast<const float4*>(a.in_f32 + off);
                    float4 q = *reinterpret_cast<const float4*>(a.in_f32 + off + 4);
                    v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w;
                    v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) xs[k * XS_STRIDE + i] = v[k];
        }
        if (!a.dgrad) {
            for (int i = tid; i < WS_ELEMS / 4; i += 256) {
                int tap = i / (CK * 16), k = (i / 16) % CK, n4 = i % 16;
                float4 wv = *reinterpret_cast<const float4*>(a.w + ((size_t)tap * 64 + c * CK + k) * 64 + n4 * 4);
                *reinterpret_cast<float4*>(ws + (tap * CK + k) * 64 + n4 * 4) = wv;
            }
        } else {
            for (int i = tid; i < WS_ELEMS; i += 256) {
                int tap = i / (CK * 64), k = (i / 64) % CK, n = i % 64;
                ws[(tap * CK + k) * 64 + n] = a.w[((size_t)(26 - tap) * 64 + n) * 64 + c * CK + k];
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < CK; ++k) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const float* xp = xs + k * XS_STRIDE + ((vx + dx) * BY + (vy + dy)) * BZ + vz0;
                    float xv[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) xv[j] = xp[j];
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz) {
                        const float* wp = ws + (((dx * 3 + dy) * 3 + dz) * CK + k) * 64;
                        float4 w0 = *reinterpret_cast<const float4*>(wp + cg * 4);
                        float4 w1 = *reinterpret_cast<const float4*>(wp + 32 + cg * 4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x = xv[j + dz];
                            acc[j][0] = fmaf(x, w0.x, acc[j][0]);
                            acc[j][1] = fmaf(x, w0.y, acc[j][1]);
                            acc[j][2] = fmaf(x, w0.z, acc[j][2]);
                            acc[j][3] = fmaf(x, w0.w, acc[j][3]);
                            acc[j][4] = fmaf(x, w1.x, acc[j][4]);
                            acc[j][5] = fmaf(x, w1.y, acc[j][5]);
                            acc[j][6] = fmaf(x, w1.z, acc[j][6]);
                            acc[j][7] = fmaf(x, w1.w, acc[j][7]);
                        }
                    }
                }
            }
        }
    }

    const int x = x0 + vx, y = y0 + vy;
    if (x >= Do || y >= Do) return;
    const int c0 = cg * 4, c1 = 32 + cg * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int z = z0 + vz0 + j;
        if (z >= Do) continue;
        if (OUT_ACT) {
            float v0[4], v1[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) { v0[n] = acc[j][n]; v1[n] = acc[j][4 + n]; }
            if (a.bias) {
#pragma unroll
                for (int n = 0; n < 4; ++n) { v0[n] += a.bias[c0 + n]; v1[n] += a.bias[c1 + n]; }
            }
            if (a.res_hi) {
                size_t o = act_off(Do, b, x, y, z);
                float r0[4], r1[4];
                act_load4(a.res_hi, a.res_lo, o + c0, r0);
                act_load4(a.res_hi, a.res_lo, o + c1, r1);
#pragma unroll
                for (int n = 0; n < 4; ++n) { v0[n] += r0[n]; v1[n] += r1[n]; }
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) { v0[n] = act_fn(v0[n], a.slope); v1[n] = act_fn(v1[n], a.slope); }
            act_store4_halo(a.out_hi, a.out_lo, Do, b, x, y, z, c0, v0, a.halo, a.ovf);
            act_store4_halo(a.out_hi, a.out_lo, Do, b, x, y, z, c1, v1, a.halo, a.ovf);
        } else {
            size_t o = ((((size_t)b * Do + x) * Do + y) * Do + z) * 64;
            *reinterpret_cast<float4*>(a.out_raw + o + c0) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            *reinterpret_cast<float4*>(a.out_raw + o + c1) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
        }
    }
}

template <bool IN_ACT, bool OUT_ACT>
cudaError_t launch_t(const Conv64Args& a, cudaStream_t s) {
    cudaError_t ea = tc_func_smem(reinterpret_cast<const void*>(conv64_simt_kernel<IN_ACT, OUT_ACT>), SMEM_BYTES);
    if (ea != cudaSuccess) return ea;
    dim3 grid((a.Do + TZ - 1) / TZ, (a.Do + TY - 1) / TY, ((a.Do + TX - 1) / TX) * a.B);
    conv64_simt_kernel<IN_ACT, OUT_ACT><<<grid, 256, SMEM_BYTES, s>>>(a);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_conv64_simt(const Conv64Args& a, cudaStream_t s) {
    const bool in_act = a.in_hi != nullptr;
    const bool out_act = a.out_hi != nullptr;
    if (in_act && out_act) return launch_t<true, true>(a, s);
    if (in_act && !out_act) return launch_t<true, false>(a, s);
    if (!in_act && out_act) return launch_t<false, true>(a, s);
    return launch_t<false, false>(a, s);
}
