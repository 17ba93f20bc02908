// Forward kernels that are HBM/L2-bound (everything except the 64->64 3x3x3 convs):
// input features, the two 3->64 stems, the 128->64 1x1 fuse, the trilinear upsample,
// the three 64->1 output heads, and fp32 <-> Act conversions.
#include "kernels.h"

namespace {

// ---- input features: SR4DFlowNet.py:10-15 ------------------------------------------
__global__ void prep_features_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                     const float* __restrict__ w, const float* __restrict__ um,
                                     const float* __restrict__ vm, const float* __restrict__ wm,
                                     float* __restrict__ feat, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = u[i], b = v[i], c = w[i], d = um[i], e = vm[i], f = wm[i];
    // tf.pow(x, 0.5) == sqrt for x >= 0
    float speed = sqrtf(a * a + b * b + c * c);
    float mag = sqrtf(d * d + e * e + f * f);
    float* o = feat + i * 6;
    o[0] = a; o[1] = b; o[2] = c;
    o[3] = mag * speed; o[4] = mag; o[5] = speed;
}

// ---- 3->64 stem conv (+bias, ReLU): SR4DFlowNet.py:17,20 ---------------------------
// 4 threads per voxel, 16 output channels each; weights [27][3][64] in shared memory.
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ feat, int ch0,
                                                        const float* __restrict__ w,
                                                        const float* __restrict__ bias, ActView out) {
    __shared__ float ws[27 * 3 * 64];
    for (int i = threadIdx.x; i < 27 * 3 * 64; i += 256) ws[i] = w[i];
    __syncthreads();
    const int P = out.D;
    const size_t nvox = (size_t)out.B * P * P * P;
    size_t vi = (size_t)blockIdx.x * 64 + (threadIdx.x >> 2);
    if (vi >= nvox) return;
    const int cq = (threadIdx.x & 3) * 16;
    int z = vi % P, y = (vi / P) % P, x = (vi / ((size_t)P * P)) % P, b = vi / ((size_t)P * P * P);
    float acc[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) acc[n] = bias[cq + n];
    for (int dx = -1; dx <= 1; ++dx) {
        int xx = min(max(x + dx, 0), P - 1);
        for (int dy = -1; dy <= 1; ++dy) {
            int yy = min(max(y + dy, 0), P - 1);
            for (int dz = -1; dz <= 1; ++dz) {
                int zz = min(max(z + dz, 0), P - 1);
                const float* f = feat + ((((size_t)b * P + xx) * P + yy) * P + zz) * 6 + ch0;
                float f0 = f[0], f1 = f[1], f2 = f[2];
                const float* wp = ws + (((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1)) * 192 + cq;
#pragma unroll
                for (int n = 0; n < 16; ++n)
                    acc[n] = fmaf(f0, wp[n], fmaf(f1, wp[64 + n], fmaf(f2, wp[128 + n], acc[n])));
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float vv[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) vv[n] = fmaxf(acc[q * 4 + n], 0.f);
        act_store4_halo(out.hi, out.lo, P, b, x, y, z, cq + q * 4, vv, true);
    }
}

// ---- 1x1 conv over concat[a(phase), b(pc)] 128->64 (+bias, ReLU): SR4DFlowNet.py:23-24
__global__ void __launch_bounds__(256) conv1x1_cat_kernel(ActView a, ActView bq, const float* __restrict__ w,
                                                          const float* __restrict__ bias, ActView out) {
    extern __shared__ float ws[];   // [128][64]
    for (int i = threadIdx.x; i < 128 * 64; i += 256) ws[i] = w[i];
    __syncthreads();
    const int D = out.D;
    const size_t nvox = (size_t)out.B * D * D * D;
    size_t vi = (size_t)blockIdx.x * 64 + (threadIdx.x >> 2);
    if (vi >= nvox) return;
    const int cq = (threadIdx.x & 3) * 16;
    int z = vi % D, y = (vi / D) % D, x = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
    size_t off = act_off(D, b, x, y, z);
    float acc[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) acc[n] = bias[cq + n];
    for (int half = 0; half < 2; ++half) {
        const __half* hi = half ? bq.hi : a.hi;
        const __half* lo = half ? bq.lo : a.lo;
        for (int c8 = 0; c8 < 8; ++c8) {
            float xv[8];
            act_load8(hi, lo, off + c8 * 8, xv);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float* wp = ws + (half * 64 + c8 * 8 + k) * 64 + cq;
#pragma unroll
                for (int n = 0; n < 16; ++n) acc[n] = fmaf(xv[k], wp[n], acc[n]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float vv[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) vv[n] = fmaxf(acc[q * 4 + n], 0.f);
        act_store4_halo(out.hi, out.lo, D, b, x, y, z, cq + q * 4, vv, true);
    }
}

// ---- trilinear upsample, align_corners=True, lerp order z,y,x: SR4DFlowNet.py:53-90 ----
// 8 threads per HR voxel, 8 channels each.
__global__ void __launch_bounds__(256) upsample_kernel(ActView in, ActView out, UpsampleTables t) {
    const int H = out.D, D = in.D;
    const size_t nvox = (size_t)out.B * H * H * H;
    size_t vi = (size_t)blockIdx.x * 32 + (threadIdx.x >> 3);
    if (vi >= nvox) return;
    const int c = (threadIdx.x & 7) * 8;
    int z = vi % H, y = (vi / H) % H, x = (vi / ((size_t)H * H)) % H, b = vi / ((size_t)H * H * H);
    const int xl = t.lo[x], xh = t.hi[x], yl = t.lo[y], yh = t.hi[y], zl = t.lo[z], zh = t.hi[z];
    const float fx = t.lerp[x], fy = t.lerp[y], fz = t.lerp[z];
    float r[2][8];
#pragma unroll
    for (int ix = 0; ix < 2; ++ix) {
        float q[2][8];
#pragma unroll
        for (int iy = 0; iy < 2; ++iy) {
            float p0[8], p1[8];
            int xx = ix ? xh : xl, yy = iy ? yh : yl;
            act_load8(in.hi, in.lo, act_off(D, b, xx, yy, zl) + c, p0);
            act_load8(in.hi, in.lo, act_off(D, b, xx, yy, zh) + c, p1);
#pragma unroll
            for (int k = 0; k < 8; ++k) q[iy][k] = p0[k] + (p1[k] - p0[k]) * fz;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) r[ix][k] = q[0][k] + (q[1][k] - q[0][k]) * fy;
    }
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = r[0][k] + (r[1][k] - r[0][k]) * fx;
    act_store4_halo(out.hi, out.lo, H, b, x, y, z, c, o, true);
    act_store4_halo(out.hi, out.lo, H, b, x, y, z, c + 4, o + 4, true);
}

// ---- 64->1 head conv, linear (+bias), writes (B,H^3,3): SR4DFlowNet.py:40,43,46,49 -------
// A CTA owns an 8x8x8 output brick.  Phase 1: every voxel of the 10x10x10 halo brick (the
// replicate halo is already materialised in the Act) is read ONCE and reduced against the 27
// tap vectors (27 dot products of length 64) into shared memory; phase 2: each output voxel
// gathers its 27 partial sums.  HBM/L2 traffic ~2x the input instead of 27x.
struct HeadArgs {
    const __half* hi[3];
    const __half* lo[3];
    const float* w[3];
    const float* b[3];
};
constexpr int HB = 8, HBP = HB + 2, HB_HALO = HBP * HBP * HBP;            // 1000 halo voxels
constexpr int HEAD_SMEM = (HB_HALO * 27 + 27 * 64) * 4;
__global__ void __launch_bounds__(512) head_out_kernel(HeadArgs a, float* __restrict__ out, int B, int H) {
    extern __shared__ float hsm[];
    float* t = hsm;                      // [1000][27]
    float* ws = hsm + HB_HALO * 27;      // [27][64]
    const int nb = (H + HB - 1) / HB;
    int bi = blockIdx.x;
    const int bz = bi % nb; bi /= nb;
    const int by = bi % nb; bi /= nb;
    const int bx = bi % nb;
    const int b = bi / nb;
    const int x0 = bx * HB, y0 = by * HB, z0 = bz * HB;
    const int Hp = H + 2;
    const int tid = threadIdx.x;
    const int ox = tid >> 6, oy = (tid >> 3) & 7, oz = tid & 7;
    float res[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < 3; ++c) {
        __syncthreads();                 // previous head's gather is done with t / ws
        for (int i = tid; i < 27 * 64; i += 512) ws[i] = a.w[c][i];
        __syncthreads();
        const __half* hi = a.hi[c];
        const __half* lo = a.lo[c];
        for (int hv = tid; hv < HB_HALO; hv += 512) {
            const int hx = hv / (HBP * HBP), hy = (hv / HBP) % HBP, hz = hv % HBP;
            const int px = x0 + hx, py = y0 + hy, pz = z0 + hz;      // padded coordinates
            float* tp = t + hv * 27;
            if (px >= Hp || py >= Hp || pz >= Hp) {
                for (int k = 0; k < 27; ++k) tp[k] = 0.f;
                continue;
            }
            const size_t off = ((((size_t)b * Hp + px) * Hp + py) * Hp + pz) * 64;
            float xv[64];
#pragma unroll
            for (int k = 0; k < 8; ++k) act_load8(hi, lo, off + k * 8, xv + k * 8);
#pragma unroll 1
            for (int tap = 0; tap < 27; ++tap) {
                const float4* wp = reinterpret_cast<const float4*>(ws + tap * 64);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float4 wv = wp[k];
                    s0 = fmaf(xv[4 * k], wv.x, s0);
                    s1 = fmaf(xv[4 * k + 1], wv.y, s1);
                    s2 = fmaf(xv[4 * k + 2], wv.z, s2);
                    s3 = fmaf(xv[4 * k + 3], wv.w, s3);
                }
                tp[tap] = (s0 + s1) + (s2 + s3);
            }
        }
        __syncthreads();
        float acc = a.b[c][0];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dz = 0; dz < 3; ++dz)
                    acc += t[(((ox + dx) * HBP + (oy + dy)) * HBP + (oz + dz)) * 27 + (dx * 3 + dy) * 3 + dz];
        res[c] = acc;
    }
    const int x = x0 + ox, y = y0 + oy, z = z0 + oz;
    if (x < H && y < H && z < H) {
        float* o = out + ((((size_t)b * H + x) * H + y) * H + z) * 3;
        o[0] = res[0]; o[1] = res[1]; o[2] = res[2];
    }
}

__global__ void pack_act_kernel(const float* __restrict__ x, ActView out) {
    const int D = out.D;
    const size_t n = (size_t)out.B * D * D * D * 16;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t vi = i >> 4;
    int c = (i & 15) * 4;
    int z = vi % D, y = (vi / D) % D, xx = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
    float4 v = *reinterpret_cast<const float4*>(x + vi * 64 + c);
    float vv[4] = {v.x, v.y, v.z, v.w};
    act_store4_halo(out.hi, out.lo, D, b, xx, y, z, c, vv, true);
}
__global__ void unpack_act_kernel(ActView in, float* __restrict__ yv) {
    const int D = in.D;
    const size_t n = (size_t)in.B * D * D * D * 16;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t vi = i >> 4;
    int c = (i & 15) * 4;
    int z = vi % D, y = (vi / D) % D, xx = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
    float vv[4];
    act_load4(in.hi, in.lo, act_off(D, b, xx, y, z) + c, vv);
    *reinterpret_cast<float4*>(yv + vi * 64 + c) = make_float4(vv[0], vv[1], vv[2], vv[3]);
}
}  // namespace

cudaError_t launch_prep_features(const float* u, const float* v, const float* w, const float* um,
                                 const float* vm, const float* wm, float* feat, int B, int P, cudaStream_t s) {
    size_t n = (size_t)B * P * P * P;
    prep_features_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(u, v, w, um, vm, wm, feat, n);
    return cudaGetLastError();
}
cudaError_t launch_stem_conv(const float* feat, int ch0, const float* w, const float* bias, ActView out,
                             cudaStream_t s) {
    size_t nvox = (size_t)out.B * out.D * out.D * out.D;
    stem_conv_kernel<<<(unsigned)((nvox + 63) / 64), 256, 0, s>>>(feat, ch0, w, bias, out);
    return cudaGetLastError();
}
cudaError_t launch_conv1x1_cat(ActView a, ActView b, const float* w, const float* bias, ActView out, cudaStream_t s) {
    size_t nvox = (size_t)out.B * out.D * out.D * out.D;
    conv1x1_cat_kernel<<<(unsigned)((nvox + 63) / 64), 256, 128 * 64 * 4, s>>>(a, b, w, bias, out);
    return cudaGetLastError();
}
cudaError_t launch_upsample(ActView in, ActView out, int r, UpsampleTables t, cudaStream_t s) {
    (void)r;
    size_t nvox = (size_t)out.B * out.D * out.D * out.D;
    upsample_kernel<<<(unsigned)((nvox + 31) / 32), 256, 0, s>>>(in, out, t);
    return cudaGetLastError();
}
cudaError_t launch_head_out(ActView h0, ActView h1, ActView h2, const float* w0, const float* w1,
                            const float* w2, const float* b0, const float* b1, const float* b2, float* out,
                            cudaStream_t s) {
    HeadArgs a;
    a.hi[0] = h0.hi; a.hi[1] = h1.hi; a.hi[2] = h2.hi;
    a.lo[0] = h0.lo; a.lo[1] = h1.lo; a.lo[2] = h2.lo;
    a.w[0] = w0; a.w[1] = w1; a.w[2] = w2;
    a.b[0] = b0; a.b[1] = b1; a.b[2] = b2;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(head_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    const int nb = (h0.D + HB - 1) / HB;
    head_out_kernel<<<(unsigned)(h0.B * nb * nb * nb), 512, HEAD_SMEM, s>>>(a, out, h0.B, h0.D);
    return cudaGetLastError();
}
cudaError_t launch_pack_act(const float* x, ActView out, cudaStream_t s) {
    size_t n = (size_t)out.B * out.D * out.D * out.D * 16;
    pack_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, out);
    return cudaGetLastError();
}
cudaError_t launch_unpack_act(ActView in, float* y, cudaStream_t s) {
    size_t n = (size_t)in.B * in.D * in.D * in.D * 16;
    unpack_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, y);
    return cudaGetLastError();
}
