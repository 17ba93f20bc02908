// Forward kernels that are HBM/L2-bound (everything except the 64->64 3x3x3 convs):
// input features, the two 3->64 stems, the 128->64 1x1 fuse, the trilinear upsample,
// the three 64->1 output heads, and fp32 <-> Act conversions.
#include "kernels.h"
#include "tc_host.h"

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
// 4 threads per run of STEM_VZ consecutive z voxels, 16 output channels each; weights [27][3][64] in shared memory.
// The kernel is bound by the shared-memory weight fetches (a 16-byte load per 4 FMAs when a thread owns one voxel:
// 12 LDS.128 per 48 FMAs, 4 LSU cycles each); with four voxels per thread every fetched weight feeds 16 FMAs and the
// six clamped feature columns of a (dx,dy) row are read once for the three dz taps of all four voxels.
// Small grids (batch 1: 54 CTAs for 148 SMs with four voxels per thread) take the one-voxel instantiation.
struct StemArgs {                 // the two stems (pc: feature channels 3..5, phase: 0..2) run as grid.y = 0 / 1 of one launch
    const float* w[2];
    const float* bias[2];
    ActView out[2];
    int ch0[2];
};
// CPT output channels per thread (64 / CPT threads per run): 16 on large grids, 8 on small ones (batch 1: twice the CTAs
// and half the dependent FMA chain per thread -- the kernel is latency-bound there)
template <int STEM_VZ, int CPT>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ feat, const StemArgs a) {
    __shared__ __align__(16) float ws[27 * 3 * 64];
    const int which = blockIdx.y;
    const float* __restrict__ w = a.w[which];
    const float* __restrict__ bias = a.bias[which];
    const ActView out = a.out[which];
    const int ch0 = a.ch0[which];
    for (int i = threadIdx.x; i < 27 * 3 * 64; i += 256) ws[i] = w[i];
    __syncthreads();
    constexpr int TPR = 64 / CPT;                                   // threads per run
    constexpr int RPC = 256 / TPR;                                  // runs per CTA pass
    const int P = out.D;
    const int nzr = (P + STEM_VZ - 1) / STEM_VZ;                    // z runs per line
    const size_t nrun = (size_t)out.B * P * P * nzr;
    const int cq = (threadIdx.x % TPR) * CPT;
    // grid-stride over groups of RPC runs: the weights are staged once per CTA
    for (size_t ri = (size_t)blockIdx.x * RPC + (threadIdx.x / TPR); ri < nrun; ri += (size_t)gridDim.x * RPC) {
        const int z0 = (int)(ri % nzr) * STEM_VZ, y = (int)((ri / nzr) % P), x = (int)((ri / ((size_t)nzr * P)) % P);
        const int b = (int)(ri / ((size_t)nzr * P * P));
        float acc[STEM_VZ][CPT];
#pragma unroll
        for (int v = 0; v < STEM_VZ; ++v)
#pragma unroll
            for (int n = 0; n < CPT; ++n) acc[v][n] = bias[cq + n];
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = min(max(x + dx, 0), P - 1);
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = min(max(y + dy, 0), P - 1);
                const float* row = feat + (((size_t)b * P + xx) * P + yy) * P * 6 + ch0;
                float f[STEM_VZ + 2][3];
#pragma unroll
                for (int j = 0; j < STEM_VZ + 2; ++j) {
                    const float* fp = row + (size_t)min(max(z0 + j - 1, 0), P - 1) * 6;
                    f[j][0] = fp[0]; f[j][1] = fp[1]; f[j][2] = fp[2];
                }
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
                    const float* wp = ws + (((dx + 1) * 3 + (dy + 1)) * 3 + dz) * 192 + cq;
#pragma unroll
                    for (int n4 = 0; n4 < CPT / 4; ++n4) {
                        const float4 w0 = *reinterpret_cast<const float4*>(wp + n4 * 4);
                        const float4 w1 = *reinterpret_cast<const float4*>(wp + 64 + n4 * 4);
                        const float4 w2 = *reinterpret_cast<const float4*>(wp + 128 + n4 * 4);
#pragma unroll
                        for (int v = 0; v < STEM_VZ; ++v) {
                            const float f0 = f[v + dz][0], f1 = f[v + dz][1], f2 = f[v + dz][2];
                            acc[v][n4 * 4 + 0] = fmaf(f0, w0.x, fmaf(f1, w1.x, fmaf(f2, w2.x, acc[v][n4 * 4 + 0])));
                            acc[v][n4 * 4 + 1] = fmaf(f0, w0.y, fmaf(f1, w1.y, fmaf(f2, w2.y, acc[v][n4 * 4 + 1])));
                            acc[v][n4 * 4 + 2] = fmaf(f0, w0.z, fmaf(f1, w1.z, fmaf(f2, w2.z, acc[v][n4 * 4 + 2])));
                            acc[v][n4 * 4 + 3] = fmaf(f0, w0.w, fmaf(f1, w1.w, fmaf(f2, w2.w, acc[v][n4 * 4 + 3])));
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < STEM_VZ; ++v) {
            if (z0 + v >= P) break;
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q) {
                float vv[4];
#pragma unroll
                for (int n = 0; n < 4; ++n) vv[n] = fmaxf(acc[v][q * 4 + n], 0.f);
                act_store4_halo(out.hi, out.lo, P, b, x, y, z0 + v, cq + q * 4, vv, true, out.ovf);
            }
        }
    }
}

// ---- 1x1 conv over concat[a(phase), b(pc)] 128->64 (+bias, ReLU): SR4DFlowNet.py:23-24
// 4 threads per group of C1_NV consecutive voxels, 16 output channels each: every 16-byte weight fetch from shared
// memory feeds 4 * C1_NV FMAs (the one-voxel version was bound by those fetches).
template <int C1_NV, int CPT>
__global__ void __launch_bounds__(256) conv1x1_cat_kernel(ActView a, ActView bq, const float* __restrict__ w,
                                                          const float* __restrict__ bias, ActView out) {
    extern __shared__ __align__(16) float ws[];   // [128][64]
    for (int i = threadIdx.x; i < 128 * 64; i += 256) ws[i] = w[i];
    __syncthreads();
    const int D = out.D;
    const size_t nvox = (size_t)out.B * D * D * D;
    const size_t ngrp = (nvox + C1_NV - 1) / C1_NV;
    constexpr int TPG = 64 / CPT;                  // threads per voxel group (CPT = 16 or 8 output channels each)
    constexpr int GPC = 256 / TPG;                 // groups per CTA pass
    const int cq = (threadIdx.x % TPG) * CPT;
    for (size_t gi = (size_t)blockIdx.x * GPC + (threadIdx.x / TPG); gi < ngrp; gi += (size_t)gridDim.x * GPC) {
        size_t off[C1_NV];
        int vx[C1_NV], vy[C1_NV], vz[C1_NV], vb[C1_NV];
#pragma unroll
        for (int v = 0; v < C1_NV; ++v) {
            size_t vi = gi * C1_NV + v;
            if (vi >= nvox) vi = nvox - 1;                       // tail: recompute the last voxel, store skipped below
            vz[v] = (int)(vi % D); vy[v] = (int)((vi / D) % D); vx[v] = (int)((vi / ((size_t)D * D)) % D);
            vb[v] = (int)(vi / ((size_t)D * D * D));
            off[v] = act_off(D, vb[v], vx[v], vy[v], vz[v]);
        }
        float acc[C1_NV][CPT];
#pragma unroll
        for (int v = 0; v < C1_NV; ++v)
#pragma unroll
            for (int n = 0; n < CPT; ++n) acc[v][n] = bias[cq + n];
        for (int half = 0; half < 2; ++half) {
            const __half* hi = half ? bq.hi : a.hi;
            const __half* lo = half ? bq.lo : a.lo;
            for (int c8 = 0; c8 < 8; ++c8) {
                float xv[C1_NV][8];
#pragma unroll
                for (int v = 0; v < C1_NV; ++v) act_load8(hi, lo, off[v] + c8 * 8, xv[v]);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4* wp = reinterpret_cast<const float4*>(ws + (half * 64 + c8 * 8 + k) * 64 + cq);
#pragma unroll
                    for (int n4 = 0; n4 < CPT / 4; ++n4) {
                        const float4 wv = wp[n4];
#pragma unroll
                        for (int v = 0; v < C1_NV; ++v) {
                            acc[v][n4 * 4 + 0] = fmaf(xv[v][k], wv.x, acc[v][n4 * 4 + 0]);
                            acc[v][n4 * 4 + 1] = fmaf(xv[v][k], wv.y, acc[v][n4 * 4 + 1]);
                            acc[v][n4 * 4 + 2] = fmaf(xv[v][k], wv.z, acc[v][n4 * 4 + 2]);
                            acc[v][n4 * 4 + 3] = fmaf(xv[v][k], wv.w, acc[v][n4 * 4 + 3]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < C1_NV; ++v) {
            if (gi * C1_NV + v >= nvox) break;
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q) {
                float vv[4];
#pragma unroll
                for (int n = 0; n < 4; ++n) vv[n] = fmaxf(acc[v][q * 4 + n], 0.f);
                act_store4_halo(out.hi, out.lo, D, vb[v], vx[v], vy[v], vz[v], cq + q * 4, vv, true, out.ovf);
            }
        }
    }
}

// ---- trilinear upsample, align_corners=True, lerp order z,y,x: SR4DFlowNet.py:53-90 ----
// 8 threads per HR z-line (b, x, y), 8 channels each, marching along z.  The four (x,y) corner columns of a line are
// fixed, and the LR z index pair (lo, hi) advances by at most one per output voxel (scale <= 1), so the thread keeps the
// corner values at LR indices i0 and i0+1 in registers and loads every LR voxel of its four columns exactly once:
// 2 instead of 8 corner loads per output voxel, same lerp expressions and order as the one-voxel-per-thread version.
// blockIdx.y = z segment (small grids: a batch-1 volume has only 2304 lines, 124 threads per SM; the segments start with
// their own corner loads).
// (SEG = false is the whole-line kernel as it was: the segmented bounds cost the batch-8 launch 60 % when they were runtime values)
template <bool SEG>
__global__ void __launch_bounds__(256) upsample_kernel(ActView in, ActView out, UpsampleTables t, int zseg) {
    const int H = out.D, D = in.D;
    const size_t nline = (size_t)out.B * H * H;
    const size_t li = (size_t)blockIdx.x * 32 + (threadIdx.x >> 3);
    if (li >= nline) return;
    const int zbeg = SEG ? blockIdx.y * zseg : 0, zend = SEG ? min(H, zbeg + zseg) : H;
    const int c = (threadIdx.x & 7) * 8;
    const int y = (int)(li % H), x = (int)((li / H) % H), b = (int)(li / ((size_t)H * H));
    const int xl = t.lo[x], xh = t.hi[x], yl = t.lo[y], yh = t.hi[y];
    const float fx = t.lerp[x], fy = t.lerp[y];
    size_t col[4];                                   // [ix][iy] column bases at z = 0
    col[0] = act_off(D, b, xl, yl, 0) + c; col[1] = act_off(D, b, xl, yh, 0) + c;
    col[2] = act_off(D, b, xh, yl, 0) + c; col[3] = act_off(D, b, xh, yh, 0) + c;
    float v0[4][8], v1[4][8];
    int i0 = SEG ? t.lo[zbeg] : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        act_load8(in.hi, in.lo, col[k] + (size_t)i0 * SR4D_C, v0[k]);
        act_load8(in.hi, in.lo, col[k] + (size_t)min(i0 + 1, D - 1) * SR4D_C, v1[k]);
    }
    for (int z = zbeg; z < zend; ++z) {
        const int zl = t.lo[z], zh = t.hi[z];
        const float fz = t.lerp[z];
        if (zl > i0) {                               // warp-uniform: every thread of the block walks the same z
            ++i0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v0[k][j] = v1[k][j];
                act_load8(in.hi, in.lo, col[k] + (size_t)min(i0 + 1, D - 1) * SR4D_C, v1[k]);
            }
        }
        const bool same = zh == zl;
        float r[2][8];
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
            float q[2][8];
#pragma unroll
            for (int iy = 0; iy < 2; ++iy) {
                const int k = ix * 2 + iy;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float p0 = v0[k][j], p1 = same ? v0[k][j] : v1[k][j];
                    q[iy][j] = p0 + (p1 - p0) * fz;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) r[ix][j] = q[0][j] + (q[1][j] - q[0][j]) * fy;
        }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = r[0][j] + (r[1][j] - r[0][j]) * fx;
        act_store8_halo(out.hi, out.lo, H, b, x, y, z, c, o, true, out.ovf);
    }
}

// ---- 64->1 head conv, linear (+bias), writes (B,H^3,3): SR4DFlowNet.py:40,43,46,49 -------
// out[x,y,z] = b + sum_t dot64(hpad[x+dx, y+dy, z+dz], w[t]).  A CTA owns a 12x12 (y,z) output tile of one head and
// marches along x: for every padded plane xp it (1) stages the 14x14 halo tile of the plane (replicate halo is
// materialised in the Act) as fp32 in shared memory, (2) computes the 27 tap dot-products of every staged voxel as
// a register-tiled 196x28x64 GEMM (thread tile 4 voxels x 7 taps, float4 shared-memory operands), (3) scatters
// them into three running output-plane accumulators (plane xp feeds outputs xp-2..xp); a finished plane is
// written out.  Every activation is read once per (y,z) tile (1.36x halo amplification instead of 1.95x).
struct HeadArgs {
    const __half* hi[3];
    const __half* lo[3];
    const float* w[3];
    const float* b[3];
};
constexpr int HO_T = 12, HO_TP = HO_T + 2, HO_NV = HO_TP * HO_TP;      // 196 staged voxels per plane
constexpr int HO_AP = 68;                                              // activation row pitch (floats)
constexpr int HO_THREADS = 224;                                        // 49 voxel groups x 4 tap groups (+ spare)
constexpr int HO_SEG = 16;                                             // output planes per CTA
constexpr int HO_LD = (HO_NV * 8 + HO_THREADS - 1) / HO_THREADS;       // staged (voxel, 8-channel) items per thread
constexpr int HEAD_SMEM = (HO_NV * HO_AP + 64 * 32 + HO_NV * 28 + 3 * HO_T * HO_T) * 4;
__global__ void __launch_bounds__(HO_THREADS, 2) head_out_kernel(HeadArgs a, float* __restrict__ out, int B, int H) {
    extern __shared__ __align__(16) float hsm[];
    float* As = hsm;                           // [196][68]
    float* Ws = As + HO_NV * HO_AP;            // [64 k][4 tap groups][8]  (7 taps + a zero)
    float* Ts = Ws + 64 * 32;                  // [196][28]
    float* acc = Ts + HO_NV * 28;              // [3][144]
    const int nt = (H + HO_T - 1) / HO_T, nseg = (H + HO_SEG - 1) / HO_SEG;
    int bi = blockIdx.x;
    const int tz = bi % nt; bi /= nt;
    const int ty = bi % nt; bi /= nt;
    const int sg = bi % nseg; bi /= nseg;
    const int c = bi % 3;
    const int b = bi / 3;
    const int y0 = ty * HO_T, z0 = tz * HO_T, xs = sg * HO_SEG, xe = min(H, xs + HO_SEG);
    const int Hp = H + 2;
    const int tid = threadIdx.x;
    const __half* hi = a.hi[c];
    const __half* lo = a.lo[c];
    for (int i = tid; i < 64 * 32; i += HO_THREADS) {
        const int k = i >> 5, tg = (i >> 3) & 3, j = i & 7;
        Ws[i] = (j < 7 && tg * 7 + j < 27) ? a.w[c][(tg * 7 + j) * 64 + k] : 0.f;
    }
    for (int i = tid; i < 3 * HO_T * HO_T; i += HO_THREADS) acc[i] = 0.f;
    const float bias = a.b[c][0];
    const int tg = tid & 3, vg = tid >> 2;                 // GEMM role: taps tg*7..+6 of voxels vg + 49*i
    // all of a thread's loads of a plane are issued back to back (the two resident CTAs of an SM overlap one CTA's
    // load phase with the other's multiply phase)
    uint4 rh[HO_LD], rl[HO_LD];
    auto fetch = [&](int xp) {
#pragma unroll
        for (int j = 0; j < HO_LD; ++j) {
            const int i = tid + j * HO_THREADS;
            const int v = i >> 3, c8 = (i & 7) * 8;
            const int yp = y0 + v / HO_TP, zp = z0 + v % HO_TP;
            rh[j] = make_uint4(0, 0, 0, 0);
            rl[j] = make_uint4(0, 0, 0, 0);
            if (i < HO_NV * 8 && yp < Hp && zp < Hp) {
                const size_t off = ((((size_t)b * Hp + xp) * Hp + yp) * Hp + zp) * 64 + c8;
                rh[j] = *reinterpret_cast<const uint4*>(hi + off);
                rl[j] = *reinterpret_cast<const uint4*>(lo + off);
            }
        }
    };
    for (int xp = xs; xp < xe + 2; ++xp) {
        fetch(xp);
        __syncthreads();                                   // previous plane's gather is done with As / Ts
#pragma unroll
        for (int j = 0; j < HO_LD; ++j) {
            const int i = tid + j * HO_THREADS;
            if (i < HO_NV * 8) {
                const int v = i >> 3, c8 = (i & 7) * 8;
                const __half2* hh = reinterpret_cast<const __half2*>(&rh[j]);
                const __half2* ll = reinterpret_cast<const __half2*>(&rl[j]);
                float xv[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 x2 = __half22float2(hh[k]), y2 = __half22float2(ll[k]);
                    xv[2 * k] = fmaf(y2.x, SR4D_LO_INV, x2.x);
                    xv[2 * k + 1] = fmaf(y2.y, SR4D_LO_INV, x2.y);
                }
                float4* dst = reinterpret_cast<float4*>(As + v * HO_AP + c8);
                dst[0] = make_float4(xv[0], xv[1], xv[2], xv[3]);
                dst[1] = make_float4(xv[4], xv[5], xv[6], xv[7]);
            }
        }
        __syncthreads();
        if (vg < 49) {
            float t[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) t[i][j] = 0.f;
#pragma unroll 4
            for (int k4 = 0; k4 < 16; ++k4) {
                float4 av[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(As + (vg + 49 * i) * HO_AP + k4 * 4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 w0 = *reinterpret_cast<const float4*>(Ws + (k4 * 4 + kk) * 32 + tg * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(Ws + (k4 * 4 + kk) * 32 + tg * 8 + 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float x = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
                        t[i][0] = fmaf(x, w0.x, t[i][0]); t[i][1] = fmaf(x, w0.y, t[i][1]);
                        t[i][2] = fmaf(x, w0.z, t[i][2]); t[i][3] = fmaf(x, w0.w, t[i][3]);
                        t[i][4] = fmaf(x, w1.x, t[i][4]); t[i][5] = fmaf(x, w1.y, t[i][5]);
                        t[i][6] = fmaf(x, w1.z, t[i][6]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 7; ++j)
                    if (tg * 7 + j < 28) Ts[(vg + 49 * i) * 28 + tg * 7 + j] = t[i][j];
        }
        __syncthreads();
        // plane xp feeds output plane xp - dx through the nine (dy,dz) taps of x-offset dx
        for (int i = tid; i < 3 * HO_T * HO_T; i += HO_THREADS) {
            const int dx = i / (HO_T * HO_T), o = i % (HO_T * HO_T);
            const int x = xp - dx;
            if (x < xs || x >= xe) continue;
            const int oy = o / HO_T, oz = o % HO_T;
            float s = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dz = 0; dz < 3; ++dz)
                    s += Ts[((oy + dy) * HO_TP + oz + dz) * 28 + dx * 9 + dy * 3 + dz];
            float* ap = acc + (x % 3) * (HO_T * HO_T) + o;
            s += *ap;
            if (dx == 2) {                                  // last contribution: emit and recycle the slot
                const int y = y0 + oy, z = z0 + oz;
                if (y < H && z < H) out[((((size_t)b * H + x) * H + y) * H + z) * 3 + c] = s + bias;
                *ap = 0.f;
            } else {
                *ap = s;
            }
        }
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
    act_store4_halo(out.hi, out.lo, D, b, xx, y, z, c, vv, true, out.ovf);
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
cudaError_t launch_stem_convs(const float* feat, const float* w_pc, const float* b_pc, ActView out_pc, const float* w_ph,
                              const float* b_ph, ActView out_ph, cudaStream_t s) {
    StemArgs a;
    a.w[0] = w_pc; a.bias[0] = b_pc; a.out[0] = out_pc; a.ch0[0] = 3;      // pc = concat[pcmr, mag, speed]   (SR4DFlowNet.py:15,17)
    a.w[1] = w_ph; a.bias[1] = b_ph; a.out[1] = out_ph; a.ch0[1] = 0;      // phase = concat[u, v, w]           (:14,20)
    const ActView& out = out_pc;
    const size_t nrun4 = (size_t)out.B * out.D * out.D * ((out.D + 3) / 4);
    const size_t ngrp4 = (nrun4 + 63) / 64;
    if (ngrp4 >= (size_t)tc_num_sms()) {
        stem_conv_kernel<4, 16><<<dim3((unsigned)(ngrp4 < 1184 ? ngrp4 : 1184), 2), 256, 0, s>>>(feat, a);
    } else {
        const size_t ngrp1 = ((size_t)out.B * out.D * out.D * out.D + 31) / 32;
        stem_conv_kernel<1, 8><<<dim3((unsigned)(ngrp1 < 1184 ? ngrp1 : 1184), 2), 256, 0, s>>>(feat, a);
    }
    return cudaGetLastError();
}
cudaError_t launch_conv1x1_cat(ActView a, ActView b, const float* w, const float* bias, ActView out, cudaStream_t s) {
    size_t nvox = (size_t)out.B * out.D * out.D * out.D;
    const size_t ngrp4 = ((nvox + 3) / 4 + 63) / 64;
    if (ngrp4 >= 2 * (size_t)tc_num_sms()) {
        conv1x1_cat_kernel<4, 16><<<(unsigned)(ngrp4 < 2368 ? ngrp4 : 2368), 256, 128 * 64 * 4, s>>>(a, b, w, bias, out);
    } else {                                       // small grids: one voxel and 8 channels per thread, eight times the CTAs
        const size_t ngrp1 = (nvox + 31) / 32;
        conv1x1_cat_kernel<1, 8><<<(unsigned)(ngrp1 < 2368 ? ngrp1 : 2368), 256, 128 * 64 * 4, s>>>(a, b, w, bias, out);
    }
    return cudaGetLastError();
}
cudaError_t launch_upsample(ActView in, ActView out, int r, UpsampleTables t, cudaStream_t s) {
    (void)r;
    size_t nline = (size_t)out.B * out.D * out.D;
    const unsigned nb = (unsigned)((nline + 31) / 32);
    int nseg = 1;
    while (nseg < 4 && (size_t)nb * nseg < 2 * (size_t)tc_num_sms() && out.D / (2 * nseg) >= 8) nseg *= 2;
    const int zseg = (out.D + nseg - 1) / nseg;
    if (nseg == 1) upsample_kernel<false><<<nb, 256, 0, s>>>(in, out, t, zseg);
    else upsample_kernel<true><<<dim3(nb, (unsigned)((out.D + zseg - 1) / zseg)), 256, 0, s>>>(in, out, t, zseg);
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
    cudaError_t ea = tc_func_smem(reinterpret_cast<const void*>(head_out_kernel), HEAD_SMEM);
    if (ea != cudaSuccess) return ea;
    const int nt = (h0.D + HO_T - 1) / HO_T, nseg = (h0.D + HO_SEG - 1) / HO_SEG;
    head_out_kernel<<<(unsigned)(h0.B * 3 * nseg * nt * nt), HO_THREADS, HEAD_SMEM, s>>>(a, out, h0.B, h0.D);
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
