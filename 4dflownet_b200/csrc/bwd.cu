// Backward kernels (fp32 CUDA cores): the gradients TF generates for
// TrainerController.train_step (TrainerController.py:209-225) via tape.gradient --
// Conv3DBackpropFilter (wgrad), BiasAddGrad, MirrorPadGrad (halo fold), Relu/LeakyReluGrad,
// ResizeBilinearGrad (upsample) -- restated per SURVEY appendix C.  Conv3DBackpropInput for
// the 64->64 layers is conv64 with dgrad=1 (conv_simt.cu).
#include "kernels.h"
#include "tc_host.h"

namespace {

__device__ __forceinline__ size_t g4_off(int D, int b, int x, int y, int z) {
    const int dp = D + 4;
    return ((((size_t)b * dp + (x + 2)) * dp + (y + 2)) * dp + (z + 2)) * 64;
}
__device__ __forceinline__ size_t raw_off(int D, int b, int px, int py, int pz) {
    const int dp = D + 2;
    return ((((size_t)b * dp + px) * dp + py) * dp + pz) * 64;
}

// deterministic second stage of every split reduction: out[j] = sum_r partial[r][j].
// block = 32 columns x 8 row phases (fixed summation order: phase-strided partial sums, then phases 0..7)
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ partial, int nrows, int ncols,
                                                          float* __restrict__ out) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ph = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (j < ncols)
        for (int r = ph; r < nrows; r += 8) s += partial[(size_t)r * ncols + j];
    red[ph][cx] = s;
    __syncthreads();
    if (ph == 0 && j < ncols) {
        for (int k = 1; k < 8; ++k) s += red[k][cx];
        out[j] = s;
    }
}

// the same reduction for a batch of partial matrices in one launch (blockIdx.y = item): the 30 weight gradients of the
// 64->64 layers, summed once at the end of the backward pass instead of with one 10-us launch per layer
__global__ void __launch_bounds__(256) reduce_rows_batched_kernel(const ReduceItem* __restrict__ items, int nrows0, int nrows1,
                                                                  int ncols) {
    // one thread per four columns walks the rows in order with seven 16-byte loads in flight (650 MB at B = 8: the first
    // version -- 32 columns x 8 row phases per block, 4-byte loads -- ran at 2.9 TB/s)
    const ReduceItem it = items[blockIdx.y];
    const int nrows = it.cls ? nrows1 : nrows0;
    const int j = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (j >= ncols) return;
    const float4* p = reinterpret_cast<const float4*>(it.partial + j);
    const size_t pitch = (size_t)ncols / 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int r = 0;
    for (; r + 7 <= nrows; r += 7) {
        float4 v[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] = __ldcs(p + (size_t)(r + k) * pitch);
#pragma unroll
        for (int k = 0; k < 7; ++k) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
    }
    for (; r < nrows; ++r) {
        const float4 v = __ldcs(p + (size_t)r * pitch);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(it.out + j) = s;
}

// the same reduction for short, wide partial matrices (few rows, e.g. the 16 wgrad slabs): one thread per column
__global__ void reduce_rows_wide_kernel(const float* __restrict__ partial, int nrows, int ncols, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    float s = 0.f;
    for (int r = 0; r < nrows; ++r) s += partial[(size_t)r * ncols + j];
    out[j] = s;
}

// block-wide max of |v| folded into *p with at most one atomic per block (all threads must call)
__device__ __forceinline__ void absmax_commit(float m, unsigned int* p) {
    __shared__ float wm[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 1; w < nw; ++w) m = fmaxf(m, wm[w]);
        const unsigned int bits = __float_as_uint(m);
        if (bits > *reinterpret_cast<volatile unsigned int*>(p)) atomicMax(p, bits);
    }
}
// ---- 64->1 head conv (SR4DFlowNet.py:40,43,46), whole backward in one pass ------------------
// With the clamp padding folded in, both gradients of out[i] = b + sum_t w[t] . hpad[i+t] share one
// channel-independent quantity, G(j,t) = sum of g[i] over the output voxels i whose tap t reads
// (clamped) input voxel j:
//     dh[j][ci] = relu'(h[j][ci]) * sum_t w[t][ci] * G(j,t)          (dgrad + MirrorPadGrad + ReluGrad)
//     dw[t][ci] = sum_j h[j][ci] * G(j,t)                            (Conv3DBackpropFilter)
//     db        = sum_i g[i]
// Per axis the set {i : clamp(i + t - 1) = j} is {j - t + 1} plus {j} again when (j==0,t==0) or
// (j==D-1,t==2).  A block walks (b,x,y) z-lines: it stages the 3x3 neighbouring g lines, builds
// G[z][27] in shared memory, then 64 channels x 4 z-phases of threads stream h once, write dh (fp32 G4
// interior) and keep dw in registers across lines.  partial[blk][29][64]: rows 0..26 = dw, row 27 = db,
// row 28 = per-channel sum of dh (the bias gradient of the head's first conv).
// HZ_MAX = z-voxels per thread (4 z-phases): 12 serves H <= 48, 32 serves H <= 128
template <int HZ_MAX>
__global__ void __launch_bounds__(256) head2_bwd_kernel(ActView h, const float* __restrict__ g, int c,
                                                        const float* __restrict__ w, float* __restrict__ out_g4,
                                                        unsigned int* amax, float* __restrict__ partial,
                                                        __half* __restrict__ split_hi, __half* __restrict__ split_lo,
                                                        int* split_exp, const unsigned int* gmax) {
    extern __shared__ __align__(16) float h2sm[];
    const int H = h.D, Hz = H + 2;
    float* gl = h2sm;                              // [3][3][H+2]
    float* Rl = h2sm + 9 * Hz;                     // [3][3][H+2]
    float* Gs = h2sm + (18 * Hz + 3) / 4 * 4;      // [H][28], 16-byte aligned rows
    const int ci = threadIdx.x & 63, q = threadIdx.x >> 6;
    float wr[27], dw[29];
    float wabs = 0.f;
#pragma unroll
    for (int t = 0; t < 27; ++t) { wr[t] = w[t * 64 + ci]; dw[t] = 0.f; wabs += fabsf(wr[t]); }
    dw[27] = dw[28] = 0.f;
    float m = 0.f;
    // scale of the split copy: |dh| <= 8 (fold multiplicity of G) * max|g| * max_ci sum_t |w[t][ci]|
    float ksplit = 0.f;
    if (split_hi) {
        h2sm[threadIdx.x] = wabs;
        __syncthreads();
        float wmax = 0.f;
        for (int i = 0; i < 64; ++i) wmax = fmaxf(wmax, h2sm[i]);
        const float bound = 8.f * wmax * __uint_as_float(*gmax);
        int e = 0;
        if (bound > 0.f && bound < 3.0e38f) e = 14 - ilogbf(bound);
        e = max(-120, min(120, e));
        ksplit = exp2f((float)e);
        if (blockIdx.x == 0 && threadIdx.x == 0) *split_exp = e;
        __syncthreads();
    }
    const int nlines = h.B * H * H;
    for (int line = blockIdx.x; line < nlines; line += gridDim.x) {
        const int y = line % H, x = (line / H) % H, b = line / (H * H);
        // this thread's voxels z = q, q+4, ...: issue the (2-byte, latency-bound) loads before building G
        __half hph[HZ_MAX], hpl[HZ_MAX];           // raw planes: converting here would wait for the loads
        const size_t lineo = act_off(H, b, x, y, 0) + ci;
#pragma unroll
        for (int u = 0; u < HZ_MAX; ++u) {
            const int z = q + 4 * u;
            if (z < H) {
                hph[u] = h.hi[lineo + (size_t)z * 64];
                hpl[u] = h.lo[lineo + (size_t)z * 64];
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 9 * Hz; i += 256) {
            const int r = i / Hz, zz = i % Hz - 1;
            const int ix = x + r / 3 - 1, iy = y + r % 3 - 1;
            float v = 0.f;
            if (ix >= 0 && ix < H && iy >= 0 && iy < H && zz >= 0 && zz < H)
                v = g[((((size_t)b * H + ix) * H + iy) * H + zz) * 3 + c];
            gl[i] = v;
        }
        __syncthreads();
        // R[tx][ty][zz]: the x/y part of the clamp sets summed once per (tx,ty) line; G then needs <= 2 terms
        for (int i = threadIdx.x; i < 9 * Hz; i += 256) {
            const int r = i / Hz, zz = i % Hz;
            const int tx = r / 3, ty = r % 3;
            const int ex = (x == 0 && tx == 0) || (x == H - 1 && tx == 2);
            const int ey = (y == 0 && ty == 0) || (y == H - 1 && ty == 2);
            float s = gl[((2 - tx) * 3 + (2 - ty)) * Hz + zz];
            if (ex) s += gl[(3 + (2 - ty)) * Hz + zz];
            if (ey) s += gl[((2 - tx) * 3 + 1) * Hz + zz];
            if (ex && ey) s += gl[4 * Hz + zz];
            Rl[i] = s;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < H * 28; i += 256) {
            const int z = i / 28, t = i % 28;
            float s;
            if (t == 27) {
                s = gl[4 * Hz + z + 1];                      // g at the voxel itself (bias gradient)
            } else {
                const int tz = t % 3;
                const float* rp = Rl + (t / 3) * Hz + z + 1;
                s = rp[1 - tz];
                if ((z == 0 && tz == 0) || (z == H - 1 && tz == 2)) s += rp[0];
            }
            Gs[i] = s;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < HZ_MAX; ++u) {
            const int z = q + 4 * u;
            if (z < H) {
                const float4* G4p = reinterpret_cast<const float4*>(Gs + z * 28);
                float Gt[28];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const float4 v = G4p[k];
                    Gt[4 * k] = v.x; Gt[4 * k + 1] = v.y; Gt[4 * k + 2] = v.z; Gt[4 * k + 3] = v.w;
                }
                const float hv = join_f16(hph[u], hpl[u]);
                float sa = 0.f;
#pragma unroll
                for (int t = 0; t < 27; ++t) {
                    sa = fmaf(wr[t], Gt[t], sa);
                    dw[t] = fmaf(hv, Gt[t], dw[t]);
                }
                dw[27] += Gt[27];
                const float d = hv > 0.f ? sa : 0.f;
                const size_t go = g4_off(H, b, x, y, z) + ci;
                out_g4[go] = d;
                if (split_hi) {
                    __half sh, sl;
                    split_f16(d * ksplit, sh, sl);
                    split_hi[go] = sh;
                    if (split_lo) split_lo[go] = sl;
                }
                dw[28] += d;
                m = fmaxf(m, fabsf(d));
            }
        }
    }
    __syncthreads();
    float* red = h2sm;                             // [4][29][64] (aliases gl / Gs)
#pragma unroll
    for (int t = 0; t < 29; ++t) red[(q * 29 + t) * 64 + ci] = dw[t];
    __syncthreads();
    for (int i = threadIdx.x; i < 29 * 64; i += 256)
        partial[(size_t)blockIdx.x * 29 * 64 + i] = red[i] + red[29 * 64 + i] + red[2 * 29 * 64 + i] + red[3 * 29 * 64 + i];
    if (amax) absmax_commit(m, amax);
}

__device__ __forceinline__ float exp2_scale(const int* e, int sign) { return e ? exp2f((float)(sign * *e)) : 1.f; }

// ---- halo fold (MirrorPadGrad) + add + activation gradient -------------------------------
// raw_i carry the power-of-two scale 2^(*e_i) of the split-fp16 gradient they were computed from
__global__ void __launch_bounds__(256) fold_act_kernel(const float* __restrict__ r0, const float* __restrict__ r1,
                                                       const float* __restrict__ r2, const int* e0, const int* e1,
                                                       const int* e2, const float* __restrict__ add,
                                                       const __half* __restrict__ shi, const __half* __restrict__ slo,
                                                       float slope, float* __restrict__ out, unsigned int* amax,
                                                       int B, int D) {
    const size_t n = (size_t)B * D * D * D * 16;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    float m = 0.f;
    if (i < n) {
    const float k0 = exp2_scale(e0, -1), k1 = exp2_scale(e1, -1), k2 = exp2_scale(e2, -1);
    size_t vi = i >> 4;
    const int c = (i & 15) * 4;
    int z = vi % D, y = (vi / D) % D, x = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
        if (dx == -1 && x != 0) continue;
        if (dx == 1 && x != D - 1) continue;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            if (dy == -1 && y != 0) continue;
            if (dy == 1 && y != D - 1) continue;
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz) {
                if (dz == -1 && z != 0) continue;
                if (dz == 1 && z != D - 1) continue;
                size_t o = raw_off(D, b, x + 1 + dx, y + 1 + dy, z + 1 + dz) + c;
                float4 a = *reinterpret_cast<const float4*>(r0 + o);
                s.x = fmaf(a.x, k0, s.x); s.y = fmaf(a.y, k0, s.y); s.z = fmaf(a.z, k0, s.z); s.w = fmaf(a.w, k0, s.w);
                if (r1) { a = *reinterpret_cast<const float4*>(r1 + o); s.x = fmaf(a.x, k1, s.x); s.y = fmaf(a.y, k1, s.y); s.z = fmaf(a.z, k1, s.z); s.w = fmaf(a.w, k1, s.w); }
                if (r2) { a = *reinterpret_cast<const float4*>(r2 + o); s.x = fmaf(a.x, k2, s.x); s.y = fmaf(a.y, k2, s.y); s.z = fmaf(a.z, k2, s.z); s.w = fmaf(a.w, k2, s.w); }
            }
        }
    }
    size_t go = g4_off(D, b, x, y, z) + c;
    if (add) {
        float4 a = *reinterpret_cast<const float4*>(add + go);
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
    if (shi) {
        float sv[4];
        act_load4(shi, slo, act_off(D, b, x, y, z) + c, sv);
        s.x *= act_grad_from_out(sv[0], slope);
        s.y *= act_grad_from_out(sv[1], slope);
        s.z *= act_grad_from_out(sv[2], slope);
        s.w *= act_grad_from_out(sv[3], slope);
    }
    *reinterpret_cast<float4*>(out + go) = s;
    m = fmaxf(fmaxf(fabsf(s.x), fabsf(s.y)), fmaxf(fabsf(s.z), fabsf(s.w)));
    }
    if (amax) absmax_commit(m, amax);
}

// ---- fp32 G4 interior -> split-fp16 copy scaled by 2^e, e chosen from the tensor's |max| so the
// largest element lands in [2^13, 2^14) (the tensor-core dgrad / wgrad operands) -----------------
__global__ void __launch_bounds__(256) g4_split_kernel(const float* __restrict__ src, const unsigned int* amax,
                                                       __half* __restrict__ hi, __half* __restrict__ lo,
                                                       int* exp_out, int B, int D) {
    const size_t n = (size_t)B * D * D * D * 16;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const float mx = __uint_as_float(*amax);
    int e = 0;
    if (mx > 0.f && mx < 3.0e38f) e = 13 - ilogbf(mx);
    e = max(-120, min(120, e));
    if (i == 0) *exp_out = e;
    if (i >= n) return;
    const float k = exp2f((float)e);
    size_t vi = i >> 4;
    const int c = (i & 15) * 4;
    int z = vi % D, y = (vi / D) % D, x = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
    const size_t go = g4_off(D, b, x, y, z) + c;
    const float4 v = *reinterpret_cast<const float4*>(src + go);
    const float vv[4] = {v.x * k, v.y * k, v.z * k, v.w * k};
    uint2 h, l;
    act_pack4(vv, h, l);
    *reinterpret_cast<uint2*>(hi + go) = h;
    if (lo) *reinterpret_cast<uint2*>(lo + go) = l;
}

// ---- 64->64 3x3x3 weight gradient: dW[t][ci][co] = sum_{b,v} Xp[b,v+t][ci] dY[b,v][co] ------
// grid (9 (dx,dy) pairs, nchunk); block 256 = 16 ci-quads x 16 co-quads, 3 dz taps each.
constexpr int WG_ZS = 32;
__global__ void __launch_bounds__(256) wgrad64_kernel(ActView xin, const float* __restrict__ dy, float* __restrict__ partial) {
    __shared__ __align__(16) float xrow[(WG_ZS + 2) * 64];
    __shared__ __align__(16) float drow[WG_ZS * 64];
    const int D = xin.D, B = xin.B;
    const int dx = blockIdx.x / 3, dyy = blockIdx.x % 3;
    const int ciq = threadIdx.x & 15, coq = threadIdx.x >> 4;
    float acc[3][4][4];
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
    const int nlines = B * D * D;
    const int nseg = (D + WG_ZS - 1) / WG_ZS;
    for (int line = blockIdx.y; line < nlines; line += gridDim.y) {
        const int y = line % D, x = (line / D) % D, b = line / (D * D);
        for (int sg = 0; sg < nseg; ++sg) {
            const int zb = sg * WG_ZS;
            const int zl = min(WG_ZS, D - zb);
            __syncthreads();
            // input rows: padded coords (x+dx, y+dy, zb .. zb+zl+1)  == interior (x+dx-1, y+dy-1, zb-1 ..)
            for (int i = threadIdx.x; i < (zl + 2) * 8; i += 256) {
                int rz = i >> 3, c8 = (i & 7) * 8;
                float v[8];
                act_load8(xin.hi, xin.lo, act_off(D, b, x + dx - 1, y + dyy - 1, zb + rz - 1) + c8, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) xrow[rz * 64 + c8 + k] = v[k];
            }
            for (int i = threadIdx.x; i < zl * 16; i += 256) {
                int rz = i >> 4, c4 = (i & 15) * 4;
                *reinterpret_cast<float4*>(drow + rz * 64 + c4) =
                    *reinterpret_cast<const float4*>(dy + g4_off(D, b, x, y, zb + rz) + c4);
            }
            __syncthreads();
            float4 a0 = *reinterpret_cast<const float4*>(xrow + 0 * 64 + ciq * 4);
            float4 a1 = *reinterpret_cast<const float4*>(xrow + 1 * 64 + ciq * 4);
            for (int z = 0; z < zl; ++z) {
                float4 a2 = *reinterpret_cast<const float4*>(xrow + (z + 2) * 64 + ciq * 4);
                float4 d = *reinterpret_cast<const float4*>(drow + z * 64 + coq * 4);
                const float av[3][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}, {a2.x, a2.y, a2.z, a2.w}};
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[t][i][j] = fmaf(av[t][i], dv[j], acc[t][i][j]);
                a0 = a1; a1 = a2;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int tap = (dx * 3 + dyy) * 3 + t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float* o = partial + ((size_t)blockIdx.y * 27 + tap) * 4096 + (ciq * 4 + i) * 64 + coq * 4;
            *reinterpret_cast<float4*>(o) = make_float4(acc[t][i][0], acc[t][i][1], acc[t][i][2], acc[t][i][3]);
        }
    }
}

// ---- bias gradient (BiasAddGrad): column sums of a G4 interior ---------------------------
// 16 threads x float4 cover one voxel's 64 channels; 16 voxels per pass, grid-stride over voxels.
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dy, int B, int D, float* __restrict__ partial) {
    const size_t nvox = (size_t)B * D * D * D;
    const int c4 = (threadIdx.x & 15) * 4, sub = threadIdx.x >> 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t vi = (size_t)blockIdx.x * 16 + sub; vi < nvox; vi += (size_t)gridDim.x * 16) {
        int z = vi % D, y = (vi / D) % D, x = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
        const float4 v = *reinterpret_cast<const float4*>(dy + g4_off(D, b, x, y, z) + c4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    __shared__ float4 red[16][16];
    red[sub][threadIdx.x & 15] = s;
    __syncthreads();
    if (threadIdx.x < 16) {
        float4 t = red[0][threadIdx.x];
        for (int k = 1; k < 16; ++k) {
            const float4 v = red[k][threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        *reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * 64 + c4) = t;
    }
}

// out[j] = sum_r partial[r][j] for a few (<= 256) columns: 1024/ncols row phases in one block, fixed order
__global__ void __launch_bounds__(1024) reduce_rows_small_kernel(const float* __restrict__ partial, int nrows, int ncols,
                                                                 float* __restrict__ out) {
    __shared__ float red[1024];
    const int nph = 1024 / ncols;
    const int j = threadIdx.x % ncols, ph = threadIdx.x / ncols;
    float s = 0.f;
    if (ph < nph)
        for (int r = ph; r < nrows; r += nph) s += partial[(size_t)r * ncols + j];
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < ncols) {
        for (int k = 1; k < nph; ++k) s += red[k * ncols + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// ---- upsample backward: d_lr = U^T d_hr, then * act'(lr) --------------------------------
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ dhr, ActView lr, float slope,
                                                           float* __restrict__ dlr, unsigned int* amax, int B, int D,
                                                           int r, UpsampleTables t) {
    const int H = D * r;
    const size_t n = (size_t)B * D * D * D * 16;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    float m = 0.f;
    if (i < n) {
    size_t vi = i >> 4;
    const int c = (i & 15) * 4;
    int z = vi % D, y = (vi / D) % D, x = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    // transposed interpolation weights of the HR indices that touch this LR index, per axis (tables read once)
    auto axis_w = [&](int i, int j) {
        return (t.lo[i] == j ? 1.f - t.lerp[i] : 0.f) + (t.hi[i] == j ? t.lerp[i] : 0.f);
    };
    const int zb = t.ibeg[z], zn = t.iend[z] - zb;
    constexpr int MAXZ = 8;                       // covers res_increase <= 4; larger factors recompute in the loop
    float wz[MAXZ];
#pragma unroll
    for (int k = 0; k < MAXZ; ++k) wz[k] = (k < zn) ? axis_w(zb + k, z) : 0.f;
    for (int ix = t.ibeg[x]; ix < t.iend[x]; ++ix) {
        const float wx = axis_w(ix, x);
        if (wx == 0.f) continue;
        for (int iy = t.ibeg[y]; iy < t.iend[y]; ++iy) {
            const float wxy = wx * axis_w(iy, y);
            if (wxy == 0.f) continue;
            const float* row = dhr + g4_off(H, b, ix, iy, zb) + c;
            if (zn <= MAXZ) {
#pragma unroll
                for (int k = 0; k < MAXZ; ++k) {
                    if (k < zn && wz[k] != 0.f) {
                        const float wgt = wxy * wz[k];
                        const float4 a = *reinterpret_cast<const float4*>(row + (size_t)k * 64);
                        s.x = fmaf(wgt, a.x, s.x); s.y = fmaf(wgt, a.y, s.y); s.z = fmaf(wgt, a.z, s.z); s.w = fmaf(wgt, a.w, s.w);
                    }
                }
            } else {
                for (int k = 0; k < zn; ++k) {
                    const float wgt = wxy * axis_w(zb + k, z);
                    if (wgt == 0.f) continue;
                    const float4 a = *reinterpret_cast<const float4*>(row + (size_t)k * 64);
                    s.x = fmaf(wgt, a.x, s.x); s.y = fmaf(wgt, a.y, s.y); s.z = fmaf(wgt, a.z, s.z); s.w = fmaf(wgt, a.w, s.w);
                }
            }
        }
    }
    float sv[4];
    act_load4(lr.hi, lr.lo, act_off(D, b, x, y, z) + c, sv);
    s.x *= act_grad_from_out(sv[0], slope);
    s.y *= act_grad_from_out(sv[1], slope);
    s.z *= act_grad_from_out(sv[2], slope);
    s.w *= act_grad_from_out(sv[3], slope);
    *reinterpret_cast<float4*>(dlr + g4_off(D, b, x, y, z) + c) = s;
    m = fmaxf(fmaxf(fabsf(s.x), fabsf(s.y)), fmaxf(fabsf(s.z), fabsf(s.w)));
    }
    if (amax) absmax_commit(m, amax);
}

// ---- 1x1 (128->64) backward ----------------------------------------------------------------
// input gradient: d_cat[v][k] = sum_co dy[v][co] W[k][co]; 8 threads per voxel, k = kg*4 + 32q + (0..3)
constexpr int C1D_NV = 4;
__global__ void __launch_bounds__(256) conv1x1_dgrad_kernel(const float* __restrict__ dy, ActView a, ActView bq,
                                                            const float* __restrict__ w, float* __restrict__ da,
                                                            float* __restrict__ db, unsigned int* amax_a,
                                                            unsigned int* amax_b) {
    extern __shared__ float wt[];   // [64 co][128 k]
    for (int i = threadIdx.x; i < 128 * 64; i += 256) {
        int k = i >> 6, co = i & 63;
        wt[co * 128 + k] = w[i];
    }
    __syncthreads();
    const int D = a.D;
    const size_t nvox = (size_t)a.B * D * D * D;
    float ma = 0.f, mb = 0.f;
    const int kg = threadIdx.x & 7;
    // grid-stride over groups of 32 runs of C1D_NV consecutive voxels (8 threads per run): the transposed weights are
    // staged once per CTA and every 16-byte weight fetch feeds 4 * C1D_NV FMAs
    const size_t nrun = (nvox + C1D_NV - 1) / C1D_NV;
    for (size_t ri = (size_t)blockIdx.x * 32 + (threadIdx.x >> 3); ri < nrun; ri += (size_t)gridDim.x * 32) {
        size_t go[C1D_NV];
#pragma unroll
        for (int v = 0; v < C1D_NV; ++v) {
            size_t vi = ri * C1D_NV + v;
            if (vi >= nvox) vi = nvox - 1;                       // tail: recomputed, store skipped below
            go[v] = g4_off(D, (int)(vi / ((size_t)D * D * D)), (int)((vi / ((size_t)D * D)) % D), (int)((vi / D) % D),
                           (int)(vi % D));
        }
        float acc[C1D_NV][4][4];
#pragma unroll
        for (int v = 0; v < C1D_NV; ++v)
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[v][q][j] = 0.f;
        for (int c4 = 0; c4 < 16; ++c4) {
            float4 d4[C1D_NV];
#pragma unroll
            for (int v = 0; v < C1D_NV; ++v) d4[v] = *reinterpret_cast<const float4*>(dy + go[v] + c4 * 4);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int co = c4 * 4 + cc;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = *reinterpret_cast<const float4*>(wt + co * 128 + q * 32 + kg * 4);
#pragma unroll
                    for (int v = 0; v < C1D_NV; ++v) {
                        const float d = cc == 0 ? d4[v].x : cc == 1 ? d4[v].y : cc == 2 ? d4[v].z : d4[v].w;
                        acc[v][q][0] = fmaf(d, wv.x, acc[v][q][0]); acc[v][q][1] = fmaf(d, wv.y, acc[v][q][1]);
                        acc[v][q][2] = fmaf(d, wv.z, acc[v][q][2]); acc[v][q][3] = fmaf(d, wv.w, acc[v][q][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < C1D_NV; ++v) {
            const size_t vi = ri * C1D_NV + v;
            if (vi >= nvox) break;
            const int z = (int)(vi % D), y = (int)((vi / D) % D), x = (int)((vi / ((size_t)D * D)) % D);
            const int b = (int)(vi / ((size_t)D * D * D));
            const size_t ao = act_off(D, b, x, y, z);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int k = q * 32 + kg * 4;
                const bool second = k >= 64;
                int kk = second ? k - 64 : k;
                float sv[4];
                act_load4(second ? bq.hi : a.hi, second ? bq.lo : a.lo, ao + kk, sv);
                float4 o = make_float4(sv[0] > 0.f ? acc[v][q][0] : 0.f, sv[1] > 0.f ? acc[v][q][1] : 0.f,
                                       sv[2] > 0.f ? acc[v][q][2] : 0.f, sv[3] > 0.f ? acc[v][q][3] : 0.f);
                *reinterpret_cast<float4*>((second ? db : da) + go[v] + kk) = o;
                const float mo = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w)));
                if (second) mb = fmaxf(mb, mo); else ma = fmaxf(ma, mo);
            }
        }
    }
    if (amax_a) { absmax_commit(ma, amax_a); __syncthreads(); absmax_commit(mb, amax_b); }
}
// weight gradient: dW[k][co] = sum_v cat[v][k] dy[v][co]; block stages 32 voxels; thread = 4 k x 8 co
constexpr int C1_VOX_PER_BLOCK = 256;
__global__ void __launch_bounds__(256) conv1x1_wgrad_kernel(const float* __restrict__ dy, ActView a, ActView bq,
                                                            float* __restrict__ partial) {
    __shared__ __align__(16) float cs[32 * 128];
    __shared__ __align__(16) float ds[32 * 64];
    const int D = a.D;
    const size_t nvox = (size_t)a.B * D * D * D;
    const int kq = threadIdx.x & 31, coq = threadIdx.x >> 5;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const size_t nstages = (nvox + 31) / 32;
    for (size_t st = blockIdx.x; st < nstages; st += gridDim.x) {
        size_t v0 = st * 32;
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 16; i += 256) {   // 32 voxels x 16 groups of 8 channels
            int vv = i >> 4, g8 = i & 15;
            size_t vi = v0 + vv;
            float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (vi < nvox) {
                int z = vi % D, y = (vi / D) % D, x = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
                size_t ao = act_off(D, b, x, y, z);
                if (g8 < 8) act_load8(a.hi, a.lo, ao + g8 * 8, v);
                else act_load8(bq.hi, bq.lo, ao + (g8 - 8) * 8, v);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) cs[vv * 128 + g8 * 8 + k] = v[k];
        }
        for (int i = threadIdx.x; i < 32 * 16; i += 256) {
            int vv = i >> 4, c4 = (i & 15) * 4;
            size_t vi = v0 + vv;
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vi < nvox) {
                int z = vi % D, y = (vi / D) % D, x = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
                d = *reinterpret_cast<const float4*>(dy + g4_off(D, b, x, y, z) + c4);
            }
            *reinterpret_cast<float4*>(ds + vv * 64 + c4) = d;
        }
        __syncthreads();
        for (int vv = 0; vv < 32; ++vv) {
            float4 cv = *reinterpret_cast<const float4*>(cs + vv * 128 + kq * 4);
            float4 d0 = *reinterpret_cast<const float4*>(ds + vv * 64 + coq * 8);
            float4 d1 = *reinterpret_cast<const float4*>(ds + vv * 64 + coq * 8 + 4);
            const float c4[4] = {cv.x, cv.y, cv.z, cv.w};
            const float d8[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(c4[i], d8[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float* o = partial + (size_t)blockIdx.x * 8192 + (kq * 4 + i) * 64 + coq * 8;
        *reinterpret_cast<float4*>(o) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
}

// ---- stem 3->64 weight gradient: dW[t][c][co] = sum_v feat[clamp(v+t)][c] * dY[v][co] ------------
// A block walks (b,x,y) z-lines: the 3x3 clamped neighbour lines of the 3-channel features are staged in
// shared memory (float4 per voxel), 64 output channels x 4 z-phases of threads keep dW in registers.
// The gradient values of a thread's voxels are requested before the feature lines are staged (one exposed global latency per
// line instead of one per voxel) and the four z-phases meet once in shared memory at the end (the first version went
// through 162 block barriers there): 159 -> ~60 us per launch at B = 8, P = 24.
constexpr int SW_ZMAX = 8;     // voxels per thread and line in registers (P <= 32); longer lines take further passes
__global__ void __launch_bounds__(256, 2) stem_wgrad_kernel(const float* __restrict__ feat, int ch0,
                                                         const float* __restrict__ dy, int B, int P,
                                                         float* __restrict__ partial) {
    extern __shared__ __align__(16) float swsm[];      // [9][P+2] float4, later [4][81][64] floats
    float4* fl = reinterpret_cast<float4*>(swsm);
    const int Pz = P + 2;
    const int co = threadIdx.x & 63, q = threadIdx.x >> 6;
    float acc[81];
#pragma unroll
    for (int t = 0; t < 81; ++t) acc[t] = 0.f;
    const int nlines = B * P * P;
    for (int line = blockIdx.x; line < nlines; line += gridDim.x) {
        const int y = line % P, x = (line / P) % P, b = line / (P * P);
        for (int zb = 0; zb < P; zb += 4 * SW_ZMAX) {
            float d[SW_ZMAX];
#pragma unroll
            for (int u = 0; u < SW_ZMAX; ++u) {
                const int z = zb + q + 4 * u;
                d[u] = z < P ? dy[g4_off(P, b, x, y, z) + co] : 0.f;
            }
            if (zb == 0) {
                __syncthreads();
                for (int i = threadIdx.x; i < 9 * Pz; i += 256) {
                    const int r = i / Pz;
                    const int zz = min(max(i % Pz - 1, 0), P - 1);
                    const int xx = min(max(x + r / 3 - 1, 0), P - 1), yy = min(max(y + r % 3 - 1, 0), P - 1);
                    const float* f = feat + ((((size_t)b * P + xx) * P + yy) * P + zz) * 6 + ch0;
                    fl[i] = make_float4(f[0], f[1], f[2], 0.f);
                }
                __syncthreads();
            }
#pragma unroll
            for (int u = 0; u < SW_ZMAX; ++u) {
                const int z = zb + q + 4 * u;
                if (z < P) {
#pragma unroll
                    for (int r = 0; r < 9; ++r)
#pragma unroll
                        for (int dz = 0; dz < 3; ++dz) {
                            const float4 f = fl[r * Pz + z + dz];
                            const int t = (r * 3 + dz) * 3;
                            acc[t] = fmaf(f.x, d[u], acc[t]);
                            acc[t + 1] = fmaf(f.y, d[u], acc[t + 1]);
                            acc[t + 2] = fmaf(f.z, d[u], acc[t + 2]);
                        }
                }
            }
        }
    }
    __syncthreads();
    float* red = swsm;                                 // [4][81][64]
#pragma unroll
    for (int t = 0; t < 81; ++t) red[(q * 81 + t) * 64 + co] = acc[t];
    __syncthreads();
    for (int i = threadIdx.x; i < 81 * 64; i += 256)
        partial[(size_t)blockIdx.x * 81 * 64 + i] = (red[i] + red[81 * 64 + i]) + (red[2 * 81 * 64 + i] + red[3 * 81 * 64 + i]);
}

__global__ void g4_from_dense_kernel(const float* __restrict__ dense, float* __restrict__ g4, unsigned int* amax,
                                     int B, int D) {
    const size_t n = (size_t)B * D * D * D * 16;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    float m = 0.f;
    if (i < n) {
        size_t vi = i >> 4;
        int c = (i & 15) * 4;
        int z = vi % D, y = (vi / D) % D, x = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
        const float4 v = *reinterpret_cast<const float4*>(dense + vi * 64 + c);
        *reinterpret_cast<float4*>(g4 + g4_off(D, b, x, y, z) + c) = v;
        m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    if (amax) absmax_commit(m, amax);
}
__global__ void dense_from_g4_kernel(const float* __restrict__ g4, float* __restrict__ dense, int B, int D) {
    const size_t n = (size_t)B * D * D * D * 16;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    size_t vi = i >> 4;
    int c = (i & 15) * 4;
    int z = vi % D, y = (vi / D) % D, x = (vi / ((size_t)D * D)) % D, b = vi / ((size_t)D * D * D);
    *reinterpret_cast<float4*>(dense + vi * 64 + c) = *reinterpret_cast<const float4*>(g4 + g4_off(D, b, x, y, z) + c);
}

inline unsigned nblocks(size_t n, int per) { return (unsigned)((n + per - 1) / per); }
constexpr unsigned MAX_RED_BLOCKS = 1184;   // 148 SMs x 8: bounds the split-reduction scratch
inline unsigned red_blocks(size_t n, int per) { unsigned b = nblocks(n, per); return b < MAX_RED_BLOCKS ? b : MAX_RED_BLOCKS; }
}  // namespace

cudaError_t launch_head2_bwd(ActView h, const float* g, int c, const float* w, float* out_g4, unsigned int* amax,
                             float* dw, float* db, float* db1, __half* split_out, int* split_exp,
                             const unsigned int* gmax, float* scratch, cudaStream_t s, bool hi_only) {
    const int H = h.D;
    const int nlines = h.B * H * H;
    const unsigned nb = nlines < 592 ? nlines : 592;           // 148 SMs x 4 resident blocks
    size_t smem = (size_t)((18 * (H + 2) + 3) / 4 * 4 + H * 28) * sizeof(float);
    if (smem < 4 * 29 * 64 * sizeof(float)) smem = 4 * 29 * 64 * sizeof(float);
    if (H > 128 || (split_out && (!split_exp || !gmax))) return cudaErrorInvalidValue;
    float* tmp = scratch + (size_t)nb * 29 * 64;   // reduced [29][64]
    __half* shi = split_out;
    __half* slo = (split_out && !hi_only) ? split_out + (size_t)h.B * (H + 4) * (H + 4) * (H + 4) * 64 : nullptr;
    if (H <= 48) head2_bwd_kernel<12><<<nb, 256, smem, s>>>(h, g, c, w, out_g4, amax, scratch, shi, slo, split_exp, gmax);
    else head2_bwd_kernel<32><<<nb, 256, smem, s>>>(h, g, c, w, out_g4, amax, scratch, shi, slo, split_exp, gmax);
    launch_reduce_rows(scratch, nb, 29 * 64, tmp, s);
    cudaMemcpyAsync(dw, tmp, 27 * 64 * sizeof(float), cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(db, tmp + 27 * 64, sizeof(float), cudaMemcpyDeviceToDevice, s);
    if (db1) cudaMemcpyAsync(db1, tmp + 28 * 64, 64 * sizeof(float), cudaMemcpyDeviceToDevice, s);
    return cudaGetLastError();
}
cudaError_t launch_fold_act(const float* raw0, const float* raw1, const float* raw2, const int* e0, const int* e1,
                            const int* e2, const float* add_g4, const __half* saved_hi, const __half* saved_lo,
                            float slope, float* out_g4, unsigned int* amax, int B, int D, cudaStream_t s) {
    size_t n = (size_t)B * D * D * D * 16;
    fold_act_kernel<<<nblocks(n, 256), 256, 0, s>>>(raw0, raw1, raw2, e0, e1, e2, add_g4, saved_hi, saved_lo, slope,
                                                    out_g4, amax, B, D);
    return cudaGetLastError();
}
cudaError_t launch_g4_split(const float* g4, const unsigned int* amax, __half* split, int* exp_out, int B, int D,
                            cudaStream_t s, bool hi_only) {
    size_t n = (size_t)B * D * D * D * 16;
    const size_t plane = (size_t)B * (D + 4) * (D + 4) * (D + 4) * 64;
    g4_split_kernel<<<nblocks(n, 256), 256, 0, s>>>(g4, amax, split, hi_only ? nullptr : split + plane, exp_out, B, D);
    return cudaGetLastError();
}
cudaError_t launch_wgrad64_simt(ActView x, const float* dy_g4, float* dw, float* scratch, int nchunk,
                                cudaStream_t s) {
    dim3 grid(9, nchunk);
    wgrad64_kernel<<<grid, 256, 0, s>>>(x, dy_g4, scratch);
    reduce_rows_kernel<<<(27 * 4096 + 31) / 32, 256, 0, s>>>(scratch, nchunk, 27 * 4096, dw);
    return cudaGetLastError();
}
cudaError_t launch_reduce_rows(const float* partial, int nrows, int ncols, float* out, cudaStream_t s) {
    if (nrows <= 32 && ncols >= 16384) reduce_rows_wide_kernel<<<(ncols + 255) / 256, 256, 0, s>>>(partial, nrows, ncols, out);
    else reduce_rows_kernel<<<(ncols + 31) / 32, 256, 0, s>>>(partial, nrows, ncols, out);
    return cudaGetLastError();
}
cudaError_t launch_reduce_rows_batched(const ReduceItem* items_dev, int nitems, int nrows0, int nrows1, int ncols, cudaStream_t s) {
    if (nitems <= 0) return cudaSuccess;
    if (ncols % 4) return cudaErrorInvalidValue;
    reduce_rows_batched_kernel<<<dim3((ncols / 4 + 255) / 256, nitems), 256, 0, s>>>(items_dev, nrows0, nrows1, ncols);
    return cudaGetLastError();
}
cudaError_t launch_bias_grad(const float* dy_g4, int B, int D, float* db, float* scratch, cudaStream_t s) {
    size_t nvox = (size_t)B * D * D * D;
    unsigned nb = red_blocks(nvox, 64);
    bias_grad_kernel<<<nb, 256, 0, s>>>(dy_g4, B, D, scratch);
    reduce_rows_small_kernel<<<1, 1024, 0, s>>>(scratch, nb, 64, db);
    return cudaGetLastError();
}
cudaError_t launch_upsample_bwd(const float* dhr_g4, ActView lr_saved, float slope, float* dlr_g4,
                                unsigned int* amax, int B, int D, int r, UpsampleTables t, cudaStream_t s) {
    size_t n = (size_t)B * D * D * D * 16;
    upsample_bwd_kernel<<<nblocks(n, 256), 256, 0, s>>>(dhr_g4, lr_saved, slope, dlr_g4, amax, B, D, r, t);
    return cudaGetLastError();
}
cudaError_t launch_conv1x1_bwd(const float* dy_g4, ActView a, ActView b, const float* w, float* da_g4,
                               float* db_g4, unsigned int* amax_a, unsigned int* amax_b, float* dw, float* dbias,
                               float* scratch, cudaStream_t s) {
    size_t nvox = (size_t)a.B * a.D * a.D * a.D;
    const unsigned ngrp = nblocks((nvox + C1D_NV - 1) / C1D_NV, 32);
    conv1x1_dgrad_kernel<<<ngrp < 592 ? ngrp : 592, 256, 128 * 64 * 4, s>>>(dy_g4, a, b, w, da_g4, db_g4, amax_a, amax_b);
    unsigned nb = red_blocks(nvox, C1_VOX_PER_BLOCK);
    conv1x1_wgrad_kernel<<<nb, 256, 0, s>>>(dy_g4, a, b, scratch);
    reduce_rows_kernel<<<8192 / 32, 256, 0, s>>>(scratch, nb, 8192, dw);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_bias_grad(dy_g4, a.B, a.D, dbias, scratch, s);
}
cudaError_t launch_stem_wgrad(const float* feat, int ch0, const float* dy_g4, int B, int P, float* dw,
                              float* db, float* scratch, cudaStream_t s) {
    const int nlines = B * P * P;
    const unsigned nb = nlines < 296 ? nlines : 296;           // 148 SMs x 2 resident blocks (registers, 81 KB of shared memory)
    size_t smem = (size_t)9 * (P + 2) * sizeof(float4);
    if (smem < (size_t)4 * 81 * 64 * sizeof(float)) smem = (size_t)4 * 81 * 64 * sizeof(float);
    cudaError_t ea = tc_func_smem(reinterpret_cast<const void*>(stem_wgrad_kernel), (int)smem);
    if (ea != cudaSuccess) return ea;
    stem_wgrad_kernel<<<nb, 256, smem, s>>>(feat, ch0, dy_g4, B, P, scratch);
    reduce_rows_kernel<<<(81 * 64 + 31) / 32, 256, 0, s>>>(scratch, nb, 81 * 64, dw);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_bias_grad(dy_g4, B, P, db, scratch, s);
}
cudaError_t launch_g4_from_dense(const float* dense, float* g4, unsigned int* amax, int B, int D, cudaStream_t s) {
    size_t n = (size_t)B * D * D * D * 16;
    g4_from_dense_kernel<<<nblocks(n, 256), 256, 0, s>>>(dense, g4, amax, B, D);
    return cudaGetLastError();
}
cudaError_t launch_dense_from_g4(const float* g4, float* dense, int B, int D, cudaStream_t s) {
    size_t n = (size_t)B * D * D * D * 16;
    dense_from_g4_kernel<<<nblocks(n, 256), 256, 0, s>>>(g4, dense, B, D);
    return cudaGetLastError();
}
