// Loss, metric, regulariser value, fused flat Adam, and the inference stitcher.
#include "kernels.h"

namespace {

// ---- per-sample reductions for TrainerController.loss_function (TrainerController.py:84-127)
// and loss_utils.calculate_relative_error (loss_utils.py:64-103).
// partial[b][blk][5] = {sum mask, sum nf, sum se*mask, sum se*nf, sum rel}; stage 2 is
// deterministic (fixed order), accumulation in double.
__global__ void __launch_bounds__(256) loss_stats_kernel(const float* __restrict__ pred, const float* __restrict__ hu,
                                                         const float* __restrict__ hv, const float* __restrict__ hw,
                                                         const float* __restrict__ mask, int nvox,
                                                         double* __restrict__ partial) {
    const int b = blockIdx.y, nblk = gridDim.x;
    double s[5] = {0, 0, 0, 0, 0};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nvox; i += nblk * 256) {
        size_t vi = (size_t)b * nvox + i;
        float pu = pred[vi * 3], pv = pred[vi * 3 + 1], pw = pred[vi * 3 + 2];
        float tu = hu[vi], tv = hv[vi], tw = hw[vi], m = mask[vi];
        float du = pu - tu, dv = pv - tv, dw = pw - tw;
        float se = du * du + dv * dv + dw * dw;
        float nf = m < 0.5f ? 1.f : 0.f;
        float diff = sqrtf(se);
        float actual = sqrtf(tu * tu + tv * tv + tw * tw);
        float rel = diff / (actual + 1e-5f);
        rel = fminf(fmaxf(rel, 0.f), 1.f);
        rel = (actual != 0.f) ? rel : diff;
        rel = rintf(rel * 1e4f) / 1e4f;          // tf.round: half-to-even
        rel = (m == 1.0f) ? rel : 0.f;
        s[0] += m; s[1] += nf; s[2] += (double)(se * m); s[3] += (double)(se * nf); s[4] += rel;
    }
    __shared__ double red[8][5];
#pragma unroll
    for (int k = 0; k < 5; ++k) s[k] = warp_sum_d(s[k]);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 5; ++k) red[threadIdx.x >> 5][k] = s[k];
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0;
        for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
        partial[((size_t)b * nblk + blockIdx.x) * 5 + threadIdx.x] = t;
    }
}
__global__ void loss_final_kernel(const double* __restrict__ partial, int nblk, float* __restrict__ per_sample,
                                  float* __restrict__ norm) {
    const int b = blockIdx.x;
    __shared__ double tot[5];
    if (threadIdx.x < 5) {
        double t = 0;
        for (int i = 0; i < nblk; ++i) t += partial[((size_t)b * nblk + i) * 5 + threadIdx.x];
        tot[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sm = (float)tot[0], snf = (float)tot[1];
        float fluid = (float)tot[2] / (sm + 1.f);
        float nonfl = (float)tot[3] / (snf + 1.f);
        float loss = fluid + nonfl;
        per_sample[b * 4 + 0] = loss;
        per_sample[b * 4 + 1] = loss;                                   // mse == loss (divergence term is 0)
        per_sample[b * 4 + 2] = (float)tot[4] / (sm + 1.f) * 100.f;
        per_sample[b * 4 + 3] = sm;
        norm[b * 2] = sm;
        norm[b * 2 + 1] = snf;
    }
}
// d loss_b / d pred = 2 (pred - y) * (mask/(sum mask + 1) + nf/(sum nf + 1))
__global__ void loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ hu,
                                 const float* __restrict__ hv, const float* __restrict__ hw,
                                 const float* __restrict__ mask, int nvox, const float* __restrict__ norm,
                                 float* __restrict__ g, unsigned int* gmax, size_t total) {
    size_t vi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float mx = 0.f;
    if (vi < total) {
    int b = vi / nvox;
    float sm = norm[b * 2], snf = norm[b * 2 + 1];
    float m = mask[vi];
    float nf = m < 0.5f ? 1.f : 0.f;
    float wgt = 2.f * (m / (sm + 1.f) + nf / (snf + 1.f));
    const float g0 = (pred[vi * 3 + 0] - hu[vi]) * wgt, g1 = (pred[vi * 3 + 1] - hv[vi]) * wgt;
    const float g2 = (pred[vi * 3 + 2] - hw[vi]) * wgt;
    g[vi * 3 + 0] = g0; g[vi * 3 + 1] = g1; g[vi * 3 + 2] = g2;
    mx = fmaxf(fabsf(g0), fmaxf(fabsf(g1), fabsf(g2)));
    }
    // max |g| of the whole tensor (bounds the head gradients' scale): one atomic per warp that has a larger value
    if (gmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0 && __float_as_uint(mx) > *reinterpret_cast<volatile unsigned int*>(gmax))
            atomicMax(gmax, __float_as_uint(mx));
    }
}

// ---- regulariser value: TrainerController.py:129-141 --------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ p, const unsigned char* __restrict__ kflag,
                                                    int64_t n, double* __restrict__ partial) {
    double s = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        if (kflag[i >> 5]) s += (double)p[i] * (double)p[i];
    __shared__ double red[8];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < 8; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}
__global__ void sumsq_final_kernel(const double* __restrict__ partial, int nblk, float coeff, float* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < nblk; ++i) t += partial[i];
        *out = (float)(t * (double)coeff);
    }
}

// ---- Keras Adam (ResourceApplyAdam): TrainerController.py:73,225 ----------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, const unsigned char* __restrict__ kflag, int64_t n4, float alpha,
                            float beta1, float beta2, float eps, float l2s, const float* __restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    // data parallel: l2s is the regulariser gradient PER SAMPLE and *count the global batch size the all-reduce
    // left in the gradient buffer's metric tail (no host round trip for ragged shards)
    if (count) l2s *= *count;
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    const float l2 = kflag[i >> 3] ? l2s : 0.f;
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float gk = fmaf(l2, P[k], G[k]);
        M[k] += (gk - M[k]) * (1.f - beta1);
        V[k] += (gk * gk - V[k]) * (1.f - beta2);
        P[k] -= alpha * M[k] / (sqrtf(V[k]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
}

// ---- PatchGenerator._patchup_with_overlap + predictor.py:99-107 ------------------------
__global__ void stitch_kernel(const float* __restrict__ pred, int nx, int ny, int nz, int H, int crop, int VX,
                              int VY, int VZ, float venc, int round_small, float* __restrict__ vol) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)VX * VY * VZ;
    if (i >= n) return;
    const int core = H - 2 * crop;
    int z = i % VZ, y = (i / VZ) % VY, x = i / ((size_t)VZ * VY);
    int px = x / core, py = y / core, pz = z / core;
    size_t patch = ((size_t)px * ny + py) * nz + pz;
    int lx = x % core + crop, ly = y % core + crop, lz = z % core + crop;
    const float* src = pred + ((((size_t)patch * H + lx) * H + ly) * H + lz) * 3;
    const float thr = venc / 2048.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = src[c] * venc;
        if (round_small && fabsf(v) < thr) v = 0.f;
        vol[c * n + i] = v;
    }
    (void)nx;
}
}  // namespace

cudaError_t launch_loss_stats(const float* pred, const float* hu, const float* hv, const float* hw,
                              const float* mask, int B, int nvox, double* partial, int nblk, float* per_sample,
                              float* norm, cudaStream_t s) {
    dim3 grid(nblk, B);
    loss_stats_kernel<<<grid, 256, 0, s>>>(pred, hu, hv, hw, mask, nvox, partial);
    loss_final_kernel<<<B, 32, 0, s>>>(partial, nblk, per_sample, norm);
    return cudaGetLastError();
}
cudaError_t launch_loss_grad(const float* pred, const float* hu, const float* hv, const float* hw,
                             const float* mask, int B, int nvox, const float* norm, float* g, unsigned int* gmax,
                             cudaStream_t s) {
    size_t total = (size_t)B * nvox;
    if (gmax) cudaMemsetAsync(gmax, 0, sizeof(unsigned int), s);
    loss_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(pred, hu, hv, hw, mask, nvox, norm, g, gmax, total);
    return cudaGetLastError();
}
cudaError_t launch_sumsq(const float* p, const unsigned char* kflag, int64_t n, double* partial, int nblk,
                         float coeff, float* out, cudaStream_t s) {
    sumsq_kernel<<<nblk, 256, 0, s>>>(p, kflag, n, partial);
    sumsq_final_kernel<<<1, 32, 0, s>>>(partial, nblk, coeff, out);
    return cudaGetLastError();
}
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, const unsigned char* kflag, int64_t n,
                        float alpha, float beta1, float beta2, float eps, float l2_scale, const float* count_dev,
                        cudaStream_t s) {
    int64_t n4 = n / 4;   // flat size is a multiple of 32 floats
    adam_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(p, g, m, v, kflag, n4, alpha, beta1, beta2, eps, l2_scale, count_dev);
    return cudaGetLastError();
}
cudaError_t launch_stitch(const float* pred, int nx, int ny, int nz, int H, int crop, int VX, int VY, int VZ,
                          float venc, int round_small, float* vol, cudaStream_t s) {
    size_t n = (size_t)VX * VY * VZ;
    stitch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pred, nx, ny, nz, H, crop, VX, VY, VZ, venc, round_small, vol);
    return cudaGetLastError();
}
