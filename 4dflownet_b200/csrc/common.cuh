// Shared device helpers and tensor views for libsr4d (sm_100a only).
//
// Activation layout ("Act"): every 64-channel feature map lives in HBM as a PAIR of
// fp16 planes, hi and lo, each [B][D+2][D+2][D+2][64] channels-last with a one-voxel
// replicate halo (the reference's tf.pad(...,'SYMMETRIC'), SR4DFlowNet.py:101-103, is
// materialised once by the producer instead of once per consumer).  value = hi + lo/2048,
// with hi = rn_fp16(x), lo = rn_fp16((x - hi) * 2048): 22 significant bits, the same 4
// bytes per element as fp32, and directly consumable as tcgen05 kind::f16 operands by the
// split-precision tensor-core convolution.
//
// Gradient layout: fp32 channels-last; "G4" buffers [B][D+4]^3[64] carry a two-voxel ZERO
// halo (what conv dgrad wants), "raw" buffers [B][D+2]^3[64] hold dgrad output on the
// padded grid before the halo is folded back (the MirrorPadGrad step).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define SR4D_C 64
#define SR4D_LO_SCALE 2048.0f
#define SR4D_LO_INV (1.0f / 2048.0f)

struct ActView {
    __half* hi;
    __half* lo;
    int B, D;   // D = interior edge; storage edge is D + 2
    // handle-owned device flag, set to 1 by any producer that had to clamp a value to the fp16 range (+-65504) while
    // splitting it (NULL: not tracked).  The fp32 reference would carry such values on; sr4d_activation_overflow()
    // reports it instead of saturating silently.
    unsigned int* ovf = nullptr;
};

__host__ __device__ inline size_t act_plane_elems(int B, int D) {
    size_t dp = (size_t)D + 2;
    return (size_t)B * dp * dp * dp * SR4D_C;
}

// element offset of channel 0 of interior voxel (x,y,z) (each in [-1, D]) of sample b
__device__ __forceinline__ size_t act_off(int D, int b, int x, int y, int z) {
    const int dp = D + 2;
    return ((((size_t)b * dp + (x + 1)) * dp + (y + 1)) * dp + (z + 1)) * SR4D_C;
}

__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    x = fminf(fmaxf(x, -65504.f), 65504.f);
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * SR4D_LO_SCALE);
}
// same, reporting whether the value was outside the representable range (NaN counts)
__device__ __forceinline__ bool split_f16_chk(float x, __half& hi, __half& lo) {
    const bool over = !(fabsf(x) <= 65504.f);
    split_f16(x, hi, lo);
    return over;
}
__device__ __forceinline__ float join_f16(__half hi, __half lo) {
    return fmaf(__half2float(lo), SR4D_LO_INV, __half2float(hi));
}

// 8 consecutive channels
__device__ __forceinline__ void act_load8(const __half* hi, const __half* lo, size_t off, float* v) {
    uint4 h = *reinterpret_cast<const uint4*>(hi + off);
    uint4 l = *reinterpret_cast<const uint4*>(lo + off);
    const __half2* hh = reinterpret_cast<const __half2*>(&h);
    const __half2* ll = reinterpret_cast<const __half2*>(&l);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 a = __half22float2(hh[i]);
        float2 b = __half22float2(ll[i]);
        v[2 * i] = fmaf(b.x, SR4D_LO_INV, a.x);
        v[2 * i + 1] = fmaf(b.y, SR4D_LO_INV, a.y);
    }
}
__device__ __forceinline__ void act_load4(const __half* hi, const __half* lo, size_t off, float* v) {
    uint2 h = *reinterpret_cast<const uint2*>(hi + off);
    uint2 l = *reinterpret_cast<const uint2*>(lo + off);
    const __half2* hh = reinterpret_cast<const __half2*>(&h);
    const __half2* ll = reinterpret_cast<const __half2*>(&l);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float2 a = __half22float2(hh[i]);
        float2 b = __half22float2(ll[i]);
        v[2 * i] = fmaf(b.x, SR4D_LO_INV, a.x);
        v[2 * i + 1] = fmaf(b.y, SR4D_LO_INV, a.y);
    }
}
__device__ __forceinline__ void act_pack4(const float* v, uint2& h, uint2& l, unsigned int* ovf = nullptr) {
    __half hh[4], ll[4];
    bool over = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) over |= split_f16_chk(v[i], hh[i], ll[i]);
    if (over && ovf) *ovf = 1u;
    h = *reinterpret_cast<uint2*>(hh);
    l = *reinterpret_cast<uint2*>(ll);
}

// Store 4 consecutive channels of interior voxel (x,y,z) and replicate them into the
// halo positions this voxel is the clamp image of (faces, edges, corners).
__device__ __forceinline__ void act_store4_halo(__half* hi, __half* lo, int D, int b, int x, int y, int z,
                                                int c, const float* v, bool halo, unsigned int* ovf = nullptr) {
    uint2 h, l;
    act_pack4(v, h, l, ovf);
    if (!halo) {
        size_t o = act_off(D, b, x, y, z) + c;
        *reinterpret_cast<uint2*>(hi + o) = h;
        *reinterpret_cast<uint2*>(lo + o) = l;
        return;
    }
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
                size_t o = act_off(D, b, x + dx, y + dy, z + dz) + c;
                *reinterpret_cast<uint2*>(hi + o) = h;
                *reinterpret_cast<uint2*>(lo + o) = l;
            }
        }
    }
}

// 8 consecutive channels with 16-byte stores (same halo replication rule)
__device__ __forceinline__ void act_store8_halo(__half* hi, __half* lo, int D, int b, int x, int y, int z,
                                                int c, const float* v, bool halo, unsigned int* ovf = nullptr) {
    __align__(16) __half hh[8];
    __align__(16) __half ll[8];
    bool over = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) over |= split_f16_chk(v[i], hh[i], ll[i]);
    if (over && ovf) *ovf = 1u;
    const uint4 H = *reinterpret_cast<const uint4*>(hh), L = *reinterpret_cast<const uint4*>(ll);
    const bool edge = halo && (x == 0 || x == D - 1 || y == 0 || y == D - 1 || z == 0 || z == D - 1);
    if (!edge) {
        const size_t o = act_off(D, b, x, y, z) + c;
        *reinterpret_cast<uint4*>(hi + o) = H;
        *reinterpret_cast<uint4*>(lo + o) = L;
        return;
    }
    for (int dx = -1; dx <= 1; ++dx) {
        if ((dx == -1 && x != 0) || (dx == 1 && x != D - 1)) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            if ((dy == -1 && y != 0) || (dy == 1 && y != D - 1)) continue;
            for (int dz = -1; dz <= 1; ++dz) {
                if ((dz == -1 && z != 0) || (dz == 1 && z != D - 1)) continue;
                const size_t o = act_off(D, b, x + dx, y + dy, z + dz) + c;
                *reinterpret_cast<uint4*>(hi + o) = H;
                *reinterpret_cast<uint4*>(lo + o) = L;
            }
        }
    }
}

__device__ __forceinline__ float act_fn(float v, float slope) { return v > 0.f ? v : v * slope; }
// derivative of the activation expressed through its OUTPUT (ReLU / LeakyReLU keep sign;
// TF's ReluGrad / LeakyReluGrad test "> 0")
__device__ __forceinline__ float act_grad_from_out(float out, float slope) { return out > 0.f ? 1.f : slope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
