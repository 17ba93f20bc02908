// tcgen05 (5th-gen tensor core) 64->64 3x3x3 convolution: interface used by the engine.
#pragma once
#include "common.cuh"

struct TcWeights;   // per-layer tensor-core operand images (split fp16, swizzled), device resident

struct TcConvArgs {
    ActView in;                 // interior edge D (storage D+2, replicate halo filled)
    ActView out;                // interior edge D
    int layer = 0;              // index into TcWeights
    int dgrad = 0;
    const float* bias = nullptr;
    const __half* res_hi = nullptr;
    const __half* res_lo = nullptr;
    float slope = 1.f;
    int halo = 1;
    float* out_raw = nullptr;          // fp32 [B][D^3][64] output (no bias/activation/halo) instead of `out`
    unsigned int* absmax = nullptr;    // with out_raw / fused: atomicMax of the |value| bit patterns written
    // fused dgrad (dgrad = 1, tc_dgrad_fusable(D)): `in` is the scaled split gradient viewed with interior edge
    // D+2; the halo fold, the skip add, the activation derivative and the accumulation happen in the epilogue:
    //   out_g4(interior) = (fold(dgrad) * 2^-(*dy_exp) + add_pre) * act'(saved; slope) + add_post
    int fused = 0;
    // dgrad only: use just the hi plane of the split gradient (one scaled fp16 value per element) -- one
    // [Wlo;Whi] x dYhi instruction per K-step instead of two.  Weights stay split; see SR4D_OPT_DGRAD_SINGLE.
    int single_b = 0;
    const int* dy_exp = nullptr;
    const float* add_pre = nullptr;    // fp32 G4 [B][D+4]^3[64] or NULL
    const float* add_post = nullptr;   // fp32 G4 or NULL (may alias out_g4)
    const __half* sav_hi = nullptr;    // saved activation planes (edge D) or NULL (no activation)
    const __half* sav_lo = nullptr;
    float* out_g4 = nullptr;
    // optional: also write the scaled split-fp16 copy [2B][D+4]^3[64] of out_g4 (what g4_split_kernel produces) with
    // an exponent derived in-kernel from max|dy| (*dy_amax), the layer's weight gain and max|add_pre| (*add_amax)
    __half* split_out = nullptr;
    int split_hi_only = 0;             // the lo plane of split_out has no consumer (single-plane dgrad + wgrad): skip it
    int* split_exp = nullptr;
    const unsigned int* dy_amax = nullptr;
    const unsigned int* add_amax = nullptr;
};

cudaError_t tc_alloc_weights(TcWeights** w, int nlayers);
void tc_free_weights(TcWeights* w);
// (re)build the operand images (forward + dgrad) of n layers in one launch: entry i is image slot layers[i],
// built from the Keras-layout fp32 kernel [27][64][64] at params + offsets[i]
cudaError_t tc_prepare_weights(TcWeights* w, const float* params, const int* layers, const long long* offsets, int n,
                               cudaStream_t s);
cudaError_t tc_conv64(TcWeights* w, const TcConvArgs& a, cudaStream_t s);
// A chain of forward layers on one grid (same batch and edge, Act -> Act, each with its own weights / bias / residual /
// slope) as ONE cooperative launch: the persistent CTAs pass a grid-wide barrier between layers instead of the launch
// boundary.  Built once per (buffers, batch) -- the parameter blocks and tensor maps live in device memory.
struct TcChain;
long tc_fwd_tiles(int Do, int B);      // tiles of one forward layer on this grid (persistent grid = min(tiles, SMs))
cudaError_t tc_chain_build(TcWeights* w, const TcConvArgs* layers, int n, TcChain** out);
cudaError_t tc_chain_launch(TcChain* c, cudaStream_t s);
int tc_chain_layers(const TcChain* c);
void tc_chain_free(TcChain* c);
// true when the fused dgrad's in-epilogue fold works for interior edge D (the halo line / column pairs
// (0,1) and (D,D+1) of the padded grid fall into one epilogue chunk and one z tile)
bool tc_dgrad_fusable(int D);

// tcgen05 weight gradient (wgrad_tc.cu): x = the layer's saved input Act (edge D), dy_split = the scaled
// split-fp16 gradient [2B][D+4]^3[64] with device exponent *dy_exp; writes tc_wgrad_slabs(B, D) partial
// dW[27][64][64] into `partial` (slab-major) for a row reduction.
int tc_wgrad_slabs(int B, int D);
// single = 1: dYhi x Xhi only, same kernel with the lo planes and the second instruction left out (kept as the
// cross-check of tc_wgrad64_single).
cudaError_t tc_wgrad64(ActView x, const __half* dy_split, const int* dy_exp, float* partial, cudaStream_t s,
                       int single = 0);
// single-operand weight gradient (wgrad_tc2.cu, SR4D_OPT_WGRAD_SINGLE = 1, the default): two x-planes of dYhi stacked on
// M against Xhi, operands re-used out of shared memory across x, accumulator chains flushed to fp32 every 384
// accumulations; writes tc_wgrad2_slabs(B, D) partial dW[27][64][64] (slab-major) for a row reduction.
int tc_wgrad2_slabs(int B, int D);
cudaError_t tc_wgrad64_single(ActView x, const __half* dy_split, const int* dy_exp, float* partial, cudaStream_t s);
// The weight gradients of several layers on the same grid (same batch and edge) in ONE launch of the stacked kernel: item i
// reads the saved input x_i and the split gradient dy_split_i (device exponent dy_exp_i) and writes its slab partials to
// partial_i.  Built once per set of buffers (tensor maps in device memory), launched once per backward pass.
struct TcWgradItem { ActView x; const __half* dy_split; const int* dy_exp; float* partial; };
struct TcWgradBatch;
cudaError_t tc_wgrad_batch_build(const TcWgradItem* items, int n, TcWgradBatch** out);
cudaError_t tc_wgrad_batch_launch(TcWgradBatch* b, cudaStream_t s);
void tc_wgrad_batch_free(TcWgradBatch* b);
