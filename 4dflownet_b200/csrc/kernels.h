// Internal (C++) launch interface between the engine and the kernel translation units.
#pragma once
#include "common.cuh"

struct Conv64Args {
    // input: split-fp16 Act planes (in_hi/in_lo) or fp32 (in_f32); buffer edge = Do + 2
    const __half* in_hi = nullptr;
    const __half* in_lo = nullptr;
    const float* in_f32 = nullptr;
    int B = 0;
    int Do = 0;                 // edge of the OUTPUT grid; in[idx + tap] feeds out[idx], tap in {0,1,2}^3
    const float* w = nullptr;   // Keras layout [27][64 ci][64 co], fp32
    int dgrad = 0;              // 1: use W'[tap][co][ci] = W[26-tap][ci][co]
    // output A: Act planes with interior edge Do (storage Do+2), optional replicate halo
    __half* out_hi = nullptr;
    __half* out_lo = nullptr;
    unsigned int* ovf = nullptr;   // fp16-range overflow flag of the output Act (ActView::ovf)
    int halo = 1;
    const float* bias = nullptr;
    const __half* res_hi = nullptr;
    const __half* res_lo = nullptr;
    float slope = 1.f;
    // output B: raw fp32 [B][Do^3][64]
    float* out_raw = nullptr;
};

struct UpsampleTables {   // device arrays of length r*D (forward) and D (backward ranges)
    const int* lo;
    const int* hi;
    const float* lerp;
    const int* ibeg;      // for LR index j: HR indices [ibeg[j], iend[j]) may touch j
    const int* iend;
};

cudaError_t launch_conv64_simt(const Conv64Args& a, cudaStream_t s);

// ---- forward small kernels (fwd_small.cu)
cudaError_t launch_prep_features(const float* u, const float* v, const float* w, const float* um,
                                 const float* vm, const float* wm, float* feat, int B, int P, cudaStream_t s);
// feat [B][P^3][6] = (u,v,w,pcmr,mag,speed); ch0 selects the 3-channel group (0: phase, 3: pc)
cudaError_t launch_stem_convs(const float* feat, const float* w_pc, const float* b_pc, ActView out_pc, const float* w_ph,
                              const float* b_ph, ActView out_ph, cudaStream_t s);
cudaError_t launch_conv1x1_cat(ActView a, ActView b, const float* w, const float* bias, ActView out, cudaStream_t s);
cudaError_t launch_upsample(ActView in, ActView out, int r, UpsampleTables t, cudaStream_t s);
// three 64->1 heads: in[c] -> out[b][voxel][c]
cudaError_t launch_head_out(ActView h0, ActView h1, ActView h2, const float* w0, const float* w1,
                            const float* w2, const float* b0, const float* b1, const float* b2, float* out,
                            cudaStream_t s);
// the same three heads on the tensor cores (head_tc.cu): tap-dot GEMM over the voxel rows of the padded Acts into
// P0..P2 [B (D+2)^3][32] fp32, then the 27-point re-indexed sum; wimg: head_tc_wimg_halves() halves of scratch for the
// weight images (rebuilt per call)
size_t head_tc_wimg_halves();
cudaError_t launch_head_out_tc(ActView h0, ActView h1, ActView h2, const float* w0, const float* w1, const float* w2,
                               const float* b0, const float* b1, const float* b2, __half* wimg, float* P0, float* P1,
                               float* P2, float* out, cudaStream_t s);
// fp32 channels-last (B,D,D,D,64) <-> Act
cudaError_t launch_pack_act(const float* x, ActView out, cudaStream_t s);
cudaError_t launch_unpack_act(ActView in, float* y, cudaStream_t s);

// ---- loss / optimizer (loss_adam.cu)
cudaError_t launch_loss_stats(const float* pred, const float* hu, const float* hv, const float* hw,
                              const float* mask, int B, int nvox, double* partial, int nblk, float* per_sample,
                              float* norm, cudaStream_t s);
// g = d(sum_b loss_b)/d pred; *gmax (optional, device) receives the bit pattern of max |g|
cudaError_t launch_loss_grad(const float* pred, const float* hu, const float* hv, const float* hw,
                             const float* mask, int B, int nvox, const float* norm, float* g, unsigned int* gmax,
                             cudaStream_t s);
cudaError_t launch_sumsq(const float* p, const unsigned char* kflag, int64_t n, double* partial, int nblk,
                         float coeff, float* out, cudaStream_t s);
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, const unsigned char* kflag, int64_t n,
                        float alpha, float beta1, float beta2, float eps, float l2_scale, const float* count_dev,
                        cudaStream_t s);
cudaError_t launch_stitch(const float* pred, int nx, int ny, int nz, int H, int crop, int VX, int VY, int VZ,
                          float venc, int round_small, float* vol, cudaStream_t s);

// ---- backward kernels (bwd.cu)
// whole backward of a 64->1 head conv: h = its saved input Act (B,H), g (B,H^3,3) channel c, w[27][64];
// writes d(pre-activation of h) = relu'(h) * dgrad (clamp padding folded in) into the G4 interior with |max|,
// dw[27][64] and db; scratch >= (592 + 1) * 28 * 64 floats
// db1 (optional, 64 floats) = per-channel sum of the written gradient (the bias gradient of the head's first conv);
// split_out (optional) = scaled split-fp16 copy [2B][H+4]^3[64] of out_g4 with *split_exp derived from *gmax
cudaError_t launch_head2_bwd(ActView h, const float* g, int c, const float* w, float* out_g4, unsigned int* amax,
                             float* dw, float* db, float* db1, __half* split_out, int* split_exp,
                             const unsigned int* gmax, float* scratch, cudaStream_t s, bool hi_only = false);
// the same backward on the tensor cores (head_bwd_tc.cu): setup once per step (scales from *gmax and the weights, weight
// images), one launch per head (writes the scaled split copy -- hi plane, lo plane unless hi_only -- its exponent and
// |max|; no fp32 copy), one finishing launch for the three heads (dw[27][64], db, db1[64] from the per-CTA partials)
bool head_bwd_tc_supported(int D);
int head_bwd_tc_grid(int B, int D);
size_t head_bwd_tc_partial_floats(int B, int D);
size_t head_bwd_tc_scales_bytes();
size_t head_bwd_tc_wimg_halves();
size_t head_bwd_tc_gplanar_floats(int B, int D);
// g (B, D^3, 3): the loss gradient; gplanar: head_bwd_tc_gplanar_floats() floats of scratch (planar scaled copy)
cudaError_t launch_head_bwd_tc_setup(const float* w0, const float* w1, const float* w2, const unsigned int* gmax, void* scales,
                                     __half* wimg, const float* g, float* gplanar, int B, int D, cudaStream_t s);
cudaError_t launch_head_bwd_tc(ActView h, const float* gplanar, int c, const void* scales, const __half* wimg, float* partial,
                               __half* split_out, int* split_exp, unsigned int* amax, bool hi_only, cudaStream_t s);
cudaError_t launch_head_bwd_tc_finish(const float* const part[3], const int ncta[3], const void* scales, float* const dw[3],
                                      float* const db[3], float* const db1[3], bool hi_only, cudaStream_t s);
// *out_bits = max(*out_bits, bits of max |x|)  (zero it first)
cudaError_t launch_absmax(const float* x, size_t n, unsigned int* out_bits, cudaStream_t s);
// scaled split-fp16 gradient (G4 layout) -> dense fp32 (B, D, D, D, 64)
cudaError_t launch_dense_from_split(const __half* split, const int* exp, bool lo_valid, float* dense, int B, int D,
                                    cudaStream_t s);
// out(G4 interior) = (fold(raw0*2^-e0 + raw1*2^-e1 + raw2*2^-e2) + add) * act'(saved); e_i are device
// exponents (NULL = 0) of the scaled split-fp16 gradients the raws were computed from; amax (optional)
// receives atomicMax of |out|
cudaError_t launch_fold_act(const float* raw0, const float* raw1, const float* raw2, const int* e0, const int* e1,
                            const int* e2, const float* add_g4, const __half* saved_hi, const __half* saved_lo,
                            float slope, float* out_g4, unsigned int* amax, int B, int D, cudaStream_t s);
// split-fp16 copy [2B][D+4]^3[64] (hi planes then lo planes, zero halo kept) of a fp32 G4 tensor, scaled by
// 2^e with e derived from *amax so max|x|*2^e is in [2^13, 2^14); *exp_out = e
// hi_only: the lo plane has no consumer (single-plane dgrad + wgrad): it is not written
cudaError_t launch_g4_split(const float* g4, const unsigned int* amax, __half* split, int* exp_out, int B, int D,
                            cudaStream_t s, bool hi_only = false);
// dW[27][64][64] (+= nothing; overwrite) from x Act and dy G4; scratch >= nchunk*27*64*64 floats
cudaError_t launch_wgrad64_simt(ActView x, const float* dy_g4, float* dw, float* scratch, int nchunk,
                                cudaStream_t s);
// out[j] = sum_r partial[r][j] (deterministic second stage of the split reductions)
cudaError_t launch_reduce_rows(const float* partial, int nrows, int ncols, float* out, cudaStream_t s);
// a batch of such reductions in one launch: item i sums nrows0 (cls 0) or nrows1 (cls 1) rows of its partial matrix
struct ReduceItem { const float* partial; float* out; int cls; int pad; };
cudaError_t launch_reduce_rows_batched(const ReduceItem* items_dev, int nitems, int nrows0, int nrows1, int ncols, cudaStream_t s);
cudaError_t launch_bias_grad(const float* dy_g4, int B, int D, float* db, float* scratch, cudaStream_t s);
cudaError_t launch_upsample_bwd(const float* dhr_g4, ActView lr_saved, float slope, float* dlr_g4,
                                unsigned int* amax, int B, int D, int r, UpsampleTables t, cudaStream_t s);
// 1x1 conv backward: dy G4 (B,D) ; a,b saved inputs (phase, pc); outputs: da,db G4 interiors already
// multiplied by relu'(a), relu'(b); dw[128][64], dbias[64]
cudaError_t launch_conv1x1_bwd(const float* dy_g4, ActView a, ActView b, const float* w, float* da_g4,
                               float* db_g4, unsigned int* amax_a, unsigned int* amax_b, float* dw, float* dbias,
                               float* scratch, cudaStream_t s);
// stem 3->64 weight gradient: feat [B][P^3][6] group ch0, dy G4 -> dw[27][3][64], db[64]
cudaError_t launch_stem_wgrad(const float* feat, int ch0, const float* dy_g4, int B, int P, float* dw,
                              float* db, float* scratch, cudaStream_t s);
cudaError_t launch_g4_from_dense(const float* dense, float* g4, unsigned int* amax, int B, int D, cudaStream_t s);
cudaError_t launch_dense_from_g4(const float* g4, float* dense, int B, int D, cudaStream_t s);
