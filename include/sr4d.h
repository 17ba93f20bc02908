/*
 * sr4d.h -- C ABI of libsr4d.so, the B200 (sm_100a) engine behind the patch-based
 * super-resolution hot path of 4DFlowNet.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Each entry point cites the reference interface it replaces (paths relative to
 * /root/reference/src).  The Python mirror of the reference's classes
 * (4dflownet_b200/Network/*.py) binds these with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative SR4D_E* code; nothing throws
 *    across the ABI; sr4d_last_error() gives the text for the calling handle.
 *  - all tensor arguments are DEVICE pointers unless the name ends in _host; the caller
 *    owns inputs/outputs, the handle owns weights, optimizer state and workspace.
 *  - work is enqueued on the cudaStream_t passed as `void* stream` (NULL = legacy
 *    default stream); no implicit synchronisation unless stated.
 *  - one handle per GPU per process; a handle is not thread safe.
 *  - there is NO CPU fallback: creation fails with SR4D_ENODEVICE unless the current
 *    device is compute capability 10.x.
 *  - layouts are the reference's: inputs (B,P,P,P) fp32 contiguous (the trailing
 *    singleton channel of the Keras inputs dropped), output (B,rP,rP,rP,3) fp32,
 *    kernels in Keras order (kx,ky,kz,Cin,Cout).
 */
#ifndef SR4D_H_
#define SR4D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR4D_OK            0
#define SR4D_EINVAL       -1   /* bad argument */
#define SR4D_ENODEVICE    -2   /* no sm_100 device / CUDA runtime failure at create */
#define SR4D_ECUDA        -3   /* a CUDA call or kernel launch failed */
#define SR4D_ENOMEM       -4   /* device allocation failed */
#define SR4D_ESTATE       -5   /* call sequence error (e.g. adam before fwd_bwd) */

/* conv implementation selector for sr4d_set_option(h, SR4D_OPT_CONV_IMPL, v) */
#define SR4D_OPT_CONV_IMPL   1
#define SR4D_CONV_AUTO       0   /* tcgen05 tensor-core kernel where it applies, SIMT elsewhere */
#define SR4D_CONV_SIMT       1   /* fp32 CUDA-core kernels everywhere (correctness anchor) */
#define SR4D_CONV_TCGEN05    2   /* force the tcgen05 split-fp16 kernel for every 64->64 conv */
#define SR4D_OPT_SAVE_ACTS   2   /* 1: forward keeps every activation (needed before backward) */
#define SR4D_OPT_PROFILE     3   /* 1: bracket every 64->64 conv launch with CUDA events on its stream */
#define SR4D_OPT_FUSED_DGRAD 4   /* 1 (default): the tensor-core dgrad folds the clamp-padding halo, adds the skip
                                    gradient and applies the activation derivative in its epilogue where the grid
                                    allows; 0: separate raw dgrad + fold kernel */
#define SR4D_OPT_FWD_CHAIN   8   /* 0..64 (default 4): runs of consecutive 64->64 forward layers go out as ONE cooperative launch with a
                                    grid-wide barrier between layers when a layer has at most this many tiles per SM (batch-1 /
                                    small-grid inference); 0 = one launch per layer.  The arithmetic is the same: outputs are
                                    bit-identical either way. */
#define SR4D_OPT_NVTX        7   /* 1: NVTX ranges around the phases of a step ("sr4d forward", "sr4d loss + backward",
                                    "sr4d adam") and around every 64->64 layer call, named after its kernel class
                                    (conv64_fwd_hr, ...); default 0 */
#define SR4D_OPT_DGRAD_SINGLE 5  /* 1 (default since round 2, validated on B200: profiles/r02_grad_parity.txt): the tensor-core
                                    dgrad consumes only the hi plane of the power-of-two-scaled split gradient (one fp16
                                    value per element; the weights stay split): one MMA per K-step instead of two.  Its
                                    rounding errors are independent per element and average out in the weight-gradient
                                    sums: flat gradient vs float64 5.97e-5 with it, 6.66e-5 without, at P=24 r=2 8/4 B=8.
                                    0: both planes (round-1 kernel path). */
#define SR4D_OPT_WGRAD_SINGLE 6  /* 1 (default since round 2): stacked single-plane weight-gradient kernel (wgrad_tc2.cu): hi
                                    planes of dY and X, two x-planes of dY on the M rows, accumulator chains cut every 384
                                    accumulations.  0: two-plane kernel (wgrad_tc.cu, slabs cut at the same chain length);
                                    2: the two-plane kernel's own hi-only mode (cross-check). */

/* kernel classes timed under SR4D_OPT_PROFILE (index into sr4d_profile_read's arrays):
 * the 64->64 3x3x3 convolution forward / input-gradient / weight-gradient, on the LR
 * (patch_size^3) and HR ((patch_size*res_increase)^3) grids */
#define SR4D_PROF_CONV64_FWD_LR    0
#define SR4D_PROF_CONV64_FWD_HR    1
#define SR4D_PROF_CONV64_DGRAD_LR  2
#define SR4D_PROF_CONV64_DGRAD_HR  3
#define SR4D_PROF_CONV64_WGRAD_LR  4
#define SR4D_PROF_CONV64_WGRAD_HR  5
#define SR4D_PROF_NCLASSES         6

typedef struct sr4d_handle sr4d_t;

typedef struct sr4d_tensor_desc {
    char    name[32];     /* "conv3d_7/kernel", "conv3d_30/bias", ... (Keras creation order) */
    int64_t offset;       /* in floats, into the flat param/grad/m/v buffers (128-byte aligned) */
    int64_t count;        /* number of floats */
    int32_t ndim;
    int32_t shape[5];     /* kernel: (kx,ky,kz,Cin,Cout); bias: (Cout) */
    int32_t is_kernel;    /* 1: carries the L2 regulariser (SR4DFlowNet.py:99,104) */
} sr4d_tensor_desc;

/* ---- life cycle -------------------------------------------------------------------
 * Replaces predictor.py:11-29 prepare_network(patch_size,res_increase,low_resblock,
 * hi_resblock) and the model construction in TrainerController.py:35-49
 * (SR4DFlowNet(res_increase).build_network(...), SR4DFlowNet.py:4-51).
 * max_batch bounds the workspace; training != 0 also allocates gradient / Adam state. */
int  sr4d_create(sr4d_t** h, int patch_size, int res_increase, int low_resblock,
                 int hi_resblock, int max_batch, int training, int device);
void sr4d_destroy(sr4d_t* h);
const char* sr4d_last_error(const sr4d_t* h);
const char* sr4d_version(void);
int  sr4d_set_option(sr4d_t* h, int option, int value);
int  sr4d_get_option(const sr4d_t* h, int option, int* value);

/* ---- parameters --------------------------------------------------------------------
 * model.trainable_variables / get_weights / set_weights / load_weights
 * (TrainerController.py:223,394; predictor.py:61) and optimizer.weights
 * (TrainerController.py:359,391). */
int64_t sr4d_param_count(const sr4d_t* h);        /* sum of tensor sizes (3342083 for 8/4) */
int64_t sr4d_flat_size(const sr4d_t* h);          /* floats in the padded flat buffers */
int     sr4d_num_tensors(const sr4d_t* h);        /* 48 for 8/4 */
int     sr4d_param_table(const sr4d_t* h, sr4d_tensor_desc* out, int capacity);
float*  sr4d_params(sr4d_t* h);                   /* borrowed device pointers, flat_size floats */
float*  sr4d_grads(sr4d_t* h);                    /* NULL unless training; sr4d_grads_size() floats */
/* The gradient buffer is flat_size floats of gradients followed by a SR4D_METRIC_TAIL-float "metric tail" the
 * engine never touches: the data-parallel caller points per_sample / l2_out of sr4d_train_fwd_bwd into its
 * rank's slot of the tail, so ONE all-reduce(SUM) of the whole buffer moves the gradients and gathers the
 * per-sample metrics and the global sample count ("13.37 MB + metric tail"). */
#define SR4D_METRIC_TAIL 4096
int64_t sr4d_grads_size(const sr4d_t* h);         /* flat_size + SR4D_METRIC_TAIL (0 unless training) */
float*  sr4d_adam_m(sr4d_t* h);
float*  sr4d_adam_v(sr4d_t* h);
/* must be called after the caller wrote into sr4d_params() so derived weight images
 * (tensor-core operand layouts) are rebuilt; sr4d_adam_step does it itself. */
int  sr4d_params_changed(sr4d_t* h, void* stream);

/* ---- forward: model.predict / model(inputs) ------------------------------------------
 * predictor.py:87-92, TrainerController.py:217,234,426.  u..w_mag: (B,P,P,P) fp32;
 * out: (B,rP,rP,rP,3) fp32.  B <= max_batch. */
int  sr4d_forward(sr4d_t* h, const float* u, const float* v, const float* w,
                  const float* u_mag, const float* v_mag, const float* w_mag,
                  float* out, int B, void* stream);

/* ---- loss + metric: TrainerController.loss_function / accuracy_function -------------
 * TrainerController.py:84-127,143-156 and loss_utils.py:64-103.  pred (B,H,H,H,3),
 * hr_u/v/w (B,H,H,H), mask (B,H,H,H); per_sample (B,4) = {loss (no l2), mse,
 * rel_err_percent, sum(mask)}. */
int  sr4d_loss_metrics(sr4d_t* h, const float* pred, const float* hr_u, const float* hr_v,
                       const float* hr_w, const float* mask, int B, float* per_sample,
                       void* stream);

/* ---- train step pieces: TrainerController.train_step (TrainerController.py:209-225) --
 * forward + loss + backward.  Leaves SUM_b grad(loss_b) (sum, not mean; no L2 term) in
 * sr4d_grads(); per_sample as above; l2_out (1 float, device) = 5e-7 * sum(kernel^2)
 * (TrainerController.py:129-141).  pred_out may be NULL. */
int  sr4d_train_fwd_bwd(sr4d_t* h, const float* u, const float* v, const float* w,
                        const float* u_mag, const float* v_mag, const float* w_mag,
                        const float* hr_u, const float* hr_v, const float* hr_w,
                        const float* mask, int B, float* per_sample, float* l2_out,
                        float* pred_out, void* stream);

/* The two halves of sr4d_train_fwd_bwd as separate calls: `with tf.GradientTape() as tape: predictions =
 * self.model(inputs, training=True)` (TrainerController.py:213-217) and `loss ...; tape.gradient`
 * (TrainerController.py:218-223).  sr4d_train_backward differentiates the activations the last
 * sr4d_train_forward of the same batch saved; options (e.g. SR4D_OPT_CONV_IMPL) may change in between, which
 * is how the parity tests feed the tensor-core backward the fp32 SIMT forward's activations. */
int  sr4d_train_forward(sr4d_t* h, const float* u, const float* v, const float* w,
                        const float* u_mag, const float* v_mag, const float* w_mag, int B,
                        float* pred_out, void* stream);
int  sr4d_train_backward(sr4d_t* h, const float* hr_u, const float* hr_v, const float* hr_w,
                         const float* mask, int B, float* per_sample, float* l2_out, void* stream);

/* Keras Adam (TrainerController.py:73,225): t = iterations+1;
 * g = grad + l2_grad_scale * w on kernels only (l2_grad_scale = B_global * 2 * 5e-7, the
 * gradient of the regulariser added to each of the B_global loss entries,
 * TrainerController.py:249). */
int  sr4d_adam_step(sr4d_t* h, float lr, float beta1, float beta2, float eps, int64_t t,
                    float l2_grad_scale, void* stream);

/* Same step with the regulariser scale taken from the device: g = grad + l2_grad_per_sample * tail[tail_index] * w,
 * where tail[tail_index] is a float in the metric tail of sr4d_grads() (the all-reduced global batch size), so
 * ragged data-parallel shards need no host synchronisation between the all-reduce and Adam. */
int  sr4d_adam_step_counted(sr4d_t* h, float lr, float beta1, float beta2, float eps, int64_t t,
                            float l2_grad_per_sample, int tail_index, void* stream);

/* ---- inference post-processing: PatchGenerator._patchup_with_overlap + predictor.py:99-107
 * Crops side_pad*r voxels per side of every predicted patch, scatters component c into
 * the stitched volume (nx*core, ny*core, nz*core) cropped to vol (VX,VY,VZ), multiplies
 * by venc and zeroes |v| < venc/2048 when round_small != 0.  pred (N,H,H,H,3) with N =
 * nx*ny*nz patches in x-major order; vol_out (3,VX,VY,VZ) fp32. */
int  sr4d_stitch(sr4d_t* h, const float* pred, int nx, int ny, int nz, int side_pad_hr,
                 int VX, int VY, int VZ, float venc, int round_small, float* vol_out,
                 void* stream);

/* ---- single-layer entry points (used by the parity tests and by ncu captures) --------
 * conv3d() of SR4DFlowNet.py:93-108 for Cin=Cout=64,k=3 on fp32 channels-last tensors:
 * x (B,D,D,D,64), kernel (3,3,3,64,64), bias (64) or NULL, residual (B,D,D,D,64) or NULL
 * (added before the activation, SR4DFlowNet.py:117), act_slope: 1 = linear, 0 = ReLU,
 * 0.2 = LeakyReLU.  impl = SR4D_CONV_SIMT / SR4D_CONV_TCGEN05. */
int  sr4d_conv64_layer(sr4d_t* h, const float* x, const float* kernel, const float* bias,
                       const float* residual, float act_slope, float* y, int B, int D,
                       int impl, void* stream);
/* upsample3d() of SR4DFlowNet.py:53-90: x (B,D,D,D,64) -> y (B,rD,rD,rD,64). */
int  sr4d_upsample_layer(sr4d_t* h, const float* x, float* y, int B, int D, int r, void* stream);
/* gradient of sr4d_conv64_layer wrt x and kernel for a given dy (B,D,D,D,64) (pre-activation
 * gradient): dx (B,D,D,D,64), dkernel (3,3,3,64,64), dbias (64).  Any output may be NULL. */
int  sr4d_conv64_layer_bwd(sr4d_t* h, const float* x, const float* kernel, const float* dy,
                           float* dx, float* dkernel, float* dbias, int B, int D, int impl,
                           void* stream);

/* whole backward of one 64->1 head convolution (SR4DFlowNet.py:40,43,46: relu -> conv3d(1 filter, 'SYMMETRIC' padding))
 * given its saved post-ReLU input x (B,D,D,D,64), kernel (3,3,3,64,1) and the loss gradient g (B,D,D,D,3) of which
 * channel c belongs to this head: dx (B,D,D,D,64) = gradient wrt the PRE-activation of x (ReluGrad applied),
 * dkernel (27*64), dbias (1), dbias_prev (64) = per-channel sum of dx.  With SR4D_CONV_TCGEN05 dx is what the
 * tensor-core consumers read: the scaled split-fp16 copy converted back to fp32 -- both planes and fp32-accurate
 * weight / bias gradients when SR4D_OPT_DGRAD_SINGLE = SR4D_OPT_WGRAD_SINGLE = 0, the hi plane and single-plane
 * arithmetic (the training default) otherwise. */
int  sr4d_head_layer_bwd(sr4d_t* h, const float* x, const float* kernel, const float* g, int c, float* dx,
                         float* dkernel, float* dbias, float* dbias_prev, int B, int D, int impl, void* stream);

/* device time (ms, CUDA events on the launch stream) and launch count per kernel class since
 * the last read; synchronises with the last recorded event.  Used by bench.py's roofline. */
int  sr4d_profile_read(sr4d_t* h, double* ms, int64_t* launches, int nclasses);

/* Activation range.  Feature maps are stored as split fp16 pairs (hi + lo/2048: 22 significant bits), whose range is
 * +-65504; the fp32 reference has no such limit.  A producer that has to clamp a value (or meets a NaN) sets a
 * handle-owned device flag instead of saturating silently.  *flag = 1 when that happened since the last reset
 * (synchronises `stream`).  Venc-normalised inputs keep activations O(1); the flag fires when training diverges. */
int  sr4d_activation_overflow(sr4d_t* h, int* flag, int reset, void* stream);

/* counters for bench.py: number of kernels this library launched since the last reset */
int64_t sr4d_launch_count(const sr4d_t* h);
void    sr4d_reset_launch_count(sr4d_t* h);

#ifdef __cplusplus
}
#endif
#endif /* SR4D_H_ */
