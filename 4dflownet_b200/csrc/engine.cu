// libsr4d engine: owns parameters, optimizer state and workspace; sequences the kernels
// of the 4DFlowNet SR graph (Network/SR4DFlowNet.py:7-51) forward and backward, and
// implements the C ABI of include/sr4d.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/sr4d.h"
#include "kernels.h"
#include "tc_host.h"
#include "conv_tc.h"

namespace {

struct Layer {
    int64_t w_off = 0, b_off = -1;
    int k = 3, cin = 64, cout = 64;
};

struct ActBuf {
    __half* base = nullptr;   // hi plane then lo plane, sized for maxB
    int D = 0;
    unsigned int* ovf = nullptr;   // the handle's fp16-range overflow flag (see ActView::ovf)
    ActView view(int B) const {
        ActView v;
        v.hi = base;
        v.lo = base + act_plane_elems(B, D);   // planes packed for the CURRENT batch
        v.B = B;
        v.D = D;
        v.ovf = ovf;
        return v;
    }
};

struct GBuf {
    float* f = nullptr;            // fp32 [B][D+4]^3[64], zero halo
    __half* s = nullptr;           // split-fp16 copy scaled by 2^(*exp): [2B][D+4]^3[64]
    unsigned int* amax = nullptr;  // device: max |f| (bit pattern)
    int* exp = nullptr;            // device: exponent of the split copy
    // s / exp point at the buffer's own storage (s0 / exp0) or, with the batched weight gradient, at the per-layer storage
    // of the layer whose weight gradient will read this tensor at the end of the backward pass
    __half* s0 = nullptr;
    int* exp0 = nullptr;
};
struct RawBuf {
    float* p = nullptr;            // fp32 [B][D+2]^3[64]
    const int* exp = nullptr;      // device exponent the values are scaled by (NULL: unscaled)
};

struct UpTab {
    int* lo = nullptr; int* hi = nullptr; float* lerp = nullptr; int* ibeg = nullptr; int* iend = nullptr;
    UpsampleTables tables() const { return UpsampleTables{lo, hi, lerp, ibeg, iend}; }
};

}  // namespace

struct sr4d_handle {
    int P = 0, r = 1, low = 0, hi = 0, maxB = 0, training = 0, device = 0, H = 0;
    std::vector<sr4d_tensor_desc> table;
    std::vector<Layer> layers;
    int64_t flat = 0, nparam = 0;
    float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr;
    unsigned char* kflag = nullptr;
    unsigned int* ovf = nullptr;       // device: set to 1 when an activation was clamped to the fp16 range
    float* feat = nullptr;
    float* tapP[3] = {nullptr, nullptr, nullptr};   // tap dot-products of the three heads [maxB (H+2)^3][32] (tensor-core heads)
    __half* head_wimg = nullptr;                    // their weight images
    // tensor-core backward of the heads (head_bwd_tc.cu): per-head scales, weight images, per-CTA partials
    void* hb_scales = nullptr;
    __half* hb_wimg = nullptr;
    float* hb_part = nullptr;
    float* hb_gplanar = nullptr;                    // planar scaled copy of the loss gradient
    // chained forward launches (conv_tc.h: TcChain), built on first use per (grid, batch); keys: kind * 65536 + B
    std::map<int, TcChain*> chains;
    // deferred second stage of the stacked weight gradient: every 64->64 layer keeps its slab partials until one batched
    // reduction at the end of the backward pass
    float* wg_part = nullptr;
    size_t wg_stride = 0;                           // floats per layer
    ReduceItem* wg_items = nullptr;                 // device, one per 64->64 layer (LR layers: cls 0, HR: cls 1)
    std::vector<int> wg_slot;                       // layer -> item index (-1: not a 64->64 layer)
    int wg_nitems = 0;
    bool wg_pending = false;                        // partials written since the last batched reduction
    // batched weight gradient: every 64->64 layer's dY split copy lives in its own buffer until ONE launch per grid at the
    // end of the backward pass computes all weight gradients (tc_wgrad_batch_*)
    std::vector<__half*> lsplit;                    // [layers]; NULL for other layers
    int* lexp = nullptr;                            // device [layers]
    std::vector<TcWgradItem> wg_list[2];            // collected during a pass: LR grid, HR grid
    std::map<long long, TcWgradBatch*> wg_batches;  // (grid class, batch, items) -> device layer table
    size_t hb_part_stride = 0;                      // floats per head
    std::vector<ActBuf> lr, hr;        // storage slots
    std::vector<int> lr_slot, hr_slot; // tensor index -> slot
    int n_lr_t = 0, n_hr_t = 0;
    UpTab up;
    // training workspace: gradient tensors (fp32 G4 + scaled split-fp16 copy for the tensor cores) and
    // dgrad outputs on the padded grid (raw, carrying the scale exponent of the gradient they came from)
    GBuf g4_lr[4];
    GBuf g4_hr[3];
    RawBuf raw_lr;
    RawBuf raw_hr[3];
    int* gmeta = nullptr;      // device: {absmax bits, exponent} per GBuf, then head_exp[2]
    int* head_exp = nullptr;
    unsigned int* gmax = nullptr;       // device: max |d loss / d pred|
    unsigned int* amax_copy = nullptr;  // device: stable copy of an accumulating tensor's |max|
    float* pred = nullptr;     // (maxB,H^3,3)
    float* gpred = nullptr;    // (maxB,H^3,3)
    float* scratch = nullptr;  // split-reduction partials
    size_t scratch_floats = 0;
    double* dpartial = nullptr;
    float* norm = nullptr;
    float* per_sample_int = nullptr;
    int wgrad_chunks = 64;
    TcWeights* tcw = nullptr;  // tensor-core operand images of the 64->64 kernels
    bool tcw_dirty = true;
    int conv_impl = SR4D_CONV_AUTO;
    int save_acts = 0;
    int fwd_chain = -1;               // SR4D_OPT_FWD_CHAIN: tiles per SM up to which forward runs are chained (-1: default / environment)
    int fused_dgrad = 1;       // SR4D_OPT_FUSED_DGRAD
    int dgrad_single = 1;      // SR4D_OPT_DGRAD_SINGLE
    int wgrad_single = 1;      // SR4D_OPT_WGRAD_SINGLE
    bool have_fwd_state = false;
    int fwd_batch = 0;         // batch of the forward whose activations are saved
    int64_t launches = 0;
    // per-kernel-class device timing (SR4D_OPT_PROFILE): event pairs on the launch stream
    int profile = 0;
    int nvtx = 0;              // SR4D_OPT_NVTX
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<int> ev_class;        // class of pair i (events 2i, 2i+1)
    std::vector<int> ev_weight;       // layer calls the pair covers (a chained launch: all its layers)
    std::string err;
};

namespace {

#define CK(h, expr, nk)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                   \
            return SR4D_ECUDA;                                                                \
        }                                                                                     \
        (h)->launches += (nk);                                                                \
    } while (0)

int fail(sr4d_t* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

const char* const kProfNames[SR4D_PROF_NCLASSES] = {"conv64_fwd_lr", "conv64_fwd_hr", "conv64_dgrad_lr", "conv64_dgrad_hr",
                                                    "conv64_wgrad_lr", "conv64_wgrad_hr"};

// Brackets the launches of one 64->64 layer call: CUDA events per kernel class (SR4D_OPT_PROFILE) and / or an NVTX
// range named after the class (SR4D_OPT_NVTX; the ranges show up in Nsight Systems / ncu --nvtx next to the kernels).
struct ProfScope {
    sr4d_t* h; cudaStream_t s; bool on; bool nvtx;
    ProfScope(sr4d_t* h_, int cls, cudaStream_t s_, int nlayers = 1) : h(h_), s(s_), on(h_->profile != 0), nvtx(h_->nvtx != 0) {
        if (nvtx) nvtxRangePushA(kProfNames[cls]);
        if (!on) return;
        if (h->ev_used + 2 > h->ev_pool.size()) {
            for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
        }
        h->ev_class.push_back(cls);
        h->ev_weight.push_back(nlayers);
        cudaEventRecord(h->ev_pool[h->ev_used], s);
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(h->ev_pool[h->ev_used + 1], s);
            h->ev_used += 2;
        }
        if (nvtx) nvtxRangePop();
    }
};
// NVTX range around a whole phase of a step (forward / loss / backward / adam)
struct NvtxPhase {
    bool on;
    NvtxPhase(const sr4d_t* h, const char* name) : on(h->nvtx != 0) { if (on) nvtxRangePushA(name); }
    ~NvtxPhase() { if (on) nvtxRangePop(); }
};

void build_table(sr4d_t* h) {
    const int C = 64;
    struct L { int k, cin, cout; bool bias; };
    std::vector<L> ls;
    ls.push_back({3, 3, C, true});
    ls.push_back({3, C, C, true});
    ls.push_back({3, 3, C, true});
    ls.push_back({3, C, C, true});
    ls.push_back({1, 2 * C, C, true});
    ls.push_back({3, C, C, true});
    for (int i = 0; i < 2 * (h->low + h->hi); ++i) ls.push_back({3, C, C, false});
    for (int i = 0; i < 3; ++i) {
        ls.push_back({3, C, C, true});
        ls.push_back({3, C, 1, true});
    }
    int64_t off = 0;
    h->nparam = 0;
    for (size_t i = 0; i < ls.size(); ++i) {
        Layer ly;
        ly.k = ls[i].k; ly.cin = ls[i].cin; ly.cout = ls[i].cout;
        char lname[24];
        if (i == 0) snprintf(lname, sizeof lname, "conv3d");
        else snprintf(lname, sizeof lname, "conv3d_%zu", i);
        sr4d_tensor_desc d;
        memset(&d, 0, sizeof d);
        snprintf(d.name, sizeof d.name, "%s/kernel", lname);
        d.offset = off;
        d.count = (int64_t)ly.k * ly.k * ly.k * ly.cin * ly.cout;
        d.ndim = 5;
        d.shape[0] = d.shape[1] = d.shape[2] = ly.k; d.shape[3] = ly.cin; d.shape[4] = ly.cout;
        d.is_kernel = 1;
        ly.w_off = off;
        h->table.push_back(d);
        h->nparam += d.count;
        off += (d.count + 31) / 32 * 32;
        if (ls[i].bias) {
            sr4d_tensor_desc b;
            memset(&b, 0, sizeof b);
            snprintf(b.name, sizeof b.name, "%s/bias", lname);
            b.offset = off;
            b.count = ly.cout;
            b.ndim = 1;
            b.shape[0] = ly.cout;
            b.is_kernel = 0;
            ly.b_off = off;
            h->table.push_back(b);
            h->nparam += b.count;
            off += (b.count + 31) / 32 * 32;
        }
        h->layers.push_back(ly);
    }
    h->flat = off;
}

template <typename T>
cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

int build_upsample_tables(sr4d_t* h) {
    const int in = h->P, out = h->H;
    std::vector<int> lo(out), hi(out), ibeg(in, out), iend(in, 0);
    std::vector<float> lerp(out);
    // TF1 legacy resize_bilinear(align_corners=True) scaler, fp32
    const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    for (int i = 0; i < out; ++i) {
        float src = (float)i * scale;
        int l = (int)floorf(src);
        int u = (int)ceilf(src);
        if (u > in - 1) u = in - 1;
        lo[i] = l; hi[i] = u; lerp[i] = src - (float)l;
        for (int j : {l, u}) {
            if (i < ibeg[j]) ibeg[j] = i;
            if (i + 1 > iend[j]) iend[j] = i + 1;
        }
    }
    if (dmalloc(&h->up.lo, out) || dmalloc(&h->up.hi, out) || dmalloc(&h->up.lerp, out) ||
        dmalloc(&h->up.ibeg, in) || dmalloc(&h->up.iend, in))
        return SR4D_ENOMEM;
    cudaMemcpy(h->up.lo, lo.data(), out * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(h->up.hi, hi.data(), out * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(h->up.lerp, lerp.data(), out * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(h->up.ibeg, ibeg.data(), in * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(h->up.iend, iend.data(), in * sizeof(int), cudaMemcpyHostToDevice);
    return SR4D_OK;
}

int alloc_act(ActBuf& b, int maxB, int D, unsigned int* ovf = nullptr) {
    b.D = D;
    b.ovf = ovf;
    size_t n = 2 * act_plane_elems(maxB, D);
    if (dmalloc(&b.base, n) != cudaSuccess) return SR4D_ENOMEM;
    return SR4D_OK;
}

// tensor indices (see forward())
inline int lr_t_of_block_t(int k) { return 6 + 2 * k; }
inline int lr_t_of_block_x(int k) { return 7 + 2 * k; }   // output of LR block k
inline int hr_t_of_block_t(int k) { return 1 + 2 * k; }
inline int hr_t_of_block_x(int k) { return 2 + 2 * k; }

int plan_buffers(sr4d_t* h) {
    h->n_lr_t = 6 + 2 * h->low;
    h->n_hr_t = 1 + 2 * h->hi + 3;
    h->lr_slot.assign(h->n_lr_t, 0);
    h->hr_slot.assign(h->n_hr_t, 0);
    int n_lr_slots, n_hr_slots;
    if (h->training) {
        for (int i = 0; i < h->n_lr_t; ++i) h->lr_slot[i] = i;
        for (int i = 0; i < h->n_hr_t; ++i) h->hr_slot[i] = i;
        n_lr_slots = h->n_lr_t;
        n_hr_slots = h->n_hr_t;
    } else {
        // inference: rotate; pc1->0 pc2->1 ph1->2 ph2->3 c1->0 f->2, t_k->0, x_{k+1}-> 1,2,1,2..
        const int base[6] = {0, 1, 2, 3, 0, 2};
        for (int i = 0; i < 6; ++i) h->lr_slot[i] = base[i];
        for (int k = 0; k < h->low; ++k) {
            h->lr_slot[lr_t_of_block_t(k)] = 0;
            h->lr_slot[lr_t_of_block_x(k)] = (k % 2 == 0) ? 1 : 2;
        }
        n_lr_slots = 4;
        h->hr_slot[0] = 0;
        for (int k = 0; k < h->hi; ++k) {
            h->hr_slot[hr_t_of_block_t(k)] = 1;
            h->hr_slot[hr_t_of_block_x(k)] = (k % 2 == 0) ? 2 : 0;
        }
        // heads: slot 1 (the block scratch, free once the trunk is final) and two extra slots
        const int head_slots[3] = {1, 3, 4};
        for (int c = 0; c < 3; ++c) h->hr_slot[1 + 2 * h->hi + c] = head_slots[c];
        n_hr_slots = 5;
    }
    h->lr.resize(n_lr_slots);
    for (auto& b : h->lr)
        if (alloc_act(b, h->maxB, h->P, h->ovf)) return SR4D_ENOMEM;
    if (h->r == 1) {
        // upsample is the identity (SR4DFlowNet.py:72-74): HR tensor 0 aliases the LR trunk
        h->hr.resize(n_hr_slots);
        for (int i = 0; i < n_hr_slots; ++i) {
            bool is_up_slot = (i == h->hr_slot[0]);
            if (h->training && is_up_slot) { h->hr[i].D = h->H; continue; }   // aliased at run time
            if (alloc_act(h->hr[i], h->maxB, h->H, h->ovf)) return SR4D_ENOMEM;
        }
    } else {
        h->hr.resize(n_hr_slots);
        for (auto& b : h->hr)
            if (alloc_act(b, h->maxB, h->H, h->ovf)) return SR4D_ENOMEM;
    }
    return SR4D_OK;
}

ActView lr_view(sr4d_t* h, int t, int B) { return h->lr[h->lr_slot[t]].view(B); }
ActView hr_view(sr4d_t* h, int t, int B) {
    if (t == 0 && h->r == 1) return lr_view(h, h->low ? lr_t_of_block_x(h->low - 1) : 5, B);
    return h->hr[h->hr_slot[t]].view(B);
}

const float* W(sr4d_t* h, int layer) { return h->params + h->layers[layer].w_off; }
const float* Bv(sr4d_t* h, int layer) { return h->layers[layer].b_off >= 0 ? h->params + h->layers[layer].b_off : nullptr; }
float* GW(sr4d_t* h, int layer) { return h->grads + h->layers[layer].w_off; }
float* GB(sr4d_t* h, int layer) { return h->grads + h->layers[layer].b_off; }

bool use_tc(sr4d_t* h, int impl_override = -1) {
    int impl = impl_override >= 0 ? impl_override : h->conv_impl;
    return impl != SR4D_CONV_SIMT;
}

int ensure_tc_weights(sr4d_t* h, cudaStream_t s) {
    if (!h->tcw_dirty) return SR4D_OK;
    std::vector<int> idx;
    std::vector<long long> off;
    for (size_t i = 0; i < h->layers.size(); ++i) {
        const Layer& ly = h->layers[i];
        if (ly.k == 3 && ly.cin == 64 && ly.cout == 64) { idx.push_back((int)i); off.push_back(ly.w_off); }
    }
    if (!idx.empty()) CK(h, tc_prepare_weights(h->tcw, h->params, idx.data(), off.data(), (int)idx.size(), s), 1);
    h->tcw_dirty = false;
    return SR4D_OK;
}

// one 64->64 3x3x3 conv layer, Act -> Act
int conv64_fwd(sr4d_t* h, int layer, ActView in, ActView out, const ActView* res, float slope, cudaStream_t s) {
    ProfScope prof(h, in.D == h->P ? SR4D_PROF_CONV64_FWD_LR : SR4D_PROF_CONV64_FWD_HR, s);
    if (use_tc(h)) {
        TcConvArgs a;
        a.in = in; a.out = out; a.layer = layer; a.dgrad = 0;
        a.bias = Bv(h, layer);
        a.res_hi = res ? res->hi : nullptr; a.res_lo = res ? res->lo : nullptr;
        a.slope = slope; a.halo = 1;
        CK(h, tc_conv64(h->tcw, a, s), 1);
        return SR4D_OK;
    }
    Conv64Args a;
    a.in_hi = in.hi; a.in_lo = in.lo; a.B = in.B; a.Do = in.D;
    a.w = W(h, layer);
    a.out_hi = out.hi; a.out_lo = out.lo; a.halo = 1; a.ovf = out.ovf;
    a.bias = Bv(h, layer);
    if (res) { a.res_hi = res->hi; a.res_lo = res->lo; }
    a.slope = slope;
    CK(h, launch_conv64_simt(a, s), 1);
    return SR4D_OK;
}

// A run of consecutive 64->64 forward layers on one grid as one chained launch (tensor-core path; SR4D_NO_CHAIN=1 keeps
// one launch per layer).  `layers` lists (layer, in, out, residual or NULL, slope).
struct FwdLayer { int layer; ActView in, out; bool has_res; ActView res; float slope; };
int conv64_fwd_run(sr4d_t* h, int kind, const std::vector<FwdLayer>& layers, cudaStream_t s) {
    static const bool no_chain = getenv("SR4D_NO_CHAIN") != nullptr;
    // chaining pays where a layer is a few tiles per SM (batch 1: launch + prologue + drain are a third of a 24^3 layer);
    // on large grids the per-layer launches are as fast and keep the per-class timing simple
    static const int env_tiles_per_sm = getenv("SR4D_CHAIN_TILES_PER_SM") ? atoi(getenv("SR4D_CHAIN_TILES_PER_SM")) : 4;
    const int max_tiles_per_sm = h->fwd_chain >= 0 ? h->fwd_chain : env_tiles_per_sm;
    int rc;
    const bool small = !layers.empty() && tc_fwd_tiles(layers[0].in.D, layers[0].in.B) <= (long)max_tiles_per_sm * tc_num_sms();
    if (!use_tc(h) || no_chain || max_tiles_per_sm == 0 || layers.size() < 2 || !small) {
        for (const auto& l : layers)
            if ((rc = conv64_fwd(h, l.layer, l.in, l.out, l.has_res ? &l.res : nullptr, l.slope, s))) return rc;
        return SR4D_OK;
    }
    const int B = layers[0].in.B;
    const int key = kind * 65536 + B;
    auto it = h->chains.find(key);
    if (it == h->chains.end()) {
        std::vector<TcConvArgs> args(layers.size());
        for (size_t i = 0; i < layers.size(); ++i) {
            TcConvArgs& a = args[i];
            a.in = layers[i].in; a.out = layers[i].out; a.layer = layers[i].layer; a.dgrad = 0;
            a.bias = Bv(h, layers[i].layer);
            a.res_hi = layers[i].has_res ? layers[i].res.hi : nullptr;
            a.res_lo = layers[i].has_res ? layers[i].res.lo : nullptr;
            a.slope = layers[i].slope; a.halo = 1;
        }
        TcChain* c = nullptr;
        CK(h, tc_chain_build(h->tcw, args.data(), (int)args.size(), &c), 0);
        it = h->chains.emplace(key, c).first;
    }
    {
        ProfScope prof(h, layers[0].in.D == h->P ? SR4D_PROF_CONV64_FWD_LR : SR4D_PROF_CONV64_FWD_HR, s, (int)layers.size());
        if (tc_chain_launch(it->second, s) == cudaSuccess) { h->launches += 1; return SR4D_OK; }
    }
    // the cooperative launch was refused (the grid cannot be co-resident here): one launch per layer from now on
    (void)cudaGetLastError();
    h->fwd_chain = 0;
    for (const auto& l : layers)
        if ((rc = conv64_fwd(h, l.layer, l.in, l.out, l.has_res ? &l.res : nullptr, l.slope, s))) return rc;
    return SR4D_OK;
}

int forward_impl(sr4d_t* h, const float* u, const float* v, const float* w, const float* um, const float* vm,
                 const float* wm, float* out, int B, cudaStream_t s) {
    if (B < 1 || B > h->maxB) return fail(h, SR4D_EINVAL, "batch size out of range (1..max_batch)");
    NvtxPhase nvtx_phase(h, "sr4d forward");
    int rc = ensure_tc_weights(h, s);
    if (rc) return rc;
    const int P = h->P;
    CK(h, launch_prep_features(u, v, w, um, vm, wm, h->feat, B, P, s), 1);
    ActView pc1 = lr_view(h, 0, B), pc2 = lr_view(h, 1, B), ph1 = lr_view(h, 2, B), ph2 = lr_view(h, 3, B);
    ActView c1 = lr_view(h, 4, B), f = lr_view(h, 5, B);
    CK(h, launch_stem_convs(h->feat, W(h, 0), Bv(h, 0), pc1, W(h, 2), Bv(h, 2), ph1, s), 1);   // SR4DFlowNet.py:17,20 (one launch)
    if ((rc = conv64_fwd(h, 1, pc1, pc2, nullptr, 0.f, s))) return rc;                  // :18
    if ((rc = conv64_fwd(h, 3, ph1, ph2, nullptr, 0.f, s))) return rc;                  // :21  (chaining these two independent
    // layers into one launch was measured slower at batch 1: 61 vs 41 us, profiles/r02_chain_cluster_ab.txt)
    CK(h, launch_conv1x1_cat(ph2, pc2, W(h, 4), Bv(h, 4), c1, s), 1);                  // :23-24
    std::vector<FwdLayer> run;
    run.push_back(FwdLayer{5, c1, f, false, ActView(), 0.f});                          // :25
    int li = 6;
    ActView x = f;
    for (int k = 0; k < h->low; ++k) {                                                  // :28-30
        ActView t = lr_view(h, lr_t_of_block_t(k), B), xo = lr_view(h, lr_t_of_block_x(k), B);
        run.push_back(FwdLayer{li, x, t, false, ActView(), 0.2f});
        run.push_back(FwdLayer{li + 1, t, xo, true, x, 0.2f});
        x = xo;
        li += 2;
    }
    if ((rc = conv64_fwd_run(h, 0, run, s))) return rc;
    ActView xh;
    if (h->r == 1) {
        xh = x;                                                                          // :72-74
    } else {
        xh = hr_view(h, 0, B);
        CK(h, launch_upsample(x, xh, h->r, h->up.tables(), s), 1);                      // :32
    }
    run.clear();
    for (int k = 0; k < h->hi; ++k) {                                                   // :35-36
        ActView t = hr_view(h, hr_t_of_block_t(k), B), xo = hr_view(h, hr_t_of_block_x(k), B);
        run.push_back(FwdLayer{li, xh, t, false, ActView(), 0.2f});
        run.push_back(FwdLayer{li + 1, t, xo, true, xh, 0.2f});
        xh = xo;
        li += 2;
    }
    ActView hd[3];
    for (int c = 0; c < 3; ++c) {                                                       // :39-46
        hd[c] = hr_view(h, 1 + 2 * h->hi + c, B);
        run.push_back(FwdLayer{li + 2 * c, xh, hd[c], false, ActView(), 0.f});
    }
    if ((rc = conv64_fwd_run(h, 1, run, s))) return rc;
    static const bool head_simt = getenv("SR4D_HEAD_SIMT") != nullptr;   // debugging aid / A-B: fp32 head_out_kernel
    if (use_tc(h) && !head_simt)                                                         // :40,43,46,49
        CK(h, launch_head_out_tc(hd[0], hd[1], hd[2], W(h, li + 1), W(h, li + 3), W(h, li + 5), Bv(h, li + 1), Bv(h, li + 3),
                                 Bv(h, li + 5), h->head_wimg, h->tapP[0], h->tapP[1], h->tapP[2], out, s), 3);
    else
        CK(h, launch_head_out(hd[0], hd[1], hd[2], W(h, li + 1), W(h, li + 3), W(h, li + 5), Bv(h, li + 1),
                              Bv(h, li + 3), Bv(h, li + 5), out, s), 1);
    h->have_fwd_state = h->training != 0 && out == h->pred;   // backward differentiates h->pred
    h->fwd_batch = B;
    return SR4D_OK;
}

// with the single-plane dgrad and the stacked single-plane wgrad nobody reads the lo plane of a split gradient
bool lo_plane_dead(const sr4d_t* h) { return h->dgrad_single == 1 && h->wgrad_single == 1; }

// A gradient tensor has just been written to g.f (and its |max| to g.amax): make the scaled
// split-fp16 copy the tensor-core dgrad / wgrad kernels consume.
int grad_ready(sr4d_t* h, GBuf& g, int B, int D, cudaStream_t s) {
    if (!use_tc(h)) return SR4D_OK;
    CK(h, launch_g4_split(g.f, g.amax, g.s, g.exp, B, D, s, lo_plane_dead(h)), 1);
    return SR4D_OK;
}

// dgrad of a 64->64 layer: dy (G4, edge D) -> raw (edge D+2)
int conv64_dgrad(sr4d_t* h, int layer, const GBuf& dy, RawBuf& raw, int B, int D, cudaStream_t s) {
    ProfScope prof(h, D == h->P ? SR4D_PROF_CONV64_DGRAD_LR : SR4D_PROF_CONV64_DGRAD_HR, s);
    if (use_tc(h)) {
        TcConvArgs a;
        a.in.hi = dy.s; a.in.lo = dy.s + act_plane_elems(B, D + 2); a.in.B = B; a.in.D = D + 2;
        a.layer = layer; a.dgrad = 1; a.single_b = h->dgrad_single;
        a.out_raw = raw.p;
        raw.exp = dy.exp;
        CK(h, tc_conv64(h->tcw, a, s), 1);
        return SR4D_OK;
    }
    Conv64Args a;
    a.in_f32 = dy.f; a.B = B; a.Do = D + 2;
    a.w = W(h, layer); a.dgrad = 1;
    a.out_raw = raw.p;
    raw.exp = nullptr;
    CK(h, launch_conv64_simt(a, s), 1);
    return SR4D_OK;
}
// batched weight gradient (default): stacked single-plane kernel, per-layer partials, nothing reduced per layer
bool wgrad_batched(const sr4d_t* h) {
    static const bool off = getenv("SR4D_WGRAD_REDUCE_EACH") != nullptr || getenv("SR4D_WGRAD_UNBATCHED") != nullptr;
    static const bool wgrad_simt = getenv("SR4D_WGRAD_SIMT") != nullptr;
    return !off && !wgrad_simt && h->conv_impl != SR4D_CONV_SIMT && h->wgrad_single == 1 && h->wg_part && h->lexp;
}
// the tensor about to be written into g is the dY of `layer` (-1: nobody's, use the buffer's own storage): with the
// batched weight gradient its split copy must survive until the end of the pass
void layer_split(sr4d_t* h, GBuf& g, int layer) {
    if (layer >= 0 && wgrad_batched(h) && h->lsplit[layer]) { g.s = h->lsplit[layer]; g.exp = h->lexp + layer; }
    else { g.s = g.s0; g.exp = g.exp0; }
}
int conv64_wgrad(sr4d_t* h, int layer, ActView x, const GBuf& dy, bool bias, cudaStream_t s) {
    if (wgrad_batched(h) && h->wg_slot[layer] >= 0 && dy.s == h->lsplit[layer]) {
        // computed by ONE launch per grid at the end of the pass (wgrad_finish); dY stays in the layer's own buffer
        TcWgradItem it;
        it.x = x; it.dy_split = dy.s; it.dy_exp = dy.exp;
        it.partial = h->wg_part + (size_t)h->wg_slot[layer] * h->wg_stride;
        h->wg_list[x.D == h->P ? 0 : 1].push_back(it);
        h->wg_pending = true;
        if (bias) CK(h, launch_bias_grad(dy.f, x.B, x.D, GB(h, layer), h->scratch, s), 2);
        return SR4D_OK;
    }
    ProfScope prof(h, x.D == h->P ? SR4D_PROF_CONV64_WGRAD_LR : SR4D_PROF_CONV64_WGRAD_HR, s);
    static const bool wgrad_simt = getenv("SR4D_WGRAD_SIMT") != nullptr;   // debugging aid
    if (use_tc(h) && !wgrad_simt) {
        static const bool immediate = getenv("SR4D_WGRAD_REDUCE_EACH") != nullptr;   // A-B: one reduce launch per layer
        if (h->wgrad_single == 1 && h->wg_part && !immediate && h->wg_slot[layer] >= 0) {
            CK(h, tc_wgrad64_single(x, dy.s, dy.exp, h->wg_part + (size_t)h->wg_slot[layer] * h->wg_stride, s), 1);
            h->wg_pending = true;                      // summed by wgrad_finish() at the end of the backward pass
        } else if (h->wgrad_single == 1) {
            CK(h, tc_wgrad64_single(x, dy.s, dy.exp, h->scratch, s), 1);
            CK(h, launch_reduce_rows(h->scratch, tc_wgrad2_slabs(x.B, x.D), 27 * 4096, GW(h, layer), s), 1);
        } else {
            CK(h, tc_wgrad64(x, dy.s, dy.exp, h->scratch, s, h->wgrad_single == 2), 1);
            CK(h, launch_reduce_rows(h->scratch, tc_wgrad_slabs(x.B, x.D), 27 * 4096, GW(h, layer), s), 1);
        }
    } else {
        CK(h, launch_wgrad64_simt(x, dy.f, GW(h, layer), h->scratch, h->wgrad_chunks, s), 2);
    }
    if (bias) CK(h, launch_bias_grad(dy.f, x.B, x.D, GB(h, layer), h->scratch, s), 2);
    return SR4D_OK;
}
// one launch sums the slab partials of every 64->64 layer's weight gradient (all layers run in every backward pass)
int wgrad_finish(sr4d_t* h, int B, cudaStream_t s) {
    if (!h->wg_pending) return SR4D_OK;
    h->wg_pending = false;
    for (int cls = 0; cls < 2; ++cls) {
        auto& list = h->wg_list[cls];
        if (list.empty()) continue;
        ProfScope prof(h, cls == 0 ? SR4D_PROF_CONV64_WGRAD_LR : SR4D_PROF_CONV64_WGRAD_HR, s, (int)list.size());
        const long long key = ((long long)cls << 40) | ((long long)B << 16) | (long long)list.size();
        auto it = h->wg_batches.find(key);
        if (it == h->wg_batches.end()) {
            TcWgradBatch* b = nullptr;
            CK(h, tc_wgrad_batch_build(list.data(), (int)list.size(), &b), 0);
            it = h->wg_batches.emplace(key, b).first;
        }
        cudaError_t e = tc_wgrad_batch_launch(it->second, s);
        list.clear();
        CK(h, e, 1);
    }
    CK(h, launch_reduce_rows_batched(h->wg_items, h->wg_nitems, tc_wgrad2_slabs(B, h->P), tc_wgrad2_slabs(B, h->H), 27 * 4096, s), 1);
    return SR4D_OK;
}
// out = (fold(raw0 [+raw1 +raw2]) + add) * act'(saved); then the split copy
int fold_act(sr4d_t* h, const RawBuf* r0, const RawBuf* r1, const RawBuf* r2, const GBuf* add, const ActView* saved,
             float slope, GBuf& out, int B, int D, cudaStream_t s) {
    CK(h, cudaMemsetAsync(out.amax, 0, sizeof(int), s), 0);
    CK(h, launch_fold_act(r0->p, r1 ? r1->p : nullptr, r2 ? r2->p : nullptr, r0->exp, r1 ? r1->exp : nullptr,
                          r2 ? r2->exp : nullptr, add ? add->f : nullptr, saved ? saved->hi : nullptr,
                          saved ? saved->lo : nullptr, slope, out.f, out.amax, B, D, s), 1);
    return grad_ready(h, out, B, D, s);
}

// does the tensor-core dgrad fold / add / activation-gradient inside its epilogue for this grid?
bool dgrad_fused(sr4d_t* h, int D) {
    return use_tc(h) && h->fused_dgrad && tc_dgrad_fusable(D);
}
// fused dgrad of `layer`: out.f (interior) = (fold(dgrad(dy)) + add_pre) * act'(saved) + add_post; updates out.amax
int conv64_dgrad_fused(sr4d_t* h, int layer, const GBuf& dy, const GBuf* add_pre, const float* add_post,
                       const ActView* saved, float slope, GBuf& out, bool with_split, int B, int D, cudaStream_t s,
                       bool want_f32 = true) {
    ProfScope prof(h, D == h->P ? SR4D_PROF_CONV64_DGRAD_LR : SR4D_PROF_CONV64_DGRAD_HR, s);
    TcConvArgs a;
    a.in.hi = dy.s; a.in.lo = dy.s + act_plane_elems(B, D + 2); a.in.B = B; a.in.D = D + 2;
    a.layer = layer; a.dgrad = 1; a.fused = 1; a.single_b = h->dgrad_single;
    a.dy_exp = dy.exp; a.add_pre = add_pre ? add_pre->f : nullptr; a.add_post = add_post;
    if (with_split) {
        a.split_out = out.s; a.split_exp = out.exp; a.dy_amax = dy.amax; a.split_hi_only = lo_plane_dead(h);
        // bound on |out|: this call's contribution plus what it is added to (skip gradient, or the in-place partial sum,
        // whose |max| the caller copied to amax_copy before this launch started updating out.amax)
        a.add_amax = add_pre ? add_pre->amax : (add_post ? h->amax_copy : nullptr);
    }
    a.sav_hi = saved ? saved->hi : nullptr; a.sav_lo = saved ? saved->lo : nullptr;
    a.slope = slope; a.out_g4 = (want_f32 || !with_split) ? out.f : nullptr; a.absmax = out.amax;
    CK(h, tc_conv64(h->tcw, a, s), 1);
    return SR4D_OK;
}
// gradient wrt the pre-activation of a 64->64 layer's input: out = (MirrorPadGrad(dgrad(dy)) + add) * act'(saved)
// (saved == NULL: no activation), followed by the split copy for the tensor-core consumers
// want_f32 = false: nobody reads the fp32 tensor out.f (every consumer takes the scaled split copy), the fused kernel may skip it
int dgrad_fold(sr4d_t* h, int layer, const GBuf& dy, const GBuf* add, const ActView* saved, float slope, GBuf& out,
               RawBuf& raw, int B, int D, cudaStream_t s, bool want_f32 = true) {
    int rc;
    if (dgrad_fused(h, D)) {
        CK(h, cudaMemsetAsync(out.amax, 0, sizeof(int), s), 0);
        // the epilogue also writes the scaled split copy (exponent from a rigorous bound): no g4_split pass
        return conv64_dgrad_fused(h, layer, dy, add, nullptr, saved, slope, out, true, B, D, s, want_f32);
    }
    if ((rc = conv64_dgrad(h, layer, dy, raw, B, D, s))) return rc;
    return fold_act(h, &raw, nullptr, nullptr, add, saved, slope, out, B, D, s);
}

// backward through `nblk` resnet blocks whose first conv is layer `l0`; bufs[0] holds the gradient wrt the
// pre-activation of the last block output on entry; on exit bufs[*result_idx] holds the gradient wrt the
// pre-activation (if slope_in >= 0, else the value) of the first block input.
// next_layer: the 64->64 layer whose dY the final result is (-1: none -- its split copy may live in the rotating buffer)
int blocks_bwd(sr4d_t* h, int nblk, int l0, bool hr, GBuf* bufs[3], RawBuf& raw, int B, int D,
               float slope_in, cudaStream_t s, int* result_idx, int next_layer) {
    int si = 0;
    int rc;
    for (int k = nblk - 1; k >= 0; --k) {
        const int la = l0 + 2 * k, lb = la + 1;
        ActView t = hr ? hr_view(h, hr_t_of_block_t(k), B) : lr_view(h, lr_t_of_block_t(k), B);
        ActView xin;
        if (hr) xin = k == 0 ? hr_view(h, 0, B) : hr_view(h, hr_t_of_block_x(k - 1), B);
        else xin = k == 0 ? lr_view(h, 5, B) : lr_view(h, lr_t_of_block_x(k - 1), B);
        GBuf& S = *bufs[si];
        GBuf& T = *bufs[(si + 1) % 3];
        GBuf& S2 = *bufs[(si + 2) % 3];
        if ((rc = conv64_wgrad(h, lb, t, S, false, s))) return rc;
        layer_split(h, T, la);                                   // T becomes dY of conv a
        // T (the gradient inside the block) is read by conv a's weight gradient and dgrad only: as the split copy when the
        // batched tensor-core weight gradient runs, so its fp32 tensor (256 B per voxel) is not written then
        if ((rc = dgrad_fold(h, lb, S, nullptr, &t, 0.2f, T, raw, B, D, s, !wgrad_batched(h)))) return rc;
        if ((rc = conv64_wgrad(h, la, xin, T, false, s))) return rc;
        layer_split(h, S2, k > 0 ? la - 1 : next_layer);         // S2 becomes dY of the previous block's conv b
        // gradient wrt x_k (post-activation) = fold + skip path; multiply by its producer's act'
        float slope = k > 0 ? 0.2f : slope_in;
        if (slope >= 0.f) rc = dgrad_fold(h, la, T, &S, &xin, slope, S2, raw, B, D, s);
        else rc = dgrad_fold(h, la, T, &S, nullptr, 1.f, S2, raw, B, D, s);
        if (rc) return rc;
        si = (si + 2) % 3;
    }
    *result_idx = si;
    return SR4D_OK;
}

int backward_impl(sr4d_t* h, const float* hu, const float* hv, const float* hw, const float* mask, int B,
                  float* per_sample, float* l2_out, cudaStream_t s) {
    if (!h->training) return fail(h, SR4D_ESTATE, "handle was not created for training");
    NvtxPhase nvtx_phase(h, "sr4d loss + backward");
    const int P = h->P, H = h->H;
    const int nvoxH = H * H * H;
    int rc;
    CK(h, launch_loss_stats(h->pred, hu, hv, hw, mask, B, nvoxH, h->dpartial, 64, per_sample, h->norm, s), 2);
    CK(h, launch_loss_grad(h->pred, hu, hv, hw, mask, B, nvoxH, h->norm, h->gpred, h->gmax, s), 1);
    if (l2_out)
        CK(h, launch_sumsq(h->params, h->kflag, h->flat, h->dpartial + 64 * 5 * h->maxB, 256, 5e-7f, l2_out, s), 2);

    const int l_hr0 = 6 + 2 * h->low;          // first HR block layer
    const int l_head = l_hr0 + 2 * h->hi;      // first head layer
    ActView trunk = h->hi > 0 ? hr_view(h, hr_t_of_block_x(h->hi - 1), B) : hr_view(h, 0, B);
    GBuf& A = h->g4_hr[0];
    // trunk gradient: sum of the three heads' input gradients; trunk producer: LeakyReLU block if hi>0,
    // the (linear) upsample if hi==0 and r>1, else the LR trunk tensor itself (r==1 aliases it).
    GBuf* hb[3] = {&h->g4_hr[1], &h->g4_hr[2], &h->g4_hr[0]};
    float slope_lr_trunk = h->low > 0 ? 0.2f : 0.f;    // producer activation of the LR trunk tensor
    const bool trunk_act = h->hi > 0 || h->r == 1;
    const float trunk_slope = h->hi > 0 ? 0.2f : (h->r == 1 ? slope_lr_trunk : 1.f);
    const bool fused_heads = dgrad_fused(h, H);
    // batched weight gradient: every gradient tensor that is some 64->64 layer's dY gets its split copy written into
    // that layer's own buffer (layer_split) so that it survives until the single wgrad launch of wgrad_finish
    for (auto& g : h->g4_lr) layer_split(h, g, -1);
    for (auto& g : h->g4_hr) layer_split(h, g, -1);
    h->wg_list[0].clear(); h->wg_list[1].clear();
    const int lr_top = h->low > 0 ? 6 + 2 * (h->low - 1) + 1 : 5;                       // last 64->64 layer of the LR trunk
    layer_split(h, *hb[0], h->hi > 0 ? l_hr0 + 2 * (h->hi - 1) + 1 : (h->r == 1 ? lr_top : -1));   // the trunk gradient
    if (fused_heads) CK(h, cudaMemsetAsync(hb[0]->amax, 0, sizeof(int), s), 0);
    static const bool head_simt = getenv("SR4D_HEAD_SIMT") != nullptr;   // debugging aid / A-B: fp32 head2_bwd_kernel
    const bool heads_tc = use_tc(h) && !head_simt && h->hb_part && head_bwd_tc_supported(H);
    if (heads_tc)
        CK(h, launch_head_bwd_tc_setup(W(h, l_head + 1), W(h, l_head + 3), W(h, l_head + 5), h->gmax, h->hb_scales,
                                       h->hb_wimg, h->gpred, h->hb_gplanar, B, H, s), 2);
    for (int c = 0; c < 3; ++c) {
        ActView hd = hr_view(h, 1 + 2 * h->hi + c, B);
        const int l1 = l_head + 2 * c, l2 = l1 + 1;
        // whole backward of the 64->1 conv, plus what its input gradient needs downstream: the bias gradient of
        // the head's first conv and (tensor-core path) the scaled split copy
        CK(h, cudaMemsetAsync(A.amax, 0, sizeof(int), s), 0);
        layer_split(h, A, l1);
        if (heads_tc)
            CK(h, launch_head_bwd_tc(hd, h->hb_gplanar, c, h->hb_scales, h->hb_wimg, h->hb_part + c * h->hb_part_stride, A.s, A.exp,
                                     A.amax, lo_plane_dead(h), s), 1);
        else
            CK(h, launch_head2_bwd(hd, h->gpred, c, W(h, l2), A.f, A.amax, GW(h, l2), GB(h, l2), GB(h, l1),
                                   use_tc(h) ? A.s : nullptr, A.exp, h->gmax, h->scratch, s, lo_plane_dead(h)), 5);
        if ((rc = conv64_wgrad(h, l1, trunk, A, false, s))) return rc;
        if (fused_heads) {
            // the three heads accumulate act'(trunk) * fold(dgrad_c) in place (the activation gradient is linear);
            // the last call also writes the split copy, bounded by this head's gain plus the |max| accumulated so far
            const bool last = c == 2;
            if (last) CK(h, cudaMemcpyAsync(h->amax_copy, hb[0]->amax, sizeof(int), cudaMemcpyDeviceToDevice, s), 0);
            if ((rc = conv64_dgrad_fused(h, l1, A, nullptr, c ? hb[0]->f : nullptr, trunk_act ? &trunk : nullptr,
                                         trunk_slope, *hb[0], last, B, H, s))) return rc;
            continue;
        }
        if ((rc = conv64_dgrad(h, l1, A, h->raw_hr[c], B, H, s))) return rc;
        if (use_tc(h) && c < 2) {
            // raw_hr[c] stays scaled by A's exponent, which the next head overwrites: keep a private copy
            CK(h, cudaMemcpyAsync(h->head_exp + c, A.exp, sizeof(int), cudaMemcpyDeviceToDevice, s), 0);
            h->raw_hr[c].exp = h->head_exp + c;
        }
    }
    if (heads_tc) {
        const int nc = head_bwd_tc_grid(B, H);
        const float* part[3] = {h->hb_part, h->hb_part + h->hb_part_stride, h->hb_part + 2 * h->hb_part_stride};
        const int ncta[3] = {nc, nc, nc};
        float* dw[3] = {GW(h, l_head + 1), GW(h, l_head + 3), GW(h, l_head + 5)};
        float* db[3] = {GB(h, l_head + 1), GB(h, l_head + 3), GB(h, l_head + 5)};
        float* db1[3] = {GB(h, l_head), GB(h, l_head + 2), GB(h, l_head + 4)};
        CK(h, launch_head_bwd_tc_finish(part, ncta, h->hb_scales, dw, db, db1, lo_plane_dead(h), s), 1);
    }
    if (!fused_heads) {
        if ((rc = fold_act(h, &h->raw_hr[0], &h->raw_hr[1], &h->raw_hr[2], nullptr, trunk_act ? &trunk : nullptr,
                           trunk_slope, *hb[0], B, H, s))) return rc;
    }
    GBuf* S;
    if (h->hi > 0) {
        int ri = 0;
        float slope_in = h->r == 1 ? slope_lr_trunk : -1.f;
        if ((rc = blocks_bwd(h, h->hi, l_hr0, true, hb, h->raw_hr[0], B, H, slope_in, s, &ri, h->r == 1 ? lr_top : -1))) return rc;
        S = hb[ri];
    } else {
        S = hb[0];
    }
    // through the upsample
    GBuf* lb[3] = {&h->g4_lr[0], &h->g4_lr[1], &h->g4_lr[2]};
    ActView lr_trunk = h->low > 0 ? lr_view(h, lr_t_of_block_x(h->low - 1), B) : lr_view(h, 5, B);
    if (h->r == 1) {
        // same grid (H == P): S is already multiplied by the LR trunk's act'; keep rotating the HR buffers
        int si = 0;
        for (int i = 0; i < 3; ++i) if (hb[i] == S) si = i;
        lb[0] = hb[si]; lb[1] = hb[(si + 1) % 3]; lb[2] = hb[(si + 2) % 3];
    } else {
        CK(h, cudaMemsetAsync(lb[0]->amax, 0, sizeof(int), s), 0);
        layer_split(h, *lb[0], lr_top);
        CK(h, launch_upsample_bwd(S->f, lr_trunk, slope_lr_trunk, lb[0]->f, lb[0]->amax, B, P, h->r, h->up.tables(), s), 1);
        if ((rc = grad_ready(h, *lb[0], B, P, s))) return rc;
    }
    int ri = 0;
    if (h->low > 0) {
        if ((rc = blocks_bwd(h, h->low, 6, false, lb, h->raw_lr, B, P, 0.f, s, &ri, 5))) return rc;
    }
    GBuf& S5 = *lb[ri];                 // d pre-activation of conv3d_5 (fuse 3x3)
    GBuf& T = *lb[(ri + 1) % 3];
    GBuf& dA = *lb[(ri + 2) % 3];
    GBuf& dB = h->g4_lr[3];
    ActView pc1 = lr_view(h, 0, B), pc2 = lr_view(h, 1, B), ph1 = lr_view(h, 2, B), ph2 = lr_view(h, 3, B);
    ActView c1 = lr_view(h, 4, B);
    if ((rc = conv64_wgrad(h, 5, c1, S5, true, s))) return rc;
    layer_split(h, T, -1);
    layer_split(h, dA, 3);              // dY of the phase branch's second conv
    layer_split(h, dB, 1);              // dY of the pc branch's second conv
    if ((rc = dgrad_fold(h, 5, S5, nullptr, &c1, 0.f, T, h->raw_lr, B, P, s))) return rc;
    CK(h, cudaMemsetAsync(dA.amax, 0, sizeof(int), s), 0);
    CK(h, cudaMemsetAsync(dB.amax, 0, sizeof(int), s), 0);
    CK(h, launch_conv1x1_bwd(T.f, ph2, pc2, W(h, 4), dA.f, dB.f, dA.amax, dB.amax, GW(h, 4), GB(h, 4), h->scratch, s), 5);
    if ((rc = grad_ready(h, dA, B, P, s))) return rc;
    if ((rc = grad_ready(h, dB, B, P, s))) return rc;
    // phase branch
    if ((rc = conv64_wgrad(h, 3, ph1, dA, true, s))) return rc;
    if ((rc = dgrad_fold(h, 3, dA, nullptr, &ph1, 0.f, T, h->raw_lr, B, P, s))) return rc;
    CK(h, launch_stem_wgrad(h->feat, 0, T.f, B, P, GW(h, 2), GB(h, 2), h->scratch, s), 4);
    // pc branch
    if ((rc = conv64_wgrad(h, 1, pc1, dB, true, s))) return rc;
    if ((rc = dgrad_fold(h, 1, dB, nullptr, &pc1, 0.f, T, h->raw_lr, B, P, s))) return rc;
    CK(h, launch_stem_wgrad(h->feat, 3, T.f, B, P, GW(h, 0), GB(h, 0), h->scratch, s), 4);
    return wgrad_finish(h, B, s);
}

void free_all(sr4d_t* h) {
    cudaFree(h->params); cudaFree(h->grads); cudaFree(h->m); cudaFree(h->v); cudaFree(h->kflag); cudaFree(h->ovf);
    cudaFree(h->feat);
    for (auto t : h->tapP) cudaFree(t);
    cudaFree(h->head_wimg);
    cudaFree(h->hb_scales); cudaFree(h->hb_wimg); cudaFree(h->hb_part); cudaFree(h->hb_gplanar);
    cudaFree(h->wg_part); cudaFree(h->wg_items);
    for (auto p : h->lsplit) cudaFree(p);
    cudaFree(h->lexp);
    for (auto& kv : h->wg_batches) tc_wgrad_batch_free(kv.second);
    h->wg_batches.clear();
    for (auto& kv : h->chains) tc_chain_free(kv.second);
    h->chains.clear();
    for (auto& b : h->lr) cudaFree(b.base);
    for (auto& b : h->hr) cudaFree(b.base);
    cudaFree(h->up.lo); cudaFree(h->up.hi); cudaFree(h->up.lerp); cudaFree(h->up.ibeg); cudaFree(h->up.iend);
    for (auto& g : h->g4_lr) { cudaFree(g.f); cudaFree(g.s0); }
    for (auto& g : h->g4_hr) { cudaFree(g.f); cudaFree(g.s0); }
    cudaFree(h->raw_lr.p);
    for (auto& r : h->raw_hr) cudaFree(r.p);
    cudaFree(h->gmeta);
    cudaFree(h->pred); cudaFree(h->gpred); cudaFree(h->scratch); cudaFree(h->dpartial); cudaFree(h->norm);
    cudaFree(h->per_sample_int);
    if (h->tcw) tc_free_weights(h->tcw);
    for (auto e : h->ev_pool) cudaEventDestroy(e);
}

}  // namespace

extern "C" {

const char* sr4d_version(void) { return "sr4d 0.1 (sm_100a)"; }

int sr4d_create(sr4d_t** out, int patch_size, int res_increase, int low_resblock, int hi_resblock, int max_batch,
                int training, int device) {
    if (!out) return SR4D_EINVAL;
    *out = nullptr;
    if (patch_size < 4 || patch_size > 128 || res_increase < 1 || res_increase > 8 || low_resblock < 0 ||
        hi_resblock < 0 || max_batch < 1)
        return SR4D_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return SR4D_ENODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SR4D_ENODEVICE;
    if (prop.major != 10) return SR4D_ENODEVICE;   // sm_100a only: no fallback paths
    if (cudaSetDevice(device) != cudaSuccess) return SR4D_ENODEVICE;

    sr4d_t* h = new sr4d_handle();
    h->P = patch_size; h->r = res_increase; h->low = low_resblock; h->hi = hi_resblock;
    h->maxB = max_batch; h->training = training; h->device = device; h->H = patch_size * res_increase;
    h->save_acts = training;
    build_table(h);
    int rc = SR4D_OK;
    const size_t nvoxP = (size_t)h->P * h->P * h->P, nvoxH = (size_t)h->H * h->H * h->H;
    do {
        if (dmalloc(&h->params, h->flat) || dmalloc(&h->kflag, h->flat / 32) ||
            dmalloc(&h->feat, (size_t)h->maxB * nvoxP * 6)) { rc = SR4D_ENOMEM; break; }
        cudaMemset(h->params, 0, h->flat * sizeof(float));
        std::vector<unsigned char> kf(h->flat / 32, 0);
        for (auto& d : h->table)
            if (d.is_kernel)
                for (int64_t i = d.offset / 32; i < (d.offset + d.count + 31) / 32; ++i) kf[i] = 1;
        cudaMemcpy(h->kflag, kf.data(), kf.size(), cudaMemcpyHostToDevice);
        if (dmalloc(&h->ovf, 1)) { rc = SR4D_ENOMEM; break; }
        cudaMemset(h->ovf, 0, sizeof(unsigned int));
        if ((rc = plan_buffers(h))) break;
        {
            const size_t prow = (size_t)h->maxB * (h->H + 2) * (h->H + 2) * (h->H + 2);
            bool bad = dmalloc(&h->head_wimg, head_tc_wimg_halves()) != cudaSuccess;
            for (auto& t : h->tapP) bad |= dmalloc(&t, prow * 32) != cudaSuccess;
            if (bad) { rc = SR4D_ENOMEM; break; }
        }
        if ((rc = build_upsample_tables(h))) break;
        // zero-initialise activations once so halos of never-written regions are defined
        for (auto& b : h->lr) if (b.base) cudaMemset(b.base, 0, 2 * act_plane_elems(h->maxB, b.D) * sizeof(__half));
        for (auto& b : h->hr) if (b.base) cudaMemset(b.base, 0, 2 * act_plane_elems(h->maxB, b.D) * sizeof(__half));
        if (dmalloc(&h->dpartial, (size_t)64 * 5 * h->maxB + 256) || dmalloc(&h->norm, (size_t)2 * h->maxB) ||
            dmalloc(&h->per_sample_int, (size_t)4 * h->maxB)) { rc = SR4D_ENOMEM; break; }
        if (tc_alloc_weights(&h->tcw, (int)h->layers.size()) != cudaSuccess) { rc = SR4D_ENOMEM; break; }
        if (training) {
            if (dmalloc(&h->grads, h->flat + SR4D_METRIC_TAIL) || dmalloc(&h->m, h->flat) || dmalloc(&h->v, h->flat)) { rc = SR4D_ENOMEM; break; }
            cudaMemset(h->grads, 0, (h->flat + SR4D_METRIC_TAIL) * sizeof(float));
            cudaMemset(h->m, 0, h->flat * sizeof(float));
            cudaMemset(h->v, 0, h->flat * sizeof(float));
            const size_t g4l = (size_t)h->maxB * (h->P + 4) * (h->P + 4) * (h->P + 4) * 64;
            const size_t g4h = (size_t)h->maxB * (h->H + 4) * (h->H + 4) * (h->H + 4) * 64;
            const size_t rwl = (size_t)h->maxB * (h->P + 2) * (h->P + 2) * (h->P + 2) * 64;
            const size_t rwh = (size_t)h->maxB * (h->H + 2) * (h->H + 2) * (h->H + 2) * 64;
            bool bad = false;
            bad |= dmalloc(&h->gmeta, 20) != cudaSuccess;
            if (!bad) cudaMemset(h->gmeta, 0, 20 * sizeof(int));
            h->head_exp = h->gmeta + 14;
            h->gmax = reinterpret_cast<unsigned int*>(h->gmeta + 16);
            h->amax_copy = reinterpret_cast<unsigned int*>(h->gmeta + 17);
            int gi = 0;
            auto alloc_g = [&](GBuf& g, size_t n) {
                // fp32 tensor and its split-fp16 copy (2 planes of n halves); halos are zeroed once and never written
                bad |= dmalloc(&g.f, n) != cudaSuccess;
                if (!bad) cudaMemset(g.f, 0, n * 4);
                bad |= dmalloc(&g.s, 2 * n) != cudaSuccess;
                if (!bad) cudaMemset(g.s, 0, 2 * n * sizeof(__half));
                g.amax = reinterpret_cast<unsigned int*>(h->gmeta + 2 * gi);
                g.exp = h->gmeta + 2 * gi + 1;
                g.s0 = g.s; g.exp0 = g.exp;
                ++gi;
            };
            for (auto& g : h->g4_lr) alloc_g(g, g4l);
            for (auto& g : h->g4_hr) alloc_g(g, g4h);
            bad |= dmalloc(&h->raw_lr.p, rwl) != cudaSuccess;
            for (auto& r : h->raw_hr) bad |= dmalloc(&r.p, rwh) != cudaSuccess;
            bad |= dmalloc(&h->pred, (size_t)h->maxB * nvoxH * 3) != cudaSuccess;
            bad |= dmalloc(&h->gpred, (size_t)h->maxB * nvoxH * 3) != cudaSuccess;
            h->scratch_floats = (size_t)h->wgrad_chunks * 27 * 4096;
            for (int D : {h->P, h->H}) {
                const size_t need = (size_t)std::max(tc_wgrad_slabs(h->maxB, D), tc_wgrad2_slabs(h->maxB, D)) * 27 * 4096;
                if (need > h->scratch_floats) h->scratch_floats = need;
            }
            size_t need2 = (size_t)1184 * 8192 + 8192;
            if (need2 > h->scratch_floats) h->scratch_floats = need2;
            bad |= dmalloc(&h->scratch, h->scratch_floats) != cudaSuccess;
            {
                // per-layer slab partials of the stacked weight gradient + the item table of the batched reduction
                std::vector<ReduceItem> items;
                h->wg_slot.assign(h->layers.size(), -1);
                h->wg_stride = (size_t)std::max(tc_wgrad2_slabs(h->maxB, h->P), tc_wgrad2_slabs(h->maxB, h->H)) * 27 * 4096;
                for (size_t i = 0; i < h->layers.size(); ++i) {
                    const auto& ly = h->layers[i];
                    if (ly.k == 3 && ly.cin == 64 && ly.cout == 64) { h->wg_slot[i] = (int)items.size(); items.push_back(ReduceItem{nullptr, nullptr, 0, 0}); }
                }
                h->wg_nitems = (int)items.size();
                h->lsplit.assign(h->layers.size(), nullptr);
                // per-layer dY storage of the batched weight gradient: 4 GB at batch 8, r = 2; geometries where it would
                // pass 24 GB (r = 4 training at batch 16: 45 GB) keep the per-layer launches instead
                size_t lbytes = 0;
                for (size_t i = 0; i < h->layers.size(); ++i)
                    if (h->wg_slot[i] >= 0) lbytes += 2 * ((int)i >= 6 + 2 * h->low ? g4h : g4l) * sizeof(__half);
                const bool per_layer = lbytes <= ((size_t)24 << 30);
                if (per_layer) {
                    bad |= dmalloc(&h->lexp, h->layers.size()) != cudaSuccess;
                    if (!bad) cudaMemset(h->lexp, 0, h->layers.size() * sizeof(int));
                }
                for (size_t i = 0; i < h->layers.size() && !bad && per_layer; ++i) {
                    if (h->wg_slot[i] < 0) continue;
                    const size_t n2 = 2 * ((int)i >= 6 + 2 * h->low ? g4h : g4l);       // hi and lo plane, zero halo
                    bad |= dmalloc(&h->lsplit[i], n2) != cudaSuccess;
                    if (!bad) cudaMemset(h->lsplit[i], 0, n2 * sizeof(__half));
                }
                if (h->wg_nitems) {
                    bad |= dmalloc(&h->wg_part, h->wg_stride * h->wg_nitems) != cudaSuccess;
                    bad |= cudaMalloc(&h->wg_items, sizeof(ReduceItem) * h->wg_nitems) != cudaSuccess;
                    if (!bad) {
                        const int l_hr0 = 6 + 2 * h->low;                       // layers from here on run on the HR grid
                        for (size_t i = 0; i < h->layers.size(); ++i)
                            if (h->wg_slot[i] >= 0)
                                items[h->wg_slot[i]] = ReduceItem{h->wg_part + (size_t)h->wg_slot[i] * h->wg_stride, GW(h, (int)i),
                                                                  (int)i >= l_hr0 ? 1 : 0, 0};
                        cudaMemcpy(h->wg_items, items.data(), sizeof(ReduceItem) * h->wg_nitems, cudaMemcpyHostToDevice);
                    }
                }
            }
            if (head_bwd_tc_supported(h->H)) {
                h->hb_part_stride = head_bwd_tc_partial_floats(h->maxB, h->H);
                bad |= cudaMalloc(&h->hb_scales, head_bwd_tc_scales_bytes()) != cudaSuccess;
                bad |= dmalloc(&h->hb_wimg, head_bwd_tc_wimg_halves()) != cudaSuccess;
                bad |= dmalloc(&h->hb_part, 3 * h->hb_part_stride) != cudaSuccess;
                bad |= dmalloc(&h->hb_gplanar, head_bwd_tc_gplanar_floats(h->maxB, h->H)) != cudaSuccess;
            }
            if (bad) { rc = SR4D_ENOMEM; break; }
        }
        if (cudaDeviceSynchronize() != cudaSuccess) { rc = SR4D_ECUDA; break; }
    } while (0);
    if (rc != SR4D_OK) {
        free_all(h);
        delete h;
        return rc;
    }
    *out = h;
    return SR4D_OK;
}

void sr4d_destroy(sr4d_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    free_all(h);
    delete h;
}

const char* sr4d_last_error(const sr4d_t* h) { return h ? h->err.c_str() : "null handle"; }

int sr4d_set_option(sr4d_t* h, int option, int value) {
    if (!h) return SR4D_EINVAL;
    switch (option) {
        case SR4D_OPT_CONV_IMPL:
            if (value < 0 || value > 2) return fail(h, SR4D_EINVAL, "bad conv impl");
            h->conv_impl = value;
            return SR4D_OK;
        case SR4D_OPT_SAVE_ACTS:
            h->save_acts = value != 0;
            for (auto& kv : h->chains) tc_chain_free(kv.second);     // the layer -> buffer mapping may differ
            h->chains.clear();
            return SR4D_OK;
        case SR4D_OPT_PROFILE:
            h->profile = value != 0;
            h->ev_used = 0;
            h->ev_class.clear();
            h->ev_weight.clear();
            return SR4D_OK;
        case SR4D_OPT_FUSED_DGRAD:
            h->fused_dgrad = value != 0;
            return SR4D_OK;
        case SR4D_OPT_NVTX:
            h->nvtx = value != 0;
            return SR4D_OK;
        case SR4D_OPT_FWD_CHAIN:
            if (value < 0 || value > 64) return fail(h, SR4D_EINVAL, "SR4D_OPT_FWD_CHAIN takes 0..64 tiles per SM");
            h->fwd_chain = value;
            return SR4D_OK;
        case SR4D_OPT_DGRAD_SINGLE:
            h->dgrad_single = value != 0;
            return SR4D_OK;
        case SR4D_OPT_WGRAD_SINGLE:
            if (value < 0 || value > 2) return fail(h, SR4D_EINVAL, "SR4D_OPT_WGRAD_SINGLE takes 0, 1 or 2");
            h->wgrad_single = value;
            return SR4D_OK;
    }
    return fail(h, SR4D_EINVAL, "unknown option");
}
int sr4d_get_option(const sr4d_t* h, int option, int* value) {
    if (!h || !value) return SR4D_EINVAL;
    if (option == SR4D_OPT_CONV_IMPL) { *value = h->conv_impl; return SR4D_OK; }
    if (option == SR4D_OPT_SAVE_ACTS) { *value = h->save_acts; return SR4D_OK; }
    if (option == SR4D_OPT_PROFILE) { *value = h->profile; return SR4D_OK; }
    if (option == SR4D_OPT_FUSED_DGRAD) { *value = h->fused_dgrad; return SR4D_OK; }
    if (option == SR4D_OPT_NVTX) { *value = h->nvtx; return SR4D_OK; }
    if (option == SR4D_OPT_FWD_CHAIN) { *value = h->fwd_chain >= 0 ? h->fwd_chain : 4; return SR4D_OK; }
    if (option == SR4D_OPT_DGRAD_SINGLE) { *value = h->dgrad_single; return SR4D_OK; }
    if (option == SR4D_OPT_WGRAD_SINGLE) { *value = h->wgrad_single; return SR4D_OK; }
    return SR4D_EINVAL;
}

int64_t sr4d_param_count(const sr4d_t* h) { return h ? h->nparam : 0; }
int64_t sr4d_flat_size(const sr4d_t* h) { return h ? h->flat : 0; }
int64_t sr4d_grads_size(const sr4d_t* h) { return h && h->training ? h->flat + SR4D_METRIC_TAIL : 0; }
int sr4d_num_tensors(const sr4d_t* h) { return h ? (int)h->table.size() : 0; }
int sr4d_param_table(const sr4d_t* h, sr4d_tensor_desc* out, int capacity) {
    if (!h || !out) return SR4D_EINVAL;
    int n = (int)h->table.size();
    if (capacity < n) return SR4D_EINVAL;
    memcpy(out, h->table.data(), n * sizeof(sr4d_tensor_desc));
    return n;
}
float* sr4d_params(sr4d_t* h) { return h ? h->params : nullptr; }
float* sr4d_grads(sr4d_t* h) { return h ? h->grads : nullptr; }
float* sr4d_adam_m(sr4d_t* h) { return h ? h->m : nullptr; }
float* sr4d_adam_v(sr4d_t* h) { return h ? h->v : nullptr; }
int sr4d_params_changed(sr4d_t* h, void* stream) {
    if (!h) return SR4D_EINVAL;
    h->tcw_dirty = true;
    return ensure_tc_weights(h, (cudaStream_t)stream);
}

int sr4d_forward(sr4d_t* h, const float* u, const float* v, const float* w, const float* um, const float* vm,
                 const float* wm, float* out, int B, void* stream) {
    if (!h || !u || !v || !w || !um || !vm || !wm || !out) return fail(h, SR4D_EINVAL, "null argument");
    cudaSetDevice(h->device);
    return forward_impl(h, u, v, w, um, vm, wm, out, B, (cudaStream_t)stream);
}

int sr4d_loss_metrics(sr4d_t* h, const float* pred, const float* hu, const float* hv, const float* hw,
                      const float* mask, int B, float* per_sample, void* stream) {
    if (!h || !pred || !hu || !hv || !hw || !mask || !per_sample) return fail(h, SR4D_EINVAL, "null argument");
    if (B < 1 || B > h->maxB) return fail(h, SR4D_EINVAL, "batch size out of range");
    cudaStream_t s = (cudaStream_t)stream;
    CK(h, launch_loss_stats(pred, hu, hv, hw, mask, B, h->H * h->H * h->H, h->dpartial, 64, per_sample, h->norm, s), 2);
    return SR4D_OK;
}

int sr4d_train_fwd_bwd(sr4d_t* h, const float* u, const float* v, const float* w, const float* um, const float* vm,
                       const float* wm, const float* hu, const float* hv, const float* hw, const float* mask, int B,
                       float* per_sample, float* l2_out, float* pred_out, void* stream) {
    if (!h || !u || !v || !w || !um || !vm || !wm || !hu || !hv || !hw || !mask || !per_sample)
        return fail(h, SR4D_EINVAL, "null argument");
    if (!h->training) return fail(h, SR4D_ESTATE, "handle was not created for training");
    cudaSetDevice(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = forward_impl(h, u, v, w, um, vm, wm, h->pred, B, s);
    if (rc) return rc;
    if (pred_out)
        CK(h, cudaMemcpyAsync(pred_out, h->pred, (size_t)B * h->H * h->H * h->H * 3 * sizeof(float),
                              cudaMemcpyDeviceToDevice, s), 0);
    return backward_impl(h, hu, hv, hw, mask, B, per_sample, l2_out, s);
}

int sr4d_train_forward(sr4d_t* h, const float* u, const float* v, const float* w, const float* um, const float* vm,
                       const float* wm, int B, float* pred_out, void* stream) {
    if (!h || !u || !v || !w || !um || !vm || !wm) return fail(h, SR4D_EINVAL, "null argument");
    if (!h->training) return fail(h, SR4D_ESTATE, "handle was not created for training");
    cudaSetDevice(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = forward_impl(h, u, v, w, um, vm, wm, h->pred, B, s);
    if (rc) return rc;
    if (pred_out)
        CK(h, cudaMemcpyAsync(pred_out, h->pred, (size_t)B * h->H * h->H * h->H * 3 * sizeof(float),
                              cudaMemcpyDeviceToDevice, s), 0);
    return SR4D_OK;
}

int sr4d_train_backward(sr4d_t* h, const float* hu, const float* hv, const float* hw, const float* mask, int B,
                        float* per_sample, float* l2_out, void* stream) {
    if (!h || !hu || !hv || !hw || !mask || !per_sample) return fail(h, SR4D_EINVAL, "null argument");
    if (!h->training) return fail(h, SR4D_ESTATE, "handle was not created for training");
    if (!h->have_fwd_state || h->fwd_batch != B)
        return fail(h, SR4D_ESTATE, "sr4d_train_backward needs a preceding sr4d_train_forward of the same batch");
    cudaSetDevice(h->device);
    int rc = ensure_tc_weights(h, (cudaStream_t)stream);   // the forward may have run with the SIMT kernels
    if (rc) return rc;
    return backward_impl(h, hu, hv, hw, mask, B, per_sample, l2_out, (cudaStream_t)stream);
}

static int adam_impl(sr4d_t* h, float lr, float beta1, float beta2, float eps, int64_t t, float l2_grad_scale,
                     const float* count_dev, cudaStream_t s) {
    if (!h->training) return fail(h, SR4D_ESTATE, "handle was not created for training");
    if (t < 1) return fail(h, SR4D_EINVAL, "t must be >= 1 (iterations + 1)");
    cudaSetDevice(h->device);
    NvtxPhase nvtx_phase(h, "sr4d adam");
    const double alpha = (double)lr * std::sqrt(1.0 - std::pow((double)beta2, (double)t)) /
                         (1.0 - std::pow((double)beta1, (double)t));
    CK(h, launch_adam(h->params, h->grads, h->m, h->v, h->kflag, h->flat, (float)alpha, beta1, beta2, eps,
                      l2_grad_scale, count_dev, s), 1);
    h->tcw_dirty = true;
    return ensure_tc_weights(h, s);
}

int sr4d_adam_step(sr4d_t* h, float lr, float beta1, float beta2, float eps, int64_t t, float l2_grad_scale,
                   void* stream) {
    if (!h) return SR4D_EINVAL;
    return adam_impl(h, lr, beta1, beta2, eps, t, l2_grad_scale, nullptr, (cudaStream_t)stream);
}

int sr4d_adam_step_counted(sr4d_t* h, float lr, float beta1, float beta2, float eps, int64_t t,
                           float l2_grad_per_sample, int tail_index, void* stream) {
    if (!h) return SR4D_EINVAL;
    if (tail_index < 0 || tail_index >= SR4D_METRIC_TAIL) return fail(h, SR4D_EINVAL, "tail_index out of range");
    if (!h->training) return fail(h, SR4D_ESTATE, "handle was not created for training");
    return adam_impl(h, lr, beta1, beta2, eps, t, l2_grad_per_sample, h->grads + h->flat + tail_index,
                     (cudaStream_t)stream);
}

int sr4d_stitch(sr4d_t* h, const float* pred, int nx, int ny, int nz, int side_pad_hr, int VX, int VY, int VZ,
                float venc, int round_small, float* vol_out, void* stream) {
    if (!h || !pred || !vol_out) return fail(h, SR4D_EINVAL, "null argument");
    const int core = h->H - 2 * side_pad_hr;
    if (core <= 0 || VX > nx * core || VY > ny * core || VZ > nz * core || VX < 1 || VY < 1 || VZ < 1)
        return fail(h, SR4D_EINVAL, "stitch geometry inconsistent");
    CK(h, launch_stitch(pred, nx, ny, nz, h->H, side_pad_hr, VX, VY, VZ, venc, round_small, vol_out,
                        (cudaStream_t)stream), 1);
    return SR4D_OK;
}

// ---- single-layer entry points ----------------------------------------------------------
int sr4d_conv64_layer(sr4d_t* h, const float* x, const float* kernel, const float* bias, const float* residual,
                      float act_slope, float* y, int B, int D, int impl, void* stream) {
    if (!h || !x || !kernel || !y || B < 1 || D < 2) return fail(h, SR4D_EINVAL, "bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ActBuf bi, bo, br;
    int rc = SR4D_OK;
    if (alloc_act(bi, B, D) || alloc_act(bo, B, D) || (residual && alloc_act(br, B, D))) rc = SR4D_ENOMEM;
    TcWeights* tw = nullptr;
    if (!rc) {
        do {
            ActView vi = bi.view(B), vo = bo.view(B), vr = br.view(B);
            cudaError_t e = launch_pack_act(x, vi, s);
            if (!e && residual) e = launch_pack_act(residual, vr, s);
            if (e) { rc = SR4D_ECUDA; h->err = cudaGetErrorString(e); break; }
            h->launches += residual ? 2 : 1;
            if (impl == SR4D_CONV_TCGEN05) {
                if (tc_alloc_weights(&tw, 1) != cudaSuccess) { rc = SR4D_ENOMEM; break; }
                { const int l0 = 0; const long long o0 = 0; e = tc_prepare_weights(tw, kernel, &l0, &o0, 1, s); }
                TcConvArgs a;
                a.in = vi; a.out = vo; a.layer = 0; a.dgrad = 0; a.bias = bias;
                a.res_hi = residual ? vr.hi : nullptr; a.res_lo = residual ? vr.lo : nullptr;
                a.slope = act_slope; a.halo = 1;
                if (!e) e = tc_conv64(tw, a, s);
                h->launches += 3;
            } else {
                Conv64Args a;
                a.in_hi = vi.hi; a.in_lo = vi.lo; a.B = B; a.Do = D; a.w = kernel;
                a.out_hi = vo.hi; a.out_lo = vo.lo; a.halo = 1; a.bias = bias;
                if (residual) { a.res_hi = vr.hi; a.res_lo = vr.lo; }
                a.slope = act_slope;
                e = launch_conv64_simt(a, s);
                h->launches += 1;
            }
            if (!e) e = launch_unpack_act(vo, y, s);
            h->launches += 1;
            if (!e) e = cudaStreamSynchronize(s);
            if (e) { rc = SR4D_ECUDA; h->err = cudaGetErrorString(e); }
        } while (0);
    }
    cudaFree(bi.base); cudaFree(bo.base); cudaFree(br.base);
    if (tw) tc_free_weights(tw);
    return rc;
}

int sr4d_upsample_layer(sr4d_t* h, const float* x, float* y, int B, int D, int r, void* stream) {
    if (!h || !x || !y || B < 1) return fail(h, SR4D_EINVAL, "bad argument");
    if (D != h->P || r != h->r || r == 1) return fail(h, SR4D_EINVAL, "upsample layer uses the handle's patch_size/res_increase (>1)");
    cudaStream_t s = (cudaStream_t)stream;
    ActBuf bi, bo;
    int rc = SR4D_OK;
    if (alloc_act(bi, B, D) || alloc_act(bo, B, D * r)) rc = SR4D_ENOMEM;
    if (!rc) {
        ActView vi = bi.view(B), vo = bo.view(B);
        cudaError_t e = launch_pack_act(x, vi, s);
        if (!e) e = launch_upsample(vi, vo, r, h->up.tables(), s);
        if (!e) e = launch_unpack_act(vo, y, s);
        if (!e) e = cudaStreamSynchronize(s);
        h->launches += 3;
        if (e) { rc = SR4D_ECUDA; h->err = cudaGetErrorString(e); }
    }
    cudaFree(bi.base); cudaFree(bo.base);
    return rc;
}

int sr4d_conv64_layer_bwd(sr4d_t* h, const float* x, const float* kernel, const float* dy, float* dx,
                          float* dkernel, float* dbias, int B, int D, int impl, void* stream) {
    if (!h || !x || !kernel || !dy || B < 1 || D < 2) return fail(h, SR4D_EINVAL, "bad argument");
    const bool tc = impl == SR4D_CONV_TCGEN05;
    cudaStream_t s = (cudaStream_t)stream;
    ActBuf bi;
    float *g4 = nullptr, *raw = nullptr, *g4o = nullptr, *scr = nullptr, *dwb = nullptr;
    __half* g4s = nullptr;
    int* meta = nullptr;
    TcWeights* tw = nullptr;
    const size_t n4 = (size_t)B * (D + 4) * (D + 4) * (D + 4) * 64, n2 = (size_t)B * (D + 2) * (D + 2) * (D + 2) * 64;
    const int nchunk = 16;
    int rc = SR4D_OK;
    if (alloc_act(bi, B, D) || dmalloc(&g4, n4) || dmalloc(&raw, n2) || dmalloc(&g4o, n4) ||
        dmalloc(&scr, (size_t)std::max(nchunk, std::max(tc_wgrad_slabs(B, D), tc_wgrad2_slabs(B, D))) * 27 * 4096 + 1184 * 64) || dmalloc(&dwb, 27 * 4096 + 64) || dmalloc(&meta, 2) ||
        (tc && (dmalloc(&g4s, 2 * n4) || tc_alloc_weights(&tw, 1) != cudaSuccess)))
        rc = SR4D_ENOMEM;
    if (!rc) {
        cudaMemsetAsync(g4, 0, n4 * 4, s);
        cudaMemsetAsync(g4o, 0, n4 * 4, s);
        cudaMemsetAsync(meta, 0, 2 * sizeof(int), s);
        if (tc) cudaMemsetAsync(g4s, 0, 2 * n4 * sizeof(__half), s);
        ActView vi = bi.view(B);
        cudaError_t e = launch_pack_act(x, vi, s);
        if (!e) e = launch_g4_from_dense(dy, g4, reinterpret_cast<unsigned int*>(meta), B, D, s);
        if (!e && dx) {
            const int* rexp = nullptr;
            if (tc) {
                e = launch_g4_split(g4, reinterpret_cast<unsigned int*>(meta), g4s, meta + 1, B, D, s);
                if (!e) { const int l0 = 0; const long long o0 = 0; e = tc_prepare_weights(tw, kernel, &l0, &o0, 1, s); }
                TcConvArgs a;
                a.in.hi = g4s; a.in.lo = g4s + act_plane_elems(B, D + 2); a.in.B = B; a.in.D = D + 2;
                a.layer = 0; a.dgrad = 1; a.out_raw = raw; a.single_b = h->dgrad_single;
                if (!e) e = tc_conv64(tw, a, s);
                rexp = meta + 1;
                h->launches += 3;
            } else {
                Conv64Args a;
                a.in_f32 = g4; a.B = B; a.Do = D + 2; a.w = kernel; a.dgrad = 1; a.out_raw = raw;
                e = launch_conv64_simt(a, s);
            }
            if (!e) e = launch_fold_act(raw, nullptr, nullptr, rexp, nullptr, nullptr, nullptr, nullptr, nullptr, 1.f,
                                        g4o, nullptr, B, D, s);
            if (!e) e = launch_dense_from_g4(g4o, dx, B, D, s);
            h->launches += 3;
        }
        if (!e && dkernel && tc) {
            if (!dx) e = launch_g4_split(g4, reinterpret_cast<unsigned int*>(meta), g4s, meta + 1, B, D, s);
            // the handle's SR4D_OPT_WGRAD_SINGLE picks the kernel, as in the training path
            if (!e) e = h->wgrad_single == 1 ? tc_wgrad64_single(vi, g4s, meta + 1, scr, s)
                                             : tc_wgrad64(vi, g4s, meta + 1, scr, s, h->wgrad_single == 2);
            if (!e) e = launch_reduce_rows(scr, h->wgrad_single == 1 ? tc_wgrad2_slabs(B, D) : tc_wgrad_slabs(B, D),
                                           27 * 4096, dwb, s);
            if (!e) e = cudaMemcpyAsync(dkernel, dwb, 27 * 4096 * 4, cudaMemcpyDeviceToDevice, s);
            h->launches += 3;
        } else if (!e && dkernel) {
            e = launch_wgrad64_simt(vi, g4, dwb, scr, nchunk, s);
            if (!e) e = cudaMemcpyAsync(dkernel, dwb, 27 * 4096 * 4, cudaMemcpyDeviceToDevice, s);
            h->launches += 2;
        }
        if (!e && dbias) {
            e = launch_bias_grad(g4, B, D, dwb + 27 * 4096, scr, s);
            if (!e) e = cudaMemcpyAsync(dbias, dwb + 27 * 4096, 64 * 4, cudaMemcpyDeviceToDevice, s);
            h->launches += 2;
        }
        if (!e) e = cudaStreamSynchronize(s);
        if (e) { rc = SR4D_ECUDA; h->err = cudaGetErrorString(e); }
    }
    cudaFree(bi.base); cudaFree(g4); cudaFree(raw); cudaFree(g4o); cudaFree(scr); cudaFree(dwb); cudaFree(g4s);
    cudaFree(meta);
    if (tw) tc_free_weights(tw);
    return rc;
}

int sr4d_head_layer_bwd(sr4d_t* h, const float* x, const float* kernel, const float* g, int c, float* dx, float* dkernel,
                        float* dbias, float* dbias_prev, int B, int D, int impl, void* stream) {
    if (!h || !x || !kernel || !g || !dx || !dkernel || !dbias || !dbias_prev || B < 1 || D < 1 || c < 0 || c > 2)
        return fail(h, SR4D_EINVAL, "bad argument");
    const bool tc = impl == SR4D_CONV_TCGEN05;
    if (tc && !head_bwd_tc_supported(D)) return fail(h, SR4D_EINVAL, "tensor-core head backward: unsupported edge");
    cudaStream_t s = (cudaStream_t)stream;
    ActBuf bi;
    float *g4 = nullptr, *scr = nullptr, *outs = nullptr, *part = nullptr, *gpl = nullptr;
    __half *g4s = nullptr, *wimg = nullptr;
    void* scales = nullptr;
    int* meta = nullptr;
    const size_t n4 = (size_t)B * (D + 4) * (D + 4) * (D + 4) * 64;
    const size_t ng = (size_t)B * D * D * D * 3;
    int rc = SR4D_OK;
    if (alloc_act(bi, B, D) || dmalloc(&g4, n4) || dmalloc(&g4s, 2 * n4) || dmalloc(&scr, (size_t)(592 + 1) * 29 * 64) ||
        dmalloc(&outs, 27 * 64 + 1 + 64) || dmalloc(&meta, 4) ||
        (tc && (dmalloc(&part, 3 * head_bwd_tc_partial_floats(B, D)) || dmalloc(&gpl, head_bwd_tc_gplanar_floats(B, D)) || dmalloc(&wimg, head_bwd_tc_wimg_halves()) ||
                cudaMalloc(&scales, head_bwd_tc_scales_bytes()) != cudaSuccess)))
        rc = SR4D_ENOMEM;
    if (!rc) {
        cudaMemsetAsync(g4, 0, n4 * 4, s);
        cudaMemsetAsync(g4s, 0, 2 * n4 * sizeof(__half), s);
        cudaMemsetAsync(meta, 0, 4 * sizeof(int), s);
        ActView vi = bi.view(B);
        unsigned int* amax = reinterpret_cast<unsigned int*>(meta);
        unsigned int* gmax = reinterpret_cast<unsigned int*>(meta + 2);
        cudaError_t e = launch_pack_act(x, vi, s);
        if (!e) e = launch_absmax(g, ng, gmax, s);
        float* dw = outs; float* db = outs + 27 * 64; float* db1 = db + 1;
        if (!e && tc) {
            e = launch_head_bwd_tc_setup(kernel, kernel, kernel, gmax, scales, wimg, g, gpl, B, D, s);
            const size_t ps = head_bwd_tc_partial_floats(B, D);
            const bool lean = lo_plane_dead(h);   // the handle's SR4D_OPT_DGRAD_SINGLE / WGRAD_SINGLE pick the kernel, as in training
            if (!e) e = launch_head_bwd_tc(vi, gpl, c, scales, wimg, part + c * ps, g4s, meta + 1, amax, lean, s);
            // the finishing kernel covers three heads: the other two get this head's partials and scratch outputs
            const float* pp[3] = {part + c * ps, part + c * ps, part + c * ps};
            const int nc = head_bwd_tc_grid(B, D);
            const int ncta[3] = {nc, nc, nc};
            float* odw[3] = {scr, scr, scr}; float* odb[3] = {scr + 2000, scr + 2000, scr + 2000};
            float* odb1[3] = {scr + 2100, scr + 2100, scr + 2100};
            odw[c] = dw; odb[c] = db; odb1[c] = db1;
            if (!e) e = launch_head_bwd_tc_finish(pp, ncta, scales, odw, odb, odb1, lean, s);
            if (!e) e = launch_dense_from_split(g4s, meta + 1, !lean, dx, B, D, s);
            h->launches += 5;
        } else if (!e) {
            e = launch_head2_bwd(vi, g, c, kernel, g4, amax, dw, db, db1, nullptr, meta + 1, gmax, scr, s, false);
            if (!e) e = launch_dense_from_g4(g4, dx, B, D, s);
            h->launches += 6;
        }
        if (!e) e = cudaMemcpyAsync(dkernel, dw, 27 * 64 * 4, cudaMemcpyDeviceToDevice, s);
        if (!e) e = cudaMemcpyAsync(dbias, db, 4, cudaMemcpyDeviceToDevice, s);
        if (!e) e = cudaMemcpyAsync(dbias_prev, db1, 64 * 4, cudaMemcpyDeviceToDevice, s);
        if (!e) e = cudaStreamSynchronize(s);
        if (e) { rc = SR4D_ECUDA; h->err = cudaGetErrorString(e); }
    }
    cudaFree(bi.base); cudaFree(g4); cudaFree(g4s); cudaFree(scr); cudaFree(outs); cudaFree(meta);
    cudaFree(part); cudaFree(wimg); cudaFree(scales); cudaFree(gpl);
    return rc;
}

int sr4d_profile_read(sr4d_t* h, double* ms, int64_t* launches, int nclasses) {
    if (!h || !ms || !launches || nclasses < SR4D_PROF_NCLASSES) return fail(h, SR4D_EINVAL, "bad argument");
    for (int i = 0; i < nclasses; ++i) { ms[i] = 0.0; launches[i] = 0; }
    if (h->ev_used) {
        cudaError_t e = cudaEventSynchronize(h->ev_pool[h->ev_used - 1]);
        if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return SR4D_ECUDA; }
    }
    for (size_t i = 0; i < h->ev_class.size(); ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]) != cudaSuccess) continue;
        ms[h->ev_class[i]] += t;
        launches[h->ev_class[i]] += h->ev_weight[i];
    }
    h->ev_used = 0;
    h->ev_class.clear();
    h->ev_weight.clear();
    return SR4D_OK;
}

int sr4d_activation_overflow(sr4d_t* h, int* flag, int reset, void* stream) {
    if (!h || !flag) return SR4D_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned int v = 0;
    cudaSetDevice(h->device);
    CK(h, cudaMemcpyAsync(&v, h->ovf, sizeof v, cudaMemcpyDeviceToHost, s), 0);
    CK(h, cudaStreamSynchronize(s), 0);
    if (reset && v) CK(h, cudaMemsetAsync(h->ovf, 0, sizeof v, s), 0);
    *flag = v != 0;
    return SR4D_OK;
}

int64_t sr4d_launch_count(const sr4d_t* h) { return h ? h->launches : 0; }
void sr4d_reset_launch_count(sr4d_t* h) { if (h) h->launches = 0; }

}  // extern "C"
