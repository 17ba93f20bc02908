"""Gradient parity of the training path against the float64 oracle (GPU box).

For each requested geometry (default: BASELINE configs[1] geometry P=24, r=2, 8/4 blocks at B=1 and B=8) prints the
flat and per-tensor relative L2 error of the summed gradient (TrainerController.py:213-223: tape.gradient of the (B,)
loss vector) for
  * torch-CPU fp32 autograd of the oracle graph (what the reference's TF fp32 graph amounts to),
  * the engine's fp32 SIMT anchor,
  * the engine's default tensor-core (tcgen05 split-fp16) path,
  * the tensor-core BACKWARD fed the SIMT forward's saved activations (identical ReLU / LeakyReLU gates): separates
    kernel error from gate flips caused by the forward's rounding,
  * the two-plane backward (SR4D_OPT_DGRAD_SINGLE = SR4D_OPT_WGRAD_SINGLE = 0) and the old kernel's hi-only mode.
Writes a text table (committed as profiles/r02_grad_parity.txt).

usage: python tools/grad_parity.py [--cases P,r,low,hi,B;...] [--out FILE]
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("4dflownet_b200")
oracle = importlib.import_module("oracle.sr4d_oracle")
L = pkg._lib


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def engine_grads(P, r, low, hi, B, params, batch, fwd_impl, bwd_impl, dgrad_single=0, wgrad_single=0, fused=1):
    eng = pkg.Engine(P, r, low, hi, max_batch=B, training=True, device=0)
    eng.set_option(L.OPT_FUSED_DGRAD, fused)
    eng.set_option(L.OPT_DGRAD_SINGLE, dgrad_single)
    eng.set_option(L.OPT_WGRAD_SINGLE, wgrad_single)
    eng.set_weights(params)
    eng.set_option(L.OPT_CONV_IMPL, fwd_impl)
    pred = eng.train_forward(batch[:6], want_pred=True)
    eng.set_option(L.OPT_CONV_IMPL, bwd_impl)
    per, l2 = eng.train_backward([b[..., 0] for b in batch[6:9]], batch[10])
    torch.cuda.synchronize()
    g = {n: v.cpu().numpy().astype(np.float64) for n, v in eng.tensor_views(eng.grads)}
    out = (g, pred.cpu().numpy(), per.cpu().numpy())
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="24,2,8,4,1;24,2,8,4,8")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "grad_parity.txt"))
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args()
    lines = []

    def emit(s=""):
        print(s, flush=True)
        lines.append(s)

    emit("# gradient parity vs the float64 oracle (tools/grad_parity.py); rel-L2 = |g - g64| / |g64|")
    emit(f"# torch CPU threads {torch.get_num_threads()}, device {torch.cuda.get_device_name(0)}")
    for case in a.cases.split(";"):
        P, r, low, hi, B = (int(x) for x in case.split(","))
        params = oracle.glorot_params(low, hi, seed=a.seed, bias_scale=0.02)
        batch = oracle.synthetic_batch(B, P, r, seed=11)
        l2c = oracle.L2_COEFF
        t0 = time.time()
        g64, met = oracle.gradients({k: v.astype(np.float64) for k, v in params.items()}, batch, r, low, hi)
        t64 = time.time() - t0
        t0 = time.time()
        g32, _ = oracle.gradients(params, batch, r, low, hi, dtype=torch.float32)
        t32 = time.time() - t0
        names = [n for n, _ in oracle.param_table(low, hi)]
        # the engine's gradient carries no L2 share (it is folded into Adam): remove it from the oracle's
        corr = {n: (B * 2 * l2c * params[n].astype(np.float64) if n.endswith("kernel") else 0.0) for n in names}
        want = {n: g64[n] - corr[n] for n in names}
        variants = {"torch-cpu fp32 autograd": ({n: np.asarray(g32[n], np.float64) - corr[n] for n in names}, None)}
        specs = [
            ("engine SIMT fp32 anchor", dict(fwd_impl=L.CONV_SIMT, bwd_impl=L.CONV_SIMT)),
            ("engine tcgen05 (default)", dict(fwd_impl=L.CONV_AUTO, bwd_impl=L.CONV_AUTO, dgrad_single=1, wgrad_single=1)),
            ("tcgen05 bwd on SIMT fwd acts", dict(fwd_impl=L.CONV_SIMT, bwd_impl=L.CONV_AUTO, dgrad_single=1, wgrad_single=1)),
            ("SIMT bwd on tcgen05 fwd acts", dict(fwd_impl=L.CONV_AUTO, bwd_impl=L.CONV_SIMT)),
            ("tcgen05 two-plane backward", dict(fwd_impl=L.CONV_AUTO, bwd_impl=L.CONV_AUTO, dgrad_single=0, wgrad_single=0)),
            ("two-plane bwd on SIMT fwd acts", dict(fwd_impl=L.CONV_SIMT, bwd_impl=L.CONV_AUTO, dgrad_single=0, wgrad_single=0)),
            ("dgrad two-plane, wgrad default", dict(fwd_impl=L.CONV_SIMT, bwd_impl=L.CONV_AUTO, dgrad_single=0, wgrad_single=1)),
            ("old hi-only wgrad on SIMT acts", dict(fwd_impl=L.CONV_SIMT, bwd_impl=L.CONV_AUTO, dgrad_single=1, wgrad_single=2)),
        ]
        preds = {}
        for label, kw in specs:
            g, pred, per = engine_grads(P, r, low, hi, B, params, batch, **kw)
            variants[label] = (g, per)
            preds[label] = pred
        emit()
        emit(f"## P={P} r={r} low={low} hi={hi} B={B}  (oracle fp64 {t64:.1f} s, fp32 {t32:.1f} s on the host)")
        pm = np.abs(met["pred"]).max()
        for label in ("engine SIMT fp32 anchor", "engine tcgen05 (default)"):
            emit(f"forward  {label:32s} max|d|/max|ref| = {np.abs(preds[label] - met['pred']).max() / pm:.3e}")
        flat_want = np.concatenate([want[n].ravel() for n in names])
        emit(f"{'variant':34s} {'flat rel-L2':>12s} {'worst tensor':>13s}  (name)   {'median tensor':>13s}")
        for label, (g, per) in variants.items():
            flat = np.concatenate([np.asarray(g[n], np.float64).ravel() for n in names])
            per_t = {n: rel(g[n], want[n]) for n in names}
            worst = max(per_t, key=per_t.get)
            emit(f"{label:34s} {rel(flat, flat_want):12.3e} {per_t[worst]:13.3e}  {worst:18s} "
                 f"{float(np.median(list(per_t.values()))):13.3e}")
        # the decisive comparison: same activations (identical gates), different backward arithmetic
        fa = np.concatenate([variants["engine SIMT fp32 anchor"][0][n].ravel() for n in names])
        for label in ("tcgen05 bwd on SIMT fwd acts", "two-plane bwd on SIMT fwd acts", "dgrad two-plane, wgrad default",
                      "old hi-only wgrad on SIMT acts"):
            fb = np.concatenate([variants[label][0][n].ravel() for n in names])
            emit(f"identical gates: {label:32s} vs SIMT backward on the same activations: flat rel-L2 = {rel(fb, fa):.3e}")
        emit("per tensor (default tcgen05 | default bwd on SIMT acts | two-plane bwd on SIMT acts | torch fp32):")
        gd = variants["engine tcgen05 (default)"][0]
        gb = variants["tcgen05 bwd on SIMT fwd acts"][0]
        g2 = variants["two-plane bwd on SIMT fwd acts"][0]
        gt = variants["torch-cpu fp32 autograd"][0]
        for n in names:
            emit(f"  {n:22s} {rel(gd[n], want[n]):10.2e} {rel(gb[n], want[n]):10.2e} {rel(g2[n], want[n]):10.2e} "
                 f"{rel(gt[n], want[n]):10.2e}")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
