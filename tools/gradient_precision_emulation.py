"""CPU emulation (runs here, no GPU): how accurate are the parameter gradients of the train step when the BACKWARD
64->64 convolutions (Conv3DBackpropInput / Conv3DBackpropFilter, 2/3 of the step's tensor work) use cheaper operand
precisions than the three-product FP16 split, with the forward pass kept at fp32 accuracy?

Reported per mode: relative L2 error of the flat gradient and the worst per-tensor relative L2 error against float64
autograd of the oracle.  fp32 autograd itself is the yardstick: ReLU / LeakyReLU gates make the gradient discontinuous in
the activations, so even fp32 is 1e-4 .. 5e-4 away from float64 on some tensors.

    python tools/gradient_precision_emulation.py [P r low hi B]      # default 12 2 2 2 2
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
oracle = importlib.import_module("oracle.sr4d_oracle")


def fp16_scaled(t):
    """Round to fp16 after scaling by a power of two that puts max|t| near 2^14 (what csrc does for gradients)."""
    m = float(t.abs().max())
    if m == 0.0:
        return t
    s = 2.0 ** (14 - int(np.floor(np.log2(m))))
    return (t * s).half().float() / s


def split16_scaled(t):
    m = float(t.abs().max())
    s = 1.0 if m == 0.0 else 2.0 ** (14 - int(np.floor(np.log2(m))))
    hi = (t * s).half().float()
    lo = ((t * s - hi) * 2048.0).half().float()
    return hi / s, lo / (s * 2048.0)


def make_conv(mode):
    class Conv(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):                     # x (B,Ci,X,Y,Z) already clamp-padded, w (Co,Ci,3,3,3)
            ctx.save_for_backward(x, w)
            return F.conv3d(x, w)

        @staticmethod
        def backward(ctx, gy):
            x, w = ctx.saved_tensors
            dgrad = lambda g, ww: torch.nn.grad.conv3d_input(x.shape, ww, g)        # noqa: E731
            wgrad = lambda xx, g: torch.nn.grad.conv3d_weight(xx, w.shape, g)       # noqa: E731
            if mode == "fp32":
                return dgrad(gy, w), wgrad(x, gy)
            if mode == "fp16 (1 pass)":
                g16, w16, x16 = fp16_scaled(gy), w.half().float(), x.half().float()
                return dgrad(g16, w16), wgrad(x16, g16)
            if mode == "fp16 split, 2 products (gradient split, W / X single)":
                gh, gl = split16_scaled(gy)
                w16, x16 = w.half().float(), x.half().float()
                return dgrad(gh, w16) + dgrad(gl, w16), wgrad(x16, gh) + wgrad(x16, gl)
            if mode == "fp16 split, 2 products (W / X split, gradient single)":
                g16 = fp16_scaled(gy)
                wh = w.half().float(); wl = w - wh
                xh = x.half().float(); xl = x - xh
                return dgrad(g16, wh) + dgrad(g16, wl.half().float()), wgrad(xh, g16) + wgrad(xl.half().float(), g16)
            if mode == "dgrad: W split x gradient single; wgrad: gradient split x X single":
                g16 = fp16_scaled(gy)
                gh, gl = split16_scaled(gy)
                wh = w.half().float(); wl = (w - wh).half().float()
                x16 = x.half().float()
                return dgrad(g16, wh) + dgrad(g16, wl), wgrad(x16, gh) + wgrad(x16, gl)
            if mode == "dgrad: W split x gradient single; wgrad: gradient single x X single":
                g16 = fp16_scaled(gy)
                wh = w.half().float(); wl = (w - wh).half().float()
                return dgrad(g16, wh) + dgrad(g16, wl), wgrad(x.half().float(), g16)
            if mode == "fp16 split, 3 products (csrc)":
                gh, gl = split16_scaled(gy)
                wh = w.half().float(); wl = ((w - wh) * 2048).half().float() / 2048
                xh = x.half().float(); xl = ((x - xh) * 2048).half().float() / 2048
                return dgrad(gh, wh) + dgrad(gl, wh) + dgrad(gh, wl), wgrad(xh, gh) + wgrad(xh, gl) + wgrad(xl, gh)
            raise ValueError(mode)
    return Conv.apply


MODES = ["fp32", "fp16 split, 3 products (csrc)", "fp16 split, 2 products (gradient split, W / X single)",
         "fp16 split, 2 products (W / X split, gradient single)",
         "dgrad: W split x gradient single; wgrad: gradient split x X single",
         "dgrad: W split x gradient single; wgrad: gradient single x X single", "fp16 (1 pass)"]


def run(P=12, r=2, low=2, hi=2, B=2, modes=MODES):
    params = oracle.glorot_params(low, hi, seed=11, bias_scale=0.02)
    batch = oracle.synthetic_batch(B, P, r, seed=5)
    ref, _ = oracle.gradients({k: v.astype(np.float64) for k, v in params.items()}, batch, r, low, hi)
    plain = oracle.conv3d
    out = {}
    for mode in modes:
        conv = make_conv(mode)

        def conv3d(x, kernel, bias, activation=None, _conv=conv):
            if tuple(kernel.shape) != (3, 3, 3, 64, 64) or x.dtype != torch.float32:
                return plain(x, kernel, bias, activation)
            xc = F.pad(x.permute(0, 4, 1, 2, 3), (1, 1, 1, 1, 1, 1), mode="replicate")
            y = _conv(xc, kernel.permute(4, 3, 0, 1, 2)).permute(0, 2, 3, 4, 1)
            if bias is not None:
                y = y + bias
            return torch.relu(y) if activation == "relu" else y
        oracle.conv3d = conv3d
        try:
            g, _ = oracle.gradients(params, batch, r, low, hi, dtype=torch.float32)
        finally:
            oracle.conv3d = plain
        flat = np.concatenate([(g[k].astype(np.float64) - ref[k]).ravel() for k in ref])
        rflat = np.concatenate([ref[k].ravel() for k in ref])
        worst = max(np.linalg.norm(g[k] - ref[k]) / (np.linalg.norm(ref[k]) + 1e-30) for k in ref)
        out[mode] = (float(np.linalg.norm(flat) / np.linalg.norm(rflat)), float(worst))
    return out


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:6]]
    res = run(*a) if a else run()
    print("backward 64->64 convolutions at reduced operand precision: gradient error vs float64 autograd")
    print(f"  {'mode':58s} {'flat rel-L2':>12s} {'worst tensor':>13s}")
    for k, (f, w) in res.items():
        print(f"  {k:58s} {f:12.2e} {w:13.2e}")
