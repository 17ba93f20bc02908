"""CPU emulation of the operand precisions a tensor-core 64->64 convolution could use (runs here, no GPU).

The thirty 64->64 3x3x3 layers of the network are evaluated with both operands rounded to TF32 / BF16 / FP16 (single
pass) or split into two FP16 parts with three products (what csrc/conv_tc.cu does: x = hi + lo/2048, fp32 accumulation),
all other layers in fp32, and the full-network output is compared with the float64 oracle.  This is the evidence behind
the choice of the split path: the north-star bar is 1e-4 of max|ref|.

    python tools/precision_emulation.py [P r low hi]        # default 24 2 8 4 (about a minute)
"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
oracle = importlib.import_module("oracle.sr4d_oracle")


def round_mantissa(x, bits):
    """Round-to-nearest-even of fp32 values to `bits` explicit mantissa bits (TF32: 10)."""
    i = x.contiguous().view(torch.int32)
    drop = 23 - bits
    bias = ((i >> drop) & 1) + ((1 << (drop - 1)) - 1)
    return ((i + bias) & ~((1 << drop) - 1)).view(torch.float32)


def split16(x):
    hi = x.half().float()
    lo = ((x - hi) * 2048.0).half().float()
    return hi, lo


MODES = {
    "fp32": lambda x, w, conv: conv(x, w),
    "tf32 (1 pass)": lambda x, w, conv: conv(round_mantissa(x, 10), round_mantissa(w, 10)),
    "bf16 (1 pass)": lambda x, w, conv: conv(x.bfloat16().float(), w.bfloat16().float()),
    "fp16 (1 pass)": lambda x, w, conv: conv(x.half().float(), w.half().float()),
    "fp16 split, 2 products (no Wlo*Xhi)": lambda x, w, conv: (lambda xs, ws: conv(xs[0], ws[0]) + conv(xs[1], ws[0]) / 2048.0)(split16(x), split16(w)),
    "fp16 split, 2 products (no Whi*Xlo)": lambda x, w, conv: (lambda xs, ws: conv(xs[0], ws[0]) + conv(xs[0], ws[1]) / 2048.0)(split16(x), split16(w)),
    "fp16 split, 3 products (conv_tc.cu)": lambda x, w, conv: (lambda xs, ws: conv(xs[0], ws[0]) + (conv(xs[0], ws[1]) + conv(xs[1], ws[0])) / 2048.0)(split16(x), split16(w)),
}


def run(P=24, r=2, low=8, hi=4, seed=1234, modes=None):
    params = oracle.glorot_params(low, hi, seed=seed, bias_scale=0.02)
    batch = oracle.synthetic_batch(1, P, r, seed=0)
    with torch.no_grad():
        ref = oracle.forward({k: torch.tensor(v, dtype=torch.float64) for k, v in params.items()},
                             [torch.tensor(b, dtype=torch.float64) for b in batch[:6]], r, low, hi).numpy()
    p32 = {k: torch.tensor(v) for k, v in params.items()}
    x32 = [torch.tensor(b) for b in batch[:6]]
    plain = oracle.conv3d
    out = {}
    for name, fn in MODES.items():
        if modes and name not in modes:
            continue

        def conv3d(x, kernel, bias, activation=None, _fn=fn):
            if tuple(kernel.shape) != (3, 3, 3, 64, 64) or x.dtype != torch.float32:
                return plain(x, kernel, bias, activation)
            y = _fn(x, kernel, lambda a, b: plain(a, b, None, None))
            if bias is not None:
                y = y + bias
            return torch.relu(y) if activation == "relu" else y
        oracle.conv3d = conv3d
        try:
            with torch.no_grad():
                y = oracle.forward(p32, x32, r, low, hi).numpy()
        finally:
            oracle.conv3d = plain
        out[name] = float(np.abs(y - ref).max() / np.abs(ref).max())
    return out


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:5]]
    res = run(*a) if a else run()
    print(f"max|d|/max|ref| of the full-network output vs the float64 oracle (bar: 1e-4)")
    for k, v in res.items():
        print(f"  {k:40s} {v:.2e}  {'ok' if v < 1e-4 else 'MISSES the bar'}")
