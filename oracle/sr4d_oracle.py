"""CPU oracle for the 4DFlowNet patch-based super-resolution hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module.  The product path (``4dflownet_b200``) never does: it fails loudly when
its CUDA library is missing.

PARITY STATUS of the floating-point graph.  The reference's arithmetic lives in
TensorFlow 2.2 / Keras (README.md:4), which is not vendored under /root/reference
and is not installable in this image (no tensorflow, keras, h5py; no network); the
reference ships no tests, golden outputs or weights.
* PINNED to the reference's own Python code: forward graph wiring, ``upsample3d``
  choreography, loss, relative-error metric, L2 term, running means and the Keras
  variable creation order.  ``tests/golden/make_graph_golden.py`` imports
  ``Network/SR4DFlowNet.py``, ``loss_utils.py`` and ``TrainerController.py``
  unmodified and executes them in float64 on a numpy stand-in for the ~30 TF
  symbols they call (``tests/golden/tf_numpy_shim.py``); this file agrees with
  the resulting ``tests/golden/graph_golden.npz`` to 1e-12
  (``tests/test_graph_golden.py``).
  The gradient of the training objective (sum over the batch of loss_b + l2, as
  ``calculate_and_update_metrics(..., 'train')`` assembles it) is checked against
  central finite differences taken through the reference code (stored in the
  same file): 2e-6.
* RESTATED, NOT PINNED ("parity unpinned" for these): the semantics of the TF
  primitives themselves (Conv3D, tf.pad SYMMETRIC, resize_bilinear with
  align_corners, LeakyReLU, tf.round), that ``tape.gradient`` of a vector target
  differentiates its sum, and Keras Adam.
  Self-checks in ``tests/test_oracle.py``: naive numpy convolution, a literal
  two-pass restatement of ``upsample3d``, finite differences, a hand-computed
  Adam example.
* PINNED: the integer tiling / data-path logic (PatchGenerator, PatchHandler3D):
  the reference's own numpy code generated ``tests/golden/patch*_golden.npz``.

Every function cites the reference file:line it follows (paths relative to
/root/reference/src).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

L2_COEFF = 5e-7          # Network/SR4DFlowNet.py:99  tf.keras.regularizers.l2(5e-7)
LEAKY_SLOPE = 0.2        # Network/SR4DFlowNet.py:113,118
CHANNELS = 64            # Network/SR4DFlowNet.py:8 (channel_nr overwritten to 64)


# ----------------------------------------------------------------------------
# parameter table (Keras creation order of
# Network/SR4DFlowNet.py:17-46)
# ----------------------------------------------------------------------------
def param_table(low_resblock: int = 8, hi_resblock: int = 4) -> List[Tuple[str, Tuple[int, ...]]]:
    """Return [(name, shape)] in Keras ``trainable_variables`` order.

    Kernel shape is Keras' (kx, ky, kz, Cin, Cout); layers are auto-named
    conv3d, conv3d_1, ... in construction order.
    """
    C = CHANNELS
    layers: List[Tuple[Tuple[int, ...], bool]] = []
    layers.append(((3, 3, 3, 3, C), True))      # :17 pc stem 1
    layers.append(((3, 3, 3, C, C), True))      # :18 pc stem 2
    layers.append(((3, 3, 3, 3, C), True))      # :20 phase stem 1
    layers.append(((3, 3, 3, C, C), True))      # :21 phase stem 2
    layers.append(((1, 1, 1, 2 * C, C), True))  # :24 1x1 fuse
    layers.append(((3, 3, 3, C, C), True))      # :25 3x3 fuse
    for _ in range(low_resblock):               # :29-30
        layers.append(((3, 3, 3, C, C), False))
        layers.append(((3, 3, 3, C, C), False))
    for _ in range(hi_resblock):                # :35-36
        layers.append(((3, 3, 3, C, C), False))
        layers.append(((3, 3, 3, C, C), False))
    for _ in range(3):                          # :39-46 three heads
        layers.append(((3, 3, 3, C, C), True))
        layers.append(((3, 3, 3, C, 1), True))
    table = []
    for i, (shape, bias) in enumerate(layers):
        lname = "conv3d" if i == 0 else f"conv3d_{i}"
        table.append((f"{lname}/kernel", shape))
        if bias:
            table.append((f"{lname}/bias", (shape[-1],)))
    return table


def glorot_params(low_resblock=8, hi_resblock=4, seed=1234, bias_scale=0.0,
                  dtype=np.float32) -> Dict[str, np.ndarray]:
    """Keras default init: glorot_uniform kernels, zero biases
    (Network/SR4DFlowNet.py:104 passes kernel_initializer=None).  ``bias_scale``>0
    gives small random biases so bias paths are exercised in tests."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in param_table(low_resblock, hi_resblock):
        if name.endswith("kernel"):
            k3 = shape[0] * shape[1] * shape[2]
            lim = math.sqrt(6.0 / (k3 * shape[3] + k3 * shape[4]))
            out[name] = rng.uniform(-lim, lim, size=shape).astype(dtype)
        else:
            out[name] = (bias_scale * rng.standard_normal(shape)).astype(dtype)
    return out


# ----------------------------------------------------------------------------
# layer primitives
# ----------------------------------------------------------------------------
def conv3d(x: torch.Tensor, kernel: torch.Tensor, bias, activation=None) -> torch.Tensor:
    """Network/SR4DFlowNet.py:93-108.  ``x`` is (B,X,Y,Z,Ci) channels-last,
    ``kernel`` Keras (kx,ky,kz,Ci,Co).  tf.pad(...,'SYMMETRIC') with p=(k-1)//2
    (p=1 == edge replicate; p=0 for k=1) then a VALID cross-correlation."""
    k = kernel.shape[0]
    p = (k - 1) // 2
    xc = x.permute(0, 4, 1, 2, 3)
    if p > 0:
        # symmetric padding with p=1 repeats the edge voxel == replicate
        assert p == 1
        xc = F.pad(xc, (p, p, p, p, p, p), mode="replicate")
    w = kernel.permute(4, 3, 0, 1, 2)
    y = F.conv3d(xc, w, bias)
    if activation == "relu":
        y = torch.relu(y)
    elif activation is not None:
        raise ValueError(activation)
    return y.permute(0, 2, 3, 4, 1)


def leaky_relu(x):
    """tf.keras.layers.LeakyReLU(alpha=0.2), Network/SR4DFlowNet.py:113,118."""
    return torch.where(x > 0, x, x * LEAKY_SLOPE)


def resnet_block(x, ka, kb):
    """Network/SR4DFlowNet.py:111-120 (scale = 1, no bias)."""
    t = conv3d(x, ka, None)
    t = leaky_relu(t)
    t = conv3d(t, kb, None)
    t = x + t * 1
    return leaky_relu(t)


def _resize_weights(in_size: int, out_size: int):
    """TF1 legacy ``resize_bilinear(align_corners=True)`` scaler, fp32:
    scale=(in-1)/(out-1); src=dst*scale; lo=floor; hi=min(ceil,in-1); lerp=src-lo."""
    scale = np.float32(in_size - 1) / np.float32(out_size - 1) if out_size > 1 else np.float32(0)
    dst = np.arange(out_size, dtype=np.float32)
    src = dst * scale                      # fp32 product, as TF computes it
    lo = np.floor(src).astype(np.int64)
    hi = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
    lerp = (src - lo.astype(np.float32)).astype(np.float32)
    return lo, hi, lerp


def _lerp_axis(x: torch.Tensor, axis: int, out_size: int) -> torch.Tensor:
    lo, hi, lerp = _resize_weights(x.shape[axis], out_size)
    a = x.index_select(axis, torch.from_numpy(lo))
    b = x.index_select(axis, torch.from_numpy(hi))
    shape = [1] * x.dim()
    shape[axis] = out_size
    t = torch.from_numpy(lerp).to(x.dtype).reshape(shape)
    return a + (b - a) * t                 # TF: top + (bottom - top) * lerp


def upsample3d(x: torch.Tensor, res_increase: int) -> torch.Tensor:
    """Network/SR4DFlowNet.py:53-90: separable trilinear, align_corners=True;
    pass 1 resizes (y,z) [z lerp inside rows, then y], pass 2 resizes x."""
    if res_increase == 1:
        return x
    r = res_increase
    _, X, Y, Z, _ = x.shape
    x = _lerp_axis(x, 3, Z * r)
    x = _lerp_axis(x, 2, Y * r)
    x = _lerp_axis(x, 1, X * r)
    return x


def upsample3d_literal(x: torch.Tensor, res_increase: int) -> torch.Tensor:
    """Literal restatement of the reshape/resize/transpose sequence of
    Network/SR4DFlowNet.py:77-89 with a hand-written 2-D legacy bilinear resize;
    used only to validate :func:`upsample3d`."""
    if res_increase == 1:
        return x
    B, X, Y, Z, C = x.shape
    r = res_increase

    def resize2d(img, oh, ow):   # img (N,H,W,C)
        ylo, yhi, yl = _resize_weights(img.shape[1], oh)
        xlo, xhi, xl = _resize_weights(img.shape[2], ow)
        yl_t = torch.from_numpy(yl).to(img.dtype).reshape(1, oh, 1, 1)
        xl_t = torch.from_numpy(xl).to(img.dtype).reshape(1, 1, ow, 1)
        top, bot = img[:, ylo], img[:, yhi]
        tl, tr = top[:, :, xlo], top[:, :, xhi]
        bl, br = bot[:, :, xlo], bot[:, :, xhi]
        t = tl + (tr - tl) * xl_t
        b = bl + (br - bl) * xl_t
        return t + (b - t) * yl_t

    s = x.reshape(-1, Y, Z, C)
    s = resize2d(s, Y * r, Z * r).reshape(-1, X, Y * r, Z * r, C)
    s = s.permute(0, 3, 2, 1, 4)
    s = s.reshape(-1, Y * r, X, C)
    s = resize2d(s, Y * r, X * r).reshape(-1, Z * r, Y * r, X * r, C)
    return s.permute(0, 3, 2, 1, 4)


# ----------------------------------------------------------------------------
# the network (Network/SR4DFlowNet.py:7-51)
# ----------------------------------------------------------------------------
def _as_param_list(params, low_resblock, hi_resblock):
    names = [n for n, _ in param_table(low_resblock, hi_resblock)]
    return [params[n] for n in names], names


def forward(params: Dict[str, torch.Tensor], inputs: Sequence[torch.Tensor], res_increase: int,
            low_resblock: int = 8, hi_resblock: int = 4, return_intermediates: bool = False):
    """inputs = [u, v, w, u_mag, v_mag, w_mag], each (B,P,P,P,1).
    Returns (B, rP, rP, rP, 3)."""
    u, v, w, u_mag, v_mag, w_mag = inputs
    inter = {}
    speed = (u ** 2 + v ** 2 + w ** 2) ** 0.5                     # :10
    mag = (u_mag ** 2 + v_mag ** 2 + w_mag ** 2) ** 0.5           # :11
    pcmr = mag * speed                                            # :12
    phase = torch.cat([u, v, w], dim=-1)                          # :14
    pc = torch.cat([pcmr, mag, speed], dim=-1)                    # :15

    def K(i):
        return params[("conv3d" if i == 0 else f"conv3d_{i}") + "/kernel"]

    def Bv(i):
        return params[("conv3d" if i == 0 else f"conv3d_{i}") + "/bias"]

    pc = conv3d(pc, K(0), Bv(0), "relu")                          # :17
    pc = conv3d(pc, K(1), Bv(1), "relu")                          # :18
    phase = conv3d(phase, K(2), Bv(2), "relu")                    # :20
    phase = conv3d(phase, K(3), Bv(3), "relu")                    # :21
    cat = torch.cat([phase, pc], dim=-1)                          # :23
    x = conv3d(cat, K(4), Bv(4), "relu")                          # :24
    x = conv3d(x, K(5), Bv(5), "relu")                            # :25
    inter["fuse"] = x
    li = 6
    for _ in range(low_resblock):                                 # :28-30
        x = resnet_block(x, K(li), K(li + 1))
        li += 2
    inter["lr_trunk"] = x
    x = upsample3d(x, res_increase)                               # :32
    inter["upsampled"] = x
    for _ in range(hi_resblock):                                  # :35-36
        x = resnet_block(x, K(li), K(li + 1))
        li += 2
    inter["hr_trunk"] = x
    outs = []
    for _ in range(3):                                            # :39-46
        h = conv3d(x, K(li), Bv(li), "relu")
        o = conv3d(h, K(li + 1), Bv(li + 1), None)
        outs.append(o)
        li += 2
    out = torch.cat(outs, dim=-1)                                 # :49
    if return_intermediates:
        return out, inter
    return out


# ----------------------------------------------------------------------------
# loss / metric / regulariser (Network/TrainerController.py, Network/loss_utils.py)
# ----------------------------------------------------------------------------
def calculate_mse(y_true, y_pred):
    """TrainerController.py:152-156: sum over the 3 components (not a mean)."""
    d = y_pred - y_true
    return (d ** 2).sum(dim=-1)


def loss_function(y_true, y_pred, mask):
    """TrainerController.py:84-127.  Returns (total[B], mse[B], 0)."""
    mse = calculate_mse(y_true, y_pred)
    non_fluid = (mask < 0.5).to(mse.dtype)                         # :96-97
    eps = 1
    fluid = (mse * mask).sum(dim=(1, 2, 3)) / (mask.sum(dim=(1, 2, 3)) + eps)          # :101-102
    nonfl = (mse * non_fluid).sum(dim=(1, 2, 3)) / (non_fluid.sum(dim=(1, 2, 3)) + eps)  # :104-105
    mse_b = fluid + nonfl
    return mse_b + 0, mse_b, 0


def calculate_relative_error(y_true, y_pred, mask):
    """loss_utils.py:64-103 (tf.round = round-half-to-even = torch.round)."""
    eps = 1e-5
    diff = torch.sqrt(((y_pred - y_true) ** 2).sum(dim=-1))
    actual = torch.sqrt((y_true ** 2).sum(dim=-1))
    rel = diff / (actual + eps)
    rel = torch.clamp(rel, 0.0, 1.0)
    rel = torch.where(actual != 0, rel, diff)
    rel = torch.round(rel * 1e4) / 1e4
    rel = torch.where(mask == 1.0, rel, torch.zeros_like(rel))
    return rel.sum(dim=(1, 2, 3)) / (mask.sum(dim=(1, 2, 3)) + 1) * 100


def regularizer_loss(params):
    """TrainerController.py:129-141: sum over layers of 5e-7 * sum(w^2); only
    kernels carry a regulariser (SR4DFlowNet.py:104)."""
    tot = 0
    for n, p in params.items():
        if n.endswith("kernel"):
            tot = tot + L2_COEFF * (p ** 2).sum()
    return tot


def train_objective(params, batch, res_increase, low_resblock=8, hi_resblock=4):
    """TrainerController.py:209-257.  ``batch`` is the 11-tuple
    (u,v,w,u_mag,v_mag,w_mag,u_hr,v_hr,w_hr,venc,mask).  Returns
    (objective scalar = sum_b(loss_b + l2), dict of per-sample metrics)."""
    u, v, w, um, vm, wm, u_hr, v_hr, w_hr, venc, mask = batch
    hires = torch.cat((u_hr, v_hr, w_hr), dim=-1)                  # :212
    pred = forward(params, [u, v, w, um, vm, wm], res_increase, low_resblock, hi_resblock)
    loss, mse, _ = loss_function(hires, pred, mask)
    rel = calculate_relative_error(hires, pred, mask)
    l2 = regularizer_loss(params)
    total = loss + l2                                              # :249 scalar broadcast on (B,)
    # tape.gradient of a (B,) target differentiates its sum       # :223
    return total.sum(), {"loss": total.detach(), "mse": mse.detach(), "rel_err": rel.detach(),
                         "l2": l2.detach(), "pred": pred.detach()}


def gradients(params_np: Dict[str, np.ndarray], batch_np, res_increase, low_resblock=8,
              hi_resblock=4, dtype=torch.float64):
    params = {k: torch.tensor(v, dtype=dtype, requires_grad=True) for k, v in params_np.items()}
    batch = [torch.tensor(np.asarray(b), dtype=dtype) for b in batch_np]
    obj, metrics = train_objective(params, batch, res_increase, low_resblock, hi_resblock)
    obj.backward()
    grads = {k: p.grad.detach().numpy() for k, p in params.items()}
    return grads, {k: v.numpy() for k, v in metrics.items()}


def adam_step(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """tf.keras.optimizers.Adam (TF 2.2 ResourceApplyAdam, non-amsgrad), used at
    TrainerController.py:73,225.  ``t`` = iterations + 1.  Returns new (p, m, v)."""
    alpha = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m = m + (g - m) * (1.0 - beta1)
    v = v + (g * g - v) * (1.0 - beta2)
    p = p - alpha * m / (np.sqrt(v) + eps)
    return p, m, v


# ----------------------------------------------------------------------------
# naive numpy checks (tiny shapes only)
# ----------------------------------------------------------------------------
def conv3d_naive(x: np.ndarray, kernel: np.ndarray, bias=None) -> np.ndarray:
    """einsum restatement of clamp-padded cross-correlation; O(27) shifted views."""
    k = kernel.shape[0]
    p = (k - 1) // 2
    xp = np.pad(x, ((0, 0), (p, p), (p, p), (p, p), (0, 0)), mode="symmetric")
    B, X, Y, Z, _ = x.shape
    out = np.zeros((B, X, Y, Z, kernel.shape[-1]), dtype=np.float64)
    for a in range(k):
        for b in range(k):
            for c in range(k):
                out += np.einsum("bxyzi,io->bxyzo", xp[:, a:a + X, b:b + Y, c:c + Z, :].astype(np.float64),
                                 kernel[a, b, c].astype(np.float64))
    if bias is not None:
        out += bias
    return out


# ----------------------------------------------------------------------------
# integer tiling logic (Network/PatchGenerator.py:53-154) -- restatement used
# to check the GPU stitcher; pinned against tests/golden/patchgen_*.npz that the
# reference's own code produced.
# ----------------------------------------------------------------------------
def tiling_plan(shape, patch_size, res_increase):
    e = patch_size - 4
    s = (patch_size - e) // 2
    pads, nrs = [], []
    for d in shape:
        dim = d + 2 * s
        res = dim % e
        pad = patch_size - res if res > 2 * s else 2 * s - res      # :62-80
        dim += pad
        pads.append(pad)
        nrs.append((dim - 2 * s) // e)                                # :97-99
    return {"side_pad": s, "stride": e, "far_pad": tuple(pads), "nr": tuple(nrs),
            "hr_padding": tuple(p * res_increase for p in pads)}


def patchify(img: np.ndarray, patch_size: int):
    plan = tiling_plan(img.shape, patch_size, 1)
    s, e = plan["side_pad"], plan["stride"]
    img = np.pad(img, [(s, s + fp) for fp in plan["far_pad"]], "constant")
    nx, ny, nz = plan["nr"]
    out = []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                out.append(img[i * e:i * e + patch_size, j * e:j * e + patch_size, k * e:k * e + patch_size])
    return np.asarray(out), plan["nr"]


def patchup(patches: np.ndarray, shape_lr, patch_size, res_increase):
    plan = tiling_plan(shape_lr, patch_size, res_increase)
    c = plan["side_pad"] * res_increase
    n = patches.shape[1] - c
    nx, ny, nz = plan["nr"]
    core = n - c
    out = np.zeros((nx * core, ny * core, nz * core), dtype=patches.dtype)
    idx = 0
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                out[i * core:(i + 1) * core, j * core:(j + 1) * core, k * core:(k + 1) * core] = \
                    patches[idx, c:n, c:n, c:n]
                idx += 1
    px, py, pz = plan["hr_padding"]
    if px > 0:
        out = out[:-px]
    if py > 0:
        out = out[:, :-py]
    if pz > 0:
        out = out[:, :, :-pz]
    return out


# ----------------------------------------------------------------------------
# synthetic data
# ----------------------------------------------------------------------------
def synthetic_batch(B, patch_size, res_increase, seed=0):
    g = np.random.default_rng(seed)
    P, H = patch_size, patch_size * res_increase
    lr = [g.uniform(-1, 1, size=(B, P, P, P, 1)).astype(np.float32) for _ in range(3)]
    mg = [g.uniform(0, 0.016, size=(B, P, P, P, 1)).astype(np.float32) for _ in range(3)]
    mask = (g.uniform(size=(B, H, H, H)) < 0.12).astype(np.float32)
    hr = [(g.standard_normal((B, H, H, H, 1)).astype(np.float32) * 0.08) * mask[..., None] for _ in range(3)]
    venc = np.full((B,), 1.5, dtype=np.float32)
    return (*lr, *mg, *hr, venc, mask)
