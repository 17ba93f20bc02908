"""Mirror of the reference's Network/loss_utils.py for the one function the training path uses,
`calculate_relative_error` (:64-103, called from TrainerController.accuracy_function :143-150): the masked relative
speed error in percent, per sample.  It runs in libsr4d's loss/metric kernel (`sr4d_loss_metrics`); the divergence
helpers of the reference file (:4-62) are dead code there (`div_weight = 0`, TrainerController.py:23,121) and are not
mirrored."""
import numpy as np
import torch

from ..engine import Engine

_ENGINES = {}


def _metrics_engine(H, B, device):
    """A minimal handle (no residual blocks, res_increase 1) whose only job is to own the metric kernel's workspace."""
    key = (int(H), torch.device(device).index if device is not None else torch.cuda.current_device())
    eng = _ENGINES.get(key)
    if eng is None or eng.max_batch < B:
        eng = Engine(H, 1, 0, 0, max_batch=max(int(B), 8), training=False, device=key[1])
        _ENGINES[key] = eng
    return eng


def _as4(a):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a)
    return a[..., 0] if a.ndim == 5 else a


def calculate_relative_error(u_pred, v_pred, w_pred, u_hi, v_hi, w_hi, binary_mask):
    """(B,) relative speed error in % -- arguments are (B,H,H,H) or (B,H,H,H,1) arrays / tensors, as in the reference."""
    up, vp, wp, uh, vh, wh, mk = (_as4(a) for a in (u_pred, v_pred, w_pred, u_hi, v_hi, w_hi, binary_mask))
    B, H = up.shape[0], up.shape[1]
    dev = up.device if up.is_cuda else None
    eng = _metrics_engine(H, B, dev)
    pred = torch.stack([t.to(eng.device, torch.float32) for t in (up, vp, wp)], dim=-1).contiguous()
    per = eng.loss_metrics(pred, uh, vh, wh, mk)
    return per[:, 2]
