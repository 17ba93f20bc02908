"""Drop-in mirror of the reference's Network/SR4DFlowNet.py (class SR4DFlowNet, :4-51)
and of the Keras Model object that predictor.prepare_network (:11-29) and
TrainerController (:35-49) build from it.  The graph itself runs in libsr4d
(hand-written sm_100a kernels); this file only adapts the Keras object protocol
(`predict`, `__call__`, `load_weights`, `save`, `get_weights`, `trainable_variables`).
"""
import os

import numpy as np
import torch

from ..engine import Engine


class SR4DFlowModel:
    """What `tf.keras.Model(input_layer, prediction)` is to the reference's callers."""

    def __init__(self, patch_size, res_increase, low_resblock=8, hi_resblock=4, max_batch=8, training=False,
                 device=None, seed=None):
        self.patch_size, self.res_increase = patch_size, res_increase
        self.low_resblock, self.hi_resblock = low_resblock, hi_resblock
        self.engine = Engine(patch_size, res_increase, low_resblock, hi_resblock, max_batch, training, device)
        self.initialize(seed)

    # Keras default initialisation: glorot_uniform kernels, zero biases (conv3d() passes
    # kernel_initializer=None, SR4DFlowNet.py:104)
    def initialize(self, seed=None):
        g = torch.Generator(device="cpu")
        if seed is not None:
            g.manual_seed(int(seed))
        ws = []
        for name, _, _, shape, is_kernel in self.engine.table:
            if is_kernel:
                k3 = shape[0] * shape[1] * shape[2]
                lim = float(np.sqrt(6.0 / (k3 * shape[3] + k3 * shape[4])))
                ws.append(((torch.rand(shape, generator=g) * 2 - 1) * lim).numpy())
            else:
                ws.append(np.zeros(shape, dtype=np.float32))
        self.engine.set_weights(ws)

    # ---- Keras Model protocol ----
    @property
    def trainable_variables(self):
        return [v for _, v in self.engine.tensor_views()]

    trainable_weights = trainable_variables

    @property
    def variable_names(self):
        return [n for n, *_ in self.engine.table]

    def get_weights(self):
        return self.engine.get_weights()

    def set_weights(self, weights):
        self.engine.set_weights(weights)

    def count_params(self):
        return self.engine.param_count

    def __call__(self, inputs, training=False):
        """model(inputs, training=...) (TrainerController.py:217,234): returns a CUDA tensor."""
        return self.engine.forward(inputs)

    def predict(self, inputs, batch_size=None):
        """model.predict([u,v,w,u_mag,v_mag,w_mag]) (predictor.py:87-92): numpy in, numpy out.
        Batches are staged through two sets of page-locked buffers so the host copies of batch k+1 / k-1 overlap
        the kernels of batch k (all device work stays on torch's current stream)."""
        n = len(inputs[0])
        eng = self.engine
        bs = eng.max_batch if batch_size is None else min(int(batch_size), eng.max_batch)
        P, H = eng.patch_size, eng.H
        out = np.empty((n, H, H, H, 3), dtype=np.float32)
        if n == 0:
            return out
        st = self._staging(bs)
        stream = torch.cuda.current_stream(eng.device)
        pending = [None, None]                       # (event, lo, hi) of the batch occupying staging set i

        def drain(i):
            if pending[i] is not None:
                ev, lo, hi = pending[i]
                ev.synchronize()
                out[lo:hi] = st[i]["out_host"][:hi - lo].numpy()
                pending[i] = None
        for k, lo in enumerate(range(0, n, bs)):
            i = k & 1
            hi = min(lo + bs, n)
            drain(i)
            hin = st[i]["in_host"]
            for c in range(6):
                hin[c, :hi - lo] = torch.from_numpy(np.ascontiguousarray(inputs[c][lo:hi], dtype=np.float32).reshape(hi - lo, P, P, P))
            st[i]["in_dev"][:, :hi - lo].copy_(hin[:, :hi - lo], non_blocking=True)
            y = eng.forward([st[i]["in_dev"][c, :hi - lo] for c in range(6)], out=st[i]["out_dev"][:hi - lo])
            st[i]["out_host"][:hi - lo].copy_(y, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            pending[i] = (ev, lo, hi)
        drain(0)
        drain(1)
        return out

    def _staging(self, bs):
        st = getattr(self, "_stage", None)
        if st is None or st[0]["in_host"].shape[1] < bs:
            eng = self.engine
            P, H = eng.patch_size, eng.H
            st = [{"in_host": torch.empty((6, bs, P, P, P), dtype=torch.float32, pin_memory=True),
                   "in_dev": torch.empty((6, bs, P, P, P), dtype=torch.float32, device=eng.device),
                   "out_dev": torch.empty((bs, H, H, H, 3), dtype=torch.float32, device=eng.device),
                   "out_host": torch.empty((bs, H, H, H, 3), dtype=torch.float32, pin_memory=True)} for _ in range(2)]
            self._stage = st
        return st

    # ---- weights on disk ----
    def save_weights(self, path):
        ws = dict(zip(self.variable_names, self.get_weights()))
        if path.endswith(".h5"):
            from ..utils import h5io
            h5io.save_keras_weights(path, ws)
        else:
            np.savez(path, **{k.replace("/", "__"): v for k, v in ws.items()})

    save = save_weights

    def load_weights(self, path):
        if path.endswith(".h5"):
            from ..utils import h5io
            ws = h5io.load_keras_weights(path, self.variable_names)
        else:
            if not os.path.exists(path) and os.path.exists(path + ".npz"):
                path = path + ".npz"
            z = np.load(path)
            ws = {k.replace("__", "/"): z[k] for k in z.files}
        self.engine.set_weights(ws)


class SR4DFlowNet:
    """Same constructor and build_network signature as the reference (SR4DFlowNet.py:4-7).
    The reference wires symbolic Keras tensors; here build_network evaluates eagerly on
    (B,P,P,P,1) arrays / CUDA tensors and returns the (B,rP,rP,rP,3) prediction."""

    def __init__(self, res_increase):
        self.res_increase = res_increase
        self.model = None

    def build_model(self, patch_size, low_resblock=8, hi_resblock=4, max_batch=8, training=False, device=None,
                    seed=None):
        self.model = SR4DFlowModel(patch_size, self.res_increase, low_resblock, hi_resblock, max_batch, training,
                                   device, seed)
        return self.model

    def build_network(self, u, v, w, u_mag, v_mag, w_mag, low_resblock=8, hi_resblock=4, channel_nr=64):
        channel_nr = 64   # the reference overwrites it too (SR4DFlowNet.py:8)
        patch_size = int(u.shape[1])
        if (self.model is None or self.model.patch_size != patch_size or self.model.low_resblock != low_resblock
                or self.model.hi_resblock != hi_resblock):
            self.build_model(patch_size, low_resblock, hi_resblock, max_batch=max(8, int(u.shape[0])))
        return self.model([u, v, w, u_mag, v_mag, w_mag])
