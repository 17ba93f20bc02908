"""Python handle over the C ABI: holds torch CUDA tensors, never computes on its own."""
import ctypes as C

import numpy as np
import torch

from . import _lib


class Sr4dError(RuntimeError):
    pass


class _DevArray:
    """Zero-copy view of handle-owned device memory through __cuda_array_interface__."""

    def __init__(self, ptr, n, owner):
        self._owner = owner   # keeps the handle alive
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """One engine per GPU per process.  Mirrors sr4d_create (include/sr4d.h), i.e. the
    constructor arguments of predictor.prepare_network / TrainerController.__init__."""

    def __init__(self, patch_size, res_increase, low_resblock=8, hi_resblock=4, max_batch=8, training=False,
                 device=None):
        if not torch.cuda.is_available():
            raise Sr4dError("no CUDA device visible: 4dflownet_b200 is B200-only and has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.patch_size, self.res_increase = int(patch_size), int(res_increase)
        self.low_resblock, self.hi_resblock = int(low_resblock), int(hi_resblock)
        self.max_batch, self.training = int(max_batch), bool(training)
        self.H = self.patch_size * self.res_increase
        h = C.c_void_p()
        rc = self.lib.sr4d_create(C.byref(h), self.patch_size, self.res_increase, self.low_resblock,
                                  self.hi_resblock, self.max_batch, int(self.training), self.device.index)
        if rc != 0:
            raise Sr4dError(f"sr4d_create failed: {_lib.ERRNAMES.get(rc, rc)}")
        self._h = h
        n = self.lib.sr4d_num_tensors(h)
        descs = (_lib.TensorDesc * n)()
        got = self.lib.sr4d_param_table(h, descs, n)
        assert got == n
        self.table = [(d.name.decode(), int(d.offset), int(d.count), tuple(d.shape[:d.ndim]), bool(d.is_kernel))
                      for d in descs]
        self.flat_size = int(self.lib.sr4d_flat_size(h))
        self.param_count = int(self.lib.sr4d_param_count(h))
        with torch.cuda.device(self.device):
            self.params = torch.as_tensor(_DevArray(self.lib.sr4d_params(h), self.flat_size, self), device=self.device)
            if self.training:
                # gradients + metric tail (include/sr4d.h SR4D_METRIC_TAIL): `grads_full` is what the data-parallel
                # all-reduce operates on, `grads` the flat_size gradient floats, `grad_tail` the caller-owned tail
                gsz = int(self.lib.sr4d_grads_size(h))
                self.grads_full = torch.as_tensor(_DevArray(self.lib.sr4d_grads(h), gsz, self), device=self.device)
                self.grads = self.grads_full[:self.flat_size]
                self.grad_tail = self.grads_full[self.flat_size:]
                self.adam_m = torch.as_tensor(_DevArray(self.lib.sr4d_adam_m(h), self.flat_size, self), device=self.device)
                self.adam_v = torch.as_tensor(_DevArray(self.lib.sr4d_adam_v(h), self.flat_size, self), device=self.device)

    # -- plumbing -------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.sr4d_last_error(self._h)
            raise Sr4dError(f"{what}: {_lib.ERRNAMES.get(rc, rc)}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.sr4d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev(self, a, shape=None):
        """float32 contiguous CUDA tensor on this engine's device (copies host data)."""
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        t = a.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        if shape is not None:
            t = t.reshape(shape)
        return t

    def set_option(self, opt, value):
        self._check(self.lib.sr4d_set_option(self._h, opt, value), "sr4d_set_option")

    def profile(self, on=True):
        self.set_option(_lib.OPT_PROFILE, int(bool(on)))

    def profile_read(self):
        """{class name: (device ms, launches)} since the last read (OPT_PROFILE must be on)."""
        n = len(_lib.PROF_CLASSES)
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        self._check(self.lib.sr4d_profile_read(self._h, ms, cnt, n), "sr4d_profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.PROF_CLASSES)}

    def activation_overflow(self, reset=True):
        """True when some activation had to be clamped to the split-fp16 range (+-65504) or was NaN since the last
        reset -- the fp32 reference would have carried the value on (synchronises the current stream)."""
        flag = C.c_int(0)
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_activation_overflow(self._h, C.byref(flag), int(bool(reset)), _stream_ptr(self.device))
        self._check(rc, "sr4d_activation_overflow")
        return bool(flag.value)

    def launch_count(self):
        return int(self.lib.sr4d_launch_count(self._h))

    def reset_launch_count(self):
        self.lib.sr4d_reset_launch_count(self._h)

    # -- parameters -------------------------------------------------------------------
    def tensor_views(self, flat=None):
        """[(name, view)] into a flat buffer, shaped like Keras' trainable_variables."""
        flat = self.params if flat is None else flat
        return [(n, flat[o:o + c].view(shape)) for n, o, c, shape, _ in self.table]

    def set_weights(self, weights):
        """weights: dict name->array or list in Keras order (model.set_weights)."""
        if isinstance(weights, dict):
            weights = [weights[n] for n, *_ in self.table]
        if len(weights) != len(self.table):
            raise ValueError(f"expected {len(self.table)} tensors, got {len(weights)}")
        views = self.tensor_views()
        staged = [torch.as_tensor(np.asarray(wv, dtype=np.float32)) for wv in weights]
        for (n, view), wv in zip(views, staged):          # validate everything before the first copy
            if tuple(wv.shape) != tuple(view.shape):
                raise ValueError(f"{n}: shape {tuple(wv.shape)} != {tuple(view.shape)} (lists must be in this "
                                 "package's table order = Keras layer-creation order)")
        for (n, view), wv in zip(views, staged):
            view.copy_(wv)
        self.params_changed()

    def get_weights(self):
        return [v.detach().cpu().numpy().copy() for _, v in self.tensor_views()]

    def params_changed(self):
        self._check(self.lib.sr4d_params_changed(self._h, _stream_ptr(self.device)), "sr4d_params_changed")

    # -- compute ------------------------------------------------------------------------
    def forward(self, inputs, out=None):
        """inputs: 6 arrays/tensors (B,P,P,P[,1]); returns a CUDA tensor (B,H,H,H,3)."""
        P = self.patch_size
        xs = [self._dev(a).reshape(-1, P, P, P) for a in inputs]
        B = xs[0].shape[0]
        if out is None:
            out = torch.empty((B, self.H, self.H, self.H, 3), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_forward(self._h, *[C.c_void_p(x.data_ptr()) for x in xs], C.c_void_p(out.data_ptr()), B,
                                       _stream_ptr(self.device))
        self._check(rc, "sr4d_forward")
        return out

    def loss_metrics(self, pred, hr_u, hr_v, hr_w, mask, per_out=None):
        H = self.H
        pred = self._dev(pred).reshape(-1, H, H, H, 3)
        B = pred.shape[0]
        t = [self._dev(a).reshape(B, H, H, H) for a in (hr_u, hr_v, hr_w, mask)]
        per = torch.empty((B, 4), device=self.device, dtype=torch.float32) if per_out is None else per_out
        if tuple(per.shape) != (B, 4) or not per.is_contiguous():
            raise ValueError("per_out must be a contiguous (B,4) tensor")
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_loss_metrics(self._h, C.c_void_p(pred.data_ptr()), *[C.c_void_p(x.data_ptr()) for x in t],
                                            B, C.c_void_p(per.data_ptr()), _stream_ptr(self.device))
        self._check(rc, "sr4d_loss_metrics")
        return per

    @staticmethod
    def _on_host(a):
        return isinstance(a, np.ndarray) or (torch.is_tensor(a) and a.device.type == "cpu")

    def _upload_targets(self, hr, mask, B):
        """Host-resident HR targets / mask (84 % of a step's input bytes) go up on a side stream into engine-owned
        staging buffers while the forward -- which does not read them -- runs on the caller's stream; returns the four
        device views and the event the backward has to wait for."""
        H = self.H
        if getattr(self, "_tgt_stage", None) is None:
            self._tgt_stage = torch.empty((4, self.max_batch, H, H, H), device=self.device, dtype=torch.float32)
            self._copy_stream = torch.cuda.Stream(self.device)
            self._copy_done = torch.cuda.Event()
        main = torch.cuda.current_stream(self.device)
        # the staging buffers were last read by the previous step's backward, enqueued on the caller's stream
        self._copy_stream.wait_stream(main)
        views = []
        with torch.cuda.stream(self._copy_stream):
            for k, a in enumerate(list(hr) + [mask]):
                if isinstance(a, np.ndarray):
                    a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
                dst = self._tgt_stage[k, :B]
                dst.copy_(a.to(torch.float32).reshape(B, H, H, H), non_blocking=True)
                views.append(dst)
            self._copy_done.record(self._copy_stream)
        return views, self._copy_done

    def train_fwd_bwd(self, inputs, hr, mask, want_pred=False, per_out=None, l2_out=None):
        """Returns (per_sample (B,4) [loss, mse, rel_err%, sum mask], l2 (1,), pred or None);
        gradients (sum over the batch, no L2 term) are left in self.grads.  per_out / l2_out: optional
        preallocated device tensors (e.g. views into self.grad_tail) that receive the metrics.

        When the HR targets and the mask arrive in host memory, the low-resolution inputs are uploaded on the
        caller's stream, the forward is enqueued (sr4d_train_forward), the targets follow on a side stream underneath
        it, and the backward (sr4d_train_backward) waits for that copy: same kernels as the single
        sr4d_train_fwd_bwd call, the bulk of the upload off the critical path."""
        P, H = self.patch_size, self.H
        host_targets = all(self._on_host(a) for a in list(hr) + [mask])
        xs = [self._dev(a).reshape(-1, P, P, P) for a in inputs]
        B = xs[0].shape[0]
        per = torch.empty((B, 4), device=self.device, dtype=torch.float32) if per_out is None else per_out
        l2 = torch.empty((1,), device=self.device, dtype=torch.float32) if l2_out is None else l2_out
        if tuple(per.shape) != (B, 4) or not per.is_contiguous() or l2.numel() != 1:
            raise ValueError("per_out must be a contiguous (B,4) tensor and l2_out a single float")
        pred = torch.empty((B, H, H, H, 3), device=self.device, dtype=torch.float32) if want_pred else None
        if host_targets and B <= self.max_batch:
            with torch.cuda.device(self.device):
                rc = self.lib.sr4d_train_forward(self._h, *[C.c_void_p(x.data_ptr()) for x in xs], B,
                                                 C.c_void_p(pred.data_ptr()) if want_pred else None,
                                                 _stream_ptr(self.device))
                self._check(rc, "sr4d_train_forward")
                tg, done = self._upload_targets(hr, mask, B)
                torch.cuda.current_stream(self.device).wait_event(done)
                rc = self.lib.sr4d_train_backward(self._h, *[C.c_void_p(t.data_ptr()) for t in tg], B,
                                                  C.c_void_p(per.data_ptr()), C.c_void_p(l2.data_ptr()),
                                                  _stream_ptr(self.device))
            self._check(rc, "sr4d_train_backward")
            return per, l2, pred
        ys = [self._dev(a).reshape(B, H, H, H) for a in hr]
        mk = self._dev(mask).reshape(B, H, H, H)
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_train_fwd_bwd(
                self._h, *[C.c_void_p(x.data_ptr()) for x in xs], *[C.c_void_p(y.data_ptr()) for y in ys],
                C.c_void_p(mk.data_ptr()), B, C.c_void_p(per.data_ptr()), C.c_void_p(l2.data_ptr()),
                C.c_void_p(pred.data_ptr()) if want_pred else None, _stream_ptr(self.device))
        self._check(rc, "sr4d_train_fwd_bwd")
        return per, l2, pred

    def train_forward(self, inputs, want_pred=False):
        """First half of train_fwd_bwd (the taped forward, TrainerController.py:213-217)."""
        P, H = self.patch_size, self.H
        xs = [self._dev(a).reshape(-1, P, P, P) for a in inputs]
        B = xs[0].shape[0]
        pred = torch.empty((B, H, H, H, 3), device=self.device, dtype=torch.float32) if want_pred else None
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_train_forward(self._h, *[C.c_void_p(x.data_ptr()) for x in xs], B,
                                             C.c_void_p(pred.data_ptr()) if want_pred else None,
                                             _stream_ptr(self.device))
        self._check(rc, "sr4d_train_forward")
        return pred

    def train_backward(self, hr, mask):
        """Second half (loss + tape.gradient, TrainerController.py:218-223) on the activations train_forward saved."""
        H = self.H
        mk = self._dev(mask).reshape(-1, H, H, H)
        B = mk.shape[0]
        ys = [self._dev(a).reshape(B, H, H, H) for a in hr]
        per = torch.empty((B, 4), device=self.device, dtype=torch.float32)
        l2 = torch.empty((1,), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_train_backward(self._h, *[C.c_void_p(y.data_ptr()) for y in ys],
                                              C.c_void_p(mk.data_ptr()), B, C.c_void_p(per.data_ptr()),
                                              C.c_void_p(l2.data_ptr()), _stream_ptr(self.device))
        self._check(rc, "sr4d_train_backward")
        return per, l2

    def adam_step(self, lr, t, l2_grad_scale, beta1=0.9, beta2=0.999, eps=1e-7):
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_adam_step(self._h, lr, beta1, beta2, eps, int(t), float(l2_grad_scale),
                                         _stream_ptr(self.device))
        self._check(rc, "sr4d_adam_step")

    def adam_step_counted(self, lr, t, l2_grad_per_sample, tail_index, beta1=0.9, beta2=0.999, eps=1e-7):
        """Adam with the regulariser scale l2_grad_per_sample * grad_tail[tail_index] read on the device."""
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_adam_step_counted(self._h, lr, beta1, beta2, eps, int(t), float(l2_grad_per_sample),
                                                 int(tail_index), _stream_ptr(self.device))
        self._check(rc, "sr4d_adam_step_counted")

    def stitch(self, pred, nr, vol_shape, side_pad_hr, venc, round_small=True):
        H = self.H
        pred = self._dev(pred).reshape(-1, H, H, H, 3)
        nx, ny, nz = (int(x) for x in nr)
        if pred.shape[0] != nx * ny * nz:
            raise ValueError("number of patches does not match nr_x*nr_y*nr_z")
        VX, VY, VZ = (int(x) for x in vol_shape)
        vol = torch.empty((3, VX, VY, VZ), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_stitch(self._h, C.c_void_p(pred.data_ptr()), nx, ny, nz, int(side_pad_hr), VX, VY, VZ,
                                      float(venc), int(bool(round_small)), C.c_void_p(vol.data_ptr()),
                                      _stream_ptr(self.device))
        self._check(rc, "sr4d_stitch")
        return vol

    # -- single layers (tests, profiling) ---------------------------------------------------
    def conv64_layer(self, x, kernel, bias=None, residual=None, act_slope=1.0, impl=_lib.CONV_SIMT):
        x = self._dev(x)
        B, D = x.shape[0], x.shape[1]
        k = self._dev(kernel).reshape(27, 64, 64)
        b = self._dev(bias) if bias is not None else None
        r = self._dev(residual) if residual is not None else None
        y = torch.empty_like(x)
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_conv64_layer(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(k.data_ptr()),
                                            C.c_void_p(b.data_ptr()) if b is not None else None,
                                            C.c_void_p(r.data_ptr()) if r is not None else None,
                                            float(act_slope), C.c_void_p(y.data_ptr()), B, D, impl,
                                            _stream_ptr(self.device))
        self._check(rc, "sr4d_conv64_layer")
        return y

    def upsample_layer(self, x):
        x = self._dev(x)
        B, D = x.shape[0], x.shape[1]
        r = self.res_increase
        y = torch.empty((B, D * r, D * r, D * r, 64), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_upsample_layer(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), B, D, r,
                                              _stream_ptr(self.device))
        self._check(rc, "sr4d_upsample_layer")
        return y

    def conv64_layer_bwd(self, x, kernel, dy, impl=_lib.CONV_SIMT):
        x, dy = self._dev(x), self._dev(dy)
        B, D = x.shape[0], x.shape[1]
        k = self._dev(kernel).reshape(27, 64, 64)
        dx = torch.empty_like(x)
        dk = torch.empty((3, 3, 3, 64, 64), device=self.device, dtype=torch.float32)
        db = torch.empty((64,), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_conv64_layer_bwd(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(k.data_ptr()),
                                                C.c_void_p(dy.data_ptr()), C.c_void_p(dx.data_ptr()),
                                                C.c_void_p(dk.data_ptr()), C.c_void_p(db.data_ptr()), B, D, impl,
                                                _stream_ptr(self.device))
        self._check(rc, "sr4d_conv64_layer_bwd")
        return dx, dk, db

    def head_layer_bwd(self, x, kernel, g, c, impl=_lib.CONV_SIMT):
        """Whole backward of one 64->1 head conv (sr4d_head_layer_bwd): x (B,D,D,D,64) = its saved post-ReLU input,
        kernel (3,3,3,64,1), g (B,D,D,D,3) loss gradient (channel c is this head's) -> dx, dkernel (27,64), dbias (1),
        dbias_prev (64)."""
        x, g = self._dev(x), self._dev(g)
        B, D = x.shape[0], x.shape[1]
        k = self._dev(kernel).reshape(27, 64)
        dx = torch.empty_like(x)
        dk = torch.empty((27, 64), device=self.device, dtype=torch.float32)
        db = torch.empty((1,), device=self.device, dtype=torch.float32)
        db1 = torch.empty((64,), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.sr4d_head_layer_bwd(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(k.data_ptr()),
                                              C.c_void_p(g.data_ptr()), int(c), C.c_void_p(dx.data_ptr()),
                                              C.c_void_p(dk.data_ptr()), C.c_void_p(db.data_ptr()),
                                              C.c_void_p(db1.data_ptr()), B, D, impl, _stream_ptr(self.device))
        self._check(rc, "sr4d_head_layer_bwd")
        return dx, dk, db, db1
