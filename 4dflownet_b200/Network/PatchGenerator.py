"""Inference tiling with the reference's PatchGenerator interface (Network/PatchGenerator.py:6-154):
zero-pad 2 voxels, pad the far side to fit, stride patch_size-4, crop 2*r per side on HR.
Integer logic, bit-exact with the reference (pinned by tests/golden/patchgen_golden.npz);
written on strided views instead of the reference's triple Python loop, and the stitch can
run on the GPU (`Engine.stitch`) to avoid the host round trip."""
import numpy as np


class PatchGenerator:
    def __init__(self, patch_size, res_increase):
        self.patch_size = patch_size
        self.effective_patch_size = patch_size - 4
        self.res_increase = res_increase
        self.padding = (0, 0, 0)
        self.nr_x = self.nr_y = self.nr_z = 0

    # -- geometry ---------------------------------------------------------------------------
    def _plan(self, shape):
        e = self.effective_patch_size
        side = (self.patch_size - e) // 2
        far, nr = [], []
        for d in shape:
            padded = d + 2 * side
            rem = padded % e
            extra = (self.patch_size - rem) if rem > 2 * side else (2 * side - rem)
            far.append(extra)
            nr.append((padded + extra - 2 * side) // e)
        return side, tuple(far), tuple(nr)

    def _pad_to_patch_size_with_overlap(self, img):
        side, far, _ = self._plan(img.shape)
        self.padding = tuple(f * self.res_increase for f in far)
        return np.pad(img, [(side, side + f) for f in far], "constant")

    def _generate_overlapping_patches(self, img, lo=None, hi=None):
        """All patches of `img` in the reference's x-major order, or only patches [lo, hi) of that list (what one
        rank of a sharded prediction needs: the windows are views, only the selected ones are copied)."""
        side, far, nr = self._plan(img.shape)
        padded = self._pad_to_patch_size_with_overlap(img)
        P, e = self.patch_size, self.effective_patch_size
        win = np.lib.stride_tricks.sliding_window_view(padded, (P, P, P))[::e, ::e, ::e]
        win = win[:nr[0], :nr[1], :nr[2]]
        if lo is None:
            return np.ascontiguousarray(win.reshape(-1, P, P, P)), nr[0], nr[1], nr[2]
        ix, iy, iz = np.unravel_index(np.arange(lo, hi), nr)
        return np.ascontiguousarray(win[ix, iy, iz]), nr[0], nr[1], nr[2]

    def patchify(self, dataset, lo=None, hi=None):
        stacks = []
        for img in (dataset.u, dataset.v, dataset.w, dataset.mag_u, dataset.mag_v, dataset.mag_w):
            s, i, j, k = self._generate_overlapping_patches(img, lo, hi)
            stacks.append(s[..., None])
        self.nr_x, self.nr_y, self.nr_z = i, j, k
        return tuple(stacks[:3]), tuple(stacks[3:])

    def patchify_device(self, dataset, device, lo=None, hi=None):
        """`patchify` with the copying done on `device`: the six volumes go up once (one page-locked staging buffer,
        one async copy), are zero-padded there and the (P,P,P) windows are gathered by a strided view -- the same
        integer plan, so the result equals `patchify` bit for bit.  Returns six (n,P,P,P) float32 tensors
        (u, v, w, mag_u, mag_v, mag_w); `lo, hi` select patches [lo, hi) of the x-major list (one rank's shard).
        On the host the six window copies of a 160x160x64 volume cost ~80 ms, more than a B200 needs for the
        forward pass of the patches of one rank."""
        import torch
        import torch.nn.functional as F
        imgs = (dataset.u, dataset.v, dataset.w, dataset.mag_u, dataset.mag_v, dataset.mag_w)
        shape = imgs[0].shape
        side, far, nr = self._plan(shape)
        self.padding = tuple(f * self.res_increase for f in far)
        self.nr_x, self.nr_y, self.nr_z = nr
        dev = torch.device(device)
        host = torch.empty((6,) + tuple(shape), dtype=torch.float32, pin_memory=dev.type == "cuda")
        for k, img in enumerate(imgs):
            host[k].copy_(torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)))
        vol = host.to(dev, non_blocking=True)
        padded = F.pad(vol, (side, side + far[2], side, side + far[1], side, side + far[0]))
        P, e = self.patch_size, self.effective_patch_size
        win = padded.unfold(1, P, e).unfold(2, P, e).unfold(3, P, e)[:, :nr[0], :nr[1], :nr[2]]
        if lo is None:
            sel = win.reshape(6, -1, P, P, P)
        else:
            ix, iy, iz = (torch.from_numpy(a).to(dev) for a in np.unravel_index(np.arange(lo, hi), nr))
            sel = win[:, ix, iy, iz]
        return tuple(sel[k].contiguous() for k in range(6))

    def count_patches(self, shape):
        nr = self._plan(shape)[2]
        return nr[0] * nr[1] * nr[2]

    def unpatchify(self, results):
        return tuple(self._patchup_with_overlap(results[..., c], self.nr_x, self.nr_y, self.nr_z) for c in range(3))

    def _patchup_with_overlap(self, patches, x, y, z):
        side_hr = (self.patch_size - self.effective_patch_size) // 2 * self.res_increase
        n = patches.shape[1] - side_hr
        core = patches[:, side_hr:n, side_hr:n, side_hr:n]
        c = core.shape[1]
        vol = core.reshape(x, y, z, c, c, c).transpose(0, 3, 1, 4, 2, 5).reshape(x * c, y * c, z * c)
        px, py, pz = self.padding
        return vol[:vol.shape[0] - px, :vol.shape[1] - py, :vol.shape[2] - pz]

    def stitched_shape(self):
        c = self.effective_patch_size * self.res_increase
        return (self.nr_x * c - self.padding[0], self.nr_y * c - self.padding[1], self.nr_z * c - self.padding[2])
