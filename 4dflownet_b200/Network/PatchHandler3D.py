"""Training-side patch iterator with the reference's interface (Network/PatchHandler3D.py:5-164).

The reference builds a tf.data pipeline (from_tensor_slices -> shuffle -> map(py_function) -> batch -> prefetch,
:20-47); here the same per-row loader (`load_patches_from_index_file`, :49-81) is driven by a thread pool that
keeps `prefetch` batches in flight and hands out the 11-tuple
    (u, v, w, mag_u, mag_v, mag_w, u_hr, v_hr, w_hr, venc, mask)
as numpy arrays (optionally page-locked torch tensors for async host->device copies).  Integer slice logic,
normalisation and the rotation augmentation (incl. its component swaps / sign flips, :166-274) follow the
reference exactly and are pinned by tests/golden/patchhandler_golden.npz.
"""
import concurrent.futures as cf

import numpy as np

from ..utils import h5io

# (plane, k) -> (source component for u, v, w; sign applied to phase images), then np.rot90(k, axes).
# Restates rotate90 / rotate180_3d (PatchHandler3D.py:166-274): e.g. plane 1 rotates the spatial axes (0, 1) yet
# swaps the v / w components -- kept as the reference has it.
_AXES = {1: (0, 1), 2: (0, 2), 3: (1, 2)}
_COMPONENTS = {
    (1, 1): ((0, 1), (2, 1), (1, -1)), (1, 2): ((0, 1), (1, -1), (2, -1)), (1, 3): ((0, 1), (2, -1), (1, 1)),
    (2, 1): ((2, -1), (1, 1), (0, 1)), (2, 2): ((0, -1), (1, 1), (2, -1)), (2, 3): ((2, 1), (1, 1), (0, -1)),
    (3, 1): ((1, -1), (0, 1), (2, 1)), (3, 2): ((0, -1), (1, -1), (2, 1)), (3, 3): ((1, 1), (0, -1), (2, 1)),
}


def apply_rotation(u, v, w, rotation_idx, plane_nr, is_phase_image):
    """PatchHandler3D.apply_rotation (:97-108): rotation_idx 1/2/3 = 90/180/270 degrees in plane 1/2/3."""
    key = (int(plane_nr), int(rotation_idx))
    if key not in _COMPONENTS:
        return u, v, w                      # unspecified plane / angle: unchanged, as in the reference
    src = (u, v, w)
    out = []
    for comp, sign in _COMPONENTS[key]:
        a = src[comp]
        if is_phase_image and sign < 0:
            a = -a
        out.append(np.rot90(a, k=key[1], axes=_AXES[key[0]]))
    return tuple(out)


def rotate_object(img, rotation_idx, plane_nr):
    """PatchHandler3D.rotate_object (:83-95): plain spatial rotation (used for the mask)."""
    if int(plane_nr) not in _AXES:
        return img
    return np.rot90(img, k=int(rotation_idx), axes=_AXES[int(plane_nr)])


class _BatchedDataset:
    """Re-iterable (one pass per epoch) batched view over the index rows."""

    def __init__(self, handler, indexes, shuffle, n_parallel, seed=None, shard=None):
        self.h, self.rows, self.shuffle = handler, list(indexes), shuffle
        self.n_parallel = n_parallel or 4
        self.rng = np.random.default_rng(seed)
        self.shard = shard              # (rank, world): load only this rank's contiguous slice of every global batch

    def __len__(self):
        n, bs = len(self.rows), self.h.batch_size
        full, tail = divmod(n, bs)
        if self.shard is not None and 0 < tail < self.shard[1]:
            return full                 # a tail batch with fewer rows than ranks is dropped (see __iter__)
        return full + (1 if tail else 0)

    def _load_batch(self, rows):
        items = [self.h.load_patches_from_index_file(r) for r in rows]
        batch = tuple(np.stack([it[k] for it in items]) for k in range(11))
        if self.h.pin_memory:
            import torch
            batch = tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in batch)
        return batch

    def __iter__(self):
        order = self.rng.permutation(len(self.rows)) if self.shuffle else np.arange(len(self.rows))
        bs = self.h.batch_size
        chunks = [[self.rows[i] for i in order[s:s + bs]] for s in range(0, len(order), bs)]
        if self.shard is not None:
            # data parallel: every rank draws the same permutation (same seed) and reads only the rows of its shard,
            # i.e. exactly parallel.shard_batch() of the global batch without loading the other ranks' patches
            from .. import parallel
            # (a ragged tail batch with fewer rows than ranks would leave a rank without work in a step that still
            # needs its all-reduce: it is dropped, at most world-1 rows per epoch)
            chunks = [c[slice(*parallel.shard_bounds(len(c), *self.shard))] for c in chunks if len(c) >= self.shard[1]]
        with cf.ThreadPoolExecutor(max_workers=self.n_parallel) as pool:
            pending = []
            it = iter(chunks)
            for _ in range(max(1, self.h.prefetch)):
                c = next(it, None)
                if c is not None:
                    pending.append(pool.submit(self._load_batch, c))
            while pending:
                fut = pending.pop(0)
                c = next(it, None)
                if c is not None:
                    pending.append(pool.submit(self._load_batch, c))
                yield fut.result()


class PatchHandler3D:
    def __init__(self, data_dir, patch_size, res_increase, batch_size, mask_threshold=0.6, pin_memory=False,
                 prefetch=2):
        self.patch_size = patch_size
        self.res_increase = res_increase
        self.batch_size = batch_size
        self.mask_threshold = mask_threshold
        self.data_directory = data_dir
        self.hr_colnames = ['u', 'v', 'w']
        self.lr_colnames = ['u', 'v', 'w']
        self.venc_colnames = ['venc_u', 'venc_v', 'venc_w']
        self.mag_colnames = ['mag_u', 'mag_v', 'mag_w']
        self.mask_colname = 'mask'
        self.pin_memory = pin_memory
        self.prefetch = prefetch

    def initialize_dataset(self, indexes, shuffle, n_parallel=None, seed=None, shard=None):
        print("Total dataset:", len(indexes), 'shuffle', shuffle)
        return _BatchedDataset(self, indexes, shuffle, n_parallel, seed, shard)

    # -- one CSV row -> one sample (PatchHandler3D.py:49-81) -------------------------------------------
    def load_patches_from_index_file(self, indexes):
        def text(x):
            x = x.numpy() if hasattr(x, "numpy") else x
            return x.decode() if isinstance(x, (bytes, np.bytes_)) else str(x)
        lr_hd5path = f'{self.data_directory}/{text(indexes[0])}'
        hd5path = f'{self.data_directory}/{text(indexes[1])}'
        idx = int(indexes[2])
        x0, y0, z0 = int(indexes[3]), int(indexes[4]), int(indexes[5])
        is_rotate, rotation_plane, rotation_degree_idx = int(indexes[6]), int(indexes[7]), int(indexes[8])
        P, r = self.patch_size, self.res_increase
        H = P * r
        patch_index = np.index_exp[idx, x0:x0 + P, y0:y0 + P, z0:z0 + P]
        hr_patch_index = np.index_exp[idx, x0 * r:x0 * r + H, y0 * r:y0 * r + H, z0 * r:z0 * r + H]
        mask_index = np.index_exp[0, x0 * r:x0 * r + H, y0 * r:y0 * r + H, z0 * r:z0 * r + H]   # mask: row 0 of the HR file
        (u, u_hr, mag_u, v, v_hr, mag_v, w, w_hr, mag_w, venc, mask) = self.load_vectorfield(
            hd5path, lr_hd5path, idx, mask_index, patch_index, hr_patch_index)
        if is_rotate > 0:
            u, v, w = apply_rotation(u, v, w, rotation_degree_idx, rotation_plane, True)
            u_hr, v_hr, w_hr = apply_rotation(u_hr, v_hr, w_hr, rotation_degree_idx, rotation_plane, True)
            mag_u, mag_v, mag_w = apply_rotation(mag_u, mag_v, mag_w, rotation_degree_idx, rotation_plane, False)
            mask = rotate_object(mask, rotation_degree_idx, rotation_plane)
        nx = np.newaxis
        return (u[..., nx], v[..., nx], w[..., nx], mag_u[..., nx], mag_v[..., nx], mag_w[..., nx],
                u_hr[..., nx], v_hr[..., nx], w_hr[..., nx], venc, mask)

    def rotate_object(self, img, rotation_idx, plane_nr):
        return rotate_object(img, rotation_idx, plane_nr)

    def apply_rotation(self, u, v, w, rotation_idx, plane_nr, is_phase_image):
        return apply_rotation(u, v, w, rotation_idx, plane_nr, is_phase_image)

    def load_vectorfield(self, hd5path, lr_hd5path, idx, mask_index, patch_index, hr_patch_index):
        """PatchHandler3D.py:110-160: HR velocities + mask from the HR file, LR velocities / magnitudes / vencs
        from the LR file; velocities / max venc, magnitudes / 4095, mask >= threshold."""
        hires, lowres, mags, vencs = [], [], [], []
        with h5io.open_file(hd5path, 'r') as hl:
            for c in self.hr_colnames:
                hires.append(hl.get(c)[hr_patch_index])
            mask = hl.get(self.mask_colname)[mask_index]
            mask = (mask >= self.mask_threshold) * 1.
        with h5io.open_file(lr_hd5path, 'r') as hl:
            for c, m, ve in zip(self.lr_colnames, self.mag_colnames, self.venc_colnames):
                lowres.append(hl.get(c)[patch_index])
                mags.append(hl.get(m)[patch_index])
                vencs.append(hl.get(ve)[idx])
        global_venc = np.max(vencs)
        hires = self._normalize(np.asarray(hires), global_venc)
        lowres = self._normalize(np.asarray(lowres), global_venc)
        mags = np.asarray(mags) / 4095.
        f32 = 'float32'
        return (lowres[0].astype(f32), hires[0].astype(f32), mags[0].astype(f32),
                lowres[1].astype(f32), hires[1].astype(f32), mags[1].astype(f32),
                lowres[2].astype(f32), hires[2].astype(f32), mags[2].astype(f32),
                global_venc.astype(f32), mask.astype(f32))

    def _normalize(self, u, venc):
        return u / venc
