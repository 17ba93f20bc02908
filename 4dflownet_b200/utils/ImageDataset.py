"""One low-resolution 4D-flow volume, normalised the way the network expects it.

Drop-in for the reference's utils/ImageDataset.py (class ImageDataset, :4-85) as used by predictor.py:50-76 and
PatchGenerator.patchify: `get_dataset_len(path)`, `load_vectorfield(path, row)` and afterwards the attributes
u, v, w, mag_u, mag_v, mag_w (float32 volumes), venc, velocity_per_px, dx, plus the column-name lists.
Velocities are divided by the largest of the three vencs of the row, magnitudes by 4095 (12-bit images); files
are opened through `h5io.open_file` (h5py when installed, the pure-Python HDF5 shim otherwise)."""
import numpy as np

from . import h5io

_AXES = ("u", "v", "w")
_MAG_FULL_SCALE = 4095.0        # magnitude images are 12-bit
_PHASE_LEVELS = 2048            # venc / 2048 = one quantisation step of the phase image


class ImageDataset:
    velocity_colnames = list(_AXES)
    venc_colnames = [f"venc_{a}" for a in _AXES]
    mag_colnames = [f"mag_{a}" for a in _AXES]
    dx_colname = "dx"

    def get_dataset_len(self, filepath):
        """Number of rows (time frames) in the file."""
        with h5io.open_file(filepath, "r") as f:
            return f[self.velocity_colnames[0]].shape[0]

    def load_vectorfield(self, filepath, idx):
        """Read row `idx` and set the normalised attributes."""
        with h5io.open_file(filepath, "r") as f:
            spacing = f[self.dx_colname][idx] if self.dx_colname in f else None
            read = lambda names: np.stack([np.asarray(f[n][idx]) for n in names])    # noqa: E731
            velocity, magnitude = read(self.velocity_colnames), read(self.mag_colnames)
            venc = np.max([np.asarray(f[n][idx]) for n in self.venc_colnames])
        self._assign(velocity / venc, magnitude / _MAG_FULL_SCALE, venc, spacing)

    def _assign(self, velocity, magnitude, venc, spacing):
        for axis, vel, mag in zip(_AXES, velocity, magnitude):
            setattr(self, axis, vel.astype(np.float32))
            setattr(self, "mag_" + axis, mag.astype(np.float32))
        self.venc = np.float32(venc)                      # needed again to de-normalise the prediction
        self.velocity_per_px = self.venc / _PHASE_LEVELS  # predictions below one phase step are zeroed
        self.dx = spacing

    def postprocess_result(self, results, zerofy=True):
        """De-normalise a prediction (x venc) and optionally zero what is below one phase quantisation step."""
        out = results * self.venc
        if zerofy:
            print(f"Zero out velocity component less than {self.velocity_per_px}")
            out[np.abs(out) < self.velocity_per_px] = 0
        return out
