"""Mirror of the reference's utils/ImageDataset.py (class ImageDataset, :4-85): loads one low-resolution
velocity + magnitude volume of a 4D-flow HDF5 file and normalises it for the network (velocity / max venc,
magnitude / 4095).  File access goes through `h5io.open_file` (h5py when installed, the pure-Python shim
otherwise); everything else is the numpy arithmetic the predictor feeds to `PatchGenerator.patchify`."""
import numpy as np

from . import h5io


class ImageDataset:
    def __init__(self):
        self.velocity_colnames = ['u', 'v', 'w']
        self.venc_colnames = ['venc_u', 'venc_v', 'venc_w']
        self.mag_colnames = ['mag_u', 'mag_v', 'mag_w']
        self.dx_colname = 'dx'

    def _normalize(self, velocity, venc):
        return velocity / venc

    def _set_images(self, velocity_images, mag_images, venc, dx):
        velocity_images = self._normalize(velocity_images, venc)
        mag_images = mag_images / 4095.          # magnitude 0 .. 1
        self.u, self.v, self.w = (velocity_images[i].astype('float32') for i in range(3))
        self.mag_u, self.mag_v, self.mag_w = (mag_images[i].astype('float32') for i in range(3))
        self.venc = venc.astype('float32')       # kept to de-normalise the prediction
        self.velocity_per_px = self.venc / 2048  # one phase-image quantum: smaller predictions are zeroed
        self.dx = dx

    def postprocess_result(self, results, zerofy=True):
        results = results * self.venc
        if zerofy:
            print(f"Zero out velocity component less than {self.velocity_per_px}")
            results[np.abs(results) < self.velocity_per_px] = 0
        return results

    def get_dataset_len(self, filepath):
        with h5io.open_file(filepath, 'r') as hl:
            return hl[self.velocity_colnames[0]].shape[0]

    def load_vectorfield(self, filepath, idx):
        lowres, mags, vencs = [], [], []
        dx = None
        with h5io.open_file(filepath, 'r') as hl:
            if self.dx_colname in hl:
                dx = hl.get(self.dx_colname)[idx]
            for vel, mag, venc in zip(self.velocity_colnames, self.mag_colnames, self.venc_colnames):
                lowres.append(np.asarray(hl.get(vel)[idx]))
                mags.append(np.asarray(hl.get(mag)[idx]))
                vencs.append(np.asarray(hl.get(venc)[idx]))
        self._set_images(np.asarray(lowres), np.asarray(mags), np.max(vencs), dx)
