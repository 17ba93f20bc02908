"""Append-along-axis-0 dataset store with the reference's interface (Network/h5util.py:5-23,
utils/prediction_utils.py:15-28).  The reference writes resizable gzip HDF5 datasets through
h5py; h5py / libhdf5 are not part of this image, so when h5py cannot be imported the same
append semantics are kept in an .npz side file `<name>.npz` (documented in DESIGN.md as the
storage-format gap; the numerical content is identical)."""
import os

import numpy as np


def _append_npz(path, col_name, dataset):
    path = path + ".npz"
    store = {}
    if os.path.exists(path):
        with np.load(path) as z:
            store = {k: z[k] for k in z.files}
    dataset = np.asarray(dataset)
    if col_name in store:
        store[col_name] = np.concatenate([store[col_name], dataset], axis=0)
    else:
        store[col_name] = dataset
    np.savez_compressed(path[:-4], **store)


def save_predictions(output_dir, output_filename, col_name, dataset, compression=None):
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, output_filename)
    try:
        import h5py
    except ImportError:
        _append_npz(path, col_name, dataset)
        return
    dataset = np.asarray(dataset)
    with h5py.File(path, "a") as hf:
        if col_name not in hf:
            hf.create_dataset(col_name, data=dataset, maxshape=(None,) + dataset.shape[1:], compression=compression)
        else:
            hf[col_name].resize(hf[col_name].shape[0] + dataset.shape[0], axis=0)
            hf[col_name][-dataset.shape[0]:] = dataset
