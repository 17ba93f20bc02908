"""Mirror of the reference's utils/prediction_utils.py (:5-28): append-along-axis-0 HDF5 result columns."""
import os

import numpy as np

from . import h5io


def save_to_h5(output_filepath, col_name, dataset, compression=None):
    dataset = np.asarray(dataset)
    if dataset.dtype == np.float64:
        dataset = dataset.astype(np.float32)     # the reference stores float32 to save space
    with h5io.open_file(output_filepath, 'a') as hf:
        if col_name not in hf:
            maxshape = (None,) + tuple(dataset.shape[1:])
            hf.create_dataset(col_name, data=dataset, maxshape=maxshape, compression=compression)
        else:
            hf[col_name].resize(hf[col_name].shape[0] + dataset.shape[0], axis=0)
            hf[col_name][-dataset.shape[0]:] = dataset


def save_predictions(output_dir, output_filename, colnames, predictions, compression=None):
    os.makedirs(output_dir, exist_ok=True)
    output_filepath = os.path.join(output_dir, output_filename)
    for i, col in enumerate(colnames):
        save_to_h5(output_filepath, col, predictions[:, :, :, :, i], compression=compression)
    print(f"Prediction saved to {output_filepath}")
