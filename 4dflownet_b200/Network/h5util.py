"""Mirror of the reference's Network/h5util.py (:5-23): append-along-axis-0 dataset store used by
TrainerController.quicksave.  Same semantics as utils/prediction_utils.save_to_h5; file access through
h5io.open_file (h5py when installed, the pure-Python HDF5 shim otherwise)."""
import os

from ..utils.prediction_utils import save_to_h5


def save_predictions(output_path, output_filename, col_name, dataset, compression=None):
    os.makedirs(output_path, exist_ok=True)
    save_to_h5(os.path.join(output_path, output_filename), col_name, dataset, compression=compression)
