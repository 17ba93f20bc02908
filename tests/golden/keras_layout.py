"""Builds an HDF5 file with the layout `tf.keras.Model.save(path)` of TensorFlow 2.2 / Keras 2.3.0-tf writes for the
reference's functional model (the file predictor.py:61 / TrainerController.py:394 hand to `load_weights`, written at
TrainerController.py:80,356), using the repo's HDF5 writer -- TensorFlow and h5py are not installable here.

What is mirrored (keras/engine/saving/hdf5_format.py of TF 2.2, `save_model_to_hdf5` + `save_weights_to_hdf5_group`):
  /                     attrs: keras_version = b'2.3.0-tf', backend = b'tensorflow', model_config = JSON bytes
                        (no training_config / optimizer_weights: the reference never compiles the model)
  /model_weights        attrs: layer_names = S-array of EVERY layer of model.layers (input layers, the TensorFlowOpLayer
                        wrappers of the raw tf ops, concatenate, conv3d*, leaky_re_lu*, add*), backend, keras_version
  /model_weights/<l>    one group per layer; attrs: weight_names = S-array ([] for weightless layers)
  /model_weights/<l>/<l>/kernel:0, bias:0     float32 datasets, kernel (kx,ky,kz,Cin,Cout)
`name_offset` reproduces a writer process whose Conv3D uid counter was not at zero (names conv3d_<k+offset>).
The order of `layer_names` is Keras' graph-depth order in a real file; readers that match by name (ours, and Keras with
by_name=True) do not depend on it, so this builder lists weightless layers first, then the conv layers in creation order.

What stays unverifiable offline: the exact bytes libhdf5 emits for such a file (object-header / B-tree versions, the
chunked encoding of an over-long layer_names attribute) -- the reader itself is exercised on libhdf5-written files
(the reference's data/example_data*.h5, tests/test_reference_fixtures.py) -- and the published pretrained weights
(README.md:23-25), which cannot be downloaded here."""
import json

import numpy as np


def conv_layer_names(n_conv, name_offset=0):
    return ["conv3d" if k + name_offset == 0 else f"conv3d_{k + name_offset}" for k in range(n_conv)]


def write_tf22_model_file(h5io, path, weights, name_offset=0):
    """weights: {'conv3d_k/kernel'|'conv3d_k/bias': array} in this package's naming (creation order)."""
    layers = []
    for n in weights:
        ly = n.split("/")[0]
        if ly not in layers:
            layers.append(ly)
    on_disk = dict(zip(layers, conv_layer_names(len(layers), name_offset)))
    weightless = ["u", "v", "w", "u_mag", "v_mag", "w_mag", "tf_op_layer_Pow", "tf_op_layer_AddV2", "concatenate",
                  "concatenate_1", "tf_op_layer_MirrorPad", "leaky_re_lu", "add", "tf_op_layer_ResizeBilinear",
                  "concatenate_3"]
    cfg = {"class_name": "Model", "config": {"name": "model", "layers": [{"name": n} for n in weightless + list(on_disk.values())]},
           "keras_version": "2.3.0-tf", "backend": "tensorflow"}
    with h5io.File(path, "w") as f:
        f.attrs["keras_version"] = np.bytes_(b"2.3.0-tf")
        f.attrs["backend"] = np.bytes_(b"tensorflow")
        f.attrs["model_config"] = np.bytes_(json.dumps(cfg).encode("utf8"))
        mw = f.create_group("model_weights")
        mw.attrs["layer_names"] = np.asarray([n.encode("utf8") for n in weightless + list(on_disk.values())])
        mw.attrs["backend"] = np.bytes_(b"tensorflow")
        mw.attrs["keras_version"] = np.bytes_(b"2.3.0-tf")
        for n in weightless:
            g = mw.create_group(n)
            g.attrs["weight_names"] = np.asarray([], dtype="S1")
        for ly, disk in on_disk.items():
            g = mw.create_group(disk)
            leaves = [n.split("/")[1] for n in weights if n.split("/")[0] == ly]
            g.attrs["weight_names"] = np.asarray([f"{disk}/{leaf}:0".encode("utf8") for leaf in leaves])
            gg = g.create_group(disk)
            for leaf in leaves:
                gg.create_dataset(f"{leaf}:0", data=np.asarray(weights[f"{ly}/{leaf}"], dtype=np.float32))
    return on_disk
