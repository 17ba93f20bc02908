"""Entry point mirroring the reference's src/predictor.py: `prepare_network` keeps the
signature (patch_size, res_increase, low_resblock, hi_resblock) (predictor.py:11) and the
`__main__` flow (predictor.py:31-116): load volume -> patchify -> batched predict -> stitch
-> denormalise -> zero small values -> save."""
import os
import time

import numpy as np

from .Network.PatchGenerator import PatchGenerator
from .Network.SR4DFlowNet import SR4DFlowModel


def prepare_network(patch_size, res_increase, low_resblock, hi_resblock, max_batch=8, device=None):
    return SR4DFlowModel(patch_size, res_increase, low_resblock, hi_resblock, max_batch=max_batch, training=False,
                         device=device)


_PINNED = {}


def _pinned_like(t):
    """Reusable page-locked host buffer for the stitched volume (a pageable D2H copy of ~150 MB costs tens of ms)."""
    import torch
    key = (tuple(t.shape), t.dtype)
    if key not in _PINNED:
        _PINNED.clear()
        _PINNED[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    return _PINNED[key]


def predict_volume(network, pgen, dataset, batch_size=8, round_small_values=True, gpu_stitch=True, all_ranks=False,
                   reuse_host_buffer=False):
    """Body of the reference's per-row loop (predictor.py:74-107).  Returns (3,X,Y,Z) fp32.
    Under torch.distributed the patch list is cut into contiguous per-rank chunks (no collective on the compute
    path): every rank tiles and predicts only its own patches, the predictions are sent to rank 0, which stitches
    and returns the volume (the other ranks return None unless all_ranks=True).  reuse_host_buffer=True returns
    a view of an internal page-locked buffer that the next call overwrites (saves one ~150 MB host copy)."""
    import torch
    from . import parallel
    eng = network.engine
    H = eng.H
    n = pgen.count_patches(dataset.u.shape)
    lo, hi = parallel.shard_bounds(n)
    # tiling on the device: one upload of the six volumes, the window copies happen there (bit-identical to patchify)
    stacks = pgen.patchify_device(dataset, eng.device, lo, hi) if parallel.world_size() > 1 else \
        pgen.patchify_device(dataset, eng.device)
    local = torch.empty((hi - lo, H, H, H, 3), device=eng.device, dtype=torch.float32)
    for i in range(0, hi - lo, batch_size):
        sl = slice(i, min(i + batch_size, hi - lo))
        eng.forward([t[sl] for t in stacks], out=local[sl])
    results = parallel.gather_rows(local, n, dst=None if all_ranks else 0)
    if results is None:
        return None
    venc = float(dataset.venc)
    if gpu_stitch:
        side_hr = (pgen.patch_size - pgen.effective_patch_size) // 2 * pgen.res_increase
        vol = eng.stitch(results, (pgen.nr_x, pgen.nr_y, pgen.nr_z), pgen.stitched_shape(), side_hr, venc,
                         round_small_values)
        host = _pinned_like(vol)
        host.copy_(vol, non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        return host.numpy() if reuse_host_buffer else host.numpy().copy()
    res = results.cpu().numpy()
    out = []
    for c in range(3):
        v = pgen._patchup_with_overlap(res[..., c], pgen.nr_x, pgen.nr_y, pgen.nr_z) * np.float32(venc)
        if round_small_values:
            v[np.abs(v) < dataset.velocity_per_px] = 0
        out.append(v)
    return np.stack(out)


def main(data_dir="../data", filename="example_data.h5", output_dir="../result", output_filename="example_result.h5",
         model_path="../models/4DFlowNet/4DFlowNet.h5", patch_size=24, res_increase=2, batch_size=8,
         round_small_values=True, low_resblock=8, hi_resblock=4):
    from .utils.ImageDataset import ImageDataset
    from .utils import prediction_utils
    from . import parallel
    local_device = parallel.init_from_env()      # under torchrun: one rank per GPU, patch list sharded per rank
    input_filepath = f"{data_dir}/{filename}"
    pgen = PatchGenerator(patch_size, res_increase)
    dataset = ImageDataset()
    nr_rows = dataset.get_dataset_len(input_filepath)
    print(f"Number of rows in dataset: {nr_rows}")
    print(f"Loading 4DFlowNet: {res_increase}x upsample")
    network = prepare_network(patch_size, res_increase, low_resblock, hi_resblock, max_batch=batch_size,
                              device=local_device)
    network.load_weights(model_path)
    if parallel.is_main():
        os.makedirs(output_dir, exist_ok=True)
    for nrow in range(nr_rows):
        print(f"\nProcessed ({nrow + 1}/{nr_rows}) - {time.ctime()}")
        dataset.load_vectorfield(input_filepath, nrow)
        t0 = time.time()
        vol = predict_volume(network, pgen, dataset, batch_size, round_small_values)
        if vol is None:
            continue            # not rank 0 of a sharded prediction
        print(f"Predicted {vol.shape[1:]} in {time.time() - t0:.2f} secs.")
        for i in range(3):
            prediction_utils.save_to_h5(f"{output_dir}/{output_filename}", dataset.velocity_colnames[i],
                                        vol[i][None], compression="gzip")
        if dataset.dx is not None:
            prediction_utils.save_to_h5(f"{output_dir}/{output_filename}", dataset.dx_colname,
                                        np.expand_dims(dataset.dx / res_increase, 0), compression="gzip")
    print("Done!")


if __name__ == "__main__":
    main()
