"""Training entry point with the reference's defaults (src/trainer.py:12-73) exposed as arguments of `main`.

Flow: read the three patch-index CSVs, wrap them in PatchHandler3D iterators (train / validation shuffled, the
benchmark set in file order so the first batch can be re-predicted after every best epoch), build a
TrainerController with the reference's positional arguments, optionally restore, train.  Under `torchrun` every
rank runs this function: all ranks draw the same global batches (same shuffle seed), each loads only the rows of its
contiguous shard (the `shard=` argument of `initialize_dataset`, equal to `parallel.shard_batch` of the global batch)
and the controller all-reduces the flat gradient buffer once per step.  `main` joins the process group itself
(`parallel.init_from_env`), binds the rank to its LOCAL_RANK device, broadcasts rank 0's initial (or restored) weights
and optimizer state, and lets only rank 0 write the model directory."""
import numpy as np

from . import parallel
from .Network.PatchHandler3D import PatchHandler3D
from .Network.TrainerController import TrainerController

CSV_COLUMNS = ("source", "target", "index", "start_x", "start_y", "start_z", "rotate", "rotation_plane",
               "rotation_degree_idx", "coverage")


def load_indexes(index_file):
    """Rows of a patch index CSV (header skipped) as a 2-D array of strings, one column per CSV_COLUMNS entry
    (reference: trainer.py:5-10)."""
    return np.genfromtxt(index_file, dtype="unicode", delimiter=",", skip_header=True)


def _iterator(csv_path, geometry, shuffle, seed=None, pinned=False, sharded=False):
    data_dir, patch_size, res_increase, batch_size, mask_threshold = geometry
    handler = PatchHandler3D(data_dir, patch_size, res_increase, batch_size, mask_threshold, pin_memory=pinned)
    kwargs = {} if seed is None else {"n_parallel": None, "seed": seed}
    if sharded and parallel.world_size() > 1:
        kwargs["shard"] = (parallel.rank(), parallel.world_size())
    return handler.initialize_dataset(load_indexes(csv_path), shuffle=shuffle, **kwargs)


def main(data_dir='../data', training_file=None, validate_file=None, benchmark_file=None, QUICKSAVE=True,
         restore=False, model_dir="../models/4DFlowNet", model_file="4DFlowNet-best.h5",
         initial_learning_rate=2e-4, epochs=60, batch_size=20, mask_threshold=0.6, network_name='4DFlowNet',
         patch_size=16, res_increase=2, low_resblock=8, hi_resblock=4, models_root="../models", seed=0):
    local_device = parallel.init_from_env()
    ranks = parallel.world_size()
    if batch_size % ranks:
        raise ValueError(f"batch_size {batch_size} must be a multiple of the number of ranks {ranks}")
    paths = {"train": training_file or f"{data_dir}/train.csv",
             "validate": validate_file or f"{data_dir}/validate.csv",
             "benchmark": f"{data_dir}/benchmark.csv" if benchmark_file is None else benchmark_file}
    geometry = (data_dir, patch_size, res_increase, batch_size, mask_threshold)

    train_batches = _iterator(paths["train"], geometry, shuffle=True, seed=seed, pinned=True, sharded=True)
    val_batches = _iterator(paths["validate"], geometry, shuffle=True, seed=seed + 1, pinned=True, sharded=True)
    bench_batches = None
    if QUICKSAVE and paths["benchmark"]:
        bench_batches = _iterator(paths["benchmark"], geometry, shuffle=False)

    print(f"4DFlowNet Patch {patch_size}, lr {initial_learning_rate}, batch {batch_size}")
    controller = TrainerController(patch_size, res_increase, initial_learning_rate, QUICKSAVE, network_name,
                                   low_resblock, hi_resblock, max_batch=max(1, batch_size // ranks),
                                   device=local_device)
    controller.init_model_dir(models_root)
    if restore:
        print(f"Restoring model {model_file}...")
        controller.restore_model(model_dir, model_file)
        print("Learning rate", controller.optimizer.lr.numpy())
    controller.sync_ranks()          # rank 0's weights / Adam state everywhere (no-op in a single process)
    controller.train_network(train_batches, val_batches, n_epoch=epochs, testset=bench_batches)
    return controller


if __name__ == "__main__":
    main()
