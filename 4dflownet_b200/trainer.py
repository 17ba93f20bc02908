"""Entry point mirroring the reference's src/trainer.py (:5-73): load the patch index CSVs, build the three
PatchHandler3D iterators, construct TrainerController with the reference's positional arguments and train.
Under `torchrun` every rank runs this script; each rank reads a disjoint slice of every global batch
(`parallel.shard_batch`) and the controller all-reduces the flat gradient buffer once per step."""
import numpy as np

from . import parallel
from .Network.PatchHandler3D import PatchHandler3D
from .Network.TrainerController import TrainerController


def load_indexes(index_file):
    """Patch index file (csv): source,target,index,start_x,start_y,start_z,rotate,rotation_plane,
    rotation_degree_idx,coverage (trainer.py:5-10)."""
    return np.genfromtxt(index_file, delimiter=',', skip_header=True, dtype='unicode')


class _Sharded:
    """Every rank iterates the same global batches (same shuffle seed) and keeps its contiguous slice."""

    def __init__(self, dataset):
        self.ds = dataset

    def __len__(self):
        return len(self.ds)

    def __iter__(self):
        for batch in self.ds:
            yield parallel.shard_batch(batch)


def main(data_dir='../data', training_file=None, validate_file=None, benchmark_file=None, QUICKSAVE=True,
         restore=False, model_dir="../models/4DFlowNet", model_file="4DFlowNet-best.h5",
         initial_learning_rate=2e-4, epochs=60, batch_size=20, mask_threshold=0.6, network_name='4DFlowNet',
         patch_size=16, res_increase=2, low_resblock=8, hi_resblock=4, models_root="../models", seed=0):
    training_file = training_file or f'{data_dir}/train.csv'
    validate_file = validate_file or f'{data_dir}/validate.csv'
    benchmark_file = benchmark_file if benchmark_file is not None else f'{data_dir}/benchmark.csv'
    world = parallel.world_size()
    if batch_size % world:
        raise ValueError(f"batch_size {batch_size} must be a multiple of the number of ranks {world}")

    trainset = load_indexes(training_file)
    valset = load_indexes(validate_file)
    z = PatchHandler3D(data_dir, patch_size, res_increase, batch_size, mask_threshold, pin_memory=True)
    trainset = _Sharded(z.initialize_dataset(trainset, shuffle=True, n_parallel=None, seed=seed))
    valdh = PatchHandler3D(data_dir, patch_size, res_increase, batch_size, mask_threshold, pin_memory=True)
    valset = _Sharded(valdh.initialize_dataset(valset, shuffle=True, n_parallel=None, seed=seed + 1))
    testset = None
    if QUICKSAVE and benchmark_file:
        ph = PatchHandler3D(data_dir, patch_size, res_increase, batch_size, mask_threshold)
        testset = ph.initialize_dataset(load_indexes(benchmark_file), shuffle=False)   # first batch saved per best epoch

    print(f"4DFlowNet Patch {patch_size}, lr {initial_learning_rate}, batch {batch_size}")
    network = TrainerController(patch_size, res_increase, initial_learning_rate, QUICKSAVE, network_name,
                                low_resblock, hi_resblock, max_batch=max(1, batch_size // world))
    network.init_model_dir(models_root)
    if restore:
        print(f"Restoring model {model_file}...")
        network.restore_model(model_dir, model_file)
        print("Learning rate", network.optimizer.lr.numpy())
    network.train_network(trainset, valset, n_epoch=epochs, testset=testset)
    return network


if __name__ == "__main__":
    main()
