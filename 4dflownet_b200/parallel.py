"""Host-side data-parallel plumbing: one process per GPU, `torch.distributed`.

The path shards on the patch / batch axis and nothing else:
  * training  -- rank k takes a contiguous shard of the global batch, leaves the SUM over its samples of the
    per-sample loss gradients in the engine's flat gradient buffer, ONE all-reduce(SUM) of that buffer per step,
    then every rank applies the identical fused Adam step with the L2 term scaled by the GLOBAL batch
    (TrainerController.py:223,249: tape.gradient of the (B,) loss vector sums over samples and the scalar l2 is
    added to each entry);
  * inference -- the patch list of a volume is cut into contiguous chunks per rank (predictor.py:82-94 is
    embarrassingly parallel), results are gathered for the unchanged stitcher.
Everything here works on CPU tensors with the gloo backend too (tests/test_parallel_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Join the process group `torchrun` describes (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*), once; returns the
    local device index this rank must use.  Without WORLD_SIZE > 1 in the environment nothing is initialised and the
    current CUDA device (or 0) is returned, so the single-process entry points behave as before.  backend: "nccl"
    when CUDA is available, else "gloo" (CPU tests)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local)
    if world > 1 and not initialized():
        # SR4D_DIST_BACKEND=gloo: several ranks on ONE GPU (NCCL refuses that), used by the single-GPU torchrun smoke test
        backend = backend or os.environ.get("SR4D_DIST_BACKEND") or ("nccl" if cuda else "gloo")
        kwargs = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kwargs)
    return local


def is_main():
    """Rank 0 owns every file the run writes (model directory, loss.csv, checkpoints, quicksave, results)."""
    return rank() == 0


def barrier():
    if world_size() > 1:
        dist.barrier()


def broadcast_(tensor, src=0):
    """In-place broadcast from `src` (initial weights / restored checkpoints must be identical on every rank)."""
    if world_size() > 1:
        dist.broadcast(tensor, src)
    return tensor


def initialized():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if initialized() else 1


def rank():
    return dist.get_rank() if initialized() else 0


def shard_bounds(n, rank_=None, world=None):
    """[lo, hi) of the contiguous chunk of `n` items owned by `rank_`; sizes differ by at most one and the
    larger chunks come first."""
    r = rank() if rank_ is None else int(rank_)
    w = world_size() if world is None else int(world)
    base, extra = divmod(int(n), w)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def shard_batch(data_pairs, rank_=None, world=None):
    """This rank's slice of a global-batch 11-tuple (PatchHandler3D.py:78-81), along axis 0."""
    n = len(data_pairs[0])
    lo, hi = shard_bounds(n, rank_, world)
    return tuple(a[lo:hi] for a in data_pairs)


def allreduce_gradients(flat_grads):
    """The one collective of a training step: SUM of the flat gradient buffer over ranks, in place."""
    if world_size() > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


class MetricTail:
    """The caller-owned floats behind the flat gradients (include/sr4d.h SR4D_METRIC_TAIL), laid out so that the
    step's ONE all-reduce(SUM) of `engine.grads_full` also gathers the metrics (SURVEY 8e "+ metric tail"):

        slot r = [count_r, l2_r, per_sample rows of rank r (max_batch x 4)]   for r < world
        last float = sum over ranks of the local batch sizes (the global batch Adam's L2 term needs,
                     TrainerController.py:249; read on the device by sr4d_adam_step_counted)

    Every rank zeroes the tail and fills only its own slot, so the SUM is a gather.  Works on CPU tensors (gloo
    tests) and on the engine's device view alike."""

    def __init__(self, tail, max_batch, rank_=None, world=None):
        self.tail = tail
        self.rank = rank() if rank_ is None else int(rank_)
        self.world = world_size() if world is None else int(world)
        self.max_batch = int(max_batch)
        self.slot = 2 + 4 * self.max_batch
        if self.world * self.slot + 1 > tail.numel():
            raise ValueError(f"metric tail of {tail.numel()} floats cannot hold {self.world} ranks x {self.max_batch} samples")
        self.count_index = tail.numel() - 1
        self._template = torch.zeros_like(tail)
        self._template_batch = None
        self._host = torch.empty(tail.numel(), dtype=tail.dtype)
        if tail.is_cuda:
            self._host = self._host.pin_memory()

    def begin(self, local_batch):
        """Reset the tail for a step of `local_batch` samples on this rank; returns the (B,4) per-sample view and the
        (1,) l2 view the engine writes its metrics into."""
        B = int(local_batch)
        if B > self.max_batch:
            raise ValueError(f"local batch {B} exceeds the tail slot ({self.max_batch})")
        if B != self._template_batch:
            t = torch.zeros(self.tail.numel(), dtype=self.tail.dtype)
            t[self.rank * self.slot] = B
            t[self.count_index] = B
            self._template.copy_(t)
            self._template_batch = B
        self.tail.copy_(self._template)
        base = self.rank * self.slot
        return self.tail[base + 2:base + 2 + 4 * B].view(B, 4), self.tail[base + 1:base + 2]

    def exchange(self):
        """All-reduce of the tail alone (validation steps, which move no gradients)."""
        if self.world > 1:
            dist.all_reduce(self.tail, op=dist.ReduceOp.SUM)

    def read_begin(self):
        """Enqueue the device->host copy of the all-reduced tail (into one of two pinned buffers) and return a token
        for read_end -- lets the caller enqueue the NEXT step before waiting for this one's metrics."""
        if getattr(self, "_ring", None) is None:
            self._ring = [self._host, torch.empty_like(self._host).pin_memory() if self.tail.is_cuda
                          else torch.empty_like(self._host)]
            self._events = [torch.cuda.Event() if self.tail.is_cuda else None for _ in range(2)]
            self._turn = 0
        k = self._turn
        self._turn ^= 1
        host = self._ring[k]
        n_used = self.world * self.slot
        host[:n_used].copy_(self.tail[:n_used], non_blocking=True)
        host[self.count_index:].copy_(self.tail[self.count_index:], non_blocking=True)
        if self.tail.is_cuda:
            self._events[k].record(torch.cuda.current_stream(self.tail.device))
        return k

    def read_end(self, token):
        """((B_global,4) per-sample metrics in rank order == global batch order, l2, B_global) of the step whose copy
        read_begin enqueued; waits for that copy only (an event, not the stream)."""
        if self.tail.is_cuda:
            self._events[token].synchronize()
        h = self._ring[token].numpy()
        rows = []
        for r in range(self.world):
            base = r * self.slot
            n = int(round(float(h[base])))
            rows.append(h[base + 2:base + 2 + 4 * n].reshape(n, 4))
        per = np.concatenate(rows, axis=0).copy()
        return per, float(h[self.rank * self.slot + 1]), int(round(float(h[self.count_index])))

    def read(self):
        """After the all-reduce: the metrics of this step, now (one small device->host copy + one event wait -- the
        step's only host round trip)."""
        return self.read_end(self.read_begin())


def l2_grad_scale(local_batch, l2_coeff=5e-7):
    """d/dw of the regulariser as the reference differentiates it: 2*l2*w added once per sample of the
    GLOBAL batch (every rank must apply the same scale after the all-reduce)."""
    return float(global_count(local_batch)) * 2.0 * l2_coeff


def global_count(local_count):
    """Sum of an integer over ranks (global batch size when shards are ragged)."""
    if world_size() == 1:
        return int(local_count)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def gather_rows(local, n_total, dst=None):
    """Gather of row-sharded tensors whose shard sizes follow `shard_bounds(n_total)`.  dst=None: all-gather, every
    rank returns the (n_total, ...) tensor; dst=k: only rank k receives it (the others return None)."""
    w = world_size()
    if w == 1:
        return local
    if dst is not None:
        me = rank()
        out = torch.empty((int(n_total),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) if me == dst else None
        ops = []
        if me == dst:
            for r in range(w):
                lo, hi = shard_bounds(n_total, r, w)
                if r == me:
                    out[lo:hi] = local
                elif hi > lo:
                    ops.append(dist.P2POp(dist.irecv, out[lo:hi], r))
        elif local.shape[0] > 0:
            ops.append(dist.P2POp(dist.isend, local.contiguous(), dst))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    chunk = -(-int(n_total) // w)
    pad = torch.zeros((chunk,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((w * chunk,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for r in range(w):
        lo, hi = shard_bounds(n_total, r, w)
        parts.append(out[r * chunk:r * chunk + (hi - lo)])
    return torch.cat(parts, dim=0)


def gather_metrics(per_sample):
    """(B_local,4) per-sample metrics -> (B_global,4) on every rank: the reference's running means
    (TrainerController.py:52-63,241-257) average over every sample of the global batch."""
    w = world_size()
    if w == 1:
        return per_sample
    n = global_count(per_sample.shape[0])
    counts = [shard_bounds(n, r, w) for r in range(w)]
    if all(hi - lo == per_sample.shape[0] for lo, hi in counts[:1]) and n == w * per_sample.shape[0]:
        out = torch.empty((n,) + tuple(per_sample.shape[1:]), dtype=per_sample.dtype, device=per_sample.device)
        dist.all_gather_into_tensor(out, per_sample.contiguous())
        return out
    return gather_rows(per_sample, n)
