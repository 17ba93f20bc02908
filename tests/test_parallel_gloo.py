"""world_size-2 gloo tests (CPU) of the data-parallel host logic: shard arithmetic, the single gradient
all-reduce reproducing the reference's summed-gradient + global-batch L2 semantics, row gathers.
The oracle stands in for the engine's gradient computation here (tests may use it; the product never does)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(fn, world=2):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn), nprocs=world, join=True)


def _entry(rank, world, port, fn):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn](rank, world)
    finally:
        dist.destroy_process_group()


def _par():
    return importlib.import_module("4dflownet_b200.parallel")


def test_shard_bounds_cover_exactly():
    par = _par()
    for n in (0, 1, 7, 12, 256, 1176):
        for w in (1, 2, 3, 8):
            got = [par.shard_bounds(n, r, w) for r in range(w)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(got[i][1] == got[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert par.shard_bounds(256, 3, 8) == (96, 128)      # config 5: 256 patches over 8 ranks


def _dp_gradient_identity(rank, world):
    par = _par()
    oracle = importlib.import_module("oracle.sr4d_oracle")
    P, r, low, hi, Bg = 6, 2, 1, 1, 4
    params = {k: v.astype(np.float64) for k, v in oracle.glorot_params(low, hi, seed=3, bias_scale=0.05).items()}
    batch = oracle.synthetic_batch(Bg, P, r, seed=9)
    names = [n for n, _ in oracle.param_table(low, hi)]
    # reference semantics on the full batch: d/dw [ sum_b loss_b + Bg * l2 ]
    g_full, met_full = oracle.gradients(params, batch, r, low, hi)
    # this rank: summed gradient of its shard WITHOUT the regulariser (what sr4d_train_fwd_bwd leaves in grads)
    shard = par.shard_batch(batch)
    Bl = len(shard[0])
    assert Bl == Bg // world
    g_loc, met_loc = oracle.gradients(params, shard, r, low, hi)
    flat = torch.cat([torch.from_numpy(np.ascontiguousarray(
        g_loc[n] - (Bl * 2 * oracle.L2_COEFF * params[n] if n.endswith("kernel") else 0.0))).reshape(-1) for n in names])
    par.allreduce_gradients(flat)                                   # the one collective
    scale = par.l2_grad_scale(Bl)
    assert abs(scale - Bg * 2 * oracle.L2_COEFF) < 1e-18
    off = 0
    for n in names:
        cnt = params[n].size
        g = flat[off:off + cnt].numpy().reshape(params[n].shape)
        off += cnt
        if n.endswith("kernel"):
            g = g + scale * params[n]
        np.testing.assert_allclose(g, g_full[n], rtol=1e-9, atol=1e-14, err_msg=n)
    # metrics: gathering the per-sample rows reproduces the full-batch vector
    per = torch.from_numpy(np.stack([met_loc["loss"], met_loc["mse"], met_loc["rel_err"], np.zeros(Bl)], 1))
    allm = par.gather_metrics(per)
    np.testing.assert_allclose(allm[:, 0].numpy() , met_full["loss"], rtol=1e-12)
    np.testing.assert_allclose(allm[:, 2].numpy(), met_full["rel_err"], rtol=1e-12)


def _ragged_gather(rank, world):
    par = _par()
    n = 7                                                           # ragged: 4 + 3
    lo, hi = par.shard_bounds(n)
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
    out = par.gather_rows(full[lo:hi].clone(), n)
    assert torch.equal(out, full)
    assert par.global_count(hi - lo) == n
    per = par.gather_metrics(full[lo:hi].clone())
    assert torch.equal(per, full)
    only0 = par.gather_rows(full[lo:hi].clone(), n, dst=0)          # point-to-point gather to one rank
    assert (only0 is None) == (rank != 0)
    if rank == 0:
        assert torch.equal(only0, full)


def _metric_tail_one_collective(rank, world):
    """The layout TrainerController.train_step uses: [flat gradients | metric tail] all-reduced ONCE; the SUM gathers
    every rank's per-sample rows, its l2 value and the global sample count (ragged shards: 3 + 2 samples)."""
    par = _par()
    flat_n, tail_n, max_b, n = 40, 4096, 3, 5
    lo, hi = par.shard_bounds(n)
    full = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4) + 0.25
    buf = torch.full((flat_n + tail_n,), float("nan"))
    buf[:flat_n] = float(rank + 1)                                   # "gradients" of this rank
    tail = par.MetricTail(buf[flat_n:], max_b)
    for rep in range(2):                                             # second pass: the tail is reset, not accumulated
        buf[:flat_n] = float(rank + 1)
        per, l2 = tail.begin(hi - lo)
        per.copy_(full[lo:hi])
        l2.fill_(0.5)
        par.allreduce_gradients(buf)                                 # the step's one collective
        assert torch.equal(buf[:flat_n], torch.full((flat_n,), float(sum(range(1, world + 1)))))
        assert float(buf[flat_n + tail.count_index]) == n            # what sr4d_adam_step_counted reads
        got, l2v, cnt = tail.read()
        assert cnt == n and l2v == 0.5
        np.testing.assert_array_equal(got, full.numpy())
    # validation steps exchange the tail alone
    per, _ = tail.begin(hi - lo)
    per.copy_(full[lo:hi] * 2)
    tail.exchange()
    np.testing.assert_array_equal(tail.read()[0], full.numpy() * 2)
    with pytest.raises(ValueError):
        par.MetricTail(torch.zeros(64), 16)                          # 2 ranks x 16 samples do not fit 64 floats


def test_metric_tail_single_collective():
    _run("_metric_tail_one_collective")


def test_dp_allreduce_reproduces_full_batch_gradient():
    _run("_dp_gradient_identity")


def test_ragged_row_gather():
    _run("_ragged_gather")


def test_metric_tail_deferred_read_ring():
    """read_begin / read_end (TrainerController.train_step folds a step's metrics after the NEXT step has been enqueued):
    two reads may be in flight; each token returns the values the tail held when its copy was enqueued, also after the
    tail has been reset and refilled for the following step (single process, CPU tensors)."""
    import importlib
    par = importlib.import_module("4dflownet_b200.parallel")
    tail_buf = torch.zeros(256)
    tail = par.MetricTail(tail_buf, max_batch=4, rank_=0, world=1)

    def fill(step, B):
        per, l2 = tail.begin(B)
        per.copy_(torch.arange(4 * B, dtype=torch.float32).view(B, 4) + 100 * step)
        l2.fill_(0.5 + step)

    fill(1, 3)
    t1 = tail.read_begin()
    fill(2, 2)                                   # the next step overwrites the tail before step 1 has been read
    t2 = tail.read_begin()
    assert t1 != t2
    per1, l2_1, n1 = tail.read_end(t1)
    per2, l2_2, n2 = tail.read_end(t2)
    assert n1 == 3 and n2 == 2 and l2_1 == 1.5 and l2_2 == 2.5
    assert per1.shape == (3, 4) and per1[0, 0] == 100 and per1[2, 3] == 111
    assert per2.shape == (2, 4) and per2[0, 0] == 200 and per2[1, 3] == 207
    fill(3, 4)
    per3, l2_3, n3 = tail.read()                 # the immediate form reuses the ring
    assert n3 == 4 and l2_3 == 3.5 and per3[3, 3] == 315
