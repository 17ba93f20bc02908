"""2-GPU data-parallel parity (runs under `gpurun --gpus 2`; skipped on a 1-GPU box): one TrainerController per
rank on its shard + one NCCL all-reduce of the flat gradient buffer == the single-GPU step on the whole batch,
and sharded inference == single-GPU inference."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import contextlib
    import io
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pkg = importlib.import_module("4dflownet_b200")
        tcm = importlib.import_module("4dflownet_b200.Network.TrainerController")
        par = importlib.import_module("4dflownet_b200.parallel")
        oracle = importlib.import_module("oracle.sr4d_oracle")     # data / weights generator only
        P, r, low, hi, Bg = 8, 2, 1, 1, 4
        params = oracle.glorot_params(low, hi, seed=5, bias_scale=0.05)
        batch = oracle.synthetic_batch(Bg, P, r, seed=6)
        with contextlib.redirect_stdout(io.StringIO()):
            ctl = tcm.TrainerController(P, r, 1e-3, False, "t", low, hi, max_batch=Bg, device=rank)
        ctl.model.set_weights(params)
        ctl.train_step(par.shard_batch(batch))
        torch.cuda.synchronize()
        w_dp = [w.copy() for w in ctl.model.get_weights()]
        mean_loss = ctl.loss_metrics['train_loss'].result()
        # every rank must hold bit-identical weights after the step
        flat = ctl.engine.params.clone()
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        same = all(torch.equal(other[0], o) for o in other)
        # sharded inference: gather_rows of per-rank predictions == local full prediction
        lo, hi_ = par.shard_bounds(Bg)
        full = ctl.engine.forward(batch[:6])
        part = ctl.engine.forward([b[lo:hi_] for b in batch[:6]])
        gathered = par.gather_rows(part, Bg)
        inf_err = float((gathered - full).abs().max())
        if rank == 0:
            # single-GPU reference: same weights, whole batch, no process group involvement
            eng = pkg.Engine(P, r, low, hi, max_batch=Bg, training=True, device=0)
            eng.set_weights(params)
            per, l2, _ = eng.train_fwd_bwd(batch[:6], [b[..., 0] for b in batch[6:9]], batch[10])
            eng.adam_step(1e-3, 1, Bg * 2 * 5e-7)
            torch.cuda.synchronize()
            w_1 = eng.get_weights()
            worst = max(float(np.abs(a - b).max()) for a, b in zip(w_dp, w_1))
            loss_1 = float((per[:, 0].double().mean() + l2.double()).item())
            q.put((same, worst, inf_err, mean_loss, loss_1))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_matches_single_gpu():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    mp.spawn(_worker, args=(2, _free_port(), q), nprocs=2, join=True)
    same, worst, inf_err, loss_dp, loss_1 = q.get()
    assert same, "ranks diverged after the all-reduced Adam step"
    # first Adam step moves each weight by <= lr = 1e-3; shard-order summation differences are ~1e-7 relative
    assert worst < 2e-5, worst
    assert inf_err == 0.0, inf_err
    assert abs(loss_dp - loss_1) < 1e-5 * abs(loss_1), (loss_dp, loss_1)


def _trainer_main_worker(rank, world, port, d, q):
    """One torchrun-style rank of trainer.main: the environment torchrun exports, gloo so both ranks can share GPU 0."""
    import contextlib
    import io
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK="0", SR4D_DIST_BACKEND="gloo")
    import torch.distributed as dist
    trainer = importlib.import_module("4dflownet_b200.trainer")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    synth = importlib.import_module("synth")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            net = trainer.main(data_dir=d, QUICKSAVE=True, initial_learning_rate=1e-3, epochs=2, batch_size=4,
                               mask_threshold=0.6, network_name="t4d", patch_size=synth.PATCH, res_increase=synth.R,
                               low_resblock=1, hi_resblock=1, models_root=os.path.join(d, "models"))
        flat = net.engine.params.cpu()                     # (gloo gathers CPU tensors)
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        q.put((rank, net.model_dir, net.engine.max_batch, net.optimizer.iterations,
               all(torch.equal(other[0], o) for o in other), float(net.loss_metrics["train_loss"].result())))
        dist.barrier()
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_trainer_main_under_two_ranks(tmp_path):
    """`torchrun --nproc-per-node 2 trainer.py` in miniature: main() joins the process group from the
    environment, every rank trains its shard (engine sized for batch_size / ranks), the ranks end with bit-identical
    weights and the same running means, ONE model directory exists and only rank 0 wrote into it (one loss.csv line
    per epoch, one quicksave file predicted from the unsharded benchmark batch in chunks of max_batch)."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    synth = importlib.import_module("synth")
    d = str(tmp_path)
    synth.make_synthetic_h5(d)
    header = "source,target,index,start_x,start_y,start_z,rotate,rotation_plane,rotation_degree_idx,coverage\n"
    for name, rows in (("train.csv", synth.ROWS), ("validate.csv", synth.ROWS[:4]), ("benchmark.csv", synth.ROWS[2:6])):
        with open(os.path.join(d, name), "w") as f:
            f.write(header + "".join(",".join(r) + "\n" for r in rows))
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    mp.spawn(_trainer_main_worker, args=(2, _free_port(), d, q), nprocs=2, join=True)
    got = sorted(q.get() for _ in range(2))
    (r0, dir0, mb0, it0, same0, loss0), (r1, dir1, mb1, it1, same1, loss1) = got
    assert dir0 == dir1 and mb0 == mb1 == 2 and it0 == it1 and same0 and same1
    assert abs(loss0 - loss1) <= 1e-6 * abs(loss0)          # both ranks averaged the metrics of the whole global batch
    models = os.listdir(os.path.join(d, "models"))
    assert len(models) == 1
    md = os.path.join(d, "models", models[0])
    rows = [ln for ln in open(os.path.join(md, "loss.csv")).read().splitlines() if ln[:1].isdigit()]
    assert len(rows) == 2                                    # not one copy per rank
    assert os.path.exists(os.path.join(md, "quicksave_t4d.h5")) and os.path.exists(os.path.join(md, "t4d-best.h5"))
