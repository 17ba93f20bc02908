"""BASELINE.json configs 1, 3 and 5 (the forward / tiled-inference cases; config 2/4 is bench.py itself).
    python tools/bench_configs.py --config 1|3|5 [--steps K]         (torchrun for N > 1)
One JSON line per run on rank 0.  Inputs are synthetic; timing = CUDA events, max over ranks."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[1, 3, 5])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("4dflownet_b200")
    predictor = importlib.import_module("4dflownet_b200.predictor")
    par = importlib.import_module("4dflownet_b200.parallel")
    dev = torch.device("cuda", local)
    g = torch.Generator().manual_seed(0)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    if args.config in (1, 3):
        P, r, B = (24, 2, 1) if args.config == 1 else (24, 4, 16)
        n = 12 if args.config == 1 else B          # config 1: the 12 patches of the 42x38x36 example volume, batch 1
        model = pkg.prepare_network(P, r, 8, 4, max_batch=B, device=local)
        xs = [(torch.rand((n, P, P, P), generator=g) * 2 - 1) for _ in range(3)] + \
             [(torch.rand((n, P, P, P), generator=g) * 0.016) for _ in range(3)]
        xd = [x.to(dev) for x in xs]
        out = torch.empty((n, P * r, P * r, P * r, 3), device=dev)

        def step():
            for i in range(0, n, B):
                model.engine.forward([x[i:i + B] for x in xd], out=out[i:i + B])
        ms = timed(step)
        flops = {2: 328.83e9, 4: 2220.4e9}[r]
        line = {"config": args.config, "workload": f"forward P={P} r={r} batch={B} ({n} patches per step)", "n_gpus": world,
                "ms_per_step": ms, "patches_per_s": n * world / ms * 1e3, "tflops_fp32_equiv": flops * n / ms / 1e9}
    else:
        # config 5: 160x160x64 volume, P=24, reference tiling (stride 20 -> 256 patches), sharded over the ranks
        class DS:
            pass
        rng = np.random.default_rng(0)
        ds = DS()
        for nme in ("u", "v", "w"):
            setattr(ds, nme, rng.uniform(-1, 1, (160, 160, 64)).astype(np.float32))
        for nme in ("mag_u", "mag_v", "mag_w"):
            setattr(ds, nme, rng.uniform(0, 0.016, (160, 160, 64)).astype(np.float32))
        ds.venc = np.float32(1.5)
        ds.velocity_per_px = ds.venc / 2048
        model = pkg.prepare_network(24, 2, 8, 4, max_batch=8, device=local)
        pg = pkg.PatchGenerator(24, 2)
        vol = [None]

        def step():
            vol[0] = predictor.predict_volume(model, pg, ds, batch_size=8, reuse_host_buffer=True)
        ms = timed(step)
        npatch = pg.nr_x * pg.nr_y * pg.nr_z
        line = {"config": 5, "workload": "tiled inference 160x160x64, P=24 stride 20 (reference tiling), host volume in -> "
                "stitched host volume out (patchify + H2D + forward + gather + GPU stitch + D2H)", "n_gpus": world,
                "patches": npatch, "ms_per_volume": ms, "patches_per_s": npatch / ms * 1e3, "out_shape": list(vol[0].shape) if vol[0] is not None else None}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
