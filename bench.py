#!/usr/bin/env python
"""Benchmark of the patch-based SR hot path (BASELINE.json metric: 24^3->48^3 SR patches/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference graph on the host CPU

Workload (BASELINE.json configs[1]): one trainer.py step -- forward + masked-MSE loss/metric +
backward + Adam -- at patch_size=24, res_increase=2, 8 low / 4 hi resblocks, batch 8 PER GPU
(weak scaling; N=8 is configs[3], global batch 64, one NCCL all-reduce of the flat gradient
buffer per step).  A forward-only leg (the predictor.py path, same geometry) is timed in the
same run and reported under "forward" because the metric also asks for the forward roofline.

One JSON line on stdout (rank 0).  `value` = patches/s with the batch already resident in HBM;
`e2e` = the same step through TrainerController.train_step from pinned HOST buffers, including
the host->device copies of the 11-tuple and the device->host read of the per-sample metrics.
The CPU oracle (oracle/) is used here only as the timed `cpu_baseline` / `--impl reference` arm.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P, R, LOW, HI = 24, 2, 8, 4
FLOPS_FWD_PER_PATCH = 328.83e9        # SURVEY 8d, sum over layers of 2*k^3*Ci*Co*D^3 (r=2)
BYTES_FWD_PER_PATCH = 1037.4e6        # SURVEY 8d, layer-wise algorithmic HBM bytes at B=8
FLOPS_CONV64 = {"lr": 2 * 27 * 64 * 64 * P ** 3, "hr": 2 * 27 * 64 * 64 * (P * R) ** 3}   # per patch per launch
METRIC = "24^3->48^3 SR patches/sec"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return {"hbm_gbs": float(d["hbm_gbs"]), "tflops_burst": float(d["bf16_tflops"]),
                    "tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    # B200_PROFILING.md fallback: 6.65 TB/s copy, 1.59 PF cuBLAS bf16 burst (~1.4 PF sustained)
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region: the sampler is started before the
    warm-up (nvidia-smi needs a moment to produce its first line, more so on an 8-GPU box) and only samples whose
    timestamp falls inside [mark_begin, mark_end] are used; if the region was too short to catch one, the samples
    taken under load during the warm-up right before it are used and the summary says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thr = index, [], None, None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=lambda: self.rows.extend(self.proc.stdout), daemon=True)
        self.thr.start()

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    @staticmethod
    def _epoch(ts):
        import datetime
        try:
            return datetime.datetime.strptime(ts.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thr.join(timeout=2)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                parsed.append((self._epoch(f[0]), float(f[1]), float(f[2]), float(f[3]),
                               [n for n, v in zip(names, f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        inside = [p for p in parsed if p[0] is not None and t0 <= p[0] <= t1 + 0.05]
        window = "timed region"
        if not inside:
            # region shorter than the sampling latency: use the last samples before it ends (warm-up under the same load)
            inside = [p for p in parsed if p[0] is None or p[0] <= t1 + 0.05][-6:]
            window = "warm-up + timed region (region too short for a sample of its own)"
        sm = [p[1] for p in inside]
        reasons = sorted({r for p in inside for r in p[4]})
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max((p[2] for p in inside), default=None),
                "power_w": statistics.median([p[3] for p in inside]) if inside else None,
                "samples": len(inside), "window": window, "reasons": reasons}


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's restatement of the reference graph on host cores
# --------------------------------------------------------------------------------------------
def cpu_train_steps(n_steps, warmup, batch=1):
    """Times `n_steps` full train steps (fwd + loss + autograd bwd + Adam) of the torch-CPU fp32
    restatement (oracle/sr4d_oracle.py) on all host cores, `batch` patches per step."""
    import numpy as np
    import torch
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently turn the CPU arm into a single-threaded run
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncpu))
    oracle = importlib.import_module("oracle.sr4d_oracle")
    params = {k: torch.tensor(v, requires_grad=True) for k, v in oracle.glorot_params(LOW, HI, seed=1234).items()}
    m = {k: np.zeros(v.shape, np.float32) for k, v in params.items()}
    v2 = {k: np.zeros(v.shape, np.float32) for k, v in params.items()}
    bt = [torch.tensor(np.asarray(b)) for b in oracle.synthetic_batch(batch, P, R, seed=0)]
    times = []
    for it in range(warmup + n_steps):
        t0 = time.perf_counter()
        for p in params.values():
            p.grad = None
        obj, _ = oracle.train_objective(params, bt, R, LOW, HI)
        obj.backward()
        with torch.no_grad():
            for k, p in params.items():
                pn, m[k], v2[k] = oracle.adam_step(p.numpy(), p.grad.numpy(), m[k], v2[k], it + 1, 1e-4)
                p.copy_(torch.from_numpy(np.asarray(pn, dtype=np.float32)))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, torch.get_num_threads()


def run_reference(args):
    """The reference graph on the host CPU, same workload as the B200 arm: one train step of `--batch` (8) patches.
    Under torchrun only rank 0 works (ONE CPU process on the box's host cores, whatever N); `n_gpus` repeats the
    launch's N so the driver can pair the lines, `cpu_processes` says what actually ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, cores = cpu_train_steps(args.steps, args.warmup, batch=args.batch)
    total = sum(times)
    val = len(times) * args.batch / total
    sample = (f"{args.batch} patches per step (full train step: fwd+loss+bwd+Adam) of the same P=24,r=2,8/4 network, "
              f"{len(times)} timed steps after {args.warmup} warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "patches/s", "n_gpus": args.gpus,
        "cpu_processes": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.gpus, args.batch),
                       note="CPU arm: one process on rank 0's host cores steps ONE shard (batch_per_gpu patches) per step; "
                            "its patches/s does not grow with N"),
        "cpu_baseline": {"value": val, "unit": "patches/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "TensorFlow is not installable here; torch-CPU (oneDNN) restatement of the "
                                 "reference graph stands in for the TF-CPU path"},
        "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, batch_per_gpu):
    return {"workload": "configs[1]: trainer.py single step (forward + masked-MSE loss/metric + backward + Adam), "
                        "patch_size=24, res_increase=2, 8 low / 4 hi resblocks, fp32",
            "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * n_gpus, "parallelism": f"dp{n_gpus}",
            "l2": "no explicit flush: one step streams ~3.3 GB of saved activations per GPU (>> 126 MB L2)"}



# --------------------------------------------------------------------------------------------
# the other BASELINE.json configs, measured in the same run (extra keys of the one JSON line)
# --------------------------------------------------------------------------------------------
def run_extra_configs(pkg, dev, local, world, timed, steps):
    """configs[0] (predictor forward, batch 1), configs[2] (r=4, batch 16) and configs[4] (tiled 160x160x64 volume,
    sharded over the ranks); configs[1]/[3] are the headline itself.  Device timing, max over ranks."""
    import numpy as np
    import torch
    predictor = importlib.import_module("4dflownet_b200.predictor")
    g = torch.Generator().manual_seed(0)
    out = {}
    for key, (Pp, r, Bc, n) in {"configs[0]": (24, 2, 1, 12), "configs[2]": (24, 4, 16, 16)}.items():
        # configs[0]: the 12 patches of the 42x38x36 example volume one by one (predictor.py batch loop at batch 1)
        model = pkg.prepare_network(Pp, r, LOW, HI, max_batch=Bc, device=local)
        xs = [(torch.rand((n, Pp, Pp, Pp), generator=g) * 2 - 1) for _ in range(3)] + \
             [(torch.rand((n, Pp, Pp, Pp), generator=g) * 0.016) for _ in range(3)]
        xd = [x.to(dev) for x in xs]
        y = torch.empty((n, Pp * r, Pp * r, Pp * r, 3), device=dev)

        def step():
            for i in range(0, n, Bc):
                model.engine.forward([x[i:i + Bc] for x in xd], out=y[i:i + Bc])
        ms = timed(step, steps, 2) / steps
        host_np = [x.numpy() for x in xs]

        def step_e2e():
            model.predict(host_np, batch_size=Bc)
        ms_e2e = timed(step_e2e, max(1, steps // 2), 1) / max(1, steps // 2)
        flops = {2: 328.83e9, 4: 2220.4e9}[r]
        out[key] = {"workload": f"forward patch_size={Pp} res_increase={r} batch={Bc} ({n} patches per step per GPU)",
                    "value": n * world / ms * 1e3, "unit": "patches/s", "ms_per_step": ms,
                    "tflops_fp32_equiv": flops * n * world / ms / 1e9,
                    "e2e": {"value": n * world / ms_e2e * 1e3, "unit": "patches/s",
                            "note": "model.predict: numpy patches in, numpy predictions out"}}
        del model, xd, y
        torch.cuda.empty_cache()

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
    model = pkg.prepare_network(24, 2, LOW, HI, max_batch=8, device=local)
    pg = pkg.PatchGenerator(24, 2)

    def step_vol():
        predictor.predict_volume(model, pg, ds, batch_size=8, reuse_host_buffer=True)
    ms = timed(step_vol, steps, 1) / steps
    npatch = pg.nr_x * pg.nr_y * pg.nr_z
    out["configs[4]"] = {"workload": "tiled inference of a 160x160x64 volume, patch_size=24 (reference tiling, stride 20: 256 "
                         "patches) sharded over the ranks: host volume in -> stitched 320x320x128 host volume out "
                         "(patchify + H2D + forward + gather + GPU stitch + D2H inside the timed region)",
                         "value": npatch / ms * 1e3, "unit": "patches/s", "ms_per_volume": ms, "patches": npatch,
                         "scaling": "strong"}
    return out


def dp_parity_preflight(pkg, tcm, dev, local, rank, world):
    """Data-parallel pre-flight (N > 1), small geometry: after the step's one all-reduce every rank must hold the
    bit-identical flat gradient, and it must equal the single-GPU gradient of the concatenated global batch up to
    fp32 summation order.  Run once with the default (tensor-core) kernels and once with the fp32 SIMT anchor."""
    import contextlib
    import io
    import numpy as np
    import torch
    import torch.distributed as dist
    par = importlib.import_module("4dflownet_b200.parallel")
    synthetic_batch = importlib.import_module("4dflownet_b200.utils.synthetic").synthetic_batch
    Pp, r, low, hi, Bl = 8, 2, 1, 1, 2
    glob = synthetic_batch(Bl * world, Pp, r, seed=77)
    shard = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in par.shard_batch(glob)]
    res = {"geometry": f"patch_size={Pp} res_increase={r} {low}/{hi} blocks, {Bl} samples per rank"}
    for label, impl in (("default", pkg._lib.CONV_AUTO), ("simt_fp32", pkg._lib.CONV_SIMT)):
        with contextlib.redirect_stdout(io.StringIO()):
            ctl = tcm.TrainerController(Pp, r, 1e-4, False, "dp", low, hi, max_batch=Bl, device=local, seed=5)
        ctl.engine.set_option(pkg._lib.OPT_CONV_IMPL, impl)
        tail = ctl._tail()
        per, l2 = tail.begin(Bl)
        ctl.engine.train_fwd_bwd(shard[:6], [a[..., 0] for a in shard[6:9]], shard[10], per_out=per, l2_out=l2)
        par.allreduce_gradients(ctl.engine.grads_full)
        flat = ctl.engine.grads.clone()
        allf = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(allf, flat)
        same = all(torch.equal(allf[0], f) for f in allf)
        _, _, count = tail.read()
        rel = None
        if rank == 0:
            eng1 = pkg.Engine(Pp, r, low, hi, max_batch=Bl * world, training=True, device=local)
            eng1.set_option(pkg._lib.OPT_CONV_IMPL, impl)
            eng1.set_weights(ctl.model.get_weights())
            eng1.train_fwd_bwd(glob[:6], [a[..., 0] for a in glob[6:9]], glob[10])
            rel = float(((flat.double() - eng1.grads.double()).norm() / eng1.grads.double().norm()).item())
            eng1.close()
        res[label] = {"bit_identical_across_ranks": bool(same), "rel_l2_vs_single_gpu": rel,
                      "global_count_in_tail": count}
        ctl.engine.close()
    if rank == 0:
        res["ok"] = bool(res["default"]["bit_identical_across_ranks"] and res["simt_fp32"]["bit_identical_across_ranks"]
                         and res["simt_fp32"]["rel_l2_vs_single_gpu"] <= 1e-6
                         and res["default"]["rel_l2_vs_single_gpu"] <= 1e-5
                         and res["default"]["global_count_in_tail"] == Bl * world)
    return res


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("4dflownet_b200")
    tcm = importlib.import_module("4dflownet_b200.Network.TrainerController")
    synthetic_batch = importlib.import_module("4dflownet_b200.utils.synthetic").synthetic_batch
    B, K, W = args.batch, args.steps, args.warmup
    dev = torch.device("cuda", local)

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        ctl = tcm.TrainerController(P, R, 1e-4, False, "4DFlowNet", LOW, HI, max_batch=B, device=local, seed=1234)
    eng = ctl.engine
    if args.two_plane_backward:
        # A/B switch: the round-1 backward (both planes of the split gradient in the dgrad, two-plane wgrad kernel).
        # The line is labelled; the default since round 2 is the single-plane backward (same gradient parity,
        # profiles/r02_grad_parity.txt).
        L = importlib.import_module("4dflownet_b200._lib")
        eng.set_option(L.OPT_DGRAD_SINGLE, 0)
        eng.set_option(L.OPT_WGRAD_SINGLE, 0)
    host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in synthetic_batch(B, P, R, seed=rank)]
    devb = [h.to(dev) for h in host]
    h2d = sum(h.numel() * 4 for i, h in enumerate(host) if i != 9)     # venc is not an input of the step
    tail = ctl._tail()
    d2h = (world * tail.slot + 1) * 4            # the metric tail read back per step (every rank's slot + the count)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_dev():
        ctl.train_step_async(devb)

    def step_e2e():
        # the call a user makes: TrainerController.train_step on a batch that lives in (pinned) host memory.  The
        # step uploads the 11-tuple, runs forward + loss + backward + the ONE all-reduce + Adam, and reads the
        # all-reduced metric tail back (one small D2H + the step's only stream synchronisation) for the running means.
        # (host tensors go in as they are: train_step uploads the low-resolution inputs on its stream and the HR
        # targets + mask on a side stream underneath the forward, engine.train_fwd_bwd)
        ctl.train_step(host)

    # ---- headline: device-resident train step --------------------------------------------------
    clocks = ClockSampler(local)
    clocks.start()
    eng.reset_launch_count()
    for _ in range(W):
        step_dev()
    barrier()
    eng.reset_launch_count()
    clocks.mark_begin()
    ms_total = timed(step_dev, K, 0)
    clocks.mark_end()
    clk = clocks.stop()
    launches = eng.launch_count()
    ms_step = ms_total / K
    value = B * world / (ms_step * 1e-3)

    # ---- e2e from pinned host buffers -------------------------------------------------------------
    ms_e2e = timed(step_e2e, K, max(1, W // 2)) / K
    e2e_val = B * world / (ms_e2e * 1e-3)

    # ---- per-kernel-class device time (CUDA events on the launch stream), same steps -------------
    eng.profile(True)
    for _ in range(K):
        step_dev()
    prof = eng.profile_read()
    eng.profile(False)
    prof_step = {k: (ms / K, n // K) for k, (ms, n) in prof.items()}

    # ---- forward-only leg (predictor path), device resident + e2e through model.predict ----------
    out = torch.empty((B, P * R, P * R, P * R, 3), device=dev)

    def fwd_dev():
        eng.forward(devb[:6], out=out)
    ms_fwd = timed(fwd_dev, K, W) / K
    eng.profile(True)
    for _ in range(K):
        fwd_dev()
    fprof = eng.profile_read()
    eng.profile(False)
    NB = 4                                            # predictor.py loops over the patch list in batches (:82-94)
    host_np = [np.concatenate([h.numpy()] * NB) for h in host[:6]]

    def fwd_e2e():
        ctl.model.predict(host_np, batch_size=B)
    ms_fwd_e2e = timed(fwd_e2e, max(1, K // 2), 1) / max(1, K // 2) / NB

    extra = {} if args.no_extra_configs else run_extra_configs(pkg, dev, local, world, timed, max(3, K // 3))
    dp = dp_parity_preflight(pkg, tcm, dev, local, rank, world) if world > 1 else None

    peaks = measured_peaks()
    if rank == 0:
        # dominant kernel class of the train step
        dom = max(prof_step, key=lambda k: prof_step[k][0])
        dom_ms, dom_n = prof_step[dom]
        grid = "hr" if dom.endswith("hr") else "lr"
        flops_launch = FLOPS_CONV64[grid] * B
        ach = flops_launch / (dom_ms / max(dom_n, 1) * 1e-3) / 1e12 if dom_n else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                with open(tpath) as f:
                    traffic = json.load(f).get(dom)
            except Exception:
                traffic = None
        peak = peaks["tflops_sustained"]
        # fp16 tensor-core products the kernel class EXECUTES per algorithmic fp32 MAC
        if "fwd" in dom or args.two_plane_backward:
            exe, exe_note = 4.0, ("the fp32-accurate split-fp16 path issues 4 fp16 tcgen05 products per fp32 MAC (3 useful + the "
                                  "64 masked-off rows of the second instruction): against a pipe that saturates at `peak` the fraction could not exceed 0.25; "
                                  "`peak` is the MEASURED sustained cuBLAS bf16 rate, which this kernel's executed rate exceeds when "
                                  "executed_frac > 1")
        elif "dgrad" in dom:
            exe, exe_note = 2.0, ("single-plane dgrad: [Wlo;Whi] x dYhi, 2 fp16 products per fp32 MAC (tensor pipe saturates at frac "
                                  "~0.5); the launch also folds the halo, adds the skip gradient, applies act' and writes the split copy")
        else:
            exe, exe_note = 4.0 / 3.0, ("stacked single-plane wgrad: dYhi x Xhi with two dY planes on M, 8 instructions per 6 useful "
                                        "tap groups (1.33 fp16 products per fp32 MAC)")
        f_hr_ms, f_hr_n = fprof["conv64_fwd_hr"]
        f_lr_ms, f_lr_n = fprof["conv64_fwd_lr"]
        fwd = {
            "value": B * world / (ms_fwd * 1e-3), "unit": "patches/s", "ms_per_step": ms_fwd,
            "e2e": {"value": B * world / (ms_fwd_e2e * 1e-3), "unit": "patches/s",
                    "h2d_bytes_per_step": 6 * B * P ** 3 * 4, "d2h_bytes_per_step": B * (P * R) ** 3 * 3 * 4,
                    "note": "model.predict on a host patch list of 4 batches (numpy in, numpy out), per batch"},
            "hbm_gbs_algorithmic": BYTES_FWD_PER_PATCH * B / (ms_fwd * 1e-3) / 1e9,
            "hbm_frac": BYTES_FWD_PER_PATCH * B / (ms_fwd * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "tflops_fp32_equiv": FLOPS_FWD_PER_PATCH * B / (ms_fwd * 1e-3) / 1e12,
            "conv64_fwd_hr": {"ms_per_launch": f_hr_ms / max(f_hr_n, 1), "launches_per_step": f_hr_n // K,
                              "tflops_fp32_equiv": FLOPS_CONV64["hr"] * B / (f_hr_ms / max(f_hr_n, 1) * 1e-3) / 1e12
                              if f_hr_n else None},
            "conv64_share_of_step": (f_hr_ms + f_lr_ms) / K / ms_fwd,
        }
        line = {
            "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world, B), **({"two_plane_backward": "round-1 backward kernels (A/B run)"}
                                                        if args.two_plane_backward else {})),
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                         "frac": ach / peak, "traffic": traffic,
                         "traffic_source": "profiles/traffic.json: dram bytes per launch of this kernel class from an "
                                           "ncu --set full capture (not re-measured in this run)",
                         "peak_source": peaks["source"] + ", sustained bf16",
                         "executed_tflops_fp16": exe * ach, "executed_frac": exe * ach / peak,
                         "ms_per_launch": dom_ms / max(dom_n, 1), "launches_per_step": dom_n,
                         "share_of_step": dom_ms / ms_step,
                         "note": "achieved = algorithmic fp32 FLOPs (2*27*64*64*B*D^3) / event time.  " + exe_note},
            "kernel_classes_ms_per_step": {k: round(v[0], 4) for k, v in prof_step.items()},
            "forward": fwd,
            "other_configs": extra,
        }
        if dp is not None:
            line["dp_parity"] = dp
        if world == 1 and not args.no_cpu_baseline:
            times, cores = cpu_train_steps(3, 1, batch=B)
            line["cpu_baseline"] = {"value": len(times) * B / sum(times), "unit": "patches/s", "cores": cores,
                                    "kind": "port", "sample": f"3 timed train steps of {B} patches (after 1 warm-up, ~20 s of "
                                    "CPU work) on the torch-CPU restatement of the reference graph (TF not installable)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="patches per GPU per step (configs[1]: 8)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the configs[0] / configs[2] / configs[4] legs (quick A/B runs)")
    ap.add_argument("--two-plane-backward", action="store_true",
                    help="A/B: SR4D_OPT_DGRAD_SINGLE = SR4D_OPT_WGRAD_SINGLE = 0 (the round-1 backward kernels)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
