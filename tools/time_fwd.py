"""GPU box: time forward / train step with CUDA events under both conv implementations."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4dflownet_b200")
synth = importlib.import_module("4dflownet_b200.utils.synthetic")
L = pkg._lib

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
r = int(sys.argv[2]) if len(sys.argv) > 2 else 2
train = (sys.argv[3] == "train") if len(sys.argv) > 3 else False
eng = pkg.Engine(24, r, 8, 4, max_batch=B, training=train, device=0)
pkg.SR4DFlowModel.initialize(type('M', (), {'engine': eng})(), seed=1)
bt = synth.synthetic_batch(B, 24, r, seed=0)
dev = [torch.tensor(np.ascontiguousarray(b)).cuda() for b in bt]
hr = [d[..., 0].contiguous() for d in dev[6:9]]
out = torch.empty((B, 24 * r, 24 * r, 24 * r, 3), device="cuda")
flops = {2: 328.83e9, 4: 2220.4e9}[r]
res = {}
for name, impl in (("simt", L.CONV_SIMT), ("tcgen05", L.CONV_TCGEN05)):
    eng.set_option(L.OPT_CONV_IMPL, impl)
    ms = timeit(lambda: eng.forward(dev[:6], out=out))
    res[name] = out.clone()
    print(f"{name:8s} forward B={B} r={r}: {ms:8.2f} ms  {B/ms*1e3:8.1f} patches/s  {flops*B/ms/1e9:7.1f} TFLOP/s(fp32-equiv)", flush=True)
    if train:
        ms = timeit(lambda: eng.train_fwd_bwd(dev[:6], hr, dev[10]), n=3, warm=1)
        print(f"{name:8s} fwd+bwd B={B}: {ms:8.2f} ms  {B/ms*1e3:8.1f} patches/s", flush=True)
d = (res["simt"] - res["tcgen05"]).abs().max().item() / res["simt"].abs().max().item()
print(f"full-network tc vs simt max rel diff: {d:.3e}")
