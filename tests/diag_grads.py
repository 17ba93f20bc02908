"""Diagnostic (GPU box): per-tensor gradient error of the CUDA path and of torch-CPU fp32 autograd,
both against the fp64 oracle; plus quick timings of forward / train step."""
import importlib, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4dflownet_b200")
oracle = importlib.import_module("oracle.sr4d_oracle")

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)

args = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else [12, 2, 2, 2, 2]
P, r, low, hi, B = args
pseed, bseed = (int(sys.argv[6]), int(sys.argv[7])) if len(sys.argv) >= 8 else (7, 11)
params = oracle.glorot_params(low, hi, seed=pseed, bias_scale=0.05 if len(sys.argv) >= 8 else 0.02)
batch = oracle.synthetic_batch(B, P, r, seed=bseed)
eng = pkg.Engine(P, r, low, hi, max_batch=B, training=True, device=0)
eng.set_weights(params)
per, l2, _ = eng.train_fwd_bwd(batch[:6], [b[..., 0] for b in batch[6:9]], batch[10])
g64, _ = oracle.gradients({k: v.astype(np.float64) for k, v in params.items()}, batch, r, low, hi)
g32, _ = oracle.gradients(params, batch, r, low, hi, dtype=torch.float32)
l2c = oracle.L2_COEFF
print(f"{'tensor':24s} {'cuda vs f64':>12s} {'torch32 vs f64':>15s}")
for name, view in eng.tensor_views(eng.grads):
    corr = (B * 2 * l2c * params[name] if name.endswith('kernel') else 0.0)
    print(f"{name:24s} {rel(view.cpu().numpy(), g64[name]-corr):12.3e} {rel(g32[name]-corr, g64[name]-corr):15.3e}")

if len(sys.argv) >= 6: sys.exit(0)
# timings, config 2 geometry
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / n
B = 8
eng2 = pkg.Engine(24, 2, 8, 4, max_batch=B, training=True, device=0)
eng2.set_weights(oracle.glorot_params(8, 4, seed=1))
bt = oracle.synthetic_batch(B, 24, 2, seed=0)
dev = [torch.tensor(np.ascontiguousarray(b)).cuda() for b in bt]
hr = [d[..., 0].contiguous() for d in dev[6:9]]
tf = timeit(lambda: eng2.forward(dev[:6]))
print(f"SIMT forward  B={B}: {tf*1e3:.1f} ms  -> {B/tf:.1f} patches/s, {328.83*B/tf/1e3:.1f} TFLOP/s")
tt = timeit(lambda: eng2.train_fwd_bwd(dev[:6], hr, dev[10]))
print(f"SIMT fwd+bwd  B={B}: {tt*1e3:.1f} ms  -> {B/tt:.1f} patches/s")
