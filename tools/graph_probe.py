"""GPU-box experiment: does replaying the forward+backward launch sequence as a CUDA graph shorten the step?"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4dflownet_b200")
synth = importlib.import_module("4dflownet_b200.utils.synthetic")
B = 8
eng = pkg.Engine(24, 2, 8, 4, max_batch=B, training=True, device=0)
pkg.SR4DFlowModel.initialize(type('M', (), {'engine': eng})(), seed=1234)
bt = [torch.tensor(np.ascontiguousarray(b)).cuda() for b in synth.synthetic_batch(B, 24, 2, seed=0)]
hr = [d[..., 0].contiguous() for d in bt[6:9]]

def step():
    eng.train_fwd_bwd(bt[:6], hr, bt[10])

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("eager fwd+bwd ms:", timeit(step))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2): step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, stream=s):
        step()
    torch.cuda.synchronize()
    print("graph fwd+bwd ms:", timeit(g.replay))
    print("eager again   ms:", timeit(step))
except Exception as e:
    print("capture failed:", repr(e)[:300])
