"""GPU box: a few train steps (forward + loss + backward + Adam) at config-2 geometry, for ncu captures."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4dflownet_b200")
synth = importlib.import_module("4dflownet_b200.utils.synthetic")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = pkg.Engine(24, 2, 8, 4, max_batch=B, training=True, device=0)
pkg.SR4DFlowModel.initialize(type('M', (), {'engine': eng})(), seed=1234)
bt = [torch.tensor(np.ascontiguousarray(b)).cuda() for b in synth.synthetic_batch(B, 24, 2, seed=0)]
hr = [d[..., 0].contiguous() for d in bt[6:9]]
for it in range(n):
    per, l2, _ = eng.train_fwd_bwd(bt[:6], hr, bt[10])
    eng.adam_step(1e-4, it + 1, B * 1e-6)
torch.cuda.synchronize()
print("ok", per[:, 0].tolist(), "launches/step", eng.launch_count() // n)
