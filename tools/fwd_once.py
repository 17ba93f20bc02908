"""GPU box: a few forward passes at config-2 geometry (for ncu captures)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4dflownet_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = pkg.Engine(24, 2, 8, 4, max_batch=B, training=False, device=0)
model_w = pkg.SR4DFlowModel.__new__(pkg.SR4DFlowModel)
model_w.engine = eng
model_w.initialize(seed=1)
g = torch.Generator().manual_seed(0)
xs = [(torch.rand((B, 24, 24, 24), generator=g) * 2 - 1).cuda() for _ in range(3)] + \
     [(torch.rand((B, 24, 24, 24), generator=g) * 0.016).cuda() for _ in range(3)]
out = torch.empty((B, 48, 48, 48, 3), device="cuda")
for _ in range(n):
    eng.forward(xs, out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
