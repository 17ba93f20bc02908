"""GPU box, TIMING EXPERIMENT (wrong numerics on purpose): sustained forward loop with the conv kernel's L2 operand
streams switched off (SR4D_TC_EXP_SKIP bit 0 = weight taps re-use stale slots, bit 1 = activation planes re-use stale
stages).  Shows how much of the sustained (power-limited) forward time the weight / activation streams cost."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4dflownet_b200")
B, n = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 150
eng = pkg.Engine(24, 2, 8, 4, max_batch=B, training=False, device=0)
pkg.SR4DFlowModel.initialize(type('M', (), {'engine': eng})(), seed=1)
g = torch.Generator().manual_seed(0)
xs = [(torch.rand((B, 24, 24, 24), generator=g) * 2 - 1).cuda() for _ in range(3)] + \
     [(torch.rand((B, 24, 24, 24), generator=g) * 0.016).cuda() for _ in range(3)]
out = torch.empty((B, 48, 48, 48, 3), device="cuda")
for _ in range(20):
    eng.forward(xs, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    eng.forward(xs, out=out)
e1.record()
torch.cuda.synchronize()
print(f"SR4D_TC_EXP_SKIP={os.environ.get('SR4D_TC_EXP_SKIP', '0')}: {e0.elapsed_time(e1) / n:.3f} ms per forward (B=8, {n} back to back), finite={bool(torch.isfinite(out).all())}")
