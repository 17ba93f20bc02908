"""GPU-box probe for the tcgen05 conv: structured test kernels to localise layout bugs."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4dflownet_b200")
L = pkg._lib

def run(case, D, B):
    eng = pkg.Engine(8, 2, 0, 0, max_batch=2, training=False, device=0)
    g = np.random.default_rng(0)
    x = g.standard_normal((B, D, D, D, 64)).astype(np.float32)
    k = np.zeros((3, 3, 3, 64, 64), np.float32)
    if case == "center_identity":
        k[1, 1, 1] = np.eye(64)
    elif case == "center_random":
        k[1, 1, 1] = g.standard_normal((64, 64)) * 0.1
    elif case == "tap_z":
        k[1, 1, 2] = np.eye(64)
    elif case == "tap_y":
        k[1, 2, 1] = np.eye(64)
    elif case == "tap_x":
        k[0, 1, 1] = np.eye(64)
    else:
        k = (g.standard_normal((3, 3, 3, 64, 64)) * 0.04).astype(np.float32)
    ys = eng.conv64_layer(x, k, impl=L.CONV_SIMT).cpu().numpy()
    yt = eng.conv64_layer(x, k, impl=L.CONV_TCGEN05).cpu().numpy()
    err = np.abs(yt - ys).max() / (np.abs(ys).max() + 1e-30)
    print(f"{case:16s} D={D:3d} B={B}: max rel err tc vs simt = {err:.3e}", flush=True)
    if err > 1e-4:
        bad = np.argwhere(np.abs(yt - ys) > 1e-3 * np.abs(ys).max())
        print("   n_bad", len(bad), "of", ys.size, "first", bad[:6].tolist())
        print("   bad co hist (mod 16):", np.bincount(bad[:, 4] % 16, minlength=16).tolist())
        print("   bad z hist:", np.bincount(bad[:, 3], minlength=D).tolist())
        print("   bad y hist:", np.bincount(bad[:, 2], minlength=D).tolist())
        print("   sample tc", yt[tuple(bad[0])], "simt", ys[tuple(bad[0])])
    eng.close()

if __name__ == "__main__":
    case, D, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    run(case, D, B)
