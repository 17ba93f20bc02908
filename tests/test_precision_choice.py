"""The operand-precision choice of the tensor-core convolution, as a regression test of its evidence
(tools/precision_emulation.py): on the CPU, with the thirty-layer structure shrunk to a small network, a single TF32 /
BF16 / FP16 pass and a two-product FP16 split miss the north-star bar of 1e-4 (max|d|/max|ref| vs the float64 oracle),
the three-product split the CUDA kernel implements is as accurate as fp32."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_three_product_split_is_the_cheapest_mode_inside_the_bar():
    spec = importlib.util.spec_from_file_location("precision_emulation", os.path.join(ROOT, "tools", "precision_emulation.py"))
    emu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emu)
    res = emu.run(8, 2, 2, 1)
    assert res["fp32"] < 2e-5
    assert res["fp16 split, 3 products (conv_tc.cu)"] < 2e-5
    for single in ("tf32 (1 pass)", "bf16 (1 pass)", "fp16 (1 pass)", "fp16 split, 2 products (no Wlo*Xhi)"):
        assert res[single] > 1e-4, (single, res[single])
