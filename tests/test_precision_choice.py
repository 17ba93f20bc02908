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


def test_single_fp16_gradient_operand_keeps_fp32_gradient_accuracy():
    """Evidence behind SR4D_OPT_DGRAD_SINGLE (tools/gradient_precision_emulation.py, small network): with W / X split, a
    single power-of-two-scaled fp16 gradient operand costs ~2e-5 of the flat gradient at 2 x 16^3 voxels (its per-voxel
    rounding errors average out in the sums, ~1/sqrt(#voxels): 0.4e-5 at 48^3, below fp32 autograd's own 1e-5 there);
    rounding W / X to one fp16 value costs 2.4e-4 at any size."""
    spec = importlib.util.spec_from_file_location("gpe", os.path.join(ROOT, "tools", "gradient_precision_emulation.py"))
    gpe = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gpe)
    res = gpe.run(8, 2, 2, 1, 2)
    fp32 = res["fp32"]
    single_grad = res["fp16 split, 2 products (W / X split, gradient single)"]
    single_all = res["fp16 (1 pass)"]
    assert res["fp16 split, 3 products (csrc)"][0] < 2 * fp32[0] + 1e-6
    assert fp32[0] < 5e-6
    assert single_grad[0] < 5e-5
    assert single_all[0] > 5 * single_grad[0]
