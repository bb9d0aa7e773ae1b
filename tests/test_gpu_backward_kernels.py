"""Per-layer parity of the BACKWARD kernels (SURVEY.md section 4 "unit (GPU)"): the weight-gradient kernel and the
data-gradient convolution of every distinct conv block of the default net (SURVEY App. A shapes, incl. the 64 / 128
channel layers that take the deep weight-gradient kernel), called through octseg_debug_backward_block, against float64
torch convolutions of the SAME inputs (bf16 mode: the inputs as the kernel sees them, i.e. rounded to bf16, so the
bound isolates the kernel from storage noise)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oct_image_segmentation_models_b200.common.synthetic import synthetic_weights
from oct_image_segmentation_models_b200.models.unet_spec import unet_blocks

pytestmark = pytest.mark.gpu
CFG = dict(input_channels=1, num_classes=4)


def bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float64)


def reference_grads(b, w_hwio, a_in, dz, round_bf16):
    """float64 autograd of z = conv(pad(upsample?(a_in)), w) for dL/dz = dz"""
    a = (bf16(a_in) if round_bf16 else torch.from_numpy(a_in).double()).permute(0, 3, 1, 2).requires_grad_(True)
    g = (bf16(dz) if round_bf16 else torch.from_numpy(dz).double()).permute(0, 3, 1, 2)
    w = (bf16(w_hwio) if round_bf16 else torch.from_numpy(w_hwio).double()).permute(3, 2, 0, 1).requires_grad_(True)
    x = F.interpolate(a, scale_factor=2, mode="nearest") if b.upsample_before else a
    pt, pl = (b.kh - 1) // 2, (b.kw - 1) // 2
    x = F.pad(x, (pl, b.kw - 1 - pl, pt, b.kh - 1 - pt))
    z = F.conv2d(x, w)
    z.backward(g)
    return (w.grad.permute(2, 3, 1, 0).numpy(), g.sum((0, 2, 3)).numpy(), a.grad.permute(0, 2, 3, 1).numpy())


@pytest.fixture(scope="module")
def engines():
    from oct_image_segmentation_models_b200.engine import UNetEngine
    w = synthetic_weights(seed=42, **CFG)
    out = {}
    for prec in ("fp32", "bf16"):
        e = UNetEngine(precision=prec, **CFG)
        e.set_weights(w)
        e.train_begin([0.5, 1.0, 2.0, 1.0], global_batch=4)
        out[prec] = e
    yield w, out
    for e in out.values():
        e.close()


# (block index, n, h, w) -- h, w = the block's INPUT grid; every (k, Cin, Cout, upsample) of the default net
CASES = [(1, 2, 32, 24), (2, 2, 32, 16), (3, 1, 32, 32), (4, 2, 16, 16), (5, 2, 32, 16), (6, 2, 16, 16), (7, 3, 16, 16),
         (8, 4, 16, 8), (9, 4, 16, 16), (10, 2, 16, 8), (11, 2, 32, 16), (12, 2, 16, 32), (13, 2, 16, 16), (14, 1, 32, 32),
         (15, 2, 32, 16), (16, 2, 16, 16), (17, 1, 32, 32), (18, 2, 32, 16), (19, 2, 32, 24), (20, 2, 32, 24), (21, 2, 48, 40),
         (9, 8, 32, 16), (11, 1, 64, 32),
         # row-walking weight gradient of the narrow layers: enough tiles per persistent CTA to wrap its TMA stage ring
         # (8 -> 8: 960 tiles of 32x32 over 148 CTAs, 6 stages), the 16 -> 16 variant, and the up-conv read from the
         # LOW-res tensor over several tiles with a partial one
         (1, 30, 256, 128), (3, 6, 128, 64), (19, 4, 64, 40), (17, 3, 64, 64), (14, 5, 64, 96), (13, 3, 32, 48)]


@pytest.mark.parametrize("idx,n,h,w", CASES)
def test_wgrad_and_dgrad_kernels_vs_float64(engines, idx, n, h, w):
    weights, eng = engines
    b = unet_blocks(**CFG)[idx]
    rng = np.random.default_rng(500 + idx)
    a_in = np.maximum(rng.normal(0.2, 1.0, size=(n, h, w, b.cin)), 0).astype(np.float32)
    oh, ow = (2 * h, 2 * w) if b.upsample_before else (h, w)
    dz = rng.normal(0.0, 1.0, size=(n, oh, ow, b.cout)).astype(np.float32)
    wk = weights[6 * idx]
    for prec, tol_w, tol_d in (("fp32", 2e-5, 2e-5), ("bf16", 4e-3, 1.2e-2)):
        dW, db, din = eng[prec].debug_backward_block(idx, a_in, dz)
        rW, rb, rin = reference_grads(b, wk, a_in, dz, round_bf16=(prec == "bf16"))
        # weight / bias gradient: fp32 accumulation of exact products of the (rounded) inputs
        eW = np.abs(dW - rW).max() / max(np.abs(rW).max(), 1e-6)
        eb = np.abs(db - rb).max() / max(np.abs(rb).max(), 1e-6)
        # data gradient: the bf16 kernel also rounds its OUTPUT to bf16 (2^-9 relative)
        ed = np.abs(din - rin).max() / max(np.abs(rin).max(), 1e-6)
        assert np.isfinite(dW).all() and np.isfinite(din).all(), prec
        assert eW <= tol_w and eb <= tol_w, (prec, "wgrad", float(eW), float(eb))
        assert ed <= tol_d, (prec, "dgrad", float(ed))


def test_tcgen05_weight_gradient_is_bit_reproducible(engines):
    """Layers with >= 64 input and output channels: split-K partial sums and the bias gradient are reduced in a fixed
    order (no atomics), so identical inputs give identical bits, run after run."""
    _, eng = engines
    b = unet_blocks(**CFG)[9]                      # 128 -> 128, 3x3
    rng = np.random.default_rng(1)
    a_in = rng.normal(0.0, 1.0, size=(8, 32, 16, b.cin)).astype(np.float32)
    dz = rng.normal(0.0, 1.0, size=(8, 32, 16, b.cout)).astype(np.float32)
    first = eng["bf16"].debug_backward_block(9, a_in, dz)
    for _ in range(3):
        again = eng["bf16"].debug_backward_block(9, a_in, dz)
        assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1])
