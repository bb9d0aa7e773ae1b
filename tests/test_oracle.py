"""Oracle self-checks (CPU): Keras/TF layer semantics the restatement must encode
(SURVEY.md App. B), checked against hand-computed tiny tensors."""
import numpy as np
import torch

from oracle import unet_oracle as uo
from oct_image_segmentation_models_b200.common.synthetic import synthetic_weights, synthetic_batch
from oct_image_segmentation_models_b200.models.unet_spec import unet_param_specs

CFG = dict(input_channels=1, num_classes=4)


def test_param_inventory_matches_survey():
    specs = unet_param_specs(**CFG)
    assert len(specs) == 134
    assert sum(int(np.prod(s)) for _, s in specs) == 489_124
    trainable = sum(int(np.prod(s)) for n, s in specs if "moving" not in n)
    assert trainable == 487_412
    assert [s for _, s in specs] == uo.oracle_param_shapes(**CFG)
    wide = unet_param_specs(input_channels=1, num_classes=4, start_neurons=64)
    assert sum(int(np.prod(s)) for _, s in wide) == 31_058_180


def test_same_padding_k2_pads_bottom_right_only():
    x = torch.arange(9, dtype=torch.float32).reshape(1, 1, 3, 3)
    w = torch.ones(2, 2, 1, 1)
    y = uo._conv_same(x, w, torch.zeros(1))[0, 0]
    # out[i,j] = x[i,j]+x[i,j+1]+x[i+1,j]+x[i+1,j+1] with zeros past the bottom/right edge
    assert y.shape == (3, 3)
    assert y[0, 0] == 0 + 1 + 3 + 4
    assert y[2, 2] == 8            # only x[2,2] is inside
    assert y[0, 2] == 2 + 5
    assert y[2, 0] == 6 + 7


def test_same_padding_k3_is_symmetric_and_cross_correlation():
    x = torch.zeros(1, 1, 3, 3)
    x[0, 0, 1, 1] = 1.0
    w = torch.arange(9, dtype=torch.float32).reshape(3, 3, 1, 1)
    y = uo._conv_same(x, w, torch.zeros(1))[0, 0]
    # cross-correlation (no flip): an impulse at the centre returns the kernel flipped
    assert torch.equal(y, torch.flip(w[:, :, 0, 0], dims=(0, 1)))


def test_hwio_channel_mapping():
    x = torch.zeros(1, 2, 1, 1)
    x[0, 1] = 1.0                                 # only input channel 1 active
    w = torch.zeros(1, 1, 2, 3)
    w[0, 0, 1, 2] = 5.0                           # in 1 -> out 2
    y = uo._conv_same(x, w, torch.zeros(3))
    assert y[0, :, 0, 0].tolist() == [0.0, 0.0, 5.0]


def test_concat_order_is_up_then_skip_and_bn_eps():
    cfg = dict(input_channels=1, num_classes=2, start_neurons=8, pool_layers=1, conv_layers=1)
    w = synthetic_weights(seed=1, **cfg)
    net = uo.OracleUNet(w, **cfg)
    kinds = [b["kind"] for b in net.blocks]
    assert kinds == ["enc", "mid", "up", "dec", "head"]
    assert net.blocks[3]["cin"] == 16 and net.blocks[3]["concat"]
    imgs, _ = synthetic_batch(0, 1, 8, 8, 2)
    p = net.predict(imgs)
    assert p.shape == (1, 8, 8, 2)
    np.testing.assert_allclose(p.sum(-1), 1.0, atol=1e-6)
    assert uo.BN_EPS == 1e-3 and uo.BN_MOMENTUM == 0.99


def test_preprocess_is_float64_division_then_float32():
    img = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1)
    x = uo.preprocess(img).numpy().reshape(-1)
    expect = (np.arange(256, dtype=np.float64) / 255.0).astype(np.float32)
    assert np.array_equal(x, expect)


def test_weighted_cce_gradient_formula():
    """d loss/d z_j = w_t (p_j/S - 1[j=t]) with zero gradient where the clip is active
    (SURVEY App. B), against autograd in float64."""
    torch.manual_seed(0)
    z = (torch.randn(5, 7, 4, dtype=torch.float64) * 6).requires_grad_(True)
    t = torch.randint(0, 4, (5, 7))
    wts = torch.tensor([0.5, 1.0, 2.0, 1.0], dtype=torch.float64)
    p = torch.softmax(z, -1)
    loss = uo.weighted_cce(p, torch.nn.functional.one_hot(t, 4).double(), wts).sum()
    loss.backward()
    with torch.no_grad():
        pt = p.gather(-1, t[..., None])[..., 0]
        active = (pt >= uo.K_EPSILON) & (pt <= 1 - uo.K_EPSILON)
        g = wts[t][..., None] * (p - torch.nn.functional.one_hot(t, 4).double())
        g = g * active[..., None]
    assert torch.allclose(z.grad, g, atol=1e-12)


def test_keras_adam_epsilon_outside_bias_correction():
    p = [torch.tensor([1.0, -2.0])]
    g = [torch.tensor([0.5, 0.25])]
    opt = uo.KerasAdam(lr=1e-3)
    opt.step(p, g)
    m = 0.1 * g[0]
    v = 0.001 * g[0] ** 2
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    expect = torch.tensor([1.0, -2.0]) - lr_t * m / (torch.sqrt(v) + 1e-7)
    assert torch.allclose(p[0], expect, atol=1e-9)


def test_training_bn_uses_batch_stats_and_unbiased_moving_var():
    cfg = dict(input_channels=1, num_classes=3, start_neurons=8, pool_layers=1, conv_layers=1)
    w = synthetic_weights(seed=3, random_bn_stats=False, **cfg)
    net = uo.OracleUNet(w, **cfg)
    imgs, labs = synthetic_batch(0, 2, 8, 8, 3)
    loss, grads, stats, _ = net.loss_and_grads(imgs, labs, [1.0, 1.0, 1.0])
    assert np.isfinite(loss)
    assert sum(g is not None for g in grads) == 4 * 4 + 2
    mean0, var0, n0 = stats[0]
    assert n0 == 2 * 8 * 8
    before = net.params[5].clone()
    net.apply_bn_moving_update(stats)
    expect = before * 0.99 + var0 * (n0 / (n0 - 1)) * 0.01
    assert torch.allclose(net.params[5], expect)
