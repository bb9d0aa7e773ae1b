"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path
(only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may use it).

PARITY UNPINNED for the network arithmetic: the reference executes this path
inside tensorflow==2.9.0 (pyproject.toml:31), which is not installable here, and
the reference ships no tests, fixtures or golden vectors.  This file restates,
on torch-CPU (oneDNN, the same kernel-library family TF uses on CPU):

  * the graph built by UNet.build_model     (reference models/unet.py:106-153,
    blocks :20-57) with Keras-2.9 layer semantics:
    Conv2D = NHWC cross-correlation with HWIO kernels, "same" padding
    (k=3: 1/1, k=2: 0 before / 1 after), BatchNormalization eps=1e-3
    momentum=0.99, MaxPooling2D 2x2 valid, UpSampling2D nearest x2,
    concatenate([up, skip]), Dropout(0.5), softmax over the last axis;
  * x/255 preprocessing                     (reference models/unet.py:87-91);
  * weighted categorical cross-entropy      (reference common/custom_losses.py:27-35)
    with the Keras SUM_OVER_BATCH_SIZE reduction (mean over all B*H*W pixels);
  * Keras optimizer_v2 Adam (eps outside the bias correction).

What *is* pinned: the min-path comparator (oracle/min_path.py) is checked against
the unmodified reference file in tests/test_oracle_golden.py (goldens made by tests/golden/make_minpath_golden.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # Keras BatchNormalization() default epsilon
BN_MOMENTUM = 0.99     # Keras default momentum
K_EPSILON = 1e-7       # K.epsilon()


# ----------------------------------------------------------------------------
# graph structure (independent restatement of reference unet.py:106-153)
# ----------------------------------------------------------------------------
def oracle_blocks(input_channels, num_classes, start_neurons=8, pool_layers=4,
                  conv_layers=2, enc_kernel=(3, 3), dec_kernel=(2, 2)) -> List[dict]:
    blocks = []
    cin = input_channels
    for i in range(pool_layers):                       # unet.py:113-121
        f = start_neurons * 2 ** i
        for j in range(conv_layers):                   # unet.py:33-34
            blocks.append(dict(kind="enc", level=i, k=tuple(enc_kernel), cin=cin, cout=f,
                               pool=(j == conv_layers - 1)))
            cin = f
    f = start_neurons * 2 ** pool_layers               # unet.py:123-129
    for j in range(conv_layers):
        blocks.append(dict(kind="mid", level=pool_layers, k=tuple(enc_kernel), cin=cin, cout=f,
                           dropout=(j == conv_layers - 1)))
        cin = f
    for i in range(pool_layers):                       # unet.py:132-140
        lvl = pool_layers - 1 - i
        f = start_neurons * 2 ** lvl
        blocks.append(dict(kind="up", level=lvl, k=tuple(dec_kernel), cin=cin, cout=f))
        cin = 2 * f
        for j in range(conv_layers):
            blocks.append(dict(kind="dec", level=lvl, k=tuple(enc_kernel), cin=cin, cout=f,
                               concat=(j == 0)))
            cin = f
    blocks.append(dict(kind="head", level=0, k=(1, 1), cin=cin, cout=num_classes))
    return blocks


def oracle_param_shapes(**cfg) -> List[Tuple[int, ...]]:
    shapes = []
    for b in oracle_blocks(**cfg):
        shapes.append((b["k"][0], b["k"][1], b["cin"], b["cout"]))
        shapes.append((b["cout"],))
        if b["kind"] != "head":
            shapes += [(b["cout"],)] * 4
    return shapes


def _conv_same(x: torch.Tensor, w_hwio: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """Keras Conv2D(strides 1, padding 'same') on NCHW x with an HWIO kernel."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    pb, pr = kh - 1 - pt, kw - 1 - pl
    if pt or pb or pl or pr:
        x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, w_hwio.permute(3, 2, 0, 1).contiguous(), bias)


def preprocess(images: np.ndarray, dtype=torch.float32) -> torch.Tensor:
    """reference unet.py:87-91 (x / 255.0 in float64) then Keras' cast to float32.
    images: [N,H,W,C] uint8 or float -> NCHW tensor."""
    x = np.asarray(images).astype(np.float64) / 255.0
    t = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))
    return t.to(torch.float32).to(dtype)


class OracleUNet:
    """Functional torch-CPU restatement; weights are a flat Keras-order list."""

    def __init__(self, weights: Sequence[np.ndarray], dtype=torch.float32, **cfg):
        self.cfg = dict(cfg)
        self.blocks = oracle_blocks(**cfg)
        self.dtype = dtype
        shapes = oracle_param_shapes(**cfg)
        assert len(shapes) == len(weights), (len(shapes), len(weights))
        for s, w in zip(shapes, weights):
            assert tuple(w.shape) == tuple(s), (w.shape, s)
        self.params: List[torch.Tensor] = [
            torch.tensor(np.asarray(w), dtype=dtype) for w in weights]

    # -- helpers -------------------------------------------------------------
    def _block_params(self):
        i = 0
        for b in self.blocks:
            n = 2 if b["kind"] == "head" else 6
            yield b, self.params[i:i + n], i
            i += n

    def trainable_mask(self) -> List[bool]:
        m = []
        for b in self.blocks:
            m += [True, True] if b["kind"] == "head" else [True, True, True, True, False, False]
        return m

    # -- forward -------------------------------------------------------------
    def forward(self, x: torch.Tensor, training: bool = False,
                dropout_mask: Optional[torch.Tensor] = None,
                params: Optional[List[torch.Tensor]] = None,
                batch_stats_out: Optional[list] = None,
                return_logits: bool = False) -> torch.Tensor:
        """x: NCHW, already preprocessed.  Returns NHWC probabilities.
        training=True uses per-batch BN statistics (biased variance) and applies
        `dropout_mask` ([N,C,h,w] of {0,1}; kept values are scaled by 2) after
        the bottleneck; a None mask means rate 0."""
        P = params if params is not None else self.params
        skips: Dict[int, torch.Tensor] = {}
        i = 0
        for b in self.blocks:
            if b["kind"] == "head":
                w, bias = P[i], P[i + 1]
                logits = _conv_same(x, w, bias)
                logits = logits.permute(0, 2, 3, 1)
                if return_logits:
                    return logits
                return torch.softmax(logits, dim=-1)
            w, bias, gamma, beta, mmean, mvar = P[i:i + 6]
            i += 6
            if b["kind"] == "up":
                x = F.interpolate(x, scale_factor=2, mode="nearest")    # UpSampling2D()
            if b.get("concat"):
                x = torch.cat([x, skips[b["level"]]], dim=1)            # unet.py:52
            z = _conv_same(x, w, bias)
            if training:
                mean = z.mean(dim=(0, 2, 3))
                var = z.var(dim=(0, 2, 3), unbiased=False)
                if batch_stats_out is not None:
                    n = z.numel() // z.shape[1]
                    batch_stats_out.append((mean.detach(), var.detach(), n))
            else:
                mean, var = mmean, mvar
            inv = torch.rsqrt(var + BN_EPS)
            y = (z - mean[None, :, None, None]) * (inv * gamma)[None, :, None, None] \
                + beta[None, :, None, None]
            x = torch.relu(y)
            if b.get("pool"):
                skips[b["level"]] = x
                x = F.max_pool2d(x, 2)
            if b.get("dropout") and training and dropout_mask is not None:
                x = x * dropout_mask.to(x.dtype) * 2.0
        raise AssertionError("no head")

    @torch.no_grad()
    def predict(self, images: np.ndarray) -> np.ndarray:
        """[N,H,W,C] u8/float images -> float32 [N,H,W,K] probabilities (inference BN)."""
        x = preprocess(images, self.dtype)
        return self.forward(x).to(torch.float32).numpy()

    # -- training ------------------------------------------------------------
    def loss_and_grads(self, images: np.ndarray, labels: np.ndarray,
                       class_weights: Sequence[float],
                       dropout_mask: Optional[np.ndarray] = None,
                       loss_scale_pixels: Optional[int] = None):
        """One forward/backward in training mode.
        labels: [N,H,W] or [N,H,W,1] integer class ids.
        Returns (loss, grads list aligned with params (None for BN moving stats),
        batch_stats list, probs)."""
        K = self.cfg["num_classes"]
        x = preprocess(images, self.dtype)
        lab = torch.from_numpy(np.asarray(labels).reshape(labels.shape[0], labels.shape[1],
                                                         labels.shape[2]).astype(np.int64))
        params = [p.clone().requires_grad_(t) for p, t in zip(self.params, self.trainable_mask())]
        stats: list = []
        dm = None if dropout_mask is None else torch.from_numpy(
            np.ascontiguousarray(dropout_mask.transpose(0, 3, 1, 2)))
        probs = self.forward(x, training=True, dropout_mask=dm, params=params,
                             batch_stats_out=stats)
        per_pixel = weighted_cce(probs, F.one_hot(lab, K).to(probs.dtype),
                                 torch.tensor(class_weights, dtype=probs.dtype))
        denom = per_pixel.numel() if loss_scale_pixels is None else loss_scale_pixels
        loss = per_pixel.sum() / denom
        loss.backward()
        grads = [p.grad.detach().clone() if p.requires_grad else None for p in params]
        return float(loss.detach()), grads, stats, probs.detach()

    def apply_bn_moving_update(self, stats):
        """Keras fused BN: moving <- moving*0.99 + batch*0.01, with the
        Bessel-corrected batch variance going into moving_variance."""
        si = 0
        for b, ps, base in self._block_params():
            if b["kind"] == "head":
                continue
            mean, var, n = stats[si]
            si += 1
            unbiased = var * (n / max(n - 1, 1))
            self.params[base + 4] = self.params[base + 4] * BN_MOMENTUM + mean * (1 - BN_MOMENTUM)
            self.params[base + 5] = self.params[base + 5] * BN_MOMENTUM + unbiased * (1 - BN_MOMENTUM)

    def get_weights(self) -> List[np.ndarray]:
        return [p.detach().to(torch.float32).numpy().copy() for p in self.params]


def weighted_cce(y_pred: torch.Tensor, y_true: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """reference common/custom_losses.py:27-35, line for line in meaning:
    renormalise, clip to [eps, 1-eps], -sum(y_true*log(y_pred)*w) over classes."""
    y_pred = y_pred / y_pred.sum(dim=-1, keepdim=True)
    y_pred = torch.clamp(y_pred, K_EPSILON, 1 - K_EPSILON)
    loss = y_true * torch.log(y_pred) * weights
    return -loss.sum(dim=-1)


class KerasAdam:
    """tf.keras.optimizers.Adam (optimizer_v2) update rule:
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t);  theta -= lr_t * m / (sqrt(v) + eps)."""

    def __init__(self, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.t = 0
        self.m: Dict[int, torch.Tensor] = {}
        self.v: Dict[int, torch.Tensor] = {}

    def step(self, params: List[torch.Tensor], grads: List[Optional[torch.Tensor]]):
        self.t += 1
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for i, (p, g) in enumerate(zip(params, grads)):
            if g is None:
                continue
            m = self.m.get(i, torch.zeros_like(p))
            v = self.v.get(i, torch.zeros_like(p))
            m = self.b1 * m + (1 - self.b1) * g
            v = self.b2 * v + (1 - self.b2) * g * g
            self.m[i], self.v[i] = m, v
            params[i] = p - lr_t * m / (torch.sqrt(v) + self.eps)
