"""Static description of the U-Net graph: layers, shapes and the Keras weight order.

The reference builds this graph with Keras layer objects
(reference: oct_image_segmentation_models/models/unet.py:20-57, 106-153).  This
module restates only its *structure* (no arithmetic) so that host code can name,
size and order the weight tensors exactly as `model.get_weights()` /
`model.layers` would:

  for each conv block   : kernel[kh,kw,Cin,Cout], bias[Cout],
                          gamma[Cout], beta[Cout], moving_mean[Cout], moving_variance[Cout]
  head 1x1 conv         : kernel[1,1,s,K], bias[K]

Block order = graph creation order: encoder level 0..P-1 (L blocks each),
bottleneck (L blocks), then per decoder level: up-conv block (dec_kernel on the
x2 nearest-upsampled tensor), L blocks on concat([up, skip]).
"""
from dataclasses import dataclass
from typing import List, Tuple


@dataclass(frozen=True)
class ConvBlockSpec:
    index: int          # position among conv layers (Keras auto-name suffix)
    role: str           # "enc", "mid", "up", "dec", "head"
    level: int          # resolution level: 0 = full res, P = bottleneck
    kh: int
    kw: int
    cin: int
    cout: int
    has_bn: bool        # False only for the head
    pool_after: bool    # 2x2 max-pool follows (last block of an encoder level)
    upsample_before: bool  # nearest x2 precedes (up-conv block)
    concat_skip_level: int  # >=0: input is concat([prev, skip(level)]) else -1
    dropout_after: bool  # Dropout(0.5) follows (last bottleneck block)


def unet_blocks(
    input_channels: int,
    num_classes: int,
    start_neurons: int = 8,
    pool_layers: int = 4,
    conv_layers: int = 2,
    enc_kernel: Tuple[int, int] = (3, 3),
    dec_kernel: Tuple[int, int] = (2, 2),
) -> List[ConvBlockSpec]:
    """Conv blocks in Keras creation order (reference unet.py:113-147)."""
    blocks: List[ConvBlockSpec] = []
    idx = 0
    cin = input_channels
    P, L, s = pool_layers, conv_layers, start_neurons
    for i in range(P):
        f = s * (2 ** i)
        for j in range(L):
            blocks.append(ConvBlockSpec(idx, "enc", i, enc_kernel[0], enc_kernel[1], cin, f,
                                        True, j == L - 1, False, -1, False))
            idx += 1
            cin = f
    f = s * (2 ** P)
    for j in range(L):
        blocks.append(ConvBlockSpec(idx, "mid", P, enc_kernel[0], enc_kernel[1], cin, f,
                                    True, False, False, -1, j == L - 1))
        idx += 1
        cin = f
    for i in range(P):
        lvl = P - 1 - i
        f = s * (2 ** lvl)
        blocks.append(ConvBlockSpec(idx, "up", lvl, dec_kernel[0], dec_kernel[1], cin, f,
                                    True, False, True, -1, False))
        idx += 1
        cin = 2 * f  # concatenate([up, skip]) -- reference unet.py:52
        for j in range(L):
            blocks.append(ConvBlockSpec(idx, "dec", lvl, enc_kernel[0], enc_kernel[1], cin, f,
                                        True, False, False, lvl if j == 0 else -1, False))
            idx += 1
            cin = f
    blocks.append(ConvBlockSpec(idx, "head", 0, 1, 1, cin, num_classes,
                                False, False, False, -1, False))
    return blocks


def _keras_suffix(i: int) -> str:
    return "" if i == 0 else f"_{i}"


def unet_param_specs(**cfg) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every weight tensor in Keras `get_weights()` order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    for b in unet_blocks(**cfg):
        cn = "conv2d" + _keras_suffix(b.index)
        out.append((f"{cn}/kernel:0", (b.kh, b.kw, b.cin, b.cout)))
        out.append((f"{cn}/bias:0", (b.cout,)))
        if b.has_bn:
            bn = "batch_normalization" + _keras_suffix(b.index)
            for w in ("gamma", "beta", "moving_mean", "moving_variance"):
                out.append((f"{bn}/{w}:0", (b.cout,)))
    return out


def config_to_spec_kwargs(model_config: dict) -> dict:
    """Pick the graph-shaping keys out of a `model_config.json` dict
    (schema: reference unet.py:93-104 + base_model.py:27-33)."""
    return dict(
        input_channels=int(model_config["input_channels"]),
        num_classes=int(model_config["num_classes"]),
        start_neurons=int(model_config.get("start_neurons", 8)),
        pool_layers=int(model_config.get("pool_layers", 4)),
        conv_layers=int(model_config.get("conv_layers", 2)),
        enc_kernel=tuple(model_config.get("enc_kernel", (3, 3))),
        dec_kernel=tuple(model_config.get("dec_kernel", (2, 2))),
    )
