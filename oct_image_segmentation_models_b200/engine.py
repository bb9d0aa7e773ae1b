"""Thin host wrapper around one liboctseg network handle (one per GPU).

numpy in / numpy out; all arithmetic happens in the CUDA library.
"""
import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _native as nat
from .models.unet_spec import unet_param_specs


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class UNetEngine:
    def __init__(self, *, input_channels: int, num_classes: int, start_neurons: int = 8,
                 pool_layers: int = 4, conv_layers: int = 2, enc_kernel=(3, 3), dec_kernel=(2, 2),
                 precision: str = "bf16", device: int = 0):
        if precision not in ("fp32", "bf16", "fp16"):
            raise ValueError("precision must be 'fp32', 'bf16' or 'fp16'")
        self.spec_kwargs = dict(input_channels=input_channels, num_classes=num_classes,
                                start_neurons=start_neurons, pool_layers=pool_layers,
                                conv_layers=conv_layers, enc_kernel=tuple(enc_kernel),
                                dec_kernel=tuple(dec_kernel))
        self.precision = precision
        self.device = device
        self.input_channels = input_channels
        self.num_classes = num_classes
        self._lib = nat.load()
        self._cfg = nat.make_config(**self.spec_kwargs)
        self.param_specs = unet_param_specs(**self.spec_kwargs)
        self._h = C.c_void_p()
        nat.check(self._lib.octseg_create(C.byref(self._cfg), device,
                                          {"fp32": nat.FP32, "bf16": nat.BF16, "fp16": nat.FP16}[precision],
                                          C.byref(self._h)))

    # ---- lifetime ----------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.octseg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -----------------------------------------------------------------
    def set_weights(self, weights: Sequence[np.ndarray]):
        if len(weights) != len(self.param_specs):
            raise ValueError(f"expected {len(self.param_specs)} weight tensors, got {len(weights)}")
        for i, ((name, shape), w) in enumerate(zip(self.param_specs, weights)):
            a = np.ascontiguousarray(w, dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
            nat.check(self._lib.octseg_set_param(self._h, i, _ptr(a), a.size))

    def get_weights(self) -> List[np.ndarray]:
        out = []
        for i, (name, shape) in enumerate(self.param_specs):
            a = np.empty(shape, dtype=np.float32)
            nat.check(self._lib.octseg_get_param(self._h, i, _ptr(a), a.size))
            out.append(a)
        return out

    # ---- inference ---------------------------------------------------------------
    def predict(self, images: np.ndarray, want_probs: bool = True, want_labels: bool = False,
                probs_out: Optional[np.ndarray] = None, labels_out: Optional[np.ndarray] = None):
        """images: [N,H,W,C] uint8 (raw) or float32 (raw 0..255 values; x/255 is applied
        on the device, reference models/unet.py:87-91).  Returns (probs, labels)."""
        if images.ndim != 4 or images.shape[3] != self.input_channels:
            raise ValueError(f"images must be [N,H,W,{self.input_channels}]")
        if images.dtype == np.uint8:
            dt = nat.U8
        else:
            images = np.asarray(images, dtype=np.float32)
            dt = nat.F32
        images = np.ascontiguousarray(images)
        n, h, w, _ = images.shape
        probs = labels = None
        if want_probs:
            probs = probs_out if probs_out is not None else np.empty((n, h, w, self.num_classes), np.float32)
        if want_labels:
            labels = labels_out if labels_out is not None else np.empty((n, h, w), np.uint8)
        nat.check(self._lib.octseg_predict_host(self._h, _ptr(images), dt, n, h, w, _ptr(probs), _ptr(labels)))
        return probs, labels

    def predict_maps(self, images: np.ndarray, bg_ilm: bool = True, bg_csi: bool = False, transposed: bool = False,
                     labels_out: Optional[np.ndarray] = None, maps_out: Optional[np.ndarray] = None):
        """uint8/float images -> (labels uint8 [N,H,W], boundary maps uint8 [N,K-1,H,W] or [N,K-1,W,H]);
        argmax and the reference's map construction run on the GPU, only 1 + (K-1) bytes per pixel
        come back instead of 4*K."""
        if images.dtype == np.uint8:
            dt = nat.U8
        else:
            images = np.asarray(images, dtype=np.float32)
            dt = nat.F32
        images = np.ascontiguousarray(images)
        n, h, w, _ = images.shape
        shape = (n, self.num_classes - 1, w, h) if transposed else (n, self.num_classes - 1, h, w)
        labels = labels_out if labels_out is not None else np.empty((n, h, w), np.uint8)
        maps = maps_out if maps_out is not None else np.empty(shape, np.uint8)
        if labels.shape != (n, h, w) or maps.shape != shape or labels.dtype != np.uint8 or maps.dtype != np.uint8 \
                or not labels.flags.c_contiguous or not maps.flags.c_contiguous:
            raise ValueError("labels_out / maps_out must be C-contiguous uint8 of shape (n,h,w) / " + str(shape))
        nat.check(self._lib.octseg_predict_maps_host(self._h, _ptr(images), dt, n, h, w, int(bg_ilm), int(bg_csi),
                                                     int(transposed), _ptr(labels), _ptr(maps)))
        return labels, maps

    def predict_maps_submit(self, images: np.ndarray, labels_out: np.ndarray, maps_out: np.ndarray, bg_ilm: bool = True,
                            bg_csi: bool = False, transposed: bool = False) -> int:
        """Asynchronous predict_maps on PINNED, C-contiguous host arrays (uint8 images [N,H,W,C]): enqueues the call and
        returns a ticket for predict_wait().  Submit batch i+1 before waiting for batch i to overlap its upload with the
        forward and the download of batch i; at most two calls in flight; the arrays must stay alive until the wait."""
        if images.dtype != np.uint8 or not images.flags.c_contiguous or not labels_out.flags.c_contiguous or not maps_out.flags.c_contiguous:
            raise ValueError("predict_maps_submit needs C-contiguous uint8 arrays")
        n, h, w, _ = images.shape
        t = C.c_int32()
        nat.check(self._lib.octseg_predict_maps_submit(self._h, _ptr(images), nat.U8, n, h, w, int(bg_ilm), int(bg_csi),
                                                       int(transposed), _ptr(labels_out), _ptr(maps_out), C.byref(t)))
        return int(t.value)

    def predict_wait(self, ticket: int):
        nat.check(self._lib.octseg_predict_wait(self._h, int(ticket)))

    def predict_preprocessed(self, x32: np.ndarray) -> np.ndarray:
        """x32: float32 [N,H,W,C] already divided by 255 on the host (the reference's
        preprocess_input_fn output after Keras' float32 cast)."""
        if x32.ndim != 4 or x32.shape[3] != self.input_channels or x32.dtype != np.float32:
            raise ValueError(f"x must be float32 [N,H,W,{self.input_channels}]")
        x32 = np.ascontiguousarray(x32)
        n, h, w, _ = x32.shape
        probs = np.empty((n, h, w, self.num_classes), np.float32)
        nat.check(self._lib.octseg_predict_host(self._h, _ptr(x32), nat.F32_PRE, n, h, w, _ptr(probs), None))
        return probs

    def predict_device(self, images_ptr: int, dtype: int, n: int, h: int, w: int,
                       probs_ptr: Optional[int], labels_ptr: Optional[int], stream: Optional[int] = None):
        """Asynchronous variant on device pointers (e.g. torch tensors' data_ptr())."""
        nat.check(self._lib.octseg_predict_device(self._h, C.c_void_p(images_ptr), dtype, n, h, w,
                                                  C.c_void_p(probs_ptr) if probs_ptr else None,
                                                  C.c_void_p(labels_ptr) if labels_ptr else None,
                                                  C.c_void_p(stream) if stream else None))

    def evaluate_counts(self, images: np.ndarray, labels: np.ndarray, class_weights=None, preprocessed: bool = False):
        """Validation pass on the device (f-4): returns (counts int64 [N,K,3] = per image and class
        (|y==c & p_c>0.5|, |p_c>0.5|, |y==c|), loss_sums float64 [N] of the weighted CE over each image's pixels)."""
        if images.dtype == np.uint8:
            dt = nat.U8
        else:
            images = np.asarray(images, dtype=np.float32)
            dt = nat.F32_PRE if preprocessed else nat.F32
        images = np.ascontiguousarray(images)
        n, h, w, _ = images.shape
        lab = np.ascontiguousarray(np.asarray(labels).reshape(n, h, w), dtype=np.uint8)
        cw = None if class_weights is None else np.ascontiguousarray(class_weights, dtype=np.float32)
        counts = np.zeros((n, self.num_classes, 3), np.int64)
        loss = np.zeros((n,), np.float64)
        nat.check(self._lib.octseg_evaluate_host(self._h, _ptr(images), dt, _ptr(lab), n, h, w, _ptr(cw), _ptr(counts),
                                                 _ptr(loss)))
        return counts, loss

    def synchronize(self):
        nat.check(self._lib.octseg_synchronize(self._h))

    # ---- training ----------------------------------------------------------------
    def train_begin(self, class_weights: Sequence[float], learning_rate=1e-3, beta_1=0.9, beta_2=0.999,
                    epsilon=1e-7, dropout_rate=0.5, dropout_seed=0, global_batch=0):
        cw = np.ascontiguousarray(class_weights, dtype=np.float32)
        if cw.size != self.num_classes:
            raise ValueError("class_weights must have num_classes entries")
        tc = nat.OctsegTrainConfig(learning_rate, beta_1, beta_2, epsilon, dropout_rate, dropout_seed,
                                   global_batch)
        nat.check(self._lib.octseg_train_begin(self._h, C.byref(tc), _ptr(cw)))

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        nat.check(self._lib.octseg_comm_init(self._h, buf, rank, world))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        nat.check(nat.load().octseg_comm_unique_id(buf))
        return bytes(buf)

    def train_step(self, images: np.ndarray, labels: np.ndarray,
                   dropout_mask: Optional[np.ndarray] = None, preprocessed: bool = False) -> float:
        """preprocessed=True: float images already divided by 255 (the batches a reference DataGenerator yields)."""
        if images.dtype == np.uint8:
            dt = nat.U8
        else:
            images = np.asarray(images, dtype=np.float32)
            dt = nat.F32_PRE if preprocessed else nat.F32
        images = np.ascontiguousarray(images)
        n, h, w, _ = images.shape
        lab = np.ascontiguousarray(labels.reshape(n, h, w), dtype=np.uint8)
        dm = None if dropout_mask is None else np.ascontiguousarray(dropout_mask, dtype=np.uint8)
        loss = C.c_float()
        nat.check(self._lib.octseg_train_step_host(self._h, _ptr(images), dt, _ptr(lab), n, h, w, _ptr(dm),
                                                   C.byref(loss)))
        return float(loss.value)

    def train_step_device(self, images_ptr: int, dtype: int, labels_ptr: int, n: int, h: int, w: int,
                          loss_ptr: Optional[int] = None, stream: Optional[int] = None,
                          mask_ptr: Optional[int] = None):
        """Asynchronous train step on device pointers (this rank's shard)."""
        nat.check(self._lib.octseg_train_step_device(
            self._h, C.c_void_p(images_ptr), dtype, C.c_void_p(labels_ptr), n, h, w,
            C.c_void_p(mask_ptr) if mask_ptr else None, C.c_void_p(loss_ptr) if loss_ptr else None,
            C.c_void_p(stream) if stream else None))

    def get_optimizer_state(self):
        """(iterations, [m per parameter], [v per parameter]); entries of non-trainable tensors are zeros."""
        it = C.c_int64()
        nat.check(self._lib.octseg_opt_iterations(self._h, 0, C.byref(it)))
        ms, vs = [], []
        for which, dst in ((0, ms), (1, vs)):
            for i, (_, shape) in enumerate(self.param_specs):
                a = np.empty(shape, dtype=np.float32)
                nat.check(self._lib.octseg_opt_state(self._h, 0, which, i, _ptr(a), a.size))
                dst.append(a)
        return int(it.value), ms, vs

    def set_optimizer_state(self, iterations: int, ms: Sequence[np.ndarray], vs: Sequence[np.ndarray]):
        it = C.c_int64(int(iterations))
        nat.check(self._lib.octseg_opt_iterations(self._h, 1, C.byref(it)))
        for which, src in ((0, ms), (1, vs)):
            for i, ((name, shape), a) in enumerate(zip(self.param_specs, src)):
                a = np.ascontiguousarray(a, dtype=np.float32)
                if tuple(a.shape) != tuple(shape):
                    raise ValueError(f"optimizer slot of {name}: expected shape {shape}, got {a.shape}")
                nat.check(self._lib.octseg_opt_state(self._h, 1, which, i, _ptr(a), a.size))

    def get_grads(self) -> List[Optional[np.ndarray]]:
        out: List[Optional[np.ndarray]] = []
        for i, (name, shape) in enumerate(self.param_specs):
            if name.endswith("moving_mean:0") or name.endswith("moving_variance:0"):
                out.append(None)
                continue
            a = np.empty(shape, dtype=np.float32)
            nat.check(self._lib.octseg_get_grad(self._h, i, _ptr(a), a.size))
            out.append(a)
        return out

    # ---- introspection -----------------------------------------------------------
    def launch_count(self) -> int:
        return int(self._lib.octseg_launch_count(self._h))

    def set_profiling(self, enable: bool):
        nat.check(self._lib.octseg_set_profiling(self._h, 1 if enable else 0))

    def block_times_ms(self) -> List[float]:
        buf = (C.c_float * 256)()
        n = C.c_int32()
        nat.check(self._lib.octseg_get_block_times(self._h, buf, 256, C.byref(n)))
        return [float(buf[i]) for i in range(n.value)]

    def layer_uses_tensor_core(self, conv_index: int, h: int, w: int) -> bool:
        return bool(self._lib.octseg_layer_uses_tensor_core(self._h, conv_index, h, w))

    def debug_backward_block(self, conv_index: int, a_in_nhwc: np.ndarray, dz_nhwc: np.ndarray):
        """Weight gradient, bias gradient and data gradient of ONE conv block through the train step's own kernels
        (needs train_begin): returns (dW [kh,kw,cin,cout], db [cout], d_in NHWC)."""
        from .models.unet_spec import unet_blocks
        b = unet_blocks(**self.spec_kwargs)[conv_index]
        a = np.ascontiguousarray(a_in_nhwc, dtype=np.float32)
        dz = np.ascontiguousarray(dz_nhwc, dtype=np.float32)
        n, h, w, c = a.shape
        oh, ow = (2 * h, 2 * w) if b.upsample_before else (h, w)
        assert c == b.cin and dz.shape == (n, oh, ow, b.cout)
        dW = np.empty((b.kh, b.kw, b.cin, b.cout), np.float32)
        db = np.empty((b.cout,), np.float32)
        din = np.empty((n, h, w, b.cin), np.float32)
        nat.check(self._lib.octseg_debug_backward_block(self._h, conv_index, _ptr(a), _ptr(dz), n, h, w, _ptr(dW), _ptr(db),
                                                        _ptr(din)))
        return dW, db, din

    def debug_conv_block(self, conv_index: int, x_nhwc: np.ndarray, path: int = 0, timed: bool = False):
        """Run one conv block (conv + folded BN + ReLU; x2 upsample first for 'up' blocks)
        on NHWC float32 input.  path 0 = CUDA-core kernel, 1 = tcgen05 kernel."""
        from .models.unet_spec import unet_blocks
        b = unet_blocks(**self.spec_kwargs)[conv_index]
        x = np.ascontiguousarray(x_nhwc, dtype=np.float32)
        n, h, w, c = x.shape
        assert c == b.cin
        oh, ow = (2 * h, 2 * w) if b.upsample_before else (h, w)
        out = np.empty((n, oh, ow, b.cout), np.float32)
        ms = C.c_float(0.0)
        nat.check(self._lib.octseg_debug_conv_block(self._h, conv_index, path, _ptr(x), n, h, w, _ptr(out),
                                                    C.byref(ms) if timed else None))
        return (out, float(ms.value)) if timed else out
