"""ctypes binding of liboctseg.so (include/octseg.h).

There is no CPU fallback: if the library is missing, or a compute entry point is
called without a CUDA device, a NativeError is raised.
"""
import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "csrc" / "liboctseg.so"


class NativeError(RuntimeError):
    pass


class OctsegConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "input_channels", "num_classes", "start_neurons", "pool_layers", "conv_layers",
        "enc_kh", "enc_kw", "dec_kh", "dec_kw")]


class OctsegTrainConfig(C.Structure):
    _fields_ = [("learning_rate", C.c_float), ("beta_1", C.c_float), ("beta_2", C.c_float),
                ("epsilon", C.c_float), ("dropout_rate", C.c_float), ("dropout_seed", C.c_uint64),
                ("global_batch", C.c_int32)]


FP32, BF16, FP16 = 0, 1, 2
U8, F32, F32_PRE = 0, 1, 2

# name -> (restype, argtypes); mirrors include/octseg.h one to one
_PROTOS = {
    "octseg_version": (C.c_int32, []),
    "octseg_last_error": (C.c_char_p, []),
    "octseg_device_count": (C.c_int32, []),
    "octseg_param_count": (C.c_int32, [C.POINTER(OctsegConfig), C.POINTER(C.c_int32)]),
    "octseg_param_info": (C.c_int32, [C.POINTER(OctsegConfig), C.c_int32, C.c_char_p, C.c_int32,
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "octseg_create": (C.c_int32, [C.POINTER(OctsegConfig), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "octseg_destroy": (C.c_int32, [C.c_void_p]),
    "octseg_set_param": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
    "octseg_get_param": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
    "octseg_predict_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p]),
    "octseg_predict_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "octseg_predict_maps_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "octseg_predict_maps_submit": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "octseg_predict_wait": (C.c_int32, [C.c_void_p, C.c_int32]),
    "octseg_evaluate_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "octseg_synchronize": (C.c_int32, [C.c_void_p]),
    "octseg_train_begin": (C.c_int32, [C.c_void_p, C.POINTER(OctsegTrainConfig), C.c_void_p]),
    "octseg_comm_unique_id": (C.c_int32, [C.c_void_p]),
    "octseg_comm_init": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "octseg_train_step_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_void_p, C.POINTER(C.c_float)]),
    "octseg_train_step_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "octseg_opt_state": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "octseg_opt_iterations": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "octseg_get_grad": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
    "octseg_min_path_segment": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "octseg_launch_count": (C.c_int64, [C.c_void_p]),
    "octseg_set_profiling": (C.c_int32, [C.c_void_p, C.c_int32]),
    "octseg_get_block_times": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32)]),
    "octseg_layer_uses_tensor_core": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "octseg_debug_backward_block": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "octseg_debug_conv_block": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_void_p, C.POINTER(C.c_float)]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)
_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("OCTSEG_LIB", str(_LIB_PATH)))


def load():
    """Load liboctseg.so (once).  Raises NativeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise NativeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C oct_image_segmentation_models_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(str(p))
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)          # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().octseg_last_error()
        raise NativeError(msg.decode() if msg else f"liboctseg call failed ({rc})")


def make_config(input_channels, num_classes, start_neurons=8, pool_layers=4, conv_layers=2,
                enc_kernel=(3, 3), dec_kernel=(2, 2)) -> OctsegConfig:
    return OctsegConfig(int(input_channels), int(num_classes), int(start_neurons), int(pool_layers),
                        int(conv_layers), int(enc_kernel[0]), int(enc_kernel[1]), int(dec_kernel[0]),
                        int(dec_kernel[1]))


def native_param_specs(cfg: OctsegConfig):
    """[(name, shape, trainable)] straight from the library (no GPU needed)."""
    lib = load()
    n = C.c_int32()
    check(lib.octseg_param_count(C.byref(cfg), C.byref(n)))
    out = []
    for i in range(n.value):
        name = C.create_string_buffer(96)
        nd, tr = C.c_int32(), C.c_int32()
        shape = (C.c_int64 * 4)()
        check(lib.octseg_param_info(C.byref(cfg), i, name, 96, C.byref(nd), shape, C.byref(tr)))
        out.append((name.value.decode(), tuple(int(shape[k]) for k in range(nd.value)), bool(tr.value)))
    return out
