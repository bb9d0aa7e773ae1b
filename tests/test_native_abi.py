"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the
header declares, agrees with the Python graph description, and refuses to compute without a
GPU (no CPU fallback)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oct_image_segmentation_models_b200 import _native as nat
from oct_image_segmentation_models_b200.models.unet_spec import unet_param_specs

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "octseg.h"


def header_functions():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(octseg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/octseg.h but not exported"
    assert set(names) == set(nat.EXPORTED_SYMBOLS), "ctypes prototypes drifted from the header"


def test_version_and_error_string():
    lib = nat.load()
    assert lib.octseg_version() >= 100
    assert isinstance(lib.octseg_last_error(), bytes)


@pytest.mark.parametrize("kw", [dict(), dict(start_neurons=64), dict(pool_layers=2, conv_layers=3),
                                dict(num_classes=7, input_channels=3)])
def test_native_graph_matches_python_spec(kw):
    args = dict(input_channels=1, num_classes=4)
    args.update(kw)
    native = nat.native_param_specs(nat.make_config(**args))
    py = unet_param_specs(**args)
    assert [(n, s) for n, s, _ in native] == py
    assert [t for _, _, t in native] == ["moving" not in n for n, _ in py]


def test_bad_config_is_an_error_not_a_crash():
    cfg = nat.make_config(1, 4, start_neurons=0)
    n = C.c_int32()
    assert nat.load().octseg_param_count(C.byref(cfg), C.byref(n)) != 0
    assert b"invalid" in nat.load().octseg_last_error()


def test_no_cpu_fallback():
    lib = nat.load()
    if lib.octseg_device_count() > 0:
        pytest.skip("GPU present")
    from oct_image_segmentation_models_b200.engine import UNetEngine
    with pytest.raises(nat.NativeError, match="no CPU fallback"):
        UNetEngine(input_channels=1, num_classes=4)


def test_product_never_imports_oracle():
    pkg = ROOT / "oct_image_segmentation_models_b200"
    for f in pkg.rglob("*.py"):
        txt = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
    for f in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert "oracle/" not in f.read_text(), f


def test_tc_descriptor_arithmetic_host_emulation():
    """Runs the host emulation of the tcgen05 kernel's smem-descriptor arithmetic and weight
    packing (csrc/hosttest/tc_emulate.cu) -- every geometry must reproduce a direct conv."""
    csrc = ROOT / "oct_image_segmentation_models_b200" / "csrc"
    exe = csrc / "hosttest" / "tc_emulate"
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart",
                    "static", "hosttest/tc_emulate.cu", "conv_tc.cu", "-o", str(exe)], cwd=csrc, check=True,
                   capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "ALL OK" in out.stdout
