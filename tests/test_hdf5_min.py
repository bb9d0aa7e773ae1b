"""Minimal HDF5 reader/writer (f-1): round trips, Keras weight layout, and structural checks against the
public HDF5 file-format specification (no h5py exists offline -> parity with real Keras files is unpinned)."""
import struct

import numpy as np
import pytest

from oct_image_segmentation_models_b200.common import hdf5_min as h5


def test_roundtrip_datasets_groups_attrs(tmp_path):
    rng = np.random.default_rng(0)
    data = {
        "a/b/c/kernel:0": rng.normal(size=(3, 3, 8, 16)).astype(np.float32),
        "a/b/bias:0": rng.normal(size=(16,)).astype(np.float32),
        "images": rng.integers(0, 256, size=(5, 7, 9, 1), dtype=np.uint8),
        "labels64": rng.integers(-5, 5, size=(4, 3)).astype(np.int64),
        "dbl": rng.normal(size=(2, 2)),
        "names": np.array([b"img_0", b"image_11"], dtype="S8"),
        "scalar": np.float32(3.5),
        "empty": np.zeros((0, 4), np.float32),
    }
    p = tmp_path / "t.hdf5"
    with h5.H5Writer(p) as f:
        f.attrs["keras_version"] = "2.9.0"
        f.attrs["answer"] = np.int32(42)
        f.attrs["vec"] = np.arange(5, dtype=np.float64)
        g = f.create_group("a")
        g.attrs["layer_names"] = [b"conv2d", b"batch_normalization_12"]
        for k, v in data.items():
            ds = f.create_dataset(k, v)
        ds.attrs["unit"] = "px"
        for i in range(40):                      # a group with many links
            f.create_dataset(f"many/d{i:02d}", np.full((2,), i, np.int32))
    r = h5.H5File(p)
    assert r.attrs["keras_version"] == b"2.9.0" and r.attrs["answer"] == 42
    assert np.array_equal(r.attrs["vec"], np.arange(5.0))
    assert list(r["a"].attrs["layer_names"]) == [b"conv2d", b"batch_normalization_12"]
    for k, v in data.items():
        got = r[k].read()
        assert got.dtype == np.asarray(v).dtype and got.shape == np.asarray(v).shape, k
        assert np.array_equal(got, v), k
    assert r["empty"].attrs["unit"] == b"px"
    assert r["many"].keys() == [f"d{i:02d}" for i in range(40)]
    assert all(int(r[f"many/d{i:02d}"].read()[0]) == i for i in range(40))
    assert r["a"].is_group and not r["images"].is_group
    assert "a/b/c" in r and "nope" not in r
    with pytest.raises(KeyError):
        r["a/zzz"]


def test_file_structure_follows_the_spec(tmp_path):
    p = tmp_path / "s.h5"
    with h5.H5Writer(p) as f:
        f.create_dataset("x", np.arange(6, dtype=np.float32).reshape(2, 3))
    b = p.read_bytes()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8] == 0 and b[13] == 8 and b[14] == 8                  # superblock v0, 8-byte offsets/lengths
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == h5.UNDEF and eof == len(b) and drv == h5.UNDEF
    root_hdr = struct.unpack_from("<Q", b, 56 + 8)[0]
    ver, nmsg, refc, hsize = struct.unpack_from("<BxHII", b, root_hdr)
    assert ver == 1 and refc == 1 and nmsg == 1
    mtype, msize = struct.unpack_from("<HH", b, root_hdr + 16)
    assert mtype == 0x0011 and msize == 16                          # symbol-table message
    bt, heap = struct.unpack_from("<QQ", b, root_hdr + 24)
    assert b[bt:bt + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    snod = struct.unpack_from("<Q", b, bt + 24 + 8)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 1
    name_off, obj = struct.unpack_from("<QQ", b, snod + 8)
    heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
    assert b[heap_data + name_off:heap_data + name_off + 2] == b"x\0"
    # dataset header: dataspace, datatype (IEEE f32 LE), fill value, contiguous layout pointing at the data
    msgs = dict(h5.H5File(p)._read_header(obj))
    assert msgs[0x0001][:2] == b"\x01\x02" and struct.unpack_from("<QQ", msgs[0x0001], 8) == (2, 3)
    assert msgs[0x0003][0] == 0x11 and struct.unpack_from("<I", msgs[0x0003], 4)[0] == 4
    lver, lcls, addr, size = struct.unpack_from("<BBQQ", msgs[0x0008])
    assert (lver, lcls, size) == (3, 1, 24)
    assert np.array_equal(np.frombuffer(b[addr:addr + 24], "<f4"), np.arange(6, dtype=np.float32))


def test_reader_handles_chunked_deflate_and_vlen_strings(tmp_path):
    """Hand-assembled pieces h5py emits by default but our writer does not: a chunked + shuffled +
    deflated dataset indexed by a v1 B-tree, a variable-length string attribute in a global heap, and
    an object-header continuation block."""
    import zlib
    buf = bytearray(96)

    def alloc(x):
        while len(buf) % 8:
            buf.append(0)
        a = len(buf)
        buf.extend(x)
        return a

    arr = np.arange(4 * 6, dtype="<i4").reshape(4, 6)
    chunk_shape = (2, 6)
    entries = []
    for ci in range(2):
        raw = arr[2 * ci:2 * ci + 2].tobytes()
        shuf = np.frombuffer(raw, np.uint8).reshape(-1, 4).T.tobytes()
        comp = zlib.compress(shuf)
        entries.append(((2 * ci, 0, 0), len(comp), alloc(comp)))
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, 2, h5.UNDEF, h5.UNDEF)
    for offs, csize, addr in entries:
        node += struct.pack("<II3Q", csize, 0, *offs) + struct.pack("<Q", addr)
    node += struct.pack("<II3Q", 0, 0, 4, 0, 0)
    bt = alloc(node)
    gcol_obj = b"hello vlen"
    gcol = b"GCOL" + struct.pack("<B3xQ", 1, 4096) + struct.pack("<HHIQ", 1, 1, 0, len(gcol_obj)) + h5._pad8(gcol_obj)
    gcol += struct.pack("<HHIQ", 0, 0, 0, 4096 - len(gcol) - 16)
    gaddr = alloc(gcol + b"\0" * (4096 - len(gcol)))
    vlen_dt = struct.pack("<B3BI", 0x19, 0x01, 0, 0, 16) + h5._encode_dtype(np.dtype("S1"))
    nm = b"note\0"
    attr = struct.pack("<BxHHH", 1, len(nm), len(vlen_dt), 8) + h5._pad8(nm) + h5._pad8(vlen_dt) + \
        h5._encode_space(()) + struct.pack("<IQI", len(gcol_obj), gaddr, 1)
    cont = alloc(h5.H5Writer._msg(0x000C, attr))
    filt = struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", 4, 0) + \
        struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 6, 0)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt) + struct.pack("<III", 2, 6, 4)
    msgs = (h5.H5Writer._msg(0x0001, h5._encode_space(arr.shape)) + h5.H5Writer._msg(0x0003, h5._encode_dtype(arr.dtype)) +
            h5.H5Writer._msg(0x000B, filt) + h5.H5Writer._msg(0x0008, layout) +
            h5.H5Writer._msg(0x0010, struct.pack("<QQ", cont, len(h5.H5Writer._msg(0x000C, attr)))))
    ds = alloc(h5.H5Writer._header(msgs, 5))
    # a root group holding the dataset, via the writer's own group machinery
    heap = bytearray(8) + h5._pad8(b"d\0") + struct.pack("<QQ", 1, 16)
    hd = alloc(bytes(heap))
    hp = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 16, hd))
    sn = alloc(b"SNOD" + struct.pack("<BxH", 1, 1) + struct.pack("<QQII16x", 8, ds, 0, 0) + b"\0" * (40 * 7))
    tr = alloc(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, h5.UNDEF, h5.UNDEF) + struct.pack("<QQQ", 0, sn, 8) + b"\0" * 512)
    root = alloc(h5.H5Writer._header(h5.H5Writer._msg(0x0011, struct.pack("<QQ", tr, hp)), 1))
    sb = h5.SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, h5.UNDEF, len(buf), h5.UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tr, hp)
    buf[:96] = sb
    p = tmp_path / "hand.h5"
    p.write_bytes(bytes(buf))
    r = h5.H5File(p)
    assert np.array_equal(r["d"].read(), arr)
    assert r["d"].attrs["note"] == b"hello vlen"


def test_keras_weight_layout_roundtrip(tmp_path):
    from oct_image_segmentation_models_b200.common.synthetic import synthetic_weights
    from oct_image_segmentation_models_b200.models.unet_spec import unet_param_specs
    cfg = dict(input_channels=1, num_classes=4, start_neurons=8, pool_layers=2)
    specs = unet_param_specs(**cfg)
    weights = synthetic_weights(seed=1, **cfg)
    layers, cur = [], None
    for (name, _), w in zip(specs, weights):
        layer, wn = name.split("/", 1)
        if cur is None or cur[0] != layer:
            cur = (layer, [])
            layers.append(cur)
        cur[1].append((wn, w))
    layers.insert(2, ("activation", []))         # Keras lists weight-less layers too
    p = tmp_path / "model_epoch03.hdf5"
    h5.save_keras_weights(p, layers, model_config='{"class_name": "Functional", "config": {"name": "unet"}}')
    r = h5.H5File(p)
    assert r.attrs["backend"] == b"tensorflow"
    assert [x.decode() for x in r["model_weights"].attrs["layer_names"]][:3] == ["conv2d", "batch_normalization", "activation"]
    assert r["model_weights/conv2d/conv2d/kernel:0"].shape == (3, 3, 1, 8)
    assert [x.decode() for x in r["model_weights/batch_normalization"].attrs["weight_names"]] == [
        "batch_normalization/gamma:0", "batch_normalization/beta:0", "batch_normalization/moving_mean:0",
        "batch_normalization/moving_variance:0"]
    from oct_image_segmentation_models_b200.models.keras_like import read_weight_file
    name, got = read_weight_file(p)
    assert name == "unet" and len(got) == len(weights)
    assert all(np.array_equal(a, b) for a, b in zip(got, weights))
