"""Minimal HDF5 reader / writer (no h5py or libhdf5 offline) -- SURVEY section 8 row f-1.

Covers what Keras `.hdf5` model files and the reference's dataset files use
(reference common/utils.py:63-69, training/training.py:319-326, common/dataset_loader.py:9-33):

  writer : superblock v0, old-style groups (v1 B-tree + local heap + one symbol-table node per
           group), v1 object headers, contiguous little-endian datasets (u8/i32/i64/f32/f64/fixed
           strings), attributes (scalars, arrays, fixed-length strings / string arrays).
  reader : the same, plus what h5py emits with its defaults: object-header continuation blocks,
           multi-level group B-trees, compact / contiguous / chunked layouts (v1 chunk B-tree) with
           deflate + shuffle filters, variable-length strings (global heap), compact "link message"
           groups, superblock v0-v3, object headers v1 and v2.

PARITY UNPINNED: no file written by real h5py/Keras exists in this sandbox; the reader is checked
against this writer and against hand-assembled byte patterns from the public HDF5 file-format
specification (tests/test_hdf5_min.py).  Layout notes: SURVEY.md App. B.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


# =====================================================================================
# datatype / dataspace encoding
# =====================================================================================
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * ((-len(b)) % 8)


def _encode_dtype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "S":
        return struct.pack("<B3BI", 0x13, 0x01, 0, 0, dt.itemsize)          # class 3 v1, null-padded ASCII
    if dt.kind in "ui":
        flags = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<B3BIHH", 0x10, flags, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<B3BIHHBBBBI", 0x11, 0x20, 0x1F, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<B3BIHHBBBBI", 0x11, 0x20, 0x3F, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    raise TypeError(f"unsupported dtype for HDF5 writer: {dt}")


def _encode_space(shape: Tuple[int, ...]) -> bytes:
    if shape == ():
        return struct.pack("<BBB5x", 1, 0, 0)
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _as_array(value) -> np.ndarray:
    if isinstance(value, str):
        value = value.encode("utf8")
    if isinstance(value, bytes):
        return np.array(value, dtype=f"S{max(1, len(value))}")
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf8")
    if a.dtype.kind == "O":
        a = np.array([x.encode("utf8") if isinstance(x, str) else x for x in a.ravel()]).reshape(a.shape)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype.kind in "uif" and a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    return a


# =====================================================================================
# writer
# =====================================================================================
class _Node:
    def __init__(self, is_group: bool):
        self.is_group = is_group
        self.children: Dict[str, "_Node"] = {}
        self.attrs: Dict[str, np.ndarray] = {}
        self.data: Optional[np.ndarray] = None
        self.addr = 0


class H5Writer:
    """Build the tree in memory, then `close()` lays it out and writes the file.

        with H5Writer(path) as f:
            f.attrs["keras_version"] = "2.9.0"
            g = f.create_group("model_weights")
            g.attrs["layer_names"] = [b"conv2d", b"batch_normalization"]
            f.create_dataset("model_weights/conv2d/conv2d/kernel:0", data=w)
    """

    class _GroupView:
        def __init__(self, node: _Node):
            self._node = node
            self.attrs = node.attrs

    def __init__(self, path):
        self.path = str(path)
        self.root = _Node(True)
        self.attrs = self.root.attrs

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()

    def _walk(self, path: str, create: bool) -> _Node:
        node = self.root
        for part in [p for p in path.split("/") if p]:
            if part not in node.children:
                if not create:
                    raise KeyError(path)
                node.children[part] = _Node(True)
            node = node.children[part]
        return node

    def create_group(self, path: str):
        return H5Writer._GroupView(self._walk(path, True))

    def require_group(self, path: str):
        return self.create_group(path)

    def create_dataset(self, path: str, data):
        parts = [p for p in path.split("/") if p]
        parent = self._walk("/".join(parts[:-1]), True)
        n = _Node(False)
        n.data = np.array(_as_array(data), order="C", copy=True)   # (ascontiguousarray would promote 0-d to 1-d)
        parent.children[parts[-1]] = n
        return H5Writer._GroupView(n)

    # ---- layout -----------------------------------------------------------------------
    @staticmethod
    def _msg(mtype: int, data: bytes) -> bytes:
        data = _pad8(data)
        return struct.pack("<HHB3x", mtype, len(data), 0) + data

    def _attr_msgs(self, node: _Node) -> bytes:
        out = b""
        for name, value in node.attrs.items():
            a = _as_array(value)
            nm = name.encode("utf8") + b"\0"
            dt, sp = _encode_dtype(a.dtype), _encode_space(a.shape)
            body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + a.tobytes()
            if len(body) > 0xFFF0:
                raise ValueError(f"attribute {name} too large for one object-header message")
            out += self._msg(0x000C, body)
        return out

    @staticmethod
    def _header(msgs: bytes, nmsgs: int) -> bytes:
        return struct.pack("<BxHII4x", 1, nmsgs, 1, len(msgs)) + msgs

    def close(self):
        buf = bytearray(96)                       # superblock placeholder

        def max_children(n: _Node) -> int:
            return max([len(n.children)] + [max_children(c) for c in n.children.values() if c.is_group])

        # "group leaf node K" is a file-wide superblock parameter: a symbol-table node holds up to
        # 2*K entries, so one node per group suffices when K >= half the largest group
        LEAF_K = max(4, (max_children(self.root) + 1) // 2)

        def alloc(b: bytes, align: int = 8) -> int:
            while len(buf) % align:
                buf.append(0)
            addr = len(buf)
            buf.extend(b)
            return addr

        def emit(node: _Node) -> int:
            nattr = len(node.attrs)
            if not node.is_group:
                a = node.data
                daddr = alloc(a.tobytes()) if a.size else UNDEF
                msgs = (self._msg(0x0001, _encode_space(a.shape)) + self._msg(0x0003, _encode_dtype(a.dtype)) +
                        self._msg(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)) +
                        self._msg(0x0008, struct.pack("<BBQQ", 3, 1, daddr, a.nbytes)) + self._attr_msgs(node))
                node.addr = alloc(self._header(msgs, 4 + nattr))
                return node.addr
            names = sorted(node.children)         # symbol-table entries are ordered by name (strcmp)
            if len(names) > 2 * LEAF_K:
                raise ValueError("too many links in one group for the minimal writer")
            for nm in names:
                emit(node.children[nm])
            # local heap: offset 0 = empty string, then the names, then one free block
            heap = bytearray(8)
            offs = {}
            for nm in names:
                offs[nm] = len(heap)
                heap.extend(_pad8(nm.encode("utf8") + b"\0"))
            free_off = len(heap)
            heap.extend(struct.pack("<QQ", 1, 16))            # last free block: next = 1 (none), size 16
            heap_data = alloc(bytes(heap))
            heap_addr = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
            snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
            for nm in names:
                ch = node.children[nm]
                if ch.is_group:
                    snod += struct.pack("<QQII", offs[nm], ch.addr, 1, 0) + struct.pack("<QQ", ch.btree, ch.heap)
                else:
                    snod += struct.pack("<QQII16x", offs[nm], ch.addr, 0, 0)
            snod += b"\0" * (40 * (2 * LEAF_K - len(names)))
            snod_addr = alloc(snod)
            last = offs[names[-1]] if names else 0
            if names:
                bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, last)
            else:
                bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 0, UNDEF, UNDEF) + struct.pack("<Q", 0)
            bt += b"\0" * (24 + (2 * 16 + 1) * 8 + 2 * 16 * 8 - len(bt))   # node sized for internal K = 16
            node.btree = alloc(bt)
            node.heap = heap_addr
            msgs = self._msg(0x0011, struct.pack("<QQ", node.btree, node.heap)) + self._attr_msgs(node)
            node.addr = alloc(self._header(msgs, 1 + nattr))
            return node.addr

        root_addr = emit(self.root)
        eof = len(buf) + ((-len(buf)) % 8)
        sb = SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", self.root.btree, self.root.heap)
        assert len(sb) == 96
        buf[:96] = sb
        buf.extend(b"\0" * (eof - len(buf)))
        with open(self.path, "wb") as fh:
            fh.write(bytes(buf))


# =====================================================================================
# reader
# =====================================================================================
class H5Object:
    def __init__(self, f: "H5File", addr: int, name: str):
        self._f, self._addr, self.name = f, addr, name
        self._msgs = f._read_header(addr)
        self.attrs = {}
        for mtype, data in self._msgs:
            if mtype == 0x000C:
                k, v = f._parse_attr(data)
                self.attrs[k] = v
        self.is_group = any(m in (0x0011, 0x0002, 0x0006) for m, _ in self._msgs) and \
            not any(m == 0x0008 for m, _ in self._msgs)

    # ---- groups -------------------------------------------------------------------------
    def _links(self) -> Dict[str, int]:
        if not hasattr(self, "_link_cache"):
            links: Dict[str, int] = {}
            for mtype, data in self._msgs:
                if mtype == 0x0011:
                    bt, heap = struct.unpack_from("<QQ", data)
                    links.update(self._f._read_group_btree(bt, heap))
                elif mtype == 0x0006:
                    nm, addr = self._f._parse_link(data)
                    if nm is not None:
                        links[nm] = addr
            self._link_cache = links
        return self._link_cache

    def keys(self) -> List[str]:
        return sorted(self._links())

    def __contains__(self, key: str) -> bool:
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str) -> "H5Object":
        obj = self
        for part in [p for p in path.split("/") if p]:
            links = obj._links()
            if part not in links:
                raise KeyError(f"{path!r} not found in {self.name!r}")
            obj = H5Object(self._f, links[part], obj.name.rstrip("/") + "/" + part)
        return obj

    # ---- datasets -----------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return self._f._parse_space(dict(self._msgs)[0x0001])

    @property
    def dtype(self):
        return self._f._parse_dtype(dict(self._msgs)[0x0003])[0]

    def read(self) -> np.ndarray:
        msgs = dict(self._msgs)
        shape = self._f._parse_space(msgs[0x0001])
        dt, vlen = self._f._parse_dtype(msgs[0x0003])
        if vlen:
            raise NotImplementedError("variable-length datasets are not supported")
        raw = self._f._read_layout(msgs[0x0008], msgs.get(0x000B), shape, dt)
        return np.frombuffer(raw, dtype=dt, count=int(np.prod(shape, dtype=np.int64))).reshape(shape).copy()

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a


class H5File(H5Object):
    def __init__(self, path):
        with open(path, "rb") as fh:
            self._buf = fh.read()
        self._gcol: Dict[int, Dict[int, bytes]] = {}
        base = self._find_superblock()
        b = self._buf
        ver = b[base + 8]
        if ver in (0, 1):
            so, sl = b[base + 13], b[base + 14]
            if so != 8 or sl != 8:
                raise NotImplementedError("only 8-byte offsets/lengths are supported")
            off = base + 24 + (4 if ver == 1 else 0)
            self._base = struct.unpack_from("<Q", b, off)[0]
            root_entry = off + 32
            root_addr = struct.unpack_from("<Q", b, root_entry + 8)[0]
        elif ver in (2, 3):
            self._base = struct.unpack_from("<Q", b, base + 12)[0]
            root_addr = struct.unpack_from("<Q", b, base + 12 + 24)[0]
        else:
            raise NotImplementedError(f"superblock version {ver}")
        super().__init__(self, root_addr, "/")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def _find_superblock(self) -> int:
        off = 0
        while off < len(self._buf):
            if self._buf[off:off + 8] == SIG:
                return off
            off = 512 if off == 0 else off * 2
        raise ValueError("not an HDF5 file")

    # ---- object headers -------------------------------------------------------------------
    def _read_header(self, addr: int) -> List[Tuple[int, bytes]]:
        b = self._buf
        addr += self._base
        msgs: List[Tuple[int, bytes]] = []
        if b[addr:addr + 4] == b"OHDR":                      # version 2
            flags = b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            szlen = 1 << (flags & 3)
            chunk = int.from_bytes(b[p:p + szlen], "little")
            p += szlen
            blocks = [(p, chunk)]
            tracked = bool(flags & 0x04)
            while blocks:
                p, size = blocks.pop(0)
                end = p + size
                while p + 4 <= end:
                    mtype = b[p]
                    msize = struct.unpack_from("<H", b, p + 1)[0]
                    p += 4 + (2 if tracked else 0)
                    data = b[p:p + msize]
                    p += msize
                    if mtype == 0x10:
                        o, ln = struct.unpack_from("<QQ", data)
                        blocks.append((o + self._base + 4, ln - 8))     # skip OCHK signature, drop checksum
                    elif mtype != 0:
                        msgs.append((mtype, data))
            return msgs
        ver, nmsgs, _, hsize = struct.unpack_from("<BxHII", b, addr)
        if ver != 1:
            raise ValueError(f"bad object header at {addr}")
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < 10000:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end:
                mtype, msize = struct.unpack_from("<HH", b, p)
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:
                    o, ln = struct.unpack_from("<QQ", data)
                    blocks.append((o + self._base, ln))
                elif mtype != 0:
                    msgs.append((mtype, data))
        return msgs

    # ---- groups ---------------------------------------------------------------------------
    def _heap_name(self, heap_addr: int, off: int) -> str:
        b = self._buf
        h = heap_addr + self._base
        assert b[h:h + 4] == b"HEAP", "bad local heap"
        data_addr = struct.unpack_from("<Q", b, h + 24)[0] + self._base
        end = b.index(b"\0", data_addr + off)
        return b[data_addr + off:end].decode("utf8")

    def _read_group_btree(self, bt_addr: int, heap_addr: int) -> Dict[str, int]:
        b = self._buf
        out: Dict[str, int] = {}
        p = bt_addr + self._base
        if b[p:p + 4] == b"SNOD":
            n = struct.unpack_from("<H", b, p + 6)[0]
            for i in range(n):
                name_off, obj = struct.unpack_from("<QQ", b, p + 8 + 40 * i)
                out[self._heap_name(heap_addr, name_off)] = obj
            return out
        assert b[p:p + 4] == b"TREE", "bad group B-tree node"
        level, used = struct.unpack_from("<BH", b, p + 5)
        q = p + 24
        for i in range(used):
            child = struct.unpack_from("<Q", b, q + 8 + 16 * i)[0]
            out.update(self._read_group_btree(child, heap_addr))
        return out

    def _parse_link(self, data: bytes):
        ver, flags = data[0], data[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = data[p]; p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nlen_size = 1 << (flags & 3)
        nlen = int.from_bytes(data[p:p + nlen_size], "little"); p += nlen_size
        name = data[p:p + nlen].decode("utf8"); p += nlen
        if ltype != 0:
            return None, 0
        return name, struct.unpack_from("<Q", data, p)[0]

    # ---- datatypes / dataspaces -------------------------------------------------------------
    def _parse_space(self, data: bytes) -> Tuple[int, ...]:
        ver, rank = data[0], data[1]
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", data, off + 8 * i)[0] for i in range(rank))

    def _parse_dtype(self, data: bytes):
        cls = data[0] & 0x0F
        b0 = data[1]
        size = struct.unpack_from("<I", data, 4)[0]
        if cls == 0:
            kind = "i" if b0 & 0x08 else "u"
            return np.dtype(("<" if not (b0 & 1) else ">") + f"{kind}{size}"), False
        if cls == 1:
            return np.dtype(("<" if not (b0 & 1) else ">") + f"f{size}"), False
        if cls == 3:
            return np.dtype(f"S{size}"), False
        if cls == 9:
            return np.dtype("O"), True                       # variable length (strings)
        if cls == 8:                                          # enum (h5py stores bool like this)
            base, _ = self._parse_dtype(data[8:])
            return base, False
        raise NotImplementedError(f"HDF5 datatype class {cls}")

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        if addr not in self._gcol:
            b = self._buf
            p = addr + self._base
            assert b[p:p + 4] == b"GCOL", "bad global heap collection"
            size = struct.unpack_from("<Q", b, p + 8)[0]
            objs: Dict[int, bytes] = {}
            q, end = p + 16, p + size
            while q + 16 <= end:
                idx, _, _, osize = struct.unpack_from("<HHIQ", b, q)
                if idx == 0:
                    break
                objs[idx] = b[q + 16:q + 16 + osize]
                q += 16 + osize + ((-osize) % 8)
            self._gcol[addr] = objs
        return self._gcol[addr][index]

    def _parse_attr(self, data: bytes):
        ver = data[0]
        if ver == 1:
            nsz, dsz, ssz = struct.unpack_from("<HHH", data, 2)
            p = 8
            name = data[p:p + nsz].split(b"\0")[0].decode("utf8"); p += nsz + ((-nsz) % 8)
            dt_raw = data[p:p + dsz]; p += dsz + ((-dsz) % 8)
            sp_raw = data[p:p + ssz]; p += ssz + ((-ssz) % 8)
        else:
            nsz, dsz, ssz = struct.unpack_from("<HHH", data, 2)
            p = 8 + (1 if ver == 3 else 0)
            name = data[p:p + nsz].split(b"\0")[0].decode("utf8"); p += nsz
            dt_raw = data[p:p + dsz]; p += dsz
            sp_raw = data[p:p + ssz]; p += ssz
        shape = self._parse_space(sp_raw) if len(sp_raw) >= 2 and sp_raw[1] > 0 else ()
        dt, vlen = self._parse_dtype(dt_raw)
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        raw = data[p:]
        if vlen:
            vals = []
            for i in range(count):
                ln, gaddr, gidx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self._global_heap_object(gaddr, gidx)[:ln])
            arr = np.array(vals, dtype=object).reshape(shape) if shape else vals[0]
        else:
            arr = np.frombuffer(raw, dtype=dt, count=count).reshape(shape).copy()
            if not shape:
                arr = arr[()]
        return name, arr

    # ---- dataset storage ----------------------------------------------------------------------
    def _read_layout(self, layout: bytes, filters: Optional[bytes], shape, dt) -> bytes:
        b = self._buf
        ver = layout[0]
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        if ver != 3:
            raise NotImplementedError(f"data layout message version {ver}")
        cls = layout[1]
        if cls == 0:                                          # compact
            size = struct.unpack_from("<H", layout, 2)[0]
            return layout[4:4 + size]
        if cls == 1:                                          # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF:
                return b"\0" * nbytes
            return b[addr + self._base:addr + self._base + nbytes]
        if cls == 2:                                          # chunked, v1 B-tree index
            rank = layout[2]
            bt = struct.unpack_from("<Q", layout, 3)[0]
            cdims = struct.unpack_from(f"<{rank}I", layout, 11)
            chunk_shape = cdims[:-1]
            pipeline = self._parse_filters(filters) if filters else []
            out = np.zeros(shape, dtype=dt)
            if bt != UNDEF:
                for offs, fmask, raw in self._iter_chunks(bt, rank):
                    for k, (fid, cd) in enumerate(reversed(pipeline)):
                        if fmask & (1 << (len(pipeline) - 1 - k)):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            n = len(raw) // dt.itemsize
                            raw = np.frombuffer(raw, np.uint8).reshape(dt.itemsize, n).T.tobytes()
                        else:
                            raise NotImplementedError(f"HDF5 filter {fid}")
                    chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(chunk_shape))).reshape(chunk_shape)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk_shape, shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
            return out.tobytes()
        raise NotImplementedError(f"data layout class {cls}")

    def _parse_filters(self, data: bytes):
        ver, n = data[0], data[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", data, p)[0]
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", data, p + 2)[0]
                flags, ncd = struct.unpack_from("<HH", data, p + 4)
                p += 8 + nlen + ((-nlen) % 8 if ver == 1 else 0)
            else:
                flags, ncd = struct.unpack_from("<HH", data, p + 2)
                p += 6
            cd = struct.unpack_from(f"<{ncd}I", data, p)
            p += 4 * ncd + (4 if (ver == 1 and ncd % 2) else 0)
            out.append((fid, cd))
        return out

    def _iter_chunks(self, addr: int, rank: int):
        b = self._buf
        p = addr + self._base
        assert b[p:p + 4] == b"TREE", "bad chunk B-tree node"
        level, used = struct.unpack_from("<BH", b, p + 5)
        q = p + 24
        key_size = 8 + 8 * rank
        for i in range(used):
            k = q + i * (key_size + 8)
            csize, fmask = struct.unpack_from("<II", b, k)
            offs = struct.unpack_from(f"<{rank}Q", b, k + 8)[:-1]
            child = struct.unpack_from("<Q", b, k + key_size)[0]
            if level > 0:
                yield from self._iter_chunks(child, rank)
            else:
                yield offs, fmask, b[child + self._base:child + self._base + csize]


# =====================================================================================
# Keras model files
# =====================================================================================
def save_keras_weights(path, layer_weights: List[Tuple[str, List[Tuple[str, np.ndarray]]]], model_config: str = "",
                       keras_version: str = "2.9.0", backend: str = "tensorflow",
                       optimizer_weights: Optional[List[Tuple[str, np.ndarray]]] = None, training_config: str = ""):
    """Write a Keras-2.x style `.hdf5`: root attrs (keras_version, backend, model_config, training_config), a
    `model_weights` group with `layer_names`, per-layer `weight_names` and datasets
    `<layer>/<layer>/<weight>:0` -- the layout Keras `load_weights` reads -- and, when
    given, an `optimizer_weights` group (attr `weight_names`, one dataset per optimizer variable) as
    `model.save()` / ModelCheckpoint add for a compiled model.  layer_weights: [(layer_name, [(weight_name,
    array), ...]), ...] in model.layers order (layers without weights appear with an empty list).
    NOTE: `model_config` is whatever string the caller passes; this package passes a stub that names the graph
    hyper-parameters, not Keras' full Functional layer graph, so `tf.keras.models.load_model` cannot rebuild the
    model from such a file -- `load_weights` into a model built by the reference's `UNet.build_model()` can."""
    with H5Writer(path) as f:
        f.attrs["keras_version"] = keras_version
        f.attrs["backend"] = backend
        if model_config:
            f.attrs["model_config"] = model_config
        if training_config:
            f.attrs["training_config"] = training_config
        if optimizer_weights:
            og = f.create_group("optimizer_weights")
            on = [n.encode("utf8") for n, _ in optimizer_weights]
            og.attrs["weight_names"] = np.array(on, dtype=f"S{max(len(n) for n in on)}")
            for n, arr in optimizer_weights:
                f.create_dataset(f"optimizer_weights/{n}", data=np.asarray(arr))
        g = f.create_group("model_weights")
        g.attrs["keras_version"] = keras_version
        g.attrs["backend"] = backend
        names = [ln.encode("utf8") for ln, _ in layer_weights]
        g.attrs["layer_names"] = np.array(names, dtype=f"S{max(len(n) for n in names)}")
        for ln, ws in layer_weights:
            lg = f.create_group(f"model_weights/{ln}")
            wn = [f"{ln}/{w}".encode("utf8") for w, _ in ws]
            lg.attrs["weight_names"] = (np.array(wn, dtype=f"S{max(len(n) for n in wn)}") if wn
                                        else np.zeros((0,), dtype="S1"))
            for w, arr in ws:
                f.create_dataset(f"model_weights/{ln}/{ln}/{w}", data=np.asarray(arr, np.float32))


def load_keras_weights(path) -> Tuple[List[Tuple[str, List[Tuple[str, np.ndarray]]]], Optional[str]]:
    """[(layer_name, [(weight_name, array), ...]), ...] in `layer_names` order, and model_config (str)."""
    def _s(x):
        return x.decode("utf8") if isinstance(x, (bytes, np.bytes_)) else str(x)

    f = H5File(path)
    root = f["model_weights"] if "model_weights" in f else f
    out = []
    for ln in [_s(x) for x in np.atleast_1d(root.attrs["layer_names"])]:
        lg = root[ln]
        ws = []
        for wn in [_s(x) for x in np.atleast_1d(lg.attrs.get("weight_names", []))]:
            ws.append((wn, lg[wn].read()))
        out.append((ln, ws))
    cfg = f.attrs.get("model_config")
    return out, (_s(cfg) if cfg is not None else None)


def load_keras_optimizer_weights(path) -> Tuple[List[Tuple[str, np.ndarray]], Optional[str]]:
    """([(variable_name, array), ...] in `weight_names` order -- empty when the file has no optimizer state --,
    training_config string or None)."""
    def _s(x):
        return x.decode("utf8") if isinstance(x, (bytes, np.bytes_)) else str(x)

    f = H5File(path)
    tc = f.attrs.get("training_config")
    if "optimizer_weights" not in f:
        return [], (_s(tc) if tc is not None else None)
    og = f["optimizer_weights"]
    out = [(n, og[n].read()) for n in [_s(x) for x in np.atleast_1d(og.attrs["weight_names"])]]
    return out, (_s(tc) if tc is not None else None)
