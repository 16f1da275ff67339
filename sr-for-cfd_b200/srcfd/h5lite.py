"""Minimal pure-Python HDF5 reader/writer (no h5py in this image).

Covers exactly the subset the reference's files use -- what h5py/Keras write
with default settings: superblock v0, old-style groups (symbol table + v1
B-tree + local heap), v1 object headers with continuation blocks, simple
dataspaces, little-endian fixed/float/fixed-length-string datatypes,
contiguous or compact layouts, and v1-v3 attributes.

Read side:  encoder weights `vanilla_encoder10_to_400_*.h5`
            (PyCFD_ML_accelerated.py:831 loads them through Keras) and the
            solver result files (layout written at PyCFD_ML_accelerated.py:523-544).
Write side: `write_h5` produces the same group/dataset/attribute layout so the
            drop-in `_save_results_hdf5` keeps the reference's output format.
"""
from __future__ import annotations

import struct
from typing import Any, Dict

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(IOError):
    pass


class Dataset:
    def __init__(self, data: np.ndarray, attrs: Dict[str, Any]):
        self.data = data
        self.attrs = attrs

    def __getitem__(self, key):
        return self.data[key]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)

    @property
    def shape(self):
        return self.data.shape


class Group(dict):
    """dict of name -> Group | Dataset, plus .attrs"""

    def __init__(self):
        super().__init__()
        self.attrs: Dict[str, Any] = {}

    def visit(self, prefix=""):
        for k, v in self.items():
            path = f"{prefix}/{k}"
            yield path, v
            if isinstance(v, Group):
                yield from v.visit(path)

    def get_path(self, path: str):
        node = self
        for part in path.strip("/").split("/"):
            if part:
                node = node[part]
        return node


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != _SIG:
            raise H5Error("not an HDF5 file")
        ver = buf[8]
        if ver not in (0, 1):
            raise H5Error(f"unsupported superblock version {ver}")
        self.O, self.L = buf[13], buf[14]
        if self.O != 8 or self.L != 8:
            raise H5Error("only 8-byte offsets/lengths supported")
        pos = 24 if ver == 0 else 28
        self.base = struct.unpack_from("<Q", buf, pos)[0]
        pos += 32
        self.root_entry = pos

    # ---- primitives -----------------------------------------------------
    def u(self, pos, n):
        return int.from_bytes(self.b[pos:pos + n], "little")

    # ---- datatypes --------------------------------------------------------
    def parse_dtype(self, pos):
        cv = self.b[pos]
        cls, bits0 = cv & 0x0F, self.b[pos + 1]
        size = self.u(pos + 4, 4)
        if cls == 0:  # fixed point
            signed = bool(bits0 & 0x08)
            order = ">" if bits0 & 1 else "<"
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), 8 + 4
        if cls == 1:  # float
            order = ">" if bits0 & 1 else "<"
            return np.dtype(f"{order}f{size}"), 8 + 12
        if cls == 3:  # fixed-length string
            return np.dtype(f"S{size}"), 8
        if cls == 9:  # variable length
            return ("vlen", size), 8
        raise H5Error(f"unsupported datatype class {cls}")

    def parse_dataspace(self, pos):
        ver, rank, flags = self.b[pos], self.b[pos + 1], self.b[pos + 2]
        if ver == 1:
            p = pos + 8
        elif ver == 2:
            if self.b[pos + 3] == 2:  # null dataspace
                return None
            p = pos + 4
        else:
            raise H5Error(f"unsupported dataspace version {ver}")
        return tuple(self.u(p + 8 * i, 8) for i in range(rank))

    def read_vlen_string(self, raw, off):
        # global heap reference: length(4) collection address(8) index(4)
        length = int.from_bytes(raw[off:off + 4], "little")
        addr = int.from_bytes(raw[off + 4:off + 12], "little")
        idx = int.from_bytes(raw[off + 12:off + 16], "little")
        if addr in (0, _UNDEF):
            return b""
        a = self.base + addr
        if self.b[a:a + 4] != b"GCOL":
            raise H5Error("bad global heap")
        csize = self.u(a + 8, 8)
        p, end = a + 16, a + csize
        while p < end:
            oidx = self.u(p, 2)
            osize = self.u(p + 8, 8)
            if oidx == idx:
                return self.b[p + 16:p + 16 + length]
            if oidx == 0:
                break
            p += 16 + ((osize + 7) // 8) * 8
        raise H5Error("global heap object not found")

    def decode(self, dtype, shape, raw):
        if shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dtype, tuple):  # vlen (strings)
            vals = [self.read_vlen_string(raw, 16 * i) for i in range(n)]
            vals = [v.decode("utf8", "replace") for v in vals]
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        arr = np.frombuffer(raw, dtype=dtype, count=n).reshape(shape)
        if not shape:
            v = arr[()]
            return v.decode("utf8", "replace") if dtype.kind == "S" else v
        return arr.copy()

    # ---- object headers ---------------------------------------------------
    def messages(self, addr):
        a = self.base + addr
        if self.b[a:a + 4] == b"OHDR":
            raise H5Error("v2 object headers not supported")
        if self.b[a] != 1:
            raise H5Error(f"unsupported object header version {self.b[a]}")
        nmsg = self.u(a + 2, 2)
        hsize = self.u(a + 8, 4)
        blocks = [(a + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize = self.u(p, 2), self.u(p + 2, 2)
                body = p + 8
                if mtype == 0x0010:
                    blocks.append((self.base + self.u(body, 8), self.u(body + 8, 8)))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    def parse_attribute(self, pos):
        ver = self.b[pos]
        nsz, tsz, ssz = self.u(pos + 2, 2), self.u(pos + 4, 2), self.u(pos + 6, 2)
        p = pos + 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = self.b[p:p + nsz].split(b"\0")[0].decode("utf8")
        p += pad(nsz)
        dtype, _ = self.parse_dtype(p)
        p += pad(tsz)
        shape = self.parse_dataspace(p)
        p += pad(ssz)
        if shape is None:
            return name, None
        n = int(np.prod(shape)) if shape else 1
        isz = 16 if isinstance(dtype, tuple) else dtype.itemsize
        return name, self.decode(dtype, shape, self.b[p:p + n * isz])

    def read_object(self, addr):
        msgs = self.messages(addr)
        attrs, dtype, shape, layout, stab = {}, None, None, None, None
        for mtype, body, msize in msgs:
            if mtype == 0x0001:
                shape = self.parse_dataspace(body)
            elif mtype == 0x0003:
                dtype, _ = self.parse_dtype(body)
            elif mtype == 0x0008:
                ver, cls = self.b[body], self.b[body + 1]
                if ver != 3:
                    raise H5Error(f"unsupported layout version {ver}")
                if cls == 1:
                    layout = ("contiguous", self.u(body + 2, 8), self.u(body + 10, 8))
                elif cls == 0:
                    sz = self.u(body + 2, 2)
                    layout = ("compact", body + 4, sz)
                else:
                    raise H5Error("chunked datasets not supported")
            elif mtype == 0x000C:
                k, v = self.parse_attribute(body)
                attrs[k] = v
            elif mtype == 0x0011:
                stab = (self.u(body, 8), self.u(body + 8, 8))
        if stab is not None:
            g = Group()
            g.attrs = attrs
            for name, child in self.iter_group(*stab):
                g[name] = self.read_object(child)
            return g
        if dtype is None or layout is None:
            g = Group()
            g.attrs = attrs
            return g
        n = int(np.prod(shape)) if shape else 1
        isz = 16 if isinstance(dtype, tuple) else dtype.itemsize
        if layout[0] == "contiguous":
            if layout[1] == _UNDEF:
                raw = b"\0" * (n * isz)
            else:
                a = self.base + layout[1]
                raw = self.b[a:a + n * isz]
        else:
            raw = self.b[layout[1]:layout[1] + layout[2]]
        return Dataset(self.decode(dtype, shape, raw), attrs)

    # ---- old-style groups -------------------------------------------------
    def heap_data(self, heap_addr):
        a = self.base + heap_addr
        if self.b[a:a + 4] != b"HEAP":
            raise H5Error("bad local heap")
        return self.base + self.u(a + 24, 8)

    def iter_group(self, btree_addr, heap_addr):
        data = self.heap_data(heap_addr)
        yield from self._iter_btree(btree_addr, data)

    def _iter_btree(self, addr, heap):
        a = self.base + addr
        if self.b[a:a + 4] != b"TREE":
            raise H5Error("bad B-tree node")
        level, used = self.b[a + 5], self.u(a + 6, 2)
        p = a + 24
        for i in range(used):
            child = self.u(p + 8, 8)  # key_i (8) then child_i (8)
            p += 16
            if level > 0:
                yield from self._iter_btree(child, heap)
            else:
                s = self.base + child
                if self.b[s:s + 4] != b"SNOD":
                    raise H5Error("bad symbol node")
                nsym = self.u(s + 6, 2)
                e = s + 8
                for _ in range(nsym):
                    name_off, obj = self.u(e, 8), self.u(e + 8, 8)
                    q = heap + name_off
                    name = self.b[q:self.b.index(b"\0", q)].decode("utf8")
                    yield name, obj
                    e += 40

    def root(self):
        obj = self.u(self.root_entry + 8, 8)
        return self.read_object(obj)


def read_h5(path: str) -> Group:
    """Read a whole file into nested Group/Dataset objects."""
    with open(path, "rb") as f:
        return _Reader(f.read()).root()


# ==========================================================================
# writer
# ==========================================================================
class _Writer:
    """Emits superblock v0 + symbol-table groups + v1 headers + contiguous data."""

    def __init__(self):
        self.buf = bytearray()

    def align(self, n=8):
        while len(self.buf) % n:
            self.buf.append(0)

    def alloc(self, data: bytes) -> int:
        self.align()
        off = len(self.buf)
        self.buf += data
        return off

    @staticmethod
    def dtype_msg(dt: np.dtype) -> bytes:
        if dt.kind == "f":
            size = dt.itemsize
            if size == 8:
                props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
                bits = bytes([0x20, 63, 0])
            elif size == 4:
                props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
                bits = bytes([0x20, 31, 0])
            else:
                raise H5Error("float size")
            return bytes([0x11]) + bits + struct.pack("<I", size) + props
        if dt.kind in "iu":
            bits = bytes([0x08 if dt.kind == "i" else 0x00, 0, 0])
            return bytes([0x10]) + bits + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
        if dt.kind == "S":
            return bytes([0x13, 0x00, 0, 0]) + struct.pack("<I", dt.itemsize)
        raise H5Error(f"unsupported dtype {dt}")

    @staticmethod
    def space_msg(shape) -> bytes:
        out = struct.pack("<BBBBI", 1, len(shape), 0, 0, 0)
        for d in shape:
            out += struct.pack("<Q", d)
        return out

    @staticmethod
    def _pad8(b: bytes) -> bytes:
        return b + b"\0" * (-len(b) % 8)

    def attr_msg(self, name: str, value) -> bytes:
        if isinstance(value, str):
            value = np.bytes_(value.encode("utf8"))
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = arr.astype("S")
        if arr.dtype.kind == "i":
            arr = arr.astype("<i8")
        if arr.dtype.kind == "f":
            arr = arr.astype("<f8")
        nm = name.encode("utf8") + b"\0"
        dt, sp = self.dtype_msg(arr.dtype), self.space_msg(arr.shape)
        head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp))
        return head + self._pad8(nm) + self._pad8(dt) + self._pad8(sp) + arr.tobytes()

    def object_header(self, msgs) -> int:
        body = b""
        for m in msgs:
            mtype, data, flags = (m + (0,))[:3]
            data = self._pad8(data)
            body += struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data
        hdr = struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4
        return self.alloc(hdr + body)

    def write_dataset(self, arr: np.ndarray, attrs) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype.kind == "f" and arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        data_off = self.alloc(arr.tobytes()) if arr.size else _UNDEF
        # the messages h5py's create_dataset emits, in its order: dataspace, datatype (constant), fill value v2
        # (allocate late, write if set, defined, no value; constant), contiguous layout v3
        msgs = [(0x0001, self.space_msg(arr.shape)), (0x0003, self.dtype_msg(arr.dtype), 1),
                (0x0005, bytes([2, 2, 2, 1, 0, 0, 0, 0]), 1),
                (0x0008, struct.pack("<BBQQ", 3, 1, data_off, arr.nbytes))]
        msgs += [(0x000C, self.attr_msg(k, v)) for k, v in attrs.items()]
        return self.object_header(msgs)

    def write_group(self, children: Dict[str, int], attrs) -> tuple[int, int, int]:
        names = sorted(children)
        heap = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += n.encode("utf8") + b"\0"
            while len(heap) % 8:
                heap.append(0)
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)  # free block: next=1 (none), size
        heap_data = self.alloc(bytes(heap))
        heap_hdr = self.alloc(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap), free_off, heap_data))
        # symbol nodes: up to 2K = 8 entries each (leaf K = 4)
        K = 4
        snods, keys = [], [0]
        for i in range(0, max(len(names), 1), 2 * K):
            chunk = names[i:i + 2 * K]
            ent = b""
            for n in chunk:
                ent += struct.pack("<QQII", offs[n], children[n], 0, 0) + b"\0" * 16
            ent += b"\0" * (40 * (2 * K - len(chunk)))
            snods.append(self.alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)) + ent))
            keys.append(offs[chunk[-1]] if chunk else 0)
        IK = 16
        if len(snods) > 2 * IK:
            raise H5Error("too many group entries for the minimal writer")
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), _UNDEF, _UNDEF)
        for i, s in enumerate(snods):
            node += struct.pack("<QQ", keys[i], s)
        node += struct.pack("<Q", keys[len(snods)])
        node += b"\0" * (16 * (2 * IK - len(snods)))
        btree = self.alloc(node)
        msgs = [(0x0011, struct.pack("<QQ", btree, heap_hdr))]
        msgs += [(0x000C, self.attr_msg(k, v)) for k, v in attrs.items()]
        return self.object_header(msgs), btree, heap_hdr

    def write_node(self, node) -> tuple[int, int, int]:
        if isinstance(node, Group) or isinstance(node, dict):
            kids = {}
            for k, v in node.items():
                kids[k] = self.write_node(v)[0]
            return self.write_group(kids, getattr(node, "attrs", {}))
        if isinstance(node, Dataset):
            return self.write_dataset(node.data, node.attrs), 0, 0
        return self.write_dataset(np.asarray(node), {}), 0, 0


def write_h5(path: str, root) -> None:
    """Write nested dict/Group of arrays/Datasets as an HDF5 file (see module doc)."""
    w = _Writer()
    w.buf += b"\0" * 96  # superblock placeholder (56 bytes + 40-byte root entry)
    obj, btree, heap = w.write_node(root if isinstance(root, (dict, Group)) else {"data": root})
    w.align()
    eof = len(w.buf)
    sb = _SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII", 0, obj, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(path, "wb") as f:
        f.write(bytes(w.buf))
