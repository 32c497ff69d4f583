"""A small pure-Python reader / writer for the subset of HDF5 that Keras 3 weight files use.

Why it exists: `.keras` checkpoints are zip archives holding `model.weights.h5`; neither h5py nor libhdf5 is in this image, and
the converter (tools/keras_to_p3w.py) must not need TensorFlow / Keras.  Keras writes one plain, contiguous, uncompressed
dataset per variable (`H5IOStore`: `group["vars"][str(i)] = value`) through h5py's default `libver='earliest'`, i.e. the
original on-disk structures of the HDF5 File Format Specification (version 3.0), which is all this module implements:

  * superblock version 0 (also 1), 8-byte offsets and lengths;
  * version-1 object headers, incl. continuation blocks (message 0x0010);
  * "old style" groups: symbol-table message (0x0011) -> version-1 B-tree of group nodes ("TREE" / "SNOD") + local heap ("HEAP");
  * datasets: dataspace (0x0001, versions 1 and 2), datatype (0x0003: fixed-point and IEEE floating point, little or big endian),
    data layout (0x0008: version 3 contiguous / compact, versions 1-2 contiguous); chunked / filtered data is rejected loudly.

The writer produces the same structures (it is how the converter's test fixture is made, and files it writes re-read bit for
bit); attributes, links, new-style groups and version-2 object headers are out of scope.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, Tuple, Union

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16  # symbols per group node = 2 * LEAF_K, children per B-tree node = 2 * INTERNAL_K

Tree = Dict[str, Union["Tree", np.ndarray]]


class H5FormatError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------------------ reader
class H5Reader:
    def __init__(self, data: bytes):
        self.b = data
        if data[:8] != SIGNATURE:
            raise H5FormatError("not an HDF5 file (signature missing at offset 0)")
        ver = data[8]
        if ver not in (0, 1):
            raise H5FormatError(f"superblock version {ver} is not supported (only the 'earliest' format Keras / h5py write by default)")
        if data[13] != 8 or data[14] != 8:
            raise H5FormatError("only 8-byte offsets / lengths are supported")
        off = 24 + (4 if ver == 1 else 0)  # v1 adds indexed-storage K + reserved
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", data, off)
        root = off + 32
        self.root_header = struct.unpack_from("<Q", data, root + 8)[0]

    # -- object headers (version 1)
    def _messages(self, addr: int) -> Iterator[Tuple[int, bytes]]:
        b = self.b
        if b[addr:addr + 4] == b"OHDR":
            raise H5FormatError("version-2 object headers are not supported (file was not written with libver='earliest')")
        version, _, nmsgs, _refcount, size = struct.unpack_from("<BBHII", b, addr)
        if version != 1:
            raise H5FormatError(f"object header version {version} at {addr} is not supported")
        chunks = [(addr + 16, size)]
        seen = 0
        while chunks and seen < nmsgs:
            pos, left = chunks.pop(0)
            end = pos + left
            while pos + 8 <= end and seen < nmsgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                seen += 1
                if mtype == 0x0010:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    chunks.append((caddr, clen))
                else:
                    yield mtype, body

    def _heap_name(self, heap_addr: int, offset: int) -> str:
        b = self.b
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("local heap signature missing")
        data_addr = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        start = data_addr + offset
        end = b.index(b"\0", start)
        return b[start:end].decode("utf-8")

    def _group_entries(self, btree: int, heap: int) -> Iterator[Tuple[str, int]]:
        b = self.b
        if b[btree:btree + 4] == b"SNOD":
            n = struct.unpack_from("<H", b, btree + 6)[0]
            for i in range(n):
                name_off, header = struct.unpack_from("<QQ", b, btree + 8 + 40 * i)
                yield self._heap_name(heap, name_off), header
            return
        if b[btree:btree + 4] != b"TREE":
            raise H5FormatError("B-tree node signature missing")
        ntype, _level, used = struct.unpack_from("<BBH", b, btree + 4)
        if ntype != 0:
            raise H5FormatError("not a group B-tree")
        pos = btree + 24
        for i in range(used):
            child = struct.unpack_from("<Q", b, pos + 8 + 16 * i)[0]  # key_i (8), child_i (8), ...
            yield from self._group_entries(child, heap)

    def _dataset(self, msgs: Dict[int, bytes]) -> np.ndarray:
        space, dtype_b, layout = msgs[0x0001], msgs[0x0003], msgs[0x0008]
        sver, rank, sflags = space[0], space[1], space[2]
        dims_at = 8 if sver == 1 else 4
        dims = struct.unpack_from("<" + "Q" * rank, space, dims_at) if rank else ()
        cls, bits0, size = dtype_b[0] & 0x0F, dtype_b[1], struct.unpack_from("<I", dtype_b, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 1:
            kind = "f"
        elif cls == 0:
            kind = "i" if bits0 & 0x08 else "u"
        else:
            raise H5FormatError(f"datatype class {cls} is not supported (only integers and IEEE floats)")
        dt = np.dtype(f"{order}{kind}{size}")
        count = int(np.prod(dims)) if rank else 1
        lver = layout[0]
        if lver == 3:
            lclass = layout[1]
            if lclass == 1:
                addr, nbytes = struct.unpack_from("<QQ", layout, 2)
                raw = b"" if addr == UNDEF else self.b[self.base + addr:self.base + addr + nbytes]
            elif lclass == 0:
                nbytes = struct.unpack_from("<H", layout, 2)[0]
                raw = layout[4:4 + nbytes]
            else:
                raise H5FormatError("chunked / filtered datasets are not supported (Keras writes contiguous data)")
        elif lver in (1, 2):
            lrank, lclass = layout[1], layout[2]
            if lclass != 1:
                raise H5FormatError("only contiguous data is supported")
            addr = struct.unpack_from("<Q", layout, 8)[0]
            raw = self.b[self.base + addr:self.base + addr + count * size]
        else:
            raise H5FormatError(f"data layout version {lver} is not supported")
        if len(raw) < count * size:
            raise H5FormatError("dataset data is truncated")
        return np.frombuffer(raw[:count * size], dtype=dt).reshape(dims).astype(dt.newbyteorder("="), copy=True)

    def _read_object(self, addr: int) -> Union[Tree, np.ndarray]:
        msgs: Dict[int, bytes] = {}
        for mtype, body in self._messages(self.base + addr):
            msgs.setdefault(mtype, body)
        if 0x0011 in msgs:
            btree, heap = struct.unpack_from("<QQ", msgs[0x0011], 0)
            out: Tree = {}
            for name, header in self._group_entries(self.base + btree, self.base + heap):
                out[name] = self._read_object(header)
            return out
        if 0x0008 in msgs and 0x0001 in msgs and 0x0003 in msgs:
            return self._dataset(msgs)
        if 0x0002 in msgs or 0x0006 in msgs:
            raise H5FormatError("new-style (link message) groups are not supported (file was not written with libver='earliest')")
        return {}

    def tree(self) -> Tree:
        root = self._read_object(self.root_header)
        if not isinstance(root, dict):
            raise H5FormatError("root object is not a group")
        return root


def read_h5(data: bytes) -> Dict[str, np.ndarray]:
    """All datasets of an HDF5 file image as {'group/sub/name': array}."""
    out: Dict[str, np.ndarray] = {}

    def walk(node: Tree, prefix: str):
        for k, v in node.items():
            p = f"{prefix}/{k}" if prefix else k
            if isinstance(v, dict):
                walk(v, p)
            else:
                out[p] = v

    walk(H5Reader(data).tree(), "")
    return out


# ------------------------------------------------------------------------------------------------------------ writer
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _object_header(msgs: bytes, nmsgs: int) -> bytes:
    return struct.pack("<BBHII4x", 1, 0, nmsgs, 1, len(msgs)) + msgs


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)  # superblock goes in last

    def alloc(self, data: bytes) -> int:
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.asarray(arr)
        if not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)  # (never for 0-d arrays: ascontiguousarray would make them 1-d)
        if arr.dtype.kind == "f":
            size = arr.dtype.itemsize
            exp_bits, mant_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
            dtype = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size) + struct.pack("<HHBBBBI", 0, size * 8, mant_bits, exp_bits, 0,
                                                                                             mant_bits, bias)
        elif arr.dtype.kind in "iu":
            size = arr.dtype.itemsize
            dtype = struct.pack("<BBBBI", 0x10, 0x08 if arr.dtype.kind == "i" else 0x00, 0, 0, size) + struct.pack("<HH", 0, size * 8)
        else:
            raise H5FormatError(f"cannot write dtype {arr.dtype}")
        raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
        data_addr = self.alloc(raw) if raw else UNDEF
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        fill = struct.pack("<BBBB", 2, 2, 0, 0)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw))
        msgs = _msg(0x0001, space) + _msg(0x0003, dtype) + _msg(0x0005, fill) + _msg(0x0008, layout)
        return self.alloc(_object_header(msgs, 4))

    def group(self, node: Tree) -> Tuple[int, int, int]:
        """Writes the group's children, then its heap, group nodes, B-tree and object header; returns (header, btree, heap)."""
        children = []
        for name in sorted(node, key=lambda s: s.encode("utf-8")):  # B-tree order = strcmp order of the link names
            child = node[name]
            if isinstance(child, dict):
                header, bt, hp = self.group(child)
                children.append((name, header, 1, struct.pack("<QQ", bt, hp)))
            else:
                children.append((name, self.dataset(np.asarray(child)), 0, b"\0" * 16))
        heap_data = bytearray(b"\0" * 8)  # offset 0: the empty name
        offsets = []
        for name, *_ in children:
            offsets.append(len(heap_data))
            heap_data += _pad8(name.encode("utf-8") + b"\0")
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, data_addr))
        per = 2 * LEAF_K
        nodes = [children[i:i + per] for i in range(0, len(children), per)] or [[]]
        if len(nodes) > 2 * INTERNAL_K:
            raise H5FormatError("too many links in one group for this writer (max 256)")
        keys, kids = [0], []
        idx = 0
        for chunk in nodes:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for (name, header, ctype, scratch) in chunk:
                body += struct.pack("<QQII", offsets[idx], header, ctype, 0) + scratch
                idx += 1
            body += b"\0" * (40 * (per - len(chunk)))
            kids.append(self.alloc(body))
            keys.append(offsets[idx - 1] if chunk else 0)
        bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(kids), UNDEF, UNDEF)
        for i in range(2 * INTERNAL_K):
            bt += struct.pack("<Q", keys[i] if i < len(keys) else 0)
            bt += struct.pack("<Q", kids[i] if i < len(kids) else UNDEF)
        bt += struct.pack("<Q", keys[len(kids)] if len(kids) < len(keys) and len(kids) == 2 * INTERNAL_K else 0)
        btree = self.alloc(bt)
        header = self.alloc(_object_header(_msg(0x0011, struct.pack("<QQ", btree, heap)), 1))
        return header, btree, heap


def write_h5(tree: Tree) -> bytes:
    """Serialises a nested {name: subgroup-dict | array} tree as an HDF5 file image (structures listed in the module docstring)."""
    w = _Writer()
    header, btree, heap = w.group(tree)
    eof = len(w.buf)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    w.buf[:96] = sb
    return bytes(w.buf)
