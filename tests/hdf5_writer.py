"""TEST INFRASTRUCTURE -- a minimal HDF5 *writer* (classic file format: superblock 0, symbol-table root group, version-1
object headers, contiguous or chunked + deflate [+ shuffle] datasets) used to produce the fixtures the pure-Python reader
`commonscenes_b200/dataset/hdf5_lite.py` is tested on.  h5py / libhdf5 do not exist in this image; this follows the HDF5 File
Format Specification independently of the reader (own layout code, forward direction), so a misreading of the specification
would have to be made twice in the same way to go unnoticed -- it is still not a substitute for files written by libhdf5."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dataspace(shape) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    big = 1 if dt.byteorder == ">" else 0
    if dt.kind == "f":
        # class 1, version 1; bit field: byte order, padding, mantissa normalisation (implied msb = 2 << 4), sign location
        exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
        bits = big | (2 << 4) | ((dt.itemsize * 8 - 1) << 8)
        props = struct.pack("<HHBBBBI", 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, bias)
        return struct.pack("<B", 0x11) + struct.pack("<I", bits)[:3] + struct.pack("<I", dt.itemsize) + props
    if dt.kind in "iu":
        bits = big | (8 if dt.kind == "i" else 0)
        props = struct.pack("<HH", 0, dt.itemsize * 8)
        return struct.pack("<B", 0x10) + struct.pack("<I", bits)[:3] + struct.pack("<I", dt.itemsize) + props
    raise ValueError(dt)


def _filters(ids, itemsize) -> bytes:
    out = struct.pack("<BB6x", 1, len(ids))
    for fid in ids:
        cd = [itemsize] if fid == 2 else [4]           # shuffle: element size; deflate: level
        out += struct.pack("<HHHH", fid, 0, 1, len(cd)) + b"".join(struct.pack("<I", v) for v in cd)
        if len(cd) % 2:
            out += b"\0" * 4
    return out


class Writer:
    def __init__(self):
        self.buf = bytearray(b"\0" * 96)                  # superblock + root symbol table entry, filled in at the end
        self.datasets = []                                # (name, object header address)

    def _alloc(self, data: bytes, align: int = 8) -> int:
        self.buf += b"\0" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _header(self, messages: bytes, nmsg: int) -> int:
        return self._alloc(struct.pack("<BxHII4x", 1, nmsg, 1, len(messages)) + messages)

    def add(self, name, arr, chunks=None, gzip=False, shuffle=False, skip_filter_on_chunk=None, split_header=False):
        arr = np.ascontiguousarray(arr)
        msgs, n = _msg(0x01, _dataspace(arr.shape)), 1
        msgs += _msg(0x03, _datatype(arr.dtype), flags=1); n += 1
        if chunks is None:
            addr = self._alloc(arr.tobytes())
            layout = struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)
        else:
            fids = ([2] if shuffle else []) + ([1] if gzip else [])
            if fids:
                msgs += _msg(0x0B, _filters(fids, arr.dtype.itemsize)); n += 1
            rank = arr.ndim
            entries = []
            grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
            for idx, offs in enumerate(np.array(np.meshgrid(*grid, indexing="ij")).reshape(rank, -1).T):
                block = np.zeros(chunks, arr.dtype)
                sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, arr.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sel)] = arr[sel]
                raw, mask = block.tobytes(), 0
                for i, fid in enumerate(fids):
                    if skip_filter_on_chunk is not None and idx == skip_filter_on_chunk and fid == 1:
                        mask |= 1 << i                   # "filter skipped for this chunk" (e.g. incompressible data)
                        continue
                    if fid == 2:
                        raw = np.frombuffer(raw, np.uint8).reshape(-1, arr.dtype.itemsize).T.tobytes()
                    else:
                        raw = zlib.compress(raw, 4)
                entries.append((tuple(int(o) for o in offs), len(raw), mask, self._alloc(raw)))
            # one leaf B-tree node (type 1): keys (size, mask, offsets.., 0) interleaved with child addresses + a final key
            node = struct.pack("<4sBBHQQ", b"TREE", 1, 0, len(entries), UNDEF, UNDEF)
            for offs, size, mask, caddr in entries:
                node += struct.pack("<II", size, mask) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
                node += struct.pack("<Q", caddr)
            node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
            baddr = self._alloc(node)
            layout = struct.pack("<BBBQ", 3, 2, rank + 1, baddr) + b"".join(struct.pack("<I", c) for c in chunks) + \
                struct.pack("<I", arr.dtype.itemsize)
        if split_header:
            # the layout message lives in a continuation block (object headers written incrementally look like this)
            cont = self._alloc(_msg(0x08, layout))
            msgs += _msg(0x10, struct.pack("<QQ", cont, len(_msg(0x08, layout)))); n += 2
        else:
            msgs += _msg(0x08, layout); n += 1
        self.datasets.append((name, self._header(msgs, n)))

    def tobytes(self) -> bytes:
        # local heap with the link names (offset 0 = empty string for the root), one SNOD, one group B-tree leaf
        names, offs = b"\0" * 8, {}
        for name, _ in sorted(self.datasets):
            offs[name] = len(names)
            names += _pad8(name.encode() + b"\0")
        heap_data = self._alloc(names)
        heap = self._alloc(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(names), UNDEF, heap_data))
        snod = struct.pack("<4sBxH", b"SNOD", 1, len(self.datasets))
        for name, oaddr in sorted(self.datasets):
            snod += struct.pack("<QQII16x", offs[name], oaddr, 0, 0)
        snod_addr = self._alloc(snod)
        last = offs[sorted(self.datasets)[-1][0]]
        btree = self._alloc(struct.pack("<4sBBHQQ", b"TREE", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, last))
        root = self._header(_msg(0x11, struct.pack("<QQ", btree, heap)), 1)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)
