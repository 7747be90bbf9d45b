"""Minimal pure-Python reader for the HDF5 files of the data pipeline (`ori_sample_grid.h5`: dataset `pc_sdf_sample`,
threedfront_dataset.py:387-391) -- h5py is not available in this image and cannot be installed.

Covers what h5py / libhdf5 write with their DEFAULT (earliest-compatible) file format, which is what the SDF pre-processing
of the reference's data produces (`create_dataset(name, data=..., compression='gzip', compression_opts=4)`):
  * superblock version 0 / 1 (8-byte offsets and lengths), root group through its symbol-table entry;
  * old-style groups: B-tree v1 (node type 0) over symbol-table nodes ("SNOD") with names in a local heap ("HEAP");
  * version-1 object headers with continuation blocks; messages: dataspace (v1 / v2), datatype (fixed-point and
    IEEE floating-point, little- or big-endian), data layout v1 - v3 (contiguous, chunked; compact), filter pipeline v1 / v2;
  * chunked datasets indexed by a B-tree v1 (node type 1), filters deflate (id 1) and shuffle (id 2), skipped filters
    honoured through the chunk's filter mask, edge chunks clipped, unallocated chunks = zeros (default fill value).
Not covered (an error, never a silent substitute): the "latest" format (superblock 2 / 3, version-2 object headers, fractal
heaps, B-tree v2 / extensible-array chunk indices), variable-length / compound types, external storage, other filters.

Layout facts follow the HDF5 File Format Specification, version 1.1 / 2.0 (sections III.A B-trees, III.C / III.D group
nodes and heaps, IV.A object headers and messages).  Pinning: the ONE genuine libhdf5-written file in this image (a MATLAB
v7.3 file among scipy's test data: 512-byte user block, superblock 0, symbol-table group, version-1 object header, version-2
layout message) is read correctly (tests/test_hdf5_cpu.py); chunked / gzip storage -- what the reference's grids use -- is
only exercised on files produced by an independent minimal writer (tests/hdf5_writer.py) that follows the same
specification, including a gzip-compressed, chunked 64^3 grid shaped like the reference's: **that part stays unpinned
against libhdf5**.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

__all__ = ["File", "read_dataset", "Hdf5Error"]

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(RuntimeError):
    pass


class _Dataset:
    def __init__(self, f: "File", name: str, shape, dtype, layout, filters):
        self._f, self.name, self.shape, self.dtype, self._layout, self._filters = f, name, tuple(shape), dtype, layout, filters

    def __getitem__(self, key):
        return self.read()[key]

    def read(self) -> np.ndarray:
        f, lay = self._f, self._layout
        n = int(np.prod(self.shape)) if self.shape else 1
        if lay["class"] == "compact":
            return np.frombuffer(lay["data"], self.dtype, n).reshape(self.shape).copy()
        if lay["class"] == "contiguous":
            if lay["address"] == _UNDEF:                         # never written: fill value (default zeros)
                return np.zeros(self.shape, self.dtype)
            raw = f._read(lay["address"], n * self.dtype.itemsize)
            return np.frombuffer(raw, self.dtype, n).reshape(self.shape).copy()
        # chunked: walk the B-tree, decode every chunk, paste its in-bounds part
        out = np.zeros(self.shape, self.dtype)
        cdims = lay["chunk"]
        if len(cdims) != len(self.shape):
            raise Hdf5Error(f"{self.name}: chunk rank {len(cdims)} != dataset rank {len(self.shape)}")
        csize = int(np.prod(cdims)) * self.dtype.itemsize
        if lay["address"] != _UNDEF:
            for offs, size, mask, addr in f._chunk_leaves(lay["address"], len(self.shape)):
                raw = f._read(addr, size)
                for i in reversed(range(len(self._filters))):    # the pipeline is undone last filter first
                    fid, cd = self._filters[i]
                    if mask & (1 << i):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else self.dtype.itemsize
                        raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes() if len(raw) % es == 0 else raw
                    else:
                        raise Hdf5Error(f"{self.name}: unsupported filter id {fid}")
                if len(raw) != csize:
                    raise Hdf5Error(f"{self.name}: chunk at {offs} decodes to {len(raw)} bytes, expected {csize}")
                chunk = np.frombuffer(raw, self.dtype).reshape(cdims)
                sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, self.shape))
                sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
                out[sel_out] = chunk[sel_in]
        return out


class File:
    """`File(path)["pc_sdf_sample"][:]` -- the two h5py calls the data pipeline makes.  Groups nest: f["a/b"]."""

    def __init__(self, path: str):
        self._fh = open(path, "rb")
        try:
            self._parse_superblock()
        except Exception:
            self._fh.close()
            raise

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ low level
    def _read(self, addr: int, n: int) -> bytes:
        self._fh.seek(self._base + addr)
        b = self._fh.read(n)
        if len(b) != n:
            raise Hdf5Error(f"truncated file: wanted {n} bytes at {addr}")
        return b

    def _parse_superblock(self):
        self._base = 0
        for off in (0, 512, 1024, 2048, 4096):                   # the superblock may sit behind a user block
            self._fh.seek(off)
            if self._fh.read(8) == _SIG:
                self._base = off
                break
        else:
            raise Hdf5Error("not an HDF5 file (no signature)")
        head = self._read(8, 16)
        version = head[0]
        if version not in (0, 1):
            raise Hdf5Error(f"superblock version {version}: only the classic format (0 / 1) is supported")
        size_offsets, size_lengths = head[5], head[6]
        if (size_offsets, size_lengths) != (8, 8):
            raise Hdf5Error("only 8-byte offsets / lengths are supported")
        pos = 8 + 16 + (4 if version == 1 else 0)                # v1 adds indexed-storage k + reserved
        # base address, free-space info address, end of file address, driver info block address
        pos += 4 * 8
        # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch (16)
        ent = self._read(pos, 40)
        self._root_header = struct.unpack_from("<Q", ent, 8)[0]

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr: int) -> List[Tuple[int, bytes]]:
        pre = self._read(addr, 16)
        if pre[0] != 1:
            raise Hdf5Error(f"object header version {pre[0]} at {addr}: only version 1 (classic format) is supported")
        nmsg, = struct.unpack_from("<H", pre, 2)
        hsize, = struct.unpack_from("<I", pre, 8)
        blocks = [(addr + 16, hsize)]
        out: List[Tuple[int, bytes]] = []
        while blocks and len(out) < nmsg:
            baddr, blen = blocks.pop(0)
            data = self._read(baddr, blen)
            p = 0
            while p + 8 <= blen and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", data, p)
                body = data[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:                                # continuation: more messages elsewhere
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr, clen))
                out.append((mtype, body))
        return out

    @staticmethod
    def _dataspace(body: bytes):
        version, rank, flags = body[0], body[1], body[2]
        p = 8 if version == 1 else 4
        return struct.unpack_from(f"<{rank}Q", body, p) if rank else ()

    @staticmethod
    def _datatype(body: bytes) -> np.dtype:
        cls, bits0 = body[0] & 0x0F, body[1]
        size, = struct.unpack_from("<I", body, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:                                             # fixed point; bit 3 of the class bit field = signed
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5Error(f"floating-point size {size}")
            return np.dtype(f"{order}f{size}")
        raise Hdf5Error(f"datatype class {cls} is not supported (fixed- and floating-point only)")

    @staticmethod
    def _filters(body: bytes):
        version, n = body[0], body[1]
        p = 8 if version == 1 else 2
        out = []
        for _ in range(n):
            fid, = struct.unpack_from("<H", body, p)
            if version == 1 or fid >= 256:
                nlen, = struct.unpack_from("<H", body, p + 2)
                p += 4
            else:
                nlen = 0
                p += 2
            _fl, ncd = struct.unpack_from("<HH", body, p)
            p += 4
            p += (nlen + 7) // 8 * 8 if version == 1 else nlen
            cd = struct.unpack_from(f"<{ncd}I", body, p)
            p += 4 * ncd
            if version == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _layout(self, body: bytes):
        version = body[0]
        if version in (1, 2):                                    # HDF5 1.6-era message: rank, class, 5 reserved, [address], sizes
            rank, cls = body[1], body[2]
            if cls == 0:
                size, = struct.unpack_from(f"<I", body, 8 + 4 * rank)
                return {"class": "compact", "data": body[12 + 4 * rank:12 + 4 * rank + size]}
            addr, = struct.unpack_from("<Q", body, 8)
            dims = struct.unpack_from(f"<{rank}I", body, 16)
            if cls == 1:
                return {"class": "contiguous", "address": addr, "size": None}
            if cls == 2:                                         # rank counts the trailing element-size entry
                return {"class": "chunked", "address": addr, "chunk": tuple(dims[:-1])}
            raise Hdf5Error(f"data layout class {cls}")
        if version != 3:
            raise Hdf5Error(f"data layout message version {version}: only versions 1-3 are supported")
        cls = body[1]
        if cls == 0:
            size, = struct.unpack_from("<H", body, 2)
            return {"class": "compact", "data": body[4:4 + size]}
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", body, 2)
            return {"class": "contiguous", "address": addr, "size": size}
        if cls == 2:
            rank = body[2]                                       # dataset rank + 1 (the last entry is the element size)
            addr, = struct.unpack_from("<Q", body, 3)
            dims = struct.unpack_from(f"<{rank}I", body, 11)
            return {"class": "chunked", "address": addr, "chunk": tuple(dims[:-1])}
        raise Hdf5Error(f"data layout class {cls}")

    # ------------------------------------------------------------------ groups
    def _group_entries(self, header_addr: int) -> Dict[str, int]:
        btree = heap = None
        for mtype, body in self._messages(header_addr):
            if mtype == 0x11:                                    # symbol table message
                btree, heap = struct.unpack_from("<QQ", body, 0)
        if btree is None:
            raise Hdf5Error("not an old-style group (no symbol table message): 'latest' format files are not supported")
        h = self._read(heap, 32)
        if h[:4] != b"HEAP":
            raise Hdf5Error("bad local heap signature")
        data_addr, = struct.unpack_from("<Q", h, 24)
        data_size, = struct.unpack_from("<Q", h, 8)
        names = self._read(data_addr, data_size)
        out: Dict[str, int] = {}

        def walk(addr):
            node = self._read(addr, 24)
            if node[:4] != b"TREE" or node[4] != 0:
                raise Hdf5Error("bad group B-tree node")
            level, used = node[5], struct.unpack_from("<H", node, 6)[0]
            body = self._read(addr + 24, (2 * used + 1) * 8)
            for i in range(used):
                child, = struct.unpack_from("<Q", body, (2 * i + 1) * 8)
                if level:
                    walk(child)
                else:
                    sn = self._read(child, 8)
                    if sn[:4] != b"SNOD":
                        raise Hdf5Error("bad symbol table node")
                    nsym, = struct.unpack_from("<H", sn, 6)
                    ents = self._read(child + 8, 40 * nsym)
                    for j in range(nsym):
                        noff, oaddr = struct.unpack_from("<QQ", ents, 40 * j)
                        end = names.index(b"\0", noff)
                        out[names[noff:end].decode()] = oaddr
        walk(btree)
        return out

    def keys(self):
        return list(self._group_entries(self._root_header))

    def __contains__(self, name: str) -> bool:
        try:
            self._resolve(name)
            return True
        except KeyError:
            return False

    def _resolve(self, name: str) -> int:
        addr = self._root_header
        for part in [p for p in name.split("/") if p]:
            ents = self._group_entries(addr)
            if part not in ents:
                raise KeyError(name)
            addr = ents[part]
        return addr

    def __getitem__(self, name: str) -> _Dataset:
        addr = self._resolve(name)
        shape = dtype = layout = None
        filters: list = []
        for mtype, body in self._messages(addr):
            if mtype == 0x01:
                shape = self._dataspace(body)
            elif mtype == 0x03:
                dtype = self._datatype(body)
            elif mtype == 0x08:
                layout = self._layout(body)
            elif mtype == 0x0B:
                filters = self._filters(body)
        if shape is None or dtype is None or layout is None:
            raise Hdf5Error(f"{name}: not a dataset (dataspace / datatype / layout message missing)")
        return _Dataset(self, name, shape, dtype, layout, filters)

    # ------------------------------------------------------------------ chunk index
    def _chunk_leaves(self, addr: int, rank: int):
        node = self._read(addr, 24)
        if node[:4] != b"TREE" or node[4] != 1:
            raise Hdf5Error("bad chunk B-tree node")
        level, used = node[5], struct.unpack_from("<H", node, 6)[0]
        key = 8 + 8 * (rank + 1)
        body = self._read(addr + 24, used * (key + 8) + key)
        for i in range(used):
            p = i * (key + 8)
            size, mask = struct.unpack_from("<II", body, p)
            offs = struct.unpack_from(f"<{rank}Q", body, p + 8)
            child, = struct.unpack_from("<Q", body, p + key)
            if level:
                yield from self._chunk_leaves(child, rank)
            else:
                yield offs, size, mask, child


def read_dataset(path: str, name: str) -> np.ndarray:
    with File(path) as f:
        return f[name].read()
