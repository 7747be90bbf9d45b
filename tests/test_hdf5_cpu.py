"""Pure-Python HDF5 reader (commonscenes_b200/dataset/hdf5_lite.py; SURVEY.md 8(f)-4: `ori_sample_grid.h5['pc_sdf_sample']`,
threedfront_dataset.py:387-391) on files produced by the independent minimal writer tests/hdf5_writer.py: contiguous and
chunked layouts, gzip / shuffle filters, ragged edge chunks, a per-chunk skipped filter, a continuation block, several dtypes
and byte orders -- and the data-pipeline entry point `load_sdf_grid` on a grid shaped and stored like the reference's."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hdf5_writer import Writer  # noqa: E402

from commonscenes_b200.dataset import formats, hdf5_lite  # noqa: E402


def _write(tmp_path, name, build):
    w = Writer()
    build(w)
    p = tmp_path / name
    p.write_bytes(w.tobytes())
    return str(p)


def test_sdf_grid_file_like_the_reference(tmp_path):
    rng = np.random.default_rng(0)
    sdf = (rng.standard_normal(64 ** 3) * 0.3).astype(np.float32)             # 262144 samples, values beyond the clamp
    other = rng.standard_normal((1000, 4)).astype(np.float32)

    def build(w):                                                              # the four datasets of the SDF pre-processing
        w.add("pc_sdf_original", other, chunks=(250, 2), gzip=True)
        w.add("pc_sdf_sample", sdf, chunks=(8192,), gzip=True)
        w.add("norm_params", np.arange(4, dtype=np.float32), chunks=(4,), gzip=True)
        w.add("sdf_params", np.arange(6, dtype=np.float64), chunks=(6,), gzip=True)
    path = _write(tmp_path, "ori_sample_grid.h5", build)
    with hdf5_lite.File(path) as f:
        assert sorted(f.keys()) == ["norm_params", "pc_sdf_original", "pc_sdf_sample", "sdf_params"]
        d = f["pc_sdf_sample"]
        assert d.shape == (64 ** 3,) and d.dtype == np.float32
        assert np.array_equal(d[:], sdf) and np.array_equal(f["pc_sdf_original"][:], other)
        assert f["sdf_params"].dtype == np.float64 and np.array_equal(f["sdf_params"][:], np.arange(6.0))
        assert "pc_sdf_sample" in f and "nope" not in f
        with pytest.raises(KeyError):
            f["nope"]
    # the data-pipeline call: (1, 64, 64, 64) fp32, clamped to +-0.2 (threedfront_dataset.py:387-391)
    g = formats.load_sdf_grid(path)
    assert g.shape == (1, 64, 64, 64) and g.dtype == torch.float32
    assert torch.equal(g, torch.from_numpy(sdf).view(1, 64, 64, 64).clamp(-0.2, 0.2))


def test_layouts_filters_and_types(tmp_path):
    rng = np.random.default_rng(1)
    a = rng.standard_normal((64, 64, 64)).astype(np.float32)
    b = rng.integers(-1000, 1000, (50, 7)).astype(np.int32)
    c = rng.standard_normal((5, 3)).astype(">f8")
    e = rng.integers(0, 255, (33,)).astype(np.uint8)

    def build(w):
        w.add("contiguous", a)
        w.add("chunked_gzip", a, chunks=(16, 16, 32), gzip=True)
        w.add("ragged_shuffle_gzip", b, chunks=(16, 4), gzip=True, shuffle=True)           # edge chunks are clipped
        w.add("big_endian", c)
        w.add("bytes_chunked_raw", e, chunks=(8,))                                           # chunked without filters
        w.add("one_chunk_unfiltered", b, chunks=(16, 4), gzip=True, skip_filter_on_chunk=3)  # filter mask bit set on one chunk
        w.add("split_header", a[:4], chunks=(2, 64, 64), gzip=True, split_header=True)       # layout message in a continuation
    path = _write(tmp_path, "mixed.h5", build)
    with hdf5_lite.File(path) as f:
        assert np.array_equal(f["contiguous"][:], a)
        assert np.array_equal(f["chunked_gzip"][:], a)
        assert np.array_equal(f["ragged_shuffle_gzip"][:], b) and f["ragged_shuffle_gzip"].dtype == np.int32
        assert np.array_equal(f["big_endian"][:], c) and f["big_endian"].dtype == np.dtype(">f8")
        assert np.array_equal(f["bytes_chunked_raw"][:], e)
        assert np.array_equal(f["one_chunk_unfiltered"][:], b)
        assert np.array_equal(f["split_header"][:], a[:4])
        assert np.array_equal(f["chunked_gzip"][3:5, ::7, 1], a[3:5, ::7, 1])               # h5py-style slicing of the read


def test_rejects_what_it_does_not_parse(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 10)
    with pytest.raises(hdf5_lite.Hdf5Error):
        hdf5_lite.File(str(p))
    w = Writer()
    w.add("d", np.zeros(4, np.float32))
    raw = bytearray(w.tobytes())
    raw[8] = 2                                            # superblock version 2 ("latest" format): refused, not guessed
    p.write_bytes(bytes(raw))
    with pytest.raises(hdf5_lite.Hdf5Error):
        hdf5_lite.File(str(p))
    p.write_bytes(w.tobytes()[:120])                      # truncated
    with pytest.raises(hdf5_lite.Hdf5Error):
        with hdf5_lite.File(str(p)) as f:
            f["d"][:]


def test_reads_a_genuine_libhdf5_file():
    """The one file in this image written by the real HDF5 library: scipy's test datum `testhdf5_7.4_GLNX86.mat` (a MATLAB
    v7.3 file = HDF5 behind a 512-byte user block; BSD-licensed scipy test data, copied to tests/golden/).  scipy's own
    expectation for it (scipy/io/matlab/tests/test_mio.py: 'testdouble' = pi / 4 * arange(9), MATLAB shape (1, 9), i.e. (9, 1)
    in HDF5's row-major order) pins superblock, user block, symbol-table group, local heap, version-1 object header,
    dataspace / datatype messages and the 1.6-era layout message of the reader against libhdf5's output."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "libhdf5_matlab73_testdouble.mat")
    with hdf5_lite.File(path) as f:
        assert f.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64
        assert np.array_equal(d[:].ravel(), np.pi / 4 * np.arange(9, dtype=float))
