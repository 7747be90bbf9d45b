"""On-disk formats either side of the shape branch (data half): what `ThreedFrontDatasetSceneGraph.__getitem__`
reads per object and per scene, restated as small host functions plus a pinned-memory prefetcher that keeps the VQ-VAE
encoder fed (64^3 fp32 = 1 MiB per object).

Reference lines (dataset/threedfront_dataset.py):
  * SDF grids  :382-391  `<model dir with 3D-FUTURE-model -> 3D-FUTURE-SDF>/ori_sample_grid.h5`, dataset `pc_sdf_sample`
                         (262144 floats) -> view (1, 64, 64, 64) -> clamp(-0.2, 0.2); objects without a model (floor, the
                         `_scene_` node :455-456) are all-zero grids
  * CLIP cache :279-287 path rule, :396-409 read + re-ordering to the scene's instance order, :480-488 write,
                :503-508 relation features looked up by the relation's words

HDF5: the reference reads the grids with h5py, which this image does not have and which cannot be installed (no network).
`load_sdf_grid` uses h5py WHEN IMPORTABLE and otherwise the pure-Python reader `dataset/hdf5_lite.py` (classic HDF5 file
format: contiguous or chunked + gzip datasets -- what the SDF pre-processing writes); an `ori_sample_grid.npy` exported next
to the .h5 (`export_sdf_npy`) is still accepted.  A file neither can read is an error, never a silent substitute.  The
container parsing is pinned against files produced by an independent minimal writer that follows the HDF5 specification
(tests/test_hdf5_cpu.py), NOT against libhdf5 (no HDF5 library or file in the image: parity unpinned); everything after the
array is read (shape, dtype, clamp, zero grids, ordering) is pinned by tests/test_formats_cpu.py.
"""
from __future__ import annotations

import os
import pickle
import queue
import threading
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

SDF_KEY = "pc_sdf_sample"
SDF_CLAMP = 0.2


# ---------------------------------------------------------------------------------------------------------------------
# SDF grids
# ---------------------------------------------------------------------------------------------------------------------
def sdf_path_for_model(model_path: str) -> str:
    """threedfront_dataset.py:387: the grid lives beside the mesh, in the 3D-FUTURE-SDF mirror of the model tree."""
    return os.path.join(model_path.replace("3D-FUTURE-model", "3D-FUTURE-SDF").rsplit("/", 1)[0], "ori_sample_grid.h5")


def _read_h5(path: str) -> np.ndarray:
    try:
        import h5py  # noqa: WPS433  (optional dependency, absent in this image)
    except ImportError:
        from . import hdf5_lite         # classic-format reader in pure Python (superblock 0, chunked + gzip datasets)
        with hdf5_lite.File(path) as f:
            return f[SDF_KEY][:].astype(np.float32)
    with h5py.File(path, "r") as f:
        return f[SDF_KEY][:].astype(np.float32)


def export_sdf_npy(h5_path: str) -> str:
    """`ori_sample_grid.h5` -> `ori_sample_grid.npy` (same 262144 floats, fp32): run once where h5py is available."""
    out = os.path.splitext(h5_path)[0] + ".npy"
    np.save(out, _read_h5(h5_path))
    return out


def load_sdf_grid(path: Optional[str], res: int = 64, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(1, res, res, res) fp32 SDF of one object, clamped to +-0.2 like the reference (:389-390); `path` None = an object
    without a model (floor / `_scene_`): zeros (:384-385, :455-456).  `path` may name the .h5 (read with h5py when importable, else with
    dataset/hdf5_lite.py; its exported .npy twin is used when the .h5 itself is missing) or a .npy directly.  `out`: optional destination (e.g. a slice of a pinned batch buffer)."""
    if out is None:
        out = torch.empty((1, res, res, res), dtype=torch.float32)
    if tuple(out.shape) != (1, res, res, res) or out.dtype != torch.float32:
        raise ValueError(f"load_sdf_grid: destination must be fp32 (1, {res}, {res}, {res}), got {out.dtype} {tuple(out.shape)}")
    if path is None:
        out.zero_()
        return out
    npy = path if path.endswith(".npy") else os.path.splitext(path)[0] + ".npy"
    if path.endswith(".npy") or (not os.path.exists(path) and os.path.exists(npy)):
        arr = np.load(npy).astype(np.float32, copy=False)
    else:
        arr = _read_h5(path)
    if arr.size != res ** 3:
        raise ValueError(f"{path}: expected {res ** 3} SDF samples, found {arr.size}")
    out.copy_(torch.from_numpy(np.ascontiguousarray(arr)).view(1, res, res, res))
    out.clamp_(min=-SDF_CLAMP, max=SDF_CLAMP)
    return out


def _have_h5py() -> bool:
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


class SdfPrefetcher:
    """Background loader of per-scene SDF batches into PINNED host buffers, with the host->device copy issued on a side
    stream one batch ahead: the VQ-VAE encoder of step i runs while the grids of step i+1 are read, clamped and copied.

        pf = SdfPrefetcher(scenes, device="cuda:0")          # scenes: iterable of lists of paths (None = zero grid)
        for sdfs in pf:                                       # (n_objects, 1, 64, 64, 64) fp32 on the device
            ...

    Two pinned buffers of `max_objects` grids rotate; a batch is only handed out once its copy has completed on the side
    stream and the consumer's stream has been made to wait for it (events, no host sync in the steady state)."""

    def __init__(self, scenes: Iterable[Sequence[Optional[str]]], device="cuda", res: int = 64, max_objects: int = 64, depth: int = 2):
        self.scenes = iter(scenes)
        self.res, self.max_objects = res, max_objects
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self._host = [torch.empty((max_objects, 1, res, res, res), dtype=torch.float32, pin_memory=self.cuda) for _ in range(depth)]
        self._free: "queue.Queue[int]" = queue.Queue()
        for i in range(depth):
            self._free.put(i)
        self._ready: "queue.Queue" = queue.Queue(maxsize=depth)
        self._stream = torch.cuda.Stream(self.device) if self.cuda else None
        self._thread = threading.Thread(target=self._work, daemon=True)
        self._thread.start()

    def _work(self):
        try:
            for paths in self.scenes:
                n = len(paths)
                if n > self.max_objects:
                    raise ValueError(f"scene with {n} objects exceeds max_objects={self.max_objects}")
                slot = self._free.get()
                buf = self._host[slot]
                for i, p in enumerate(paths):
                    load_sdf_grid(p, self.res, out=buf[i])
                self._ready.put((slot, n, None))
            self._ready.put(None)
        except Exception as e:       # surfaced to the consumer, not swallowed
            self._ready.put((None, 0, e))

    def __iter__(self):
        return self

    def __next__(self) -> torch.Tensor:
        item = self._ready.get()
        if item is None:
            raise StopIteration
        slot, n, err = item
        if err is not None:
            raise err
        host = self._host[slot][:n]
        if not self.cuda:
            out = host.clone()
            self._free.put(slot)
            return out
        with torch.cuda.stream(self._stream):
            dev = host.to(self.device, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._stream)
        torch.cuda.current_stream(self.device).wait_event(done)
        dev.record_stream(torch.cuda.current_stream(self.device))
        done.synchronize()              # the pinned slot may be rewritten once the copy has left it
        self._free.put(slot)
        return dev


# ---------------------------------------------------------------------------------------------------------------------
# cached CLIP features
# ---------------------------------------------------------------------------------------------------------------------
def clip_feats_path(root_3dfront: str, scan_id: str, large: bool, recompute: bool = False) -> str:
    """threedfront_dataset.py:279-287."""
    name = "CLIP_{}.pkl".format(scan_id) if large else "CLIP_small_{}.pkl".format(scan_id)
    path = os.path.join(root_3dfront, scan_id, name)
    return path + "tmp" if recompute else path


def write_clip_feats(path: str, instance_feats: np.ndarray, instance_order: Sequence[int], rel_feats: Dict[str, np.ndarray]) -> None:
    """threedfront_dataset.py:480-488: {'instance_feats': (n_instances + 1, 512) with the room's feature last,
    'instance_order': instance ids in the order the features were computed, 'rel_feats': {relation words: (512,)}}."""
    with open(path, "wb") as f:
        pickle.dump({"instance_feats": instance_feats, "instance_order": instance_order, "rel_feats": rel_feats}, f)


def read_clip_feats(path: str, instances_order: Sequence[int], words: Optional[Sequence[str]] = None):
    """threedfront_dataset.py:399-409 (+ :503-508 when `words` is given): the cached instance features re-ordered to this
    scene's `instances_order`, the room's feature appended last, and the relation features looked up by the relations'
    words.  Returns (text_feats: list of (512,) arrays, rel_feats: list of (512,) arrays or the raw dict when words is None)."""
    with open(path, "rb") as f:
        dic = pickle.load(f)
    feats = dic["instance_feats"]
    order = np.asarray(dic["instance_order"])
    ordered = [feats[:-1][inst == order] for inst in instances_order]
    ordered.append(feats[-1][np.newaxis, :])
    text_feats: List[np.ndarray] = list(np.concatenate(ordered, axis=0))
    rel = dic["rel_feats"]
    if words is None:
        return text_feats, rel
    return text_feats, [rel[w] for w in words]
