"""Batch API over the C ABI: numpy in, numpy out (host buffers) plus raw
device-pointer variants.  Everything here forwards to libvoxb200.so; nothing
is computed in Python."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import NoiseParams, TerrainOpts, VoxError, check

VOXELS = 32768
ROWS = 1024
INSTANCE_DTYPE = np.dtype([("t_matrix", "<f4", (4, 4)), ("texture_index", "<u4")])   # renderer/instance.rs:1-6
assert INSTANCE_DTYPE.itemsize == 68


def default_noise_params(**kw) -> NoiseParams:
    p = NoiseParams()
    _lib.lib().vx_default_noise_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def _coords(c) -> np.ndarray:
    c = np.ascontiguousarray(c, dtype=np.int32)
    if c.ndim == 1:
        c = c.reshape(-1, 3)
    if c.ndim != 2 or c.shape[1] != 3:
        raise ValueError("coords must be int32 [n][3]")
    return c


@dataclass
class MeshBatch:
    """Per-chunk quad counts, exclusive offsets and the instance bytes (ChunkMesh buffers back to back)."""
    counts: np.ndarray
    offsets: np.ndarray
    instances: Optional[np.ndarray]
    total: int
    blocks: Optional[np.ndarray] = None

    def chunk(self, i: int) -> np.ndarray:
        return self.instances[int(self.offsets[i]):int(self.offsets[i + 1])]


class PinnedBuffer:
    """cudaHostAlloc'd bytes exposed as numpy (vx_host_alloc / vx_host_free)."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        check(_lib.lib().vx_host_alloc(C.byref(p), nbytes))
        self.ptr, self.nbytes = p.value, nbytes
        raw = (C.c_uint8 * max(nbytes, 1)).from_address(self.ptr)
        raw._vx_owner = self          # numpy views reach `raw` through .base: they keep the allocation alive
        self.array = np.ctypeslib.as_array(raw)[:nbytes]

    def view(self, dtype, count=None):
        a = self.array.view(dtype)
        return a if count is None else a[:count]

    def free(self):
        if self.ptr:
            _lib.lib().vx_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """vx_ctx: one or more GPUs; batch calls shard by contiguous index range."""

    def __init__(self, devices=None, n_devices: int = 1):
        L = _lib.lib()
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            check(L.vx_create(arr, len(devices), C.byref(h)))
        else:
            check(L.vx_create(None, n_devices, C.byref(h)))
        self._h = h
        self._L = L

    def close(self):
        if getattr(self, "_h", None):
            self._L.vx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    @property
    def num_devices(self) -> int:
        return self._L.vx_num_devices(self._h)

    @property
    def launch_count(self) -> int:
        return int(self._L.vx_launch_count(self._h))

    def sync(self):
        check(self._L.vx_sync(self._h))

    # ---- fill: Chunk::new / generate_terrain (chunk.rs:21-46) ----------------
    def fill(self, coords, params: Optional[NoiseParams] = None, blocks=True, noise=False,
             occupancy=False, census=False) -> dict:
        coords = _coords(coords)
        n = coords.shape[0]
        out = {}
        b = np.empty((n, VOXELS), np.uint8) if blocks else None
        f = np.empty((n, VOXELS), np.float32) if noise else None
        o = np.empty((n, ROWS), np.uint32) if occupancy else None
        nz = C.c_uint64(0)
        check(self._L.vx_fill(self._h, _ptr(coords), n, C.byref(params) if params else None,
                              _ptr(b), _ptr(f), _ptr(o), C.addressof(nz) if census else None))
        if blocks:
            out["blocks"] = b
        if noise:
            out["noise"] = f
        if occupancy:
            out["occupancy"] = o
        if census:
            out["near_zero"] = int(nz.value)
        return out

    # ---- mesh of given blocks: Chunk::create_mesh (chunk.rs:48-57) ------------
    def mesh(self, coords, blocks, nbr_idx=None, mesh_mask=None, instances=True,
             capacity: Optional[int] = None, out: Optional[np.ndarray] = None) -> MeshBatch:
        coords = _coords(coords)
        n = coords.shape[0]
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(n, VOXELS)
        ni = None if nbr_idx is None else np.ascontiguousarray(nbr_idx, dtype=np.int32).reshape(n, 4)
        mm = None if mesh_mask is None else np.ascontiguousarray(mesh_mask, dtype=np.uint8).reshape(n)

        def call(inst, cap, counts, offsets, total):
            return self._L.vx_mesh(self._h, _ptr(coords), n, _ptr(blocks), _ptr(ni), _ptr(mm), _ptr(counts),
                                   _ptr(offsets), _ptr(inst), cap, C.addressof(total))
        return self._run_mesh(call, n, instances, capacity, out)

    # ---- fused: the worker closure chunk_builder.rs:80-84 ---------------------
    def gen_mesh(self, coords, params: Optional[NoiseParams] = None, nbr_mask=None, instances=True,
                 blocks=False, capacity: Optional[int] = None, out: Optional[np.ndarray] = None,
                 vertical: bool = False, dirt_rule: bool = False) -> MeshBatch:
        """vertical / dirt_rule: the vx_gen_mesh_ex extension (mask bits 4, 5 = -y, +y; DIRT under cover)."""
        coords = _coords(coords)
        n = coords.shape[0]
        nm = None if nbr_mask is None else np.ascontiguousarray(nbr_mask, dtype=np.uint8).reshape(n)
        b = np.empty((n, VOXELS), np.uint8) if blocks else None
        opts = TerrainOpts(int(vertical), int(dirt_rule)) if (vertical or dirt_rule) else None

        def call(inst, cap, counts, offsets, total):
            if opts is not None:
                return self._L.vx_gen_mesh_ex(self._h, _ptr(coords), n, C.byref(params) if params else None,
                                              C.byref(opts), _ptr(nm), _ptr(counts), _ptr(offsets), _ptr(inst), cap,
                                              C.addressof(total), _ptr(b))
            return self._L.vx_gen_mesh(self._h, _ptr(coords), n, C.byref(params) if params else None, _ptr(nm),
                                       _ptr(counts), _ptr(offsets), _ptr(inst), cap, C.addressof(total), _ptr(b))
        res = self._run_mesh(call, n, instances, capacity, out)
        res.blocks = b
        return res

    def gen_mesh_packed(self, coords, params: Optional[NoiseParams] = None, nbr_mask=None,
                        out: Optional[np.ndarray] = None) -> MeshBatch:
        """Extension (not the reference wire format): one uint32 record per quad, see include/voxb200.h."""
        coords = _coords(coords)
        n = coords.shape[0]
        nm = None if nbr_mask is None else np.ascontiguousarray(nbr_mask, dtype=np.uint8).reshape(n)
        counts = np.zeros(n, np.uint32)
        offsets = np.zeros(n + 1, np.uint64)
        total = C.c_uint64(0)
        quads = out.view(np.uint32).reshape(-1) if out is not None else np.empty(max(1024, n * 8192), np.uint32)

        def call(q):
            return self._L.vx_gen_mesh_packed(self._h, _ptr(coords), n, C.byref(params) if params else None, _ptr(nm),
                                              _ptr(counts), _ptr(offsets), _ptr(q), q.shape[0], C.addressof(total))
        rc = call(quads)
        if rc == _lib.VX_ERR_CAPACITY and out is None:
            quads = np.empty(int(total.value), np.uint32)
            rc = call(quads)
        check(rc)
        t = int(total.value)
        return MeshBatch(counts, offsets, quads[:t], t)

    def _run_mesh(self, call, n, instances, capacity, out) -> MeshBatch:
        counts = np.zeros(n, np.uint32)
        offsets = np.zeros(n + 1, np.uint64)
        total = C.c_uint64(0)
        if not instances:
            check(call(None, 0, counts, offsets, total))
            return MeshBatch(counts, offsets, None, int(total.value))
        if out is not None:
            inst = out.view(INSTANCE_DTYPE).reshape(-1)
        else:
            cap = capacity if capacity is not None else max(1024, n * 8192)
            inst = np.empty(cap, INSTANCE_DTYPE)
        rc = call(inst, inst.shape[0], counts, offsets, total)
        if rc == _lib.VX_ERR_CAPACITY and out is None and capacity is None:
            inst = np.empty(int(total.value), INSTANCE_DTYPE)     # exact size known now
            rc = call(inst, inst.shape[0], counts, offsets, total)
        check(rc)
        t = int(total.value)
        return MeshBatch(counts, offsets, inst[:t], t)

    # ---- frustum: Frustum::contains (frustum_culling.rs:16-40) ----------------
    def frustum_cull(self, planes, coords, bits=False) -> np.ndarray:
        planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, 6, 4)
        coords = _coords(coords)
        F, n = planes.shape[0], coords.shape[0]
        if bits:
            out = np.zeros((F, (n + 31) // 32), np.uint32)
            check(self._L.vx_frustum_cull_bits(self._h, _ptr(planes), F, _ptr(coords), n, _ptr(out)))
        else:
            out = np.zeros((F, n), np.uint8)
            check(self._L.vx_frustum_cull(self._h, _ptr(planes), F, _ptr(coords), n, _ptr(out)))
        return out

    # ---- device-pointer variants (pointers are ints, e.g. tensor.data_ptr()) ---
    def fill_device(self, coords_ptr, n, params=None, blocks_ptr=None, noise_ptr=None, occ_ptr=None,
                    near_zero_ptr=None, stream=None):
        check(self._L.vx_fill_device(self._h, coords_ptr, n, C.byref(params) if params else None, blocks_ptr,
                                     noise_ptr, occ_ptr, near_zero_ptr, stream))

    def mesh_device(self, coords_ptr, n, blocks_ptr, counts_ptr, offsets_ptr, inst_ptr, capacity,
                    nbr_idx_ptr=None, mesh_mask_ptr=None, status_ptr=None, stream=None):
        check(self._L.vx_mesh_device(self._h, coords_ptr, n, blocks_ptr, nbr_idx_ptr, mesh_mask_ptr, counts_ptr,
                                     offsets_ptr, inst_ptr, capacity, status_ptr, stream))

    def gen_mesh_device(self, coords_ptr, n, counts_ptr, offsets_ptr, inst_ptr, capacity, params=None,
                        nbr_mask_ptr=None, status_ptr=None, blocks_ptr=None, stream=None):
        check(self._L.vx_gen_mesh_device(self._h, coords_ptr, n, C.byref(params) if params else None,
                                         nbr_mask_ptr, counts_ptr, offsets_ptr, inst_ptr, capacity, status_ptr,
                                         blocks_ptr, stream))

    def emit_again_device(self, coords_ptr, n, counts_ptr, offsets_ptr, inst_ptr, capacity, status_ptr=None,
                          stream=None):
        check(self._L.vx_emit_again_device(self._h, coords_ptr, n, counts_ptr, offsets_ptr, inst_ptr, capacity,
                                           status_ptr, stream))

    def frustum_cull_device(self, planes_ptr, n_frames, coords_ptr, n, visible_ptr=None, bits_ptr=None,
                            stream=None):
        check(self._L.vx_frustum_cull_device(self._h, planes_ptr, n_frames, coords_ptr, n, visible_ptr, bits_ptr,
                                             stream))

    KERNEL_KINDS = ("fill", "halo", "resolve", "count", "scan", "emit", "pack", "frustum")

    def set_timing(self, on: bool):
        check(self._L.vx_set_timing(self._h, int(on)))

    def get_timing(self) -> dict:
        """{kind: (total_ms, launches)} since the last call (CUDA events on the launching stream)."""
        ms = (C.c_double * 8)()
        ln = (C.c_uint64 * 8)()
        check(self._L.vx_get_timing(self._h, C.addressof(ms), C.addressof(ln)))
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(self.KERNEL_KINDS)}

    def measure_peak(self, kind: int) -> float:
        v = C.c_double(0)
        check(self._L.vx_measure_peak(self._h, kind, C.byref(v)))
        return float(v.value)


def planes_from_matrix(m) -> np.ndarray:
    """Frustum::matrix_to_planes (frustum_culling.rs:47-85); m column-major 4x4."""
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    out = np.empty((6, 4), np.float32)
    check(_lib.lib().vx_planes_from_matrix(_ptr(m), _ptr(out)))
    return out


_default: Optional[Context] = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context()
    return _default
