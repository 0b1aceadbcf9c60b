"""ctypes front-end for the CPU oracle (oracle/vox_oracle.c).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product
package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvoxoracle.so")

VOXELS = 32768
GRASS, DIRT, AIR = 0, 1, 2
LEFT, RIGHT, BACK, FRONT, DOWN, UP = range(6)


class NoiseParams(C.Structure):
    _fields_ = [("seed", C.c_int32), ("freq", C.c_float), ("gain", C.c_float),
                ("lacunarity", C.c_float), ("octaves", C.c_int32), ("simplex_scale", C.c_float)]


QUAD_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4"), ("face", "<u4"), ("tex", "<u4")])
INSTANCE_DTYPE = np.dtype([("t_matrix", "<f4", (4, 4)), ("texture_index", "<u4")])
assert INSTANCE_DTYPE.itemsize == 68


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vox_oracle.c")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("vox_oracle.c", "vox_oracle.h") if os.path.exists(os.path.join(_HERE, f)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    assert os.path.exists(_SO), src
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, i32p = C.c_void_p, C.c_size_t, C.c_void_p
        L.vxo_default_params.argtypes = [C.POINTER(NoiseParams)]
        L.vxo_simplex_3d.restype = C.c_float
        L.vxo_simplex_3d.argtypes = [C.c_float] * 3 + [C.c_int32]
        L.vxo_fbm_3d.restype = C.c_float
        L.vxo_fbm_3d.argtypes = [C.c_float] * 5 + [C.c_int32, C.c_int32, C.c_float]
        L.vxo_noise_chunk.argtypes = [i32p, C.POINTER(NoiseParams), vp]
        L.vxo_noise_chunk_simd.argtypes = [i32p, C.POINTER(NoiseParams), vp]
        L.vxo_have_avx2.restype = C.c_int
        L.vxo_terrain_from_noise.argtypes = [vp, vp]
        L.vxo_generate_terrain.argtypes = [i32p, C.POINTER(NoiseParams), vp]
        L.vxo_cull_unseen_triangles.restype = sz
        L.vxo_cull_unseen_triangles.argtypes = [i32p, vp, C.POINTER(vp), vp, sz]
        L.vxo_quad_to_raw.argtypes = [vp, vp]
        L.vxo_face_rotation.argtypes = [C.c_uint32, vp]
        L.vxo_create_mesh.restype = sz
        L.vxo_create_mesh.argtypes = [i32p, vp, C.POINTER(vp), vp, sz]
        L.vxo_matrix_to_planes.argtypes = [vp, vp]
        L.vxo_frustum_contains.restype = C.c_int
        L.vxo_frustum_contains.argtypes = [vp, i32p]
        L.vxo_frustum_cull.argtypes = [vp, sz, i32p, sz, vp]
        L.vxo_gen_mesh_batch.restype = sz
        L.vxo_gen_mesh_batch.argtypes = [i32p, sz, C.POINTER(NoiseParams), vp, C.c_int, C.c_int,
                                         vp, vp, vp, sz, vp]
        L.vxo_gen_mesh_batch_ex.restype = sz
        L.vxo_gen_mesh_batch_ex.argtypes = [i32p, sz, C.POINTER(NoiseParams), vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                            vp, vp, vp, sz, vp]
        L.vxo_dirt_rule.argtypes = [vp, vp]
        L.vxo_create_mesh_ex.restype = sz
        L.vxo_create_mesh_ex.argtypes = [i32p, vp, C.POINTER(vp), vp, sz]
        L.vxo_mesh_batch.restype = sz
        L.vxo_mesh_batch.argtypes = [i32p, sz, vp, vp, C.c_int, vp, vp, vp, sz]
        _lib = L
    return _lib


def default_params(**kw) -> NoiseParams:
    p = NoiseParams()
    lib().vxo_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _coord(c):
    return np.ascontiguousarray(np.asarray(c, dtype=np.int32).reshape(3))


def noise_chunk(coord, params=None, simd=False) -> np.ndarray:
    p = params or default_params()
    out = np.empty(VOXELS, dtype=np.float32)
    c = _coord(coord)
    (lib().vxo_noise_chunk_simd if simd else lib().vxo_noise_chunk)(_ptr(c), C.byref(p), _ptr(out))
    return out


def terrain_from_noise(noise) -> np.ndarray:
    noise = np.ascontiguousarray(noise, dtype=np.float32)
    out = np.empty(VOXELS, dtype=np.uint8)
    lib().vxo_terrain_from_noise(_ptr(noise), _ptr(out))
    return out


def generate_terrain(coord, params=None) -> np.ndarray:
    p = params or default_params()
    out = np.empty(VOXELS, dtype=np.uint8)
    c = _coord(coord)
    lib().vxo_generate_terrain(_ptr(c), C.byref(p), _ptr(out))
    return out


def _adj_array(adj):
    arr = (C.c_void_p * 4)()
    keep = []
    for k in range(4):
        a = adj[k] if adj is not None else None
        if a is None:
            arr[k] = None
        else:
            a = np.ascontiguousarray(a, dtype=np.uint8).reshape(VOXELS)
            keep.append(a)
            arr[k] = a.ctypes.data
    return arr, keep


def cull_unseen_triangles(coord, cubes, adj=None) -> np.ndarray:
    cubes = np.ascontiguousarray(cubes, dtype=np.uint8).reshape(VOXELS)
    arr, keep = _adj_array(adj)
    out = np.empty(98304, dtype=QUAD_DTYPE)
    c = _coord(coord)
    n = lib().vxo_cull_unseen_triangles(_ptr(c), _ptr(cubes), arr, _ptr(out), 98304)
    return out[:n].copy()


def create_mesh(coord, cubes, adj=None) -> np.ndarray:
    cubes = np.ascontiguousarray(cubes, dtype=np.uint8).reshape(VOXELS)
    arr, keep = _adj_array(adj)
    out = np.empty(98304, dtype=INSTANCE_DTYPE)
    c = _coord(coord)
    n = lib().vxo_create_mesh(_ptr(c), _ptr(cubes), arr, _ptr(out), 98304)
    return out[:n].copy()


def face_rotation(face: int) -> np.ndarray:
    R = np.empty((3, 3), dtype=np.float32)
    lib().vxo_face_rotation(face, _ptr(R))
    return R


def matrix_to_planes(m) -> np.ndarray:
    """m: 4x4 float32, column-major storage m[c][r] (nalgebra `data.0`)."""
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    out = np.empty((6, 4), dtype=np.float32)
    lib().vxo_matrix_to_planes(_ptr(m), _ptr(out))
    return out


def frustum_contains(planes, coord) -> bool:
    planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(24)
    c = _coord(coord)
    return bool(lib().vxo_frustum_contains(_ptr(planes), _ptr(c)))


def frustum_cull(planes, coords) -> np.ndarray:
    planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, 24)
    coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
    out = np.empty((planes.shape[0], coords.shape[0]), dtype=np.uint8)
    lib().vxo_frustum_cull(_ptr(planes), planes.shape[0], _ptr(coords), coords.shape[0], _ptr(out))
    return out


def gen_mesh_batch(coords, params=None, nbr_mask=None, nthreads=None, use_simd=True,
                   want_instances=True, want_cubes=False, vertical=False, dirt_rule=False):
    """Returns dict(counts, offsets, instances|None, cubes|None, total).
    vertical / dirt_rule: the SURVEY 8(f4) extension (vxo_gen_mesh_batch_ex)."""
    if vertical or dirt_rule:
        return _gen_mesh_batch_ex(coords, params, nbr_mask, nthreads, use_simd, want_instances, want_cubes,
                                  vertical, dirt_rule)
    p = params or default_params()
    coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
    n = coords.shape[0]
    nthreads = nthreads or os.cpu_count() or 1
    counts = np.zeros(n, dtype=np.uint32)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    cubes = np.empty((n, VOXELS), dtype=np.uint8) if want_cubes else None
    mask = None if nbr_mask is None else np.ascontiguousarray(nbr_mask, dtype=np.uint8)
    if not want_instances:
        total = lib().vxo_gen_mesh_batch(_ptr(coords), n, C.byref(p), _ptr(mask), nthreads,
                                         int(use_simd), _ptr(counts), _ptr(offsets), None, 0,
                                         _ptr(cubes))
        inst = None
    else:
        # two calls would double the work; size from a counting pass on small
        # inputs only (tests) -- benches use want_instances=False.
        total = lib().vxo_gen_mesh_batch(_ptr(coords), n, C.byref(p), _ptr(mask), nthreads,
                                         int(use_simd), _ptr(counts), _ptr(offsets), None, 0,
                                         _ptr(cubes))
        inst = np.empty(int(total), dtype=INSTANCE_DTYPE)
        total = lib().vxo_gen_mesh_batch(_ptr(coords), n, C.byref(p), _ptr(mask), nthreads,
                                         int(use_simd), _ptr(counts), _ptr(offsets), _ptr(inst),
                                         int(total), _ptr(cubes))
    if total == C.c_size_t(-1).value:
        raise RuntimeError("oracle batch failed")
    return dict(counts=counts, offsets=offsets, instances=inst, cubes=cubes, total=int(total))


def _gen_mesh_batch_ex(coords, params, nbr_mask, nthreads, use_simd, want_instances, want_cubes, vertical, dirt_rule):
    p = params or default_params()
    coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
    n = coords.shape[0]
    nthreads = nthreads or os.cpu_count() or 1
    counts = np.zeros(n, dtype=np.uint32)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    cubes = np.empty((n, VOXELS), dtype=np.uint8) if want_cubes else None
    mask = None if nbr_mask is None else np.ascontiguousarray(nbr_mask, dtype=np.uint8)

    def call(inst, cap):
        return lib().vxo_gen_mesh_batch_ex(_ptr(coords), n, C.byref(p), _ptr(mask), int(vertical), int(dirt_rule),
                                           nthreads, int(use_simd), _ptr(counts), _ptr(offsets), _ptr(inst), cap,
                                           _ptr(cubes))
    total = call(None, 0)
    inst = None
    if want_instances:
        inst = np.empty(int(total), dtype=INSTANCE_DTYPE)
        total = call(inst, int(total))
    if total == C.c_size_t(-1).value:
        raise RuntimeError("oracle batch failed")
    return dict(counts=counts, offsets=offsets, instances=inst, cubes=cubes, total=int(total))


def dirt_rule(cubes, above=None) -> np.ndarray:
    out = np.ascontiguousarray(cubes, dtype=np.uint8).reshape(VOXELS).copy()
    ab = None if above is None else np.ascontiguousarray(above, dtype=np.uint8).reshape(VOXELS)
    lib().vxo_dirt_rule(_ptr(out), _ptr(ab))
    return out


def mesh_batch(coords, cubes, nbr_idx, nthreads=None, want_instances=True):
    coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
    n = coords.shape[0]
    cubes = np.ascontiguousarray(cubes, dtype=np.uint8)
    nbr_idx = np.ascontiguousarray(nbr_idx, dtype=np.int32).reshape(n, 4)
    nthreads = nthreads or os.cpu_count() or 1
    counts = np.zeros(n, dtype=np.uint32)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    total = lib().vxo_mesh_batch(_ptr(coords), n, _ptr(cubes), _ptr(nbr_idx), nthreads,
                                 _ptr(counts), _ptr(offsets), None, 0)
    inst = None
    if want_instances:
        inst = np.empty(int(total), dtype=INSTANCE_DTYPE)
        total = lib().vxo_mesh_batch(_ptr(coords), n, _ptr(cubes), _ptr(nbr_idx), nthreads,
                                     _ptr(counts), _ptr(offsets), _ptr(inst), int(total))
    return dict(counts=counts, offsets=offsets, instances=inst, total=int(total))
