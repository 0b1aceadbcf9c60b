"""Loader + ctypes signatures for libvoxb200.so (include/voxb200.h).

The library is built in-tree by csrc/Makefile (nvcc, sm_100a).  There is no
CPU fallback: if the library cannot be built/loaded, or no CUDA device is
present, compute calls raise VoxError.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libvoxb200.so")
CSRC = os.path.join(_HERE, "csrc")

VX_OK, VX_ERR_BAD_ARG, VX_ERR_CAPACITY, VX_ERR_CUDA = 0, -1, -2, -3
VX_ERR_NO_DEVICE, VX_ERR_OOM, VX_ERR_BAD_BLOCK = -4, -5, -6
_CODES = {-1: "VX_ERR_BAD_ARG", -2: "VX_ERR_CAPACITY", -3: "VX_ERR_CUDA", -4: "VX_ERR_NO_DEVICE",
          -5: "VX_ERR_OOM", -6: "VX_ERR_BAD_BLOCK"}


class VoxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s: %s" % (_CODES.get(code, code), msg))
        self.code = code


class NoiseParams(C.Structure):
    """perlin_noise.rs:14-18"""
    _fields_ = [("seed", C.c_int32), ("freq", C.c_float), ("gain", C.c_float),
                ("lacunarity", C.c_float), ("octaves", C.c_int32), ("simplex_scale", C.c_float)]


class TerrainOpts(C.Structure):
    """vx_terrain_opts (extension)"""
    _fields_ = [("vertical_neighbours", C.c_uint32), ("dirt_rule", C.c_uint32)]


def build(force: bool = False) -> str:
    """Compile libvoxb200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", "Makefile"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "voxb200.h"))
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libvoxb200.so failed:\n" + r.stdout[-4000:])
    return SO_PATH


_lib = None

# name -> (restype, argtypes); every symbol include/voxb200.h declares
_vp, _sz, _i = C.c_void_p, C.c_size_t, C.c_int
SIGNATURES = {
    "vx_version": (_i, []),
    "vx_last_error": (C.c_char_p, []),
    "vx_device_count": (_i, []),
    "vx_create": (_i, [_vp, _i, C.POINTER(_vp)]),
    "vx_destroy": (None, [_vp]),
    "vx_num_devices": (_i, [_vp]),
    "vx_default_noise_params": (None, [C.POINTER(NoiseParams)]),
    "vx_launch_count": (C.c_uint64, [_vp]),
    "vx_host_alloc": (_i, [C.POINTER(_vp), _sz]),
    "vx_host_free": (_i, [_vp]),
    "vx_fill": (_i, [_vp, _vp, _sz, C.POINTER(NoiseParams), _vp, _vp, _vp, _vp]),
    "vx_mesh": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vx_gen_mesh": (_i, [_vp, _vp, _sz, C.POINTER(NoiseParams), _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "vx_gen_mesh_ex": (_i, [_vp, _vp, _sz, C.POINTER(NoiseParams), C.POINTER(TerrainOpts), _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "vx_gen_mesh_ex_device": (_i, [_vp, _vp, _sz, C.POINTER(NoiseParams), C.POINTER(TerrainOpts), _vp, _vp, _vp, _vp, _sz,
                                   _vp, _vp, _vp]),
    "vx_gen_mesh_packed": (_i, [_vp, _vp, _sz, C.POINTER(NoiseParams), _vp, _vp, _vp, _vp, _sz, _vp]),
    "vx_gen_mesh_packed_device": (_i, [_vp, _vp, _sz, C.POINTER(NoiseParams), _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "vx_planes_from_matrix": (_i, [_vp, _vp]),
    "vx_frustum_cull": (_i, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "vx_frustum_cull_bits": (_i, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "vx_fill_device": (_i, [_vp, _vp, _sz, C.POINTER(NoiseParams), _vp, _vp, _vp, _vp, _vp]),
    "vx_mesh_device": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "vx_gen_mesh_device": (_i, [_vp, _vp, _sz, C.POINTER(NoiseParams), _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "vx_emit_again_device": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp]),
    "vx_frustum_cull_device": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp]),
    "vx_sync": (_i, [_vp]),
    "vx_measure_peak": (_i, [_vp, _i, C.POINTER(C.c_double)]),
    "vx_set_timing": (_i, [_vp, _i]),
    "vx_get_timing": (_i, [_vp, _vp, _vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            build()
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)       # AttributeError = missing export
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int) -> int:
    if code < 0:
        raise VoxError(code, (lib().vx_last_error() or b"").decode("utf-8", "replace"))
    return code
