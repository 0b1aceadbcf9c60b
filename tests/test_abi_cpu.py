"""CPU: the C-ABI library builds, loads and exports every symbol include/voxb200.h
declares; host-side logic that needs no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "voxb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(vox):
    L = C.CDLL(vox.SO_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export " + n
    # and the Python binding covers exactly the declared set
    assert sorted(vox._lib.SIGNATURES) == names


def test_version_and_defaults(vox):
    L = vox._lib.lib()
    assert L.vx_version() == 110
    p = vox.default_noise_params()
    # perlin_noise.rs:14-18
    assert (p.seed, p.octaves) == (1, 3)
    assert np.float32(p.freq) == np.float32(0.09) and p.gain == 2.5 and p.lacunarity == 0.5
    assert np.float32(p.simplex_scale) == np.float32(32.69587493801679)
    # the Rust mirror of the struct has the same fields in the same order
    rs = open(os.path.join(ROOT, "wgpu-voxel-engine_b200", "rust", "voxb200-sys", "src", "lib.rs")).read()
    body = rs.split("pub struct vx_noise_params {")[1].split("}")[0]
    assert re.findall(r"pub (\w+):", body) == [f for f, _ in type(p)._fields_]


def test_instance_layout(vox):
    # renderer/instance.rs:1-6: 16 f32 + u32, repr(C), 68 bytes
    d = vox.INSTANCE_DTYPE
    assert d.itemsize == 68 and d.fields["t_matrix"][1] == 0 and d.fields["texture_index"][1] == 64


def test_no_device_is_an_error_not_a_fallback(vox):
    L = vox._lib.lib()
    if L.vx_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(vox.VoxError) as e:
        vox.Context()
    assert e.value.code == vox._lib.VX_ERR_NO_DEVICE
    h = C.c_void_p()
    assert L.vx_create(None, 1, C.byref(h)) == vox._lib.VX_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.vx_last_error() or b"no CUDA device" in L.vx_last_error()


def test_planes_from_matrix_matches_oracle(vox, oracle, golden):
    # host bookkeeping (24 adds), frustum_culling.rs:47-85
    rng = np.random.default_rng(0)
    mats = [rng.standard_normal(16).astype(np.float32) for _ in range(8)]
    mats += [np.array([int(h, 16) for h in fx["matrix_colmajor"]], np.uint32).view(np.float32)
             for fx in golden["frustum"]]
    for m in mats:
        a = vox.planes_from_matrix(m)
        b = oracle.matrix_to_planes(m)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_coordinate_conventions(vox):
    c = vox.ChunkCoord3D(-3, 0, 7)
    assert c.to_world_position_i32() == (-96, 0, 224)                   # coordinate.rs:60-65
    assert vox.world_to_chunk_coord(-1, 31, 32) == (-1, 0, 1)           # coordinate.rs:13-18 (>> 5)
    assert vox.ChunkCoord3D(1 << 30, 0, 0).to_world_position_i32()[0] == 0   # wrapping i32 product
    assert sorted([vox.ChunkCoord3D(1, 0, 0), vox.ChunkCoord3D(0, 5, 0)])[0] == (0, 5, 0)   # Ord (x,y,z)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "wgpu-voxel-engine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile", ".rs")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "vox_oracle" not in text and "libvoxoracle" not in text and "import oracle" not in text \
                    and "from oracle" not in text, os.path.join(dirpath, f)


def test_rust_bindings_cover_the_header():
    # the Rust FFI crate cannot be compiled here (no rustc): keep it in lock-step with the header
    rs = open(os.path.join(ROOT, "wgpu-voxel-engine_b200", "rust", "voxb200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (vx_[a-z0-9_]+)\s*\(", rs))
    assert declared == set(_declared_symbols())


def test_cpp_host_mirror_compiles(vox, tmp_path):
    import subprocess
    host = os.path.join(ROOT, "wgpu-voxel-engine_b200", "host")
    exe = str(tmp_path / "host_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(host, "host_demo.cpp"),
                           "-L" + os.path.dirname(vox.SO_PATH), "-lvoxb200",
                           "-Wl,-rpath," + os.path.dirname(vox.SO_PATH), "-o", exe])
    if vox._lib.lib().vx_device_count() <= 0:
        r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stdout          # loud failure, no fallback
