"""csrc/host_copy.hpp -- the copy threads and the page-residency probe behind downloads into pageable caller memory
(api.cu: d2h_pageable, use_bounce) -- compiled for the host and exercised without a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parallel_copier_and_residency_probe(tmp_path):
    exe = str(tmp_path / "host_copy_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread",
                           os.path.join(ROOT, "tests", "host_shim", "host_copy_test.cpp"), "-o", exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    assert "host_copy ok" in out.stdout
