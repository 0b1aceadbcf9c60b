"""Where does the bounce-buffer path towards pageable memory start to pay?  vx_fill with block output (32 KiB per chunk)
into FRESH and into REUSED numpy memory for growing lists; run with VX_BOUNCE_MIN_KB=64 (always bounce) and
VX_BOUNCE_MIN_KB=100000000 (never)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import voxb200
L = voxb200._lib.lib()
ctx = voxb200.Context(devices=[0])
for n in (8, 16, 32, 64, 128, 256, 512, 1024):
    c = np.array([(i % 32, 0, i // 32) for i in range(n)], np.int32)
    reused = np.zeros(n * 32768, np.uint8)
    res = {}
    for kind in ("fresh", "reused"):
        best = 1e9
        for r in range(6):
            dst = np.empty(n * 32768, np.uint8) if kind == "fresh" else reused
            t0 = time.perf_counter()
            rc = L.vx_fill(ctx.handle, c.ctypes.data_as(C.c_void_p), n, None, dst.ctypes.data_as(C.c_void_p), None, None, None)
            dt = time.perf_counter() - t0
            assert rc == 0
            if r:
                best = min(best, dt)
        res[kind] = best
    print("min_kb=%s  %5d chunks %6.2f MB: fresh %7.0f us (%5.1f GB/s)  reused %7.0f us (%5.1f GB/s)" % (
        os.environ.get("VX_BOUNCE_MIN_KB", "default"), n, n * 32768 / 1e6, res["fresh"] * 1e6, n * 32768 / res["fresh"] / 1e9,
        res["reused"] * 1e6, n * 32768 / res["reused"] / 1e9))
