import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import voxb200
ctx = voxb200.Context(devices=[0])
xs = np.arange(-16, 16, dtype=np.int32)
z, x = np.meshgrid(xs, xs, indexing="ij")
coords = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)
fast = ctx.fill(coords, blocks=True, occupancy=True)
slow = ctx.fill(coords, blocks=True, noise=True)
d = fast["blocks"] != slow["blocks"]
print("mismatching voxels", int(d.sum()), "of", d.size, "chunks affected", int(d.any(axis=1).sum()))
if d.any():
    ci, vi = np.nonzero(d)
    yy, zz, xx = vi // 1024, (vi // 32) % 32, vi % 32
    print("y hist", np.bincount(yy, minlength=32))
    print("y%4 hist", np.bincount(yy % 4, minlength=4), "x%4", np.bincount(xx % 4, minlength=4), "z%4", np.bincount(zz % 4, minlength=4))
    n = slow["noise"][ci, vi]
    print("abs noise at mismatches: min %.3g median %.3g max %.3g" % (np.abs(n).min(), np.median(np.abs(n)), np.abs(n).max()))
    for k in range(min(8, len(ci))):
        print(coords[ci[k]], int(xx[k]), int(yy[k]), int(zz[k]), "fast", fast["blocks"][ci[k], vi[k]], "slow", slow["blocks"][ci[k], vi[k]], "noise", n[k])
