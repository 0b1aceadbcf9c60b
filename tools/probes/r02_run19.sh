set -x
for v in g1t128 g1t64 g1t128u1 g1t128u3 g1t128u6 g1t256; do python tools/probes/frustum_ab.py tools/probes/build/libvox_$v.so; done
