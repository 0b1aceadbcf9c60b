set -x
for v in base sf sf128 t128 g1 sf128u1 sf128u3; do python tools/probes/frustum_ab.py tools/probes/build/libvox_$v.so; done
