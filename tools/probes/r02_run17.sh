set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "frustum" 2>&1 | tail -3
for v in b4 b8m b2 b6 b4u6 b4u2 b8u2; do python tools/probes/frustum_ab.py tools/probes/build/libvox_$v.so; done
