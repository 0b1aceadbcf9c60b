set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "frustum or scheduler or device_resident" 2>&1 | tail -3
for v in default t64 t256 nosf b6 b3; do python tools/probes/frustum_ab.py tools/probes/build/libvox_$v.so; done
