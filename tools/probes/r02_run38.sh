timeout 100 python -m pytest tests -m gpu -x -q -k "frustum" 2>&1 | tail -2
python tools/probes/frustum_ab.py tools/probes/build/libvox_pair0.so
python tools/probes/frustum_ab.py tools/probes/build/libvox_pair1.so
