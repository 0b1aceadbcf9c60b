set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "frustum or scheduler" 2>&1 | tail -3
python tools/probes/frustum_ab.py
