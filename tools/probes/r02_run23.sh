set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "mesh or scheduler or chunk or cpp or concurrent" 2>&1 | tail -5
timeout 300 tools/probes/pool_probe
