set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "short_list or scheduler or concurrent or thread or lane or cpp" 2>&1 | tail -3
timeout 300 tools/probes/pool_probe
