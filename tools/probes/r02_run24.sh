set -x
tools/probes/pool_probe batch
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
