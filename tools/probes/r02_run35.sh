timeout 900 python -m pytest tests -m gpu -x -q -k "short or scheduler or pageable or cpp or concurrent or thread" 2>&1 | tail -3
python tools/probes/latency_probe.py 2>&1 | head -8
