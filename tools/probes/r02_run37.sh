timeout 900 python -m pytest tests -m gpu -x -q -k "pageable or short or ring or scheduler or multi_gpu or fill" 2>&1 | tail -3
python tools/probes/bounce_threshold_probe.py
tools/probes/pool_probe batch | grep -v vx_mesh
python tools/probes/latency_probe.py 2>&1 | grep -E "^n=|vx_fill"
