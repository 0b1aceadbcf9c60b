set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "pageable or short or ring or multi_gpu or sharded" 2>&1 | tail -3
tools/probes/pool_probe batch
for t in memcheck racecheck; do timeout 900 compute-sanitizer --tool $t python tools/sanitize_smoke.py 2>&1 | tail -3; done
