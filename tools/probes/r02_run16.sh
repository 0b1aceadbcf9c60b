set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "frustum or scheduler or device_resident" 2>&1 | tail -3
for f in 1 8 64 1000; do python tools/frustum_probe.py $f; done
timeout 600 python tools/bench_configs.py > gpurun_out/r02_configs.jsonl 2>&1; grep frustum gpurun_out/r02_configs.jsonl | cut -c1-330
timeout 120 ncu --set full --clock-control none --import-source on -k regex:vx_frustum_one -c 1 -o gpurun_out/r02_frustum1 python tools/frustum_probe.py 1 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:vx_frustum_kernel -c 1 -o gpurun_out/r02_frustum256 python tools/frustum_probe.py 256 > /dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
