set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "frustum or scheduler or mirror" 2>&1 | tail -4
timeout 300 python tools/bench_configs.py 2>&1 | grep -i "frustum" | cut -c1-420
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/frustum_probe.py 1 2>/dev/null | grep -i "frustum" | cut -c100-300
timeout 120 ncu --set full --clock-control none --import-source on -k regex:vx_frustum_one -c 1 -o gpurun_out/r02e_frustum1 python tools/frustum_probe.py 1 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:vx_frustum_kernel -c 1 -o gpurun_out/r02e_frustum256 python tools/frustum_probe.py 256 > /dev/null 2>&1
