set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/bench_configs.py 2>&1 | grep -i "frustum" | cut -c1-420 | head -4
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/frustum_probe.py 1 2>/dev/null | grep -i "frustum" | cut -c1-300
timeout 120 ncu --set full --clock-control none --import-source on -k regex:vx_frustum_one -c 1 -o gpurun_out/r02d_frustum1 python tools/frustum_probe.py 1 > /dev/null 2>&1
