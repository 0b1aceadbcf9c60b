set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b14.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/b14.json')); print(d['value'], d['one_stream']['value'], d['kernel_ms'], d['roofline_mesh']['frac'], d['roofline_mesh']['steady_state']['frac'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vx_mesh_emit -s 2 -c 1 -o gpurun_out/r02_emit python bench.py --region 256 --steps 1 --warmup 3 --device-only > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -3
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py 2>&1 | tail -3
