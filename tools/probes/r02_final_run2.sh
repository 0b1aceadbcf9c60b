# final evidence set of round 2 on the tree with the extreme-corner frustum kernels (one B200)
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 400 gpurun_out/r02_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['link_frac'], d['kernel_ms'], d['roofline']['frac'], d['roofline_mesh']['frac'], d['roofline_mesh']['steady_state']['frac'], d['clocks'], d['parity']['oracle_sample'])"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --region 256 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:vx_frustum_one -c 1 -o gpurun_out/r02_frustum1 python tools/frustum_probe.py 1 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:vx_frustum_kernel -c 1 -o gpurun_out/r02_frustum256 python tools/frustum_probe.py 256 > /dev/null 2>&1
timeout 600 python tools/bench_configs.py > gpurun_out/r02_configs.jsonl 2>&1; tail -3 gpurun_out/r02_configs.jsonl | cut -c1-250
python tools/probes/frustum_ab.py
for t in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $t python tools/sanitize_smoke.py 2>&1 | tail -3; done > gpurun_out/r02_sanitizer.txt 2>&1; cat gpurun_out/r02_sanitizer.txt
