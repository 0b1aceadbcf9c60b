set -x
for v in 0 1 2; do VX_FILL_VARIANT=$v python tools/probes/fill_ab.py; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fill or simplex_scale or fast" 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -c 800 gpurun_out/bench_r02b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02b.json')); print(d['value'], d['one_stream']['value'], d['kernel_ms'], d['roofline']['frac'], d['roofline_mesh']['frac'], d['parity'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vx_fill_lattice2 -s 2 -c 1 -o gpurun_out/r02b_fill python bench.py --region 256 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for F in 1 1 8; do timeout 100 python tools/frustum_probe.py $F 2>&1 | tail -1; done
