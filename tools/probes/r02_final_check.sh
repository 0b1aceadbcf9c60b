set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 300 gpurun_out/r02_bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_default.json')); print(d['value'], d['steps'], d['ms_per_step'], d['e2e']['value'], d['e2e']['link_frac'], d['e2e_pageable_sample']['value'], d['roofline']['frac'], d['roofline_mesh']['frac'], d['gpu_launches'], d['clocks']['reasons'])"
python bench.py --impl reference | cut -c1-300
