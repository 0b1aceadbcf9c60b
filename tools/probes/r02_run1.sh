set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; tail -c 1500 gpurun_out/bench_r02a.err; head -c 6000 gpurun_out/bench_r02a.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r02a_ref.json 2>/dev/null; cat gpurun_out/bench_r02a_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --region 256 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vx_fill_lattice -s 2 -c 1 -o gpurun_out/r02a_fill python bench.py --region 256 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vx_mesh_emit -s 2 -c 1 -o gpurun_out/r02a_emit python bench.py --region 256 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 python tools/bench_configs.py > gpurun_out/r02a_configs.jsonl 2>&1; tail -4 gpurun_out/r02a_configs.jsonl
