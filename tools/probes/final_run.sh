set -x
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_fill_lattice -c 1 -o gpurun_out/final_fill python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 900 python tools/verify_sign_only.py --region 1024 --seeds 1,2,12345 > gpurun_out/verify.jsonl 2>&1; tail -3 gpurun_out/verify.jsonl
for t in memcheck racecheck synccheck; do timeout 600 compute-sanitizer --tool $t python tools/sanitize_smoke.py 2>&1 | tail -4; done > gpurun_out/sanitizer.txt 2>&1; cat gpurun_out/sanitizer.txt
timeout 600 python tools/bench_configs.py > gpurun_out/configs.jsonl 2>&1; tail -2 gpurun_out/configs.jsonl
