set -x
for t in memcheck racecheck synccheck initcheck; do timeout 900 compute-sanitizer --tool $t python tools/sanitize_smoke.py 2>&1 | tail -4; done > gpurun_out/r02_sanitizer.txt 2>&1; cat gpurun_out/r02_sanitizer.txt
timeout 1500 python tools/verify_sign_only.py --region 1024 --seeds 1,2,12345 > gpurun_out/r02_verify_scaled.jsonl 2>&1; tail -3 gpurun_out/r02_verify_scaled.jsonl
timeout 600 python tools/verify_sign_only.py --region 1024 --seeds 1 --scale 1.0 > gpurun_out/r02_verify_raw.jsonl 2>&1; tail -1 gpurun_out/r02_verify_raw.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --region 256 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
