for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N tools/bench_region_multigpu.py --region 1024 2>gpurun_out/region_$N.err | tail -1 > gpurun_out/region_$N.json
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_1.json
timeout 300 python tools/bench_region_multigpu.py --region 1024 2>/dev/null | tail -1 > gpurun_out/region_1.json
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.load(open("gpurun_out/scale_%d.json"%n)); r=json.load(open("gpurun_out/region_%d.json"%n))
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], "| region ms", r["ms_max_over_ranks"], r["chunks_per_s"], r["quads"])
    except Exception as e:
        print(n, "failed", e)
PY
