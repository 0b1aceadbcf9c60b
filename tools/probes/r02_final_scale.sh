set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "multi_gpu" 2>&1 | tail -2
for n in 8 4 2 1; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err; tail -c 200 gpurun_out/r02_scale_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/r02_scale_n$n.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['d2h_gbs'], d['e2e']['link_peak_gbs'], d['parity'].get('sharded_equals_single_gpu'), d['parity']['region_digest'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_scale_ref_n8.json 2>/dev/null; cut -c1-300 gpurun_out/r02_scale_ref_n8.json
timeout 300 python tools/bench_multidev_host.py --rings 8 --steps 5 > gpurun_out/r02_multidev_host.jsonl 2>&1; cat gpurun_out/r02_multidev_host.jsonl | cut -c1-250
