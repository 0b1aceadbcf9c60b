set -x
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "multi_gpu" 2>&1 | tail -3
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_r02_n$n.json 2> gpurun_out/bench_r02_n$n.err; tail -c 300 gpurun_out/bench_r02_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_n$n.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['d2h_gbs'], d['e2e']['link_peak_gbs'], d['parity']['sharded_equals_single_gpu'], d['parity']['oracle_sample'])"
done
timeout 300 python tools/bench_multidev_host.py --rings 8 --steps 5 2>&1 | tail -4
