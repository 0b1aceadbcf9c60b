set -x
n=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err; tail -c 300 gpurun_out/r02_scale_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/r02_scale_n$n.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['d2h_gbs'], d['e2e']['link_peak_gbs'], d['e2e_pageable_sample']['value'], d['parity'].get('sharded_equals_single_gpu'), d['parity']['region_digest'])"
