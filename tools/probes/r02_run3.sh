set -x
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "multi_gpu or frustum or mesh_edge or emit_again" 2>&1 | tail -4
timeout 300 python tools/bench_configs.py 2>&1 | grep -i "frustum" | cut -c1-200 | head -4
timeout 600 python bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02c_n2.json 2> gpurun_out/bench_r02c_n2.err; tail -c 1500 gpurun_out/bench_r02c_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02c_n2.json')); print(d['n_gpus'], d['value'], d['e2e'], d['parity'])"
timeout 300 python tools/bench_multidev_host.py --rings 8 --steps 5 2>&1 | tail -3
timeout 120 ncu --set full --clock-control none --import-source on -k regex:vx_frustum_one -s 3 -c 1 -o gpurun_out/r02c_frustum1 python tools/frustum_probe.py 1 > /dev/null 2>&1
