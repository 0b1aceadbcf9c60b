set -x
export CUDA_DEVICE_MAX_CONNECTIONS=32
python bench.py --device-only --steps 3 --warmup 3
VX_EMIT_BG_SMEM=49152 python bench.py --device-only --steps 3 --warmup 3
VX_EMIT_BG_SMEM=81920 python bench.py --device-only --steps 3 --warmup 3
VX_EMIT_BG_SMEM=116736 python bench.py --device-only --steps 3 --warmup 3
VX_BENCH_PIPE=3 VX_EMIT_BG_SMEM=116736 python bench.py --device-only --steps 3 --warmup 3
