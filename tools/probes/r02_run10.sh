set -x
VX_BENCH_PIPE=3 python bench.py --device-only --steps 3 --warmup 3
VX_BENCH_PIPE=3 VX_EMIT_BG_SMEM=116736 python bench.py --device-only --steps 3 --warmup 3
VX_BENCH_PIPE=3 VX_EMIT_BG_SMEM=81920 python bench.py --device-only --steps 3 --warmup 3
VX_BENCH_PIPE=4 VX_EMIT_BG_SMEM=116736 python bench.py --device-only --steps 3 --warmup 3
