set -x
python bench.py --device-only --steps 3 --warmup 3
VX_EMIT_BG_SMEM=116736 python bench.py --device-only --steps 3 --warmup 3
VX_EMIT_BG_SMEM=81920 python bench.py --device-only --steps 3 --warmup 3
VX_EMIT_BG_SMEM=57344 python bench.py --device-only --steps 3 --warmup 3
VX_EMIT_BG_SMEM=49152 python bench.py --device-only --steps 3 --warmup 3
