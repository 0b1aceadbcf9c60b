set -x
for z in 1 2 4; do VX_FILL_ZB=$z python tools/probes/fill_ab.py; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fill or simplex_scale or fast or ring or region" 2>&1 | tail -3
timeout 600 python tools/verify_sign_only.py --region 1024 --seeds 1,2 2>&1 | tail -2 | cut -c1-300
python bench.py --device-only --steps 3 --warmup 3
