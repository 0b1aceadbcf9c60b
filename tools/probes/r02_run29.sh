set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_properties.py -x -q -k "fill or simplex_scale or fast or ring or region or sign or fused" 2>&1 | tail -3
timeout 900 python tools/verify_sign_only.py --region 1024 --seeds 1,2,12345 2>&1 | tail -3 | cut -c1-300
timeout 600 python tools/verify_sign_only.py --region 1024 --seeds 1 --scale 1.0 2>&1 | tail -1 | cut -c1-300
timeout 600 python tools/verify_sign_only.py --region 512 --seeds 1 --origin 3000000,0,-3000000 2>&1 | tail -1 | cut -c1-300
python bench.py --device-only --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('kernel_ms'), d['roofline']['frac'])"
