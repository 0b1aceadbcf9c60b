set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "mesh or ring or region or vertical or halo or sharded" 2>&1 | tail -3
python bench.py --region 512 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('kernel_ms'), d['one_stream']['value'])"
python tools/probes/emit_ab.py
