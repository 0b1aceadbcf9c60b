set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --device-only --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('kernel_ms'), d.get('one_stream'))"
