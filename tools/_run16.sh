python -m pytest tests/test_modes.py tests/test_cabi.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload pm1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pm1', d['value'], d['unit'], 'ms_per_term', d['ms_per_term'], 'ms_per_mul', d['ms_per_mul'], 'terms', d['terms'])"
python bench.py --workload pm1 > gpurun_out/r2y_pm1.json 2>/dev/null
