# scratch driver: gradient parity tests + short bench (+ per-kernel time of the contraction from a CUPTI timeline)
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_ref_parity.py tests/test_gpu_api.py -x -q -m gpu -k "grad or nlml or twin or plan or partition" 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['ms_per_step'], d['value'], d['roofline']['frac'])"
timeout 200 python tools/timeline.py 2>&1 | grep -i "grad_contract\|grid_tables\|grad_point\|grad_finish" | head
