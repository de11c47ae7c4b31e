# usage: bash tools/ws_probe.sh WS ...   (one LFM_GEMM_WS value per argument; scratch driver of the warp-specialised GEMM experiments)
export PYTHONUNBUFFERED=1
for cfg in "$@"; do
  export LFM_GEMM_WS=$cfg
  echo "=== WS=$cfg"
  if [ "$cfg" != "0" ] && [ -z "$SKIP_TESTS" ]; then
    timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense" 2>&1 | tail -2
  fi
  [ -z "$SKIP_SYRK" ] && LFM_GEMM_FORCE=3 timeout 120 python tools/syrk_probe.py 2>&1 | grep '"cold": 0' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('K', d['K'], 'm', d['m'], 'us %.1f' % d['us'], 'TF %.2f' % d['tflops'])"
  timeout 200 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
