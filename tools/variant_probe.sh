# scratch: A/B of prebuilt library variants (build_variants/liblfm_<tag>.so copied over the product .so on the box only)
export PYTHONUNBUFFERED=1
cp dis_project_b200/liblfm_b200.so /tmp/orig.so
for tag in "$@"; do
  cp build_variants/liblfm_$tag.so dis_project_b200/liblfm_b200.so
  echo "=== $tag"
  timeout 200 python bench.py --steps 30 --warmup 3 --no-secondary --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['ms_per_step'], d['value'], d['roofline']['frac'])"
  timeout 200 python tools/timeline.py > /dev/null 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/timeline_eager.json'))
for e in d:
    if 'grad_contract_kernel<true>' in e[0]: print('contract us', e[3])
PY
done
cp /tmp/orig.so dis_project_b200/liblfm_b200.so
