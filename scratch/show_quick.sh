tag=$1
cat gpurun_out/${tag}_pytest.log | tail -3
for m in f16x3f f16; do python - gpurun_out/${tag}_bench_$m.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), 'e2e', round(d['e2e']['value']))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
