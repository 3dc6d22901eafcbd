# A/B on one B200: bench.py (short) under two environments.  bash tools/gpu_ab.sh <tag> "<env A>" "<env B>" [pytest -k expr]
TAG=${1:-ab}; EA=${2:-}; EB=${3:-}; K=${4:-}
mkdir -p gpurun_out; P=gpurun_out/${TAG}
if [ -n "$K" ]; then timeout 300 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "$K" > ${P}_pytest.log 2>&1; echo "pytest exit=$?"; tail -5 ${P}_pytest.log; fi
for V in A B; do
  if [ $V = A ]; then E="$EA"; else E="$EB"; fi
  env $E timeout 600 python bench.py --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline --no-eager-baseline --dump-kernels ${P}_${V}_kernels.csv > ${P}_${V}_bench.json 2> ${P}_${V}_bench.err; echo "bench $V ($E) exit=$?"
  python - <<PY
import json
try:
    d=json.load(open('${P}_${V}_bench.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'launches',d.get('gpu_launches'),d['clocks']['sm_mhz'])
    print(d['forward_breakdown_ms'], d['forward_ms_sum_of_kernels'])
except Exception as e:
    print('bench parse failed',e); print(open('${P}_${V}_bench.err').read()[-2000:])
PY
done
