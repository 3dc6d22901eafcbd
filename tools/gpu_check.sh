# One quick B200 call: GPU tests (optionally a subset: K="expr"), then a short bench of the headline config with the
# per-kernel table.  bash tools/gpu_check.sh <tag> [pytest -k expression]
TAG=${1:-chk}; K=${2:-}
mkdir -p gpurun_out; P=gpurun_out/${TAG}
if [ -n "$K" ]; then timeout 300 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "$K" > ${P}_pytest.log 2>&1
else timeout 300 python -m pytest tests -q -x -m gpu -p no:cacheprovider > ${P}_pytest.log 2>&1; fi
echo "pytest exit=$?"; tail -15 ${P}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline --dump-kernels ${P}_kernels.csv > ${P}_bench.json 2> ${P}_bench.err; echo "bench exit=$?"
python - <<PY
import json
try:
    d=json.load(open('${P}_bench.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'launches',d.get('gpu_launches'),d['clocks'])
    print(d['forward_breakdown_ms'], d['forward_ms_sum_of_kernels'])
except Exception as e:
    print('bench parse failed',e); print(open('${P}_bench.err').read()[-2000:])
PY
