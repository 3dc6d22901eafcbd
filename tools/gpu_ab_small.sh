# A/B of the small-batch shapes under two environments.  bash tools/gpu_ab_small.sh <tag> "<env A>" "<env B>"
TAG=${1:-abs}; EA=${2:-}; EB=${3:-}
mkdir -p gpurun_out; P=gpurun_out/${TAG}
for V in A B; do
  if [ $V = A ]; then E="$EA"; else E="$EB"; fi
  for SHAPE in "224 1" "512 2"; do
    set -- $SHAPE
    env $E timeout 300 python bench.py --steps 3 --warmup 3 --res $1 --batch $2 --no-cpu-baseline --no-eager-baseline > ${P}_${V}_$1.json 2> ${P}_${V}_$1.err
    python -c "
import json; d=json.load(open('${P}_${V}_$1.json')); print('$V', '$E', 'res $1 batch $2', 'ms/step', d['ms_per_step']/100, 'img/s', d['value'], 'frac', d['roofline']['frac'])"
  done
done
