# Multi-GPU evidence (profiles/r02_*_{N}gpu.json): bash tools/evidence_r02_multi.sh N
N=$1; mkdir -p gpurun_out; P=gpurun_out/r02
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 > ${P}_bench_${N}gpu.json 2> ${P}_bench_${N}gpu.err; echo "weak exit=$?"
if [ "$N" != "8" ]; then timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 --global-batch 256 > ${P}_bench_gb256_${N}gpu.json 2> ${P}_bench_gb256_${N}gpu.err; echo "strong exit=$?"; fi
if [ "$N" = "8" ]; then timeout 900 $RUN bench.py --gpus $N --steps 5 --warmup 3 --res 512 --batch 2 > ${P}_bench_512_b2_${N}gpu.json 2> ${P}_bench_512_${N}gpu.err; echo "512 exit=$?"; fi
timeout 900 $RUN bench.py --gpus $N --mode train --steps 8 --warmup 3 > ${P}_train_${N}gpu.json 2> ${P}_train_${N}gpu.err; echo "train exit=$?"
timeout 900 $RUN bench.py --gpus $N --mode train --steps 8 --warmup 3 --comm-bf16 > ${P}_train_bf16wire_${N}gpu.json 2> ${P}_train_bf16_${N}gpu.err; echo "train bf16 exit=$?"
for f in ${P}_*_${N}gpu.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], round(d['value'],2), d['unit'], 'ms/step', round(d['ms_per_step'],2), d.get('collective') and {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['collective'].items() if k in ('allreduce_alone_ms','bus_gbs','overlap_fraction','exposed_ms')})
except Exception as e: print(sys.argv[1], 'unreadable', e)
PY
done
