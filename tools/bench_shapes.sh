# Extra bench lines at the other BASELINE.json shapes that fit one GPU (not the headline).
mkdir -p gpurun_out; P=gpurun_out/${TAG:-s3}
timeout 120 python bench.py --no-cpu-baseline > ${P}_bench_quick.json 2> ${P}_bench_quick.err; echo "bench exit=$?"
timeout 150 python bench.py --res 512 --batch 16 --steps 2 --warmup 3 --no-cpu-baseline > ${P}_bench_512_b16.json 2> ${P}_bench_512_b16.err; echo "512b16 exit=$?"
timeout 100 python bench.py --res 512 --batch 2 --steps 3 --warmup 3 --no-cpu-baseline > ${P}_bench_512_b2.json 2> ${P}_bench_512_b2.err; echo "512b2 exit=$?"
timeout 200 python bench.py --res 256 --batch 128 --steps 2 --warmup 3 --no-cpu-baseline > ${P}_bench_256_b128.json 2> ${P}_bench_256_b128.err; echo "b128 exit=$?"
python - <<PY
import json
for f in ["bench_quick", "bench_512_b16", "bench_512_b2", "bench_256_b128"]:
    try:
        d = json.load(open("${P}_%s.json" % f))
        print(f, round(d["value"], 3), round(d["e2e"]["value"], 3), round(d["ms_per_step"], 1), round(d["roofline"]["frac"], 3),
              d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "ERR", e)
PY
for f in ${P}_bench_*.err; do tail -n 3 $f; done
