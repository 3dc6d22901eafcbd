# (compute-sanitizer is closed on this GPU pool, so there is no memcheck leg.)
# Third-session evidence (profiles/r01s3_*): tests, smoke, both bench arms, per-kernel table, ncu launch list of the
# bench command and full captures of the two fused SDE kernels.  The conv / attention captures of r01s2 still describe
# the same kernel code (csrc/conv_gemm*, attention.cu, linattn_fused.cu unchanged since).
mkdir -p gpurun_out; P=gpurun_out/s3fin
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > ${P}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -2 ${P}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 ${P}_smoke.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_reference.json 2> ${P}_ref.err; echo "ref exit=$?"
timeout 400 python bench.py --dump-kernels ${P}_kernels_cuda_events.csv > ${P}_bench.json 2> ${P}_bench.err; echo "bench exit=$?"
timeout 200 python tools/test_um.py > ${P}_test_um.log 2>&1; echo "test_um exit=$?"; tail -3 ${P}_test_um.log
timeout 200 python tools/test_um.py --batch 4 > ${P}_test_um_batch4.log 2>&1; echo "test_um batch4 exit=$?"; tail -3 ${P}_test_um_batch4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${P}_launches_ncu_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > ${P}_ncu_bench.log 2>&1; echo "ncu list exit=$?"
for K in sde_step random_states; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o ${P}_$K python tools/profile_forward.py > ${P}_ncu_$K.log 2>&1; echo "ncu $K exit=$?"; done
python -c "
import json; d=json.load(open('${P}_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline_sde']['frac'], d['clocks'])"
