mkdir -p gpurun_out; P=gpurun_out/fin
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > ${P}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -2 ${P}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 ${P}_smoke.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_reference.json 2> ${P}_ref.err; echo "ref exit=$?"
timeout 400 python bench.py --dump-kernels ${P}_kernels_cuda_events.csv > ${P}_bench.json 2> ${P}_bench.err; echo "bench exit=$?"
IDIFF_LIB_PATH=instancediff_b200/libidiff_prof.so timeout 200 python tools/prof_layers.py > ${P}_role_cycles.txt 2>&1; echo "prof exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${P}_launches_ncu_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > ${P}_ncu_bench.log 2>&1; echo "ncu list exit=$?"
for L in c a b; do IDIFF_LAYER=$L timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 7 -c 1 -f -o ${P}_conv_$L python tools/run_layer.py > ${P}_ncu_$L.log 2>&1; echo "ncu $L exit=$?"; done
for K in la_out la_ctx la_merge; do IDIFF_LA_CASE=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o ${P}_$K python tools/run_linattn.py > ${P}_ncu_$K.log 2>&1; echo "ncu $K exit=$?"; done
for K in self_attention stem_tc rowwise_kernel sde_step head_conv3; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o ${P}_$K python tools/profile_forward.py > ${P}_ncu_$K.log 2>&1; echo "ncu $K exit=$?"; done
ls -la gpurun_out/fin_* | awk '{s+=$5} END {print s/1e6 " MB"}'
