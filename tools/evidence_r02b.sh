# Round 2, second session: evidence of the final tree on ONE B200 (profiles/r02b_*), every step under a tight timeout.
mkdir -p gpurun_out; P=gpurun_out/r02b
timeout 300 python -m pytest tests -q -m gpu -p no:cacheprovider > ${P}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -2 ${P}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 ${P}_smoke.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file ${P}_smoke_launches_ncu.csv python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke_ncu.log 2>&1; echo "smoke ncu exit=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > ${P}_bench_reference.json 2> ${P}_ref.err; echo "ref exit=$?"
timeout 400 python bench.py --steps 5 --warmup 3 --dump-kernels ${P}_kernels_cuda_events.csv > ${P}_bench.json 2> ${P}_bench.err; echo "bench exit=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${P}_launches_ncu_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > ${P}_ncu_bench.log 2>&1; echo "ncu list exit=$?"
python -c "
import json; d=json.load(open('${P}_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_conv_kxk']['frac_of_burst'], d['clocks'], d['forward_breakdown_ms'])"
export IDIFF_PROFILE_STEPS=1
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'conv_gemm_kernel|conv3_rowpair_kernel' --csv --log-file ${P}_conv_dram.csv python tools/profile_forward.py > ${P}_ncu_dram.log 2>&1; echo "ncu dram exit=$?"
python tools/ncu_conv_traffic.py ${P}_conv_dram.csv "ncu capture of tools/profile_forward.py (B=32, 256x256, one forward), round 2 second session, final tree" && cp profiles/conv_dram_traffic.json ${P}_conv_dram_traffic.json; echo "traffic exit=$?"
IDIFF_LA_CASE=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:'la_out2_kernel' -s 2 -c 1 -f -o ${P}_la_out2 python tools/run_linattn.py > /dev/null 2>&1; echo "ncu la_out2 exit=$?"
timeout 200 ncu --set full --clock-control none -k regex:'chan_ln_gn_kernel' -c 1 -f -o ${P}_chan_ln_gn python tools/profile_forward.py > /dev/null 2>&1; echo "ncu chan_ln_gn exit=$?"
python tools/ncu_select.py ${P}_ncu_full_selected.csv ${P}_la_out2.ncu-rep ${P}_chan_ln_gn.ncu-rep > /dev/null; echo "select exit=$?"
python tools/ncu_lines.py ${P}_la_out2.ncu-rep 40 > ${P}_la_out2_ncu_source_hotlines.txt 2>&1
IDIFF_LIB_PATH=instancediff_b200/libidiff_prof.so IDIFF_LA_PROF=1 timeout 100 python tools/prof_linattn.py > ${P}_la_out2_role_cycles.txt 2>&1; echo "roles exit=$?"
timeout 60 python tools/run_linattn.py > ${P}_linattn_alone.txt 2>&1; IDIFF_LA_OUT_V1=1 timeout 60 python tools/run_linattn.py >> ${P}_linattn_alone.txt 2>&1
rm -f ${P}_conv_dram.csv ${P}_chan_ln_gn.ncu-rep
ls -la gpurun_out | grep r02b; du -sh gpurun_out
