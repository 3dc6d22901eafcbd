# la_out2 iteration loop: tests, timing, role counters, one full ncu capture of la_out2<64>
mkdir -p gpurun_out; P=gpurun_out/${1:-laq}
timeout 240 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "linattn or forward_layerwise" > ${P}_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 ${P}_pytest.log
timeout 120 python tools/run_linattn.py 2>&1 | tee ${P}_plain.log
IDIFF_LIB_PATH=instancediff_b200/libidiff_prof.so IDIFF_LA_PROF=1 timeout 120 python tools/prof_linattn.py 2>&1 | tee ${P}_roles.txt
#IDIFF_LA_CASE=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'la_out2_kernel' -s 2 -c 1 -f -o ${P}_out2 python tools/run_linattn.py > /dev/null 2>&1; echo "ncu out2 exit=$?"
#python tools/ncu_select.py ${P}_selected.csv ${P}_out2.ncu-rep > /dev/null; grep -E "gpu__time_duration|issue_active|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|bank_conflicts" ${P}_selected.csv | cut -d, -f3-5
