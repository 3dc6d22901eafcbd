# linear-attention kernels alone: launch list (v1 and v2 output pass) + full captures of la_out2<64> and la_ctx<64>
mkdir -p gpurun_out; P=gpurun_out/${1:-laprof}
python tools/run_linattn.py > ${P}_plain.log 2>&1; cat ${P}_plain.log
IDIFF_LA_OUT_V1=1 python tools/run_linattn.py > ${P}_plain_v1.log 2>&1; cat ${P}_plain_v1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'la_' --csv --log-file ${P}_launches.csv python tools/run_linattn.py > /dev/null 2>&1; echo "list exit=$?"
IDIFF_LA_OUT_V1=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'la_' --csv --log-file ${P}_launches_v1.csv python tools/run_linattn.py > /dev/null 2>&1; echo "list v1 exit=$?"
IDIFF_LA_CASE=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'la_out2_kernel' -s 2 -c 1 -f -o ${P}_out2 python tools/run_linattn.py > /dev/null 2>&1; echo "ncu out2 exit=$?"
IDIFF_LA_CASE=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'la_ctx_kernel' -s 2 -c 1 -f -o ${P}_ctx python tools/run_linattn.py > /dev/null 2>&1; echo "ncu ctx exit=$?"
python tools/ncu_select.py ${P}_selected.csv ${P}_out2.ncu-rep ${P}_ctx.ncu-rep; echo "select exit=$?"
python tools/ncu_lines.py ${P}_out2.ncu-rep 45 > ${P}_out2_hotlines.txt 2>&1
python tools/ncu_lines.py ${P}_ctx.ncu-rep 35 > ${P}_ctx_hotlines.txt 2>&1
ls -la gpurun_out | grep ${1:-laprof}
