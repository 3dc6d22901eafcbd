# Round-2 evidence on ONE B200 (profiles/r02_*): tests, smoke (+ its launch list), both bench arms, per-kernel table, ncu
# launch list of the bench command, DRAM bytes of the conv launches, full captures of the dominant kernels (converted to
# CSV summaries on the box: gpurun_out/ may carry 64 MiB back).   bash tools/evidence_r02.sh [light|ncu|all]
MODE=${1:-all}
mkdir -p gpurun_out; P=gpurun_out/r02
if [ "$MODE" != "ncu" ]; then
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > ${P}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -2 ${P}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file ${P}_smoke_launches_ncu.csv python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke_ncu.log 2>&1; echo "smoke exit=$?"; tail -1 ${P}_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > ${P}_bench_reference.json 2> ${P}_ref.err; echo "ref exit=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --dump-kernels ${P}_kernels_cuda_events.csv > ${P}_bench.json 2> ${P}_bench.err; echo "bench exit=$?"
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > ${P}_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file ${P}_launches_ncu_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > ${P}_ncu_bench.log 2>&1; echo "ncu list exit=$?"
IDIFF_LIB_PATH=instancediff_b200/libidiff_prof.so timeout 200 python tools/prof_rowpair.py > ${P}_rowpair_role_cycles.txt 2>&1; echo "roles exit=$?"
python -c "
import json; d=json.load(open('${P}_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_conv_kxk']['frac_of_burst'], d['clocks'])"
fi
if [ "$MODE" != "light" ]; then
export IDIFF_PROFILE_STEPS=1
timeout 300 python tools/profile_forward.py > ${P}_pf.log 2>&1 && timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'conv_gemm_kernel|conv3_rowpair_kernel' --csv --log-file ${P}_conv_dram.csv python tools/profile_forward.py > ${P}_ncu_dram.log 2>&1; echo "ncu dram exit=$?"
python tools/ncu_conv_traffic.py ${P}_conv_dram.csv "ncu capture of tools/profile_forward.py (B=32, 256x256, one forward), round 2 final tree" && cp profiles/conv_dram_traffic.json ${P}_conv_dram_traffic.json; echo "traffic exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv3_rowpair_kernel' -s 1 -c 2 -f -o ${P}_rowpair python tools/profile_forward.py > ${P}_ncu_rowpair.log 2>&1; echo "ncu rowpair exit=$?"
# generic engine: one launch per N tile (64: 128->64 3x3 concat; 128: 192->128 3x3; 256: 384->256 3x3) and the hottest glue kernels
timeout 600 ncu --set full --clock-control none -k regex:'conv_gemm_kernel<64, 3, 0, 0>' -s 1 -c 1 -f -o ${P}_conv_nt64 python tools/profile_forward.py > ${P}_ncu_c64.log 2>&1; echo "ncu nt64 exit=$?"
timeout 600 ncu --set full --clock-control none -k regex:'conv_gemm_kernel<128, 3, 0, 0>' -s 1 -c 1 -f -o ${P}_conv_nt128 python tools/profile_forward.py > ${P}_ncu_c128.log 2>&1; echo "ncu nt128 exit=$?"
timeout 600 ncu --set full --clock-control none -k regex:'conv_gemm_kernel<256, 3, 0, 0>' -s 3 -c 1 -f -o ${P}_conv_nt256 python tools/profile_forward.py > ${P}_ncu_c256.log 2>&1; echo "ncu nt256 exit=$?"
timeout 600 ncu --set full --clock-control none -k regex:'la_out_kernel|la_ctx_kernel' -c 2 -f -o ${P}_linattn python tools/profile_forward.py > ${P}_ncu_la.log 2>&1; echo "ncu linattn exit=$?"
python tools/ncu_select.py ${P}_ncu_full_selected.csv ${P}_rowpair.ncu-rep ${P}_conv_nt64.ncu-rep ${P}_conv_nt128.ncu-rep ${P}_conv_nt256.ncu-rep ${P}_linattn.ncu-rep; echo "select exit=$?"
python tools/ncu_lines.py ${P}_rowpair.ncu-rep 40 > ${P}_rowpair_ncu_source_hotlines.txt 2>&1
rm -f ${P}_conv_nt64.ncu-rep ${P}_conv_nt128.ncu-rep ${P}_conv_nt256.ncu-rep ${P}_linattn.ncu-rep ${P}_conv_dram.csv
ls -la gpurun_out | head -40; du -sh gpurun_out
fi
