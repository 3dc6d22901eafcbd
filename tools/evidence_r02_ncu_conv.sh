mkdir -p gpurun_out; P=gpurun_out/r02
export IDIFF_PROFILE_STEPS=1
timeout 300 python tools/profile_forward.py > ${P}_pf.log 2>&1 || exit 1
i=0
for spec in "7 c256_32" "41 c384to256_64" "50 c192to128_128" "57 c128to64_256" "58 c1x1_128to64_256"; do set -- $spec
timeout 600 ncu --set full --clock-control none -k regex:'conv_gemm_kernel' -s $1 -c 1 -f -o ${P}_conv_$2 python tools/profile_forward.py > ${P}_ncu_$2.log 2>&1; echo "ncu $2 exit=$?"; done
timeout 600 ncu --set full --clock-control none -k regex:'self_attention_kernel|block_tail|stem_tc|head_conv3|sde_step' -c 6 -f -o ${P}_glue python tools/profile_forward.py > ${P}_ncu_glue.log 2>&1; echo "ncu glue exit=$?"
python tools/ncu_select.py ${P}_ncu_full_selected_conv.csv ${P}_conv_c256_32.ncu-rep ${P}_conv_c384to256_64.ncu-rep ${P}_conv_c192to128_128.ncu-rep ${P}_conv_c128to64_256.ncu-rep ${P}_conv_c1x1_128to64_256.ncu-rep ${P}_glue.ncu-rep; echo "select exit=$?"
rm -f ${P}_conv_*.ncu-rep ${P}_glue.ncu-rep; du -sh gpurun_out
