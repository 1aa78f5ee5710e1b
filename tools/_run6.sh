mkdir -p gpurun_out
K='regex:warp_|mip_|flow_compose|splat_|feature_distance|bias_relu_pool|tent_down|adam_ema|tv_fwd|tv_bwd|nn_argmin|up2_k4|down2_k4|demod_umma'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 0 -o gpurun_out/misc -f python tools/prof_misc.py > gpurun_out/prof_misc.log 2>&1
python tools/ncu_summary.py gpurun_out/misc.ncu-rep > gpurun_out/misc_ncu_summary.txt 2>&1
rm -f gpurun_out/misc.ncu-rep
for DT in f32 bf16; do
  DT=$DT B=32 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:blur_nhwc|styled_tail|rowwise_nhwc|noise_bias_act_nhwc' -o gpurun_out/nhwc_$DT -f python tools/prof_nhwc.py > gpurun_out/prof_nhwc_$DT.log 2>&1
  python tools/ncu_summary.py gpurun_out/nhwc_$DT.ncu-rep > gpurun_out/nhwc_b32_${DT}_ncu_summary.txt 2>&1
  rm -f gpurun_out/nhwc_$DT.ncu-rep
done
grep -c "^==" gpurun_out/misc_ncu_summary.txt gpurun_out/nhwc_b32_f32_ncu_summary.txt gpurun_out/nhwc_b32_bf16_ncu_summary.txt; tail -2 gpurun_out/prof_misc.log
