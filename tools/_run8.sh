mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_perceptual.py tests/test_nhwc_gpu.py tests/test_styled_fused_gpu.py tests/test_networks_gpu.py -x -q -m gpu 2>&1 | tail -4) > gpurun_out/t8.log
for DT in f32 bf16; do
  B=5 DT=$DT timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_b5_${DT}.csv python tools/step_profile.py > gpurun_out/step_b5_${DT}.log 2>&1
  python tools/launch_summary.py gpurun_out/step_b5_${DT}.csv 70 > gpurun_out/step_b5_${DT}_launches.txt
  rm -f gpurun_out/step_b5_${DT}.csv
done
(timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/bench8.err | tail -1) > gpurun_out/bench8.json
tail -3 gpurun_out/t8.log; head -3 gpurun_out/step_b5_f32_launches.txt; cut -c1-200 gpurun_out/bench8.json
