mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_reference_kernels_gpu.py tests/test_stylegan2_ops_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/t2.log
(timeout 600 python tools/opbench.py --ref --batch 32 --graph --json gpurun_out/opbench_ref_b32_graph.json 2>&1 | tail -60) > gpurun_out/opbench_ref_b32_graph.txt
for DT in f32 bf16; do
  B=32 DT=$DT timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_b32_${DT}.csv python tools/step_profile.py > gpurun_out/step_${DT}.log 2>&1
  python tools/launch_summary.py gpurun_out/step_b32_${DT}.csv 90 > gpurun_out/step_b32_${DT}_launches.txt
  rm -f gpurun_out/step_b32_${DT}.csv
done
tail -3 gpurun_out/t2.log; head -3 gpurun_out/step_b32_f32_launches.txt
