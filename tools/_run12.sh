mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_nhwc_gpu.py tests/test_styled_fused_gpu.py tests/test_stylegan2_ops_gpu.py tests/test_reference_kernels_gpu.py tests/test_networks_gpu.py -x -q -m gpu 2>&1 | tail -5) > gpurun_out/t12.log
for DT in f32 bf16; do
  DT=$DT B=32 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k 'regex:blur_nhwc|styled_tail' --log-file gpurun_out/nhwc_time_$DT.csv python tools/prof_nhwc.py > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/nhwc_time_$DT.csv 20 > gpurun_out/nhwc_time_$DT.txt; rm -f gpurun_out/nhwc_time_$DT.csv
done
(timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/bench12.err | tail -1) > gpurun_out/bench12.json
tail -3 gpurun_out/t12.log; cat gpurun_out/nhwc_time_bf16.txt | cut -c1-120; cut -c1-200 gpurun_out/bench12.json
