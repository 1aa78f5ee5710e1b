mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_splat.py tests/test_sampling_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/t3.log
for M in 1 2 0; do
  echo "== GG_SPLAT_MODE=$M" >> gpurun_out/splat_modes.txt
  GG_SPLAT_MODE=$M GG_OPBENCH_ONLY=splat timeout 300 python tools/opbench.py --ref --batch 32 --graph 2>&1 | grep splat >> gpurun_out/splat_modes.txt
done
timeout 600 python tools/configbench.py --config 5 --json gpurun_out/config5_1gpu_f32.json > gpurun_out/config5.log 2>&1
timeout 600 python tools/configbench.py --config 5 --dtype bf16 --json gpurun_out/config5_1gpu_bf16.json >> gpurun_out/config5.log 2>&1
timeout 600 python tools/configbench.py --config 4 --json gpurun_out/config4_1gpu.json > gpurun_out/config4.log 2>&1
tail -3 gpurun_out/t3.log; cat gpurun_out/splat_modes.txt; tail -2 gpurun_out/config5.log | cut -c1-300; tail -1 gpurun_out/config4.log | cut -c1-600
