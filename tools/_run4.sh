mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/t4.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12) > gpurun_out/smoke4.log
timeout 600 python tools/configbench.py --config 5 --json gpurun_out/config5_1gpu_f32.json > gpurun_out/config5.log 2>&1
timeout 600 python tools/configbench.py --config 5 --dtype bf16 --json gpurun_out/config5_1gpu_bf16.json >> gpurun_out/config5.log 2>&1
timeout 600 python tools/configbench.py --config 4 --json gpurun_out/config4_1gpu.json > gpurun_out/config4.log 2>&1
(timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/bench4.err | tail -1) > gpurun_out/bench4.json
GG_OPBENCH_ONLY=splat timeout 300 python tools/opbench.py --ref --batch 32 --graph 2>&1 | grep splat > gpurun_out/splat_final.txt
tail -4 gpurun_out/t4.log; tail -3 gpurun_out/smoke4.log; grep -h '"value"' gpurun_out/config5.log | cut -c1-200; cut -c1-400 gpurun_out/bench4.json; cat gpurun_out/splat_final.txt
