mkdir -p gpurun_out
timeout 600 python tools/configbench.py --config 5 --json gpurun_out/config5_1gpu_f32.json > gpurun_out/config5.log 2>&1
timeout 600 python tools/configbench.py --config 5 --dtype bf16 --json gpurun_out/config5_1gpu_bf16.json >> gpurun_out/config5.log 2>&1
timeout 600 python tools/configbench.py --config 4 --json gpurun_out/config4_1gpu.json > gpurun_out/config4.log 2>&1
(timeout 400 python bench.py --batch 5 --no-cpu-baseline 2> gpurun_out/bench_b5.err | tail -1) > gpurun_out/bench_b5.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gg:: -o gpurun_out/misc -f python tools/prof_misc.py > gpurun_out/prof_misc.log 2>&1
python tools/ncu_summary.py gpurun_out/misc.ncu-rep > gpurun_out/misc_ncu_summary.txt 2>&1
rm -f gpurun_out/misc.ncu-rep
grep -h '"value"' gpurun_out/config5.log | cut -c1-220; tail -2 gpurun_out/config5.log | cut -c1-300; tail -1 gpurun_out/config4.log | cut -c1-300; cut -c1-300 gpurun_out/bench_b5.json; tail -3 gpurun_out/prof_misc.log; grep -c "^==" gpurun_out/misc_ncu_summary.txt
