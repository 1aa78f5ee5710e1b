mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_networks_gpu.py tests/test_reference_dropin_gpu.py -x -q -m gpu 2>&1 | tail -15) > gpurun_out/t9.log
(timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/bench9.err | tail -1) > gpurun_out/bench9.json
(timeout 400 python bench.py --batch 5 --no-cpu-baseline 2> gpurun_out/bench9_b5.err | tail -1) > gpurun_out/bench9_b5.json
tail -6 gpurun_out/t9.log; cut -c1-200 gpurun_out/bench9.json; cut -c1-200 gpurun_out/bench9_b5.json; tail -3 gpurun_out/bench9.err
