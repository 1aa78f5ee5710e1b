mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench_2gpu.err | grep '^{' | tail -1) > gpurun_out/bench_2gpu.json
(timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --batch 5 2> gpurun_out/bench_2gpu_b5.err | grep '^{' | tail -1) > gpurun_out/bench_2gpu_b5.json
(timeout 600 $TR --master-port 29513 tools/configbench.py --config 5 --json gpurun_out/config5_2gpu_f32.json 2>&1 | tail -3) > gpurun_out/config5_2gpu.log
(timeout 600 $TR --master-port 29514 tools/configbench.py --config 4 --json gpurun_out/config4_2gpu.json 2>&1 | tail -3) > gpurun_out/config4_2gpu.log
(timeout 300 $TR --master-port 29515 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2> gpurun_out/bench_ref.err | grep '^{' | tail -1) > gpurun_out/bench_ref_2gpu.json
cut -c1-250 gpurun_out/bench_2gpu.json; cut -c1-250 gpurun_out/bench_2gpu_b5.json; cut -c1-200 gpurun_out/config5_2gpu.log; cut -c1-200 gpurun_out/config4_2gpu.log; cut -c1-300 gpurun_out/bench_ref_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
