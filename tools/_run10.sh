mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_${N}gpu.err | grep '^{' | tail -1) > gpurun_out/bench_${N}gpu.json
(timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --batch 5 2> gpurun_out/bench_${N}gpu_b5.err | grep '^{' | tail -1) > gpurun_out/bench_${N}gpu_b5.json
cut -c1-250 gpurun_out/bench_${N}gpu.json; cut -c1-250 gpurun_out/bench_${N}gpu_b5.json; tail -4 gpurun_out/bench_${N}gpu.err
