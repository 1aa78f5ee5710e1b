mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/t11.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -45) > gpurun_out/smoke11.log
(timeout 400 python bench.py 2> gpurun_out/bench11.err | tail -1) > gpurun_out/bench11.json
(timeout 400 python bench.py --batch 5 --no-cpu-baseline 2> gpurun_out/bench11_b5.err | tail -1) > gpurun_out/bench11_b5.json
for DT in f32 bf16; do
  B=32 DT=$DT timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_b32_${DT}.csv python tools/step_profile.py > gpurun_out/step_${DT}.log 2>&1
  python tools/launch_summary.py gpurun_out/step_b32_${DT}.csv 90 > gpurun_out/step_b32_${DT}_launches_final.txt
  rm -f gpurun_out/step_b32_${DT}.csv
done
tail -3 gpurun_out/t11.log; tail -2 gpurun_out/smoke11.log; cut -c1-200 gpurun_out/bench11.json; cut -c1-200 gpurun_out/bench11_b5.json; head -2 gpurun_out/step_b32_f32_launches_final.txt
