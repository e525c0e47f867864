#!/bin/bash
# One GPU-box round: CoNg diagnostics, the -m gpu suite, the two bench lines, smoke.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
( time timeout 200 python scripts/gpu_cong_diag.py ) > gpurun_out/diag.log 2>&1
( time timeout 480 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest.log 2>&1
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest.log
( time timeout 180 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_knlm.json 2> gpurun_out/bench_knlm.err
( time timeout 180 python bench.py --model cong --steps 10 --warmup 3 --cpu-sample 2048 ) > gpurun_out/bench_cong.json 2> gpurun_out/bench_cong.err
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
if [ -n "$KB_NCU" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r1b_launches_cong.csv python bench.py --model cong --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_l.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:viterbi_cong_kernel -s 1 -c 1 -o gpurun_out/r1b_viterbi_cong python bench.py --model cong --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_f.log 2>&1
fi
tail -c 1500 gpurun_out/diag.log; tail -n 15 gpurun_out/pytest.log; cat gpurun_out/bench_knlm.json | cut -c1-600; cat gpurun_out/bench_cong.json | cut -c1-600; tail -n 4 gpurun_out/smoke.log
