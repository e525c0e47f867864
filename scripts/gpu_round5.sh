#!/bin/bash
# GPU-box round 5: full verification of the shipped build + bench lines + ncu captures for profiles/.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
( timeout 200 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_knlm.json 2> gpurun_out/bench_knlm.err
( timeout 200 python bench.py --model cong --steps 20 --warmup 3 --cpu-sample 8192 ) > gpurun_out/bench_cong.json 2> gpurun_out/bench_cong.err
( timeout 100 python scripts/gpu_timing.py knlm knlm ) > gpurun_out/timing_knlm.log 2>&1
( timeout 100 python scripts/gpu_timing.py cong cong ) > gpurun_out/timing_cong.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r1b_launches_knlm.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_lk.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_kernel -s 1 -c 1 -o gpurun_out/r1b_viterbi_knlm python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_fk.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_cong_kernel -s 1 -c 1 -o gpurun_out/r1b_viterbi_cong python bench.py --model cong --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_fc.log 2>&1
tail -n 4 gpurun_out/pytest.log; tail -n 3 gpurun_out/smoke.log
cut -c1-300 gpurun_out/bench_knlm.json; cut -c1-300 gpurun_out/bench_cong.json
tail -n 1 gpurun_out/timing_*.log | cut -c1-300; tail -n 2 gpurun_out/ncu_fk.log gpurun_out/ncu_fc.log | cut -c1-200
