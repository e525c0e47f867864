#!/bin/bash
# One GPU-box round (run through gpurun): CoNg diagnostics, the -m gpu suite, smoke, both bench lines, per-sentence timing and -
# with KB_NCU=1 - the ncu launch lists and full captures that profiles/ summarises.  Everything lands in gpurun_out/.
# Kernel experiments: build another viterbi object with `make -C kiwi_b200/csrc variant NAME=x VFLAGS="-D..."` (or variant_cong)
# and select it with KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_x.so; KIWI_B200_CARVEOUT=<pct> overrides the smem/L1 split.
mkdir -p gpurun_out
( time timeout 200 python scripts/gpu_cong_diag.py ) > gpurun_out/diag.log 2>&1
( time timeout 480 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
( timeout 200 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_knlm.json 2> gpurun_out/bench_knlm.err
( timeout 200 python bench.py --model cong --steps 20 --warmup 3 ) > gpurun_out/bench_cong.json 2> gpurun_out/bench_cong.err
( timeout 100 python scripts/gpu_timing.py knlm knlm ) > gpurun_out/timing_knlm.log 2>&1
( timeout 100 python scripts/gpu_timing.py cong cong ) > gpurun_out/timing_cong.log 2>&1
if [ -n "$KB_NCU" ]; then
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_knlm.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_lk.log 2>&1
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_cong.csv python bench.py --model cong --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_lc.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_kernel -s 1 -c 1 -o gpurun_out/viterbi_knlm python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_fk.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_cong_kernel -s 1 -c 1 -o gpurun_out/viterbi_cong python bench.py --model cong --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_fc.log 2>&1
fi
tail -c 1200 gpurun_out/diag.log; tail -n 8 gpurun_out/pytest.log; tail -n 3 gpurun_out/smoke.log
cut -c1-300 gpurun_out/bench_knlm.json; cut -c1-300 gpurun_out/bench_cong.json; tail -n 1 gpurun_out/timing_*.log | cut -c1-300
