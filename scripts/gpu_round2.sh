#!/bin/bash
# GPU-box round 2: parity suite on the top1-mode item pipeline, variant benches, ncu captures of the CoNg kernel.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
( timeout 150 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_knlm.json 2> gpurun_out/bench_knlm.err
( timeout 100 python scripts/gpu_timing.py knlm knlm ) > gpurun_out/timing_knlm.log 2>&1
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_s512.so timeout 100 python scripts/gpu_timing.py knlm knlm_s512 ) > gpurun_out/timing_knlm_s512.log 2>&1
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_solo256.so timeout 100 python scripts/gpu_timing.py knlm knlm_solo256 ) > gpurun_out/timing_knlm_solo256.log 2>&1
( timeout 100 python scripts/gpu_timing.py cong cong ) > gpurun_out/timing_cong.log 2>&1
for v in s512 s1024 s1536b3 solo64 solo256 solo256b3 nolock; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_knlm_$v.json 2> gpurun_out/bench_knlm_$v.err
done
( timeout 150 python bench.py --model cong --steps 10 --warmup 3 --cpu-sample 8192 ) > gpurun_out/bench_cong.json 2> gpurun_out/bench_cong.err
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_cgnopipe.so timeout 120 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_cong_cgnopipe.json 2> gpurun_out/bench_cong_cgnopipe.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r1b_launches_cong.csv python bench.py --model cong --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_cong_kernel -s 1 -c 1 -o gpurun_out/r1b_viterbi_cong python bench.py --model cong --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_f.log 2>&1
tail -n 6 gpurun_out/pytest.log
for f in gpurun_out/bench_knlm.json gpurun_out/bench_knlm_s512.json gpurun_out/bench_knlm_s1024.json gpurun_out/bench_knlm_s1536b3.json gpurun_out/bench_knlm_solo64.json gpurun_out/bench_knlm_solo256.json gpurun_out/bench_knlm_solo256b3.json gpurun_out/bench_knlm_nolock.json gpurun_out/bench_cong.json gpurun_out/bench_cong_cgnopipe.json; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["value"]), "e2e", round(d["e2e"]["value"]), "vit ms", round(d["roofline"]["kernel_ms_per_launch"],2), "retried", d["config"]["retried_sentences_per_step"])
PY
done
tail -n 3 gpurun_out/ncu_l.log gpurun_out/ncu_f.log; tail -n 2 gpurun_out/timing_*.log
