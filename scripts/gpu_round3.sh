#!/bin/bash
# GPU-box round 3: occupancy / lockstep experiments (all parity-checked by the -m gpu suite on the default build first).
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
( timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_knlm.json 2> gpurun_out/bench_knlm.err
for v in ls2 ls4 ls8 b3 ls4b3; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_knlm_$v.json 2> gpurun_out/bench_knlm_$v.err
done
( timeout 150 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_cong.json 2> gpurun_out/bench_cong.err
for v in cgls4 cgb2 cgu64b3; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 120 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_cong_$v.json 2> gpurun_out/bench_cong_$v.err
done
( timeout 100 python scripts/gpu_timing.py cong cong ) > gpurun_out/timing_cong.log 2>&1
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_ls4.so timeout 100 python scripts/gpu_timing.py knlm knlm_ls4 ) > gpurun_out/timing_knlm_ls4.log 2>&1
tail -n 4 gpurun_out/pytest.log
for f in gpurun_out/bench_knlm.json gpurun_out/bench_knlm_ls2.json gpurun_out/bench_knlm_ls4.json gpurun_out/bench_knlm_ls8.json gpurun_out/bench_knlm_b3.json gpurun_out/bench_knlm_ls4b3.json gpurun_out/bench_cong.json gpurun_out/bench_cong_cgls4.json gpurun_out/bench_cong_cgb2.json gpurun_out/bench_cong_cgu64b3.json; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["value"]), "e2e", round(d["e2e"]["value"]), "vit ms", round(d["roofline"]["kernel_ms_per_launch"],2), "retried", d["config"]["retried_sentences_per_step"])
PY
done
tail -n 1 gpurun_out/timing_*.log | cut -c1-400
