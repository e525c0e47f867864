#!/bin/bash
# GPU-box round 6: shared-memory carve-out / item-buffer experiments (Knlm), parity suite first.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
run() { # name lib carveout
  ( KIWI_B200_LIB=$2 KIWI_B200_CARVEOUT=$3 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_knlm_$1.json 2> gpurun_out/bench_knlm_$1.err
}
( timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_knlm_default.json 2> gpurun_out/bench_knlm_default.err
for c in 50 58 64 72 86 100; do run co$c kiwi_b200/libkiwi_b200.so $c; done
for c in 50 58 64 100; do run ic384co$c kiwi_b200/variants/libkiwi_b200_ic384.so $c; done
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_ic384.so timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_knlm_ic384.json 2> gpurun_out/bench_knlm_ic384.err
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_ic256.so timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_knlm_ic256.json 2> gpurun_out/bench_knlm_ic256.err
for c in 50 64 100; do ( KIWI_B200_CARVEOUT=$c timeout 120 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_cong_co$c.json 2> gpurun_out/bench_cong_co$c.err; done
tail -n 3 gpurun_out/pytest.log
for f in gpurun_out/bench_knlm_*.json gpurun_out/bench_cong_co*.json; do echo -n "$f "; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["value"]), "e2e", round(d["e2e"]["value"]), "vit ms", round(d["roofline"]["kernel_ms_per_launch"],2))
PY
done
