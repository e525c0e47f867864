#!/bin/bash
# Round 2, GPU call A: the -m gpu suite on the new kernels, bench lines for configs 2 / 3 / 4, the reference arm, kernel variants
# (lockstep on / off x resident blocks), pass-size experiment for e2e, per-sentence timing.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r2a_smi.txt 2>&1
nproc > $O/r2a_nproc.txt; cat /sys/fs/cgroup/cpu.max >> $O/r2a_nproc.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2a_pytest.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2a_smoke.log 2>&1
( timeout 300 python bench.py --steps 20 --warmup 3 ) > $O/r2a_bench_cfg2.json 2> $O/r2a_bench_cfg2.err
for v in nolock nolock5 nolock6 lock5 nolock5w8; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu ) > $O/r2a_var_$v.json 2> $O/r2a_var_$v.err
done
for v in cgnolock cgnolock4; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 200 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > $O/r2a_var_$v.json 2> $O/r2a_var_$v.err
done
( timeout 200 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > $O/r2a_bench_cong8192.json 2> $O/r2a_bench_cong8192.err
( timeout 400 python bench.py --config 3 --steps 3 --warmup 3 ) > $O/r2a_bench_cfg3.json 2> $O/r2a_bench_cfg3.err
( timeout 600 python bench.py --config 4 --steps 3 --warmup 3 ) > $O/r2a_bench_cfg4.json 2> $O/r2a_bench_cfg4.err
( timeout 400 python bench.py --impl reference --steps 5 --warmup 1 ) > $O/r2a_bench_ref.json 2> $O/r2a_bench_ref.err
for ps in 4096 2048; do
  ( KIWI_B200_PASS_SENT=$ps timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu ) > $O/r2a_pass_$ps.json 2> $O/r2a_pass_$ps.err
done
( timeout 100 python scripts/gpu_timing.py knlm r2a_knlm ) > $O/r2a_timing_knlm.log 2>&1
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_nolock.so timeout 100 python scripts/gpu_timing.py knlm r2a_knlm_nolock ) > $O/r2a_timing_knlm_nolock.log 2>&1
tail -n 15 $O/r2a_pytest.log; tail -n 3 $O/r2a_smoke.log
for f in $O/r2a_bench_*.json $O/r2a_var_*.json $O/r2a_pass_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("value %.0f e2e %.0f ms/step %.2f vit_ms %s lat_ms %s frac %s cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
