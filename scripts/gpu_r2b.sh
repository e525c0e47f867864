#!/bin/bash
# Round 2, GPU call B: -m gpu suite after the std::sort tie fix, TMA-staged candidate rows (default build) vs variants, ncu launch list
# + full capture of viterbi_kernel, per-sentence timing.
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2b_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu ) > $O/r2b_bench_cfg2.json 2> $O/r2b_bench_cfg2.err
for v in w8 w8notma w8s1536 w8s1536i256 w16 lockw8; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu ) > $O/r2b_var_$v.json 2> $O/r2b_var_$v.err
done
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_cgw8.so timeout 200 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > $O/r2b_var_cgw8.json 2> $O/r2b_var_cgw8.err
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_w8.so timeout 100 python scripts/gpu_timing.py knlm r2b_w8 ) > $O/r2b_timing_w8.log 2>&1
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_w8s1536.so timeout 100 python scripts/gpu_timing.py knlm r2b_w8s1536 ) > $O/r2b_timing_w8s1536.log 2>&1
export KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_w8.so
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file $O/r2b_launches_knlm.csv python bench.py --steps 3 --warmup 2 --no-cpu > $O/r2b_ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:viterbi_kernel -s 1 -c 1 -o $O/r2b_viterbi_knlm python bench.py --steps 2 --warmup 1 --no-cpu > $O/r2b_ncu_f.log 2>&1
unset KIWI_B200_LIB
tail -n 12 $O/r2b_pytest.log
for f in $O/r2b_bench_*.json $O/r2b_var_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("value %.0f e2e %.0f ms/step %.2f vit_ms %s lat_ms %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
tail -n 2 $O/r2b_timing_*.log | cut -c1-400
