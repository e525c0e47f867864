#!/bin/bash
# Round 2, GPU call G: flush pipelining reverted (measured slower), recycled result buffers; last block-shape / inlining variants.
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2g_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu ) > $O/r2g_bench_cfg2.json 2> $O/r2g_bench_cfg2.err
for v in inl w11 w13 i384; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu ) > $O/r2g_var_$v.json 2> $O/r2g_var_$v.err
done
( timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu ) > $O/r2g_bench_cfg2_again.json 2> $O/r2g_bench_cfg2_again.err
tail -n 5 $O/r2g_pytest.log
for f in $O/r2g_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("value %.0f e2e %.0f ms/step %.2f vit_ms %s lat_ms %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
