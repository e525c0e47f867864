#!/bin/bash
# Round 2, GPU call K: result token arrays in page-locked memory, device-to-host copies land in the caller's array (no host staging copy):
# parity suite + end-to-end numbers of configs 2 / 3.
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2k_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu ) > $O/r2k_bench_cfg2.json 2> $O/r2k_bench_cfg2.err
( timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu ) > $O/r2k_bench_cfg2b.json 2> $O/r2k_bench_cfg2b.err
( timeout 500 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu ) > $O/r2k_bench_cfg3.json 2> $O/r2k_bench_cfg3.err
tail -n 6 $O/r2k_pytest.log
for f in $O/r2k_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("value %.0f e2e %.0f (%.3f) ms/step %.2f vit_ms %s lat_ms %s frac %s" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
