#!/bin/bash
# Round 2, GPU call L: the SkipBigram build (viterbi_sbg_kernel) - parity with the reference's ModelType::sbg vectors, bench line of the
# SkipBigram configuration with the reference beside it; then the whole -m gpu suite and the config-2 line with the final library.
mkdir -p gpurun_out; O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sbg.py -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/r2l_pytest_sbg.log 2>&1
echo "pytest rc=$?" >> $O/r2l_pytest_sbg.log
tail -n 15 $O/r2l_pytest_sbg.log
( timeout 600 python bench.py --config 6 --steps 3 --warmup 3 --cpu-sample 1024 ) > $O/r2l_bench_cfg6.json 2> $O/r2l_bench_cfg6.err
( time timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_sbg.py ) > $O/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2l_pytest.log
tail -n 6 $O/r2l_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 3 ) > $O/r2l_bench_cfg2.json 2> $O/r2l_bench_cfg2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file $O/r2l_launches_sbg.csv python bench.py --config 6 --steps 2 --warmup 2 --no-cpu > $O/r2l_ncu_l.log 2>&1
for f in $O/r2l_bench_*.json $O/r2l_ref_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("value %.0f e2e %.0f ms/step %.2f vit_ms %s lat_ms %s frac %s cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
