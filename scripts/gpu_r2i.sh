#!/bin/bash
# Round 2, GPU call I: the SHIPPED build - parity suite, smoke, bench lines of configs 2 / 3 / 4 with cpu baselines, reference arms,
# ncu launch lists + full captures (viterbi_kernel, viterbi_cong_kernel, lattice_kernel), per-sentence timing.
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2i_pytest.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2i_smoke.log 2>&1
( timeout 300 python bench.py --steps 20 --warmup 3 ) > $O/r2i_bench_cfg2.json 2> $O/r2i_bench_cfg2.err
( timeout 500 python bench.py --config 3 --steps 5 --warmup 3 ) > $O/r2i_bench_cfg3.json 2> $O/r2i_bench_cfg3.err
( timeout 700 python bench.py --config 4 --steps 5 --warmup 3 ) > $O/r2i_bench_cfg4.json 2> $O/r2i_bench_cfg4.err
( timeout 400 python bench.py --impl reference --steps 5 --warmup 1 ) > $O/r2i_ref_cfg2.json 2> $O/r2i_ref_cfg2.err
( timeout 400 python bench.py --impl reference --config 3 --steps 5 --warmup 1 ) > $O/r2i_ref_cfg3.json 2> $O/r2i_ref_cfg3.err
( timeout 600 python bench.py --impl reference --config 4 --steps 5 --warmup 1 ) > $O/r2i_ref_cfg4.json 2> $O/r2i_ref_cfg4.err
( timeout 100 python scripts/gpu_timing.py knlm r2e ) > $O/r2i_timing.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 48 --csv --log-file $O/r2i_launches_knlm.csv python bench.py --steps 3 --warmup 2 --no-cpu > $O/r2i_ncu_l.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 48 --csv --log-file $O/r2i_launches_cong.csv python bench.py --model cong --steps 3 --warmup 2 --no-cpu > $O/r2i_ncu_lc.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:viterbi_kernel -s 1 -c 1 -o $O/r2i_viterbi_knlm python bench.py --steps 2 --warmup 1 --no-cpu > $O/r2i_ncu_f.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:viterbi_cong_kernel -s 1 -c 1 -o $O/r2i_viterbi_cong python bench.py --model cong --steps 2 --warmup 1 --no-cpu > $O/r2i_ncu_fc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lattice_kernel -s 1 -c 1 -o $O/r2i_lattice python bench.py --steps 2 --warmup 1 --no-cpu > $O/r2i_ncu_fl.log 2>&1
tail -n 6 $O/r2i_pytest.log; tail -n 3 $O/r2i_smoke.log
for f in $O/r2i_bench_*.json $O/r2i_var_*.json $O/r2i_ref_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("value %.0f e2e %.0f ms/step %.2f vit_ms %s lat_ms %s frac %s cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
tail -n 2 $O/r2i_timing.log | cut -c1-500
