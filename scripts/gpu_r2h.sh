#!/bin/bash
# Round 2, GPU call H: inlining of single-call-site functions (evaluate / stagePaths / fixupGroup / the first Knlm step of a flush round).
mkdir -p gpurun_out; O=gpurun_out
for v in inl inl2 inl3 inl4 inl2w10 inl2tma; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu ) > $O/r2h_var_$v.json 2> $O/r2h_var_$v.err
done
( timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu ) > $O/r2h_default.json 2> $O/r2h_default.err
for v in cginl cginl2; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 200 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > $O/r2h_var_$v.json 2> $O/r2h_var_$v.err
done
( timeout 200 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > $O/r2h_default_cong.json 2> $O/r2h_default_cong.err
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_inl4.so timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "not config3 and not config4" ) > $O/r2h_pytest_inl4.log 2>&1
echo "pytest rc=$?" >> $O/r2h_pytest_inl4.log
tail -n 4 $O/r2h_pytest_inl4.log
for f in $O/r2h_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("value %.0f e2e %.0f ms/step %.2f vit_ms %s lat_ms %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
