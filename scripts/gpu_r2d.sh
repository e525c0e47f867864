#!/bin/bash
# Round 2, GPU call D: block shapes x team code on/off x mode-2 staging x TMA, team shares, launch-order key; parity suite on the default build.
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2d_pytest.log
run() { # name lib team lpt
  ( KIWI_B200_TEAM_PERMILLE=$3 KIWI_B200_LPT=$4 KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$2.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu ) > $O/r2d_$1.json 2> $O/r2d_$1.err
}
for v in A B C E F G I; do run ${v}_cost $v 0 cost; done
run A_len A 0 len
run C_len C 0 len
for tp in 0 15 30 60 120; do run D_t$tp D $tp cost; done
for tp in 30 60; do run H_t$tp H $tp cost; done
for v in cgA cgB cgC; do
  ( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 200 python bench.py --model cong --steps 10 --warmup 3 --no-cpu ) > $O/r2d_$v.json 2> $O/r2d_$v.err
done
( KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_C.so timeout 100 python scripts/gpu_timing.py knlm r2d_C ) > $O/r2d_timing_C.log 2>&1
( KIWI_B200_TEAM_PERMILLE=30 KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_D.so timeout 100 python scripts/gpu_timing.py knlm r2d_D30 ) > $O/r2d_timing_D30.log 2>&1
# parity of the mode-2 pipeline + team mode on the hardware
( KIWI_B200_TEAM_PERMILLE=200 KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_D.so timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "not config3" ) > $O/r2d_pytest_D.log 2>&1
echo "pytest rc=$?" >> $O/r2d_pytest_D.log
tail -n 6 $O/r2d_pytest.log; tail -n 6 $O/r2d_pytest_D.log
for f in $O/r2d_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("value %.0f e2e %.0f ms/step %.2f vit_ms %s lat_ms %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
tail -n 2 $O/r2d_timing_*.log | cut -c1-500
