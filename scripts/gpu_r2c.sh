#!/bin/bash
# Round 2, GPU call C: parity suite (std::sort + unordered_set order restated on the device), team mode (heaviest sentences by a team of
# warps) at several shares, TMA-only staging buffers, block shapes.
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2c_pytest.log
for tp in 0 10 30 60 120 250; do
  ( KIWI_B200_TEAM_PERMILLE=$tp timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu ) > $O/r2c_team_$tp.json 2> $O/r2c_team_$tp.err
done
for v in notma w8b1 w12b1 w4b4 w8b2t2 w8b2t8; do
  for tp in 0 60; do
    ( KIWI_B200_TEAM_PERMILLE=$tp KIWI_B200_LIB=kiwi_b200/variants/libkiwi_b200_$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu ) > $O/r2c_var_${v}_$tp.json 2> $O/r2c_var_${v}_$tp.err
  done
done
( KIWI_B200_TEAM_PERMILLE=60 timeout 100 python scripts/gpu_timing.py knlm r2c_team60 ) > $O/r2c_timing_team60.log 2>&1
( KIWI_B200_TEAM_PERMILLE=0 timeout 100 python scripts/gpu_timing.py knlm r2c_team0 ) > $O/r2c_timing_team0.log 2>&1
# parity of team mode on the hardware: the whole suite again with every 4th sentence in a team
( KIWI_B200_TEAM_PERMILLE=250 timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "not config3 and not config4" ) > $O/r2c_pytest_team.log 2>&1
echo "pytest rc=$?" >> $O/r2c_pytest_team.log
for tp in 0 60; do
  ( KIWI_B200_TEAM_PERMILLE=$tp timeout 400 python bench.py --config 3 --steps 3 --warmup 2 --no-cpu ) > $O/r2c_cfg3_$tp.json 2> $O/r2c_cfg3_$tp.err
  ( KIWI_B200_TEAM_PERMILLE=$tp timeout 600 python bench.py --config 4 --steps 3 --warmup 2 --no-cpu ) > $O/r2c_cfg4_$tp.json 2> $O/r2c_cfg4_$tp.err
done
tail -n 12 $O/r2c_pytest.log; tail -n 8 $O/r2c_pytest_team.log
for f in $O/r2c_team_*.json $O/r2c_var_*.json $O/r2c_cfg*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("value %.0f e2e %.0f ms/step %.2f vit_ms %s lat_ms %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
tail -n 2 $O/r2c_timing_*.log | cut -c1-600
