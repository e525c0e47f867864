#!/bin/bash
# Round 2, GPU call J: work-queue build of the Viterbi kernels (persistent warps draw sentences from a global counter) and solo blocks for
# the heaviest sentences, against the shipped static assignment on the same box.
mkdir -p gpurun_out; O=gpurun_out
Q=kiwi_b200/variants/libkiwi_b200_q.so; QC=kiwi_b200/variants/libkiwi_b200_qc.so
( timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu ) > $O/r2j_default.json 2> $O/r2j_default.err
( KIWI_B200_LIB=$Q timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu ) > $O/r2j_q.json 2> $O/r2j_q.err
for solo in 1,1 2,1 4,1 4,2 8,4; do
  ( KIWI_B200_SOLO=$solo KIWI_B200_LIB=$Q timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu ) > $O/r2j_q_solo_$solo.json 2> $O/r2j_q_solo_$solo.err
done
( KIWI_B200_SOLO=4,1 KIWI_B200_LIB=$Q timeout 100 python scripts/gpu_timing.py knlm r2j_solo41 ) > $O/r2j_timing_solo41.log 2>&1
( timeout 400 python bench.py --config 4 --steps 4 --warmup 3 --no-cpu ) > $O/r2j_default_cfg4.json 2> $O/r2j_default_cfg4.err
( KIWI_B200_LIB=$Q timeout 400 python bench.py --config 4 --steps 4 --warmup 3 --no-cpu ) > $O/r2j_q_cfg4.json 2> $O/r2j_q_cfg4.err
( timeout 400 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu ) > $O/r2j_default_cfg3.json 2> $O/r2j_default_cfg3.err
( KIWI_B200_LIB=$QC timeout 400 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu ) > $O/r2j_qc_cfg3.json 2> $O/r2j_qc_cfg3.err
( KIWI_B200_SOLO=4,1 KIWI_B200_LIB=$Q timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "not cong and not config3" ) > $O/r2j_pytest_q.log 2>&1
echo "pytest rc=$?" >> $O/r2j_pytest_q.log
tail -n 4 $O/r2j_pytest_q.log
for f in $O/r2j_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("value %.0f e2e %.0f ms/step %.2f vit_ms %s lat_ms %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("lattice_ms_per_step"), r.get("frac")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
tail -n 2 $O/r2j_timing_solo41.log | cut -c1-600
