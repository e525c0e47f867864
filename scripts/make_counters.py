#!/usr/bin/env python3
"""Work counters of the byte model for the bench batches, from the instrumented oracle.
Writes profiles/counters_r1.json (Knlm) and profiles/counters_r1b_cong.json (CoNg, incl. the gather-GEMM row / MAC counters);
bench.py turns them into algorithmic bytes per sentence."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kiwi_b200.synth import synth_batch, SEED
from kiwi_b200 import bytemodel
from tests.orc import Oracle, CONG_IMAGE
out = {}
for batch in [8192]:
    o = Oracle()
    t0 = time.time()
    for s in synth_batch(batch, SEED): o.analyze(s)
    c = o.work_counters(); o.close()
    c["oracle_seconds"] = time.time() - t0
    c["lattice_bytes_per_sentence"] = bytemodel.lattice_bytes(c) / c["sentences"]
    c["viterbi_bytes_per_sentence"] = bytemodel.viterbi_bytes(c) / c["sentences"]
    out["batch%d_seed%d" % (batch, SEED)] = c
    print(c)
json.dump(out, open(os.path.join(ROOT, "profiles", "counters_r1.json"), "w"), indent=1)

out = {}
for batch in [8192]:
    o = Oracle(CONG_IMAGE)
    t0 = time.time()
    for s in synth_batch(batch, SEED): o.analyze(s)
    c = o.work_counters(); c.update(o.cong_counters()); o.close()
    c["oracle_seconds"] = time.time() - t0
    c["lattice_bytes_per_sentence"] = bytemodel.lattice_bytes(c) / c["sentences"]
    c["viterbi_bytes_per_sentence"] = (bytemodel.viterbi_bytes(c) + bytemodel.cong_bytes(c)) / c["sentences"]
    out["batch%d_seed%d" % (batch, SEED)] = c
    print(c)
json.dump(out, open(os.path.join(ROOT, "profiles", "counters_r1b_cong.json"), "w"), indent=1)
