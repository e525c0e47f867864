#!/usr/bin/env python3
"""Every SkipBigram golden vector (tests/golden/sbg_*) against the oracle restatement, bit for bit: python scripts/sbg_oracle_check.py [inputs_written inputs_web inputs_ref_tests]"""
import sys, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from tests.orc import Oracle, SBG_IMAGE
from tests.goldenio import read_golden, read_inputs
import time
o=Oracle(SBG_IMAGE)
names = sys.argv[1:] or ["inputs_written", "inputs_web", "inputs_ref_tests"]
for name in names:
    texts=read_inputs(name); gold=read_golden("sbg_"+name)
    bad=[]
    t0=time.time()
    for i,(t,g) in enumerate(zip(texts,gold)):
        toks,score=o.analyze(t)
        if [x[:4] for x in toks]!=[x[:4] for x in g["tokens"]] or np.float32(score)!=np.float32(g["score"]) or [np.float32(x[4]) for x in toks] != [np.float32(x[4]) for x in g["tokens"]]:
            bad.append((i,len(t),float(score),g["score"]))
    print(name,len(gold),"bad",bad,"%.1fs"%(time.time()-t0))
