#!/usr/bin/env python3
"""GPU-box diagnostic for the CoNg path (writes gpurun_out/cong_diag.txt): stage-level self-test of the int8 scorer
(dp4a dots, epilogues, context-trie steps, tensor-core tile) against the oracle, then every golden sentence whose tokens
or score differ from the reference vectors, with both token lists.  Never a bench; tests/ hold the assertions."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import kiwi_b200
from tests.orc import Oracle, CONG_IMAGE
from tests.goldenio import read_golden, read_inputs, read_cong_qgemm

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "cong_diag.txt"), "w")
def P(*a):
    print(*a, file=out); out.flush()

try:
    kw = kiwi_b200.Kiwi(CONG_IMAGE)
    orc = Oracle(CONG_IMAGE)
    # ---- stage level
    recs = [r for r in read_cong_qgemm() if r[0] == "P"]
    ctx = [r[2] for r in recs]; wid = [r[3] for r in recs]; node = [r[1] for r in recs]
    o = kw.debug_cong(ctx, wid, node)
    bad_ll = [i for i, r in enumerate(recs) if np.float32(o["eps"][i, 0]) != np.float32(r[4])]
    bad_st = [i for i, r in enumerate(recs) if (int(o["node"][i]), int(o["ctx"][i])) != (r[5], r[6])]
    bad_dot = []; bad_eps = []
    for i in range(len(recs)):
        acc, eps = orc.cong_pair(ctx[i], wid[i])
        if int(o["dot"][i]) != acc: bad_dot.append(i)
        if [np.float32(x) for x in o["eps"][i]] != eps: bad_eps.append(i)
    P("stage: scalar ll mismatches %d, step mismatches %d, dot mismatches %d, epilogue mismatches %d of %d" % (len(bad_ll), len(bad_st), len(bad_dot), len(bad_eps), len(recs)))
    for i in (bad_dot[:5] + bad_st[:5] + bad_eps[:5]):
        acc, eps = orc.cong_pair(ctx[i], wid[i])
        P("  i", i, "in", (node[i], ctx[i], wid[i]), "gpu dot", int(o["dot"][i]), "want", acc, "gpu eps", [float(x).hex() for x in o["eps"][i]], "want", [float(x).hex() for x in eps],
          "gpu next", (int(o["node"][i]), int(o["ctx"][i])), "want", (recs[i][5], recs[i][6]))
    rs = np.random.RandomState(11)
    for n in (64, 40, 17, 5):
        c = rs.randint(0, 8192, size=n); w = rs.randint(0, 60000, size=n)
        t = kw.debug_cong(c, w, np.zeros(n, np.int32))["tile"]
        want = np.array([[orc.cong_pair(int(c[r]), int(w[k]))[0] for k in range(min(n, 32))] for r in range(min(n, 64))], np.int32)
        nb = int((t != want).sum())
        P("tile n=%d: %d mismatching entries of %d" % (n, nb, want.size))
        if nb:
            r, k = np.argwhere(t != want)[0]
            P("  first at row %d col %d: gpu %d want %d;  row0 gpu %s want %s" % (r, k, t[r, k], want[r, k], t[0, :8].tolist(), want[0, :8].tolist()))
            # is it a permutation of the expected layout?  print where want[0,0] and want[1,0], want[0,1] landed
            for (rr, kk) in [(0, 0), (1, 0), (0, 1), (8, 0), (0, 2)]:
                if rr < want.shape[0] and kk < want.shape[1]:
                    P("  want[%d,%d]=%d found at" % (rr, kk, want[rr, kk]), np.argwhere(t == want[rr, kk])[:4].tolist())
    # ---- whole analyses
    for name in ["inputs_web", "inputs_written", "inputs_ref_tests", "inputs_dialect_typos"]:
        texts = read_inputs(name); gold = read_golden("cong_" + name)
        try:
            res = kw.analyze_batch(texts)
        except Exception as e:
            P(name, "analyze_batch failed:", repr(e))
            # find the offending sentences one by one (bounded)
            nfail = 0
            for i, t in enumerate(texts):
                try: kw.analyze_batch([t])
                except Exception as e2:
                    nfail += 1
                    if nfail <= 8: P("  sentence", i, repr(t), "->", repr(e2))
            P("  failing sentences:", nfail)
            continue
        ntok = nscore = nexact = 0
        for i, (t, g) in enumerate(zip(texts, gold)):
            got = [(int(k["morph_id"]), int(k["tag"]), int(k["position"]), int(k["length"])) for k in res.sentence(i)]
            want = [x[:4] for x in g["tokens"]]
            if got != want:
                ntok += 1
                if ntok <= 12:
                    d = next((k for k in range(min(len(got), len(want))) if got[k] != want[k]), min(len(got), len(want)))
                    P("TOK", name, i, repr(t)); P("  first diff at token", d, "gpu", got[max(0, d - 1):d + 3], "want", want[max(0, d - 1):d + 3], "scores gpu %r want %r" % (float(res.scores[i]), g["score"]))
            elif abs(float(res.scores[i]) - g["score"]) > 1e-4 * max(1.0, abs(g["score"])):
                nscore += 1
                if nscore <= 8: P("SCORE", name, i, repr(t), float(res.scores[i]), g["score"])
            else:
                nexact += int(np.float32(res.scores[i]) == np.float32(g["score"]))
        P("%s: %d sentences, token mismatches %d, score mismatches %d, bit-exact scores %d" % (name, len(texts), ntok, nscore, nexact))
except Exception:
    P("diag crashed:\n" + traceback.format_exc())
out.close()
print(open(os.path.join(ROOT, "gpurun_out", "cong_diag.txt")).read()[-3000:])
