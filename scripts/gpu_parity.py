#!/usr/bin/env python3
"""GPU parity driver: CUDA path (through the C ABI) vs the oracle restatement on a text file (one sentence per
line, optional tab-separated gold column ignored).  Prints mismatch statistics and the first differences."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kiwi_b200
from tests.orc import Oracle, IMAGE

def read_lines(path, limit=None):
    out = []
    with open(path, encoding="utf-8", errors="surrogatepass") as f:
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if "\t" in line: line = line.split("\t")[0]
            out.append(line)
            if limit and len(out) >= limit: break
    return out

def main():
    path = sys.argv[1]
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else None
    check_lattice = "--lattice" in sys.argv
    texts = read_lines(path, limit)
    kw = kiwi_b200.Kiwi(IMAGE)
    orc = Oracle()
    t0 = time.time()
    res = kw.analyze_batch(texts)
    t1 = time.time()
    print("gpu batch: %d sentences in %.3f s; timings %s" % (len(texts), t1 - t0, res.timings_ms))
    bad_tok = bad_score = bad_lat = 0
    shown = 0
    for i, t in enumerate(texts):
        otoks, oscore = orc.analyze(t)
        g = res.sentence(i)
        gt = [(int(k["morph_id"]), int(k["tag"]), int(k["position"]), int(k["length"])) for k in g]
        ot = [x[:4] for x in otoks]
        tok_ok = gt == ot
        score_ok = abs(float(res.scores[i]) - oscore) <= 1e-4 * max(1.0, abs(oscore))
        exact = np.float32(res.scores[i]) == np.float32(oscore)
        if not tok_ok: bad_tok += 1
        if not score_ok: bad_score += 1
        lat_ok = True
        if check_lattice:
            gl = kw.debug_lattice(t); ol = orc.lattice(t)
            lat_ok = gl.shape == ol.shape and bool((gl == ol).all())
            if not lat_ok: bad_lat += 1
        if (not tok_ok or not score_ok or not lat_ok) and shown < 8:
            shown += 1
            print("MISMATCH line %d: %r" % (i, t[:60]))
            print("  gpu score %r oracle %r exact=%s" % (float(res.scores[i]), oscore, exact))
            if not tok_ok:
                for a, b in zip(gt, ot):
                    if a != b: print("   first diff gpu", a, "oracle", b); break
                print("   ntok gpu %d oracle %d" % (len(gt), len(ot)))
            if not lat_ok:
                print("   lattice shapes", gl.shape, ol.shape)
                for r in range(min(len(gl), len(ol))):
                    if (gl[r] != ol[r]).any(): print("   first lattice diff row", r, gl[r].tolist(), ol[r].tolist()); break
    print(json.dumps(dict(file=os.path.basename(path), n=len(texts), token_mismatch=bad_tok, score_mismatch=bad_score, lattice_mismatch=bad_lat)))

if __name__ == "__main__":
    main()
