#!/usr/bin/env python3
"""Work counters of the byte model for the bench batches of every configuration, from the instrumented oracle.
Writes profiles/counters_r2.json: one entry per bench.py --config, aggregated over ALL rotating batches of that configuration
(config 5: the first 65536 of the 1 M sentences); bench.py turns them into algorithmic bytes per sentence."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from kiwi_b200.synth import SEED
from kiwi_b200 import bytemodel
from tests.orc import Oracle, TypoOracle, TYPO_IMAGES
from concurrent.futures import ProcessPoolExecutor


def count(args):
    cid, first_block, n = args
    cfg = bench.CONFIGS[cid]
    o = Oracle(bench.image_path(cfg["model"]))
    if cfg["typo"]: o.set_typo(TypoOracle(TYPO_IMAGES[cfg["typo"]]))
    for s in bench.gen_sentences(cfg, first_block, n, SEED): o.analyze(s)
    c = o.work_counters()
    if cfg["model"] == "cong": c.update(o.cong_counters())
    o.close()
    return c


def main():
    only = [int(x) for x in sys.argv[1:]] or sorted(bench.CONFIGS)
    path = os.path.join(ROOT, "profiles", "counters_r2.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for cid in only:
        cfg = bench.CONFIGS[cid]
        nblk = (cfg["batch"] + bench.BLOCK - 1) // bench.BLOCK
        if cfg["scaling"] == "strong":
            tasks = [(cid, b, bench.BLOCK) for b in range(8)]; coverage = "first 65536 of %d sentences" % cfg["batch"]
        else:
            tasks = [(cid, r * nblk + b if r else b, min(bench.BLOCK, cfg["batch"] - b * bench.BLOCK)) for r in range(cfg["rotate"]) for b in range(nblk)]
            tasks = [(cid, (1000 * 0 + r) * nblk + b, n) for r in range(cfg["rotate"]) for b, n in [(b, min(bench.BLOCK, cfg["batch"] - b * bench.BLOCK)) for b in range(nblk)]]
            coverage = "all %d rotating batches of rank 0, %d sentences" % (cfg["rotate"], cfg["rotate"] * cfg["batch"])
        t0 = time.time()
        with ProcessPoolExecutor(min(8, len(tasks))) as ex: parts = list(ex.map(count, tasks))
        agg = {}
        for p in parts:
            for k, v in p.items(): agg[k] = agg.get(k, 0) + v
        agg["oracle_seconds"] = time.time() - t0
        extra = bytemodel.cong_bytes(agg) if cfg["model"] == "cong" else 0.0
        agg["lattice_bytes_per_sentence"] = bytemodel.lattice_bytes(agg) / agg["sentences"]
        agg["viterbi_bytes_per_sentence"] = (bytemodel.viterbi_bytes(agg) + extra) / agg["sentences"]
        agg["coverage"] = coverage
        out["cfg%d" % cid] = agg
        print(cid, agg, flush=True)
        json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
