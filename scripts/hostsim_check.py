#!/usr/bin/env python3
"""Runs sentences of a golden input file through the 32-lane host simulation of the device pipeline
(tests/hostsim/libpipeline_sim32.so: lattice.cu + viterbi.cu + emit.cu compiled as C++, one OS thread per lane) and compares
tokens and scores with the reference's golden vectors.  No GPU needed; ~1-60 s per sentence.

  python scripts/hostsim_check.py {plain|typo} <inputs name> <line> [<line> ...]      e.g.  typo inputs_dialect_typos 98 469
  HS32_MODEL=cong   use the CoNg model image and the cong_* vectors
  HS32_ASCENDING=1  schedule the lowest runnable lane first (default: highest first); HS32_TRACE=1  per-stage trace on stderr
A divergent collective or a lane that never reaches one aborts with the call sites (addr2line -e tests/hostsim/libpipeline_sim32.so)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.goldenio import read_golden, read_inputs
from tests.orc import IMAGE, TYPO_IMAGES, CONG_IMAGE


def main():
    if len(sys.argv) < 4 or sys.argv[1] not in ("plain", "typo"):
        print(__doc__); return 2
    cong = os.environ.get("HS32_MODEL", "") == "cong"
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libpipeline_sim32.so"))
    lib.hs32_open.restype = C.c_void_p; lib.hs32_open.argtypes = [C.c_char_p]
    lib.hs32_set_typo.argtypes = [C.c_void_p, C.c_char_p, C.c_float]
    lib.hs32_analyze.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p]
    h = lib.hs32_open(os.fsencode(CONG_IMAGE if cong else IMAGE))
    assert h, "cannot open the model image: run __graft_entry__.build()"
    match_all = (1 << 0) | (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 5) | (1 << 23) | (1 << 16)
    cap = 4096
    morph = np.zeros(cap, np.uint32); tag = np.zeros(cap, np.uint8); pos = np.zeros(cap, np.uint32); ln = np.zeros(cap, np.uint16); sc = np.zeros(cap, np.float32)
    typo = sys.argv[1] == "typo"
    if typo:
        assert lib.hs32_set_typo(h, os.fsencode(TYPO_IMAGES["basic"]), 2.5) == 0
    name = sys.argv[2]
    texts = read_inputs(name); gold = read_golden(("cong_" if cong else "") + ("typo6_" if typo else "") + name)
    bad = 0
    for i in [int(x) for x in sys.argv[3:]]:
        t, g = texts[i], gold[i]
        u = np.ascontiguousarray(np.frombuffer(t.encode("utf-16-le", "surrogatepass"), dtype="<u2"))
        s = C.c_float(0); nn = C.c_int(0)
        t0 = time.time()
        n = lib.hs32_analyze(h, u.ctypes.data, len(u), match_all, morph.ctypes.data, tag.ctypes.data, pos.ctypes.data, ln.ctypes.data, sc.ctypes.data, cap,
                             C.byref(s), C.byref(nn), None)
        dt = time.time() - t0
        if n < 0:
            print(i, "FAILED with status", n, repr(t[:40])); bad += 1; continue
        got = [(int(morph[k]), int(tag[k]), int(pos[k]), int(ln[k])) for k in range(n)]
        ok = got == [x[:4] for x in g["tokens"]]
        exact = ok and all(np.float32(sc[k]) == np.float32(g["tokens"][k][4]) for k in range(n)) and np.float32(s.value) == np.float32(g["score"])
        bad += not exact
        print(i, "len", len(u), "nodes", nn.value, "tokens", n, "ok" if ok else "TOKENS DIFFER", "scores exact" if exact else "scores differ", "%.1fs" % dt, flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
