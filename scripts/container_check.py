#!/usr/bin/env python3
"""Container-level check of the Knlm kernel (no GPU): per-node path counts of the simulated kernel (tests/hostsim) against the oracle in its
two container modes - the reference's insertOptimized as it BEHAVES (default) and as its comments intend (ORC_BUCKET_SEARCH_ALL=1) - on
bench sentences where the two differ (a bucket reaches 64 states and a state recurs: DESIGN.md section 2 item 3).

  python scripts/container_check.py [idx,idx,...]      default: the six sentences of the first 4096 of the config-2 batch where they differ
  python scripts/container_check.py scan N             lists the sentences among the first N where the oracle's two modes differ"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEFAULT = [785, 2344, 2348, 2989, 3209, 4059]


def texts(n=4096):
    from kiwi_b200.synth import synth_batch
    return synth_batch(n)


def child(role, idx):
    T = texts(max(idx) + 1 if max(idx) >= 4096 else 4096)
    if role == "hs":
        os.environ["HS32_TRACE"] = "1"
        import numpy as np, ctypes as C
        from scripts.hostsim_sweep import _handle, MATCH_ALL
        lib, h = _handle("plain"); cap = 8192
        morph = np.zeros(cap, np.uint32); tag = np.zeros(cap, np.uint8); pos = np.zeros(cap, np.uint32); ln = np.zeros(cap, np.uint16); sc = np.zeros(cap, np.float32)
        for i in idx:
            sys.stderr.write("SENT %d\n" % i); sys.stderr.flush()
            u = np.ascontiguousarray(np.frombuffer(T[i].encode("utf-16-le", "surrogatepass"), dtype="<u2")); s = C.c_float(0); nn = C.c_int(0)
            lib.hs32_analyze(h, u.ctypes.data, len(u), MATCH_ALL, morph.ctypes.data, tag.ctypes.data, pos.ctypes.data, ln.ctypes.data, sc.ctypes.data, cap, C.byref(s), C.byref(nn), None)
    else:
        os.environ["ORC_TRACE_NODES"] = "1"
        if role == "all": os.environ["ORC_BUCKET_SEARCH_ALL"] = "1"
        from tests.orc import Oracle, IMAGE
        o = Oracle(IMAGE)
        for i in idx:
            sys.stderr.write("SENT %d\n" % i); sys.stderr.flush()
            o.analyze(T[i])


def run(role, idx):
    txt = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", role, ",".join(map(str, idx))], capture_output=True, text=True).stderr
    out = {}; cur = None
    for line in txt.splitlines():
        if line.startswith("SENT "): cur = int(line[5:]); out[cur] = []
        elif "<- [orc]" in line: out[cur] += [int(x) for x in line.split("<-")[0].split()]
        elif line.startswith("[hs32] paths per node:"): out[cur] += [int(x) for x in line.split(":")[1].split()]
    return out


def contains(seq, sub):      # the simulator also lists every chunk's begin node (1 path) and end node (0): the oracle's counts are a subsequence
    it = iter(seq)
    return all(any(v == w for w in it) for v in sub)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child": return child(sys.argv[2], [int(x) for x in sys.argv[3].split(",")])
    if len(sys.argv) > 2 and sys.argv[1] == "scan":
        n = int(sys.argv[2]); idx = list(range(n))
        q = run("quirk", idx); a = run("all", idx)
        print("sentences whose per-node path counts differ between the two container modes:", [i for i in idx if q.get(i) != a.get(i)])
        return 0
    idx = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else DEFAULT
    hs = run("hs", idx); q = run("quirk", idx); a = run("all", idx)
    good = 0
    for i in idx:
        follows = contains(hs[i], q[i]) and sum(hs[i]) - sum(q[i]) <= 64
        intended = contains(hs[i], a[i]) and q[i] != a[i]
        print("sentence %d: paths over all nodes - kernel %d, reference behaviour %d, intended behaviour %d -> %s" % (i, sum(hs[i]), sum(q[i]), sum(a[i]), "follows the reference" if follows and not intended else "DIFFERS"))
        good += int(follows and not intended)
    print("container check: %d/%d sentences follow the reference's behaviour" % (good, len(idx)))
    return 0 if good == len(idx) else 1


if __name__ == "__main__":
    sys.exit(main())
