#!/usr/bin/env python3
"""Parallel sweep of the 32-lane host simulation over golden sentences (no GPU): every sentence of the named golden input
files through lattice.cu + viterbi.cu + emit.cu compiled as C++ (tests/hostsim), tokens and bit-exact scores against the
unmodified reference's vectors.  The CPU-side regression net for kernel edits.

  python scripts/hostsim_sweep.py [plain|typo|cong] [--files inputs_web,inputs_written,...] [--stride K] [--jobs J]"""
import ctypes as C, os, sys, time, argparse
import numpy as np
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.goldenio import read_golden, read_inputs
from tests.orc import IMAGE, TYPO_IMAGES, CONG_IMAGE, SBG_IMAGE

MATCH_ALL = (1 << 0) | (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 5) | (1 << 23) | (1 << 16)


_H = {}      # one simulator handle per worker process and mode (opening the 135 MB image is the expensive part)


def _handle(mode):
    if mode in _H: return _H[mode]
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostsim", os.environ.get("HS32_LIB", "libpipeline_sim32.so")))      # HS32_LIB=libpipeline_sim32_redo.so: forced group redo
    lib.hs32_open.restype = C.c_void_p; lib.hs32_open.argtypes = [C.c_char_p]
    lib.hs32_set_typo.argtypes = [C.c_void_p, C.c_char_p, C.c_float]
    lib.hs32_analyze.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p]
    h = lib.hs32_open(os.fsencode(CONG_IMAGE if mode == "cong" else SBG_IMAGE if mode == "sbg" else IMAGE))
    assert h
    if mode == "typo": assert lib.hs32_set_typo(h, os.fsencode(TYPO_IMAGES["basic"]), 2.5) == 0
    if mode == "block":      # AnalyzeOption::blocklist: the morpheme ids the reference resolved for the vectors (tests/golden/MANIFEST.json)
        import json
        ids = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "MANIFEST.json")))["blocklist"]["morpheme_ids"], dtype=np.uint32)
        lib.hs32_set_blocklist.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        assert lib.hs32_set_blocklist(h, ids.ctypes.data, len(ids)) == 0
    _H[mode] = (lib, h)
    return _H[mode]


_BENCH = {}


def _fuzz_texts(n, seed=20240917):
    """random strings that look nothing like the bench text: random Hangul syllables / jamo / compatibility jamo, ASCII words, digits and
    punctuation runs, emoji and surrogate pairs, URLs / mentions / hashtags, odd whitespace, repeated characters, lengths 0 .. 120"""
    rs = np.random.RandomState(seed)
    pools = [lambda: "".join(chr(0xAC00 + int(rs.randint(0, 11172))) for _ in range(rs.randint(1, 8))),
             lambda: "".join(chr(0xAC00 + 588 * int(rs.randint(0, 19)) + 28 * int(rs.randint(0, 21))) for _ in range(rs.randint(1, 6))),
             lambda: "".join(chr(int(rs.choice([0x1100, 0x1161, 0x11A8, 0x3131, 0x314F, 0x3147])) + int(rs.randint(0, 12))) for _ in range(rs.randint(1, 4))),
             lambda: "".join(chr(int(rs.randint(97, 123))) for _ in range(rs.randint(1, 9))),
             lambda: "".join(chr(int(rs.randint(48, 58))) for _ in range(rs.randint(1, 7))) + rs.choice(["", ".", ",000", "%", "년", "개", "-1"]),
             lambda: rs.choice([".", "..", "...", "!", "?!", ",", "\"", "'", "(", ")", "[", "]", "~", "-", "·", "ㅋㅋㅋ", "ㅠㅠ", "^^", "※", "①", "●", "1)"]),
             lambda: rs.choice(["\U0001F600", "\U0001F44D\U0001F3FD", "\u2764\uFE0F", "\U0001F1F0\U0001F1F7", "\u263A"]),
             lambda: rs.choice(["https://a.b/c?d=1", "www.kiwi.co.kr", "a@b.com", "@user_1", "#태그", "#tag2", "010-1234-5678", "2024.01.02.", "3:45"]),
             lambda: rs.choice(["하다", "했다", "합니다", "이다", "것", "수", "있다", "없다", "에서", "으로", "는", "을", "를", "이", "가", "도", "만", "요", "죠", "네요"]),
             lambda: chr(0xAC00 + int(rs.randint(0, 11172))) * int(rs.randint(2, 12))]
    seps = [" ", " ", " ", "", "", "  ", "\t", "\n", "\u00A0", "\u3000"]
    out = []
    for _ in range(n):
        k = int(rs.randint(0, 14)); t = ""
        for _j in range(k):
            t += pools[int(rs.randint(0, len(pools)))]() + seps[int(rs.randint(0, len(seps)))]
            if len(t) > 120: break
        out.append(t[:120])
    return out


def _bench_case(mode, n):
    """sentences of the bench batch of this mode + the ORACLE's analysis of them (no golden vectors exist for synthetic text)"""
    key = (mode, n)
    if key not in _BENCH:
        import bench
        from kiwi_b200.synth import SEED
        from tests.orc import Oracle, TypoOracle
        cfg = bench.CONFIGS[{"plain": 2, "cong": 3, "typo": 4}[mode]]
        o = Oracle(bench.image_path(cfg["model"]))
        if cfg["typo"]: o.set_typo(TypoOracle(TYPO_IMAGES[cfg["typo"]]))
        _BENCH[key] = (bench.gen_sentences(cfg, 0, n, SEED), o)
    return _BENCH[key]


def work(args):
    mode, name, idxs = args
    cong = mode == "cong"; typo = mode == "typo"
    lib, h = _handle(mode)
    if name.startswith("bench:") or name.startswith("fuzz:"):
        if name.startswith("fuzz:"):
            _, orc = _bench_case(mode, 1); texts = _fuzz_texts(int(name[5:]))
        else:
            texts, orc = _bench_case(mode, int(name[6:]))
        cap = 8192
        morph = np.zeros(cap, np.uint32); tag = np.zeros(cap, np.uint8); pos = np.zeros(cap, np.uint32); ln = np.zeros(cap, np.uint16); sc = np.zeros(cap, np.float32)
        bad = []
        for i in idxs:
            t = texts[i]
            u = np.ascontiguousarray(np.frombuffer(t.encode("utf-16-le", "surrogatepass"), dtype="<u2"))
            s = C.c_float(0); nn = C.c_int(0)
            n = lib.hs32_analyze(h, u.ctypes.data, len(u), MATCH_ALL, morph.ctypes.data, tag.ctypes.data, pos.ctypes.data, ln.ctypes.data, sc.ctypes.data, cap, C.byref(s), C.byref(nn), None)
            otoks, oscore = orc.analyze(t)
            if n < 0: bad.append((name, i, "status %d" % n)); continue
            got = [(int(morph[k]), int(tag[k]), int(pos[k]), int(ln[k])) for k in range(n)]
            if got != [x[:4] for x in otoks]: bad.append((name, i, "tokens")); continue
            if not (all(np.float32(sc[k]) == np.float32(otoks[k][4]) for k in range(n)) and np.float32(s.value) == np.float32(oscore)): bad.append((name, i, "scores"))
        return len(idxs), bad
    cap = 8192
    morph = np.zeros(cap, np.uint32); tag = np.zeros(cap, np.uint8); pos = np.zeros(cap, np.uint32); ln = np.zeros(cap, np.uint16); sc = np.zeros(cap, np.float32)
    texts = read_inputs(name); gold = read_golden(("cong_" if cong else "") + ("sbg_" if mode == "sbg" else "") + ("open_" if mode == "open" else "") + ("block_" if mode == "block" else "") + ("typo6_" if typo else "") + name)
    bad = []
    for i in idxs:
        t, g = texts[i], gold[i]
        u = np.ascontiguousarray(np.frombuffer(t.encode("utf-16-le", "surrogatepass"), dtype="<u2"))
        s = C.c_float(0); nn = C.c_int(0)
        n = lib.hs32_analyze(h, u.ctypes.data, len(u), MATCH_ALL | (0x80000000 if mode == "open" else 0), morph.ctypes.data, tag.ctypes.data, pos.ctypes.data, ln.ctypes.data, sc.ctypes.data, cap, C.byref(s), C.byref(nn), None)
        if n < 0: bad.append((name, i, "status %d" % n)); continue
        got = [(int(morph[k]), int(tag[k]), int(pos[k]), int(ln[k])) for k in range(n)]
        if got != [x[:4] for x in g["tokens"]]: bad.append((name, i, "tokens")); continue
        if not (all(np.float32(sc[k]) == np.float32(g["tokens"][k][4]) for k in range(n)) and np.float32(s.value) == np.float32(g["score"])): bad.append((name, i, "scores"))
    return len(idxs), bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", nargs="?", default="plain", choices=["plain", "typo", "cong", "sbg", "open", "block"])
    ap.add_argument("--files", default="inputs_web,inputs_written,inputs_ref_tests,inputs_dialect_typos")
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--maxlen", type=int, default=400, help="skip longer inputs (the pathological reference tests take minutes)")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    a = ap.parse_args()
    tasks = []
    for name in a.files.split(","):
        if name.startswith("bench:") or name.startswith("fuzz:"):      # --files bench:8192 : the first N sentences of the mode's bench batch against the oracle; fuzz:N : N random strings
            idx = list(range(0, int(name.split(":")[1]), a.stride)); per = max(1, len(idx) // (a.jobs * 4))
            for k in range(0, len(idx), per): tasks.append((a.mode, name, idx[k:k + per]))
            continue
        texts = read_inputs(name)
        nGold = len(read_golden("sbg_" + name)) if a.mode == "sbg" else len(texts)      # (the sbg vectors of inputs_ref_tests stop before the pathological inputs)
        idx = [i for i in range(0, min(len(texts), nGold), a.stride) if len(texts[i]) <= a.maxlen]
        per = max(1, len(idx) // (a.jobs * 4))
        for k in range(0, len(idx), per): tasks.append((a.mode, name, idx[k:k + per]))
    t0 = time.time(); total = 0; bad = []
    with ProcessPoolExecutor(a.jobs) as ex:
        for n, b in ex.map(work, tasks): total += n; bad += b
    print("%s: %d sentences, %d mismatches, %.1f s" % (a.mode, total, len(bad), time.time() - t0))
    for b in bad[:40]: print("  MISMATCH", b)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
