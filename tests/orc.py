"""ctypes access to the ORACLE (oracle/liboracle.so, the CPU restatement).  Test infrastructure only."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")
IMAGE = os.path.join(ROOT, "oracle", "_ref", "models", "knlm_small.img")
CONG_IMAGE = os.path.join(ROOT, "oracle", "_ref", "models", "cong_small.img")
SBG_IMAGE = os.path.join(ROOT, "oracle", "_ref", "models", "sbg_small.img")


TYPO_IMAGES = {"kat": os.path.join(ROOT, "oracle", "_ref", "models", "typo_kat.img"), "basic": os.path.join(ROOT, "oracle", "_ref", "models", "typo_basic.img")}


class TypoOracle:
    """restated PreparedTypoTransformer::generateGraph over a flat typo image (oracle/restate/typo.hpp)"""
    def __init__(self, image_path):
        self.lib = C.CDLL(ORACLE_LIB)
        self.lib.orc_typo_open.restype = C.c_void_p
        self.lib.orc_typo_open.argtypes = [C.c_char_p]
        self.lib.orc_typo_close.argtypes = [C.c_void_p]
        self.lib.orc_typo_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        self.h = self.lib.orc_typo_open(os.fsencode(image_path))
        if not self.h:
            raise RuntimeError("oracle: cannot open typo image " + image_path)

    def graph(self, text: str):
        u = np.ascontiguousarray(np.frombuffer(text.encode("utf-16-le", "surrogatepass"), dtype="<u2"))
        rows = np.zeros((8192, 9), np.int32); nl = C.c_int(0)
        n = self.lib.orc_typo_graph(self.h, u.ctypes.data, len(u), rows.ctypes.data, len(rows), C.byref(nl))
        if n < 0:
            raise RuntimeError("oracle typo graph failed")
        return nl.value, rows[:n].tolist()

    def close(self):
        if self.h:
            self.lib.orc_typo_close(self.h); self.h = None


class Oracle:
    def __init__(self, image_path=IMAGE):
        self.lib = C.CDLL(ORACLE_LIB)
        self.lib.orc_open.restype = C.c_void_p
        self.lib.orc_open.argtypes = [C.c_char_p]
        self.lib.orc_close.argtypes = [C.c_void_p]
        self.lib.orc_analyze.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(C.c_float)]
        self.lib.orc_lattice.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self.lib.orc_counters.argtypes = [C.c_void_p, C.c_void_p]
        self.h = self.lib.orc_open(os.fsencode(image_path))
        if not self.h:
            raise RuntimeError("oracle: cannot open image " + image_path)
        self.cap = 1 << 16
        self.morph = np.zeros(self.cap, np.uint32); self.tag = np.zeros(self.cap, np.uint8); self.pos = np.zeros(self.cap, np.uint32)
        self.length = np.zeros(self.cap, np.uint16); self.score = np.zeros(self.cap, np.float32)

    def analyze(self, text: str):
        blob = np.ascontiguousarray(np.frombuffer(text.encode("utf-16-le", "surrogatepass"), dtype="<u2"))
        s = C.c_float(0)
        n = self.lib.orc_analyze(self.h, blob.ctypes.data, len(blob), self.morph.ctypes.data, self.tag.ctypes.data, self.pos.ctypes.data,
                                 self.length.ctypes.data, self.score.ctypes.data, self.cap, C.byref(s))
        if n < 0:
            raise RuntimeError("oracle analyze failed")
        toks = [(int(self.morph[i]), int(self.tag[i]), int(self.pos[i]), int(self.length[i]), float(self.score[i])) for i in range(n)]
        return toks, float(s.value)

    def lattice(self, text: str) -> np.ndarray:
        blob = np.ascontiguousarray(np.frombuffer(text.encode("utf-16-le", "surrogatepass"), dtype="<u2"))
        rows = np.zeros((1 << 16, 9), np.int32)
        n = self.lib.orc_lattice(self.h, blob.ctypes.data, len(blob), rows.ctypes.data, len(rows))
        if n < 0:
            raise RuntimeError("oracle lattice failed")
        return rows[:n].copy()

    def set_blocklist(self, ids):
        """AnalyzeOption::blocklist as morpheme ids (what kiwi_morphset_add resolves to)"""
        a = np.ascontiguousarray(np.asarray(list(ids), dtype=np.uint32))
        self.lib.orc_set_blocklist.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.lib.orc_set_blocklist(self.h, a.ctypes.data, len(a))

    def set_open_ending(self, on=True):
        """AnalyzeOption::openEnding (no end-of-sentence step on the last chunk)"""
        self.lib.orc_set_open_ending.argtypes = [C.c_void_p, C.c_int]
        self.lib.orc_set_open_ending(self.h, 1 if on else 0)

    def counters(self):
        out = np.zeros(6, np.uint64)
        self.lib.orc_counters(self.h, out.ctypes.data)
        return dict(zip(["lmSteps", "pairs", "inserts", "pathsOut", "candEvals", "evalCalls"], [int(x) for x in out]))

    def counters2(self):
        """{top1Mode, bucketFull, mediumMode, maxIncoming}: how often the > 512-path `top1` container, the 'container is full' rule and
        the 4-bucket medium container were used since open"""
        out = np.zeros(4, np.uint64)
        self.lib.orc_counters2.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.orc_counters2(self.h, out.ctypes.data)
        return dict(zip(["top1Mode", "bucketFull", "mediumMode", "maxIncoming"], [int(x) for x in out]))

    WORK_FIELDS = ["sentences", "rawUnits", "normUnits", "trieVisits", "trieProbes", "trieHits", "candForms", "nodesBuilt", "nodesFinal",
                   "candEntries", "candEvals", "lmSteps", "lmHops", "lmProbes", "pairs", "pathsWritten", "pathsKept", "tokens"]

    def work_counters(self):
        self.lib.orc_work_counters.argtypes = [C.c_void_p, C.c_void_p]
        out = np.zeros(18, np.uint64)
        self.lib.orc_work_counters(self.h, out.ctypes.data)
        return dict(zip(self.WORK_FIELDS, [int(x) for x in out]))

    # ---- CoNg scorer pieces (oracle/restate/cong.hpp)
    def cong_pair(self, ctx: int, wid: int):
        """-> (acc - hsum, [E_scalar, E_small, E_gemv]) for one (context row, output row) pair"""
        self.lib.orc_cong_pair.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        acc = C.c_int32(); eps = (C.c_float * 3)()
        if self.lib.orc_cong_pair(self.h, ctx, wid, C.byref(acc), eps) != 0:
            raise RuntimeError("oracle: not a CoNg image or index out of range")
        return acc.value, [np.float32(eps[k]) for k in range(3)]

    def cong_epilogue(self, m: int, n: int) -> int:
        self.lib.orc_cong_epilogue.argtypes = [C.c_int, C.c_int]
        return int(self.lib.orc_cong_epilogue(m, n))

    def cong_step(self, node: int, wid: int):
        """one context-trie transition -> (node', contextIdx')"""
        self.lib.orc_cong_step.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_uint32]
        self.lib.orc_cong_step.restype = C.c_uint32
        nd = C.c_int32(node)
        ctx = self.lib.orc_cong_step(self.h, C.byref(nd), wid)
        return nd.value, int(ctx)

    def cong_counters(self):
        self.lib.orc_cong_counters.argtypes = [C.c_void_p, C.c_void_p]
        out = np.zeros(2, np.uint64)
        self.lib.orc_cong_counters(self.h, out.ctypes.data)
        return {"cgRows": int(out[0]), "cgMacs": int(out[1])}

    def set_typo(self, typo, threshold: float = 2.5):
        """analyse with a typo lattice (AnalyzeOption::withTypoTransformer); `typo` is a TypoOracle or None"""
        self.lib.orc_set_typo.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        self._typo = typo       # keep the image alive
        self.lib.orc_set_typo(self.h, typo.h if typo else None, threshold)

    def reset_history(self):
        """fresh `top1` container (the reference's thread_local unordered_set of a new process)"""
        self.lib.orc_reset_history.argtypes = [C.c_void_p]
        self.lib.orc_reset_history(self.h)

    def close(self):
        if self.h:
            self.lib.orc_close(self.h); self.h = None
