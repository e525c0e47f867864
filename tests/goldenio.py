"""Reader for the golden dumps written by oracle/ref_build/tools/dump_golden.cpp (test infrastructure)."""
import gzip, os

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_inputs(name):
    with open(os.path.join(HERE, name + ".txt"), encoding="utf-8", errors="surrogatepass") as f:
        return [l.rstrip("\n").rstrip("\r").split("\t")[0] for l in f]


def read_golden(name):
    """-> list of dict(tokens=[(morph, tag, pos, len, score)], score=float, lattice=[[9 ints]], chunks=[...])"""
    out = []
    cur = None; chunk_idx = -1; kept_chunk = -1
    with gzip.open(os.path.join(HERE, name + ".golden.txt.gz"), "rt") as f:
        for line in f:
            p = line.split()
            k = p[0]
            if k == "S":
                cur = dict(tokens=[], score=float.fromhex(p[3]), lattice=[], paths=[], norm_len=int(p[5]))
                out.append(cur); kept_chunk = -1
            elif k == "T":
                cur["tokens"].append((int(p[1]) & 0xFFFFFFFF, int(p[2]), int(p[3]), int(p[4]), float.fromhex(p[5])))
                if len(p) > 6:      # (wordPosition, sentPosition, lineNumber, subSentPosition, pairedToken) and the surface form (TokenInfo::str)
                    cur.setdefault("forms", []).append((tuple(int(x) for x in p[6].split(",")), line.rstrip("\n").split(" ", 7)[7] if len(p) > 7 else ""))
            elif k == "Y":      # typo-tolerant dumps: TokenInfo::typoCost per token
                cur["typo_costs"] = [float.fromhex(x) for x in p[1:]]
            elif k == "C":
                n_nodes = int(p[4])
                cur["_keep"] = n_nodes > 2
                if cur["_keep"]: kept_chunk += 1
            elif k == "N":
                if cur["_keep"]:
                    form, uoff, ulen, prev, sib, st, en, se = (int(x) for x in p[1:9])
                    cur["lattice"].append([form, uoff if ulen else -1, ulen, prev, sib, st, en, se, kept_chunk])
            elif k == "P":
                cur["paths"].append(dict(score=float.fromhex(p[1]), prev=int(p[2]), cur=int(p[3]), toks=[]))
            elif k == "K":
                cur["paths"][-1]["toks"].append((int(p[1]), int(p[2]), int(p[3]), float.fromhex(p[4])))
    return out


def read_cong_qgemm():
    """records of tests/golden/cong_qgemm.golden.txt.gz: ("Q", m, n, ctxIds, outIds, floats) | ("P", node, ctx, wid, ll, node2, ctx2)"""
    with gzip.open(os.path.join(HERE, "cong_qgemm.golden.txt.gz"), "rt") as f:
        for line in f:
            p = line.split()
            if p[0] == "Q":
                m, n = int(p[1]), int(p[2])
                yield ("Q", m, n, [int(x) for x in p[3:3 + m]], [int(x) for x in p[3 + m:3 + m + n]], [float.fromhex(x) for x in p[3 + m + n:]])
            else:
                yield ("P", int(p[1]), int(p[2]), int(p[3]), float.fromhex(p[4]), int(p[5]), int(p[6]))


def read_typo_graphs(key):
    """tests/golden/<key>.golden.txt.gz -> list of (normLen, [rows of 9 ints {endPos, typoCost bits, prev, sibling, continualTypoIdx, dialect, fromPool, off, len}])"""
    import struct
    out = []
    with gzip.open(os.path.join(HERE, key + ".golden.txt.gz"), "rt") as f:
        for line in f:
            p = line.split()
            if p[0] == "G":
                out.append((int(p[2]), []))
            else:
                bits = struct.unpack("<i", struct.pack("<f", float.fromhex(p[2])))[0]
                out[-1][1].append([int(p[1]), bits, int(p[3]), int(p[4]), int(p[5]), int(p[6]), 1 if p[7] == "R" else 0, int(p[8]), int(p[9])])
    return out
