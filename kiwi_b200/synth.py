"""Seeded synthetic Korean batches with the length distribution of the reference's eval_data/web.txt
(SURVEY.md section 8d): draw a target length from the empirical first-column UTF-16 lengths of web.txt, then
concatenate whole eojeols (space separated surface words) sampled with replacement from web.txt + written.txt
until the target is reached, cutting at a word boundary.  The source sentences are the committed fixtures
tests/golden/inputs_web.txt / inputs_written.txt (first columns of the reference's eval files)."""
import os
import random
from typing import List

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0x4B315731          # "K1W1"


def _lines(name):
    with open(os.path.join(_ROOT, "tests", "golden", name), encoding="utf-8") as f:
        return [l.rstrip("\n") for l in f if l.strip()]


def u16len(s: str) -> int:
    return len(s.encode("utf-16-le", "surrogatepass")) // 2


def synth_batch(n: int, seed: int = SEED) -> List[str]:
    web = _lines("inputs_web.txt")
    written = _lines("inputs_written.txt")
    lengths = [u16len(l) for l in web]
    words = [w for l in web + written for w in l.split(" ") if w]
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        target = rng.choice(lengths)
        parts, cur = [], 0
        while cur < target:
            w = rng.choice(words)
            add = u16len(w) + (1 if parts else 0)
            if parts and cur + add > target:
                break
            parts.append(w)
            cur += add
        out.append(" ".join(parts))
    return out
