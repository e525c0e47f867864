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


def synth_batch(n: int, seed: int = SEED, typo_frac: float = 0.0) -> List[str]:
    """typo_frac > 0 (BASELINE.json config 4: 0.3): that share of the sentences draws its eojeols from web_with_typos.txt instead
    (tests/golden/inputs_web_typos.txt); with typo_frac == 0 the random stream is the one round 1 used."""
    web = _lines("inputs_web.txt")
    written = _lines("inputs_written.txt")
    lengths = [u16len(l) for l in web]
    words = [w for l in web + written for w in l.split(" ") if w]
    typo_words = [w for l in _lines("inputs_web_typos.txt") for w in l.split(" ") if w] if typo_frac > 0 else None
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        target = rng.choice(lengths)
        src = typo_words if (typo_words and rng.random() < typo_frac) else words
        parts, cur = [], 0
        while cur < target:
            w = rng.choice(src)
            add = u16len(w) + (1 if parts else 0)
            if parts and cur + add > target:
                break
            parts.append(w)
            cur += add
        out.append(" ".join(parts))
    return out
