#!/usr/bin/env python3
"""Rewrite the reference's gold evaluation files (`text \t form/TAG form/TAG ...`, eval_data/*.txt) into the
tab-separated trainer format parsed at /root/reference/src/KiwiBuilder.cpp:767-847
(`<ignored> \t form \t TAG \t form \t TAG ...`, blank line = end of sentence), and optionally append a
seeded synthetic corpus so that the fabricated Knlm gets a realistic n-gram count.

TEST INFRASTRUCTURE (model fabrication; the LFS model binaries are absent from the snapshot).
usage: gold_to_corpus.py OUT.txt [--synth N --seed S] GOLD.txt...
"""
import sys, random

def parse_gold(path):
    sents = []
    for line in open(path, encoding='utf-8'):
        line = line.rstrip('\n')
        if '\t' not in line: continue
        gold = line.split('\t')[1]
        toks = []
        for t in gold.split(' '):
            if '/' not in t: continue
            form, tag = t.rsplit('/', 1)
            if not form or not tag: continue
            toks.append((form, tag))
        if toks: sents.append(toks)
    return sents

def main():
    args = sys.argv[1:]
    out = args.pop(0)
    synth, seed = 0, 1
    files = []
    while args:
        a = args.pop(0)
        if a == '--synth': synth = int(args.pop(0))
        elif a == '--seed': seed = int(args.pop(0))
        else: files.append(a)
    sents = []
    for f in files: sents += parse_gold(f)
    rng = random.Random(seed)
    with open(out, 'w', encoding='utf-8') as fo:
        def emit(toks):
            fo.write('\t' + '\t'.join(f'{a}\t{b}' for a, b in toks) + '\n\n')
        for s in sents: emit(s)
        # synthetic sentences: splice random gold sentence halves -> new n-grams across the seam, same unigrams
        for _ in range(synth):
            a, b = rng.choice(sents), rng.choice(sents)
            i, j = rng.randrange(len(a) + 1), rng.randrange(len(b) + 1)
            s = a[:i] + b[j:]
            if len(s) >= 2: emit(s)

if __name__ == '__main__':
    main()
