#!/usr/bin/env python3
"""Generate oracle/_ref/gen/KiwiBuilder_nosbg.cpp from the reference's src/KiwiBuilder.cpp.

TEST INFRASTRUCTURE.  The reference's SkipBigram *trainer* (src/SkipBigramTrainer.hpp, pulled in by
src/KiwiBuilder.cpp:18) needs full Eigen, which is an empty submodule in the snapshot.  Training SBG is not
on the analysed hot path, so the generated copy drops (a) that include, (b) the SBDataFeeder helper
(KiwiBuilder.cpp:1252-1284) and (c) replaces the body of the SBG-training constructor
`KiwiBuilder::KiwiBuilder(const string& modelPath, const ModelBuildArgs& args)` (KiwiBuilder.cpp:1286-1477)
with a throw.  Nothing else is touched; the output lives only under the git-ignored oracle/_ref/.
"""
import re, sys
src, dst = sys.argv[1], sys.argv[2]
text = open(src, encoding='utf-8-sig').read()
text = text.replace('#include "SkipBigramTrainer.hpp"\n', '', 1)

def cut_braced(text, start_idx):
    """return index just past the brace block that opens at/after start_idx"""
    i = text.index('{', start_idx)
    depth = 0
    while True:
        ch = text[i]
        if ch == '{': depth += 1
        elif ch == '}':
            depth -= 1
            if depth == 0: return i + 1
        i += 1

# (b) namespace kiwi { template<class Vid> class SBDataFeeder ... }
m = re.search(r'namespace kiwi\s*\{\s*template<class Vid>\s*class SBDataFeeder', text)
assert m, 'SBDataFeeder block not found'
end = cut_braced(text, m.start())
text = text[:m.start()] + text[end:]

# (c) the SBG-training constructor
sig = 'KiwiBuilder::KiwiBuilder(const string& modelPath, const ModelBuildArgs& args)'
s = text.index(sig)
body_open = text.index('\n{', s) + 1          # skip the `: KiwiBuilder{ modelPath }` delegating initialiser
end = cut_braced(text, body_open)
text = text[:body_open] + '{\n\tthrow std::runtime_error{ "SkipBigram training is not available in the oracle build (Eigen absent)" };\n}' + text[end:]
open(dst, 'w', encoding='utf-8').write(text)
