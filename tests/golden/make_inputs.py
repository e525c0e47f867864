#!/usr/bin/env python3
"""Builds tests/golden/inputs_ref_tests.txt: one input sentence per line, taken from
 (a) every UTF-16 string literal (u"...") that the reference's own tests feed to the analyzer
     (/root/reference/test/test_cpp.cpp, test_c.cpp, test_typo.cpp), and
 (b) hand-written edge cases for the lattice builder (empty / blank lines, pattern matches, emoji
     sequences, surrogates, old Hangul, very long unbroken runs, chunk-boundary shapes).
Run here (needs /root/reference); the output is committed so the GPU box never reads the reference."""
import re, sys, os
REF = '/root/reference/test'
out = []
seen = set()
def add(s):
    s = s.replace('\r', ' ').replace('\n', ' ').replace('\t', ' ')   # one sentence per line, tab = field separator
    if s in seen: return
    seen.add(s); out.append(s)

lit = re.compile(r'u"((?:[^"\\]|\\.)*)"')
def unescape(s):
    def rep(m):
        e = m.group(1)
        if e[0] == 'u': return chr(int(e[1:5], 16))
        if e[0] == 'U': return chr(int(e[1:9], 16))
        if e[0] == 'x': return chr(int(e[1:], 16))
        return {'n': '\n', 't': '\t', 'r': '\r', '"': '"', "'": "'", '\\': '\\', '0': '\0'}.get(e, e)
    return re.sub(r'\\(u[0-9a-fA-F]{4}|U[0-9a-fA-F]{8}|x[0-9a-fA-F]{1,4}|.)', rep, s)

for fn in ['test_cpp.cpp', 'test_c.cpp', 'test_typo.cpp']:
    src = open(os.path.join(REF, fn), encoding='utf-8-sig').read()
    for m in lit.finditer(src):
        s = unescape(m.group(1))
        if '\0' in s or not s: continue
        add(s)

edge = [
    '', ' ', '   ', '.', '?!', '...', 'a', '1', '가', '각', 'ㄱ', 'ᆨ', '​', '안녕', '안녕하세요.', '안녕하세요. 반갑습니다.',
    '안녕하세요.  반갑습니다!   정말요?  네.', '가.나.다.라 마바사', 'ab. cd. ef. gh. ij. kl.', '1. 2. 3. 4. 5.',
    'https://github.com/bab2min/Kiwi 에서 받으세요', 'http://a.bc', 'http://a.bc:8080/x?y=1. 끝', 'mail@example.com 으로 보내주세요', '@kiwi_bot 안녕 #해시태그 #tag2.',
    '전화번호는 010-1234-5678 입니다', '2024.01.02. 3:45 경', '1,234,567.89원', '3.14', '1.', '12,34', '1/2/3', 'U.S.A. 에서', 'Mr. Kim', 'e.g. this',
    '😀', '😀😀😀', '👨‍👩‍👧‍👦 가족', '👍🏽 좋아요', '☺️ 웃음', '©️ ® ™', '🇰🇷 대한민국', 'ㅋㅋㅋㅋㅋ 웃기다 ㅠㅠ', 'ㅇㅇ ㄴㄴ', '됬다 됐다', '했읍니다', '하겠슴다',
    '𠀀𠀁 한자 漢字 かな カナ Ελληνικά кириллица', 'ＡＢＣ１２３', 'αβγ δ', '\ud83d', 'abc\ud83dabc', '\udc00x',
    'ᄀᆞᆷ ᄒᆞᆫ글 옛한글', 'ᄒᆞᆫ', '〮〯',
    '"따옴표" \'작은따옴표\' (괄호) [대괄호] {중괄호} <꺽쇠> 《겹》 「낫표」', '"열린 따옴표', '\'\'', '가. 나) (다) ① ② 1) 2)',
    '가'*50, '가나다라마바사아자차카타파하'*20, 'a'*300, '가 '*100, '1'*200, '.'*100, '가.'*60,
    'ㄴ다', '잇다', '갔ㄴ데', '먹었엌ㅋㅋ', '웃기닼ㅋㅋ', '안녕하세욯ㅎ', '바다ㅅ가', '나뭇잎', '했다고 하ㅂ니다',
    '아버지가방에들어가신다', '나는 학교에 간다', '그는 책을 읽었다.', '이것은 사과이고 저것은 배다', '먹고 싶다', '할 수 있다', '봤어요', '해서 좋았음',
    'ㅏㅑㅓㅕ', '  앞 공백', '뒤 공백  ', ' 가 나 ', '가 나　다', '가‍나', '가-나~다', '가,나;다:라/마',
    'Kiwi는 C++로 작성된 형태소 분석기입니다.', 'GPU(B200)에서 8192문장/초?', 'x=y+1', '100% 50$ #1 @2',
]
for s in edge: add(s)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'inputs_ref_tests.txt'), 'w', encoding='utf-8', errors='surrogatepass') as f:
    for s in out: f.write(s + '\n')
print(len(out), 'inputs')

# ---- round 2: tests/golden/inputs_web_typos.txt = first column of eval_data/web_with_typos.txt (the typo'd sentences BASELINE.json
# config 4 draws 30 % of its synthetic batch from)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'inputs_web_typos.txt'), 'w', encoding='utf-8') as f:
    for line in open('/root/reference/eval_data/web_with_typos.txt', encoding='utf-8'):
        if line.strip(): f.write(line.split('\t')[0].rstrip('\n') + '\n')
