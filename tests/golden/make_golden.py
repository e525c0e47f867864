#!/usr/bin/env python3
"""Regenerates the golden vectors from the UNMODIFIED reference (oracle/_ref/dump_golden, built by
oracle/ref_build/Makefile from /root/reference) on the committed input files.  Run here (needs oracle/_ref);
the .golden.txt.gz outputs are committed so that neither the CPU suite nor the GPU box needs the reference.
Format: see oracle/ref_build/tools/dump_golden.cpp.  MANIFEST.json records the md5 of the model image the
vectors belong to (the fabricated Knlm model is rebuilt deterministically by oracle/Makefile)."""
import gzip, hashlib, json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
IMAGE = os.path.join(REFDIR, "models", "knlm_small.img")
CONG_IMAGE = os.path.join(REFDIR, "models", "cong_small.img")
SBG_IMAGE = os.path.join(REFDIR, "models", "sbg_small.img")
manifest = {"image_md5": hashlib.md5(open(IMAGE, "rb").read()).hexdigest(), "model": "knlm_small (fabricated, see oracle/Makefile)",
            "cong_image_md5": hashlib.md5(open(CONG_IMAGE, "rb").read()).hexdigest(), "cong_model": "cong_small (fabricated, see oracle/Makefile)",
            "sbg_image_md5": hashlib.md5(open(SBG_IMAGE, "rb").read()).hexdigest(), "sbg_model": "sbg_small (fabricated, see oracle/Makefile)",
            "reference_arch": "avx2", "files": {}}
# Knlm vectors: <name>.golden.txt.gz; CoNg / SkipBigram vectors (same inputs, ModelType::cong / ::sbg): cong_<name> / sbg_<name>.golden.txt.gz
for model, mtype, prefix in [("knlm_small", "knlm", ""), ("cong_small", "cong", "cong_"), ("sbg_small", "sbg", "sbg_")]:
    for name in ["inputs_ref_tests", "inputs_web", "inputs_written", "inputs_dialect_typos"]:
        if mtype == "sbg" and name == "inputs_dialect_typos": continue      # the reference needs minutes per sentence on some of them (history states do not merge)
        src = os.path.join(HERE, name + ".txt")
        tmp = os.path.join("/tmp", prefix + name + ".golden.txt")
        env = dict(os.environ, KIWI_ARCH_TYPE="avx2", KB_MODEL_TYPE=mtype)
        # SkipBigram states rarely merge, so the thread-local `top1` set of the reference is in use all the time and its history shows in
        # the results: these vectors are made with one fresh thread per sentence (dump_golden.cpp, KB_FRESH_THREAD)
        if mtype == "sbg": env["KB_FRESH_THREAD"] = "1"
        # SkipBigram states carry an 8-token history and rarely merge: on the pathological repeated-syllable inputs at the end
        # of inputs_ref_tests the REFERENCE itself needs tens of GB, so the sbg vectors stop before them
        limit = ["449"] if (mtype == "sbg" and name == "inputs_ref_tests") else []
        subprocess.run([os.path.join(REFDIR, "dump_golden"), os.path.join(REFDIR, "models", model), src, tmp] + limit, check=True, env=env, timeout=300)
        data = open(tmp, "rb").read()
        with gzip.GzipFile(os.path.join(HERE, prefix + name + ".golden.txt.gz"), "wb", mtime=0) as f:
            f.write(data)
        manifest["files"][prefix + name] = {"lines": data.count(b"\nS ") + (1 if data.startswith(b"S ") else 0), "md5": hashlib.md5(data).hexdigest()}
# AnalyzeOption::openEnding (no end-of-sentence step on the chunk that ends the text) with the Knlm model: open_<name>.golden.txt.gz
for name in ["inputs_written", "inputs_web"]:
    tmp = os.path.join("/tmp", "open_" + name + ".golden.txt")
    env = dict(os.environ, KIWI_ARCH_TYPE="avx2", KB_MODEL_TYPE="knlm", KB_OPEN_ENDING="1")
    subprocess.run([os.path.join(REFDIR, "dump_golden"), os.path.join(REFDIR, "models", "knlm_small"), os.path.join(HERE, name + ".txt"), tmp], check=True, env=env, timeout=600)
    data = open(tmp, "rb").read()
    with gzip.GzipFile(os.path.join(HERE, "open_" + name + ".golden.txt.gz"), "wb", mtime=0) as f:
        f.write(data)
    manifest["files"]["open_" + name] = {"lines": data.count(b"\nS ") + (1 if data.startswith(b"S ") else 0), "md5": hashlib.md5(data).hexdigest()}
# AnalyzeOption::blocklist with the Knlm model: block_<name>.golden.txt.gz; the (form, tag) list and the morpheme ids the reference resolved
# it to are recorded in the manifest
BLOCK_SPEC = "하/VV;이/VCP;는/JX;을/JKO;것/NNB;있/VV;에서/JKB"
for name in ["inputs_written", "inputs_web"]:
    tmp = os.path.join("/tmp", "block_" + name + ".golden.txt")
    env = dict(os.environ, KIWI_ARCH_TYPE="avx2", KB_MODEL_TYPE="knlm", KB_BLOCKLIST=BLOCK_SPEC)
    out = subprocess.run([os.path.join(REFDIR, "dump_golden"), os.path.join(REFDIR, "models", "knlm_small"), os.path.join(HERE, name + ".txt"), tmp], check=True, env=env, timeout=600, capture_output=True, text=True).stdout
    manifest["blocklist"] = {"spec": BLOCK_SPEC, "morpheme_ids": [int(x) for x in [l for l in out.splitlines() if l.startswith("BLOCKLIST")][0].split()[1:]]}
    data = open(tmp, "rb").read()
    with gzip.GzipFile(os.path.join(HERE, "block_" + name + ".golden.txt.gz"), "wb", mtime=0) as f:
        f.write(data)
    manifest["files"]["block_" + name] = {"lines": data.count(b"\nS ") + (1 if data.startswith(b"S ") else 0), "md5": hashlib.md5(data).hexdigest()}
# typo-tolerant analysis (BASELINE.json config 4: AnalyzeOption::typoTransformer = basicTypoSet.prepare(true), typoThreshold 2.5,
# typoCostWeight 6) with the Knlm model: typo6_<name>.golden.txt.gz
for name in ["inputs_ref_tests", "inputs_web", "inputs_written", "inputs_dialect_typos"]:
    tmp = os.path.join("/tmp", "typo6_" + name + ".golden.txt")
    env = dict(os.environ, KIWI_ARCH_TYPE="avx2", KB_MODEL_TYPE="knlm", KB_TYPO="basic")
    subprocess.run([os.path.join(REFDIR, "dump_golden"), os.path.join(REFDIR, "models", "knlm_small"), os.path.join(HERE, name + ".txt"), tmp], check=True, env=env, timeout=600)
    data = open(tmp, "rb").read()
    with gzip.GzipFile(os.path.join(HERE, "typo6_" + name + ".golden.txt.gz"), "wb", mtime=0) as f:
        f.write(data)
    manifest["files"]["typo6_" + name] = {"lines": data.count(b"\nS ") + (1 if data.startswith(b"S ") else 0), "md5": hashlib.md5(data).hexdigest()}
# typo graphs (PreparedTypoTransformer::generateGraph of the unmodified reference, node for node): the rules and sentence of the
# reference's own test KiwiTypo.GenerateGraph (test/test_typo.cpp:8-22, 11 nodes) and the default basic typo set on real text
for tset, name in [("kat", "inputs_typo_kat"), ("basic", "inputs_typo_kat"), ("basic", "inputs_web"), ("basic", "inputs_written"), ("basic", "inputs_dialect_typos")]:
    tmp = "/tmp/typo_%s_%s.golden.txt" % (tset, name)
    subprocess.run([os.path.join(REFDIR, "typo_tool"), "graphs", tset, os.path.join(HERE, name + ".txt"), tmp], check=True, timeout=300)
    data = open(tmp, "rb").read()
    key = "typo_%s_%s" % (tset, name)
    with gzip.GzipFile(os.path.join(HERE, key + ".golden.txt.gz"), "wb", mtime=0) as f:
        f.write(data)
    manifest["files"][key] = {"lines": data.count(b"\nG ") + (1 if data.startswith(b"G ") else 0), "md5": hashlib.md5(data).hexdigest()}
for tset in ["kat", "basic"]:
    manifest["typo_%s_image_md5" % tset] = hashlib.md5(open(os.path.join(REFDIR, "models", "typo_%s.img" % tset), "rb").read()).hexdigest()
# known-answer vectors of the CoNg int8 scorer and context trie (oracle/ref_build/tools/cong_probe.cpp)
subprocess.run([os.path.join(REFDIR, "cong_probe"), os.path.join(REFDIR, "models", "cong_small"), "/tmp/cong_qgemm.golden.txt"], check=True)
data = open("/tmp/cong_qgemm.golden.txt", "rb").read()
with gzip.GzipFile(os.path.join(HERE, "cong_qgemm.golden.txt.gz"), "wb", mtime=0) as f:
    f.write(data)
manifest["files"]["cong_qgemm"] = {"lines": data.count(b"\n"), "md5": hashlib.md5(data).hexdigest()}
json.dump(manifest, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1)
print(manifest)
