// Flatten a reference-built kiwi::Kiwi (any model directory the reference can load) into the
// position-independent model image described in include/kiwi_b200_image.h.
// TEST/FIXTURE INFRASTRUCTURE: links oracle/_ref/libkiwi_ref.so and is compiled with -fno-access-control
// to read the private arrays of kiwi::Kiwi (include/kiwi/Kiwi.h:176-208), utils::FrozenTrie
// (include/kiwi/FrozenTrie.h:94-99) and lm::KnLangModel (src/Knlm.hpp:28-36).  The tool is deliberately
// "dumb": it copies raw fields only; every derived quantity is computed by the product at load time.
// The Kiwi object is built with ArchType::balanced, whose nst::prepare keeps keys sorted ascending
// (src/search.cpp:238-292), which is the order the device kernels binary-search.
// With model type "cong" the trie/forms/morphemes still come from the balanced-arch Kiwi, and the language model
// from a second Kiwi built with ModelType::cong on ArchType::avx2 (the quantized CoNg model only exists for the
// SIMD archs, src/ArchAvailable.h:50-66); its arch-specific key packets are read back through nst::extractKV
// (src/search.h:94-110) and re-sorted ascending.
// usage: flatten_model <model_dir> <out.img> [model_name] [knlm|cong|sbg]
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>
#include <kiwi/Kiwi.h>
#include <kiwi/ScriptType.h>
#include <kiwi/Utils.h>
#include "Knlm.hpp"
#include "CoNgramModel.hpp"
#include "SkipBigramModel.hpp"
#include <algorithm>
#include "../../../include/kiwi_b200_image.h"

using namespace kiwi;

template<class T> static void putSection(std::vector<char>& blob, kb2_section& sec, const std::vector<T>& v)
{
	while (blob.size() % 256) blob.push_back(0);
	sec.offset = blob.size();
	sec.nbytes = v.size() * sizeof(T);
	const char* p = reinterpret_cast<const char*>(v.data());
	blob.insert(blob.end(), p, p + sec.nbytes);
}

template<class KeyType>
static bool dumpKnlm(const lm::ILangModel* base, kb2_header& h,
	std::vector<kb2_kn_node>& nodes, std::vector<uint32_t>& keys, std::vector<int32_t>& values,
	std::vector<int32_t>& root, std::vector<uint32_t>& htx)
{
	auto* m = dynamic_cast<const lm::KnLangModel<ArchType::balanced, KeyType>*>(base);
	if (!m) return false;
	const auto& hd = m->getHeader();
	const size_t nNodes = m->num_non_leaf_nodes;
	const size_t nEdges = hd.num_nodes - 1;
	const size_t htxVocab = m->value_data - m->all_value_data.get();
	nodes.resize(nNodes);
	for (size_t i = 0; i < nNodes; ++i)
	{
		const auto& n = m->node_data[i];
		nodes[i] = kb2_kn_node{ (uint32_t)n.num_nexts, (int32_t)n.lower, n.next_offset, n.ll, n.gamma };
	}
	keys.resize(nEdges);
	values.resize(nEdges);
	for (size_t i = 0; i < nEdges; ++i)
	{
		keys[i] = m->key_data[i];
		values[i] = m->value_data[i];
	}
	root.assign(m->all_value_data.get(), m->all_value_data.get() + htxVocab);
	if (m->htx_data) htx.assign(m->htx_data, m->htx_data + hd.vocab_size);
	h.kn_num_nodes = (uint32_t)nNodes;
	h.kn_num_edges = (uint32_t)nEdges;
	h.kn_htx_vocab = (uint32_t)htxVocab;
	h.kn_has_htx = m->htx_data ? 1 : 0;
	h.kn_order = hd.order;
	h.kn_bos_node = (int32_t)m->bos_node_idx;
	h.kn_unk_ll = m->unk_ll;
	h.lang_vocab_size = (uint32_t)hd.vocab_size;
	return true;
}

struct CongDump
{
	std::vector<kb2_cg_node> nodes; std::vector<uint32_t> keys; std::vector<int32_t> values, root;
	std::vector<uint8_t> ctxEmb, outEmb; std::vector<uint32_t> invVocab; std::vector<float> outBias;
};

template<class KeyType, class VlKeyType>
static bool dumpCong(const lm::ILangModel* base, kb2_header& h, CongDump& d)
{
	using Model = lm::CoNgramModel<ArchType::avx2, KeyType, VlKeyType, 0, true>;
	auto* m = dynamic_cast<const Model*>(base);
	if (!m) return false;
	const auto& hd = m->getHeader();
	// non-leaf nodes are contiguous in DFS order: walk them while the highest reachable index grows
	size_t count = 1;
	for (size_t i = 0; i < count; ++i)
	{
		const auto& n = m->nodeData[i];
		std::vector<std::pair<uint32_t, int32_t>> kv(n.numNexts);
		for (size_t j = 0; j < n.numNexts; ++j)
		{
			auto p = nst::extractKV<ArchType::avx2, VlKeyType, int32_t>(&m->alignedKeyValueData[n.nextOffset], n.numNexts, j);
			kv[j] = std::make_pair((uint32_t)p.first, p.second);
			if (p.second > 0) count = std::max(count, i + (size_t)p.second + 1);
		}
		std::sort(kv.begin(), kv.end());
		d.nodes.push_back(kb2_cg_node{ (int32_t)n.lower, n.value, (uint32_t)d.keys.size(), (uint32_t)n.numNexts });
		for (auto& p : kv) { d.keys.push_back(p.first); d.values.push_back(p.second); }
	}
	d.root.assign(m->allRootValueData.get(), m->allRootValueData.get() + hd.vocabSize);
	const size_t cs = m->contextEmbStride(), os = m->outputEmbStride();
	if (cs != (size_t)hd.dim + 8 || os != (size_t)hd.dim + 8) throw std::runtime_error{ "unexpected CoNg row stride" };
	d.ctxEmb.assign(m->contextEmbPtr, m->contextEmbPtr + hd.contextSize * cs);
	d.outEmb.assign(m->outputEmbPtr, m->outputEmbPtr + hd.vocabSize * os);
	if (m->invertedContextVocabPtr) for (size_t i = 0; i < hd.vocabSize; ++i) d.invVocab.push_back((uint32_t)m->invertedContextVocabPtr[i]);
	if (m->outputEmbBiasPtr) d.outBias.assign(m->outputEmbBiasPtr, m->outputEmbBiasPtr + hd.vocabSize);
	h.cg_num_nodes = (uint32_t)d.nodes.size();
	h.cg_num_edges = (uint32_t)d.keys.size();
	h.cg_root_size = (uint32_t)hd.vocabSize;
	h.cg_dim = hd.dim;
	h.cg_context_size = (uint32_t)hd.contextSize;
	h.cg_key_size = hd.keySize;
	h.cg_flags = hd.flags;
	h.lang_vocab_size = (uint32_t)hd.vocabSize;
	return true;
}

struct SbgDump { std::vector<uint32_t> ptrs, keys; std::vector<float> comps, discnts; std::vector<uint8_t> valid; };

template<class KeyType>
static const lm::ILangModel* dumpSbg(const lm::ILangModel* base, kb2_header& h, SbgDump& d)
{
	using Model = lm::SkipBigramModel<ArchType::balanced, KeyType, 8>;
	auto* m = dynamic_cast<const Model*>(base);
	if (!m) return nullptr;
	const auto& hd = m->getHeader();
	const size_t total = m->ptrs[hd.vocabSize];
	for (size_t i = 0; i <= hd.vocabSize; ++i) d.ptrs.push_back((uint32_t)m->ptrs[i]);
	for (size_t i = 0; i < total; ++i) { d.keys.push_back((uint32_t)m->keyData[i]); d.comps.push_back(m->compensations[i]); }
	for (size_t i = 0; i < hd.vocabSize; ++i)
	{
		d.discnts.push_back(m->discnts[i]); d.valid.push_back(m->vocabValidness[i]);
		for (size_t j = m->ptrs[i] + 1; j < m->ptrs[i + 1]; ++j) if (!(m->keyData[j - 1] < m->keyData[j])) throw std::runtime_error{ "sbg keys are not ascending" };
	}
	h.sb_vocab_size = (uint32_t)hd.vocabSize; h.sb_window_size = hd.windowSize; h.sb_num_pairs = (uint32_t)total;
	return &m->knlm;
}

int main(int argc, char** argv)
{
	if (argc < 3) { std::cerr << "usage: flatten_model <model_dir> <out.img> [name]\n"; return 2; }
	setenv("KIWI_ARCH_TYPE", "balanced", 1);
	try
	{
		const bool wantSbg = argc > 4 && std::string{ argv[4] } == "sbg";
		KiwiBuilder kb{ argv[1], 1, BuildOption::default_, wantSbg ? ModelType::sbg : ModelType::knlm };
		Kiwi kw = kb.build();

		kb2_header h;
		std::memset(&h, 0, sizeof(h));
		h.magic = KB2_IMAGE_MAGIC;
		h.version = KB2_IMAGE_VERSION;
		const bool cong = argc > 4 && std::string{ argv[4] } == "cong";
		const bool sbgModel = argc > 4 && std::string{ argv[4] } == "sbg";
		h.model_type = (uint32_t)(cong ? ModelType::cong : sbgModel ? ModelType::sbg : ModelType::knlm);
		std::strncpy(h.model_name, argc > 3 ? argv[3] : argv[1], sizeof(h.model_name) - 1);

		// ---- form trie
		const auto& ft = kw.formTrie;
		std::vector<kb2_trie_node> tnodes(ft.numNodes);
		std::vector<uint16_t> tkeys(ft.nextKeys.get(), ft.nextKeys.get() + ft.numNexts);
		std::vector<int32_t> tdiffs(ft.nextDiffs.get(), ft.nextDiffs.get() + ft.numNexts);
		const Form* formBase = kw.forms.data();
		for (size_t i = 0; i < ft.numNodes; ++i)
		{
			const auto& n = ft.nodes[i];
			const Form* v = ft.values[i];
			int32_t value;
			if (!v) value = KB2_TRIE_NONE;
			else if (ft.hasSubmatch(v)) value = KB2_TRIE_SUBMATCH;
			else value = (int32_t)(v - formBase);
			tnodes[i] = kb2_trie_node{ n.nextOffset, n.lower, value, (uint16_t)n.numNexts, n.depth };
			for (size_t j = 1; j < n.numNexts; ++j)
			{
				if (!(tkeys[n.nextOffset + j - 1] < tkeys[n.nextOffset + j])) throw std::runtime_error{ "trie keys are not ascending" };
			}
		}

		// ---- forms
		std::vector<kb2_form> forms(kw.forms.size());
		std::vector<uint16_t> fchars;
		std::vector<uint32_t> fcands;
		const Morpheme* morphBase = kw.morphemes.data();
		for (size_t i = 0; i < kw.forms.size(); ++i)
		{
			const auto& f = kw.forms[i];
			kb2_form o;
			std::memset(&o, 0, sizeof(o));
			o.str_off = (uint32_t)fchars.size();
			o.str_len = (uint16_t)f.form.size();
			fchars.insert(fchars.end(), f.form.begin(), f.form.end());
			o.cand_off = (uint32_t)fcands.size();
			o.cand_cnt = (uint16_t)f.candidate.size();
			for (auto* c : f.candidate) fcands.push_back((uint32_t)(c - morphBase));
			o.num_spaces = (uint16_t)f.numSpaces;
			o.dialect = (uint16_t)f.dialect;
			o.vowel = (uint8_t)f.vowel;
			o.polar = (uint8_t)f.polar;
			o.form_hash = f.formHash;
			o.flags = (f.zCodaAppendable ? KB2_FORM_ZCODA : 0) | (f.zSiotAppendable ? KB2_FORM_ZSIOT : 0)
				| (f.hasJClass ? KB2_FORM_HASJ : 0) | (f.hasAnyFullMorphemes ? KB2_FORM_HASFULL : 0);
			if (f.form.size() > 0xFFFF || f.candidate.size() > 0xFFFF) throw std::runtime_error{ "form too large" };
			forms[i] = o;
		}

		// ---- morphemes
		std::vector<kb2_morph> morphs(kw.morphemes.size());
		std::vector<kb2_chunk> chunks;
		for (size_t i = 0; i < kw.morphemes.size(); ++i)
		{
			const auto& m = kw.morphemes[i];
			kb2_morph o;
			std::memset(&o, 0, sizeof(o));
			if (m.kform)
			{
				const Form* f = reinterpret_cast<const Form*>(reinterpret_cast<const char*>(m.kform) - offsetof(Form, form));
				if (f < formBase || f >= formBase + kw.forms.size()) throw std::runtime_error{ "kform outside forms[]" };
				o.form_idx = (int32_t)(f - formBase);
			}
			else o.form_idx = -1;
			o.combined = m.combined;
			o.chunk_off = (uint32_t)chunks.size();
			o.chunk_cnt = (uint8_t)m.chunks.size();
			if (m.chunks.size() > 255) throw std::runtime_error{ "too many chunks" };
			for (size_t c = 0; c < m.chunks.size(); ++c)
			{
				const auto& p = m.chunks.getSecond(c);
				chunks.push_back(kb2_chunk{ (uint32_t)(m.chunks[c] - morphBase), p.first, p.second, 0 });
			}
			o.lm_morpheme_id = m.lmMorphemeId;
			o.orig_morpheme_id = m.origMorphemeId;
			o.user_score = m.userScore;
			o.dialect = (uint16_t)m.dialect;
			o.tag = (uint8_t)m.tag;
			o.vowel = (uint8_t)m.vowel;
			o.polar = (uint8_t)m.polar;
			o.flags = (m.complex ? KB2_MORPH_COMPLEX : 0) | (m.saisiot ? KB2_MORPH_SAISIOT : 0);
			o.sense_id = m.senseId;
			o.combine_socket = m.combineSocket;
			morphs[i] = o;
		}

		// ---- Knlm
		std::vector<kb2_kn_node> knodes; std::vector<uint32_t> kkeys, khtx; std::vector<int32_t> kvals, kroot;
		const auto* lmBase = kw.langMdl.get();
		SbgDump sb;
		if (sbgModel)
		{
			const lm::ILangModel* inner = dumpSbg<uint32_t>(lmBase, h, sb);
			if (!inner) inner = dumpSbg<uint16_t>(lmBase, h, sb);
			if (!inner) throw std::runtime_error{ "language model is not a balanced-arch SkipBigramModel (window 8)" };
			lmBase = inner;
		}
		CongDump cg;
		std::unique_ptr<Kiwi> kwCong;
		if (cong)
		{
			setenv("KIWI_ARCH_TYPE", "avx2", 1);
			KiwiBuilder kbc{ argv[1], 1, BuildOption::default_, ModelType::cong };
			kwCong = std::make_unique<Kiwi>(kbc.build());
			setenv("KIWI_ARCH_TYPE", "balanced", 1);
			if (kwCong->morphemes.size() != kw.morphemes.size() || kwCong->forms.size() != kw.forms.size()) throw std::runtime_error{ "cong / knlm builds differ" };
			const auto* cb = kwCong->langMdl.get();
			if (!dumpCong<uint16_t, uint16_t>(cb, h, cg) && !dumpCong<uint32_t, uint16_t>(cb, h, cg) && !dumpCong<uint32_t, uint32_t>(cb, h, cg))
				throw std::runtime_error{ "language model is not an avx2 quantized CoNgramModel without window" };
		}
		else
		if (!dumpKnlm<uint16_t>(lmBase, h, knodes, kkeys, kvals, kroot, khtx)
			&& !dumpKnlm<uint32_t>(lmBase, h, knodes, kkeys, kvals, kroot, khtx)
			&& !dumpKnlm<uint8_t>(lmBase, h, knodes, kkeys, kvals, kroot, khtx)
			&& !dumpKnlm<uint64_t>(lmBase, h, knodes, kkeys, kvals, kroot, khtx))
		{
			throw std::runtime_error{ "language model is not a balanced-arch KnLangModel" };
		}

		// ---- scalars
		h.n_trie_nodes = (uint32_t)tnodes.size();
		h.n_trie_edges = (uint32_t)tkeys.size();
		h.n_forms = (uint32_t)forms.size();
		h.n_morphs = (uint32_t)morphs.size();
		h.n_chunks = (uint32_t)chunks.size();
		h.default_tag_size = (uint32_t)defaultTagSize;
		h.postag_max = (uint32_t)POSTag::max;
		if ((size_t)POSTag::max > KB2_POSTAG_MAX) throw std::runtime_error{ "POSTag::max too large" };
		for (size_t i = 0; i < 6; ++i) h.special_morph_ids[i] = (uint32_t)kw.specialMorphIds[i];
		for (int b = 0; b < 2; ++b) for (size_t t = 0; t < (size_t)POSTag::max; ++t)
		{
			h.tag_left_boundary[b][t] = kw.tagScorer.evalLeftBoundary(b != 0, (POSTag)t);
		}
		const auto& gc = kw.globalConfig;
		h.config = kb2_config{ gc.cutOffThreshold, gc.oovRuleScale, gc.oovRuleBias, gc.spacePenalty, gc.typoCostWeight,
			gc.maxUnkFormSize, gc.maxUnkFormSizeFollowedByJClass, gc.spaceTolerance, gc.integrateAllomorph ? 1u : 0u };

		// ---- code-point attribute runs
		std::vector<kb2_chr_run> runs;
		for (uint32_t c = 0; c <= 0x10FFFF; ++c)
		{
			kb2_chr_run r{ c, (uint8_t)identifySpecialChr((char32_t)c), (uint8_t)chr2ScriptType((char32_t)c), 0, 0 };
			if (c < 0x10000 && isSpace((char16_t)c)) r.flags |= KB2_CHR_SPACE;
			if (isEmoji((char32_t)c, 0) == 1) r.flags |= KB2_CHR_EMOJI1;
			else if (isEmoji((char32_t)c, 0xfe0f) == 2) r.flags |= KB2_CHR_EMOJI2;
			if (runs.empty() || runs.back().cls != r.cls || runs.back().script != r.script || runs.back().flags != r.flags) runs.push_back(r);
		}
		h.n_chr_runs = (uint32_t)runs.size();
		h.script_latin = (uint32_t)ScriptType::latin;
		h.script_variation_selectors = (uint32_t)ScriptType::variation_selectors;

		std::vector<char> blob(sizeof(kb2_header), 0);
		putSection(blob, h.sec[KB2_SEC_TRIE_NODES], tnodes);
		putSection(blob, h.sec[KB2_SEC_TRIE_KEYS], tkeys);
		putSection(blob, h.sec[KB2_SEC_TRIE_DIFFS], tdiffs);
		putSection(blob, h.sec[KB2_SEC_FORMS], forms);
		putSection(blob, h.sec[KB2_SEC_FORM_CHARS], fchars);
		putSection(blob, h.sec[KB2_SEC_FORM_CANDS], fcands);
		putSection(blob, h.sec[KB2_SEC_MORPHS], morphs);
		putSection(blob, h.sec[KB2_SEC_MORPH_CHUNKS], chunks);
		putSection(blob, h.sec[KB2_SEC_KN_NODES], knodes);
		putSection(blob, h.sec[KB2_SEC_KN_KEYS], kkeys);
		putSection(blob, h.sec[KB2_SEC_KN_VALUES], kvals);
		putSection(blob, h.sec[KB2_SEC_KN_ROOT], kroot);
		putSection(blob, h.sec[KB2_SEC_KN_HTX], khtx);
		putSection(blob, h.sec[KB2_SEC_CHR_RUNS], runs);
		putSection(blob, h.sec[KB2_SEC_CG_NODES], cg.nodes);
		putSection(blob, h.sec[KB2_SEC_CG_KEYS], cg.keys);
		putSection(blob, h.sec[KB2_SEC_CG_VALUES], cg.values);
		putSection(blob, h.sec[KB2_SEC_CG_ROOT], cg.root);
		putSection(blob, h.sec[KB2_SEC_CG_CTX_EMB], cg.ctxEmb);
		putSection(blob, h.sec[KB2_SEC_CG_OUT_EMB], cg.outEmb);
		putSection(blob, h.sec[KB2_SEC_CG_INV_VOCAB], cg.invVocab);
		putSection(blob, h.sec[KB2_SEC_CG_OUT_BIAS], cg.outBias);
		putSection(blob, h.sec[KB2_SEC_SB_PTRS], sb.ptrs);
		putSection(blob, h.sec[KB2_SEC_SB_KEYS], sb.keys);
		putSection(blob, h.sec[KB2_SEC_SB_COMPS], sb.comps);
		putSection(blob, h.sec[KB2_SEC_SB_DISCNTS], sb.discnts);
		putSection(blob, h.sec[KB2_SEC_SB_VALID], sb.valid);
		while (blob.size() % 256) blob.push_back(0);
		h.total_bytes = blob.size();
		std::memcpy(blob.data(), &h, sizeof(h));
		std::ofstream ofs{ argv[2], std::ios_base::binary };
		ofs.write(blob.data(), blob.size());
		std::cerr << "image: " << blob.size() << " bytes; trie nodes " << tnodes.size() << " edges " << tkeys.size()
			<< "; forms " << forms.size() << "; morphemes " << morphs.size() << "; knlm nodes " << knodes.size()
			<< " edges " << kkeys.size() << " vocab " << h.lang_vocab_size << " htxVocab " << h.kn_htx_vocab
			<< " bos " << h.kn_bos_node << " unk_ll " << h.kn_unk_ll
			<< "; cong nodes " << h.cg_num_nodes << " edges " << h.cg_num_edges << " dim " << h.cg_dim << " contexts " << h.cg_context_size
			<< " keySize " << h.cg_key_size << " flags " << h.cg_flags << "; sbg pairs " << h.sb_num_pairs << std::endl;
	}
	catch (const std::exception& e)
	{
		std::cerr << "flatten_model failed: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
