// ORACLE (test infrastructure, never linked into the product): read-only view of a kiwi_b200 model image
// (include/kiwi_b200_image.h) plus the per-code-point attribute lookup rebuilt from its run table.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/kiwi_b200_image.h"

namespace orc
{
	// work counters of the byte model; filled when a pointer is installed
	struct WorkCounters
	{
		uint64_t sentences = 0, rawUnits = 0, normUnits = 0;
		uint64_t trieVisits = 0, trieProbes = 0, trieHits = 0, candForms = 0;      // V_t, sum ceil(log2(k+1)), H_t, C
		uint64_t nodesBuilt = 0, nodesFinal = 0;                                    // L, L'
		uint64_t candEntries = 0, candEvals = 0;                                    // sum cand (ids read), sum C_v (morpheme records)
		uint64_t lmSteps = 0, lmHops = 0, lmProbes = 0;                             // progress() calls, H_lm, sum ceil(log2(k+1))
		uint64_t pairs = 0, pathsWritten = 0, pathsKept = 0, tokens = 0;           // P_read, P (container inserts), surviving, T
		uint64_t cgRows = 0, cgMacs = 0;                                            // CoNg: unique rows gathered per node, sum m*n*dim
	};
	inline uint32_t ceilLog2p1(uint32_t k) { uint32_t r = 0; while ((1u << r) < k + 1) ++r; return r; }
	// reference POSTag values used by name (include/kiwi/Types.h:195-227)
	enum Tag : uint8_t
	{
		T_unknown = 0, T_nng, T_nnp, T_nnb, T_vv, T_va, T_mag, T_nr, T_np, T_vx, T_mm, T_maj, T_ic,
		T_xpn, T_xsn, T_xsv, T_xsa, T_xsm, T_xr, T_vcp, T_vcn,
		T_sf, T_sp, T_ss, T_sso, T_ssc, T_se, T_so, T_sw, T_sb, T_sl, T_sh, T_sn,
		T_w_url, T_w_email, T_w_mention, T_w_hashtag, T_w_serial, T_w_emoji,
		T_jks, T_jkc, T_jkg, T_jko, T_jkb, T_jkv, T_jkq, T_jx, T_jc,
		T_ep, T_ef, T_ec, T_etn, T_etm, T_z_coda, T_z_siot,
		T_user0, T_user1, T_user2, T_user3, T_user4, T_p, T_max,
		T_irregular = 0x80,
	};
	inline uint8_t clearIrregular(uint8_t t) { return t & 0x7F; }
	inline bool isIrregular(uint8_t t) { return (t & 0x80) != 0; }
	inline bool isEClass(uint8_t t) { return T_ep <= t && t <= T_etm; }               // include/kiwi/TagUtils.h
	inline bool isJClass(uint8_t t) { return T_jks <= t && t <= T_jc; }
	inline bool isNNClass(uint8_t t) { return T_nng <= t && t <= T_nnb; }
	inline bool isVerbClass(uint8_t t)                                               // src/TagUtils.cpp:21-25
	{
		t = clearIrregular(t);
		return t == T_vv || t == T_va || t == T_vx || t == T_xsv || t == T_xsa || t == T_vcp || t == T_vcn;
	}

	// CondVowel / CondPolarity (include/kiwi/Types.h:243-270)
	enum { CV_none = 0, CV_any, CV_vowel, CV_vocalic, CV_vocalic_h, CV_non_vowel, CV_non_vocalic, CV_non_vocalic_h, CV_applosive };
	enum { CP_none = 0, CP_positive, CP_negative, CP_non_adj };

	struct Image
	{
		std::vector<char> blob;
		const kb2_header* h = nullptr;
		const kb2_trie_node* trieNodes = nullptr;
		const uint16_t* trieKeys = nullptr;
		const int32_t* trieDiffs = nullptr;
		const kb2_form* forms = nullptr;
		const uint16_t* formChars = nullptr;
		const uint32_t* formCands = nullptr;
		const kb2_morph* morphs = nullptr;
		const kb2_chunk* chunks = nullptr;
		const kb2_kn_node* knNodes = nullptr;
		const uint32_t* knKeys = nullptr;
		const int32_t* knValues = nullptr;
		const int32_t* knRoot = nullptr;
		const uint32_t* knHtx = nullptr;
		const kb2_chr_run* runs = nullptr;
		std::vector<uint8_t> bmpCls, bmpScript, bmpFlags;

		template<class T> const T* sec(int id) const { return reinterpret_cast<const T*>(blob.data() + h->sec[id].offset); }

		void load(const std::string& path)
		{
			FILE* f = std::fopen(path.c_str(), "rb");
			if (!f) throw std::runtime_error("cannot open image " + path);
			std::fseek(f, 0, SEEK_END);
			const long n = std::ftell(f);
			std::fseek(f, 0, SEEK_SET);
			blob.resize((size_t)n);
			if (std::fread(blob.data(), 1, (size_t)n, f) != (size_t)n) { std::fclose(f); throw std::runtime_error("short read"); }
			std::fclose(f);
			h = reinterpret_cast<const kb2_header*>(blob.data());
			if (h->magic != KB2_IMAGE_MAGIC || h->version != KB2_IMAGE_VERSION) throw std::runtime_error("bad image magic/version");
			trieNodes = sec<kb2_trie_node>(KB2_SEC_TRIE_NODES);
			trieKeys = sec<uint16_t>(KB2_SEC_TRIE_KEYS);
			trieDiffs = sec<int32_t>(KB2_SEC_TRIE_DIFFS);
			forms = sec<kb2_form>(KB2_SEC_FORMS);
			formChars = sec<uint16_t>(KB2_SEC_FORM_CHARS);
			formCands = sec<uint32_t>(KB2_SEC_FORM_CANDS);
			morphs = sec<kb2_morph>(KB2_SEC_MORPHS);
			chunks = sec<kb2_chunk>(KB2_SEC_MORPH_CHUNKS);
			knNodes = sec<kb2_kn_node>(KB2_SEC_KN_NODES);
			knKeys = sec<uint32_t>(KB2_SEC_KN_KEYS);
			knValues = sec<int32_t>(KB2_SEC_KN_VALUES);
			knRoot = sec<int32_t>(KB2_SEC_KN_ROOT);
			knHtx = h->kn_has_htx ? sec<uint32_t>(KB2_SEC_KN_HTX) : nullptr;
			runs = sec<kb2_chr_run>(KB2_SEC_CHR_RUNS);
			bmpCls.resize(0x10000); bmpScript.resize(0x10000); bmpFlags.resize(0x10000);
			for (uint32_t i = 0; i < h->n_chr_runs; ++i)
			{
				const uint32_t s = runs[i].start, e = i + 1 < h->n_chr_runs ? runs[i + 1].start : 0x110000;
				for (uint32_t c = s; c < e && c < 0x10000; ++c) { bmpCls[c] = runs[i].cls; bmpScript[c] = runs[i].script; bmpFlags[c] = runs[i].flags; }
			}
		}

		const kb2_chr_run& run(uint32_t c) const
		{
			size_t lo = 0, hi = h->n_chr_runs;            // last run with start <= c
			while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (runs[mid].start <= c) lo = mid; else hi = mid; }
			return runs[lo];
		}
		uint8_t cls(uint32_t c) const { return c < 0x10000 ? bmpCls[c] : run(c).cls; }          // identifySpecialChr
		uint8_t script(uint32_t c) const { return c < 0x10000 ? bmpScript[c] : run(c).script; } // chr2ScriptType
		uint8_t flags(uint32_t c) const { return c < 0x10000 ? bmpFlags[c] : (c <= 0x10FFFF ? run(c).flags : 0); }
		bool isSpace(uint16_t c) const { return (bmpFlags[c] & KB2_CHR_SPACE) != 0; }
		int isEmoji(uint32_t c0, uint32_t c1) const                                              // src/ScriptType.cpp:569-753
		{
			const uint8_t f = flags(c0);
			if (f & KB2_CHR_EMOJI1) return 1;
			if (!(c1 == 0xfe0f || (0x1f3fb <= c1 && c1 <= 0x1f3ff))) return 0;
			return (f & KB2_CHR_EMOJI2) ? 2 : 0;
		}

		const uint16_t* formStr(int32_t formIdx) const { return formChars + forms[formIdx].str_off; }
		uint32_t formLen(int32_t formIdx) const { return forms[formIdx].str_len; }
	};

	inline bool isHangulSyllable(uint32_t c) { return 0xAC00 <= c && c < 0xD7A4; }   // include/kiwi/Utils.h:64-77
	inline bool isHangulCoda(uint32_t c) { return 0x11A8 <= c && c < 0x11A8 + 27; }
	inline bool isHighSurrogate(uint32_t c) { return (c & 0xFC00) == 0xD800; }
	inline bool isLowSurrogate(uint32_t c) { return (c & 0xFC00) == 0xDC00; }
	inline uint32_t mergeSurrogate(uint32_t h, uint32_t l) { return (((h & 0x3FF) << 10) | (l & 0x3FF)) + 0x10000; }
}
