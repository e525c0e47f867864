// ORACLE (test infrastructure): CPU restatement of BestPathFinder<KnLangModel>::findBestPath, topN == 1,
// /root/reference/src/PathEvaluator.hpp:1178-1419 with
//   PathEvaluator (non-transposed)  PathEvaluator.hpp:324-635
//   RuleBasedScorer / insertToPathContainer / FormEvaluator  PathEvaluator.hpp:88-311
//   BucketedHashContainer (top1Small / top1Medium)  src/BestPathContainer.hpp:291-483
//   generateTokenList / isDisconnected  PathEvaluator.hpp:1038-1176
//   FeatureTestor  src/FeatureTestor.cpp:6-104,  UnkFormScorer::ruleBasedScore  src/UnkFormScorer.cpp:28-51
// Not restated (out of the top-1 default path): topN > 1 heaps, dialects, pretokenized spans,
// chr-model OOV scorers.  The > 512-incoming-path `top1` mode iterates a thread_local unordered_set whose
// bucket count depends on previously analysed sentences; it is restated with insertion order and counted.
#pragma once
#include <cmath>
#include <cstdlib>
#include <unordered_set>
#include "lattice.hpp"
#include "feature.hpp"
#include "knlm.hpp"
#include "cong.hpp"
#include "sbg.hpp"

namespace orc
{
	static constexpr uint8_t commonRootId = 0xFF;

	struct WordLL      // src/BestPathContainer.hpp:21-67
	{
		int32_t lmState = 0;          // Knlm: node index.  CoNg: context-trie node (the only field state equality looks at)
		uint32_t ctxIdx = 0;          // CoNg only: CoNgramState::contextIdx, carried along but not compared (CoNgramModel.hpp:491-494)
		SbHist sb;                    // SkipBigram only: ring of the last valid tokens, part of the state's identity
		uint64_t hashv = 0;           // Hash<WordLL> of the reference, kept for the `top1` container (an std::unordered_set there)
		uint8_t cmpSb = 0;
		uint8_t hashByte = 0;         // BucketedHashContainer::hashes: low byte of the hash the entry was APPENDED with (not refreshed by an overwrite)
		uint8_t prevRootId = 0, spState = 0, rootId = 0;
		int32_t morpheme = -1;
		float accScore = 0, firstChunkScore = 0, accTypoCost = 0, accDialectCost = 0;
		int32_t parentNode = -1, parentIdx = -1;
		uint32_t wid = 0;
		uint16_t ownFormId = 0;
		uint8_t combineSocket = 0;
	};

	struct PathTok { uint32_t morph; uint32_t begin, end; float wordScore; uint32_t nodeId; int32_t strOff; uint32_t strLen; };
	struct PathResult { std::vector<PathTok> path; float score = 0; uint8_t prevState = 0, curState = 0; };

	struct Counters { uint64_t lmSteps = 0, lmHops = 0, pairs = 0, inserts = 0, pathsOut = 0, top1Mode = 0, bucketFull = 0, maxNodePre = 0, maxIncoming = 0, mediumMode = 0, evalCalls = 0, candEvals = 0, maxCont = 0; };

	// per-evaluate() shape statistics (design input for the CUDA kernel's fast path; enabled by the oracle C API on request)
	struct EvalStats
	{
		// rows: one per evaluate() call -> {P, nCands, lmSteps, flags}; flags: 1 fork cand, 2 socket path, 4 shortcut cand, 8 second (ignoreCond) pass, 16 left-half cand, 32 socket-chunk cand
		std::vector<uint32_t> rows;
	};

	inline uint8_t hashSbTypeOrder(uint8_t type, uint8_t order) { return ((type << 1) ^ (type >> 7) ^ order) % 63 + 1; }   // PathEvaluator.hpp:83-86

	inline size_t getSBType(const u16* form, size_t len)      // src/Utils.cpp:264-298 (on the un-joined kform)
	{
		size_t format = 0, group = 0;
		uint32_t chr = form[0];
		if (form[len - 1] == '.') format = 1;
		else if (form[len - 1] == ')')
		{
			if (form[0] == '(') { chr = form[1]; format = 2; }
			else format = 3;
		}
		if (0xAC00 <= chr && chr <= 0xD7A3) group = 1;
		else if (0x3131 <= chr && chr <= 0x314E) group = 2;
		else if ('0' <= chr && chr <= '9') group = 3;
		else if (0x2160 <= chr && chr <= 0x216B) group = 4;
		else if (0x2170 <= chr && chr <= 0x217B) group = 5;
		else if (0x2460 <= chr && chr <= 0x2473) return 24;
		else if (0x2780 <= chr && chr <= 0x2789) return 24;
		else if (0x2776 <= chr && chr <= 0x277F) return 25;
		else if (0x278A <= chr && chr <= 0x2793) return 25;
		else if (0x2474 <= chr && chr <= 0x2487) return 26;
		else if (0x2488 <= chr && chr <= 0x249B) return 27;
		return format | (group << 2);
	}

	struct Viterbi
	{
		const Image& im;
		Knlm lm;
		Cong cg;
		Sbg sbgm;
		const bool sbg;                // model_type == ModelType::sbg: Knlm + SkipBigram (non-transposed evaluator, state = node + history ring)
		const bool cong;               // model_type == ModelType::cong: transposed evaluation (PathEvaluator.hpp:837-1036, CoNgramModel.cpp:17-317)
		kb2_config cfg;
		Counters* cnt = nullptr;
		WorkCounters* wc = nullptr;
		EvalStats* es = nullptr;

		// per-call state
		const LNode* graph = nullptr; size_t graphSize = 0;
		const u16* norm = nullptr;
		std::vector<std::vector<WordLL>> cache;
		std::vector<std::pair<int32_t, uint32_t>> ownFormList;     // (offset into norm or ~formIdx, len)
		std::vector<uint8_t> uniqStates;
		bool splitSaisiot = false, mergeSaisiot = false, splitComplex = false;

		explicit Viterbi(const Image& _im) : im{ _im }, lm{ _im }, cg{ _im }, sbgm{ _im }, sbg{ _im.h->model_type == 3 }, cong{ _im.h->model_type == 4 }, cfg{ _im.h->config } {}

		// --- small accessors
		const kb2_morph& M(int32_t id) const { return im.morphs[id]; }
		bool isSingle(const kb2_morph& m) const { return m.chunk_cnt == 0 || (m.flags & (KB2_MORPH_COMPLEX | KB2_MORPH_SAISIOT)); }
		const u16* kformPtr(const kb2_morph& m) const { return m.form_idx >= 0 ? im.formStr(m.form_idx) : nullptr; }
		uint32_t kformLen(const kb2_morph& m) const { return m.form_idx >= 0 ? im.formLen(m.form_idx) : 0; }
		// AnalyzeOption::blocklist, Morpheme::hasMorpheme (include/kiwi/Form.h:187-196): the candidate's combined morpheme or one of its chunks is listed
		std::vector<uint32_t> blocklist;      // sorted morpheme ids; empty = none
		bool blocked(int32_t id) const
		{
			if (blocklist.empty()) return false;
			const auto& m = M(id);
			if (std::binary_search(blocklist.begin(), blocklist.end(), (uint32_t)(id + m.combined))) return true;
			for (uint32_t c = 0; c < m.chunk_cnt; ++c) if (std::binary_search(blocklist.begin(), blocklist.end(), (uint32_t)im.chunks[m.chunk_off + c].morph)) return true;
			return false;
		}
		bool hasComplex(int32_t id) const                                                   // Form.h:176-185
		{
			const auto& m = M(id);
			if (M(id + m.combined).flags & KB2_MORPH_COMPLEX) return true;
			for (uint32_t c = 0; c < m.chunk_cnt; ++c) if (M(im.chunks[m.chunk_off + c].morph).flags & KB2_MORPH_COMPLEX) return true;
			return false;
		}
		void ownForm(uint16_t id, const u16*& p, uint32_t& len) const
		{
			const auto& o = ownFormList[id - 1];
			if (o.first >= 0) p = norm + o.first; else p = im.formStr(~o.first);
			len = o.second;
		}

		// PathEvaluator.hpp:22-44
		bool hasLeftBoundary(const LNode* node) const
		{
			const LNode* prev = node - node->prev;
			if (prev->endPos == 0) return true;
			if (prev->endPos < node->startPos) return true;
			if (prev->uformLen)
			{
				const u16 c = norm[prev->uformOff + prev->uformLen - 1];
				const uint8_t tag = im.cls(c);
				if (tag == T_ssc || c == '"' || c == '\'') return false;
				if (T_sf <= tag && tag <= T_sb) return true;
			}
			return false;
		}

		struct RuleScorer      // PathEvaluator.hpp:88-184
		{
			int specialType; size_t sbType; int sbOrder;
			bool vowelE, infJ, badPairOfL, positiveE, contractableE, snEndswithPoint;
			uint8_t condP;
		};

		RuleScorer makeRuleScorer(int32_t curId, const LNode* node) const
		{
			const auto& m = M(curId);
			const u16* kf = kformPtr(m); const uint32_t kl = kformLen(m);
			RuleScorer r;
			r.specialType = 6;
			for (int i = 0; i < 6; ++i) if ((uint32_t)curId == im.h->special_morph_ids[i]) { r.specialType = i; break; }
			r.sbType = m.tag == T_sb ? getSBType(kf, kl) : 0;
			r.sbOrder = r.sbType ? m.sense_id : 0;
			const u16 k0 = kl ? kf[0] : 0;
			r.vowelE = isEClass(m.tag) && kf && (0xC544 <= k0 && k0 <= 0xC774);
			r.infJ = (m.tag == T_jks || m.tag == T_jkc) && kl == 1 && kf[0] == 0xAC00;
			r.badPairOfL = (k0 == 0xC73C || k0 == 0xB290 || (0xC0AC <= k0 && k0 <= 0xC2DC));
			r.positiveE = isEClass(m.tag) && node->form >= 0 && im.formStr(node->form)[0] == 0xC544 && true;
			r.contractableE = isEClass(m.tag) && kf && kl && kf[0] == 0xC5B4;
			r.snEndswithPoint = m.tag == T_sn && node->uformLen && norm[node->uformOff + node->uformLen - 1] == '.';
			r.condP = m.polar;
			return r;
		}

		float ruleScore(const RuleScorer& r, int32_t prevId, uint8_t prevSp) const
		{
			const auto& pm = M(prevId);
			const u16* kf = kformPtr(pm); const uint32_t kl = kformLen(pm);
			float acc = 0;
			if (r.vowelE && isIrregular(pm.tag)) acc -= 10;
			if (r.infJ && pm.tag == T_np && kl == 1 && (kf[0] == 0xB098 || kf[0] == 0xB108 || kf[0] == 0xC800)) acc -= 5;
			if (r.badPairOfL && isVerbClass(pm.tag) && kf && kl && kf[kl - 1] == 0x11AF) acc -= 7;
			if (r.positiveE && !(isVerbClass(pm.tag) && ftPolar(kf, kf + kl, CP_positive))) acc -= 100;
			if (r.contractableE && isVerbClass(pm.tag) && kf && kl && !isHangulCoda(kf[kl - 1])) acc -= 3;
			if (r.condP == CP_non_adj && (pm.tag == T_va || pm.tag == T_xsa)) acc -= 10;
			const uint8_t sq = prevSp & 1, dq = (prevSp >> 1) & 1, bh = prevSp >> 2;
			if (r.specialType <= 2) { if ((uint8_t)r.specialType != sq) acc -= 2; }
			else if (r.specialType <= 5) { if ((uint8_t)(r.specialType - 3) != dq) acc -= 2; }
			if (r.sbType == 5) acc -= 5;
			if (r.sbType && isEClass(pm.tag) && pm.tag != T_ef) acc -= 10;
			if (r.sbType && bh == hashSbTypeOrder((uint8_t)r.sbType, (uint8_t)r.sbOrder)) acc += 3;
			if (r.snEndswithPoint && (pm.tag == T_unknown || pm.tag == T_ef || pm.tag == T_sf)) acc -= 5;
			return acc;
		}

		// ---- BucketedHashContainer, src/BestPathContainer.hpp:291-483
		// The `top1` mode (> 512 incoming paths, BestPathContainer.hpp:229-276) is an std::unordered_set in the reference: its
		// iteration order is libstdc++'s bucket order, which depends on the bucket count the set has grown to.  The restatement
		// uses the same library container with the reference's hash and equality; the analyzer starts every sentence with a
		// fresh set (resetHistory()), which is what reproduces the reference's dumps best (measured on the SkipBigram vectors,
		// where this container is used hundreds of times per sentence; letting the set keep its growth across sentences like
		// a thread_local would, or renewing it per candidate, both match fewer sentences).
		struct WHash { size_t operator()(const WordLL& w) const { return (size_t)w.hashv; } };
		struct WEq
		{
			bool operator()(const WordLL& a, const WordLL& b) const
			{
				return a.prevRootId == b.prevRootId && a.spState == b.spState && a.lmState == b.lmState && (!a.cmpSb || a.sb == b.sb);
			}
		};
		struct Container
		{
			std::vector<WordLL> buckets[4];
			std::unordered_set<WordLL, WHash, WEq> top1;
			int mode = 0;   // 0 small (1 bucket), 1 medium (4 buckets), 2 "top1" (unordered_set)
			void clear() { for (auto& b : buckets) b.clear(); top1.clear(); }
		};
		Container cont;
		void resetHistory(size_t buckets = 0) { cont.top1 = std::unordered_set<WordLL, WHash, WEq>{}; if (buckets) cont.top1.rehash(buckets); }

		void contInsert(uint8_t prevRootId, uint8_t rootId, int32_t morph, float accScore, float firstChunkScore,
			float accTypoCost, float accDialectCost, int32_t pNode, int32_t pIdx, uint8_t parentRootId, int32_t lmState, uint8_t spState, uint32_t ctxIdx = 0, const SbHist* sbh = nullptr)
		{
			if (cnt) cnt->inserts++;
			if (wc) wc->pathsWritten++;
			uint64_t h = (uint64_t)(int64_t)lmState;                                       // Knlm.hpp:1170-1178 std::hash<int32_t>
			if (sbg && sbh)                                                                 // SkipBigramModel.hpp:188-203: history folded into the Knlm hash
			{
				for (int i = 0; i < 8; ++i) h = (uint64_t)sbh->hist[i] ^ ((h << 3) | (h >> 61));
			}
			if (cong)                                                                       // CoNgramModel.hpp:505-541 Hash<uint32_t>(state.node)
			{
				const uint64_t v = (uint32_t)lmState;
				h = (v * 2305843009213693951ull) ^ ((v << 33) | (v >> 31));
			}
			h = ((uint16_t)prevRootId | ((uint16_t)spState << 8)) ^ ((h << 3) | (h >> 61)); // BestPathContainer.hpp:79-84
			if (cont.mode == 2)
			{
				WordLL w;
				w.morpheme = morph; w.accScore = accScore; w.firstChunkScore = firstChunkScore; w.accTypoCost = accTypoCost;
				w.accDialectCost = accDialectCost; w.parentNode = pNode; w.parentIdx = pIdx; w.lmState = lmState; w.spState = spState; w.ctxIdx = ctxIdx;
				if (sbh) { w.sb = *sbh; w.cmpSb = sbg ? 1 : 0; }
				w.rootId = parentRootId;
				w.prevRootId = prevRootId;
				if (rootId != commonRootId) w.rootId = rootId;
				w.hashv = h;
				auto ins = cont.top1.emplace(w);
				if (!ins.second)
				{
					auto& target = const_cast<WordLL&>(*ins.first);      // as the reference does: the key fields are equal
					if (accScore > target.accScore) target = w;
				}
				return;
			}
			const size_t bucket = cont.mode == 1 ? ((h >> 8) & 3) : 0;
			auto& value = cont.buckets[bucket];
			// insertOptimized<avx2> (BestPathContainer.hpp:316-383 with nst::findAll<avx2>, search.cpp:948-968), as it BEHAVES on x86-64:
			//  - the candidates of the first 64 entries are findAll(hash, min(n, 64), h): a byte-compare mask ANDed with ((size_t)1 << size) - 1.
			//    For size == 64 that shift is by the register width: x86 masks the count to 0, the mask becomes 0 and NO entry of the first
			//    64 is ever a candidate once the bucket holds 64 entries;
			//  - the candidates of entries 64.. are findAll(hash + 64, n - 64, h) (again empty when n - 64 == 64), and the equality test of
			//    candidate i looks at value[i] where it means value[64 + i]: when value[i] is the new state, entry 64 + i - some other state
			//    with the same hash byte, or an earlier duplicate - is the one that gets compared by score and overwritten;
			//  - for 32 < size < 64 the low movemask is sign-extended: when byte 31 matches, every position 32 .. size-1 becomes a candidate.
			//  - no candidate matches: the state is appended (a second time, if it lives among the first 64) while there is room.
			// Knlm / CoNg buckets rarely reach 64 states; SkipBigram states (8-token ring) rarely merge and live in this regime.
			static const bool searchAll = std::getenv("ORC_BUCKET_SEARCH_ALL") != nullptr;      // (experiments: what the code comments intend)
			const size_t nVal = value.size();
			auto eq = [&](size_t k) { return value[k].prevRootId == prevRootId && value[k].spState == spState && value[k].lmState == lmState && (!(sbg && sbh) || value[k].sb == *sbh); };
			auto candMask = [&](size_t from, size_t size) -> uint64_t      // findAllAVX2(hash + from, size, h)
			{
				uint64_t lo = 0, hi = 0;
				for (size_t k = 0; k < size && k < 32; ++k) if (value[from + k].hashByte == (uint8_t)h) lo |= 1ull << k;
				if (size <= 32) return lo & ((1ull << size) - 1);
				for (size_t k = 32; k < size; ++k) if (value[from + k].hashByte == (uint8_t)h) hi |= 1ull << k;
				// NB: the reference loads 32 bytes past `size` as well (stale hash bytes of the array); they fall to the final mask
				if (lo & 0x80000000ull) lo |= 0xFFFFFFFF00000000ull;      // (size_t)(int)movemask: sign extension
				const uint64_t mask = size >= 64 ? 0ull : ((1ull << size) - 1);      // shift count taken mod 64 by the hardware
				return (lo | hi) & mask;
			};
			size_t it = nVal;
			if (searchAll) { for (size_t k = 0; k < nVal; ++k) if (eq(k)) { it = k; break; } }
			else
			{
				uint64_t b0 = candMask(0, std::min<size_t>(nVal, 64));
				for (; b0 && it == nVal; b0 &= b0 - 1) { const size_t k = (size_t)__builtin_ctzll(b0); if (eq(k)) it = k; }
				if (it == nVal && nVal > 64)
				{
					uint64_t b1 = candMask(64, nVal - 64);
					for (; b1 && it == nVal; b1 &= b1 - 1) { const size_t k = (size_t)__builtin_ctzll(b1); if (eq(k)) it = k + 64; }
				}
			}
			if (it >= value.size())
			{
				if (cont.mode == 2 || value.size() < 128)
				{
					WordLL w;
					w.morpheme = morph; w.accScore = accScore; w.firstChunkScore = firstChunkScore; w.accTypoCost = accTypoCost;
					w.accDialectCost = accDialectCost; w.parentNode = pNode; w.parentIdx = pIdx; w.lmState = lmState; w.spState = spState; w.ctxIdx = ctxIdx;
					if (sbh) w.sb = *sbh;
					w.rootId = parentRootId;
					w.prevRootId = prevRootId;
					if (rootId != commonRootId) w.rootId = rootId;
					w.hashByte = (uint8_t)h;
					value.push_back(w);
				}
				else if (cnt) cnt->bucketFull++;
			}
			else
			{
				auto& t = value[it];
				if (accScore > t.accScore)
				{
					t.morpheme = morph; t.accScore = accScore; t.firstChunkScore = firstChunkScore; t.accTypoCost = accTypoCost;
					t.accDialectCost = accDialectCost; t.parentNode = pNode; t.parentIdx = pIdx; t.lmState = lmState; t.spState = spState; t.ctxIdx = ctxIdx;
					if (sbh) t.sb = *sbh;
					t.rootId = parentRootId;
					if (rootId != commonRootId) t.rootId = rootId;
				}
			}
		}

		// PathEvaluator.hpp:193-251
		void insertToPathContainer(int32_t curId, int32_t lmState, float score, float firstChunkScore, const LNode* node,
			const WordLL& prevPath, int32_t pNode, int32_t pIdx, const RuleScorer& rs, uint32_t ctxIdx = 0, const SbHist* sbh = nullptr)
		{
			auto insert = [&](uint8_t rootId)
			{
				uint8_t spState = prevPath.spState;
				if (rootId != commonRootId) spState = uniqStates[rootId];
				const float rsc = ruleScore(rs, (int32_t)prevPath.wid, spState);
				const float candScoreWithRule = score + rsc;
				const float firstChunkScoreWithRule = firstChunkScore + rsc;
				if (rs.specialType == 0) spState |= 1;
				else if (rs.specialType == 1) spState &= ~1;
				else if (rs.specialType == 3) spState |= 2;
				else if (rs.specialType == 4) spState &= ~2;
				if (rs.sbType) spState = (spState & 3) | (uint8_t)(hashSbTypeOrder((uint8_t)rs.sbType, (uint8_t)(rs.sbOrder + 1)) << 2);
				const float curDialectCost = 0.f;      // standard dialect only
				contInsert(prevPath.rootId, rootId, curId, candScoreWithRule - curDialectCost, firstChunkScoreWithRule - curDialectCost,
					prevPath.accTypoCost + node->typoCost, prevPath.accDialectCost + curDialectCost, pNode, pIdx, prevPath.rootId, lmState, spState, ctxIdx, sbh);
			};
			const bool quote = rs.specialType == 0 || rs.specialType == 1 || rs.specialType == 3 || rs.specialType == 4;
			if ((rs.sbType || quote) && prevPath.rootId == commonRootId)
			{
				for (uint8_t rootId = 0; rootId < uniqStates.size(); ++rootId) insert(rootId);
			}
			else insert(commonRootId);
		}

		float lmNext(int32_t& state, uint32_t wid) { if (cnt) cnt->lmSteps++; return lm.progress(state, wid); }
		float lmNext(int32_t& state, SbHist& sbh, uint32_t wid)
		{
			if (!sbg) return lmNext(state, wid);
			if (cnt) cnt->lmSteps++;
			return sbgm.next(lm, state, sbh, wid);
		}

		// PathEvaluator.hpp:514-634
		void evalSingleMorpheme(std::vector<WordLL>& resultOut, size_t nodeIdx, size_t ownFormId, int32_t curId,
			float ignoreCondScore, float nodeLevelDiscount)
		{
			const LNode* node = graph + nodeIdx;
			const auto& cur = M(curId);
			const uint32_t langVocabSize = im.h->lang_vocab_size;
			int32_t lastMorph; uint32_t firstWid;
			if (isSingle(cur)) { lastMorph = cur.combined ? curId + cur.combined : curId; firstWid = cur.lm_morpheme_id; }
			else { lastMorph = (int32_t)im.chunks[cur.chunk_off + cur.chunk_cnt - 1].morph; firstWid = M(im.chunks[cur.chunk_off].morph).lm_morpheme_id; }
			uint32_t lastSeqId;
			if ((uint32_t)lastMorph >= langVocabSize && (uint32_t)lastMorph < im.h->n_morphs) lastSeqId = (uint32_t)lastMorph;
			else lastSeqId = M(lastMorph).lm_morpheme_id;

			cont.clear();
			const float additionalScore = cur.user_score + nodeLevelDiscount + im.h->tag_left_boundary[hasLeftBoundary(node) ? 1 : 0][clearIrregular(cur.tag)];
			const RuleScorer rs = makeRuleScorer(curId, node);
			const bool allowedSpaceBetweenChunk = cfg.space_tolerance > 0;

			const LNode* prev = node->prev ? node - node->prev : nullptr;
			for (; prev; prev = prev->sibling ? prev + prev->sibling : nullptr)
			{
				const int32_t pNode = (int32_t)(prev - graph);
				auto& pc = cache[pNode];
				for (size_t pi = 0; pi < pc.size(); ++pi)
				{
					const WordLL& prevPath = pc[pi];
					if (cnt) cnt->pairs++;
					if (wc) wc->pairs++;
					if (M(prevPath.morpheme).tag == T_z_siot && (!isNNClass(cur.tag) || prev->endPos < node->startPos)) continue;
					float candScore = prevPath.accScore + additionalScore;
					float firstChunkScore = additionalScore;
					if (prevPath.combineSocket)
					{
						if (prevPath.combineSocket != cur.combine_socket || isSingle(cur)) continue;
						if (prev->endPos < node->startPos)
						{
							if (allowedSpaceBetweenChunk) candScore -= cfg.space_penalty;
							else continue;
						}
						const auto& pw = M((int32_t)prevPath.wid);
						firstWid = M((int32_t)prevPath.wid + pw.combined).lm_morpheme_id;      // NB: persists for later pairs, as in the reference (:590)
					}
					// FormEvaluator, PathEvaluator.hpp:253-311
					{
						const u16* lf; uint32_t ll;
						const auto& pwm = M((int32_t)prevPath.wid);
						if (prevPath.ownFormId) ownForm(prevPath.ownFormId, lf, ll);
						else if (pwm.form_idx >= 0 && kformLen(pwm)) { lf = kformPtr(pwm); ll = kformLen(pwm); }
						else { const auto& pm = M(prevPath.morpheme); lf = kformPtr(pm); ll = kformLen(pm); }
						const bool leftSSC = ll && im.cls(lf[ll - 1]) == T_ssc;
						const uint8_t prevTag = M(prevPath.morpheme).tag;
						if (prevTag == T_ssc || leftSSC) {}
						else if (ignoreCondScore != 0)
						{
							candScore += (ftVowel(lf, lf + ll, cur.vowel) && ftPolar(lf, lf + ll, cur.polar)) ? 0 : ignoreCondScore;
						}
						else
						{
							if (!(ftVowel(lf, lf + ll, cur.vowel) && ftPolar(lf, lf + ll, cur.polar))) continue;
						}
					}
					int32_t cLmState = prevPath.lmState;
					SbHist cSb = prevPath.sb;
					if (cur.combine_socket && isSingle(cur)) {}
					else
					{
						if (M((int32_t)firstWid).tag == T_p) continue;
						float ll = lmNext(cLmState, cSb, firstWid);
						candScore += ll;
						firstChunkScore += ll;
						if (!isSingle(cur))
						{
							bool prohibited = false;
							for (uint32_t i = 1; i < cur.chunk_cnt; ++i)
							{
								const uint32_t wid = M((int32_t)im.chunks[cur.chunk_off + i].morph).lm_morpheme_id;
								if (M((int32_t)wid).tag == T_p) { prohibited = true; break; }
								ll = lmNext(cLmState, cSb, wid);
								candScore += ll;
							}
							if (prohibited) continue;
						}
					}
					insertToPathContainer(curId, cLmState, candScore, firstChunkScore, node, prevPath, pNode, (int32_t)pi, rs, 0, sbg ? &cSb : nullptr);
				}
			}
			if (cnt) { size_t tot = 0; for (auto& b : cont.buckets) tot += b.size(); cnt->maxCont = std::max<uint64_t>(cnt->maxCont, tot); }
			// writeTo, BestPathContainer.hpp:451-469
			auto emit = [&](const WordLL& p)
			{
				resultOut.push_back(p);
				auto& np = resultOut.back();
				np.wid = lastSeqId;
				if (isSingle(cur)) { np.combineSocket = cur.combine_socket; np.ownFormId = (uint16_t)ownFormId; }
			};
			{ static const bool tb = std::getenv("ORC_TRACE_BIG") != nullptr; if (tb && cont.mode != 2) { size_t mx = 0, tot = 0; for (auto& b : cont.buckets) { mx = std::max(mx, b.size()); tot += b.size(); } std::fprintf(stderr, "[cont] mode %d total %zu maxbucket %zu\n", cont.mode, tot, mx); } }
			if (std::getenv("ORC_TRACE_CAND")) std::fprintf(stderr, "[cand] node %zu cand %d mode %d E %zu buckets %zu %zu %zu %zu\n", nodeIdx, curId, cont.mode, cont.mode == 2 ? cont.top1.size() : cont.buckets[0].size() + cont.buckets[1].size() + cont.buckets[2].size() + cont.buckets[3].size(), cont.buckets[0].size(), cont.buckets[1].size(), cont.buckets[2].size(), cont.buckets[3].size());
			if (cont.mode == 2) { for (auto& p : cont.top1) emit(p); }      // libstdc++ iteration order, as in the reference
			else for (auto& bk : cont.buckets) for (auto& p : bk) emit(p);
		}


		// FormEvaluator, PathEvaluator.hpp:253-311 (the CoNg evaluator calls it in a different place than evalSingleMorpheme)
		bool formEval(const WordLL& prevPath, const kb2_morph& cur, float ignoreCondScore, float& score) const
		{
			const u16* lf; uint32_t ll;
			const auto& pwm = M((int32_t)prevPath.wid);
			if (prevPath.ownFormId) ownForm(prevPath.ownFormId, lf, ll);
			else if (pwm.form_idx >= 0 && kformLen(pwm)) { lf = kformPtr(pwm); ll = kformLen(pwm); }
			else { const auto& pm = M(prevPath.morpheme); lf = kformPtr(pm); ll = kformLen(pm); }
			const bool leftSSC = ll && im.cls(lf[ll - 1]) == T_ssc;
			const uint8_t prevTag = M(prevPath.morpheme).tag;
			if (prevTag == T_ssc || leftSSC) return true;
			const bool ok = ftVowel(lf, lf + ll, cur.vowel) && ftPolar(lf, lf + ll, cur.polar);
			if (ignoreCondScore != 0) { score += ok ? 0 : ignoreCondScore; return true; }
			return ok;
		}

		uint32_t lastSeqIdOf(int32_t curId) const            // CoNgramModel.cpp:147-168 (same rule as PathEvaluator.hpp:536-556)
		{
			const auto& cur = M(curId);
			int32_t lastMorph;
			if (isSingle(cur)) lastMorph = cur.combined ? curId + cur.combined : curId;
			else lastMorph = (int32_t)im.chunks[cur.chunk_off + cur.chunk_cnt - 1].morph;
			if ((uint32_t)lastMorph >= im.h->lang_vocab_size && (uint32_t)lastMorph < im.h->n_morphs) return (uint32_t)lastMorph;
			return M(lastMorph).lm_morpheme_id;
		}

		void writeTo(std::vector<WordLL>& resultOut, int32_t curId, uint32_t lastSeqId, size_t ownFormId)     // BestPathContainer.hpp:451-469
		{
			const auto& cur = M(curId);
			if (cnt) { size_t tot = 0; for (auto& b : cont.buckets) tot += b.size(); cnt->maxCont = std::max<uint64_t>(cnt->maxCont, tot); }
			auto emit = [&](const WordLL& p)
			{
				resultOut.push_back(p);
				auto& np = resultOut.back();
				np.wid = lastSeqId;
				if (isSingle(cur)) { np.combineSocket = cur.combine_socket; np.ownFormId = (uint16_t)ownFormId; }
			};
			if (cont.mode == 2) { for (auto& p : cont.top1) emit(p); }      // libstdc++ iteration order, as in the reference
			else for (auto& bk : cont.buckets) for (auto& p : bk) emit(p);
		}

		// MorphemeEvaluator<CoNgramState>::eval, src/CoNgramModel.cpp:17-317
		void evalCong(std::vector<WordLL>& resultOut, size_t nodeIdx, size_t ownFormId, const std::vector<int32_t>& morphs,
			float ignoreCondScore, float nodeLevelDiscount)
		{
			const LNode* node = graph + nodeIdx;
			struct PP { const LNode* prev; int32_t pNode, pi; };
			std::vector<PP> regularPrev, combiningPrev;
			for (const LNode* prev = node->prev ? node - node->prev : nullptr; prev; prev = prev->sibling ? prev + prev->sibling : nullptr)
			{
				const int32_t pNode = (int32_t)(prev - graph);
				for (size_t pi = 0; pi < cache[pNode].size(); ++pi)
				{
					(cache[pNode][pi].combineSocket ? combiningPrev : regularPrev).push_back(PP{ prev, pNode, (int32_t)pi });
				}
			}
			std::vector<int32_t> regularMorphs, combiningL, combiningR;
			std::vector<uint32_t> nextWids;
			for (int32_t curId : morphs)
			{
				const auto& cur = M(curId);
				if (cur.combine_socket) { (isSingle(cur) ? combiningL : combiningR).push_back(curId); continue; }
				const uint32_t firstWid = isSingle(cur) ? cur.lm_morpheme_id : M((int32_t)im.chunks[cur.chunk_off].morph).lm_morpheme_id;
				if (M((int32_t)firstWid).tag == T_p) continue;
				regularMorphs.push_back(curId);
				nextWids.push_back(firstWid);
			}
			// progressMatrix (CoNgramModel.cpp:124-139, 1494-1611): all (prevState x firstWid) scores and next states at once
			const size_t P = regularPrev.size(), W = nextWids.size();
			std::vector<float> scores(P * W);
			std::vector<int32_t> nextNode(P * W);
			std::vector<uint32_t> nextCtx(P * W);
			if (P && W)
			{
				if (P == 1 && W == 1)
				{
					const WordLL& pp = cache[regularPrev[0].pNode][regularPrev[0].pi];
					nextNode[0] = pp.lmState; nextCtx[0] = pp.ctxIdx;
					scores[0] = cg.next(nextNode[0], nextCtx[0], nextWids[0]);
					if (cnt) cnt->lmSteps++;
				}
				else
				{
					std::vector<uint32_t> uc, uw(nextWids);
					for (auto& pp : regularPrev) uc.push_back(cache[pp.pNode][pp.pi].ctxIdx);
					std::sort(uc.begin(), uc.end()); uc.erase(std::unique(uc.begin(), uc.end()), uc.end());
					std::sort(uw.begin(), uw.end()); uw.erase(std::unique(uw.begin(), uw.end()), uw.end());
					const auto ep = Cong::epilogueOf(uc.size(), uw.size());
					if (wc) { wc->cgRows += uc.size() + uw.size(); wc->cgMacs += (uint64_t)uc.size() * uw.size() * cg.dim; }
					for (size_t i = 0; i < P; ++i)
					{
						const WordLL& pp = cache[regularPrev[i].pNode][regularPrev[i].pi];
						for (size_t j = 0; j < W; ++j)
						{
							scores[i * W + j] = cg.finish(pp.ctxIdx, nextWids[j], ep);
							nextNode[i * W + j] = pp.lmState;
							nextCtx[i * W + j] = cg.step(nextNode[i * W + j], nextWids[j]);
							if (cnt) cnt->lmSteps++;
							if (wc) wc->lmSteps++;
						}
					}
				}
			}
			const bool allowedSpaceBetweenChunk = cfg.space_tolerance > 0;
			const float lb = 0;
			(void)lb;
			for (size_t curIdx = 0; curIdx < regularMorphs.size(); ++curIdx)
			{
				const int32_t curId = regularMorphs[curIdx];
				const auto& cur = M(curId);
				cont.clear();
				const size_t length = isSingle(cur) ? 1 : cur.chunk_cnt;
				const RuleScorer rs = makeRuleScorer(curId, node);
				const float morphScore = cur.user_score + nodeLevelDiscount + im.h->tag_left_boundary[hasLeftBoundary(node) ? 1 : 0][clearIrregular(cur.tag)];
				for (size_t prevId = 0; prevId < P; ++prevId)
				{
					const auto& rp = regularPrev[prevId];
					const WordLL& prevPath = cache[rp.pNode][rp.pi];
					if (cnt) cnt->pairs++;
					if (wc) wc->pairs++;
					int32_t stNode = nextNode[prevId * W + curIdx]; uint32_t stCtx = nextCtx[prevId * W + curIdx];
					float score = prevPath.accScore + morphScore + scores[prevId * W + curIdx];
					const float firstChunkScore = morphScore + scores[prevId * W + curIdx];
					if (!formEval(prevPath, cur, ignoreCondScore, score)) continue;
					if (M(prevPath.morpheme).tag == T_z_siot && (!isNNClass(cur.tag) || rp.prev->endPos < node->startPos)) continue;
					bool prohibited = false;
					for (size_t i = 1; i < length; ++i)
					{
						const uint32_t wid = M((int32_t)im.chunks[cur.chunk_off + i].morph).lm_morpheme_id;
						if (M((int32_t)wid).tag == T_p) { prohibited = true; break; }
						score += cg.next(stNode, stCtx, wid);
						if (cnt) cnt->lmSteps++;
					}
					if (prohibited) continue;
					insertToPathContainer(curId, stNode, score, firstChunkScore, node, prevPath, rp.pNode, rp.pi, rs, stCtx);
				}
				writeTo(resultOut, curId, lastSeqIdOf(curId), ownFormId);
			}
			for (int32_t curId : combiningL)
			{
				const auto& cur = M(curId);
				cont.clear();
				const RuleScorer rs = makeRuleScorer(curId, node);
				const float morphScore = cur.user_score + nodeLevelDiscount + im.h->tag_left_boundary[hasLeftBoundary(node) ? 1 : 0][clearIrregular(cur.tag)];
				for (auto& rp : regularPrev)
				{
					const WordLL& prevPath = cache[rp.pNode][rp.pi];
					if (cnt) cnt->pairs++;
					if (wc) wc->pairs++;
					float score = prevPath.accScore + morphScore;
					const float firstChunkScore = morphScore;
					if (!formEval(prevPath, cur, ignoreCondScore, score)) continue;
					insertToPathContainer(curId, prevPath.lmState, score, firstChunkScore, node, prevPath, rp.pNode, rp.pi, rs, prevPath.ctxIdx);
				}
				writeTo(resultOut, curId, lastSeqIdOf(curId), ownFormId);
			}
			for (int32_t curId : combiningR)
			{
				const auto& cur = M(curId);
				cont.clear();
				const size_t length = isSingle(cur) ? 1 : cur.chunk_cnt;
				const RuleScorer rs = makeRuleScorer(curId, node);
				const float morphScore = cur.user_score + nodeLevelDiscount + im.h->tag_left_boundary[hasLeftBoundary(node) ? 1 : 0][clearIrregular(cur.tag)];
				for (auto& rp : combiningPrev)
				{
					const WordLL& prevPath = cache[rp.pNode][rp.pi];
					if (cnt) cnt->pairs++;
					if (wc) wc->pairs++;
					float score = prevPath.accScore + morphScore;
					float firstChunkScore = 0;
					if (prevPath.combineSocket != cur.combine_socket || isSingle(cur)) continue;
					if (rp.prev->endPos < node->startPos)
					{
						if (allowedSpaceBetweenChunk) score -= cfg.space_penalty;
						else continue;
					}
					const auto& pw = M((int32_t)prevPath.wid);
					const uint32_t firstWid = M((int32_t)prevPath.wid + pw.combined).lm_morpheme_id;
					if (!formEval(prevPath, cur, ignoreCondScore, score)) continue;
					int32_t stNode = prevPath.lmState; uint32_t stCtx = prevPath.ctxIdx;
					score += (firstChunkScore = cg.next(stNode, stCtx, firstWid));
					if (cnt) cnt->lmSteps++;
					firstChunkScore += morphScore;
					bool prohibited = false;
					for (size_t i = 1; i < length; ++i)
					{
						const uint32_t wid = M((int32_t)im.chunks[cur.chunk_off + i].morph).lm_morpheme_id;
						if (M((int32_t)wid).tag == T_p) { prohibited = true; break; }
						score += cg.next(stNode, stCtx, wid);
						if (cnt) cnt->lmSteps++;
					}
					if (prohibited) continue;
					insertToPathContainer(curId, stNode, score, firstChunkScore, node, prevPath, rp.pNode, rp.pi, rs, stCtx);
				}
				writeTo(resultOut, curId, lastSeqIdOf(curId), ownFormId);
			}
		}

		// PathEvaluator<LmState, transposed>::operator(), PathEvaluator.hpp:860-1035
		void evaluateCong(size_t nodeIdx, size_t ownFormId, const uint32_t* cands, size_t nCands, float nodeLevelDiscount, size_t totalPrevPathes)
		{
			const LNode* node = graph + nodeIdx;
			auto& nCache = cache[nodeIdx];
			int32_t zCodaMorph = -1, zSiotMorph = -1;
			std::vector<int32_t> validMorphCands;
			for (size_t ci = 0; ci < nCands; ++ci)
			{
				const int32_t curId = (int32_t)cands[ci];
				const auto& cur = M(curId);
				if (splitComplex && hasComplex(curId)) continue;
				if (blocked(curId)) continue;
				if (cur.dialect != 0) continue;
				if (cur.tag == T_z_coda) { zCodaMorph = curId; continue; }
				if (cur.tag == T_z_siot) { zSiotMorph = curId; continue; }
				if (!isSingle(cur))
				{
					const u16* kf = kformPtr(cur);
					const auto& c0 = M((int32_t)im.chunks[cur.chunk_off].morph);
					if (node->prev && (node - node->prev)->endPos < node->startPos
						&& kf && kformLen(cur) == 1 && (kf[0] == 0xB2E4 || kf[0] == 0xAC8C || kf[0] == 0xC9C0)
						&& kformPtr(c0) && kformLen(c0) == 1 && kformPtr(c0)[0] == 0xD558)
					{
						continue;
					}
				}
				validMorphCands.push_back(curId);
			}
			auto shortcut = [&](int32_t morphId, bool coda)
			{
				const auto& cur = M(morphId);
				for (const LNode* prev = node->prev ? node - node->prev : nullptr; prev; prev = prev->sibling ? prev + prev->sibling : nullptr)
				{
					const int32_t pNode = (int32_t)(prev - graph);
					for (size_t pi = 0; pi < cache[pNode].size(); ++pi)
					{
						const WordLL& p = cache[pNode][pi];
						const uint8_t lastTag = M((int32_t)p.wid).tag;
						if (coda) { if (!isJClass(lastTag) && !isEClass(lastTag)) continue; }
						else { if (!isNNClass(lastTag)) continue; }
						WordLL np = p;
						np.accScore += cur.user_score * cfg.typo_cost_weight;
						np.accTypoCost -= cur.user_score;
						np.parentNode = pNode; np.parentIdx = (int32_t)pi;
						np.morpheme = (int32_t)cur.lm_morpheme_id;
						np.wid = cur.lm_morpheme_id;
						nCache.push_back(np);
					}
				}
			};
			for (int ignoreCond = 0; ignoreCond < 2; ++ignoreCond)
			{
				if (zCodaMorph >= 0) shortcut(zCodaMorph, true);
				if (zSiotMorph >= 0 && (splitSaisiot || mergeSaisiot)) shortcut(zSiotMorph, false);
				if (cnt) cnt->candEvals += validMorphCands.size();
				if (wc) wc->candEvals += validMorphCands.size();
				if (totalPrevPathes <= 128) cont.mode = 0;
				else if (totalPrevPathes <= 512) { cont.mode = 1; if (cnt) cnt->mediumMode++; }
				else { cont.mode = 2; if (cnt) cnt->top1Mode++; }
				evalCong(nCache, nodeIdx, ownFormId, validMorphCands, ignoreCond ? -10.f : 0.f, nodeLevelDiscount);
				if (!nCache.empty()) break;
			}
		}

		// PathEvaluator::operator(), PathEvaluator.hpp:347-512
		void evaluate(size_t nodeIdx, size_t ownFormId, const uint32_t* cands, size_t nCands, float unkFormDiscount)
		{
			const LNode* node = graph + nodeIdx;
			auto& nCache = cache[nodeIdx];
			float whitespaceDiscount = 0;
			if (node->uformLen == 0 && node->form >= 0 && im.formLen(node->form) && node->spaceErrors) whitespaceDiscount = -cfg.space_penalty * node->spaceErrors;
			const float typoDiscount = -node->typoCost * cfg.typo_cost_weight;
			const float nodeLevelDiscount = whitespaceDiscount + typoDiscount + unkFormDiscount;
			size_t totalPrevPathes = 0;
			for (const LNode* prev = node->prev ? node - node->prev : nullptr; prev; prev = prev->sibling ? prev + prev->sibling : nullptr) totalPrevPathes += cache[prev - graph].size();

			uint32_t esFlags = 0; const uint64_t esLm0 = cnt ? cnt->lmSteps : 0;
			if (es)
			{
				for (const LNode* prev = node->prev ? node - node->prev : nullptr; prev; prev = prev->sibling ? prev + prev->sibling : nullptr)
					for (auto& pp : cache[prev - graph]) if (pp.combineSocket) esFlags |= 2;
				for (size_t ci = 0; ci < nCands; ++ci)
				{
					const auto& cur = M((int32_t)cands[ci]);
					if (cur.tag == T_z_coda || cur.tag == T_z_siot) esFlags |= 4;
					if (cur.combine_socket) esFlags |= isSingle(cur) ? 16 : 32;
					const RuleScorer r = makeRuleScorer((int32_t)cands[ci], node);
					if (r.sbType || r.specialType == 0 || r.specialType == 1 || r.specialType == 3 || r.specialType == 4) esFlags |= 1;
				}
			}
			if (cong) evaluateCong(nodeIdx, ownFormId, cands, nCands, nodeLevelDiscount, totalPrevPathes);
			else
			for (int ignoreCond = 0; ignoreCond < 2; ++ignoreCond)
			{
				if (ignoreCond) esFlags |= 8;
				for (size_t ci = 0; ci < nCands; ++ci)
				{
					const int32_t curId = (int32_t)cands[ci];
					const auto& cur = M(curId);
					if (splitComplex && hasComplex(curId)) continue;
					if (blocked(curId)) continue;
					if (cur.dialect != 0) continue;          // allowedDialect == standard
					if (cur.tag == T_z_coda || cur.tag == T_z_siot)
					{
						if (cur.tag == T_z_siot && !(splitSaisiot || mergeSaisiot)) continue;
						for (const LNode* prev = node->prev ? node - node->prev : nullptr; prev; prev = prev->sibling ? prev + prev->sibling : nullptr)
						{
							const int32_t pNode = (int32_t)(prev - graph);
							for (size_t pi = 0; pi < cache[pNode].size(); ++pi)
							{
								const WordLL& p = cache[pNode][pi];
								const uint8_t lastTag = M((int32_t)p.wid).tag;
								if (cur.tag == T_z_coda) { if (!isJClass(lastTag) && !isEClass(lastTag)) continue; }
								else { if (!isNNClass(lastTag)) continue; }
								WordLL np = p;
								np.accScore += cur.user_score * cfg.typo_cost_weight;
								np.accTypoCost -= cur.user_score;
								np.parentNode = pNode; np.parentIdx = (int32_t)pi;
								np.morpheme = (int32_t)cur.lm_morpheme_id;
								np.wid = cur.lm_morpheme_id;
								nCache.push_back(np);
							}
						}
						continue;
					}
					if (!isSingle(cur))
					{
						const u16* kf = kformPtr(cur);
						const auto& c0 = M((int32_t)im.chunks[cur.chunk_off].morph);
						if (node->prev && (node - node->prev)->endPos < node->startPos
							&& kf && kformLen(cur) == 1 && (kf[0] == 0xB2E4 || kf[0] == 0xAC8C || kf[0] == 0xC9C0)
							&& kformPtr(c0) && kformLen(c0) == 1 && kformPtr(c0)[0] == 0xD558)
						{
							continue;
						}
					}
					if (cnt) cnt->candEvals++;
					if (wc) wc->candEvals++;
					if (totalPrevPathes <= 128) cont.mode = 0;
					else if (totalPrevPathes <= 512) { cont.mode = 1; if (cnt) cnt->mediumMode++; }
					else { cont.mode = 2; if (cnt) cnt->top1Mode++; }
					evalSingleMorpheme(nCache, nodeIdx, ownFormId, curId, ignoreCond ? -10.f : 0.f, nodeLevelDiscount);
				}
				if (!nCache.empty()) break;
			}

			if (es) { es->rows.push_back((uint32_t)totalPrevPathes); es->rows.push_back((uint32_t)nCands); es->rows.push_back((uint32_t)((cnt ? cnt->lmSteps : 0) - esLm0)); es->rows.push_back(esFlags); es->rows.push_back((uint32_t)nCache.size()); }
			if (cnt) { cnt->maxNodePre = std::max<uint64_t>(cnt->maxNodePre, nCache.size()); cnt->maxIncoming = std::max<uint64_t>(cnt->maxIncoming, totalPrevPathes); cnt->evalCalls++; }
			std::vector<float> maxScores(1 + uniqStates.size(), -INFINITY);
			for (auto& c : nCache)
			{
				if (M(c.morpheme).combine_socket) continue;
				const size_t rootId = c.rootId == commonRootId ? 0 : c.rootId + 1;
				maxScores[rootId] = std::max(maxScores[rootId], c.accScore);
			}
			size_t validCount = 0;
			for (size_t i = 0; i < nCache.size(); ++i)
			{
				const size_t rootId = nCache[i].rootId == commonRootId ? 0 : nCache[i].rootId + 1;
				if (nCache[i].accScore + cfg.cut_off_threshold < maxScores[rootId]) continue;
				if (validCount != i) nCache[validCount] = nCache[i];
				validCount++;
			}
			nCache.resize(validCount);
		}

		// src/UnkFormScorer.cpp:28-51
		float unkFormScore(const u16* form, size_t len) const
		{
			float penalty = 0;
			if (len > 0)
			{
				uint32_t chrs[2] = { 0, 0 };
				for (size_t i = 0, j = 0; i < len && j < 2; ++j)
				{
					if (isHighSurrogate(form[i])) { chrs[j] = mergeSurrogate(form[i], i + 1 < len ? form[i + 1] : 0); i += 2; }
					else { chrs[j] = form[i]; ++i; }
				}
				if (im.isEmoji(chrs[0], chrs[1])) penalty = -10;
			}
			return penalty - (len * cfg.oov_rule_scale + cfg.oov_rule_bias);
		}

		// PathEvaluator.hpp:1159-1176
		bool isDisconnected(std::vector<uint8_t>& reachable, size_t scanStart) const
		{
			if (reachable[scanStart - 1]) return false;
			std::fill(reachable.begin() + scanStart, reachable.end(), 0);
			for (size_t i = scanStart; i < graphSize; ++i)
			{
				for (const LNode* prev = graph[i].prev ? &graph[i] - graph[i].prev : nullptr; prev; prev = prev->sibling ? prev + prev->sibling : nullptr)
				{
					if (reachable[prev - graph]) { reachable[i] = 1; break; }
				}
			}
			return reachable[graphSize - 1] == 0;
		}

		uint32_t unify(int32_t morph) const            // PathEvaluator.hpp:1054-1058
		{
			if (!((uint32_t)morph < im.h->lang_vocab_size) || M(morph).combined) return (uint32_t)morph;
			return M(morph).lm_morpheme_id;
		}

		// PathEvaluator.hpp:1038-1157
		std::vector<PathTok> generateTokenList(const WordLL& result) const
		{
			std::vector<std::pair<const WordLL*, int32_t>> steps;   // (path, node index)
			{
				int32_t n = result.parentNode, i = result.parentIdx;
				while (true)
				{
					const WordLL* s = &cache[n][i];
					if (s->parentNode < 0) break;
					steps.emplace_back(s, n);
					n = s->parentNode; i = s->parentIdx;
				}
			}
			std::vector<PathTok> ret;
			const WordLL* prev = &cache[steps.back().first->parentNode][steps.back().first->parentIdx];
			for (auto it = steps.rbegin(); it != steps.rend(); ++it)
			{
				const WordLL* cur = it->first;
				const float scoreDiff = cur->accScore - prev->accScore;
				float typoCostDiff = cur->accTypoCost - prev->accTypoCost;
				const auto& m = M(cur->morpheme);
				const bool single = m.chunk_cnt == 0 || (m.flags & (KB2_MORPH_COMPLEX | KB2_MORPH_SAISIOT));
				const size_t numNewTokens = (splitSaisiot && (m.flags & KB2_MORPH_SAISIOT)) || !single ? m.chunk_cnt : 1;
				const LNode& g = graph[it->second];
				const float firstScore = cur->firstChunkScore + typoCostDiff * cfg.typo_cost_weight;
				const float restScores = numNewTokens > 1 ? (scoreDiff - cur->firstChunkScore) / (numNewTokens - 1) : 0;
				const uint32_t nodeId = (uint32_t)it->second;
				auto chunkTok = [&](size_t ch, float sc)
				{
					const auto& c = im.chunks[m.chunk_off + ch];
					ret.push_back(PathTok{ unify((int32_t)c.morph), g.startPos + c.begin, g.startPos + c.end, sc, nodeId, -1, 0 });
				};
				if (splitSaisiot && (m.flags & KB2_MORPH_SAISIOT))
				{
					for (size_t ch = 0; ch < numNewTokens; ++ch) chunkTok(ch, ch == 0 ? firstScore : restScores);
					ret.back().end = g.endPos;
				}
				else if (single)
				{
					PathTok t{ unify(cur->morpheme), g.startPos, g.endPos, firstScore, nodeId, -1, 0 };
					if (cur->ownFormId) { const auto& o = ownFormList[cur->ownFormId - 1]; t.strOff = o.first; t.strLen = o.second; }
					ret.push_back(t);
				}
				else if (m.combine_socket)
				{
					ret.back().morph = ret.back().morph + M((int32_t)ret.back().morph).combined;
					ret.back().end = g.startPos + im.chunks[m.chunk_off].end;
					ret.back().wordScore = firstScore;
					for (size_t ch = 1; ch < numNewTokens; ++ch) chunkTok(ch, restScores);
					ret.back().end = g.endPos;
				}
				else
				{
					for (size_t ch = 0; ch < numNewTokens; ++ch) chunkTok(ch, ch == 0 ? firstScore : restScores);
					ret.back().end = g.endPos;
				}
				prev = cur;
			}
			return ret;
		}

		// PathEvaluator.hpp:1178-1419
		std::vector<PathResult> findBestPath(const std::vector<uint8_t>& prevSpStates, const u16* normForm, const LNode* g, size_t gs,
			bool openEnding, uint32_t matchOptions)
		{
			graph = g; graphSize = gs; norm = normForm;
			splitComplex = (matchOptions >> 22) & 1; splitSaisiot = (matchOptions >> 25) & 1; mergeSaisiot = (matchOptions >> 26) & 1;
			cache.assign(gs, {});
			ownFormList.clear();
			std::vector<uint8_t> reachable(gs, 0);
			const uint32_t unknownNodeCands[2] = { T_nng + 1u, T_nnp + 1u };     // getDefaultMorphemeId, Kiwi.h:64-67
			const uint32_t unknownNodeLCands[1] = { T_nnp + 1u };
			uniqStates = prevSpStates;
			std::sort(uniqStates.begin(), uniqStates.end());
			uniqStates.erase(std::unique(uniqStates.begin(), uniqStates.end()), uniqStates.end());
			if (prevSpStates.empty()) uniqStates.push_back(0);

			{
				WordLL bos;
				bos.morpheme = 0; bos.lmState = cong ? 0 : im.h->kn_bos_node; bos.ctxIdx = 0; bos.rootId = commonRootId;      // CoNgramState(const ILangModel*): node 0, contextIdx 0
				cache[0].push_back(bos);
				reachable[0] = 1;
			}
			for (size_t i = 1; i + 1 < gs; ++i)
			{
				const LNode* node = &graph[i];
				size_t ownFormId = 0;
				if (node->uformLen) { ownFormList.emplace_back(node->uformOff, node->uformLen); ownFormId = ownFormList.size(); }
				if (node->form >= 0)
				{
					const auto& f = im.forms[node->form];
					if (wc) wc->candEntries += f.cand_cnt;
					evaluate(i, ownFormId, im.formCands + f.cand_off, f.cand_cnt, 0.f);
					bool allPartial = true;
					for (uint32_t c = 0; c < f.cand_cnt; ++c)
					{
						const auto& m = M((int32_t)im.formCands[f.cand_off + c]);
						if (!(m.combine_socket || !(m.chunk_cnt == 0 || (m.flags & (KB2_MORPH_COMPLEX | KB2_MORPH_SAISIOT))))) { allPartial = false; break; }
					}
					if (node->typoCost == 0 && node->typoFormId == 0 && allPartial)
					{
						ownFormList.emplace_back(~node->form, f.str_len);
						ownFormId = ownFormList.size();
						const float unkScore = unkFormScore(im.formStr(node->form), f.str_len);
						evaluate(i, ownFormId, unknownNodeLCands, 1, unkScore);
					}
					reachable[i] = 0;
					for (auto& p : cache[i]) if (!p.combineSocket) { reachable[i] = 1; break; }
					if (isDisconnected(reachable, i + 1))
					{
						ownFormList.emplace_back((int32_t)node->startPos, node->endPos - node->startPos);
						ownFormId = ownFormList.size();
						const float unkScore = unkFormScore(norm + node->startPos, node->endPos - node->startPos);
						evaluate(i, ownFormId, unknownNodeCands, 2, unkScore);
					}
				}
				else
				{
					const float unkScore = unkFormScore(norm + node->uformOff, node->uformLen);
					evaluate(i, ownFormId, unknownNodeCands, 2, unkScore);
				}
				if (cnt) cnt->pathsOut += cache[i].size();
				if (wc) wc->pathsKept += cache[i].size();
				if (std::getenv("ORC_TRACE_NODES")) std::fprintf(stderr, " %zu", cache[i].size());
				if (const char* dn = std::getenv("ORC_DUMP_NODE")) if ((size_t)std::atoi(dn) == i) for (auto& q : cache[i]) std::fprintf(stderr, "[path] morph %d lm %d acc %a root %u sp %u hist %u %u %u %u %u %u %u %u pos %u\n", q.morpheme, q.lmState, q.accScore, q.rootId, q.spState, q.sb.hist[0], q.sb.hist[1], q.sb.hist[2], q.sb.hist[3], q.sb.hist[4], q.sb.hist[5], q.sb.hist[6], q.sb.hist[7], (unsigned)q.sb.pos);
			}
			if (std::getenv("ORC_TRACE_NODES")) std::fprintf(stderr, " <- [orc] paths per node (1..)\n");

			// end node
			auto& cand = cache.back();
			const LNode* endNode = graph + gs - 1;
			for (const LNode* prev = endNode->prev ? endNode - endNode->prev : nullptr; prev; prev = prev->sibling ? prev + prev->sibling : nullptr)
			{
				const int32_t pNode = (int32_t)(prev - graph);
				for (size_t pi = 0; pi < cache[pNode].size(); ++pi)
				{
					const WordLL& p = cache[pNode][pi];
					if (p.combineSocket) continue;
					const auto& pm = M(p.morpheme);
					if (!(pm.chunk_cnt == 0 || (pm.flags & (KB2_MORPH_COMPLEX | KB2_MORPH_SAISIOT))))
					{
						if (pm.chunk_cnt <= (pm.combine_socket ? 2u : 1u))
						{
							if (!(pm.vowel == CV_none)) continue;      // FeatureTestor::isMatched(nullptr, vowel)
						}
					}
					if (pm.tag == T_z_siot) continue;
					float c = p.accScore;
					float firstChunkScore = 0;
					int32_t st = p.lmState; uint32_t stCtx = p.ctxIdx;
					if (!openEnding)
					{
						SbHist stSb = p.sb;
						c += (firstChunkScore = cong ? cg.next(st, stCtx, 1) : lmNext(st, stSb, 1));
						if (p.spState & 1) c -= 2;
						if (p.spState & 2) c -= 2;
					}
					WordLL w;
					w.morpheme = -1; w.accScore = c; w.firstChunkScore = firstChunkScore; w.accTypoCost = p.accTypoCost; w.accDialectCost = p.accDialectCost;
					w.parentNode = pNode; w.parentIdx = (int32_t)pi; w.lmState = p.lmState; w.rootId = p.rootId;
					if (p.rootId == commonRootId)
					{
						for (size_t i = 0; i < uniqStates.size(); ++i) { w.spState = uniqStates[i]; w.rootId = (uint8_t)i; cand.push_back(w); }
					}
					else { w.spState = p.spState; cand.push_back(w); }
				}
			}
			if (std::getenv("ORC_TRACE_END")) { std::fprintf(stderr, "[orc] end cands %zu:", cand.size()); for (auto& c : cand) std::fprintf(stderr, " (%d,%d,%a,n%d,i%d)", c.rootId, c.spState, c.accScore, c.parentNode, c.parentIdx); std::fprintf(stderr, "\n"); }
			std::sort(cand.begin(), cand.end(), [](const WordLL& a, const WordLL& b)
			{
				if (a.rootId < b.rootId) return true;
				if (a.rootId > b.rootId) return false;
				if (a.spState < b.spState) return true;
				if (a.spState > b.spState) return false;
				return a.accScore > b.accScore;
			});
			std::vector<PathResult> ret;
			size_t numUniq = 0;
			{
				std::vector<std::pair<uint8_t, uint8_t>> u;
				for (auto& c : cand) u.emplace_back(c.rootId, c.spState);
				std::sort(u.begin(), u.end());
				numUniq = std::unique(u.begin(), u.end()) - u.begin();
			}
			const size_t perGroup = (size_t)std::ceil(2 / (double)numUniq);
			size_t startIdx = 0;
			std::pair<uint8_t, uint8_t> prevKey{ 0, 0 };
			if (!cand.empty()) prevKey = { cand[0].rootId, cand[0].spState };
			for (size_t i = 0; i < cand.size(); ++i)
			{
				std::pair<uint8_t, uint8_t> curKey{ cand[i].rootId, cand[i].spState };
				if (prevKey != curKey) { startIdx = i; prevKey = curKey; }
				if (i - startIdx < perGroup)
				{
					PathResult r;
					r.path = generateTokenList(cand[i]);
					r.score = cand[i].accScore; r.prevState = uniqStates[cand[i].rootId]; r.curState = cand[i].spState;
					ret.push_back(std::move(r));
				}
			}
			std::sort(ret.begin(), ret.end(), [](const PathResult& a, const PathResult& b) { return a.score > b.score; });
			return ret;
		}
	};
}
