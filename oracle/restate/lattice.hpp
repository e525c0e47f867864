// ORACLE (test infrastructure): CPU restatement of the reference's default ("new") Splitter, no typo
// transformer, no pretokenized spans:  /root/reference/src/KTrie.cpp:709-1521 (+ appendNewNode 16-43,
// removeUnconnected 240-299, countSpaceErrors 316-328, isDiscontinuous 566-574).
// Works on the flat model image; node order, prev/sibling offsets and positions follow the reference exactly.
#pragma once
#include "prep.hpp"
#include "typo.hpp"

namespace orc
{
	struct LNode      // KGraphNode, src/KTrie.h:57-77
	{
		int32_t form = -1;          // form index, -1 = none
		int32_t uformOff = -1;      // absolute offset into the normalized sentence, -1 = empty uform
		uint32_t uformLen = 0;
		uint32_t prev = 0, sibling = 0;
		uint32_t startPos = 0, endPos = 0;
		float typoCost = 0;
		uint32_t typoFormId = 0;
		uint32_t spaceErrors = 0;
	};

	struct Splitter
	{
		const Image& im;
		Pat pat;
		// options
		uint32_t matchOptions = 0;
		size_t maxUnkFormSize = 6, maxUnkFormSizeFollowedByJClass = (uint32_t)-1, spaceTolerance = 0;
		// state
		std::vector<std::pair<uint32_t, uint32_t>> endPosMap;
		struct MP { size_t end; uint32_t len; uint8_t tag; };
		std::vector<MP> matchedPatterns;
		size_t nextMatchedPattern = 0;
		std::vector<uint32_t> nsToPos, posToNs;
		std::vector<LNode> out;
		const u16* rawStr = nullptr; size_t rawLen = 0; size_t startOffset = 0;
		std::vector<int32_t> candidates;

		WorkCounters* wc = nullptr;
		// typo lattice (AnalyzeOption::typoTransformer / typoThreshold): a flat typo image, or nullptr for the plain 2-node graph
		const TypoImage* typoImg = nullptr;
		float typoThreshold = 2.5f;
		explicit Splitter(const Image& _im) : im{ _im }, pat{ _im } {}

		static constexpr uint32_t MATCH_ZCODA = 1u << 23, MATCH_SPLIT_SAISIOT = 1u << 25, MATCH_MERGE_SAISIOT = 1u << 26;

		// KTrie.cpp:16-43
		bool appendNewNode(size_t startPos, size_t endPos, int32_t form, int32_t uoff, uint32_t ulen, float typoCost = 0)
		{
			if (endPosMap[startPos].first == endPosMap[startPos].second) return false;
			const size_t newId = out.size();
			LNode n;
			n.form = form; n.uformOff = ulen ? uoff : -1; n.uformLen = ulen;
			n.startPos = (uint16_t)startPos; n.endPos = (uint16_t)endPos;     // ctor takes uint16_t (KTrie.h:67-70)
			n.typoCost = typoCost;
			n.prev = (uint32_t)(newId - endPosMap[startPos].first);
			out.push_back(n);
			if (n.endPos >= endPosMap.size()) return true;
			auto& e = endPosMap[n.endPos];
			if (e.first == e.second) { e.first = (uint32_t)newId; e.second = (uint32_t)newId + 1; }
			else
			{
				out[e.second - 1].sibling = (uint32_t)(newId - (e.second - 1));
				e.second = (uint32_t)newId + 1;
			}
			return true;
		}

		// KTrie.cpp:766-858; returns the stop position (relative to `str`)
		size_t preparePattern(const u16* str, size_t len)
		{
			size_t n = 0, continuousNonSpaceCount = 0;
			uint8_t lastChrType = T_unknown;
			for (; n < len; ++n)
			{
				{
					auto m = pat.match(n ? str[n - 1] : (u16)' ', str + n, str + len, matchOptions);
					if (m.second != T_unknown)
					{
						matchedPatterns.push_back(MP{ n + m.first, (uint32_t)m.first, m.second });
						n += m.first - 1;
						continue;
					}
				}
				const u16 c = str[n];
				uint32_t c32 = c;
				if (isHighSurrogate(c32) && n + 1 < len) c32 = mergeSurrogate(c32, str[n + 1]);
				const uint8_t chrType = im.cls(c32);
				if (chrType == T_unknown) continuousNonSpaceCount = 0;
				else continuousNonSpaceCount++;
				if (chrType == T_unknown && n >= (lastChrType == T_sf ? 4u : 4096u))
				{
					if (!im.isSpace(str[n - 3]) && !im.isSpace(str[n - 2])) break;
				}
				else if (continuousNonSpaceCount >= 1024) break;
				if (c32 >= 0x10000) ++n;
				lastChrType = chrType;
			}
			for (size_t i = 0; i < n; ++i)
			{
				if (!im.isSpace(str[i]))
				{
					posToNs.push_back((uint32_t)nsToPos.size());
					nsToPos.push_back((uint32_t)i);
					if (isHighSurrogate(str[i]) && i + 1 < n)
					{
						posToNs.push_back((uint32_t)nsToPos.size());
						nsToPos.push_back((uint32_t)++i);
					}
				}
				else posToNs.push_back((uint32_t)nsToPos.size());
			}
			posToNs.push_back((uint32_t)nsToPos.size());
			std::sort(matchedPatterns.begin(), matchedPatterns.end(), [](const MP& a, const MP& b)
			{
				if (a.end != b.end) return a.end < b.end;
				if (a.len != b.len) return a.len < b.len;
				return a.tag < b.tag;
			});
			nextMatchedPattern = 0;
			return n;
		}

		size_t formSizeWithoutSpace(int32_t f) const { return im.forms[f].str_len - im.forms[f].num_spaces; }

		// KTrie.cpp:897-905
		bool hasFormAlready(size_t startPos, size_t endPos) const
		{
			const uint32_t scanStart = std::max(endPosMap[endPos].first, (uint32_t)1), scanEnd = endPosMap[endPos].second;
			if (endPosMap[endPos].first == (uint32_t)-1) return false;      // [max(-1,1), -1) is an empty range
			for (uint32_t i = scanStart; i < scanEnd; ++i)
			{
				const auto& g = out[i];
				const size_t sp = g.endPos - (g.uformLen == 0 ? formSizeWithoutSpace(g.form) : g.uformLen);
				if (g.endPos == endPos && sp == startPos && g.typoCost == 0 && (g.form < 0 || (im.forms[g.form].flags & KB2_FORM_HASFULL))) return true;
			}
			return false;
		}

		// KTrie.cpp:907-919
		std::pair<bool, bool> isZFollowable(size_t pos) const
		{
			if (pos >= nsToPos.size()) return { false, false };
			const uint32_t scanStart = endPosMap[pos].first, scanEnd = endPosMap[pos].second;
			bool zc = false, zs = false;
			if (scanStart == (uint32_t)-1) return { false, false };
			for (uint32_t i = scanStart; i < scanEnd; ++i)
			{
				const auto& g = out[i];
				if (g.endPos != pos || g.form < 0) continue;
				zc = zc || (im.forms[g.form].flags & KB2_FORM_ZCODA);
				zs = zs || (im.forms[g.form].flags & KB2_FORM_ZSIOT);
			}
			return { zc, zs };
		}

		void appendRaw(size_t sNs, size_t eNs)   // helper: unknown-form node over rawStr[nsToPos[s] .. nsToPos[e-1]]
		{
			size_t off = nsToPos[sNs], len = nsToPos[eNs - 1] + 1 - nsToPos[sNs];
			while (len && im.isSpace(rawStr[off + len - 1])) --len;
			appendNewNode(sNs, eNs, -1, (int32_t)(startOffset + off), (uint32_t)len);
		}

		// KTrie.cpp:921-953
		void insertUnkForm(size_t startPos, size_t endPos, bool hasJClass)
		{
			if (startPos >= endPos || hasFormAlready(startPos, endPos)) return;
			size_t lastPos = out.back().endPos;
			if (lastPos < endPos)
			{
				if (lastPos && isHangulCoda(rawStr[nsToPos[lastPos]])) lastPos--;
				if (lastPos != startPos && !hasFormAlready(lastPos, endPos)) appendRaw(lastPos, endPos);
			}
			const size_t newNodeLength = endPos - startPos;
			const size_t lengthLimit = hasJClass ? maxUnkFormSizeFollowedByJClass : maxUnkFormSize;
			if (newNodeLength <= lengthLimit) appendRaw(startPos, endPos);
		}

		// KTrie.cpp:316-328
		size_t countSpaceErrors(int32_t form, const uint32_t* first, const uint32_t* last) const
		{
			const u16* f = im.formStr(form);
			size_t n = 0, spaceOffset = 0;
			const size_t size = last - first;
			for (size_t i = 1; i < size; ++i)
			{
				const bool hasSpace = first[i] - first[i - 1] > 1;
				if (hasSpace && f[i + spaceOffset] != ' ') ++n;
				spaceOffset += f[i + spaceOffset] == ' ' ? 1 : 0;
			}
			return n;
		}

		// KTrie.cpp:955-996 (no continual / lengthening typos)
		void flushCandidates(size_t endPosition, ptrdiff_t startPosOffset, size_t unkFormStartNsPos, size_t lastSpaceBoundaryNsPos, float typoCost)
		{
			for (const int32_t cand : candidates)
			{
				if (wc) wc->candForms++;
				const size_t nBegin = endPosition - formSizeWithoutSpace(cand) + startPosOffset;
				const size_t nEnd = endPosition;
				const u16* fs = im.formStr(cand);
				if (!isHangulCoda(fs[0]))
				{
					const bool isSTag = im.formLen(cand) == 1 && im.cls(fs[0]) >= T_sf && im.cls(fs[0]) <= T_sw;
					const bool hj = (im.forms[cand].flags & KB2_FORM_HASJ) || isSTag;
					if (lastSpaceBoundaryNsPos < nBegin) insertUnkForm(lastSpaceBoundaryNsPos, nBegin, hj);
					insertUnkForm(unkFormStartNsPos, nBegin, hj);
				}
				size_t spaceErrors = 0;
				if ((spaceErrors = countSpaceErrors(cand, &nsToPos[nBegin], &nsToPos[nEnd])) <= spaceTolerance)
				{
					if (appendNewNode(nBegin, nEnd, cand, -1, 0, typoCost)) out.back().spaceErrors = (uint32_t)spaceErrors;
				}
			}
			candidates.clear();
		}

		// KTrie.cpp:566-574
		static bool isDiscontinuous(uint8_t prevTag, uint8_t curTag, uint8_t prevScript, uint8_t curScript)
		{
			if ((prevTag == T_sl || prevTag == T_sh || prevTag == T_sw) && (curTag == T_sl || curTag == T_sh || curTag == T_sw)) return prevScript != curScript;
			return prevTag != curTag;
		}

		// trie helpers, src/FrozenTrie.hpp:14-29,55-58 with sorted keys (src/search.cpp:238-292 `balanced`)
		int32_t nextOpt(int32_t node, u16 c) const
		{
			const auto& n = im.trieNodes[node];
			if (wc) { wc->trieVisits++; wc->trieProbes += ceilLog2p1(n.num_nexts); }
			const u16* keys = im.trieKeys + n.next_offset;
			const u16* it = std::lower_bound(keys, keys + n.num_nexts, c);
			if (it == keys + n.num_nexts || *it != c) return -1;
			if (wc) wc->trieHits++;
			return node + im.trieDiffs[n.next_offset + (it - keys)];
		}
		int32_t failOf(int32_t node) const { return im.trieNodes[node].fail ? node + im.trieNodes[node].fail : -1; }

		void specialRunNode(size_t specialStartNsPos, size_t rawEnd, size_t endNs, uint8_t lastChrType)
		{
			size_t off = nsToPos[specialStartNsPos], len = rawEnd - nsToPos[specialStartNsPos];
			while (len && im.isSpace(rawStr[off + len - 1])) --len;
			if (appendNewNode(specialStartNsPos, endNs, -1, (int32_t)(startOffset + off), (uint32_t)len))
			{
				out.back().form = im.trieNodes[lastChrType].value;      // trie.value((size_t)tag): node #tag holds the default form
			}
		}

		// KTrie.cpp:998-1412 specialised to the 2-node typo graph [empty, whole chunk] (873-895)
		void search()
		{
			const size_t formSize = rawLen;
			uint32_t prevChr = 0;
			uint8_t lastChrType = T_unknown, lastScriptType = 0;
			size_t specialStartNsPos = 0, unkFormStartNsPos = 0, lastSpaceBoundaryNsPos = 0;
			int32_t curNode = 0;
			const float typoCost = 0;
			for (size_t j = 0; j < formSize; ++j)
			{
				const u16 c = rawStr[j];
				uint32_t c32 = c;
				if (isHighSurrogate(c32) && j + 1 < formSize) c32 = mergeSurrogate(c32, rawStr[j + 1]);
				{
					const bool isInPattern = nextMatchedPattern != matchedPatterns.size() &&
						j >= matchedPatterns[nextMatchedPattern].end - matchedPatterns[nextMatchedPattern].len;
					uint8_t chrType = im.cls(c32);
					uint8_t scriptType = im.script(c32);
					if (lastChrType == T_sw && (c32 == 0x200d || (0x1f3fb <= c32 && c32 <= 0x1f3ff) || scriptType == im.h->script_variation_selectors))
					{
						chrType = lastChrType;
						scriptType = lastScriptType;
					}
					if (isDiscontinuous(lastChrType, isInPattern ? (uint8_t)T_unknown : chrType, lastScriptType, scriptType)
						|| lastChrType == T_sso || lastChrType == T_ssc)
					{
						if (lastChrType != T_max && lastChrType != T_unknown)
						{
							if (lastChrType != T_ss)
							{
								const bool hj = T_sf <= lastChrType && lastChrType <= T_sw;
								if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, specialStartNsPos, hj);
								insertUnkForm(unkFormStartNsPos, specialStartNsPos, hj);
								specialRunNode(specialStartNsPos, j, posToNs[j], lastChrType);
							}
						}
						unkFormStartNsPos = specialStartNsPos;
						specialStartNsPos = posToNs[j];
						if (T_sf <= lastChrType && lastChrType <= T_sw) lastSpaceBoundaryNsPos = specialStartNsPos;
					}
					else if (chrType == T_max)
					{
						unkFormStartNsPos = specialStartNsPos;
					}
					lastChrType = isInPattern ? (uint8_t)T_unknown : chrType;
					lastScriptType = scriptType;
					if (c32 >= 0x10000) {}
					else
					{
						if (chrType == T_unknown)   // whitespace
						{
							if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, posToNs[j + 1], true);
							insertUnkForm(unkFormStartNsPos, posToNs[j + 1], true);
							lastSpaceBoundaryNsPos = specialStartNsPos = unkFormStartNsPos = posToNs[j + 1];
							prevChr = c32;
							continue;
						}
						const size_t pos = j;
						const auto zf = isZFollowable(posToNs[pos]);
						if ((matchOptions & MATCH_ZCODA) && zf.first && isHangulCoda(c) && (pos + 1 >= rawLen || !isHangulSyllable(rawStr[pos + 1])))
						{
							candidates.push_back((int32_t)(im.h->default_tag_size + (c - 0x11A8) - 1));
						}
						else if ((matchOptions & (MATCH_SPLIT_SAISIOT | MATCH_MERGE_SAISIOT)) && zf.second && c == 0x11BA && pos + 1 < rawLen && isHangulSyllable(rawStr[pos + 1]))
						{
							candidates.push_back((int32_t)(im.h->default_tag_size + (0x11BA - 0x11A8) - 1));
						}
					}
				}
				if (nextMatchedPattern != matchedPatterns.size())
				{
					const size_t currentEnd = j + (c32 >= 0x10000 ? 2 : 1);
					while (nextMatchedPattern != matchedPatterns.size() && matchedPatterns[nextMatchedPattern].end == currentEnd)
					{
						const auto mp = matchedPatterns[nextMatchedPattern];
						const size_t matchedStart = mp.end - mp.len;
						const bool hj = T_w_url <= mp.tag && mp.tag <= T_w_emoji;
						if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, posToNs[matchedStart], hj);
						insertUnkForm(unkFormStartNsPos, posToNs[matchedStart], hj);
						if (appendNewNode(posToNs[matchedStart], posToNs[mp.end], -1, (int32_t)(startOffset + matchedStart), (uint32_t)(mp.end - matchedStart)))
						{
							out.back().form = im.trieNodes[mp.tag].value;
						}
						++nextMatchedPattern;
					}
				}
				if (c32 >= 0x10000)
				{
					++j;
					prevChr = c32;
					continue;
				}
				prevChr = c32;

				int32_t nextNode = nextOpt(curNode, c);
				while (nextNode < 0)
				{
					curNode = failOf(curNode);
					if (curNode < 0) break;
					nextNode = nextOpt(curNode, c);
				}
				if (nextNode >= 0)
				{
					curNode = nextNode;
					for (int32_t sub = curNode; sub >= 0; sub = failOf(sub))
					{
						const int32_t v = im.trieNodes[sub].value;
						if (wc && sub != curNode) wc->trieVisits++;
						if (v == KB2_TRIE_NONE) break;
						else if (v != KB2_TRIE_SUBMATCH) candidates.push_back(v);    // minFormLen == 0 without typos
					}
				}
				else curNode = 0;

				flushCandidates(posToNs[j + 1], 0, unkFormStartNsPos, lastSpaceBoundaryNsPos, typoCost);
			}
			if (lastChrType != T_max && lastChrType != T_unknown)
			{
				if (lastChrType != T_ss)
				{
					const bool hj = T_sf <= lastChrType && lastChrType <= T_sw;
					if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, specialStartNsPos, hj);
					insertUnkForm(unkFormStartNsPos, specialStartNsPos, hj);
					specialRunNode(specialStartNsPos, rawLen, posToNs[rawLen], lastChrType);
					unkFormStartNsPos = specialStartNsPos;
					if (hj) lastSpaceBoundaryNsPos = posToNs[rawLen];
				}
			}
			// search() tail, KTrie.cpp:1434-1452 (the single end state)
			const size_t totEndPos = nsToPos.back() + 1;
			if (rawLen == totEndPos)
			{
				if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, posToNs[totEndPos], true);
				insertUnkForm(unkFormStartNsPos, posToNs[totEndPos], true);
			}
			appendNewNode(nsToPos.size(), nsToPos.size() + 1, -1, -1, 0);
			out.back().endPos = (uint32_t)nsToPos.size();
		}


		// ---- the general search over a typo graph (KTrie.cpp:998-1452, lengtheningTypoTolerant == false, no continual typos, no
		// pretokenized spans).  One SearchState per (typo-graph node, surviving trie state); `search()` above is this loop for the
		// 2-node graph.  Differences that matter: special-character runs are flushed at the end of EVERY zero-cost graph node,
		// trie candidates of a replaced segment are only collected at its last character and must span it (`minFormLen`), and
		// a candidate's start is shifted by the length difference of the replacements it crossed (`startPosOffset`).
		struct SState
		{
			int32_t node = 0; float accumulatedCost = 0; uint32_t minFormLen = 0; int32_t startPosOffset = 0;
			uint32_t specialStartNsPos = 0, unkFormStartNsPos = 0, lastSpaceBoundaryNsPos = 0; uint32_t lastChr = 0;
		};

		void progressTypoNode(const TypoNode& prevT, const TypoNode& tn, const SState& state, std::vector<SState>& curStates)
		{
			float typoCost = state.accumulatedCost + tn.typoCost;
			if (typoCost > typoThreshold) return;
			const u16* form = tn.fromPool ? typoImg->pool + tn.off : rawStr + tn.off;
			const size_t formSize = tn.len;
			uint32_t prevChr = state.lastChr;
			uint8_t lastChrType = prevChr ? im.cls(prevChr) : (uint8_t)T_unknown;
			uint8_t lastScriptType = prevChr ? im.script(prevChr) : 0;
			size_t specialStartNsPos = state.specialStartNsPos, unkFormStartNsPos = state.unkFormStartNsPos, lastSpaceBoundaryNsPos = state.lastSpaceBoundaryNsPos;
			size_t minFormLen = state.minFormLen;
			int32_t startPosOffset = state.startPosOffset;
			if (tn.typoCost > 0) startPosOffset += (int32_t)((ptrdiff_t)formSize - (ptrdiff_t)(tn.endPos - prevT.endPos));
			int32_t curNode = state.node;
			for (size_t j = 0; j < formSize; ++j)
			{
				const u16 c = form[j];
				uint32_t c32 = c;
				if (isHighSurrogate(c32) && j + 1 < formSize) c32 = mergeSurrogate(c32, form[j + 1]);
				const size_t pos = tn.endPos + j - formSize;
				if (typoCost == 0)
				{
					const bool isInPattern = nextMatchedPattern != matchedPatterns.size() &&
						pos >= matchedPatterns[nextMatchedPattern].end - matchedPatterns[nextMatchedPattern].len;
					uint8_t chrType = im.cls(c32);
					uint8_t scriptType = im.script(c32);
					if (lastChrType == T_sw && (c32 == 0x200d || (0x1f3fb <= c32 && c32 <= 0x1f3ff) || scriptType == im.h->script_variation_selectors))
					{
						chrType = lastChrType;
						scriptType = lastScriptType;
					}
					if (isDiscontinuous(lastChrType, isInPattern ? (uint8_t)T_unknown : chrType, lastScriptType, scriptType)
						|| lastChrType == T_sso || lastChrType == T_ssc)
					{
						if (lastChrType != T_max && lastChrType != T_unknown)
						{
							if (lastChrType != T_ss)
							{
								const bool hj = T_sf <= lastChrType && lastChrType <= T_sw;
								if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, specialStartNsPos, hj);
								insertUnkForm(unkFormStartNsPos, specialStartNsPos, hj);
								specialRunNode(specialStartNsPos, pos, posToNs[pos], lastChrType);
							}
						}
						unkFormStartNsPos = specialStartNsPos;
						specialStartNsPos = posToNs[pos];
						if (T_sf <= lastChrType && lastChrType <= T_sw) lastSpaceBoundaryNsPos = specialStartNsPos;
					}
					else if (chrType == T_max)
					{
						unkFormStartNsPos = specialStartNsPos;
					}
					lastChrType = isInPattern ? (uint8_t)T_unknown : chrType;
					lastScriptType = scriptType;
					if (c32 >= 0x10000) {}
					else
					{
						if (chrType == T_unknown)   // whitespace
						{
							if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, posToNs[pos + 1], true);
							insertUnkForm(unkFormStartNsPos, posToNs[pos + 1], true);
							lastSpaceBoundaryNsPos = specialStartNsPos = unkFormStartNsPos = posToNs[pos + 1];
							prevChr = c32;
							continue;
						}
						const auto zf = isZFollowable(posToNs[pos]);
						if ((matchOptions & MATCH_ZCODA) && zf.first && isHangulCoda(c) && (pos + 1 >= rawLen || !isHangulSyllable(rawStr[pos + 1])))
						{
							candidates.push_back((int32_t)(im.h->default_tag_size + (c - 0x11A8) - 1));
						}
						else if ((matchOptions & (MATCH_SPLIT_SAISIOT | MATCH_MERGE_SAISIOT)) && zf.second && c == 0x11BA && pos + 1 < rawLen && isHangulSyllable(rawStr[pos + 1]))
						{
							candidates.push_back((int32_t)(im.h->default_tag_size + (0x11BA - 0x11A8) - 1));
						}
					}
				}
				else
				{
					if (im.isSpace((u16)c32) && c32 < 0x10000)
					{
						lastSpaceBoundaryNsPos = specialStartNsPos = unkFormStartNsPos = posToNs[pos + 1];
						prevChr = c32;
						continue;
					}
				}
				if (tn.typoCost == 0 && nextMatchedPattern != matchedPatterns.size())
				{
					const size_t currentEnd = pos + (c32 >= 0x10000 ? 2 : 1);
					while (nextMatchedPattern != matchedPatterns.size() && matchedPatterns[nextMatchedPattern].end == currentEnd)
					{
						const auto mp = matchedPatterns[nextMatchedPattern];
						const size_t matchedStart = mp.end - mp.len;
						const bool hj = T_w_url <= mp.tag && mp.tag <= T_w_emoji;
						if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, posToNs[matchedStart], hj);
						insertUnkForm(unkFormStartNsPos, posToNs[matchedStart], hj);
						if (appendNewNode(posToNs[matchedStart], posToNs[mp.end], -1, (int32_t)(startOffset + matchedStart), (uint32_t)(mp.end - matchedStart)))
						{
							out.back().form = im.trieNodes[mp.tag].value;
						}
						++nextMatchedPattern;
					}
				}
				if (c32 >= 0x10000)
				{
					++j;
					prevChr = c32;
					continue;
				}
				prevChr = c32;

				if (minFormLen > 0 || tn.typoCost > 0) ++minFormLen;
				int32_t nextNode = nextOpt(curNode, c);
				while (nextNode < 0)
				{
					curNode = failOf(curNode);
					if (curNode < 0) break;
					nextNode = nextOpt(curNode, c);
				}
				if (nextNode >= 0)
				{
					curNode = nextNode;
					// with a typo only candidates that cover the whole replaced segment are searched, at its last character
					if (tn.typoCost == 0 || j == formSize - 1)
					{
						if (typoCost > 0 && im.trieNodes[curNode].depth < minFormLen) {}      // early pruning
						else
						{
							for (int32_t sub = curNode; sub >= 0; sub = failOf(sub))
							{
								const int32_t v = im.trieNodes[sub].value;
								if (wc && sub != curNode) wc->trieVisits++;
								if (v == KB2_TRIE_NONE) break;
								else if (v != KB2_TRIE_SUBMATCH)
								{
									if (im.formLen(v) < minFormLen) break;
									candidates.push_back(v);
								}
							}
						}
					}
				}
				else
				{
					if (typoCost == 0) curNode = 0;
					else return;
				}
				flushCandidates(posToNs[pos + 1], startPosOffset, unkFormStartNsPos, lastSpaceBoundaryNsPos, typoCost);
			}
			if (typoCost == 0 && lastChrType != T_max && lastChrType != T_unknown)
			{
				if (lastChrType != T_ss)
				{
					const bool hj = T_sf <= lastChrType && lastChrType <= T_sw;
					if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, specialStartNsPos, hj);
					insertUnkForm(unkFormStartNsPos, specialStartNsPos, hj);
					specialRunNode(specialStartNsPos, tn.endPos, posToNs[tn.endPos], lastChrType);
					unkFormStartNsPos = specialStartNsPos;
					if (hj) lastSpaceBoundaryNsPos = posToNs[tn.endPos];
				}
			}
			if (typoCost > 0 && im.trieNodes[curNode].depth < minFormLen) {}      // early pruning
			else
			{
				SState ns;
				ns.node = curNode; ns.accumulatedCost = typoCost; ns.minFormLen = (uint32_t)minFormLen; ns.startPosOffset = startPosOffset;
				ns.specialStartNsPos = (uint32_t)specialStartNsPos; ns.unkFormStartNsPos = (uint32_t)unkFormStartNsPos; ns.lastSpaceBoundaryNsPos = (uint32_t)lastSpaceBoundaryNsPos;
				ns.lastChr = prevChr;
				curStates.push_back(ns);
			}
		}

		void searchTypo(const std::vector<TypoNode>& graph)
		{
			const size_t totEndPos = nsToPos.back() + 1;
			std::vector<std::vector<SState>> states(graph.size());
			states[0].emplace_back();
			for (size_t i = 1; i < graph.size(); ++i)
			{
				const TypoNode& tn = graph[i];
				auto& curStates = states[i];
				if (tn.prevOffset)
				{
					for (size_t p = i - tn.prevOffset;; p += graph[p].siblingOffset)
					{
						for (const auto& st : states[p]) progressTypoNode(graph[p], tn, st, curStates);
						if (!graph[p].siblingOffset) break;
					}
				}
				if (tn.typoCost == 0 && tn.endPos == totEndPos)
				{
					for (const auto& st : curStates)
					{
						if (st.lastSpaceBoundaryNsPos < st.unkFormStartNsPos) insertUnkForm(st.lastSpaceBoundaryNsPos, posToNs[totEndPos], true);
						insertUnkForm(st.unkFormStartNsPos, posToNs[totEndPos], true);
					}
				}
			}
			appendNewNode(nsToPos.size(), nsToPos.size() + 1, -1, -1, 0);
			out.back().endPos = (uint32_t)nsToPos.size();
		}

		// KTrie.cpp:240-299
		void removeUnconnected(std::vector<LNode>& ret) const
		{
			const size_t gs = out.size();
			std::vector<uint8_t> connected(gs, 0);
			std::vector<uint32_t> queue;
			queue.push_back((uint32_t)gs - 1);
			connected[gs - 1] = 1;
			for (size_t qi = 0; qi < queue.size(); ++qi)
			{
				const auto& node = out[queue[qi]];
				const uint32_t scanStart = endPosMap[node.startPos].first, scanEnd = endPosMap[node.startPos].second;
				if (scanStart == (uint32_t)-1) continue;
				for (uint32_t i = scanStart; i < scanEnd; ++i)
				{
					if (out[i].endPos != node.startPos) continue;
					if (connected[i]) continue;
					queue.push_back(i);
					connected[i] = 1;
				}
			}
			std::vector<size_t> sorted(gs), inverted(gs);
			for (size_t i = 0; i < gs; ++i) sorted[i] = i;
			std::stable_sort(sorted.begin(), sorted.end(), [&](size_t a, size_t b)
			{
				if (connected[a] != connected[b]) return connected[a] > connected[b];
				return out[a].endPos < out[b].endPos;
			});
			for (size_t i = 0; i < gs; ++i) inverted[sorted[i]] = i;
			size_t connectedCnt = 0;
			for (auto c : connected) connectedCnt += c;
			for (size_t i = 0; i < connectedCnt; ++i)
			{
				const size_t idx = sorted[i];
				LNode nn = out[idx];
				if (nn.prev) nn.prev = (uint32_t)(i - inverted[idx - nn.prev]);
				if (nn.sibling)
				{
					const size_t ns = inverted[idx + nn.sibling];
					if (ns >= connectedCnt) nn.sibling = 0;
					else nn.sibling = (uint32_t)(ns - i);
				}
				ret.push_back(nn);
			}
		}

		// splitByTrieUsingTypo, KTrie.cpp:1467-1521.  `str` = rest of the normalized sentence from `startOff`.
		size_t split(std::vector<LNode>& ret, const u16* str, size_t len, size_t startOff)
		{
			endPosMap.clear(); matchedPatterns.clear(); nsToPos.clear(); posToNs.clear(); out.clear(); candidates.clear();
			startOffset = startOff;
			size_t stopPos = preparePattern(str, len);
			if (nsToPos.empty())
			{
				ret.emplace_back();
				ret.emplace_back();
				while (stopPos < len && im.isSpace(str[stopPos])) ++stopPos;
				return stopPos + startOff;
			}
			rawStr = str; rawLen = stopPos;
			endPosMap.assign(nsToPos.size() + 1, std::make_pair((uint32_t)-1, (uint32_t)-1));
			endPosMap[0] = std::make_pair(0u, 1u);
			out.emplace_back();
			if (typoImg)
			{
				// buildTypoGraph (KTrie.cpp:873-895): the graph over the chunk [0, stopPos)
				TypoGraph tg{ *typoImg };
				const auto graph = tg.generate(std::u16string(reinterpret_cast<const char16_t*>(str), stopPos));
				searchTypo(graph);
			}
			else search();
			removeUnconnected(ret);
			if (wc) { wc->nodesBuilt += out.size(); wc->nodesFinal += ret.size(); }
			for (size_t i = 1; i + 1 < ret.size(); ++i)
			{
				auto& r = ret[i];
				r.startPos = nsToPos[r.startPos] + (uint32_t)startOff;
				r.endPos = nsToPos[r.endPos - 1] + 1 + (uint32_t)startOff;
			}
			ret.back().startPos = ret.back().endPos = (uint32_t)(startOff + stopPos);
			return stopPos + startOff;
		}
	};
}
