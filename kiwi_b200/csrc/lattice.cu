// kiwi_b200 kernel A: morpheme-lattice construction, one warp per sentence.
//
// Replaces (reference file:line under /root/reference/):
//   normalizeHangulWithPosition src/StrUtils.h:493-520, normalizeCoda src/StrUtils.h:637-710
//   Splitter::preparePattern src/KTrie.cpp:766-858 + matchPattern src/PatternMatcher.cpp:54-384
//   Splitter::search/progressNode src/KTrie.cpp:998-1452 (2-node typo graph, no pretokenized spans)
//   flushCandidates 955-996, insertUnkForm 921-953, hasFormAlready 897-905, isZFollowable 907-919,
//   appendNewNode 16-43, countSpaceErrors 316-328, removeUnconnected 240-299, writeResult 1454-1464
//   FrozenTrie::Node::nextOpt/fail/val src/FrozenTrie.hpp:14-29,55-58
//
// Execution model: the splitter is an order-dependent state machine, so its control flow is kept
// warp-uniform (every lane holds the same state; loads of one address are a single broadcast request) and
// the 32 lanes are used where the reference loops: key search inside a trie node, scans over the nodes that
// end at one position (hasFormAlready / isZFollowable), normalisation, index tables, and the final
// reachability + counting sort that replaces removeUnconnected's BFS + stable_sort.  Stores are issued by
// lane 0 and ordered against later loads with __syncwarp().
#include <cuda_runtime.h>
#include "kb_model.h"
#include "kb_batch.h"

namespace kb
{
	__constant__ DevModel c_m;
	// lanes per sentence.  tests/hostsim compiles this file as plain C++ with a one-lane "warp" (KB_HOSTSIM, see
	// tests/hostsim/shim/cuda_runtime.h) to check the kernel's logic against the golden lattices without a GPU.
#if defined(KB_HOSTSIM) && KB_HOSTSIM == 1
	static constexpr uint32_t KB_W = 1;
#else
	static constexpr uint32_t KB_W = 32;
#endif
#if defined(KB_HOSTSIM) && KB_HOSTSIM == 32
#define KB_CHECK_UNIFORM(v) simt::check_uniform((uint32_t)(v), __LINE__)      // tests/hostsim: warp-uniform state really is uniform
#else
#define KB_CHECK_UNIFORM(v) ((void)0)
#endif
	static constexpr unsigned FULL = 0xFFFFFFFFu;
	static constexpr uint32_t NPOS = 0xFFFFFFFFu;
	static constexpr uint32_t MATCH_NORMALIZE_CODA = 1u << 16, MATCH_ZCODA = 1u << 23, MATCH_SPLIT_SAISIOT = 1u << 25, MATCH_MERGE_SAISIOT = 1u << 26;

	// ------------------------------------------------------------------------------------------------
	// pattern matchers (src/PatternMatcher.cpp).  `p` = sentence text, [i, n) = remaining range.
	// ------------------------------------------------------------------------------------------------
	struct PatDev
	{
		const uint16_t* p;
		uint32_t n;          // `last`
		__device__ PatDev(const uint16_t* _p, uint32_t _n) : p{ _p }, n{ _n } {}

		static __device__ bool isAlpha(uint32_t c) { return ('A' <= c && c <= 'Z') || ('a' <= c && c <= 'z'); }
		static __device__ bool isUpperAlpha(uint32_t c) { return 'A' <= c && c <= 'Z'; }
		static __device__ bool isDigit(uint32_t c) { return ('0' <= c && c <= '9') || (0xff10 <= c && c <= 0xff19); }
		static __device__ bool alnum(uint32_t c) { return isAlpha(c) || ('0' <= c && c <= '9'); }
		static __device__ bool emailAccount(uint32_t c) { return alnum(c) || c == '-' || c == '.' || c == '_' || c == '%' || c == '+'; }
		static __device__ bool alphaNumDotDash(uint32_t c) { return alnum(c) || c == '-' || c == '.'; }
		static __device__ bool domain(uint32_t c) { return alnum(c) || c == '-' || c == '@' || c == ':' || c == '%' || c == '.' || c == '_' || c == '+' || c == '~' || c == '#' || c == '='; }
		static __device__ bool path(uint32_t c) { return domain(c) || c == '(' || c == ')' || c == '!' || c == '?' || c == '&' || c == '/'; }
		static __device__ bool hashtags(uint32_t c)
		{
			switch (c) { case '#': case ' ': case '\t': case '\n': case '\r': case '\v': case '\f': case '.': case ',': case '(': case ')': case '[': case ']': case '<': case '>': case '{': case '}': return false; }
			return true;
		}
		static __device__ bool spaceSet(uint32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }
		__device__ bool lit(uint32_t i, const char* s, uint32_t len) const
		{
			if (n - i < len) return false;
			for (uint32_t k = 0; k < len; ++k) if (p[i + k] != (uint16_t)s[k]) return false;
			return true;
		}

		__device__ uint32_t testUrl(uint32_t first) const
		{
			uint32_t b;
			if (lit(first, "http://", 7)) b = first + 7;
			else if (lit(first, "https://", 8)) b = first + 8;
			else return 0;
			int state = 0;
			uint32_t lastMatched = first;
			if (b == n || !domain(p[b])) return 0;
			++b;
			for (; b != n && domain(p[b]); ++b)
			{
				if (p[b] == '.') state = 1;
				else if (isAlpha(p[b]))
				{
					if (state > 0) ++state;
					if (state >= 3) lastMatched = b + 1;
				}
				else state = 0;
			}
			if (lastMatched == first) return 0;
			b = lastMatched;
			if (b != n && p[b] == ':')
			{
				++b;
				if (b == n || !isDigit(p[b])) return 0;
				++b;
				while (b != n && isDigit(p[b])) ++b;
			}
			if (b != n && p[b] == '/')
			{
				++b;
				while (b != n && path(p[b])) ++b;
			}
			else
			{
				if (b != n && !spaceSet(p[b])) return 0;
			}
			if (p[b - 1] == '.' || p[b - 1] == ':') --b;
			return b - first;
		}
		__device__ uint32_t testEmail(uint32_t first) const
		{
			uint32_t b = first;
			if (b == n || !emailAccount(p[b])) return 0;
			++b;
			while (b != n && emailAccount(p[b])) ++b;
			if (b == n || p[b] != '@') return 0;
			++b;
			int state = 0;
			uint32_t lastMatched = first;
			if (b == n || !alphaNumDotDash(p[b])) return 0;
			++b;
			for (; b != n && alphaNumDotDash(p[b]); ++b)
			{
				if (p[b] == '.') state = 1;
				else if (isAlpha(p[b]))
				{
					if (state > 0) ++state;
					if (state >= 3) lastMatched = b + 1;
				}
				else state = 0;
			}
			return lastMatched - first;
		}
		__device__ uint32_t testMention(uint32_t first) const
		{
			uint32_t b = first;
			if (b == n || p[b] != '@') return 0;
			++b;
			if (b == n || !isAlpha(p[b])) return 0;
			++b;
			while (b != n && emailAccount(p[b])) ++b;
			if (p[b - 1] == '.' || p[b - 1] == '%' || p[b - 1] == '+' || p[b - 1] == '-') --b;
			if (b - first <= 3) return 0;
			return b - first;
		}
		__device__ uint32_t testHashtag(uint32_t first) const
		{
			uint32_t b = first;
			if (b == n || p[b] != '#') return 0;
			++b;
			if (b == n || !hashtags(p[b])) return 0;
			++b;
			while (b != n && hashtags(p[b])) ++b;
			return b - first;
		}
		__device__ uint32_t testNumeric(uint32_t left, uint32_t first) const
		{
			uint32_t b = first;
			bool hasComma = false;
			if (b == n || !isDigit(p[b])) return 0;
			while (b != n && isDigit(p[b])) ++b;
			while (b != n && p[b] == ',')
			{
				++b;
				if (b + 2 >= n || !isDigit(p[b]) || !isDigit(p[b + 1]) || !isDigit(p[b + 2])) return b - 1 - first;
				b += 3;
				hasComma = true;
			}
			if (b == n || isSpaceChr(c_m, p[b]) || isHangulSyllable(p[b])) return b - first;
			if (p[b] == '.')
			{
				++b;
				if (!hasComma && !alphaNumDotDash(left) && (b == n || !alphaNumDotDash(p[b]))) return b - first;
				if (b == n || !isDigit(p[b])) return b - 1 - first;
				while (b != n && isDigit(p[b])) ++b;
			}
			if (b == n || p[b] != '.') return b - first;
			return 0;
		}
		__device__ uint32_t testSerial(uint32_t first) const
		{
			uint32_t b = first;
			if (b == n || !isDigit(p[b])) return 0;
			while (b != n && isDigit(p[b])) ++b;
			if (b == n) return 0;
			uint32_t sep;
			if (p[b] == ':' || p[b] == '.' || p[b] == '-' || p[b] == '/') sep = p[b];
			else return 0;
			++b;
			if (b != n && p[b] == ' ') ++b;
			if (b == n || !isDigit(p[b])) return 0;
			++b;
			while (b != n && isDigit(p[b])) ++b;
			if (sep == '.' && (b == n || p[b] != sep)) return 0;
			while (b != n && p[b] == sep)
			{
				++b;
				if (b != n && p[b] == ' ') ++b;
				if (b == n || !isDigit(p[b]))
				{
					if (p[b - 1] == ' ') --b;
					return b - first;
				}
				++b;
				while (b != n && isDigit(p[b])) ++b;
			}
			if (p[b - 1] == ' ') --b;
			return b - first;
		}
		__device__ uint32_t testAbbr(uint32_t first) const
		{
			uint32_t b = first;
			if (b == n || !isAlpha(p[b])) return 0;
			uint32_t l = 0;
			while (b != n && isAlpha(p[b])) ++b, ++l;
			if (b == n) return 0;
			if (p[b] == '.') ++b;
			else return 0;
			if (b != n && p[b] == ' ')
			{
				if (l > (isUpperAlpha(p[first]) ? 5u : 3u)) return 0;
				return b - first;
			}
			else
			{
				if (l > 5) return 0;
			}
			while (b != n && isAlpha(p[b]))
			{
				l = 0;
				while (b != n && isAlpha(p[b])) ++b, ++l;
				if (l > 5) return 0;
				if (b != n && p[b] == '.') ++b;
				else return b - first;
			}
			if (p[b - 1] == ' ') --b;
			return b - first;
		}
		// p[n] may be read by the reference (`*b` at last); the normalized buffer is zero-padded for that.
		__device__ uint32_t testEmoji(uint32_t first) const
		{
			uint32_t b = first;
			while (b + 1 < n)
			{
				uint32_t c0 = 0, c1 = 0;
				uint32_t b1 = b;
				if (isHighSurrogate(p[b1])) { c0 = mergeSurrogate(p[b1], p[b1 + 1]); b1 += 2; }
				else c0 = p[b1++];
				uint32_t b2 = b1;
				if (b2 < n)
				{
					if (isHighSurrogate(p[b2]) && b2 + 1 < n) { c1 = mergeSurrogate(p[b2], p[b2 + 1]); b2 += 2; }
					else c1 = p[b2++];
				}
				const int r = isEmoji(c_m, c0, c1);
				if (r == 1) b = b1;
				else if (r == 2) b = b2;
				else break;
				if (b == n) return b - first;
				if (0xfe00 <= p[b] && p[b] <= 0xfe0f)
				{
					++b;
					if (b == n) return b - first;
				}
				else if (b + 1 < n && isHighSurrogate(p[b]))
				{
					c1 = mergeSurrogate(p[b], p[b + 1]);
					if (0x1f3fb <= c1 && c1 <= 0x1f3ff)
					{
						b += 2;
						if (b == n) return b - first;
					}
				}
				if (p[b] == 0x200d) { ++b; continue; }
				break;
			}
			return b - first;
		}

		// PatternMatcherImpl::match, :366-378.  returns length, tag in `tag`
		__device__ uint32_t match(uint32_t left, uint32_t i, uint32_t opt, uint32_t& tag) const
		{
			// quick reject: every tester needs an ASCII / full-width-digit first unit, or an emoji-capable one.
			const uint32_t c = p[i];
			if (c >= 0x80 && !(0xff10 <= c && c <= 0xff19))
			{
				if (!(opt & (1u << 5))) return 0;
				// emoji test only; hashtag/others cannot start here ('#', '@', digits, letters are ASCII)
				const uint32_t size = (i + 1 < n) ? testEmoji(i) : 0;
				if (size) { tag = T_w_emoji; return size; }
				return 0;
			}
			uint32_t size;
			if ((opt & (1u << 4)) && (size = testSerial(i))) { tag = T_w_serial; return size; }
			if ((size = testNumeric(left, i))) { tag = T_sn; return size; }
			if ((opt & (1u << 2)) && (size = testHashtag(i))) { tag = T_w_hashtag; return size; }
			if ((opt & (1u << 1)) && (size = testEmail(i))) { tag = T_w_email; return size; }
			if ((opt & (1u << 3)) && (size = testMention(i))) { tag = T_w_mention; return size; }
			if ((opt & (1u << 0)) && (size = testUrl(i))) { tag = T_w_url; return size; }
			if ((opt & (1u << 5)) && (size = testEmoji(i))) { tag = T_w_emoji; return size; }
			if ((size = testAbbr(i))) { tag = T_sl; return size; }
			return 0;
		}
	};

	// ------------------------------------------------------------------------------------------------
	struct Builder
	{
		const BatchView& bv;
		const uint32_t lane;
		// sentence
		uint32_t s;
		uint16_t* norm; uint32_t normLen;
		uint32_t W;
		// chunk
		const uint16_t* raw; uint32_t rawLen; uint32_t startOffset;
		uint32_t* nsToPos; uint32_t* posToNs; uint2* endPosMap; uint32_t* ctr; DPattern* pats;
		uint32_t nNs, nPats, nextPat;
		DNode* out; uint32_t nOut, outCap;
		uint32_t lastEndPos;      // out.back().endPos
		uint32_t err;

		__device__ Builder(const BatchView& _bv, uint32_t _lane) : bv{ _bv }, lane{ _lane }, err{ 0 } {}

		// ---- appendNewNode, KTrie.cpp:16-43 ---------------------------------------------------------
		__device__ bool appendNewNode(uint32_t startPos, uint32_t endPos, int32_t form, uint32_t uoff, uint32_t ulen, float typoCost, uint32_t spaceErrors)
		{
			const uint2 es = endPosMap[startPos];
			if (es.x == es.y) return false;
			const uint32_t newId = nOut;
			if (newId >= outCap) { err = ST_NODE_OVERFLOW; return false; }
			// relative links (DNode::prev / sibling) and DPath::node are 16-bit: a chunk with more nodes is reported, never wrapped
			if (newId >= 0xFFFFu) { err = ST_TOO_LONG; return false; }
			// the scans that decided this append (hasFormAlready / isZFollowable / `es`) read endPosMap and out[]: every lane
			// must be past them before lane 0 changes those arrays (found by the 32-lane host simulation, tests/hostsim)
			__syncwarp();
			if (lane == 0)
			{
				DNode nd;
				nd.form = form; nd.uform_off = uoff; nd.uform_len = ulen;
				nd.start_pos = startPos; nd.end_pos = endPos;
				nd.prev = (uint16_t)(newId - es.x); nd.sibling = 0;
				nd.space_errors = (uint16_t)spaceErrors; nd.reserved = 0; nd.typo_cost = typoCost;
				out[newId] = nd;
			}
			nOut = newId + 1;
			lastEndPos = endPos;
			if (endPos < nNs + 1)
			{
				const uint2 ee = endPosMap[endPos];
				if (lane == 0)
				{
					if (ee.x == ee.y) endPosMap[endPos] = make_uint2(newId, newId + 1);
					else
					{
						out[ee.y - 1].sibling = (uint16_t)(newId - (ee.y - 1));
						endPosMap[endPos] = make_uint2(ee.x, newId + 1);
					}
				}
			}
			__syncwarp();
			return true;
		}

		// ---- hasFormAlready, KTrie.cpp:897-905 (lanes scan the nodes registered at endPos) ----------
		__device__ bool hasFormAlready(uint32_t startPos, uint32_t endPos) const
		{
			const uint2 e = endPosMap[endPos];
			if (e.x == NPOS) return false;
			const uint32_t scanStart = e.x > 1u ? e.x : 1u;
			bool found = false;
			for (uint32_t base = scanStart; base < e.y; base += KB_W)
			{
				const uint32_t i = base + lane;
				bool hit = false;
				if (i < e.y)
				{
					const DNode g = out[i];
					const uint32_t size = g.uform_len == 0 ? (uint32_t)c_m.forms[g.form].size_no_space : g.uform_len;
					hit = g.end_pos == endPos && g.end_pos - size == startPos && g.typo_cost == 0.f
						&& (g.form < 0 || (c_m.forms[g.form].flags & FF_HASFULL));
				}
				if (__any_sync(FULL, hit)) { found = true; break; }
			}
			return found;
		}

		// ---- isZFollowable, KTrie.cpp:907-919: returns bit0 zCoda, bit1 zSiot ------------------------
		__device__ uint32_t isZFollowable(uint32_t pos) const
		{
			if (pos >= nNs) return 0;
			const uint2 e = endPosMap[pos];
			if (e.x == NPOS) return 0;
			uint32_t acc = 0;
			for (uint32_t base = e.x; base < e.y; base += KB_W)
			{
				const uint32_t i = base + lane;
				uint32_t f = 0;
				if (i < e.y)
				{
					const DNode g = out[i];
					if (g.end_pos == pos && g.form >= 0) f = c_m.forms[g.form].flags & (FF_ZCODA | FF_ZSIOT);
				}
				acc |= __reduce_or_sync(FULL, f);
			}
			return acc;
		}

		__device__ void appendRaw(uint32_t sNs, uint32_t eNs)
		{
			const uint32_t off = nsToPos[sNs];
			uint32_t len = nsToPos[eNs - 1] + 1 - off;
			while (len && isSpaceChr(c_m, raw[off + len - 1])) --len;
			appendNewNode(sNs, eNs, -1, startOffset + off, len, 0.f, 0);
		}

		// ---- insertUnkForm, KTrie.cpp:921-953 ---------------------------------------------------------
		__device__ void insertUnkForm(uint32_t startPos, uint32_t endPos, bool hasJClass)
		{
			if (startPos >= endPos || hasFormAlready(startPos, endPos)) return;
			uint32_t lastPos = lastEndPos;
			if (lastPos < endPos)
			{
				if (lastPos && isHangulCoda(raw[nsToPos[lastPos]])) lastPos--;
				if (lastPos != startPos && !hasFormAlready(lastPos, endPos)) appendRaw(lastPos, endPos);
			}
			const uint32_t newNodeLength = endPos - startPos;
			const uint32_t lengthLimit = hasJClass ? c_m.cfg.max_unk_form_size_followed_by_jclass : c_m.cfg.max_unk_form_size;
			if (newNodeLength <= lengthLimit) appendRaw(startPos, endPos);
		}

		// ---- countSpaceErrors, KTrie.cpp:316-328 ------------------------------------------------------
		__device__ uint32_t countSpaceErrors(int32_t form, uint32_t nBegin, uint32_t nEnd) const
		{
			const uint32_t size = nEnd - nBegin;
			if (size < 2) return 0;
			// no gap inside the span -> no error can be counted
			if (nsToPos[nEnd - 1] - nsToPos[nBegin] == size - 1) return 0;
			const uint16_t* f = c_m.form_chars + c_m.forms_raw[form].str_off;
			uint32_t cnt = 0, spaceOffset = 0;
			for (uint32_t i = 1; i < size; ++i)
			{
				const bool hasSpace = nsToPos[nBegin + i] - nsToPos[nBegin + i - 1] > 1;
				const uint16_t fc = f[i + spaceOffset];
				if (hasSpace && fc != ' ') ++cnt;
				spaceOffset += fc == ' ' ? 1 : 0;
			}
			return cnt;
		}

		// ---- one candidate of flushCandidates, KTrie.cpp:955-996 --------------------------------------
		__device__ void flushCandidate(int32_t cand, uint32_t endPosition, uint32_t unkFormStartNsPos, uint32_t lastSpaceBoundaryNsPos,
			int32_t startPosOffset = 0, float typoCost = 0.f)
		{
			const DForm f = c_m.forms[cand];
			const uint32_t nBegin = endPosition - f.size_no_space + (uint32_t)startPosOffset;
			const uint32_t nEnd = endPosition;
			if (!(f.flags & FF_FIRST_IS_CODA))
			{
				const bool hj = (f.flags & FF_HASJ_OR_STAG) != 0;
				if (lastSpaceBoundaryNsPos < nBegin) insertUnkForm(lastSpaceBoundaryNsPos, nBegin, hj);
				insertUnkForm(unkFormStartNsPos, nBegin, hj);
			}
			const uint32_t spaceErrors = countSpaceErrors(cand, nBegin, nEnd);
			if (spaceErrors <= c_m.cfg.space_tolerance) appendNewNode(nBegin, nEnd, cand, 0, 0, typoCost, spaceErrors);
		}

		// ---- trie step: child of `node` for key c, -1 if none (FrozenTrie.hpp:14-22).  Root: direct table. --
		__device__ int32_t nextOpt(int32_t node, uint32_t c) const
		{
			if (node == 0) return c_m.trie_root_next[c];
			const kb2_trie_node nd = c_m.trie_nodes[node];
			for (uint32_t base = 0; base < nd.num_nexts; base += KB_W)
			{
				const uint32_t i = base + lane;
				const uint32_t k = i < nd.num_nexts ? (uint32_t)c_m.trie_keys[nd.next_offset + i] : 0x10000u;
				const unsigned hit = __ballot_sync(FULL, k == c);
				if (hit) return node + c_m.trie_diffs[nd.next_offset + base + (__ffs(hit) - 1)];
				// keys ascend: stop when the last key of this tile is already larger than c
				if (__any_sync(FULL, k > c)) break;
			}
			return -1;
		}

		__device__ void specialRunNode(uint32_t specialStartNsPos, uint32_t rawEnd, uint32_t endNs, uint32_t lastChrType)
		{
			const uint32_t off = nsToPos[specialStartNsPos];
			uint32_t len = rawEnd - off;
			while (len && isSpaceChr(c_m, raw[off + len - 1])) --len;
			appendNewNode(specialStartNsPos, endNs, c_m.trie_nodes[lastChrType].value, startOffset + off, len, 0.f, 0);
		}

		static __device__ bool isDiscontinuous(uint32_t prevTag, uint32_t curTag, uint32_t prevScript, uint32_t curScript)
		{
			if ((prevTag == T_sl || prevTag == T_sh || prevTag == T_sw) && (curTag == T_sl || curTag == T_sh || curTag == T_sw)) return prevScript != curScript;
			return prevTag != curTag;
		}

		// ---- preparePattern, KTrie.cpp:766-858.  `str` = norm + startOffset, `len` = rest of the sentence --
		__device__ uint32_t preparePattern(const uint16_t* str, uint32_t len)
		{
			PatDev pat{ str, len };
			uint32_t n = 0, continuousNonSpaceCount = 0;
			uint32_t lastChrType = T_unknown;
			nPats = 0;
			for (; n < len; ++n)
			{
				{
					uint32_t tag = T_unknown;
					const uint32_t ml = pat.match(n ? str[n - 1] : (uint32_t)' ', n, bv.match_options, tag);
					if (tag != T_unknown)
					{
						if (lane == 0) pats[nPats] = DPattern{ n + ml, ml, tag };
						++nPats;
						n += ml - 1;
						continue;
					}
				}
				const uint32_t c = str[n];
				uint32_t c32 = c;
				if (isHighSurrogate(c32) && n + 1 < len) c32 = mergeSurrogate(c32, str[n + 1]);
				const uint32_t chrType = attrCls(chrAttr(c_m, c32));
				if (chrType == T_unknown) continuousNonSpaceCount = 0;
				else continuousNonSpaceCount++;
				if (chrType == T_unknown && n >= (lastChrType == T_sf ? 4u : 4096u))
				{
					if (!isSpaceChr(c_m, str[n - 3]) && !isSpaceChr(c_m, str[n - 2])) break;
				}
				else if (continuousNonSpaceCount >= 1024) break;
				if (c32 >= 0x10000) ++n;
				lastChrType = chrType;
			}
			if (n > len) n = len;     // a trailing unpaired high surrogate cannot advance past the end (c32 < 0x10000 then), defensive
			__syncwarp();
			// nsToPos / posToNs (837-854), lane-parallel: a unit is "non-space" unless isSpace; a low unit that
			// follows a high surrogate is always kept with it.
			uint32_t nsCount = 0;
			for (uint32_t base = 0; base < n; base += KB_W)
			{
				const uint32_t i = base + lane;
				bool keep = false;
				if (i < n)
				{
					const uint16_t c = str[i];
					keep = !isSpaceChr(c_m, c);
					// the reference never tests the unit after a non-space high surrogate (i + 1 < n)
					if (!keep && i > 0 && isHighSurrogate(str[i - 1]) && !isSpaceChr(c_m, str[i - 1]))
					{
						// str[i-1] is a high surrogate that was itself consumed as a "first" unit only if it is not
						// the second unit of an earlier pair; surrogates are never spaces, so chains resolve by parity.
						uint32_t k = i - 1, run = 1;
						while (k > 0 && isHighSurrogate(str[k - 1])) { --k; ++run; }
						if (run & 1) keep = true;
					}
				}
				const unsigned mask = __ballot_sync(FULL, keep);
				const uint32_t before = __popc(mask & ((1u << lane) - 1));
				if (i < n)
				{
					posToNs[i] = nsCount + before;
					if (keep) nsToPos[nsCount + before] = i;
				}
				nsCount += __popc(mask);
			}
			if (lane == 0) posToNs[n] = nsCount;
			nNs = nsCount;
			// sort(matchedPatterns): they are produced with strictly increasing end (each match consumes its span)
			nextPat = 0;
			__syncwarp();
			return n;
		}

		// ---- the walk over one chunk: progressNode (998-1412) + tail of search (1434-1452) -------------
		__device__ void search()
		{
			const uint32_t opt = bv.match_options;
			uint32_t lastChrType = T_unknown, lastScriptType = 0;
			uint32_t specialStartNsPos = 0, unkFormStartNsPos = 0, lastSpaceBoundaryNsPos = 0;
			int32_t curNode = 0;
			DPattern np = nPats ? pats[0] : DPattern{ NPOS, 0, 0 };
			for (uint32_t j = 0; j < rawLen; ++j)
			{
				const uint32_t c = raw[j];
				uint32_t c32 = c;
				if (isHighSurrogate(c32) && j + 1 < rawLen) c32 = mergeSurrogate(c32, raw[j + 1]);
				const bool havePat = nextPat != nPats;
				{
					const bool isInPattern = havePat && j >= np.end - np.len;
					const uint32_t attr = chrAttr(c_m, c32);
					uint32_t chrType = attrCls(attr), scriptType = attrScript(attr);
					if (lastChrType == T_sw && (c32 == 0x200d || (0x1f3fb <= c32 && c32 <= 0x1f3ff) || scriptType == c_m.script_variation_selectors))
					{
						chrType = lastChrType;
						scriptType = lastScriptType;
					}
					if (isDiscontinuous(lastChrType, isInPattern ? (uint32_t)T_unknown : chrType, lastScriptType, scriptType)
						|| lastChrType == T_sso || lastChrType == T_ssc)
					{
						if (lastChrType != T_max && lastChrType != T_unknown && lastChrType != T_ss)
						{
							const bool hj = T_sf <= lastChrType && lastChrType <= T_sw;
							if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, specialStartNsPos, hj);
							insertUnkForm(unkFormStartNsPos, specialStartNsPos, hj);
							specialRunNode(specialStartNsPos, j, posToNs[j], lastChrType);
						}
						unkFormStartNsPos = specialStartNsPos;
						specialStartNsPos = posToNs[j];
						if (T_sf <= lastChrType && lastChrType <= T_sw) lastSpaceBoundaryNsPos = specialStartNsPos;
					}
					else if (chrType == T_max)
					{
						unkFormStartNsPos = specialStartNsPos;
					}
					lastChrType = isInPattern ? (uint32_t)T_unknown : chrType;
					lastScriptType = scriptType;

					if (c32 < 0x10000)
					{
						if (chrType == T_unknown)      // whitespace
						{
							const uint32_t nx = posToNs[j + 1];
							if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, nx, true);
							insertUnkForm(unkFormStartNsPos, nx, true);
							lastSpaceBoundaryNsPos = specialStartNsPos = unkFormStartNsPos = nx;
							continue;
						}
						// z-coda / saisiot built-in forms (1126-1135)
						if ((opt & MATCH_ZCODA) && isHangulCoda(c) && (j + 1 >= rawLen || !isHangulSyllable(raw[j + 1])))
						{
							if (isZFollowable(posToNs[j]) & FF_ZCODA)
							{
								// pushed before the trie candidates of this position -> flushed first
								zCand = (int32_t)(c_m.default_tag_size + (c - 0x11A8) - 1);
							}
						}
						else if ((opt & (MATCH_SPLIT_SAISIOT | MATCH_MERGE_SAISIOT)) && c == 0x11BA && j + 1 < rawLen && isHangulSyllable(raw[j + 1]))
						{
							if (isZFollowable(posToNs[j]) & FF_ZSIOT) zCand = (int32_t)(c_m.default_tag_size + (0x11BA - 0x11A8) - 1);
						}
					}
				}
				if (havePat)
				{
					const uint32_t currentEnd = j + (c32 >= 0x10000 ? 2 : 1);
					while (nextPat != nPats && np.end == currentEnd)
					{
						const uint32_t matchedStart = np.end - np.len;
						const bool hj = T_w_url <= np.tag && np.tag <= T_w_emoji;
						const uint32_t ms = posToNs[matchedStart];
						if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, ms, hj);
						insertUnkForm(unkFormStartNsPos, ms, hj);
						appendNewNode(ms, posToNs[np.end], c_m.trie_nodes[np.tag].value, startOffset + matchedStart, np.len, 0.f, 0);
						++nextPat;
						if (nextPat != nPats) np = pats[nextPat];
					}
				}
				if (c32 >= 0x10000)
				{
					++j;
					continue;
				}

				// Aho-Corasick step (1283-1289)
				int32_t nextNode = nextOpt(curNode, c);
				while (nextNode < 0)
				{
					const int32_t fl = c_m.trie_nodes[curNode].fail;
					if (!fl) { curNode = -1; break; }
					curNode += fl;
					nextNode = nextOpt(curNode, c);
				}
				const uint32_t endPosition = posToNs[j + 1];
				if (zCand >= 0)
				{
					flushCandidate(zCand, endPosition, unkFormStartNsPos, lastSpaceBoundaryNsPos);
					zCand = -1;
				}
				if (nextNode >= 0)
				{
					curNode = nextNode;
					// all suffix matches, longest first (1302-1311); each is flushed immediately, which is
					// equivalent to the reference's collect-then-flush because flushing never touches the trie state.
					for (int32_t sub = curNode; ; )
					{
						const kb2_trie_node sn = c_m.trie_nodes[sub];
						if (sn.value == KB2_TRIE_NONE) break;
						if (sn.value != KB2_TRIE_SUBMATCH) flushCandidate(sn.value, endPosition, unkFormStartNsPos, lastSpaceBoundaryNsPos);
						if (!sn.fail) break;
						sub += sn.fail;
					}
				}
				else curNode = 0;
			}
			if (lastChrType != T_max && lastChrType != T_unknown && lastChrType != T_ss)
			{
				const bool hj = T_sf <= lastChrType && lastChrType <= T_sw;
				if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, specialStartNsPos, hj);
				insertUnkForm(unkFormStartNsPos, specialStartNsPos, hj);
				specialRunNode(specialStartNsPos, rawLen, posToNs[rawLen], lastChrType);
				unkFormStartNsPos = specialStartNsPos;
				if (hj) lastSpaceBoundaryNsPos = posToNs[rawLen];
			}
			const uint32_t totEndPos = nsToPos[nNs - 1] + 1;
			if (rawLen == totEndPos)
			{
				if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, posToNs[totEndPos], true);
				insertUnkForm(unkFormStartNsPos, posToNs[totEndPos], true);
			}
			// EOS node (1449-1450): appended with endPos = nNs + 1 (never registered), then endPos of out.back() := nNs
			appendNewNode(nNs, nNs + 1, -1, 0, 0, 0.f, 0);
			if (lane == 0) out[nOut - 1].end_pos = nNs;
			__syncwarp();
		}
		int32_t zCand = -1;

		// =================================================================================================
		// Typo lattice (BASELINE config 4).  (1) the typo graph of the chunk, PreparedTypoTransformer::generateGraph
		// (src/TypoTransformer.cpp:810-1039, appendNewNode 594-629) for the standard dialect; (2) the general walk of
		// Splitter::search / progressNode (src/KTrie.cpp:998-1452) with one search state per (graph node, surviving trie
		// state).  search() above is (2) for the 2-node graph.  Without continual / lengthening typo sets and
		// pretokenized spans, as the rest of this kernel.
		// The graph is built by lane 0 alone (a sequential Aho-Corasick walk with data-dependent appends); the walk
		// over it is warp-uniform like search() and shares its lane-parallel helpers.
		// =================================================================================================
		DTypoNode* tg; DTypoNode* tgTmp; uint32_t* tgRemap; uint2* tgRange; DTypoState* tgStates; DTypoMatch* tgMatches;
		uint32_t tgCap, tgStateCap, nTg;

		__device__ int32_t typoNext(int32_t node, uint32_t c) const
		{
			const kb2_typo_node nd = bv.typo.nodes[node];
			uint32_t lo = 0, hi = nd.num_nexts;
			const uint16_t* k = bv.typo.keys + nd.next_offset;
			while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (k[mid] < c) lo = mid + 1; else hi = mid; }
			if (lo == nd.num_nexts || k[lo] != c) return -1;
			return node + bv.typo.diffs[nd.next_offset + lo];
		}

		// appendNewNode of the typo graph (TypoTransformer.cpp:594-629); `map` = endPosMap over [mapOffset, mapOffset + mapSize)
		// holding {first, last} node ids per end position, NPOS = none.  startPos == NPOS: link to the previous node.
		__device__ bool tgAppend(uint32_t& n, uint2* map, uint32_t mapSize, uint32_t mapOffset, bool fromPool, uint32_t off, uint32_t len,
			uint32_t startPos, uint32_t endPos, float cost)
		{
			if (startPos != NPOS && (startPos < mapOffset || map[startPos - mapOffset].x == NPOS)) return false;
			if (n >= tgCap) { err = ST_TYPO_OVERFLOW; return false; }
			DTypoNode nn;
			nn.end_pos = endPos; nn.typo_cost = cost; nn.sibling = 0; nn.off = off; nn.len = (uint16_t)len; nn.from_pool = fromPool ? 1 : 0; nn.continual_idx = 0;
			nn.prev = startPos == NPOS ? n - 1 : map[startPos - mapOffset].x;
			const uint32_t newId = n++;
			tgTmp[newId] = nn;
			if (endPos >= mapSize + mapOffset) return true;
			uint2 slot = map[endPos - mapOffset];
			if (slot.x == NPOS) slot.x = newId;
			else tgTmp[slot.y].sibling = newId;
			slot.y = newId;
			map[endPos - mapOffset] = slot;
			return true;
		}

		// insertBranch (TypoTransformer.cpp:842-1002): the matches [0, nM) overlap transitively
		__device__ void tgInsertBranch(uint32_t& n, uint2* map, uint2& mapBack, uint32_t& last, uint32_t nM, const uint16_t* str)
		{
			DTypoMatch* M = tgMatches;
			const uint32_t totStartPos = M[0].end_pos - M[0].pat_len;
			const uint32_t totEndPos = M[nM - 1].end_pos;
			const uint32_t mapSize = totEndPos - last + 1;
			for (uint32_t i = 0; i < mapSize; ++i) map[i] = make_uint2(NPOS, NPOS);
			map[0] = mapBack;
			// break points: the branch start and every match end, ascending and unique (flags over [last, totEndPos] in ctr)
			uint32_t* flag = ctr;
			for (uint32_t i = 0; i < mapSize; ++i) flag[i] = 0;
			flag[totStartPos - last] = 1;
			for (uint32_t i = 0; i < nM; ++i) flag[M[i].end_pos - last] = 1;
			// sort the matches by start position (ties carry different ends, whose relative order cannot reach the sorted graph)
			for (uint32_t i = 1; i < nM; ++i)
			{
				const DTypoMatch m = M[i];
				const uint32_t ms = m.end_pos - m.pat_len;
				uint32_t j = i;
				while (j > 0 && M[j - 1].end_pos - M[j - 1].pat_len > ms) { M[j] = M[j - 1]; --j; }
				M[j] = m;
			}
			if (last < totStartPos) tgAppend(n, map, mapSize, last, false, last, totStartPos - last, last, totStartPos, 0.f);
			{
				uint32_t prevBp = totStartPos;
				for (uint32_t q = totStartPos + 1; q <= totEndPos; ++q)
				{
					if (!flag[q - last]) continue;
					tgAppend(n, map, mapSize, last, false, prevBp, q - prevBp, prevBp, q, 0.f);
					prevBp = q;
				}
			}
			const float continualThr = bv.typo.continual_threshold;
			for (uint32_t mi = 0; mi < nM && !err; ++mi)
			{
				const DTypoMatch m = M[mi];
				const uint32_t e = m.end_pos, s = e - m.pat_len;
				// continualIdx: first unit of the replacement -> (index, node) (TypoTransformer.cpp:905-960)
				uint16_t ckey[8]; uint32_t cnode[8]; uint32_t nCont = 0;
				for (uint32_t j = 0; j < m.size; ++j)
				{
					const kb2_typo_repl repl = bv.typo.repls[m.repl_off + j];
					if (repl.dialect != 0) continue;      // allowedDialect == standard
					if (repl.left_cond == 2 /* CondVowel::vowel */)
					{
						if (s == 0 || !isHangulSyllable(str[s - 1])) continue;
					}
					else if (repl.left_cond == 1 /* any */)
					{
						if (s == 0) continue;
					}
					else if (repl.left_cond == 9 /* continual */ || repl.left_cond == 10 /* boundary */)
					{
						const bool continual = repl.left_cond == 9;
						if (continual && (s == 0 || !isHangulSyllable(str[s - 1]))) continue;
						if (continual && !(continualThr < 3.0e38f)) continue;
						const float scale = continual ? continualThr : 1.f;
						const uint16_t key = bv.typo.pool[repl.str_off];
						uint32_t k = 0;
						while (k < nCont && ckey[k] != key) ++k;
						if (k == nCont)
						{
							if (nCont == 8) { err = ST_TYPO_OVERFLOW; break; }
							if (tgAppend(n, map, mapSize, last, true, repl.str_off, 1, s, NPOS, repl.cost * scale / 2))
							{
								tgTmp[n - 1].end_pos = e;
								tgTmp[n - 1].continual_idx = (uint8_t)(nCont + 1);
								ckey[nCont] = key; cnode[nCont] = n - 1; ++nCont;
								if (tgAppend(n, map, mapSize, last, true, repl.str_off + 1, repl.length - 1, NPOS, e, repl.cost * scale / 2)) tgTmp[n - 1].prev = cnode[k];
							}
						}
						else
						{
							if (tgAppend(n, map, mapSize, last, true, repl.str_off + 1, repl.length - 1, NPOS, e, repl.cost * scale / 2)) tgTmp[n - 1].prev = cnode[k];
						}
						continue;
					}
					else
					{
						if (!ftVowel(s == 0, s ? str[s - 1] : 0, repl.left_cond)) continue;
					}
					tgAppend(n, map, mapSize, last, true, repl.str_off, repl.length, s, e, repl.cost);
				}
			}
			mapBack = map[mapSize - 1];
			last = totEndPos;
		}

		// generateGraph over raw[0, rawLen) -> tg[0, nTg), sorted by end position with relative prev / sibling offsets
		__device__ __noinline__ void genTypoGraph()      // out of line: the plain walk keeps its register allocation
		{
			uint32_t n = 0;
			if (lane == 0)
			{
				uint2* map = endPosMap;        // free until search() (re)initialises it; rawLen + 2 <= W entries
				const uint16_t* str = raw;
				uint2 mapBack = make_uint2(0, 0);
				uint32_t last = 0, nM = 0;
				{
					DTypoNode bos; bos.end_pos = 0; bos.typo_cost = 0.f; bos.prev = 0; bos.sibling = 0; bos.off = 0; bos.len = 0; bos.from_pool = 0; bos.continual_idx = 0;
					tgTmp[n++] = bos;
				}
				int32_t node = typoNext(0, 0);
				if (node < 0) err = ST_INTERNAL;
				for (uint32_t i = 0; i < rawLen && !err; ++i)
				{
					const uint32_t c = str[i];
					int32_t nnode = typoNext(node, c);
					while (nnode < 0)
					{
						const int32_t fl = bv.typo.nodes[node].fail;
						if (fl) { node += fl; nnode = typoNext(node, c); }
						else { node = 0; break; }
					}
					if (nnode < 0) continue;
					node = nnode;
					const int32_t v = bv.typo.nodes[node].value;
					if (v == -1) continue;
					const uint32_t endPos = i + 1;
					// a sub-match-only node carries patLength = -1 in the reference: its start wraps far beyond any end position
					if (nM && (v < 0 || tgMatches[nM - 1].end_pos < endPos - bv.typo.pats[v].pat_len))
					{
						tgInsertBranch(n, map, mapBack, last, nM, str);
						nM = 0;
					}
					for (int32_t sub = node; ; )
					{
						const kb2_typo_node sn = bv.typo.nodes[sub];
						if (sn.value == -1) break;
						if (sn.value != -2)
						{
							if (nM >= tgCap) { err = ST_TYPO_OVERFLOW; break; }
							const kb2_typo_pat pt = bv.typo.pats[sn.value];
							tgMatches[nM++] = DTypoMatch{ endPos, pt.repl_off, pt.size, pt.pat_len };
						}
						if (!sn.fail) break;
						sub += sn.fail;
					}
				}
				if (nM && !err) tgInsertBranch(n, map, mapBack, last, nM, str);
				if (!err)
				{
					// the tail segment: registered nowhere, its end position becomes the chunk length (TypoTransformer.cpp:1014-1018)
					map[0] = mapBack;
					if (tgAppend(n, map, 1, last, false, last, rawLen - last, last, rawLen + 1, 0.f)) tgTmp[n - 1].end_pos = rawLen;
				}
				if (!err)
				{
					// stable sort by end position (a counting sort over [0, rawLen]) and renumbering of the links (1020-1036)
					uint32_t* cnt = ctr;
					for (uint32_t i = 0; i <= rawLen + 1; ++i) cnt[i] = 0;
					for (uint32_t i = 0; i < n; ++i) cnt[tgTmp[i].end_pos + 1]++;
					for (uint32_t i = 1; i <= rawLen + 1; ++i) cnt[i] += cnt[i - 1];
					for (uint32_t i = 0; i < n; ++i) tgRemap[i] = cnt[tgTmp[i].end_pos]++;
					for (uint32_t i = 0; i < n; ++i)
					{
						DTypoNode nd = tgTmp[i];
						const uint32_t ni = tgRemap[i];
						nd.prev = ni - tgRemap[nd.prev];
						if (nd.sibling) nd.sibling = tgRemap[nd.sibling] - ni;
						tg[ni] = nd;
					}
				}
			}
			__syncwarp();
			nTg = __shfl_sync(FULL, n, 0);
			err = __shfl_sync(FULL, err, 0);
		}

		// progressNode (KTrie.cpp:998-1412) for one (graph node, incoming state); returns false when the state dies
		__device__ bool progressTypoNode(const DTypoNode& prevT, const DTypoNode& tn, const DTypoState& state, DTypoState& outState)
		{
			const uint32_t opt = bv.match_options;
			const float typoCost = state.acc_cost + tn.typo_cost;
			if (typoCost > bv.typo.threshold) return false;
			const uint16_t* form = tn.from_pool ? bv.typo.pool + tn.off : raw + tn.off;
			const uint32_t formSize = tn.len;
			uint32_t prevChr = state.last_chr;
			uint32_t lastChrType = T_unknown, lastScriptType = 0;
			if (prevChr) { const uint32_t a = chrAttr(c_m, prevChr); lastChrType = attrCls(a); lastScriptType = attrScript(a); }
			uint32_t specialStartNsPos = state.special_start, unkFormStartNsPos = state.unk_form_start, lastSpaceBoundaryNsPos = state.last_space_boundary;
			uint32_t minFormLen = state.min_form_len;
			int32_t startPosOffset = state.start_pos_offset;
			if (tn.typo_cost > 0.f) startPosOffset += (int32_t)formSize - (int32_t)(tn.end_pos - prevT.end_pos);
			int32_t curNode = state.node;
			for (uint32_t j = 0; j < formSize; ++j)
			{
				const uint32_t c = form[j];
				uint32_t c32 = c;
				if (isHighSurrogate(c32) && j + 1 < formSize) c32 = mergeSurrogate(c32, form[j + 1]);
				const uint32_t pos = tn.end_pos + j - formSize;
				int32_t zc = -1;
				if (typoCost == 0.f)
				{
					const bool isInPattern = nextPat != nPats && pos >= pats[nextPat].end - pats[nextPat].len;
					const uint32_t attr = chrAttr(c_m, c32);
					uint32_t chrType = attrCls(attr), scriptType = attrScript(attr);
					if (lastChrType == T_sw && (c32 == 0x200d || (0x1f3fb <= c32 && c32 <= 0x1f3ff) || scriptType == c_m.script_variation_selectors))
					{
						chrType = lastChrType;
						scriptType = lastScriptType;
					}
					if (isDiscontinuous(lastChrType, isInPattern ? (uint32_t)T_unknown : chrType, lastScriptType, scriptType)
						|| lastChrType == T_sso || lastChrType == T_ssc)
					{
						if (lastChrType != T_max && lastChrType != T_unknown && lastChrType != T_ss)
						{
							const bool hj = T_sf <= lastChrType && lastChrType <= T_sw;
							if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, specialStartNsPos, hj);
							insertUnkForm(unkFormStartNsPos, specialStartNsPos, hj);
							specialRunNode(specialStartNsPos, pos, posToNs[pos], lastChrType);
						}
						unkFormStartNsPos = specialStartNsPos;
						specialStartNsPos = posToNs[pos];
						if (T_sf <= lastChrType && lastChrType <= T_sw) lastSpaceBoundaryNsPos = specialStartNsPos;
					}
					else if (chrType == T_max)
					{
						unkFormStartNsPos = specialStartNsPos;
					}
					lastChrType = isInPattern ? (uint32_t)T_unknown : chrType;
					lastScriptType = scriptType;
					if (c32 < 0x10000)
					{
						if (chrType == T_unknown)      // whitespace
						{
							const uint32_t nx = posToNs[pos + 1];
							if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, nx, true);
							insertUnkForm(unkFormStartNsPos, nx, true);
							lastSpaceBoundaryNsPos = specialStartNsPos = unkFormStartNsPos = nx;
							prevChr = c32;
							continue;
						}
						if ((opt & MATCH_ZCODA) && isHangulCoda(c) && (pos + 1 >= rawLen || !isHangulSyllable(raw[pos + 1])))
						{
							if (isZFollowable(posToNs[pos]) & FF_ZCODA) zc = (int32_t)(c_m.default_tag_size + (c - 0x11A8) - 1);
						}
						else if ((opt & (MATCH_SPLIT_SAISIOT | MATCH_MERGE_SAISIOT)) && c == 0x11BA && pos + 1 < rawLen && isHangulSyllable(raw[pos + 1]))
						{
							if (isZFollowable(posToNs[pos]) & FF_ZSIOT) zc = (int32_t)(c_m.default_tag_size + (0x11BA - 0x11A8) - 1);
						}
					}
				}
				else
				{
					if (c32 < 0x10000 && isSpaceChr(c_m, (uint16_t)c32))
					{
						lastSpaceBoundaryNsPos = specialStartNsPos = unkFormStartNsPos = posToNs[pos + 1];
						prevChr = c32;
						continue;
					}
				}
				if (tn.typo_cost == 0.f && nextPat != nPats)
				{
					const uint32_t currentEnd = pos + (c32 >= 0x10000 ? 2 : 1);
					while (nextPat != nPats && pats[nextPat].end == currentEnd)
					{
						const DPattern np = pats[nextPat];
						const uint32_t matchedStart = np.end - np.len;
						const bool hj = T_w_url <= np.tag && np.tag <= T_w_emoji;
						const uint32_t ms = posToNs[matchedStart];
						if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, ms, hj);
						insertUnkForm(unkFormStartNsPos, ms, hj);
						appendNewNode(ms, posToNs[np.end], c_m.trie_nodes[np.tag].value, startOffset + matchedStart, np.len, 0.f, 0);
						++nextPat;
					}
				}
				if (c32 >= 0x10000)
				{
					++j;
					prevChr = c32;
					continue;
				}
				prevChr = c32;

				if (minFormLen > 0 || tn.typo_cost > 0.f) ++minFormLen;
				int32_t nextNode = nextOpt(curNode, c);
				while (nextNode < 0)
				{
					const int32_t fl = c_m.trie_nodes[curNode].fail;
					if (!fl) { curNode = -1; break; }
					curNode += fl;
					nextNode = nextOpt(curNode, c);
				}
				const uint32_t endPosition = posToNs[pos + 1];
				// the built-in z-coda / saisiot form is pushed before the trie candidates of this position (1126-1135) -> flushed first
				if (zc >= 0) flushCandidate(zc, endPosition, unkFormStartNsPos, lastSpaceBoundaryNsPos, startPosOffset, typoCost);
				if (nextNode >= 0)
				{
					curNode = nextNode;
					// with a typo only the forms that cover the whole replaced segment are collected, at its last character (1291-1311)
					if ((tn.typo_cost == 0.f || j == formSize - 1) && !(typoCost > 0.f && c_m.trie_nodes[curNode].depth < minFormLen))
					{
						for (int32_t sub = curNode; ; )
						{
							const kb2_trie_node sn = c_m.trie_nodes[sub];
							if (sn.value == KB2_TRIE_NONE) break;
							if (sn.value != KB2_TRIE_SUBMATCH)
							{
								if ((uint32_t)c_m.forms_raw[sn.value].str_len < minFormLen) break;
								flushCandidate(sn.value, endPosition, unkFormStartNsPos, lastSpaceBoundaryNsPos, startPosOffset, typoCost);
							}
							if (!sn.fail) break;
							sub += sn.fail;
						}
					}
				}
				else
				{
					if (typoCost == 0.f) curNode = 0;
					else return false;
				}
			}
			if (typoCost == 0.f && lastChrType != T_max && lastChrType != T_unknown && lastChrType != T_ss)
			{
				const bool hj = T_sf <= lastChrType && lastChrType <= T_sw;
				if (lastSpaceBoundaryNsPos < unkFormStartNsPos) insertUnkForm(lastSpaceBoundaryNsPos, specialStartNsPos, hj);
				insertUnkForm(unkFormStartNsPos, specialStartNsPos, hj);
				specialRunNode(specialStartNsPos, tn.end_pos, posToNs[tn.end_pos], lastChrType);
				unkFormStartNsPos = specialStartNsPos;
				if (hj) lastSpaceBoundaryNsPos = posToNs[tn.end_pos];
			}
			if (typoCost > 0.f && c_m.trie_nodes[curNode].depth < minFormLen) return false;      // early pruning (1404)
			outState.node = curNode; outState.acc_cost = typoCost; outState.min_form_len = minFormLen; outState.start_pos_offset = startPosOffset;
			outState.special_start = specialStartNsPos; outState.unk_form_start = unkFormStartNsPos; outState.last_space_boundary = lastSpaceBoundaryNsPos;
			outState.last_chr = prevChr;
			return true;
		}

		// Splitter::search (KTrie.cpp:1414-1452) over tg[0, nTg)
		__device__ __noinline__ void searchTypo()
		{
			const uint32_t totEndPos = nsToPos[nNs - 1] + 1;
			uint32_t nStates = 0;
			if (lane == 0)
			{
				DTypoState s0; s0.node = 0; s0.acc_cost = 0.f; s0.min_form_len = 0; s0.start_pos_offset = 0;
				s0.special_start = 0; s0.unk_form_start = 0; s0.last_space_boundary = 0; s0.last_chr = 0;
				tgStates[0] = s0;
				tgRange[0] = make_uint2(0, 1);
			}
			nStates = 1;
			__syncwarp();
			for (uint32_t i = 1; i < nTg && !err; ++i)
			{
				const DTypoNode tn = tg[i];
				const uint32_t first = nStates;
				if (tn.prev)
				{
					for (uint32_t p = i - tn.prev; !err; )
					{
						const DTypoNode pt = tg[p];
						const uint2 r = tgRange[p];
						for (uint32_t si = r.x; si < r.y; ++si)
						{
							const DTypoState st = tgStates[si];
							DTypoState ns;
							if (!progressTypoNode(pt, tn, st, ns)) continue;
							if (err) break;
							if (nStates >= tgStateCap) { err = ST_TYPO_OVERFLOW; break; }
							if (lane == 0) tgStates[nStates] = ns;
							++nStates;
							__syncwarp();
						}
						if (!pt.sibling) break;
						p += pt.sibling;
					}
				}
				if (lane == 0) tgRange[i] = make_uint2(first, nStates);
				__syncwarp();
				if (tn.typo_cost == 0.f && tn.end_pos == totEndPos)
				{
					for (uint32_t si = first; si < nStates; ++si)
					{
						const DTypoState st = tgStates[si];
						if (st.last_space_boundary < st.unk_form_start) insertUnkForm(st.last_space_boundary, posToNs[totEndPos], true);
						insertUnkForm(st.unk_form_start, posToNs[totEndPos], true);
					}
				}
			}
			// EOS node, as search()
			appendNewNode(nNs, nNs + 1, -1, 0, 0, 0.f, 0);
			if (lane == 0) out[nOut - 1].end_pos = nNs;
			__syncwarp();
		}

		// ---- removeUnconnected + writeResult (240-299, 1454-1464), see DESIGN.md "lattice ordering" ------
		// A node is connected iff it is the last node or it ends at a position where a connected node starts;
		// all nodes ending at one position share that fate, so the stable sort by (connected, endPos) equals a
		// counting sort by endPos over the connected nodes and the relative links become closed forms:
		//   prev' = newIndex - firstIndexOf(endPos == startPos),  sibling' = 1 while the next node ends at the same position.
		__device__ uint32_t finalize(DNode* dst, uint32_t dstCap, uint32_t* newIndex, uint32_t stopPos)
		{
			uint32_t* needed = ctr;                     // [nNs + 2] : 0/1 flags, later counts / bases
			for (uint32_t i = lane; i < nNs + 2; i += KB_W) needed[i] = 0;
			__syncwarp();
			const uint32_t lastId = nOut - 1;
			const DNode lastNode = out[lastId];
			if (lane == 0) needed[lastNode.start_pos] = 1;
			__syncwarp();
			// positions descending: mark the start positions of every node that ends at a needed position
			for (int32_t p = (int32_t)lastNode.start_pos; p >= 0; --p)
			{
				if (!needed[p]) continue;
				const uint2 e = endPosMap[p];
				if (e.x == NPOS) continue;
				for (uint32_t base = e.x; base < e.y; base += KB_W)
				{
					const uint32_t i = base + lane;
					if (i < e.y)
					{
						const DNode g = out[i];
						if (g.end_pos == (uint32_t)p && i != lastId) needed[g.start_pos] = 1;    // benign same-value races
					}
				}
				__syncwarp();
			}
			// count connected nodes per end position (the last node is handled separately: it is always last)
			uint32_t* cnt = posToNs;                     // posToNs is dead after search(); [nNs + 2]
			for (uint32_t i = lane; i < nNs + 2; i += KB_W) cnt[i] = 0;
			__syncwarp();
			for (uint32_t base = 0; base < lastId; base += KB_W)
			{
				const uint32_t i = base + lane;
				if (i < lastId)
				{
					const DNode g = out[i];
					if (needed[g.end_pos] && g.end_pos <= nNs) atomicAdd(&cnt[g.end_pos], 1u);
				}
			}
			__syncwarp();
			// exclusive prefix over positions -> base index of each end position
			uint32_t running = 0;
			for (uint32_t base = 0; base < nNs + 1; base += KB_W)
			{
				const uint32_t i = base + lane;
				const uint32_t v = i < nNs + 1 ? cnt[i] : 0;
				uint32_t incl = v;
				for (int d = 1; d < (int)KB_W; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (lane >= (uint32_t)d) incl += t; }
				if (i < nNs + 1) cnt[i] = running + incl - v;
				running += __shfl_sync(FULL, incl, KB_W - 1);
			}
			const uint32_t connectedCnt = running + 1;    // + the last node
			__syncwarp();
			if (connectedCnt > dstCap) { err = ST_NODE_OVERFLOW; return 0; }
			// `fill[p]` = next free slot among the nodes ending at p, in original index order
			uint32_t* fill = needed;                      // reuse: store base+taken; we still need "needed" -> encode: needed stays in high bit
			// copy needed flag into bit 31 of cnt-based fill array
			for (uint32_t i = lane; i < nNs + 1; i += KB_W) fill[i] = (needed[i] ? 0x80000000u : 0u) | cnt[i];
			__syncwarp();
			for (uint32_t base = 0; base < lastId; base += KB_W)
			{
				const uint32_t i = base + lane;
				DNode g;
				bool conn = false;
				if (i < lastId)
				{
					g = out[i];
					conn = g.end_pos <= nNs && (fill[g.end_pos] & 0x80000000u);
				}
				const uint32_t key = conn ? g.end_pos : (0xFFFF0000u + lane);
				const unsigned grp = __match_any_sync(FULL, key);
				uint32_t slot = 0;
				if (conn)
				{
					const uint32_t first = fill[g.end_pos] & 0x7FFFFFFFu;
					slot = first + __popc(grp & ((1u << lane) - 1));
				}
				__syncwarp();
				if (conn && (grp & ((1u << lane) - 1)) == 0) fill[g.end_pos] += __popc(grp);
				__syncwarp();
				if (i < lastId) newIndex[i] = conn ? slot : NPOS;
				if (conn)
				{
					DNode o = g;
					const uint32_t groupBase = cnt[g.end_pos];
					const uint32_t groupEnd = g.end_pos + 1 <= nNs ? cnt[g.end_pos + 1] : running;
					o.prev = slot == 0 ? 0 : (uint16_t)(slot - cnt[g.start_pos]);
					o.sibling = (slot + 1 < groupEnd) ? 1 : 0;
					(void)groupBase;
					if (slot != 0)
					{
						o.start_pos = nsToPos[g.start_pos] + startOffset;
						o.end_pos = nsToPos[g.end_pos - 1] + 1 + startOffset;
					}
					dst[slot] = o;
				}
			}
			if (lane == 0)
			{
				DNode o = lastNode;
				const uint32_t slot = connectedCnt - 1;
				o.prev = (uint16_t)(slot - cnt[lastNode.start_pos]);
				o.sibling = 0;
				o.start_pos = o.end_pos = startOffset + stopPos;
				dst[slot] = o;
				newIndex[lastId] = slot;
			}
			__syncwarp();
			return connectedCnt;
		}
	};

	// ------------------------------------------------------------------------------------------------
	static __device__ void lattice_sentence(const BatchView& bv, const uint32_t slot, const uint32_t lane)
	{
		const uint32_t s = bv.order[slot];

		const uint32_t t0 = bv.text_off[s], t1 = bv.text_off[s + 1];
		const uint32_t n = t1 - t0;
		const uint32_t W = 2 * n + 4;
		const size_t wbase = 2 * (size_t)t0 + 4 * (size_t)s;
		const size_t nbase = (size_t)bv.nodes_per_unit * wbase;
		const uint32_t nodeCap = bv.nodes_per_unit * W;
		const uint16_t* text = bv.text + t0;
		uint16_t* norm = bv.norm + wbase;
		uint32_t* posTable = bv.pos_table + t0 + s;

		Builder b{ bv, lane };
		b.s = s; b.norm = norm; b.W = W;

		// ---- normalizeHangulWithPosition (StrUtils.h:493-520), lane-parallel with a ballot prefix
		uint32_t outPos = 0;
		for (uint32_t base = 0; base < n; base += KB_W)
		{
			const uint32_t i = base + lane;
			uint32_t c = i < n ? text[i] : 0;
			if (c == 0xB42C) c = 0xB410;
			uint32_t coda = 0;
			if (0xAC00 <= c && c < 0xD7A4) coda = (c - 0xAC00) % 28;
			const unsigned two = __ballot_sync(FULL, i < n && coda != 0);
			const uint32_t before = __popc(two & ((1u << lane) - 1));
			const uint32_t o = outPos + lane + before;
			if (i < n)
			{
				posTable[i] = o;
				norm[o] = (uint16_t)(c - coda);
				if (coda) norm[o + 1] = (uint16_t)(coda + 0x11A7);
			}
			const uint32_t valid = min(KB_W, n - base);
			outPos += valid + __popc(two);
		}
		const uint32_t normLen = outPos;
		if (lane == 0) { posTable[n] = normLen; norm[normLen] = 0; norm[normLen + 1] = 0; bv.norm_len[s] = normLen; }
		__syncwarp();
		// ---- normalizeCoda (StrUtils.h:637-710): it[-1] is rewritten from the ORIGINAL pair (before = *it of the
		// previous iteration, read before any write to it), so positions are independent.
		if (bv.match_options & MATCH_NORMALIZE_CODA)
		{
			const uint32_t codaToOnset[27] = {
				0x3131, 0x3131, 0x3145, 0x3134, 0x3148, 0x314E, 0x3137, 0x3139, 0x3131, 0x3141, 0x3142, 0x3145, 0x314C, 0x314D,
				0x314E, 0x3141, 0x3142, 0x3145, 0x3145, 0x3145, 0x3147, 0x3148, 0x314A, 0x314B, 0x314C, 0x314D, 0x314E };
			const uint32_t codaConv[27] = {
				0, 0x11A8, 0x11A8, 0, 0x11AB, 0x11AB, 0, 0, 0x11AF, 0x11AF, 0x11AF, 0x11AF, 0x11AF, 0x11AF,
				0x11AF, 0, 0, 0x11B8, 0, 0x11BA, 0, 0, 0, 0, 0, 0, 0 };
			for (uint32_t base = 1; base < normLen; base += KB_W)
			{
				const uint32_t i = base + lane;
				uint32_t nv = 0; bool wr = false;
				if (i < normLen)
				{
					const uint32_t before = norm[i - 1], cur = norm[i];
					if (0x11A8 <= before && before <= 0x11C2)
					{
						const uint32_t off = before - 0x11A8;
						if (cur == codaToOnset[off]) { wr = true; nv = codaConv[off] ? codaConv[off] : cur; }
					}
				}
				__syncwarp();
				if (wr) norm[i - 1] = (uint16_t)nv;
				__syncwarp();
			}
		}
		__syncwarp();

		// ---- chunk loop of Kiwi::analyze (src/Kiwi.cpp:1095-1119)
		DChunk* chunks = bv.chunks + (wbase >> 2) + 2 * (size_t)s;
		const uint32_t chunkCap = (W >> 2) + 2;
		uint32_t nChunks = 0, nodeOff = 0, splitEnd = 0;
		b.nsToPos = bv.ns_to_pos + wbase; b.posToNs = bv.pos_to_ns + wbase; b.endPosMap = bv.end_pos_map + wbase;
		b.ctr = bv.ctr + wbase; b.pats = bv.patterns + wbase;
		b.out = bv.build_nodes + nbase; b.outCap = nodeCap;
		const bool typo = bv.typo.nodes != nullptr;
		if (typo)
		{
			const size_t gbase = (size_t)bv.typo.graph_per_unit * wbase, sbase = (size_t)bv.typo.states_per_unit * wbase;
			b.tgCap = bv.typo.graph_per_unit * W; b.tgStateCap = bv.typo.states_per_unit * W;
			b.tgTmp = bv.typo.tmp + gbase; b.tg = bv.typo.graph + gbase; b.tgRemap = bv.typo.remap + gbase; b.tgRange = bv.typo.state_range + gbase;
			b.tgMatches = bv.typo.matches + gbase; b.tgStates = bv.typo.states + sbase;
		}
		DNode* finalNodes = bv.nodes + nbase;
		uint32_t* newIndex = bv.new_index + nbase;
		while (splitEnd < normLen && !b.err)
		{
			const uint16_t* str = norm + splitEnd;
			const uint32_t len = normLen - splitEnd;
			b.startOffset = splitEnd;
			uint32_t stopPos = b.preparePattern(str, len);
			if (b.nNs == 0)
			{
				while (stopPos < len && isSpaceChr(c_m, str[stopPos])) ++stopPos;
				splitEnd += stopPos;
				continue;       // 2-node graph: skipped by the caller (Kiwi.cpp:1119)
			}
			b.raw = str; b.rawLen = stopPos;
			if (typo)
			{
				b.genTypoGraph();      // uses endPosMap / ctr as scratch: before they are initialised for the walk
				if (b.err) break;
			}
			for (uint32_t i = lane; i < b.nNs + 1; i += KB_W) b.endPosMap[i] = make_uint2(NPOS, NPOS);
			__syncwarp();
			if (lane == 0)
			{
				b.endPosMap[0] = make_uint2(0, 1);
				DNode bos; bos.form = -1; bos.uform_off = 0; bos.uform_len = 0; bos.start_pos = 0; bos.end_pos = 0;
				bos.prev = 0; bos.sibling = 0; bos.space_errors = 0; bos.reserved = 0; bos.typo_cost = 0.f;
				b.out[0] = bos;
			}
			b.nOut = 1; b.lastEndPos = 0;
			__syncwarp();
			if (typo) b.searchTypo();
			else b.search();
			if (b.err) break;
			const uint32_t cnt = b.finalize(finalNodes + nodeOff, nodeCap - nodeOff, newIndex, stopPos);
			if (b.err) break;
			if (cnt > 2)
			{
				if (nChunks >= chunkCap) { b.err = ST_CHUNK_OVERFLOW; break; }
				if (lane == 0) chunks[nChunks] = DChunk{ splitEnd, splitEnd + stopPos, nodeOff, cnt };
				++nChunks;
				nodeOff += cnt;
			}
			splitEnd += stopPos;
			__syncwarp();
		}
		if (lane == 0) { bv.n_chunks[s] = nChunks; bv.status[s] = b.err; }
	}

#ifndef KB_HOSTSIM
	__global__ void __launch_bounds__(128) lattice_kernel(const BatchView bv)
	{
		const uint32_t warpsPerBlock = blockDim.x >> 5;
		const uint32_t slot = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5);
		if (slot >= bv.n_sent) return;
		lattice_sentence(bv, slot, threadIdx.x & 31);
	}
#endif

	cudaError_t set_model_lattice(const DevModel& m) { return cudaMemcpyToSymbol(c_m, &m, sizeof(DevModel)); }

	cudaError_t launch_lattice(const DevModel&, const BatchView& bv, cudaStream_t stream)
	{
		if (bv.n_sent == 0) return cudaSuccess;
#if defined(KB_HOSTSIM) && KB_HOSTSIM == 32
		(void)stream;
		for (uint32_t slot = 0; slot < bv.n_sent; ++slot) simt::launch(1, 32, [&] { lattice_sentence(bv, slot, simt::lane); });
		return cudaSuccess;
#elif defined(KB_HOSTSIM)
		for (uint32_t slot = 0; slot < bv.n_sent; ++slot) lattice_sentence(bv, slot, 0);
		(void)stream;
		return cudaSuccess;
#else
		const uint32_t warpsPerBlock = 4;
		const uint32_t blocks = (bv.n_sent + warpsPerBlock - 1) / warpsPerBlock;
		lattice_kernel<<<blocks, warpsPerBlock * 32, 0, stream>>>(bv);
		return cudaGetLastError();
#endif
	}
}
