// ORACLE (test infrastructure): CPU restatement of PreparedTypoTransformer::generateGraph — the typo DAG over one
// normalised chunk (BASELINE.json config 4) — over a flat typo image (include/kiwi_b200_typo.h).
//   generateGraph            /root/reference/src/TypoTransformer.cpp:810-1039
//   appendNewNode (typo)     src/TypoTransformer.cpp:594-629
//   FrozenTrie walk          include/kiwi/FrozenTrie.h:55-99, src/FrozenTrie.hpp:14-58
//   normalizeHangul          include/kiwi/Utils.h:131-166
// Pinned by the reference's own known-answer test (KiwiTypo.GenerateGraph, test/test_typo.cpp:8-22: 11 nodes) and by
// node-for-node dumps of the unmodified reference (tests/golden/typo_*.golden.txt.gz).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/kiwi_b200_typo.h"
#include "feature.hpp"      // ftVowel (FeatureTestor::isMatched(CondVowel)), isHangulSyllable

namespace orc
{
	struct TypoNode      // TypoGraphNode, include/kiwi/TypoTransformer.h:130-156
	{
		bool fromPool = false; uint32_t off = 0, len = 0;      // form: view into the analysed string or into the replacement pool
		uint32_t endPos = 0; float typoCost = 0; uint32_t prevOffset = 0, siblingOffset = 0; uint8_t continualTypoIdx = 0; uint16_t dialect = 0;
	};

	inline std::u16string normalizeHangulPlain(const std::u16string& s)      // Utils.h:131-166
	{
		std::u16string ret;
		for (char16_t c : s)
		{
			if (c == 0xB42C) c = 0xB410;
			if (0xAC00 <= c && c < 0xD7A4)
			{
				const int coda = (c - 0xAC00) % 28;
				ret.push_back((char16_t)(c - coda));
				if (coda) ret.push_back((char16_t)(coda + 0x11A7));
			}
			else if (!ret.empty() && 0x1100 <= ret.back() && ret.back() < 0x1100 + 19 && 0x1161 <= c && c < 0x1176)
			{
				ret.back() = (char16_t)(0xAC00 + ((ret.back() - 0x1100) * 21 * 28) + ((c - 0x1161) * 28));
			}
			else ret.push_back(c);
		}
		return ret;
	}

	struct TypoImage
	{
		std::vector<char> blob;
		const kb2_typo_header* h = nullptr;
		const kb2_typo_node* nodes = nullptr; const uint16_t* keys = nullptr; const int32_t* diffs = nullptr;
		const kb2_typo_pat* pats = nullptr; const kb2_typo_repl* repls = nullptr; const uint16_t* pool = nullptr;

		void load(const std::string& path)
		{
			FILE* f = std::fopen(path.c_str(), "rb");
			if (!f) throw std::runtime_error("cannot open typo image " + path);
			std::fseek(f, 0, SEEK_END); const long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
			blob.resize((size_t)n);
			if (std::fread(blob.data(), 1, (size_t)n, f) != (size_t)n) { std::fclose(f); throw std::runtime_error("short read"); }
			std::fclose(f);
			h = reinterpret_cast<const kb2_typo_header*>(blob.data());
			if (h->magic != KB2_TYPO_MAGIC) throw std::runtime_error("bad typo image magic");
			size_t o = sizeof(kb2_typo_header);
			auto take = [&](size_t bytes) { o = (o + 15) / 16 * 16; const char* p = blob.data() + o; o += bytes; return p; };
			nodes = reinterpret_cast<const kb2_typo_node*>(take(sizeof(kb2_typo_node) * h->n_nodes));
			keys = reinterpret_cast<const uint16_t*>(take(2 * (size_t)h->n_edges));
			diffs = reinterpret_cast<const int32_t*>(take(4 * (size_t)h->n_edges));
			pats = reinterpret_cast<const kb2_typo_pat*>(take(sizeof(kb2_typo_pat) * h->n_pats));
			repls = reinterpret_cast<const kb2_typo_repl*>(take(sizeof(kb2_typo_repl) * h->n_repls));
			pool = reinterpret_cast<const uint16_t*>(take(2 * (size_t)h->n_pool));
		}

		// Node::nextOpt: child for key c, -1 = none
		int32_t next(int32_t node, uint16_t c) const
		{
			const kb2_typo_node& n = nodes[node];
			const uint16_t* k = keys + n.next_offset;
			const uint16_t* it = std::lower_bound(k, k + n.num_nexts, c);
			if (it == k + n.num_nexts || *it != c) return -1;
			return node + diffs[n.next_offset + (it - k)];
		}
		int32_t fail(int32_t node) const { return nodes[node].fail ? node + nodes[node].fail : -1; }
	};

	enum { TCV_none = 0, TCV_any = 1, TCV_vowel = 2, TCV_continual = 9, TCV_boundary = 10 };      // CondVowel, Types.h:260-273

	struct TypoGraph
	{
		const TypoImage& im;
		explicit TypoGraph(const TypoImage& _im) : im{ _im } {}

		using EndPosMap = std::vector<std::pair<uint32_t, uint32_t>>;
		static constexpr uint32_t npos = 0xFFFFFFFFu;
		static constexpr size_t none = (size_t)-1;

		// appendNewNode, TypoTransformer.cpp:594-629 (prev / sibling offsets stay absolute until the final renumbering)
		static bool append(std::vector<TypoNode>& nodes, EndPosMap& endPosMap, size_t mapOffset, bool fromPool, uint32_t off, uint32_t len,
			size_t startPos, size_t endPos, float cost = 0)
		{
			if (startPos != none && endPosMap[startPos - mapOffset].first == npos) return false;
			const size_t newId = nodes.size();
			TypoNode nn; nn.fromPool = fromPool; nn.off = off; nn.len = len; nn.endPos = (uint32_t)endPos; nn.typoCost = cost;
			nodes.push_back(nn);
			TypoNode& nnode = nodes.back();
			if (startPos == none) nnode.prevOffset = (uint32_t)(newId - 1);
			else nnode.prevOffset = endPosMap[startPos - mapOffset].first;
			if ((size_t)nnode.endPos >= endPosMap.size() + mapOffset) return true;
			auto& slot = endPosMap[nnode.endPos - mapOffset];
			if (slot.first == npos) slot.first = (uint32_t)newId;
			else nodes[slot.second].siblingOffset = (uint32_t)newId;
			slot.second = (uint32_t)newId;
			return true;
		}

		struct Match { size_t endPos; kb2_typo_pat pat; };

		std::vector<TypoNode> generate(const std::u16string& str) const
		{
			const float continualTypoThreshold = im.h->continual_typo_threshold;
			std::vector<TypoNode> temp;
			std::vector<Match> matches;
			std::vector<size_t> breakPoints;
			EndPosMap endPosMap;
			endPosMap.emplace_back(0, 0);
			size_t last = 0;
			{ TypoNode bos; temp.push_back(bos); }

			auto insertBranch = [&]()
			{
				const size_t totStartPos = matches[0].endPos - matches[0].pat.pat_len;
				const size_t totEndPos = matches.back().endPos;
				const auto v = endPosMap.back();
				endPosMap.assign((totEndPos - last) + 1, std::make_pair(npos, npos));
				endPosMap[0] = v;

				breakPoints.clear();
				breakPoints.push_back(totStartPos);
				for (auto& m : matches) breakPoints.push_back(m.endPos);
				breakPoints.push_back(totEndPos);
				std::sort(breakPoints.begin(), breakPoints.end());
				breakPoints.erase(std::unique(breakPoints.begin(), breakPoints.end()), breakPoints.end());

				std::sort(matches.begin(), matches.end(), [](const Match& a, const Match& b) { return a.endPos - a.pat.pat_len < b.endPos - b.pat.pat_len; });

				if (last < totStartPos) append(temp, endPosMap, last, false, (uint32_t)last, (uint32_t)(totStartPos - last), last, totStartPos);
				for (size_t i = 1; i < breakPoints.size(); ++i)
					append(temp, endPosMap, last, false, (uint32_t)breakPoints[i - 1], (uint32_t)(breakPoints[i] - breakPoints[i - 1]), breakPoints[i - 1], breakPoints[i]);

				for (auto& m : matches)
				{
					const size_t e = m.endPos, s = e - m.pat.pat_len;
					std::unordered_map<uint16_t, std::pair<size_t, size_t>> continualIdx;      // first unit of the replacement -> (idx, node)
					for (uint32_t j = 0; j < m.pat.size; ++j)
					{
						const kb2_typo_repl& repl = im.repls[m.pat.repl_off + j];
						if (repl.dialect != 0) continue;      // allowedDialect == standard
						if (repl.left_cond == TCV_vowel)
						{
							if (s == 0 || !isHangulSyllable(str[s - 1])) continue;
						}
						else if (repl.left_cond == TCV_any)
						{
							if (s == 0) continue;
						}
						else if (repl.left_cond == TCV_continual || repl.left_cond == TCV_boundary)
						{
							if (repl.left_cond == TCV_continual && (s == 0 || !isHangulSyllable(str[s - 1]))) continue;
							if (repl.left_cond == TCV_continual && !std::isfinite(continualTypoThreshold)) continue;
							const float scale = repl.left_cond == TCV_continual ? continualTypoThreshold : 1.f;
							auto ins = continualIdx.emplace(im.pool[repl.str_off], std::make_pair(continualIdx.size() + 1, (size_t)0));
							auto& idxAndNode = ins.first->second;
							if (ins.second)
							{
								if (append(temp, endPosMap, last, true, repl.str_off, 1, s, none, repl.cost * scale / 2))
								{
									temp.back().endPos = (uint32_t)e;
									temp.back().continualTypoIdx = (uint8_t)idxAndNode.first;
									temp.back().dialect = repl.dialect;
									idxAndNode.second = temp.size() - 1;
									if (append(temp, endPosMap, last, true, repl.str_off + 1, repl.length - 1, none, e, repl.cost * scale / 2))
									{
										temp.back().prevOffset = (uint32_t)idxAndNode.second;
										temp.back().dialect = repl.dialect;
									}
								}
								else continualIdx.erase(ins.first);
							}
							else
							{
								if (append(temp, endPosMap, last, true, repl.str_off + 1, repl.length - 1, none, e, repl.cost * scale / 2))
								{
									temp.back().prevOffset = (uint32_t)idxAndNode.second;
									temp.back().dialect = repl.dialect;
								}
							}
							continue;
						}
						else
						{
							const u16* b = reinterpret_cast<const u16*>(str.data());
							if (!ftVowel(b, b + s, repl.left_cond)) continue;
						}
						if (append(temp, endPosMap, last, true, repl.str_off, repl.length, s, e, repl.cost)) temp.back().dialect = repl.dialect;
					}
				}
				last = totEndPos;
				matches.clear();
			};

			int32_t node = im.next(0, 0);
			if (node < 0) throw std::runtime_error("typo trie has no start node");
			for (size_t i = 0; i < str.size(); ++i)
			{
				int32_t nnode = im.next(node, str[i]);
				while (nnode < 0)
				{
					node = im.fail(node);
					if (node >= 0) nnode = im.next(node, str[i]);
					else { node = 0; break; }
				}
				if (nnode < 0) continue;
				node = nnode;
				const int32_t v = im.nodes[node].value;
				if (v == -1) continue;
				const size_t endPos = i + 1;
				// the value of a sub-match-only node carries patLength = -1 in the reference (size_t arithmetic wraps); restate that
				const size_t patLen = v >= 0 ? im.pats[v].pat_len : (size_t)(uint32_t)-1;
				const size_t startPos = endPos - patLen;
				if (!matches.empty() && matches.back().endPos < startPos) insertBranch();
				for (int32_t sub = node; sub >= 0; sub = im.fail(sub))
				{
					const int32_t sv = im.nodes[sub].value;
					if (sv == -1) break;
					if (sv == -2) continue;
					matches.push_back(Match{ endPos, im.pats[sv] });
				}
			}
			if (!matches.empty()) insertBranch();
			{
				const auto v = endPosMap.back();
				endPosMap.assign(1, v);
			}
			append(temp, endPosMap, last, false, (uint32_t)last, (uint32_t)(str.size() - last), last, str.size() + 1);
			temp.back().endPos = (uint32_t)str.size();

			std::vector<size_t> sortIdx(temp.size()), reverseIdx(temp.size());
			std::iota(sortIdx.begin(), sortIdx.end(), 0);
			std::stable_sort(sortIdx.begin(), sortIdx.end(), [&](size_t a, size_t b) { return temp[a].endPos < temp[b].endPos; });
			for (size_t i = 0; i < temp.size(); ++i) reverseIdx[sortIdx[i]] = i;
			std::vector<TypoNode> out;
			out.reserve(temp.size());
			for (size_t i = 0; i < temp.size(); ++i)
			{
				out.push_back(temp[sortIdx[i]]);
				TypoNode& n = out.back();
				n.prevOffset = (uint32_t)(i - reverseIdx[n.prevOffset]);
				if (n.siblingOffset != 0) n.siblingOffset = (uint32_t)(reverseIdx[n.siblingOffset] - i);
			}
			return out;
		}
	};
}
