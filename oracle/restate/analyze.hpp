// ORACLE (test infrastructure): CPU restatement of the chunk loop of Kiwi::analyze (topN == 1, default
// AnalyzeOption), /root/reference/src/Kiwi.cpp:1014-1158, and of the parts of insertPathIntoResults
// (src/Kiwi.cpp:615-783) that decide which chunk paths are stitched together and how a PathNode becomes an
// output token (position/length mapping :734-737, script re-tagging :590-605, space pseudo-tokens :700).
#pragma once
#include "viterbi.hpp"

namespace orc
{
	struct Token { uint32_t morph; uint8_t tag; uint32_t position; uint16_t length; float score; };

	struct ChunkDump { size_t start, end; std::vector<LNode> nodes; std::vector<PathResult> paths; };

	struct AnalyzeResult
	{
		std::vector<Token> tokens;
		float score = 0;
		std::vector<ChunkDump> chunks;
		size_t normLen = 0;
	};

	struct Analyzer
	{
		const Image& im;
		Splitter splitter;
		Viterbi viterbi;
		uint32_t matchOptions = (1u << 0) | (1u << 1) | (1u << 2) | (1u << 3) | (1u << 4) | (1u << 5) | (1u << 23) | (1u << 16);   // Match::allWithNormalizing
		bool keepChunks = true;
		WorkCounters work;
		void enableWorkCounters() { splitter.wc = &work; viterbi.wc = &work; viterbi.lm.wc = &work; viterbi.cg.wc = &work; viterbi.sbgm.wc = &work; }

		explicit Analyzer(const Image& _im) : im{ _im }, splitter{ _im }, viterbi{ _im }
		{
			splitter.matchOptions = matchOptions;
			splitter.maxUnkFormSize = _im.h->config.max_unk_form_size;
			splitter.maxUnkFormSizeFollowedByJClass = _im.h->config.max_unk_form_size_followed_by_jclass;
			splitter.spaceTolerance = _im.h->config.space_tolerance;
		}

		struct Ret { std::vector<Token> tokens; float score = 0; };

		void appendTokens(std::vector<Token>& out, const PathResult& r, const std::vector<u16>& norm, const std::vector<uint32_t>& positionTable) const
		{
			for (const auto& s : r.path)
			{
				const u16* str = nullptr; uint32_t strLen = s.strLen;
				if (strLen) str = s.strOff >= 0 ? norm.data() + s.strOff : im.formStr(~s.strOff);
				if (strLen && str[0] == ' ') continue;
				const auto& m = im.morphs[s.morph];
				Token t;
				t.morph = s.morph;
				t.tag = m.tag;
				const size_t beginPos = (std::upper_bound(positionTable.begin(), positionTable.end(), s.begin) - positionTable.begin()) - 1;
				const size_t endPos = std::lower_bound(positionTable.begin(), positionTable.end(), s.end) - positionTable.begin();
				t.position = (uint32_t)beginPos;
				t.length = (uint16_t)(endPos - beginPos);
				t.score = s.wordScore;
				// updateTokenInfoScript, src/Kiwi.cpp:590-605.  info.str = joinHangul(s.str.empty() ? *kform : s.str)
				if (t.tag == T_sl || t.tag == T_sh || t.tag == T_sw || t.tag == T_w_emoji)
				{
					const bool hasKform = m.form_idx >= 0 && im.formLen(m.form_idx) > 0;
					if (!hasKform && strLen)
					{
						uint32_t c = str[0];
						if (isHighSurrogate(c)) c = mergeSurrogate(c, strLen > 1 ? str[1] : 0);
						if (im.script(c) == im.h->script_latin) t.tag = T_sl;
					}
				}
				out.push_back(t);
			}
		}

		bool openEnding = false;
		AnalyzeResult analyze(const u16* text, size_t len)
		{
			AnalyzeResult res;
			std::vector<u16> norm; std::vector<uint32_t> positionTable;
			normalizeHangulWithPosition(text, text + len, norm, positionTable);
			if (matchOptions & (1u << 16)) normalizeCoda(norm);
			res.normLen = norm.size();
			work.sentences++; work.rawUnits += len; work.normUnits += norm.size();
			if (!std::getenv("ORC_KEEP_TOP1_HISTORY")) viterbi.resetHistory();      // every sentence starts with a fresh `top1` container (see viterbi.hpp)

			std::vector<Ret> ret;
			std::vector<uint8_t> spStatesByRet;
			size_t splitEnd = 0;
			while (splitEnd < norm.size())
			{
				ChunkDump ch;
				ch.start = splitEnd;
				splitEnd = splitter.split(ch.nodes, norm.data() + splitEnd, norm.size() - splitEnd, splitEnd);
				ch.end = splitEnd;
				if (ch.nodes.size() > 2)
				{
					// AnalyzeOption::openEnding: no end-of-sentence step on the chunk that ends the text (src/Kiwi.cpp:1122-1131)
					ch.paths = viterbi.findBestPath(spStatesByRet, norm.data(), ch.nodes.data(), ch.nodes.size(), openEnding && splitEnd == norm.size(), matchOptions);
					// insertPathIntoResults, topN == 1
					const auto& pathes = ch.paths;
					std::vector<size_t> parentMap;
					if (ret.empty())
					{
						const size_t n = std::min(pathes.size(), (size_t)2);
						ret.resize(n);
						spStatesByRet.resize(n);
						for (size_t i = 0; i < n; ++i) parentMap.push_back(i);
					}
					else
					{
						uint32_t prevParents[256] = { 0 };
						std::vector<uint8_t> selected(pathes.size());
						for (size_t i = 0; i < ret.size(); ++i)
						{
							const uint8_t st = spStatesByRet[i];
							auto findFrom = [&](size_t from) { size_t k = from; for (; k < pathes.size(); ++k) if (pathes[k].prevState == st) break; return k; };
							size_t parent = findFrom(prevParents[st]);
							if (parent >= pathes.size() && prevParents[st]) parent = findFrom(0);
							parentMap.push_back(parent);
							if (parent < pathes.size()) { selected[parent] = 1; prevParents[st] = (uint32_t)parent + 1; }
						}
						for (size_t i = 0; i < pathes.size(); ++i)
						{
							if (selected[i]) continue;
							size_t parent = std::find(spStatesByRet.begin(), spStatesByRet.end(), pathes[i].prevState) - spStatesByRet.begin();
							if (parent < ret.size())
							{
								ret.push_back(ret[parent]);
								spStatesByRet.push_back(spStatesByRet[parent]);
								parentMap.push_back(i);
							}
							else throw std::runtime_error("unreachable stitch branch");
						}
					}
					uint32_t spStateCnt[256] = { 0 };
					size_t validTarget = 0;
					for (size_t i = 0; i < ret.size(); ++i)
					{
						if (parentMap[i] < pathes.size() && spStateCnt[pathes[parentMap[i]].curState] < 1)
						{
							if (validTarget != i) ret[validTarget] = std::move(ret[i]);
						}
						else continue;
						const auto& r = pathes[parentMap[i]];
						appendTokens(ret[validTarget].tokens, r, norm, positionTable);
						ret[validTarget].score += r.score;
						spStatesByRet[validTarget] = r.curState;
						spStateCnt[r.curState]++;
						validTarget++;
					}
					std::vector<size_t> idx(validTarget);
					for (size_t i = 0; i < validTarget; ++i) idx[i] = i;
					std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return ret[a].score > ret[b].score; });
					std::vector<Ret> sortedRet; std::vector<uint8_t> sortedSp;
					const size_t maxCands = std::min((size_t)2, validTarget);
					for (size_t i = 0; i < maxCands; ++i) { sortedRet.push_back(std::move(ret[idx[i]])); sortedSp.push_back(spStatesByRet[idx[i]]); }
					ret = std::move(sortedRet);
					spStatesByRet = std::move(sortedSp);
				}
				if (keepChunks) res.chunks.push_back(std::move(ch));
			}
			std::sort(ret.begin(), ret.end(), [](const Ret& a, const Ret& b) { return a.score > b.score; });
			if (!ret.empty()) { res.tokens = std::move(ret[0].tokens); res.score = ret[0].score; }
			work.tokens += res.tokens.size();
			return res;
		}
	};
}
