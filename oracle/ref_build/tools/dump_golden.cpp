// Golden-vector dumper: runs the UNMODIFIED reference (oracle/_ref/libkiwi_ref.so) on UTF-8 input lines and
// writes (A) the public result of kiwi::Kiwi::analyze(line, topN=1, AnalyzeOption{}) and (B) the stage-level
// intermediates of the same call, obtained by driving the reference's own seams exactly as
// Kiwi::analyze does (src/Kiwi.cpp:1095-1141): FnSplitByTrie -> KGraphNode[] per chunk, FnFindBestPath ->
// PathResult[] per chunk.  TEST INFRASTRUCTURE; compiled with -fno-access-control to reach the private
// function-pointer slots (include/kiwi/Kiwi.h:204-208).
//
// Output (text, floats as C99 hex so they round-trip bit-exactly):
//   S <idx> <nTokens> <score> <nChunks> <normLen>
//   T <morphId> <tag> <position> <length> <wordScore> <wordPosition>,<sentPosition>,<lineNumber>,<subSentPosition>,<pairedToken> <form utf-8>   x nTokens
//   C <chunkIdx> <startOffset> <endOffset> <nNodes> <nPaths>
//   N <formIdx|-1> <uformOff|-1> <uformLen> <prev> <sibling> <startPos> <endPos> <spaceErrors> <typoCost>   x nNodes
//   P <score> <prevState> <curState> <nTok>
//   K <morphId> <begin> <end> <wordScore> <nodeId> <hasStr>                x nTok
// The model type follows the environment variable KB_MODEL_TYPE (knlm, default, cong or sbg).  KB_TYPO=basic analyses with the
// default basic typo set prepared as the evaluator's `--typo` does (tools/Evaluator.cpp:79-144): option.typoTransformer =
// getDefaultTypoSet(basicTypoSet).prepare(true), typoThreshold 2.5, KiwiConfig::typoCostWeight 6 (BASELINE.json config 4).
// usage: dump_golden <model_dir> <input.txt> <out.txt> [maxLines]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <thread>
#include <unordered_set>
#include <algorithm>
#include <kiwi/Kiwi.h>
#include <kiwi/TypoTransformer.h>
#include "StrUtils.h"
#include "KTrie.h"
#include "PathEvaluator.h"

using namespace kiwi;

// UTF-8 for the dump; unpaired surrogates (edge-case inputs) become U+FFFD instead of an exception
static std::string toUtf8Lenient(const std::u16string& s)
{
	std::string out;
	for (size_t i = 0; i < s.size(); ++i)
	{
		uint32_t c = s[i];
		if (0xD800 <= c && c < 0xDC00 && i + 1 < s.size() && 0xDC00 <= s[i + 1] && s[i + 1] < 0xE000) { c = 0x10000 + ((c - 0xD800) << 10) + (s[i + 1] - 0xDC00); ++i; }
		else if (0xD800 <= c && c < 0xE000) c = 0xFFFD;
		if (c < 0x80) out.push_back((char)c);
		else if (c < 0x800) { out.push_back((char)(0xC0 | (c >> 6))); out.push_back((char)(0x80 | (c & 63))); }
		else if (c < 0x10000) { out.push_back((char)(0xE0 | (c >> 12))); out.push_back((char)(0x80 | ((c >> 6) & 63))); out.push_back((char)(0x80 | (c & 63))); }
		else { out.push_back((char)(0xF0 | (c >> 18))); out.push_back((char)(0x80 | ((c >> 12) & 63))); out.push_back((char)(0x80 | ((c >> 6) & 63))); out.push_back((char)(0x80 | (c & 63))); }
	}
	return out;
}

int main(int argc, char** argv)
{
	if (argc < 4) { std::cerr << "usage: dump_golden <model_dir> <input.txt> <out.txt> [maxLines]\n"; return 2; }
	if (!getenv("KIWI_ARCH_TYPE")) setenv("KIWI_ARCH_TYPE", "avx2", 1);
	const size_t maxLines = argc > 4 ? std::stoul(argv[4]) : (size_t)-1;
	try
	{
		const char* mt = getenv("KB_MODEL_TYPE");
		KiwiBuilder kb{ argv[1], 1, BuildOption::default_, (mt && std::string{ mt } == "cong") ? ModelType::cong : (mt && std::string{ mt } == "sbg") ? ModelType::sbg : ModelType::knlm };
		Kiwi kw = kb.build();
		std::ifstream ifs{ argv[2] };
		FILE* fo = std::fopen(argv[3], "w");
		std::string line;
		size_t idx = 0;
		AnalyzeOption option;
		if (getenv("KB_OPEN_ENDING")) option.openEnding = true;      // (vectors open_<name>: AnalyzeOption::openEnding)
		// KB_BLOCKLIST="form/TAG;form/TAG": AnalyzeOption::blocklist built like kiwi_morphset_add does (capi/kiwi_c.cpp:1796-1811); the resolved
		// morpheme ids go to stdout as one line `BLOCKLIST id id ...`
		std::unordered_set<const Morpheme*> blockSet;
		if (const char* bl = getenv("KB_BLOCKLIST"))
		{
			std::string spec{ bl }; size_t pos = 0;
			while (pos < spec.size())
			{
				size_t e = spec.find(';', pos); if (e == spec.npos) e = spec.size();
				const std::string item = spec.substr(pos, e - pos); pos = e + 1;
				const size_t sl = item.rfind('/');
				const std::u16string form = utf8To16(item.substr(0, sl));
				const POSTag tag = sl == item.npos ? POSTag::unknown : toPOSTag(utf8To16(item.substr(sl + 1)));
				for (auto* m : kw.findMorphemes(form, tag)) blockSet.insert(m);
			}
			std::vector<size_t> ids; for (auto* m : blockSet) ids.push_back(kw.morphToId(m));
			std::sort(ids.begin(), ids.end());
			std::printf("BLOCKLIST"); for (auto i : ids) std::printf(" %zu", i); std::printf("\n");
			option.blocklist = &blockSet;
		}
		PreparedTypoTransformer ptt;
		if (const char* ty = getenv("KB_TYPO"))
		{
			if (std::string{ ty } != "basic") throw std::runtime_error{ "KB_TYPO: only `basic` is known" };
			ptt = getDefaultTypoSet(DefaultTypoSet::basicTypoSet).prepare(true);
			option.typoTransformer = &ptt;
		}
		const KiwiConfig config = kw.globalConfig;
		// KB_FRESH_THREAD=1: every sentence is analysed on a thread of its own.  The reference keeps its `top1` path container (an
		// std::unordered_set, BestPathContainer.hpp:229-276) in thread-local storage; its bucket count - and with it the order in which a
		// later sentence's paths come out - depends on the sentences the thread has analysed before.  A fresh thread gives every sentence the
		// state of a thread's first analysis, which is the definition the oracle and the CUDA path use.
		const bool freshThread = getenv("KB_FRESH_THREAD") != nullptr;
		while (std::getline(ifs, line) && idx < maxLines)
		{
			auto body = [&]()
			{
			if (!line.empty() && line.back() == '\r') line.pop_back();
			const auto tab = line.find('\t');
			if (tab != line.npos) line = line.substr(0, tab);
			const std::u16string str = utf8To16(line);

			// (A) public API
			auto res = kw.analyze(str, 1, option);
			const auto& tokens = res[0].first;

			// (B) stage level, mirroring src/Kiwi.cpp:1028-1141 (no pretokenized spans, no typo transformer)
			KString normalizedStr;
			Vector<uint32_t> positionTable;
			normalizeHangulWithPosition(str.begin(), str.end(), std::back_inserter(normalizedStr), std::back_inserter(positionTable));
			if (!!(option.match & Match::normalizeCoda)) normalizeCoda(normalizedStr.begin(), normalizedStr.end());

			struct Chunk { size_t start, end; Vector<KGraphNode> nodes; Vector<PathResult> paths; };
			std::vector<Chunk> chunks;
			Vector<SpecialState> spStatesByRet;
			// the carried special-state set only depends on scores/states, which we recompute like insertPathIntoResults
			std::vector<std::pair<float, uint8_t>> retStates; // (accumulated score, state)
			size_t splitEnd = 0;
			const PretokenizedSpanGroup::Span* ptFirst = nullptr;
			while (splitEnd < normalizedStr.size())
			{
				Chunk ch;
				ch.start = splitEnd;
				splitEnd = (*reinterpret_cast<FnSplitByTrie>(kw.dfSplitByTrie))(
					ch.nodes, kw.forms.data(), kw.typoPtrs.data(), kw.formTrie,
					U16StringView{ normalizedStr.data() + splitEnd, normalizedStr.size() - splitEnd },
					splitEnd, option.match, option.allowedDialects,
					config.maxUnkFormSize, config.maxUnkFormSizeFollowedByJClass, config.spaceTolerance,
					option.typoTransformer, option.typoThreshold, kw.continualTypoCost, kw.lengtheningTypoCost,
					ptFirst, ptFirst);
				ch.end = splitEnd;
				if (ch.nodes.size() > 2)
				{
					ch.paths = (*reinterpret_cast<FnFindBestPath>(kw.dfFindBestPath))(
						&kw, config, spStatesByRet, normalizedStr, ch.nodes.data(), ch.nodes.size(), 1,
						(size_t)(option.match & Match::oovMask),
						option.openEnding && splitEnd == normalizedStr.size(),
						!!(option.match & Match::splitComplex), !!(option.match & Match::splitSaisiot), !!(option.match & Match::mergeSaisiot),
						option.blocklist, option.allowedDialects, option.dialectCost, nullptr);

					// evolve spStatesByRet with the selection rules of insertPathIntoResults (src/Kiwi.cpp:629-782), topN = 1
					const auto& pathes = ch.paths;
					std::vector<size_t> parentMap;
					std::vector<std::pair<float, uint8_t>> ret = retStates;
					if (ret.empty())
					{
						const size_t n = std::min(pathes.size(), (size_t)2);
						ret.assign(n, std::make_pair(0.f, (uint8_t)0));
						for (size_t i = 0; i < n; ++i) parentMap.push_back(i);
					}
					else
					{
						UnorderedMap<uint8_t, uint32_t> prevParents;
						Vector<uint8_t> selected(pathes.size());
						for (size_t i = 0; i < ret.size(); ++i)
						{
							auto pred = [&](const PathResult& p) { return (uint8_t)p.prevState == ret[i].second; };
							size_t parent = std::find_if(pathes.begin() + prevParents[ret[i].second], pathes.end(), pred) - pathes.begin();
							if (parent >= pathes.size() && prevParents[ret[i].second]) parent = std::find_if(pathes.begin(), pathes.end(), pred) - pathes.begin();
							parentMap.push_back(parent);
							if (parent < pathes.size()) { selected[parent] = 1; prevParents[ret[i].second] = parent + 1; }
						}
						const size_t origSize = ret.size();
						for (size_t i = 0; i < pathes.size(); ++i)
						{
							if (selected[i]) continue;
							size_t parent = 0;
							for (; parent < ret.size(); ++parent) if (ret[parent].second == (uint8_t)pathes[i].prevState) break;
							if (parent < ret.size()) { ret.push_back(ret[parent]); parentMap.push_back(i); }
						}
						(void)origSize;
					}
					UnorderedMap<uint8_t, uint32_t> spStateCnt;
					std::vector<std::pair<float, uint8_t>> kept;
					for (size_t i = 0; i < ret.size(); ++i)
					{
						if (!(parentMap[i] < pathes.size() && spStateCnt[pathes[parentMap[i]].curState] < 1)) continue;
						const auto& r = pathes[parentMap[i]];
						kept.emplace_back(ret[i].first + r.score, (uint8_t)r.curState);
						spStateCnt[r.curState]++;
					}
					std::vector<size_t> order(kept.size());
					for (size_t i = 0; i < order.size(); ++i) order[i] = i;
					std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return kept[a].first > kept[b].first; });
					retStates.clear();
					spStatesByRet.clear();
					for (size_t i = 0; i < std::min((size_t)2, kept.size()); ++i)
					{
						retStates.push_back(kept[order[i]]);
						SpecialState s;
						reinterpret_cast<uint8_t&>(s) = kept[order[i]].second;
						spStatesByRet.push_back(s);
					}
				}
				chunks.emplace_back(std::move(ch));
			}

			std::fprintf(fo, "S %zu %zu %a %zu %zu\n", idx, tokens.size(), res[0].second, chunks.size(), normalizedStr.size());
			for (auto& t : tokens)
			{
				std::fprintf(fo, "T %zu %u %u %u %a %u,%u,%u,%u,%d %s\n", t.morph ? kw.morphToId(t.morph) : (size_t)-1, (unsigned)t.tag, t.position, (unsigned)t.length, t.score,
					(unsigned)t.wordPosition, (unsigned)t.sentPosition, (unsigned)t.lineNumber, (unsigned)t.subSentPosition, (int)t.pairedToken, toUtf8Lenient(t.str).c_str());
			}
			if (getenv("KB_TYPO"))      // typo-tolerant dumps only: TokenInfo::typoCost of every token
			{
				std::fprintf(fo, "Y");
				for (auto& t : tokens) std::fprintf(fo, " %a", t.typoCost);
				std::fprintf(fo, "\n");
			}
			for (size_t c = 0; c < chunks.size(); ++c)
			{
				auto& ch = chunks[c];
				std::fprintf(fo, "C %zu %zu %zu %zu %zu\n", c, ch.start, ch.end, ch.nodes.size(), ch.paths.size());
				for (auto& n : ch.nodes)
				{
					const long uoff = n.uform.empty() ? -1 : (long)(n.uform.data() - normalizedStr.data());
					std::fprintf(fo, "N %ld %ld %zu %u %u %u %u %u %a\n",
						n.form ? (long)(n.form - kw.forms.data()) : -1L, uoff, n.uform.size(),
						n.prev, n.sibling, n.startPos, n.endPos, n.spaceErrors, n.typoCost);
				}
				for (auto& p : ch.paths)
				{
					std::fprintf(fo, "P %a %u %u %zu\n", p.score, (unsigned)(uint8_t)p.prevState, (unsigned)(uint8_t)p.curState, p.path.size());
					for (auto& k : p.path)
					{
						std::fprintf(fo, "K %zu %u %u %a %u %d\n", kw.morphToId(k.morph), k.begin, k.end, k.wordScore, k.nodeId, k.str.empty() ? 0 : 1);
					}
				}
			}
			};
			if (freshThread) { std::thread th(body); th.join(); } else body();
			++idx;
		}
		std::fclose(fo);
		std::cerr << "dumped " << idx << " lines\n";
	}
	catch (const std::exception& e)
	{
		std::cerr << "dump_golden failed: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
