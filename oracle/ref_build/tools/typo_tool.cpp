// Typo-transformer flattening + golden graphs (TEST / FIXTURE INFRASTRUCTURE: links oracle/_ref/libkiwi_ref.so, compiled with
// -fno-access-control to read kiwi::PreparedTypoTransformer's private trie / replacement table,
// include/kiwi/TypoTransformer.h:205-212).
//   typo_tool flatten <set> <out.img>                  set = basic (getDefaultTypoSet(basicTypoSet), prepared like kiwi_typo_prepare
//                                                      does for analysis: prepare(true)) | kat (the three rules of the reference's own
//                                                      test KiwiTypo.GenerateGraph, test/test_typo.cpp:8-22)
//   typo_tool graphs <set> <input.txt> <out.txt> [max]  per input line (first tab column): normalizeHangul, generateGraph, dump
//       G <idx> <normLen> <nNodes>
//       N <endPos> <typoCost hex> <prevOffset> <siblingOffset> <continualTypoIdx> <dialect> <src> <off> <len>     src = S (view into the string) | R (pool)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>
#include <algorithm>
#include <kiwi/Kiwi.h>
#include <kiwi/TypoTransformer.h>
#include "StrUtils.h"
#include "../../../include/kiwi_b200_typo.h"

using namespace kiwi;

static PreparedTypoTransformer makeSet(const std::string& name)
{
	if (name == "kat")
	{
		TypoTransformer tt;
		tt.addTypo(u"ㅐ", u"ㅚ");
		tt.addTypo(u"레", u"뢰");
		tt.addTypo(u"뢨", u"룄");
		return tt.prepare(true);
	}
	if (name == "basic") return getDefaultTypoSet(DefaultTypoSet::basicTypoSet).prepare(true);
	throw std::runtime_error{ "unknown typo set " + name };
}

template<class T> static void put(std::vector<char>& blob, const std::vector<T>& v)
{
	while (blob.size() % 16) blob.push_back(0);
	const char* p = reinterpret_cast<const char*>(v.data());
	blob.insert(blob.end(), p, p + v.size() * sizeof(T));
}

int main(int argc, char** argv)
{
	if (argc < 4) { std::cerr << "usage: typo_tool flatten <set> <out.img> | graphs <set> <input.txt> <out.txt> [max]\n"; return 2; }
	setenv("KIWI_ARCH_TYPE", "balanced", 1);
	try
	{
		const std::string cmd = argv[1];
		PreparedTypoTransformer ptt = makeSet(argv[2]);
		const char16_t* pool = ptt.strPool.data();
		if (cmd == "flatten")
		{
			const auto& ft = ptt.patTrie;
			std::vector<kb2_typo_node> nodes(ft.numNodes);
			std::vector<uint16_t> keys(ft.nextKeys.get(), ft.nextKeys.get() + ft.numNexts);
			std::vector<int32_t> diffs(ft.nextDiffs.get(), ft.nextDiffs.get() + ft.numNexts);
			std::vector<kb2_typo_pat> pats;
			for (size_t i = 0; i < ft.numNodes; ++i)
			{
				const auto& n = ft.nodes[i];
				const auto& v = ft.values[i];
				int32_t value = -1;
				if (ft.hasSubmatch(v)) value = -2;
				else if (!ft.isNull(v))
				{
					value = (int32_t)pats.size();
					pats.push_back(kb2_typo_pat{ (uint32_t)(v.repl - ptt.replacements.data()), v.size, v.patLength });
				}
				nodes[i] = kb2_typo_node{ n.nextOffset, n.lower, value, (uint16_t)n.numNexts, n.depth };
				// the pattern trie is frozen for ArchType::none (keys in build order): sort every node's (key, diff) pairs ascending
				std::vector<std::pair<uint16_t, int32_t>> kv(n.numNexts);
				for (size_t j = 0; j < n.numNexts; ++j) kv[j] = std::make_pair(keys[n.nextOffset + j], diffs[n.nextOffset + j]);
				std::sort(kv.begin(), kv.end());
				for (size_t j = 0; j < n.numNexts; ++j) { keys[n.nextOffset + j] = kv[j].first; diffs[n.nextOffset + j] = kv[j].second; }
			}
			std::vector<kb2_typo_repl> repls;
			for (auto& r : ptt.replacements) repls.push_back(kb2_typo_repl{ (uint32_t)(r.str - pool), r.length, r.cost, (uint8_t)r.leftCond, 0, (uint16_t)r.dialect });
			std::vector<uint16_t> poolv(ptt.strPool.begin(), ptt.strPool.end());
			kb2_typo_header h;
			std::memset(&h, 0, sizeof(h));
			h.magic = KB2_TYPO_MAGIC; h.n_nodes = (uint32_t)nodes.size(); h.n_edges = (uint32_t)keys.size(); h.n_pats = (uint32_t)pats.size();
			h.n_repls = (uint32_t)repls.size(); h.n_pool = (uint32_t)poolv.size();
			h.continual_typo_threshold = ptt.continualTypoThreshold; h.lengthening_typo_threshold = ptt.lengtheningTypoThreshold;
			std::vector<char> blob(reinterpret_cast<const char*>(&h), reinterpret_cast<const char*>(&h) + sizeof(h));
			put(blob, nodes); put(blob, keys); put(blob, diffs); put(blob, pats); put(blob, repls); put(blob, poolv);
			std::ofstream ofs{ argv[3], std::ios_base::binary };
			ofs.write(blob.data(), blob.size());
			std::cerr << "typo image " << argv[2] << ": nodes " << nodes.size() << " edges " << keys.size() << " patterns " << pats.size() << " replacements " << repls.size()
				<< " pool " << poolv.size() << " continual " << h.continual_typo_threshold << " lengthening " << h.lengthening_typo_threshold << std::endl;
			return 0;
		}
		if (cmd == "graphs" && argc >= 5)
		{
			const size_t maxLines = argc > 5 ? std::stoul(argv[5]) : (size_t)-1;
			std::ifstream ifs{ argv[3] };
			FILE* fo = std::fopen(argv[4], "w");
			std::string line; size_t idx = 0;
			while (std::getline(ifs, line) && idx < maxLines)
			{
				if (!line.empty() && line.back() == '\r') line.pop_back();
				const auto tab = line.find('\t');
				if (tab != line.npos) line = line.substr(0, tab);
				std::u16string raw;
				try { raw = utf8To16(line); } catch (...) { ++idx; std::fprintf(fo, "G %zu 0 0\n", idx - 1); continue; }
				std::u16string nstr;
				normalizeHangul(nstr, std::u16string_view{ raw });
				std::vector<TypoGraphNode> graph;
				const size_t n = ptt.generateGraph(nstr, graph);
				std::fprintf(fo, "G %zu %zu %zu\n", idx, nstr.size(), n);
				for (auto& g : graph)
				{
					const bool inStr = g.form.data() >= nstr.data() && g.form.data() <= nstr.data() + nstr.size();
					const long off = inStr ? (long)(g.form.data() - nstr.data()) : (long)(g.form.data() - pool);
					std::fprintf(fo, "N %u %a %u %u %u %u %c %ld %zu\n", g.endPos, g.typoCost, g.prevOffset, g.siblingOffset, (unsigned)g.continualTypoIdx, (unsigned)g.dialect,
						inStr ? 'S' : 'R', off, g.form.size());
				}
				++idx;
			}
			std::fclose(fo);
			std::cerr << "dumped " << idx << " typo graphs\n";
			return 0;
		}
		std::cerr << "bad command\n";
		return 2;
	}
	catch (const std::exception& e) { std::cerr << "typo_tool failed: " << e.what() << std::endl; return 1; }
}
