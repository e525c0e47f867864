// ORACLE driver (test infrastructure): prints the restatement's results in the exact text format of
// oracle/ref_build/tools/dump_golden.cpp, so the two files can be compared line by line.
// usage: oracle_main <model.img> <input.txt> <out.txt> [maxLines]
#include <fstream>
#include <iostream>
#include <string>
#include "analyze.hpp"

static std::vector<uint16_t> utf8To16(const std::string& s)
{
	std::vector<uint16_t> out;
	for (size_t i = 0; i < s.size();)
	{
		uint32_t c = (uint8_t)s[i]; size_t n = 1;
		if (c >= 0xF0) { c &= 7; n = 4; } else if (c >= 0xE0) { c &= 15; n = 3; } else if (c >= 0xC0) { c &= 31; n = 2; }
		for (size_t k = 1; k < n && i + k < s.size(); ++k) c = (c << 6) | ((uint8_t)s[i + k] & 63);
		i += n;
		if (c >= 0x10000) { c -= 0x10000; out.push_back((uint16_t)(0xD800 | (c >> 10))); out.push_back((uint16_t)(0xDC00 | (c & 0x3FF))); }
		else out.push_back((uint16_t)c);
	}
	return out;
}

int main(int argc, char** argv)
{
	if (argc < 4) { std::cerr << "usage: oracle_main <model.img> <input.txt> <out.txt> [maxLines]\n"; return 2; }
	const size_t maxLines = argc > 4 ? std::stoul(argv[4]) : (size_t)-1;
	orc::Image im;
	im.load(argv[1]);
	orc::Analyzer an{ im };
	orc::Counters cnt;
	an.viterbi.cnt = &cnt;
	std::ifstream ifs{ argv[2] };
	FILE* fo = std::fopen(argv[3], "w");
	std::string line;
	size_t idx = 0;
	while (std::getline(ifs, line) && idx < maxLines)
	{
		if (!line.empty() && line.back() == '\r') line.pop_back();
		const auto tab = line.find('\t');
		if (tab != line.npos) line = line.substr(0, tab);
		auto str = utf8To16(line);
		const uint64_t po0 = cnt.pathsOut;
		auto res = an.analyze(str.data(), str.size());
		{ static double maxRatio = 0; double r = (double)(cnt.pathsOut - po0) / (double)(res.normLen + 1); if (r > maxRatio) { maxRatio = r; std::cerr << "ratio " << r << " at line " << idx << " normLen " << res.normLen << "\n"; } }
		std::fprintf(fo, "S %zu %zu %a %zu %zu\n", idx, res.tokens.size(), res.score, res.chunks.size(), res.normLen);
		for (auto& t : res.tokens) std::fprintf(fo, "T %u %u %u %u %a\n", t.morph, (unsigned)t.tag, t.position, (unsigned)t.length, t.score);
		for (size_t c = 0; c < res.chunks.size(); ++c)
		{
			auto& ch = res.chunks[c];
			std::fprintf(fo, "C %zu %zu %zu %zu %zu\n", c, ch.start, ch.end, ch.nodes.size(), ch.paths.size());
			for (auto& n : ch.nodes)
			{
				std::fprintf(fo, "N %ld %ld %zu %u %u %u %u %u %a\n", (long)n.form, n.uformLen ? (long)n.uformOff : -1L, (size_t)n.uformLen,
					n.prev, n.sibling, n.startPos, n.endPos, n.spaceErrors, n.typoCost);
			}
			for (auto& p : ch.paths)
			{
				std::fprintf(fo, "P %a %u %u %zu\n", p.score, (unsigned)p.prevState, (unsigned)p.curState, p.path.size());
				for (auto& k : p.path) std::fprintf(fo, "K %u %u %u %a %u %d\n", k.morph, k.begin, k.end, k.wordScore, k.nodeId, k.strLen ? 1 : 0);
			}
		}
		++idx;
	}
	std::fclose(fo);
	std::cerr << "oracle: " << idx << " lines; lmSteps " << cnt.lmSteps << " pairs " << cnt.pairs << " inserts " << cnt.inserts
		<< " pathsOut " << cnt.pathsOut << " top1Mode " << cnt.top1Mode << " bucketFull " << cnt.bucketFull
		<< " maxNodePre " << cnt.maxNodePre << " maxIncoming " << cnt.maxIncoming << " mediumMode " << cnt.mediumMode << " evalCalls " << cnt.evalCalls << " candEvals " << cnt.candEvals << " maxCont " << cnt.maxCont << std::endl;
	return 0;
}
