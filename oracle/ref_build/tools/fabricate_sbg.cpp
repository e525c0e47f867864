// Fabricate a format-exact skipbigram.mdl next to the fabricated sj.knlm (the reference's trainer needs full Eigen and
// a large corpus; the file format itself is simple).  TEST INFRASTRUCTURE: links oracle/_ref/libkiwi_ref.so.
//   SkipBigramModelHeader                      /root/reference/include/kiwi/SkipBigramModel.h:9-13
//   layout read by the loader (non-quantized)  src/SkipBigramModel.hpp:40-105:
//     header | KeyType kSizes[vocab] | KeyType keys[total] (ascending per target) | float discnts[vocab] |
//     float compensations[total] | uint8 vocabValidness[vocab]
// Targets and their history keys are the co-occurrences (window 8) in the reference's own analyses of the given
// corpora with the Knlm model; ~65 % of the pairs are kept so that both hits and misses occur.  Values are seeded:
// discnts in [-0.7, 0], compensations in [-7, -0.5] (natural-log probabilities like the Knlm scores).
// usage: fabricate_sbg <knlm_model_dir> <out_dir> <corpus.txt>...
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <kiwi/Kiwi.h>
#include <kiwi/SkipBigramModel.h>
#include "StrUtils.h"

using namespace kiwi;

static uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }

int main(int argc, char** argv)
{
	if (argc < 4) { std::cerr << "usage: fabricate_sbg <knlm_model_dir> <out_dir> <corpus>...\n"; return 2; }
	if (!getenv("KIWI_ARCH_TYPE")) setenv("KIWI_ARCH_TYPE", "avx2", 1);
	try
	{
		KiwiBuilder kb{ argv[1], 1, BuildOption::default_, ModelType::knlm };
		Kiwi kw = kb.build();
		const uint32_t vocab = (uint32_t)kw.langMdl->vocabSize();
		std::map<uint32_t, std::set<uint32_t>> pairs;      // target -> history keys
		std::set<uint32_t> seen;
		for (int a = 3; a < argc; ++a)
		{
			std::ifstream ifs{ argv[a] };
			std::string line;
			while (std::getline(ifs, line))
			{
				const auto tab = line.find('\t');
				if (tab != line.npos) line = line.substr(0, tab);
				if (line.empty()) continue;
				auto res = kw.analyze(utf8To16(line), 1, AnalyzeOption{});
				std::vector<uint32_t> ids;
				for (auto& t : res[0].first) if (t.morph && t.morph->lmMorphemeId < vocab) ids.push_back(t.morph->lmMorphemeId);
				for (size_t i = 0; i < ids.size(); ++i)
				{
					seen.insert(ids[i]);
					for (size_t j = i > 8 ? i - 8 : 0; j < i; ++j)
					{
						if (mix(((uint64_t)ids[i] << 32) | ids[j]) % 100 < 65) pairs[ids[i]].insert(ids[j]);
					}
				}
			}
		}
		lm::SkipBigramModelHeader h;
		std::memset(&h, 0, sizeof(h));
		h.vocabSize = vocab; h.keySize = 4; h.windowSize = 8; h.compressed = 0; h.quantize = 0;
		std::vector<uint32_t> kSizes(vocab, 0), keys;
		std::vector<float> discnts(vocab), comps;
		std::vector<uint8_t> valid(vocab, 0);
		for (uint32_t w = 0; w < vocab; ++w)
		{
			discnts[w] = -0.7f * (float)(mix(w * 31ull + 7) % 1024) / 1024.f;
			if (seen.count(w)) valid[w] = 1;
			auto it = pairs.find(w);
			if (it == pairs.end()) continue;
			kSizes[w] = (uint32_t)it->second.size();
			for (auto k : it->second)
			{
				keys.push_back(k);
				comps.push_back(-0.5f - 6.5f * (float)(mix(((uint64_t)w << 32) ^ k ^ 0x5bd1e995) % 4096) / 4096.f);
			}
		}
		std::ofstream ofs{ std::string{ argv[2] } + "/skipbigram.mdl", std::ios_base::binary };
		ofs.write((const char*)&h, sizeof(h));
		ofs.write((const char*)kSizes.data(), kSizes.size() * 4);
		ofs.write((const char*)keys.data(), keys.size() * 4);
		ofs.write((const char*)discnts.data(), discnts.size() * 4);
		ofs.write((const char*)comps.data(), comps.size() * 4);
		ofs.write((const char*)valid.data(), valid.size());
		std::cerr << "sbg model: vocab " << vocab << ", valid targets " << seen.size() << ", pairs " << keys.size() << std::endl;
	}
	catch (const std::exception& e) { std::cerr << "fabricate_sbg failed: " << e.what() << std::endl; return 1; }
	return 0;
}
