// Fabricate a format-exact CoNg model directory (cong.mdl next to the fabricated sj.morph) with the reference's
// OWN builder, because models/cong/base/cong.mdl is a git-LFS pointer in the snapshot.
// TEST INFRASTRUCTURE: links oracle/_ref/libkiwi_ref.so (the unmodified reference).
//   lm::CoNgramModelBase::build(contextDefinition, embedding, ...)   /root/reference/src/CoNgramModel.cpp:1660-2030
//   how the reference's own tool calls it                            tools/cong_builder.cpp:13-24
//   embedding.bin layout (header + int8 tables + fp16 scales/biases)  src/CoNgramModel.cpp:1865-1919
//   context definition: "<clusterId>\t<id>\t<id>..." per line         src/CoNgramModel.cpp:1695-1730
// The vocabulary (output rows) is the LM vocabulary of the Knlm model directory given as input, so the copied
// sj.morph (whose lmMorphemeId fields index that vocabulary) stays consistent.  Contexts are the 1..3-grams of
// LM ids the reference itself produces when it analyses the given corpora with the Knlm model; a hash keeps
// ~70 % of the 2/3-grams so that back-off, leaf and miss transitions of the context trie are all exercised.
// Embeddings are seeded pseudo-random int8 in [-64, 63]: with that range the saturating AVX2 `maddubs`
// emulation (src/archImpl/avx2_qgemm.hpp:24-30) equals the exact VNNI dot product, so every x86 arch of the
// reference produces the same integers.
// usage: fabricate_cong <knlm_model_dir> <out_dir> <dim> <contextSize> <corpus.txt>...
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <kiwi/Kiwi.h>
#include <kiwi/CoNgramModel.h>
#include "StrUtils.h"

using namespace kiwi;

static uint16_t f2h(float f)        // round-to-nearest-even float -> IEEE half (normal range only, enough for our constants)
{
	uint32_t x; std::memcpy(&x, &f, 4);
	const uint32_t sign = (x >> 16) & 0x8000;
	int32_t e = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
	uint32_t m = x & 0x7FFFFF;
	if (e <= 0) return (uint16_t)sign;
	if (e >= 31) return (uint16_t)(sign | 0x7C00);
	uint32_t h = (uint32_t)(e << 10) | (m >> 13);
	const uint32_t rem = m & 0x1FFF;
	if (rem > 0x1000 || (rem == 0x1000 && (h & 1))) ++h;
	return (uint16_t)(sign | h);
}

static uint64_t mix(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

int main(int argc, char** argv)
{
	if (argc < 6) { std::cerr << "usage: fabricate_cong <knlm_model_dir> <out_dir> <dim> <contextSize> <corpus>...\n"; return 2; }
	if (!getenv("KIWI_ARCH_TYPE")) setenv("KIWI_ARCH_TYPE", "avx2", 1);
	const std::string inDir = argv[1], outDir = argv[2];
	const uint32_t dim = (uint32_t)std::stoul(argv[3]), contextSize = (uint32_t)std::stoul(argv[4]);
	try
	{
		KiwiBuilder kb{ inDir, 1, BuildOption::default_, ModelType::knlm };
		Kiwi kw = kb.build();
		const uint32_t vocab = (uint32_t)kw.langMdl->vocabSize();

		// ---- contexts from the reference's own analyses
		std::map<std::vector<uint32_t>, uint32_t> ctx;
		size_t nLines = 0;
		for (int a = 5; a < argc; ++a)
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
				for (auto& t : res[0].first)
				{
					if (!t.morph) continue;
					const uint32_t id = t.morph->lmMorphemeId;
					if (id < vocab) ids.push_back(id);
				}
				for (size_t i = 0; i < ids.size(); ++i)
				{
					for (size_t n = 1; n <= 3 && n <= i + 1; ++n)
					{
						std::vector<uint32_t> g(ids.begin() + (i + 1 - n), ids.begin() + i + 1);
						uint64_t h = n;
						for (auto v : g) h = mix(h ^ v);
						if (n > 1 && (h >> 8) % 10 >= 7) continue;
						ctx.emplace(std::move(g), (uint32_t)(mix(h) % (contextSize - 1)));      // clusterId in [0, contextSize-2]; stored +1
					}
				}
				++nLines;
			}
		}
		// a 2/3-gram is only reachable if its prefix exists as a node: the builder's trie.build inserts the whole key
		// path, so no extra care is needed.
		const std::string ctxPath = outDir + "/cong_context.tsv", embPath = outDir + "/cong_embedding.bin";
		{
			std::ofstream ofs{ ctxPath };
			for (auto& p : ctx)
			{
				ofs << p.second;
				for (auto v : p.first) ofs << '\t' << v;
				ofs << '\n';
			}
		}

		// ---- embedding.bin.  The header's windowSize must be > 0 even for the window-less model type: build() always
		// writes the per-context confidence / validTokenSum halves, and the loader only skips them when
		// header.windowSize > 0 (src/CoNgramModel.cpp:2004-2005 vs 646-660) - the shipped cong.mdl serves both
		// ModelType::cong and ModelType::congGlobal (window 7, CoNgramModel.cpp:2893-2898).
		const uint32_t windowSize = 7;
		{
			std::ofstream ofs{ embPath, std::ios_base::binary };
			auto w32 = [&](uint32_t v) { ofs.write((const char*)&v, 4); };
			w32(dim); w32(contextSize); w32(vocab); w32(windowSize); w32(8); w32(0);
			auto randRows = [&](size_t rows, uint64_t seed)
			{
				std::vector<int8_t> v(rows * dim);
				for (size_t i = 0; i < v.size(); ++i) v[i] = (int8_t)((int)(mix(seed * 0x100000001B3ull + i) % 128) - 64);
				ofs.write((const char*)v.data(), v.size());
			};
			auto halfs = [&](size_t n, uint64_t seed, float lo, float hi)
			{
				std::vector<uint16_t> v(n);
				for (size_t i = 0; i < n; ++i) v[i] = f2h(lo + (hi - lo) * (float)(mix(seed * 0x9E37ull + i) % 4096) / 4096.f);
				ofs.write((const char*)v.data(), n * 2);
			};
			const float s = std::sqrt(2.f / (37.f * 37.f * std::sqrt((float)dim)));   // ll std ~ 2
			randRows(contextSize, 1);                      // contextEmb
			halfs(contextSize, 2, s * 0.5f, s * 1.5f);      // contextEmbScale
			halfs(contextSize, 3, 4.f, 9.f);                // contextEmbBias (loaded negated)
			halfs(contextSize, 4, 0.f, 1.f);                // contextValidTokenSum
			halfs(contextSize, 5, 0.f, 1.f);                // contextConfidence
			randRows(vocab, 6);                            // distantEmb (unused without a window)
			halfs(vocab, 7, s * 0.5f, s * 1.5f);            // distantEmbScale
			halfs(vocab, 8, 4.f, 9.f);                      // distantEmbBias
			halfs(vocab, 9, 0.f, 1.f);                      // distantConfidence
			halfs(windowSize, 12, 0.f, 1.f);                // positionConfidence
			randRows(vocab, 10);                           // outputEmb
			halfs(vocab, 11, s * 0.5f, s * 1.5f);           // outputEmbScale
			std::vector<uint8_t> mask(vocab, 0);
			ofs.write((const char*)mask.data(), mask.size());
		}

		auto mem = lm::CoNgramModelBase::build(ctxPath, embPath, (size_t)-1, true, true, nullptr);
		mem.writeToFile(outDir + "/cong.mdl");
		std::cerr << "cong model: vocab " << vocab << " contexts " << ctx.size() << " from " << nLines << " lines, dim " << dim
			<< ", contextSize " << contextSize << ", " << mem.size() << " bytes" << std::endl;
	}
	catch (const std::exception& e)
	{
		std::cerr << "fabricate_cong failed: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
