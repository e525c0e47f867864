// Known-answer dump of the reference's CoNg int8 scorer (test infrastructure; links the unmodified reference):
//   Q <m> <n> <ctx ids...> <out ids...> <m*n hex floats>      qgemm::scatteredGEMMOpt<ArchType::avx2> on the model's own tables,
//                                                              exactly as progressMatrixNoWindow calls it (src/CoNgramModel.cpp:1575-1579)
//   P <node> <ctx> <wid> <ll hex> <node'> <ctx'>              lm::CoNgramModel::progressOneStep (scalar `next`, CoNgramModel.cpp:869-903)
// The reference's tests hold no vectors for these functions, so this dump is the pin for the oracle's
// epilogue / context-trie restatement (tests/golden/cong_qgemm.golden.txt.gz).
// usage: cong_probe <cong_model_dir> <out.txt>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>
#include <kiwi/Kiwi.h>
#include "CoNgramModel.hpp"
#include "qgemm.h"

using namespace kiwi;

static uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }

template<class Model> static bool run(const lm::ILangModel* base, FILE* fo)
{
	auto* m = dynamic_cast<const Model*>(base);
	if (!m) return false;
	const auto& hd = m->getHeader();
	uint64_t seed = 1;
	for (size_t mm = 1; mm <= 10; ++mm) for (size_t nn = 1; nn <= 10; ++nn) for (int rep = 0; rep < 3; ++rep)
	{
		std::vector<int32_t> a(mm), b(nn);
		for (auto& v : a) v = (int32_t)(mix(seed++) % hd.contextSize);
		for (auto& v : b) v = (int32_t)(mix(seed++) % hd.vocabSize);
		std::vector<float> c(((mm + 7) / 8 * 8) * ((nn + 7) / 8 * 8) + 64, 0.f);
		qgemm::scatteredGEMMOpt<ArchType::avx2>(mm, nn, hd.dim, m->getContextQuantEmb(0), a.data(), m->contextEmbStride(),
			m->getOutputQuantEmb(0), b.data(), m->outputEmbStride(), c.data(), nn);
		std::fprintf(fo, "Q %zu %zu", mm, nn);
		for (auto v : a) std::fprintf(fo, " %d", v);
		for (auto v : b) std::fprintf(fo, " %d", v);
		for (size_t i = 0; i < mm * nn; ++i) std::fprintf(fo, " %a", c[i]);
		std::fprintf(fo, "\n");
	}
	// scalar steps: random walks through the context trie
	int32_t node = 0; uint32_t ctx = 0;
	for (int i = 0; i < 4000; ++i)
	{
		const uint64_t r = mix(seed++);
		uint32_t wid = (uint32_t)(r % hd.vocabSize);
		if (i % 3) wid = (uint32_t)(mix(r) % 3000);       // frequent ids keep the walk inside the trie
		const int32_t n0 = node; const uint32_t c0 = ctx;
		const float ll = m->progressOneStep(node, ctx, wid);
		std::fprintf(fo, "P %d %u %u %a %d %u\n", n0, c0, wid, ll, node, ctx);
		if (i % 17 == 0) { node = 0; ctx = 0; }
	}
	return true;
}

int main(int argc, char** argv)
{
	if (argc < 3) { std::cerr << "usage: cong_probe <cong_model_dir> <out.txt>\n"; return 2; }
	setenv("KIWI_ARCH_TYPE", "avx2", 1);
	try
	{
		KiwiBuilder kb{ argv[1], 1, BuildOption::default_, ModelType::cong };
		Kiwi kw = kb.build();
		FILE* fo = std::fopen(argv[2], "w");
		const auto* b = kw.langMdl.get();
		if (!run<lm::CoNgramModel<ArchType::avx2, uint16_t, uint16_t, 0, true>>(b, fo)
			&& !run<lm::CoNgramModel<ArchType::avx2, uint32_t, uint16_t, 0, true>>(b, fo)
			&& !run<lm::CoNgramModel<ArchType::avx2, uint32_t, uint32_t, 0, true>>(b, fo)) throw std::runtime_error{ "not an avx2 quantized CoNg model" };
		std::fclose(fo);
	}
	catch (const std::exception& e) { std::cerr << "cong_probe failed: " << e.what() << std::endl; return 1; }
	return 0;
}
