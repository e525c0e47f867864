// CPU baseline timer: the UNMODIFIED reference (oracle/_ref/libkiwi_ref.so) on a sentence file.
// TEST/BENCH INFRASTRUCTURE.  Protocol = MorphEvaluator::eval (/root/reference/tools/Evaluator.cpp:315-329):
// wall time of `for line: kiwi.analyze(line, 1, option)`; with threads > 1 the reference's own batch mode
// Kiwi::analyze(topN, reader, receiver, option) (include/kiwi/Kiwi.h:402-454) on a pool of `threads`.
// Prints one JSON line: {"sentences": n, "threads": t, "repeats": r, "seconds": best, "sent_per_s": v, "arch": "..."}
// usage: ref_bench <model_dir> <input.txt> <threads> <repeats> [maxLines]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <kiwi/Kiwi.h>
#include "StrUtils.h"

using namespace kiwi;

int main(int argc, char** argv)
{
	if (argc < 5) { std::cerr << "usage: ref_bench <model_dir> <input.txt> <threads> <repeats> [maxLines]\n"; return 2; }
	if (!getenv("KIWI_ARCH_TYPE")) setenv("KIWI_ARCH_TYPE", "avx2", 1);
	const size_t threads = std::stoul(argv[3]), repeats = std::stoul(argv[4]);
	const size_t maxLines = argc > 5 ? std::stoul(argv[5]) : (size_t)-1;
	try
	{
		const char* mt = getenv("KB_MODEL_TYPE");      // knlm (default), cong or sbg
		KiwiBuilder kb{ argv[1], threads <= 1 ? 0 : threads, BuildOption::default_, (mt && std::string{ mt } == "cong") ? ModelType::cong : (mt && std::string{ mt } == "sbg") ? ModelType::sbg : ModelType::knlm };
		Kiwi kw = kb.build();
		std::vector<std::u16string> lines;
		{
			std::ifstream ifs{ argv[2] };
			std::string line;
			while (std::getline(ifs, line) && lines.size() < maxLines)
			{
				const auto tab = line.find('\t');
				if (tab != line.npos) line = line.substr(0, tab);
				auto u = utf8To16(line);
				if (u.empty()) u = u" ";      // the batch reader treats an empty string as end of input
				lines.emplace_back(std::move(u));
			}
		}
		AnalyzeOption option;
		double best = 1e30;
		size_t tokens = 0;
		for (size_t r = 0; r < repeats; ++r)
		{
			tokens = 0;
			const auto t0 = std::chrono::steady_clock::now();
			if (threads <= 1)
			{
				for (auto& l : lines) { auto res = kw.analyze(l, 1, option); tokens += res[0].first.size(); }
			}
			else
			{
				size_t idx = 0;
				kw.analyze(1, [&]() { return idx < lines.size() ? lines[idx++] : std::u16string{}; },
					[&](std::vector<TokenResult>&& res) { tokens += res[0].first.size(); }, option);
			}
			const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			if (s < best) best = s;
		}
		std::printf("{\"sentences\": %zu, \"threads\": %zu, \"repeats\": %zu, \"seconds\": %.6f, \"sent_per_s\": %.2f, \"tokens\": %zu, \"arch\": \"%s\"}\n",
			lines.size(), threads, repeats, best, lines.size() / best, tokens, archToStr(kw.archType()));
	}
	catch (const std::exception& e)
	{
		std::cerr << "ref_bench failed: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
