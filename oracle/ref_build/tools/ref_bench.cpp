// CPU baseline timer: the UNMODIFIED reference (oracle/_ref/libkiwi_ref.so) on a sentence file.
// TEST/BENCH INFRASTRUCTURE.  Protocol = MorphEvaluator::eval (/root/reference/tools/Evaluator.cpp:315-329):
// wall time of `for line: kiwi.analyze(line, 1, option)`; with threads > 1 the reference's own batch mode
// Kiwi::analyze(topN, reader, receiver, option) (include/kiwi/Kiwi.h:402-454) on a pool of `threads`.
// All passes run inside ONE process (model built once); best and median of `repeats` passes are reported.
//   KB_MODEL_TYPE = knlm (default) | cong | sbg        KB_TYPO = basic: option.typoTransformer = basicTypoSet.prepare(true) (`--typo`, tools/Evaluator.cpp:124-137)
//   KB_SINGLE_SAMPLE = K: additionally time the single-thread protocol on the first K lines (1 pass)
//   KB_DUMP = path: after timing, write every sentence's tokens as binary rows {u32 n_tokens, f32 score, n x {u32 morph, u32 position, f32 score,
//             u16 length, u8 tag, u8 0}} (the device's packed token row) for the full-batch parity test
// Prints one JSON line.   usage: ref_bench <model_dir> <input.txt> <threads> <repeats> [maxLines]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <kiwi/Kiwi.h>
#include <kiwi/TypoTransformer.h>
#include "StrUtils.h"

using namespace kiwi;

int main(int argc, char** argv)
{
	if (argc < 5) { std::cerr << "usage: ref_bench <model_dir> <input.txt> <threads> <repeats> [maxLines]\n"; return 2; }
	if (!getenv("KIWI_ARCH_TYPE")) setenv("KIWI_ARCH_TYPE", "avx2", 1);
	const size_t threads = std::stoul(argv[3]), repeats = std::max<size_t>(1, std::stoul(argv[4]));
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
		PreparedTypoTransformer ptt;
		const char* ty = getenv("KB_TYPO");
		if (ty && *ty)
		{
			if (std::string{ ty } != "basic") throw std::runtime_error{ "KB_TYPO: only `basic` is known" };
			ptt = getDefaultTypoSet(DefaultTypoSet::basicTypoSet).prepare(true);
			option.typoTransformer = &ptt;
		}
		std::vector<double> secs;
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
			secs.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
		}
		std::vector<double> sorted = secs;
		std::sort(sorted.begin(), sorted.end());
		const double best = sorted.front(), median = sorted[sorted.size() / 2];
		double singleRate = 0; size_t singleN = 0;
		if (const char* ss = getenv("KB_SINGLE_SAMPLE"))
		{
			singleN = std::min<size_t>(lines.size(), std::stoul(ss));
			const auto t0 = std::chrono::steady_clock::now();
			size_t tk = 0;
			for (size_t i = 0; i < singleN; ++i) { auto res = kw.analyze(lines[i], 1, option); tk += res[0].first.size(); }
			const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			singleRate = singleN / s;
			(void)tk;
		}
		if (const char* dp = getenv("KB_DUMP"))
		{
			FILE* fo = std::fopen(dp, "wb");
			if (!fo) throw std::runtime_error{ "cannot open KB_DUMP file" };
			const uint32_t n = (uint32_t)lines.size();
			std::fwrite(&n, 4, 1, fo);
			auto writeRes = [&](const TokenResult& r)
			{
				const uint32_t nt = (uint32_t)r.first.size(); const float sc = r.second;
				std::fwrite(&nt, 4, 1, fo); std::fwrite(&sc, 4, 1, fo);
				for (auto& t : r.first)
				{
					struct { uint32_t morph, position; float score; uint16_t length; uint8_t tag, zero; } row;
					row.morph = t.morph ? (uint32_t)kw.morphToId(t.morph) : 0xFFFFFFFFu; row.position = t.position; row.score = t.score;
					row.length = (uint16_t)t.length; row.tag = (uint8_t)t.tag; row.zero = 0;
					std::fwrite(&row, 16, 1, fo);
				}
			};
			if (threads <= 1) { for (auto& l : lines) { auto res = kw.analyze(l, 1, option); writeRes(res[0]); } }
			else
			{
				size_t idx = 0;      // the receiver is called in input order (ordered pool delivery, Kiwi.h:402-454)
				kw.analyze(1, [&]() { return idx < lines.size() ? lines[idx++] : std::u16string{}; },
					[&](std::vector<TokenResult>&& res) { writeRes(res[0]); }, option);
			}
			std::fclose(fo);
		}
		std::printf("{\"sentences\": %zu, \"threads\": %zu, \"repeats\": %zu, \"seconds\": %.6f, \"seconds_median\": %.6f, \"sent_per_s\": %.2f, \"sent_per_s_median\": %.2f, "
			"\"single_thread_sent_per_s\": %.2f, \"single_thread_sample\": %zu, \"tokens\": %zu, \"arch\": \"%s\", \"typo\": \"%s\"}\n",
			lines.size(), threads, repeats, best, median, lines.size() / best, lines.size() / median, singleRate, singleN, tokens, archToStr(kw.archType()), ty ? ty : "");
	}
	catch (const std::exception& e)
	{
		std::cerr << "ref_bench failed: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
