// CPU check (tests/test_emulations.py): the two libstdc++ behaviours the kernels restate on the device - std::sort's permutation of
// equal keys (kiwi_b200/csrc/std_sort_emu.h) and std::unordered_set's iteration order (kiwi_b200/csrc/unordered_emu.h) - against the
// real library containers, on random inputs with many ties / collisions, re-used sets and adversarial (median-of-3 killer) orders.
#include "../../kiwi_b200/csrc/std_sort_emu.h"
#include "../../kiwi_b200/csrc/unordered_emu.h"
#include <algorithm>
#include <cstdio>
#include <random>
#include <unordered_set>
#include <vector>

struct Key { unsigned long long code; int id; bool operator==(const Key& o) const { return id == o.id; } };
struct H { size_t operator()(const Key& k) const { return (size_t)k.code; } };

int main(int argc, char** argv)
{
	const int scale = argc > 1 ? std::atoi(argv[1]) : 1;
	long bad = 0, sorts = 0, sets = 0;
	{
		std::mt19937 rng(123);
		for (int trial = 0; trial < 20000 * scale; ++trial)
		{
			const long n = trial % 7 == 0 ? rng() % 3000 : rng() % 200;
			const int nkeys = 1 + rng() % (trial % 3 == 0 ? 4 : 50);
			std::vector<kb::SortRec> a(n), b;
			for (long i = 0; i < n; ++i) { a[i].key = rng() % nkeys; a[i].idx = (uint32_t)i; a[i].pad = 0; }
			if (trial % 11 == 0) std::sort(a.begin(), a.end(), [](auto& x, auto& y) { return x.key < y.key; });
			if (trial % 13 == 0) std::sort(a.begin(), a.end(), [](auto& x, auto& y) { return x.key > y.key; });
			b = a;
			std::sort(b.begin(), b.end(), [](const kb::SortRec& x, const kb::SortRec& y) { return x.key < y.key; });
			kb::stdSortEmu(a.data(), n);
			++sorts;
			for (long i = 0; i < n; ++i) if (a[i].idx != b[i].idx) { ++bad; break; }
		}
		for (int n : { 1000, 5000, 20000 })
		{
			std::vector<kb::SortRec> a(n), b; const int k = n / 2; std::vector<int> v(n);
			for (int i = 1; i <= k; ++i) { if (i % 2 == 1) { v[i - 1] = i; v[i] = k + i; } v[k + i - 1] = 2 * i; }
			for (int i = 0; i < n; ++i) { a[i].key = v[i]; a[i].idx = i; a[i].pad = 0; }
			b = a; std::sort(b.begin(), b.end(), [](const kb::SortRec& x, const kb::SortRec& y) { return x.key < y.key; });
			kb::stdSortEmu(a.data(), n); ++sorts;
			for (int i = 0; i < n; ++i) if (a[i].idx != b[i].idx) { ++bad; break; }
		}
	}
	{
		std::mt19937_64 rng(7);
		for (int trial = 0; trial < 300 * scale; ++trial)
		{
			std::unordered_set<Key, H> s;      // fresh per "sentence", re-used (cleared, never shrunk) by its containers
			uint32_t B = 1;
			const int rounds = 1 + rng() % 12;
			for (int r = 0; r < rounds; ++r)
			{
				s.clear();
				const int n = (rng() % 5 == 0) ? rng() % 3000 : rng() % 200;
				std::vector<unsigned long long> codes(n);
				const int mode = 1 + rng() % 3;
				for (int i = 0; i < n; ++i) codes[i] = mode == 1 ? (rng() % 64) * 8 + (rng() % 3) : rng();
				for (int i = 0; i < n; ++i) { s.insert(Key{ codes[i], i }); if (i > 0 && rng() % 3 == 0) { const int d = rng() % i; s.insert(Key{ codes[d], d }); } }
				std::vector<int32_t> next(n + 1), order(n + 1);
				const uint32_t after = kb::unorderedBucketsAfter(B, n);
				std::vector<int32_t> buckets(after ? after : 1);
				const bool ok = kb::unorderedSetOrder(codes.data(), n, B, next.data(), buckets.data(), order.data());
				++sets;
				if (!ok || B != s.bucket_count() || B != after) { ++bad; continue; }
				int j = 0; bool diff = false;
				for (auto& k : s) { if (order[j] != k.id) { diff = true; break; } ++j; }
				if (diff || j != n) ++bad;
			}
		}
	}
	std::printf("sorts %ld sets %ld mismatching %ld\n", sorts, sets, bad);
	return bad != 0;
}
