// kiwi_b200: libstdc++'s std::sort, restated step for step (bits/stl_algo.h: __introsort_loop with __unguarded_partition_pivot /
// __move_median_to_first, depth limit 2 * floor(log2 n) with the heapsort fallback of bits/stl_heap.h, then __final_insertion_sort
// with the threshold 16).
//
// Why: BestPathFinder::findBestPath sorts its end-node candidates with std::sort by (rootId, spState, score desc)
// (/root/reference/src/PathEvaluator.hpp:1359-1368) and keeps the first entries of every group.  std::sort is not stable, so WHICH
// of several equal-score candidates comes first is decided by the algorithm's exact sequence of swaps - two candidate morphemes with
// the same LM id and tag tie regularly (16 of the 8192 bench sentences).  The sort only looks at keys, so sorting (key, index)
// records with the same algorithm reproduces the reference's permutation.  Sequential (one lane); n is the number of end-node
// candidates of a chunk (tens to hundreds).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KB_SORT_HD __host__ __device__ inline
#else
#define KB_SORT_HD inline
#endif

namespace kb
{
	struct alignas(16) SortRec { unsigned long long key; uint32_t idx; uint32_t pad; };      // comp(a, b) = a.key < b.key

	namespace sortemu
	{
		KB_SORT_HD bool lt(const SortRec& a, const SortRec& b) { return a.key < b.key; }
		KB_SORT_HD void swp(SortRec* a, long i, long j) { const SortRec t = a[i]; a[i] = a[j]; a[j] = t; }

		// bits/stl_heap.h
		KB_SORT_HD void pushHeap(SortRec* first, long holeIndex, long topIndex, SortRec value)
		{
			long parent = (holeIndex - 1) / 2;
			while (holeIndex > topIndex && lt(first[parent], value))
			{
				first[holeIndex] = first[parent];
				holeIndex = parent;
				parent = (holeIndex - 1) / 2;
			}
			first[holeIndex] = value;
		}
		KB_SORT_HD void adjustHeap(SortRec* first, long holeIndex, long len, SortRec value)
		{
			const long topIndex = holeIndex;
			long secondChild = holeIndex;
			while (secondChild < (len - 1) / 2)
			{
				secondChild = 2 * (secondChild + 1);
				if (lt(first[secondChild], first[secondChild - 1])) --secondChild;
				first[holeIndex] = first[secondChild];
				holeIndex = secondChild;
			}
			if ((len & 1) == 0 && secondChild == (len - 2) / 2)
			{
				secondChild = 2 * (secondChild + 1);
				first[holeIndex] = first[secondChild - 1];
				holeIndex = secondChild - 1;
			}
			pushHeap(first, holeIndex, topIndex, value);
		}
		KB_SORT_HD void heapSort(SortRec* first, long len)      // __partial_sort(first, last, last) = make_heap + sort_heap
		{
			if (len >= 2)
			{
				long parent = (len - 2) / 2;
				while (true)
				{
					const SortRec value = first[parent];
					adjustHeap(first, parent, len, value);
					if (parent == 0) break;
					--parent;
				}
			}
			long last = len;
			while (last > 1)
			{
				--last;
				const SortRec value = first[last];
				first[last] = first[0];
				adjustHeap(first, 0, last, value);
			}
		}
		KB_SORT_HD void moveMedianToFirst(SortRec* a, long result, long x, long y, long z)
		{
			if (lt(a[x], a[y]))
			{
				if (lt(a[y], a[z])) swp(a, result, y);
				else if (lt(a[x], a[z])) swp(a, result, z);
				else swp(a, result, x);
			}
			else if (lt(a[x], a[z])) swp(a, result, x);
			else if (lt(a[y], a[z])) swp(a, result, z);
			else swp(a, result, y);
		}
		KB_SORT_HD long unguardedPartition(SortRec* a, long first, long last, long pivot)
		{
			while (true)
			{
				while (lt(a[first], a[pivot])) ++first;
				--last;
				while (lt(a[pivot], a[last])) --last;
				if (!(first < last)) return first;
				swp(a, first, last);
				++first;
			}
		}
		KB_SORT_HD void unguardedLinearInsert(SortRec* a, long last)
		{
			const SortRec val = a[last];
			long next = last - 1;
			while (lt(val, a[next])) { a[last] = a[next]; last = next; --next; }
			a[last] = val;
		}
		KB_SORT_HD void insertionSort(SortRec* a, long first, long last)
		{
			if (first == last) return;
			for (long i = first + 1; i != last; ++i)
			{
				if (lt(a[i], a[first]))
				{
					const SortRec val = a[i];
					for (long k = i; k > first; --k) a[k] = a[k - 1];
					a[first] = val;
				}
				else unguardedLinearInsert(a, i);
			}
		}
	}

	// std::sort(a, a + n) of libstdc++
	KB_SORT_HD void stdSortEmu(SortRec* a, long n)
	{
		using namespace sortemu;
		if (n <= 0) return;
		long lg = 0; for (long t = n; t > 1; t >>= 1) ++lg;
		// __introsort_loop with the recursion on the right part turned into an explicit stack
		struct Frame { int32_t first, last, depth; };
		Frame stack[48]; int sp = 0;      // right parts waiting: at most one per recursion level (<= 2 * log2 n + 1)
		stack[sp++] = Frame{ 0, (int32_t)n, (int32_t)(2 * lg) };
		while (sp)
		{
			Frame f = stack[--sp];
			long first = f.first, last = f.last, depth = f.depth;
			while (last - first > 16)
			{
				if (depth == 0) { heapSort(a + first, last - first); break; }
				--depth;
				const long mid = first + (last - first) / 2;
				moveMedianToFirst(a, first, first + 1, mid, last - 1);
				const long cut = unguardedPartition(a, first + 1, last, first);
				// the reference recurses into [cut, last) FIRST and then continues with [first, cut): the two parts are disjoint, so the order
				// in which they are processed does not change the result - push the right part, keep looping on the left
				if (sp < 48) stack[sp++] = Frame{ (int32_t)cut, (int32_t)last, (int32_t)depth };
				last = cut;
			}
		}
		// __final_insertion_sort
		if (n > 16)
		{
			insertionSort(a, 0, 16);
			for (long i = 16; i != n; ++i) unguardedLinearInsert(a, i);
		}
		else insertionSort(a, 0, n);
	}
}
