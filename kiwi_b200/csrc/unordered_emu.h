// kiwi_b200: the iteration order of libstdc++'s std::unordered_set, restated (bits/hashtable.h: _M_insert_unique_node /
// _M_insert_bucket_begin, _M_rehash_aux(unique keys), _Prime_rehash_policy with max_load_factor 1 and growth factor 2).
//
// Why: with more than 512 incoming paths the reference collects a candidate's paths in BestPathConatiner<top1> - an
// std::unordered_set - and writes them out in the set's ITERATION order (/root/reference/src/BestPathContainer.hpp:229-276).  That
// order fixes the order of the node's paths and through it every later tie-break (first-inserted wins in the containers, std::sort of
// equal end-node scores), so it is part of the result.  The order is a function of (a) the bucket count the set has grown to - the
// set is cleared, never shrunk, between uses - and (b) the hash codes of the distinct keys in first-insertion order; the kernels know
// both: keys are appended in first-insertion order, and the bucket count is a per-sentence state that starts at 1 like the oracle's
// (the reference's thread_local set additionally remembers earlier sentences of its thread: not reproducible by definition).
// Sequential (one lane); n = distinct keys of one candidate's container (hundreds).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KB_UO_HD __host__ __device__ inline
#else
#define KB_UO_HD inline
#endif

namespace kb
{
	// bucket counts libstdc++ walks through from a default-constructed set: 1 -> 13 on the first insertion, then the smallest prime of its
	// __prime_list that is >= twice the current count, whenever an insertion would push the size above the bucket count
	KB_UO_HD uint32_t unorderedNextBuckets(uint32_t b)
	{
		const uint32_t chain[] = { 1u, 13u, 29u, 59u, 127u, 257u, 541u, 1109u, 2357u, 5087u, 10273u, 20753u, 42043u, 85229u, 172933u, 351061u, 712697u, 1447153u, 2938679u };
		for (unsigned i = 0; i + 1 < sizeof(chain) / sizeof(chain[0]); ++i) if (chain[i] == b) return chain[i + 1];
		return 0;      // not on the chain: the caller treats it as an internal error
	}
	// the bucket count after a container that held `n` distinct keys, starting from `b`
	KB_UO_HD uint32_t unorderedBucketsAfter(uint32_t b, uint32_t n)
	{
		while (n > (b == 1 ? 0u : b)) { b = unorderedNextBuckets(b); if (!b) return 0; }
		return b;
	}

	// codes[i]: hash code of the i-th distinct key in first-insertion order.  bucketCount: in = the set's bucket count before the first
	// insertion, out = after the last.  next[n], buckets[bucketCount after]: scratch.  order[j] = insertion index of the j-th element in
	// iteration order.  Returns false when the bucket chain is exhausted.
	KB_UO_HD bool unorderedSetOrder(const unsigned long long* codes, int32_t n, uint32_t& bucketCount, int32_t* next, int32_t* buckets, int32_t* order)
	{
		const int32_t EMPTY = -2, BEFORE_BEGIN = -1;
		uint32_t B = bucketCount;
		int32_t beginNext = -1;      // _M_before_begin._M_nxt
		for (uint32_t b = 0; b < B; ++b) buckets[b] = EMPTY;
		for (int32_t i = 0; i < n; ++i)
		{
			if ((uint32_t)i + 1 > (B == 1 ? 0u : B))
			{
				// _M_rehash_aux(n, unique keys): relink every node into the new bucket array, walking the old list from its head
				const uint32_t nb = unorderedNextBuckets(B);
				if (!nb) return false;
				B = nb;
				for (uint32_t b = 0; b < B; ++b) buckets[b] = EMPTY;
				int32_t p = beginNext; beginNext = -1;
				uint32_t bbeginBkt = 0;
				while (p >= 0)
				{
					const int32_t nx = next[p];
					const uint32_t bkt = (uint32_t)(codes[p] % B);
					if (buckets[bkt] == EMPTY)
					{
						next[p] = beginNext; beginNext = p;
						buckets[bkt] = BEFORE_BEGIN;
						if (next[p] >= 0) buckets[bbeginBkt] = p;
						bbeginBkt = bkt;
					}
					else
					{
						const int32_t before = buckets[bkt];
						if (before == BEFORE_BEGIN) { next[p] = beginNext; beginNext = p; }
						else { next[p] = next[before]; next[before] = p; }
					}
					p = nx;
				}
			}
			// _M_insert_bucket_begin
			const uint32_t bkt = (uint32_t)(codes[i] % B);
			if (buckets[bkt] != EMPTY)
			{
				const int32_t before = buckets[bkt];
				if (before == BEFORE_BEGIN) { next[i] = beginNext; beginNext = i; }
				else { next[i] = next[before]; next[before] = i; }
			}
			else
			{
				next[i] = beginNext; beginNext = i;
				if (next[i] >= 0) buckets[(uint32_t)(codes[next[i]] % B)] = i;
				buckets[bkt] = BEFORE_BEGIN;
			}
		}
		int32_t j = 0;
		for (int32_t p = beginNext; p >= 0; p = next[p]) order[j++] = p;
		bucketCount = B;
		return j == n;
	}
}
