// ORACLE (test infrastructure): CPU restatement of KnLangModel::progress,
// /root/reference/src/Knlm.cpp:44-130, over the flat image arrays (keys sorted ascending per node, i.e. the
// reference's ArchType::balanced layout, src/search.cpp:238-292).  Float additions are in reference order.
#pragma once
#include "image.hpp"

namespace orc
{
	struct Knlm
	{
		const Image& im;
		mutable WorkCounters* wc = nullptr;
		explicit Knlm(const Image& _im) : im{ _im } {}

		bool search(const kb2_kn_node& n, uint32_t key, int32_t& v) const
		{
			if (wc) wc->lmProbes += ceilLog2p1(n.num_nexts);
			const uint32_t* keys = im.knKeys + n.next_offset;
			const uint32_t* it = std::lower_bound(keys, keys + n.num_nexts, key);
			if (it == keys + n.num_nexts || *it != key) return false;
			v = im.knValues[n.next_offset + (it - keys)];
			return true;
		}

		static float asFloat(int32_t v) { float f; std::memcpy(&f, &v, 4); return f; }

		int32_t htxFallback(uint32_t next) const
		{
			if (!im.knHtx) return 0;
			int32_t lv;
			if (search(im.knNodes[0], im.knHtx[next], lv)) return lv;
			return 0;
		}

		float progress(int32_t& nodeIdx, uint32_t next) const
		{
			if (wc) wc->lmSteps++;
			float acc = 0;
			while (1)
			{
				int32_t v;
				const kb2_kn_node* node = &im.knNodes[nodeIdx];
				if (wc) wc->lmHops++;
				if (nodeIdx == 0)
				{
					v = im.knRoot[next];
					if (v == 0)
					{
						if (im.knHtx) nodeIdx = htxFallback(next);
						return acc + im.h->kn_unk_ll;
					}
				}
				else
				{
					if (!search(*node, next, v))
					{
						acc += node->gamma;
						nodeIdx += node->lower;
						continue;
					}
				}
				if (v > 0)
				{
					nodeIdx += v;
					return acc + im.knNodes[nodeIdx].ll;
				}
				else
				{
					while (node->lower)
					{
						node += node->lower;
						if (wc) wc->lmHops++;
						int32_t lv;
						if (search(*node, next, lv))
						{
							if (lv > 0)
							{
								node += lv;
								nodeIdx = (int32_t)(node - im.knNodes);
								return acc + asFloat(v);
							}
						}
					}
					nodeIdx = im.knHtx ? htxFallback(next) : 0;
					return acc + asFloat(v);
				}
			}
		}
	};
}
