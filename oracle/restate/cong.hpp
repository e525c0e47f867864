// ORACLE (test infrastructure): CPU restatement of the quantized CoNg language model without window
// (lm::CoNgramModel<arch, KeyType, VlKeyType, 0, true>) over the flat image arrays:
//   progressContextNode / progressContextNodeVl   /root/reference/src/CoNgramModel.hpp:271-385
//   progress (scalar `next`, no-window branch)     src/CoNgramModel.cpp:869-903
//   progressMatrixNoWindow (sort+unique, gather GEMM, shape-dependent epilogue)   src/CoNgramModel.cpp:1494-1611
//   qgemm::scatteredGEMMOpt dispatch               src/qgemm.hpp:157-201
//   AVX2 / AVX-VNNI kernels and their epilogues    src/archImpl/avx2_qgemm.hpp:63-477, src/archImpl/avx2.cpp:20-70
// The integer dot product is exact (u8 x s8 -> s32).  The float epilogue of the reference depends on WHICH
// kernel the (uniqueContexts m, uniqueOutputs n) shape dispatches to; this file reproduces each association
// (pinned by tests/golden/cong_qgemm.golden.txt.gz, dumped from the unmodified reference):
//   E_scalar  ((float)(acc - hsum) * cs) * os + cb                 two roundings, no FMA   (CoNgramModel.cpp TU: baseline x86-64)
//   E_small   fma((float)(acc - hsum) * cs, os, cb)                m <= 3 && n <= 3, and the generic baseline (avx2 TU, -mfma, contracted)
//   E_gemv    fma((float)(acc - hsum) * os, cs, cb)                _mm_fmadd_ps(_mm_mul_ps(cvt, bScale), aScale, aBias): n == 1, or m >= 4 && n == 2
#pragma once
#include <cmath>
#include "image.hpp"

namespace orc
{
	struct Cong
	{
		const Image& im;
		const kb2_cg_node* nodes = nullptr; const uint32_t* keys = nullptr; const int32_t* values = nullptr; const int32_t* root = nullptr;
		const uint8_t* ctxEmb = nullptr; const uint8_t* outEmb = nullptr; const uint32_t* invVocab = nullptr; const float* outBias = nullptr;
		uint32_t dim = 0, stride = 0, keySize = 0;
		mutable WorkCounters* wc = nullptr;

		explicit Cong(const Image& _im) : im{ _im }
		{
			const auto* h = im.h;
			if (!h->cg_num_nodes) return;
			nodes = im.sec<kb2_cg_node>(KB2_SEC_CG_NODES); keys = im.sec<uint32_t>(KB2_SEC_CG_KEYS); values = im.sec<int32_t>(KB2_SEC_CG_VALUES);
			root = im.sec<int32_t>(KB2_SEC_CG_ROOT);
			ctxEmb = im.sec<uint8_t>(KB2_SEC_CG_CTX_EMB); outEmb = im.sec<uint8_t>(KB2_SEC_CG_OUT_EMB);
			invVocab = h->sec[KB2_SEC_CG_INV_VOCAB].nbytes ? im.sec<uint32_t>(KB2_SEC_CG_INV_VOCAB) : nullptr;
			outBias = h->sec[KB2_SEC_CG_OUT_BIAS].nbytes ? im.sec<float>(KB2_SEC_CG_OUT_BIAS) : nullptr;
			dim = h->cg_dim; stride = dim + 8; keySize = h->cg_key_size;
		}

		// nst::searchKV: value of `key` among the node's children, 0 when absent
		int32_t search(const kb2_cg_node& n, uint32_t key) const
		{
			if (wc) wc->lmProbes += ceilLog2p1(n.num_nexts);
			const uint32_t* k = keys + n.next_offset;
			const uint32_t* it = std::lower_bound(k, k + n.num_nexts, key);
			if (it == k + n.num_nexts || *it != key) return 0;
			return values[n.next_offset + (it - k)];
		}

		// CoNgramModel.hpp:314-385
		uint32_t stepVl(int32_t& nodeIdx, uint32_t next) const
		{
			while (1)
			{
				int32_t v;
				const kb2_cg_node* node = &nodes[nodeIdx];
				if (wc) wc->lmHops++;
				if (nodeIdx != 0)
				{
					if ((v = search(*node, next)) == 0)
					{
						if (!node->lower) return 0;
						nodeIdx += node->lower;
						continue;
					}
				}
				else
				{
					v = next < im.h->cg_root_size ? root[next] : 0;
					if (v == 0) return 0;
				}
				if (v > 0)
				{
					nodeIdx += v;
					return nodes[nodeIdx].value;
				}
				while (node->lower)
				{
					node += node->lower;
					int32_t lv;
					if (node != nodes)
					{
						if ((lv = search(*node, next)) != 0)
						{
							if (lv > 0)
							{
								node += lv;
								nodeIdx = (int32_t)(node - nodes);
								return (uint32_t)-v;
							}
						}
					}
					else
					{
						lv = next < im.h->cg_root_size ? root[next] : 0;
						if (lv > 0)
						{
							nodeIdx = lv;
							return (uint32_t)-v;
						}
					}
				}
				nodeIdx = 0;
				return (uint32_t)-v;
			}
		}

		// CoNgramModel.hpp:271-297 (keySize 3 = 16-bit keys, ids >= tMax are split into a surrogate pair)
		uint32_t step(int32_t& nodeIdx, uint32_t next) const
		{
			if (invVocab) next = invVocab[next];
			if (keySize != 3) return stepVl(nodeIdx, next);
			const uint32_t tMax = (1u << 16) - (1u << 10) * 2;
			if (next < tMax) return stepVl(nodeIdx, next);
			next -= tMax;
			const uint32_t high = next >> 10, low = next & 0x3FF;
			stepVl(nodeIdx, tMax + high);
			return stepVl(nodeIdx, tMax + (1u << 10) + low);
		}

		int32_t dotMinusHsum(uint32_t ctx, uint32_t wid) const
		{
			const uint8_t* a = ctxEmb + (size_t)ctx * stride;
			const int8_t* b = reinterpret_cast<const int8_t*>(outEmb + (size_t)wid * stride);
			int32_t acc = 0;
			for (uint32_t k = 0; k < dim; ++k) acc += (int32_t)a[k] * (int32_t)b[k];
			int32_t hsum; std::memcpy(&hsum, b + dim + 4, 4);
			return acc - hsum;
		}
		float ctxScale(uint32_t ctx) const { float f; std::memcpy(&f, ctxEmb + (size_t)ctx * stride + dim, 4); return f; }
		float ctxBias(uint32_t ctx) const { float f; std::memcpy(&f, ctxEmb + (size_t)ctx * stride + dim + 4, 4); return f; }
		float outScale(uint32_t wid) const { float f; std::memcpy(&f, outEmb + (size_t)wid * stride + dim, 4); return f; }

		enum Epilogue { E_scalar = 0, E_small = 1, E_gemv = 2 };
		// which kernel scatteredGEMMOpt<avx2> runs for m unique contexts x n unique outputs (ldc == n), qgemm.hpp:157-201
		static Epilogue epilogueOf(size_t m, size_t n)
		{
			if (m <= 3 && n <= 3) return E_small;
			if (n == 1) return E_gemv;
			if (m >= 4 && n == 2) return E_gemv;
			return E_small;              // scatteredGEMMBaseline inside the avx2 TU (GEMV3/GEMV4 are not specialised there)
		}
		float finish(uint32_t ctx, uint32_t wid, Epilogue e) const
		{
			const float x = (float)dotMinusHsum(ctx, wid);
			const float cs = ctxScale(ctx), os = outScale(wid), cb = ctxBias(ctx);
			switch (e)
			{
			case E_scalar: { volatile float t = x * cs; volatile float u = t * os; return u + cb; }
			case E_small: { volatile float t = x * cs; return std::fmaf(t, os, cb); }
			default: { volatile float t = x * os; return std::fmaf(t, cs, cb); }
			}
		}

		// CoNgramState::next -> CoNgramModel::progress, CoNgramModel.cpp:869-903
		float next(int32_t& node, uint32_t& ctx, uint32_t wid) const
		{
			if (wc) wc->lmSteps++;
			float ll = finish(ctx, wid, E_scalar);
			if (outBias) ll += outBias[wid];
			ctx = step(node, wid);
			return ll;
		}
	};
}
