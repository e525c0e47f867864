// ORACLE (test infrastructure): CPU restatement of the SkipBigram model on top of Knlm,
//   SbgState::nextImpl / SkipBigramModel::evaluate   /root/reference/src/SkipBigramModel.hpp:113-185
//   logSumExp<avx2> over 16 floats                   src/MathFunc.hpp:12-32 with the AVX2 packet operators of
//                                                    src/SIMD.hpp:100-160 (ldexpf_fast, expf: Cephes polynomial with FMA),
//                                                    443-480 (redmaxbf, redsumf = (lo128 + hi128) then movehl / shuffle adds)
// The final `std::log` is the C library's logf, exactly as in the reference.  State = Knlm node + an 8-slot ring of the
// last valid tokens; equality compares node, ring position and the whole ring (SkipBigramModel.hpp:156-159).
#pragma once
#include <cmath>
#include "knlm.hpp"

namespace orc
{
	struct SbHist
	{
		uint8_t pos = 0;
		uint32_t hist[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
		bool operator==(const SbHist& o) const { return pos == o.pos && std::memcmp(hist, o.hist, sizeof(hist)) == 0; }
	};

	struct Sbg
	{
		const Image& im;
		const uint32_t* ptrs = nullptr; const uint32_t* keys = nullptr; const float* comps = nullptr; const float* discnts = nullptr; const uint8_t* valid = nullptr;
		float logWindowSize = 0;
		mutable WorkCounters* wc = nullptr;

		explicit Sbg(const Image& _im) : im{ _im }
		{
			if (!im.h->sb_vocab_size) return;
			ptrs = im.sec<uint32_t>(KB2_SEC_SB_PTRS); keys = im.sec<uint32_t>(KB2_SEC_SB_KEYS); comps = im.sec<float>(KB2_SEC_SB_COMPS);
			discnts = im.sec<float>(KB2_SEC_SB_DISCNTS); valid = im.sec<uint8_t>(KB2_SEC_SB_VALID);
			logWindowSize = std::log((float)im.h->sb_window_size);
		}

		// simd::OperatorBase<avx2>::expf on one lane (maddf = fused multiply-add)
		static float expLane(float _x)
		{
			float x = std::fmax(std::fmin(_x, 88.723f), -88.723f);
			const float m = std::floor(std::fmaf(x, 1.44269504088896341f, 0.5f));
			float r = std::fmaf(m, -0.693359375f, x);
			r = std::fmaf(m, 2.12194440e-4f, r);
			const float r2 = r * r, r3 = r2 * r;
			float y = std::fmaf(1.9875691500E-4f, r, 1.3981999507E-3f);
			float y1 = std::fmaf(4.1665795894E-2f, r, 1.6666665459E-1f);
			const float y2 = r + 1.0f;
			y = std::fmaf(y, r, 8.3334519073E-3f);
			y1 = std::fmaf(y1, r, 5.0000001201E-1f);
			y = std::fmaf(y, r3, y1);
			y = std::fmaf(y, r2, y2);
			// ldexpf_fast: y * 2^m with the biased exponent clamped to [0, 255]
			const int32_t e = (int32_t)std::nearbyint(std::fmin(std::fmax(m + 127.f, 0.f), 255.f));
			const uint32_t bits = (uint32_t)e << 23;
			float p; std::memcpy(&p, &bits, 4);
			const float v = y * p;
			return v > _x ? v : _x;          // _mm256_max_ps(a, b): a > b ? a : b
		}

		static float logSumExp16(const float* arr)
		{
			float mx = arr[0];
			for (int i = 1; i < 16; ++i) mx = arr[i] > mx ? arr[i] : mx;
			float s[8];
			for (int i = 0; i < 8; ++i) s[i] = 0.f + expLane(arr[i] - mx);
			for (int i = 0; i < 8; ++i) s[i] = s[i] + expLane(arr[8 + i] - mx);
			const float t0 = s[0] + s[4], t1 = s[1] + s[5], t2 = s[2] + s[6], t3 = s[3] + s[7];
			const float sum = (t0 + t2) + (t1 + t3);
			return std::log(sum) + mx;
		}

		bool isValidVocab(uint32_t k) const { return k < im.h->sb_vocab_size && valid[k]; }

		// SkipBigramModel::evaluate with cnt == windowSize == 8
		float evaluate(const uint32_t* history, uint32_t next, float base) const
		{
			if (!valid[next]) return base;
			alignas(32) float arr[16];
			for (int i = 0; i < 8; ++i) arr[i] = base;
			for (int i = 8; i < 16; ++i) arr[i] = -INFINITY;
			const uint32_t b = ptrs[next], e = ptrs[next + 1];
			for (int i = 0; i < 8; ++i)
			{
				arr[i] = discnts[history[i]] + base;
				const uint32_t* it = std::lower_bound(keys + b, keys + e, history[i]);
				if (wc) wc->lmProbes += ceilLog2p1(e - b);
				if (it != keys + e && *it == history[i]) arr[i + 8] = comps[it - keys];
			}
			return logSumExp16(arr) - logWindowSize;
		}

		// SbgState::nextImpl
		float next(const Knlm& lm, int32_t& node, SbHist& st, uint32_t wid) const
		{
			float ll = lm.progress(node, wid);
			if (isValidVocab(wid))
			{
				const float ll0 = ll;
				if (ll > -13) ll = evaluate(st.hist, wid, ll);
				if (std::getenv("ORC_TRACE_SBG")) std::fprintf(stderr, "[orc] sbg wid %u base %a -> %a hist %u %u %u %u %u %u %u %u pos %u\n", wid, ll0, ll, st.hist[0], st.hist[1], st.hist[2], st.hist[3], st.hist[4], st.hist[5], st.hist[6], st.hist[7], (unsigned)st.pos);
				st.hist[st.pos] = wid;
				st.pos = (uint8_t)((st.pos + 1) % 8);
			}
			return ll;
		}
	};
}
