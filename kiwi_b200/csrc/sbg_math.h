// kiwi_b200: the float arithmetic of the SkipBigram score, restated operation by operation so that device results equal the
// reference's AVX2 build bit for bit:
//   logSumExp<avx2> over 16 floats   /root/reference/src/MathFunc.hpp:12-32 with the packet operators of src/SIMD.hpp:100-160
//                                    (expf: Cephes polynomial with fused multiply-adds, ldexpf_fast) and 443-480 (redmaxbf, redsumf)
//   the final std::log               glibc's logf as the FMA-capable x86-64 hosts run it (sysdeps/ieee754/flt-32/e_logf.c built with
//                                    -mfma: every multiply-add of the polynomial is fused; 16-entry table __logf_data).  The table and
//                                    the operation order were read off this image's libm.so.6 (2.39); tests/native/emu_check.cpp compares
//                                    sbgLogf with the C library for every float in [1, 16] - the only range a sum of 16 exponentials
//                                    with maximum exp(0) can take.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define KB_SM_HD __host__ __device__ inline
#else
#define KB_SM_HD inline
#endif

namespace kb
{
	KB_SM_HD uint32_t sbgBits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
	KB_SM_HD float sbgFloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

	// logf for 1 <= x <= 16 (no zero / subnormal / negative / inf / nan handling: the caller's sum is in that range)
	KB_SM_HD float sbgLogf(float x)
	{
		// { invc, logc } of __logf_data.tab, then ln2 and the polynomial
		const double T[16][2] = {
			{ 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
			{ 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2 }, { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
			{ 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3 },
			{ 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
			{ 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1p+0, 0x0p+0 },
			{ 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 }, { 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4 },
			{ 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 }, { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },
			{ 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 }, { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 } };
		const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
		const uint32_t ix = sbgBits(x);
		if (ix == 0x3f800000u) return 0.f;
		const uint32_t tmp = ix - 0x3f330000u;
		const uint32_t i = (tmp >> 19) & 15u;
		const int32_t k = (int32_t)tmp >> 23;
		const uint32_t iz = ix - (tmp & 0xff800000u);
		const double z = (double)sbgFloat(iz);
		const double y0 = fma((double)k, Ln2, T[i][1]);
		const double r = fma(z, T[i][0], -1.0);
		double y = fma(A1, r, A2);
		const double r2 = r * r;
		const double t = r + y0;
		y = fma(A0, r2, y);
		y = fma(r2, y, t);
		return (float)y;
	}

	// simd::OperatorBase<avx2>::expf on one lane
	KB_SM_HD float sbgExpLane(float _x)
	{
		const float x = fmaxf(fminf(_x, 88.723f), -88.723f);
		const float m = floorf(fmaf(x, 1.44269504088896341f, 0.5f));
		float r = fmaf(m, -0.693359375f, x);
		r = fmaf(m, 2.12194440e-4f, r);
		const float r2 = r * r, r3 = r2 * r;
		float y = fmaf(1.9875691500E-4f, r, 1.3981999507E-3f);
		float y1 = fmaf(4.1665795894E-2f, r, 1.6666665459E-1f);
		const float y2 = r + 1.0f;
		y = fmaf(y, r, 8.3334519073E-3f);
		y1 = fmaf(y1, r, 5.0000001201E-1f);
		y = fmaf(y, r3, y1);
		y = fmaf(y, r2, y2);
		// ldexpf_fast: y * 2^m with the biased exponent clamped to [0, 255] (m is integral)
		const int32_t e = (int32_t)fminf(fmaxf(m + 127.f, 0.f), 255.f);
		const float p = sbgFloat((uint32_t)e << 23);
		const float v = y * p;
		return v > _x ? v : _x;      // _mm256_max_ps(a, b)
	}

	// logSumExp<avx2>(arr, 16): maximum, two packets of exp(arr - max) added lane-wise, the horizontal sum (lo128 + hi128, then
	// movehl / shuffle adds), log, + max
	KB_SM_HD float sbgLogSumExp16(const float* arr)
	{
		float mx = arr[0];
		for (int i = 1; i < 16; ++i) mx = arr[i] > mx ? arr[i] : mx;
		float s[8];
		for (int i = 0; i < 8; ++i) s[i] = 0.f + sbgExpLane(arr[i] - mx);
		for (int i = 0; i < 8; ++i) s[i] = s[i] + sbgExpLane(arr[8 + i] - mx);
		const float t0 = s[0] + s[4], t1 = s[1] + s[5], t2 = s[2] + s[6], t3 = s[3] + s[7];
		const float sum = (t0 + t2) + (t1 + t3);
		return sbgLogf(sum) + mx;
	}
}
