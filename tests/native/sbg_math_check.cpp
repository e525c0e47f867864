// CPU check (tests/test_emulations.py): kiwi_b200/csrc/sbg_math.h against the C library's logf for every float in [1, 16] and against
// the oracle's logSumExp16 (oracle/restate/sbg.hpp: the same restatement with std::log as the reference calls it) on random score arrays.
#include "../../kiwi_b200/csrc/sbg_math.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

static float expLaneRef(float _x)      // oracle/restate/sbg.hpp Sbg::expLane, verbatim arithmetic
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
	const int32_t e = (int32_t)std::nearbyint(std::fmin(std::fmax(m + 127.f, 0.f), 255.f));
	const uint32_t bits = (uint32_t)e << 23;
	float p; std::memcpy(&p, &bits, 4);
	const float v = y * p;
	return v > _x ? v : _x;
}

int main()
{
	long bad = 0, n = 0;
	for (uint32_t u = 0x3f800000u; u <= 0x41800000u; ++u)
	{
		float x; std::memcpy(&x, &u, 4);
		volatile float xv = x;      // (keep the compiler from folding the library call)
		const float a = kb::sbgLogf(x), b = std::log((float)xv);
		if (std::memcmp(&a, &b, 4)) { if (bad < 5) std::printf("logf mismatch at %a: %a vs %a\n", x, a, b); ++bad; }
		++n;
	}
	std::mt19937 rng(7);
	std::uniform_real_distribution<float> d(-30.f, 0.f);
	long lse = 0;
	for (int t = 0; t < 2000000; ++t)
	{
		float arr[16];
		const float base = d(rng);
		for (int i = 0; i < 8; ++i) arr[i] = base + d(rng) * 0.2f;
		for (int i = 8; i < 16; ++i) arr[i] = (rng() & 3) ? -INFINITY : d(rng);
		float mx = arr[0];
		for (int i = 1; i < 16; ++i) mx = arr[i] > mx ? arr[i] : mx;
		float s[8];
		for (int i = 0; i < 8; ++i) s[i] = 0.f + expLaneRef(arr[i] - mx);
		for (int i = 0; i < 8; ++i) s[i] = s[i] + expLaneRef(arr[8 + i] - mx);
		const float sum = ((s[0] + s[4]) + (s[2] + s[6])) + ((s[1] + s[5]) + (s[3] + s[7]));
		volatile float sv = sum;
		const float ref = std::log((float)sv) + mx, got = kb::sbgLogSumExp16(arr);
		if (std::memcmp(&ref, &got, 4)) { if (bad < 10) std::printf("logSumExp mismatch: %a vs %a\n", got, ref); ++bad; }
		++lse;
	}
	std::printf("logf values %ld logsumexp arrays %ld mismatches %ld\n", n, lse, bad);
	return bad ? 1 : 0;
}
