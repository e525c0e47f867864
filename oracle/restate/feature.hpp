// ORACLE (test infrastructure): FeatureTestor::isMatched restated, /root/reference/src/FeatureTestor.cpp:6-80
#pragma once
#include "image.hpp"

namespace orc
{
	using u16 = uint16_t;
	// ---- FeatureTestor, src/FeatureTestor.cpp
	inline bool ftVowel(const u16* b, const u16* e, uint8_t vowel)                          // :6-60
	{
		if (vowel == CV_none) return true;
		if (b == e) return false;
		if (vowel == CV_any) return true;
		const u16 c = e[-1];
		if (vowel == CV_applosive)
		{
			switch (c) { case 0x11A8: case 0x11A9: case 0x11AA: case 0x11AE: case 0x11B8: case 0x11B9: case 0x11BA: case 0x11BB: case 0x11BD: case 0x11BE: case 0x11BF: case 0x11C0: case 0x11C1: return true; }
			return false;
		}
		if (!(0xAC00 <= c && c <= 0xD7A4) && !(0x11A8 <= c && c <= 0x11C2)) return true;
		switch (vowel)
		{
		case CV_vocalic_h: if (c == 0x11C2) return true; [[fallthrough]];
		case CV_vocalic: if (c == 0x11AF) return true; [[fallthrough]];
		case CV_vowel: if (0x11A8 <= c && c <= 0x11C2) return false; return true;
		case CV_non_vocalic_h: if (c == 0x11C2) return false; [[fallthrough]];
		case CV_non_vocalic: if (c == 0x11AF) return false; [[fallthrough]];
		case CV_non_vowel: if (0xAC00 <= c && c <= 0xD7A4) return false; return true;
		default: return false;
		}
	}
	inline bool ftPolar(const u16* b, const u16* e, uint8_t polar)                           // :62-80
	{
		if (polar == CP_none || polar == CP_non_adj) return true;
		if (b == e) return true;
		for (const u16* it = e - 1; it >= b; --it)
		{
			const u16 c = *it;
			if (0x11A8 <= c && c <= 0x11C2) continue;
			if (c == 0x1161 || c == 0x1163 || c == 0x1169 || c == 0x116D || c == 0x119E) return polar == CP_positive;
			if (!(0xAC00 <= c && c <= 0xD7A4)) break;
			const int v = ((c - 0xAC00) / 28) % 21;
			if (v == 0 || v == 2 || v == 8 || v == 12) return polar == CP_positive;
			if (v == 18 && it == e - 1) continue;
			return polar == CP_negative;
		}
		return polar == CP_negative;
	}

}
