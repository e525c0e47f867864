// ORACLE (test infrastructure): CPU restatement of the pre-lattice steps of the reference.
//   normalizeHangulWithPosition  /root/reference/src/StrUtils.h:493-520
//   normalizeCoda                src/StrUtils.h:637-710
//   matchPattern and its testers src/PatternMatcher.cpp:54-384 (charset grammar: src/pattern.hpp:176-226)
#pragma once
#include "image.hpp"

namespace orc
{
	using u16 = uint16_t;

	// src/StrUtils.h:493-520
	inline void normalizeHangulWithPosition(const u16* first, const u16* last, std::vector<u16>& out, std::vector<uint32_t>& pos)
	{
		uint32_t s = 0;
		for (; first != last; ++first)
		{
			u16 c = *first;
			pos.push_back(s);
			if (c == 0xB42C) c = 0xB410;
			if (0xAC00 <= c && c < 0xD7A4)
			{
				const int coda = (c - 0xAC00) % 28;
				out.push_back((u16)(c - coda)); s++;
				if (coda) { out.push_back((u16)(coda + 0x11A7)); s++; }
			}
			else { out.push_back(c); s++; }
		}
		pos.push_back(s);
	}

	// src/StrUtils.h:637-710
	inline void normalizeCoda(std::vector<u16>& s)
	{
		static const u16 codaToOnset[27] = {
			0x3131, 0x3131, 0x3145, 0x3134, 0x3148, 0x314E, 0x3137, 0x3139, 0x3131, 0x3141, 0x3142, 0x3145, 0x314C, 0x314D,
			0x314E, 0x3141, 0x3142, 0x3145, 0x3145, 0x3145, 0x3147, 0x3148, 0x314A, 0x314B, 0x314C, 0x314D, 0x314E };
		static const u16 codaConv[27] = {
			0, 0x11A8, 0x11A8, 0, 0x11AB, 0x11AB, 0, 0, 0x11AF, 0x11AF, 0x11AF, 0x11AF, 0x11AF, 0x11AF,
			0x11AF, 0, 0, 0x11B8, 0, 0x11BA, 0, 0, 0, 0, 0, 0, 0 };
		u16 before = 0;
		for (size_t i = 0; i < s.size(); ++i)
		{
			if (0x11A8 <= before && before <= 0x11C2)
			{
				const int off = before - 0x11A8;
				if (s[i] == codaToOnset[off])
				{
					if (codaConv[off]) s[i - 1] = codaConv[off];
					else s[i - 1] = s[i];
				}
			}
			before = s[i];
		}
	}

	// ---- pattern matcher, src/PatternMatcher.cpp
	struct Pat
	{
		const Image& im;
		explicit Pat(const Image& _im) : im{ _im } {}

		static bool isAlpha(u16 c) { return ('A' <= c && c <= 'Z') || ('a' <= c && c <= 'z'); }     // :38-41
		static bool isUpperAlpha(u16 c) { return 'A' <= c && c <= 'Z'; }
		static bool isDigit(u16 c) { return ('0' <= c && c <= '9') || (0xff10 <= c && c <= 0xff19); } // :48-51
		static bool alnum(u16 c) { return isAlpha(c) || ('0' <= c && c <= '9'); }
		// charsets of PatternMatcherImpl::md (:16-21)
		static bool emailAccount(u16 c) { return alnum(c) || c == '-' || c == '.' || c == '_' || c == '%' || c == '+'; }
		static bool alphaNumDotDash(u16 c) { return alnum(c) || c == '-' || c == '.'; }
		static bool domain(u16 c) { return alnum(c) || c == '-' || c == '@' || c == ':' || c == '%' || c == '.' || c == '_' || c == '+' || c == '~' || c == '#' || c == '='; }
		static bool path(u16 c) { return alnum(c) || c == '-' || c == '(' || c == ')' || c == '@' || c == ':' || c == '%' || c == '_' || c == '+' || c == '.' || c == '~' || c == '#' || c == '!' || c == '?' || c == '&' || c == '/' || c == '='; }
		static bool hashtags(u16 c)
		{
			switch (c) { case '#': case ' ': case '\t': case '\n': case '\r': case '\v': case '\f': case '.': case ',': case '(': case ')': case '[': case ']': case '<': case '>': case '{': case '}': return false; }
			return true;
		}
		static bool spaceSet(u16 c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }
		static bool startsWith(const u16* f, const u16* l, const char* lit)
		{
			size_t n = std::strlen(lit);
			if ((size_t)(l - f) < n) return false;
			for (size_t i = 0; i < n; ++i) if (f[i] != (u16)lit[i]) return false;
			return true;
		}

		size_t testUrl(const u16* first, const u16* last) const                                  // :54-117
		{
			const u16* b = first;
			if (startsWith(first, last, "http://")) b = first + 7;
			else if (startsWith(first, last, "https://")) b = first + 8;
			else return 0;
			int state = 0;
			const u16* lastMatched = first;
			if (b == last || !domain(*b)) return 0;
			++b;
			for (; b != last && domain(*b); ++b)
			{
				if (*b == '.') state = 1;
				else if (isAlpha(*b))
				{
					if (state > 0) ++state;
					if (state >= 3) lastMatched = b + 1;
				}
				else state = 0;
			}
			if (lastMatched == first) return 0;
			b = lastMatched;
			if (b != last && *b == ':')
			{
				++b;
				if (b == last || !isDigit(*b)) return 0;
				++b;
				while (b != last && isDigit(*b)) ++b;
			}
			if (b != last && *b == '/')
			{
				++b;
				while (b != last && path(*b)) ++b;
			}
			else
			{
				if (b != last && !spaceSet(*b)) return 0;
			}
			if (b[-1] == '.' || b[-1] == ':') --b;
			return b - first;
		}

		size_t testEmail(const u16* first, const u16* last) const                                // :119-150
		{
			const u16* b = first;
			if (b == last || !emailAccount(*b)) return 0;
			++b;
			while (b != last && emailAccount(*b)) ++b;
			if (b == last || *b != '@') return 0;
			++b;
			int state = 0;
			const u16* lastMatched = first;
			if (b == last || !alphaNumDotDash(*b)) return 0;
			++b;
			for (; b != last && alphaNumDotDash(*b); ++b)
			{
				if (*b == '.') state = 1;
				else if (isAlpha(*b))
				{
					if (state > 0) ++state;
					if (state >= 3) lastMatched = b + 1;
				}
				else state = 0;
			}
			return lastMatched - first;
		}

		size_t testMention(const u16* first, const u16* last) const                              // :152-168
		{
			const u16* b = first;
			if (b == last || *b != '@') return 0;
			++b;
			if (b == last || !isAlpha(*b)) return 0;
			++b;
			while (b != last && emailAccount(*b)) ++b;
			if (b[-1] == '.' || b[-1] == '%' || b[-1] == '+' || b[-1] == '-') --b;
			if (b - first <= 3) return 0;
			return b - first;
		}

		size_t testHashtag(const u16* first, const u16* last) const                              // :170-183
		{
			const u16* b = first;
			if (b == last || *b != '#') return 0;
			++b;
			if (b == last || !hashtags(*b)) return 0;
			++b;
			while (b != last && hashtags(*b)) ++b;
			return b - first;
		}

		size_t testNumeric(u16 left, const u16* first, const u16* last) const                    // :185-219
		{
			const u16* b = first;
			bool hasComma = false;
			if (b == last || !isDigit(*b)) return 0;
			while (b != last && isDigit(*b)) ++b;
			while (b != last && *b == ',')
			{
				++b;
				if (b + 2 >= last || !isDigit(b[0]) || !isDigit(b[1]) || !isDigit(b[2])) return b - 1 - first;
				b += 3;
				hasComma = true;
			}
			if (b == last || im.isSpace(*b) || isHangulSyllable(*b)) return b - first;
			if (*b == '.')
			{
				++b;
				if (!hasComma && !alphaNumDotDash(left) && (b == last || !alphaNumDotDash(*b))) return b - first;
				if (b == last || !isDigit(*b)) return b - 1 - first;
				while (b != last && isDigit(*b)) ++b;
			}
			if (b == last || (*b != '.')) return b - first;
			return 0;
		}

		size_t testSerial(const u16* first, const u16* last) const                               // :221-258
		{
			const u16* b = first;
			if (b == last || !isDigit(*b)) return 0;
			while (b != last && isDigit(*b)) ++b;
			if (b == last) return 0;
			u16 sep = 0;
			if (*b == ':' || *b == '.' || *b == '-' || *b == '/') sep = *b;
			else return 0;
			++b;
			if (b != last && *b == ' ') ++b;
			if (b == last || !isDigit(*b)) return 0;
			++b;
			while (b != last && isDigit(*b)) ++b;
			if (sep == '.' && (b == last || *b != sep)) return 0;
			while (b != last && *b == sep)
			{
				++b;
				if (b != last && *b == ' ') ++b;
				if (b == last || !isDigit(*b))
				{
					if (b[-1] == ' ') --b;
					return b - first;
				}
				++b;
				while (b != last && isDigit(*b)) ++b;
			}
			if (b[-1] == ' ') --b;
			return b - first;
		}

		size_t testAbbr(const u16* first, const u16* last) const                                 // :260-295
		{
			const u16* b = first;
			if (b == last || !isAlpha(*b)) return 0;
			size_t l = 0;
			while (b != last && isAlpha(*b)) ++b, ++l;
			if (b == last) return 0;
			if (*b == '.') ++b;
			else return 0;
			if (b != last && *b == ' ')
			{
				if (l > (isUpperAlpha(*first) ? 5u : 3u)) return 0;
				return b - first;
			}
			else
			{
				if (l > 5) return 0;
			}
			while (b != last && isAlpha(*b))
			{
				l = 0;
				while (b != last && isAlpha(*b)) ++b, ++l;
				if (l > 5) return 0;
				if (b != last && *b == '.') ++b;
				else return b - first;
			}
			if (b[-1] == ' ') --b;
			return b - first;
		}

		size_t testEmoji(const u16* first, const u16* last) const                                // :297-364
		{
			const u16* b = first;
			while (b + 1 < last)
			{
				uint32_t c0 = 0, c1 = 0;
				const u16* b1 = b;
				if (isHighSurrogate(*b1)) { c0 = mergeSurrogate(b1[0], b1[1]); b1 += 2; }
				else c0 = *b1++;
				const u16* b2 = b1;
				if (b2 < last)
				{
					if (isHighSurrogate(*b2) && b2 + 1 < last) { c1 = mergeSurrogate(b2[0], b2[1]); b2 += 2; }
					else c1 = *b2++;
				}
				const int r = im.isEmoji(c0, c1);
				if (r == 1) b = b1;
				else if (r == 2) b = b2;
				else break;
				if (b == last) return b - first;
				if (0xfe00 <= *b && *b <= 0xfe0f)
				{
					++b;
					if (b == last) return b - first;
				}
				else if (b + 1 < last && isHighSurrogate(b[0]))
				{
					c1 = mergeSurrogate(b[0], b[1]);
					if (0x1f3fb <= c1 && c1 <= 0x1f3ff)
					{
						b += 2;
						if (b == last) return b - first;
					}
				}
				if (*b == 0x200d) { ++b; continue; }
				break;
			}
			return b - first;
		}

		// returns (length, tag); match option bits: include/kiwi/PatternMatcher.h:12-17
		std::pair<size_t, uint8_t> match(u16 left, const u16* first, const u16* last, uint32_t opt) const   // :366-378
		{
			size_t size;
			if ((opt & (1 << 4)) && (size = testSerial(first, last))) return { size, T_w_serial };
			if ((size = testNumeric(left, first, last))) return { size, T_sn };
			if ((opt & (1 << 2)) && (size = testHashtag(first, last))) return { size, T_w_hashtag };
			if ((opt & (1 << 1)) && (size = testEmail(first, last))) return { size, T_w_email };
			if ((opt & (1 << 3)) && (size = testMention(first, last))) return { size, T_w_mention };
			if ((opt & (1 << 0)) && (size = testUrl(first, last))) return { size, T_w_url };
			if ((opt & (1 << 5)) && (size = testEmoji(first, last))) return { size, T_w_emoji };
			if ((size = testAbbr(first, last))) return { size, T_sl };
			return { 0, T_unknown };
		}
	};
}
