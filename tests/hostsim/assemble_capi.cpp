// TEST INFRASTRUCTURE: C entry over the product's host-side result assembly (kiwi_b200/csrc/assemble.h) so that the CPU
// suite can run it on the reference's golden token lists without a GPU.  Built by oracle/Makefile (g++, host only).
#include <cstdint>
#include <string>
#include <vector>
#include "../../kiwi_b200/csrc/assemble.h"

extern "C" int kb_asm_run(int n, const uint32_t* pos, const uint32_t* len, const uint8_t* tag, const uint16_t* formBlob, const uint32_t* formOff,
	const uint8_t* isYo, const uint32_t* wordPosIn, const uint16_t* text, uint32_t textLen, int32_t* out)
{
	std::vector<kb::AsmTok> t((size_t)n);
	for (int i = 0; i < n; ++i)
	{
		t[i].position = pos[i]; t[i].length = len[i]; t[i].tag = tag[i];
		t[i].form.assign(reinterpret_cast<const char16_t*>(formBlob) + formOff[i], formOff[i + 1] - formOff[i]);
		t[i].kformIsYo = isYo[i] != 0; t[i].wordPosition = wordPosIn[i];
	}
	kb::fillPaired(t);
	kb::fillSentLine(t, kb::newlinePositions(text, textLen));
	for (int i = 0; i < n; ++i)
	{
		out[5 * i + 0] = (int32_t)t[i].wordPosition; out[5 * i + 1] = (int32_t)t[i].sentPosition; out[5 * i + 2] = (int32_t)t[i].lineNumber;
		out[5 * i + 3] = (int32_t)t[i].subSentPosition; out[5 * i + 4] = t[i].pairedToken == 0xFFFFFFFFu ? -1 : (int32_t)t[i].pairedToken;
	}
	return 0;
}

// surface form of an own-substring token: returns the number of UTF-16 units written to out
extern "C" int kb_own_form(const uint16_t* text, uint32_t textLen, int normalizeCoda, uint32_t position, uint32_t length, int beginsAtCoda, int endsBeforeCoda,
	uint16_t* out, int cap)
{
	const kb::NormText nt = kb::normalizeWithPosition(text, textLen, normalizeCoda != 0);
	const std::u16string f = kb::ownSubstringForm(nt, position, length, beginsAtCoda != 0, endsBeforeCoda != 0);
	if ((int)f.size() > cap) return -1;
	for (size_t i = 0; i < f.size(); ++i) out[i] = (uint16_t)f[i];
	return (int)f.size();
}
