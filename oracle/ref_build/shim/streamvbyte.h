// Scalar restatement of the StreamVByte codec (fast-pack/streamvbyte @7c472d7d, an EMPTY submodule in the
// reference snapshot).  Published format (Lemire, Kurz, Rupp, "Stream VByte", IPL 2018): ceil(n/4) control
// bytes, two bits per integer (LSB first), followed by the data bytes, little endian.
//   classic 1234 variant: code c -> c+1 bytes.      0124 variant: code {0,1,2,3} -> {0,1,2,4} bytes.
// TEST INFRASTRUCTURE ONLY (lets src/CoNgramModel.cpp compile into oracle/_ref).  Call sites:
// src/CoNgramModel.cpp:436,452,461,465 (decode) and 1854-1859,2254-2255 (encode).
#pragma once
#include <cstddef>
#include <cstdint>

static inline size_t streamvbyte_max_compressedbytes(size_t length)
{
	return (length + 3) / 4 + length * 4;
}

static inline size_t svb_shim_encode(const uint32_t* in, uint32_t length, uint8_t* out, bool v0124)
{
	uint8_t* ctrl = out;
	uint8_t* data = out + (length + 3) / 4;
	for (uint32_t i = 0; i < length; ++i)
	{
		if ((i & 3) == 0) ctrl[i >> 2] = 0;
		const uint32_t v = in[i];
		unsigned code, nbytes;
		if (v0124)
		{
			if (v == 0) { code = 0; nbytes = 0; }
			else if (v < (1u << 8)) { code = 1; nbytes = 1; }
			else if (v < (1u << 16)) { code = 2; nbytes = 2; }
			else { code = 3; nbytes = 4; }
		}
		else
		{
			if (v < (1u << 8)) { code = 0; nbytes = 1; }
			else if (v < (1u << 16)) { code = 1; nbytes = 2; }
			else if (v < (1u << 24)) { code = 2; nbytes = 3; }
			else { code = 3; nbytes = 4; }
		}
		ctrl[i >> 2] |= (uint8_t)(code << ((i & 3) * 2));
		for (unsigned b = 0; b < nbytes; ++b) *data++ = (uint8_t)(v >> (8 * b));
	}
	return (size_t)(data - out);
}

static inline size_t svb_shim_decode(const uint8_t* in, uint32_t* out, uint32_t length, bool v0124)
{
	const uint8_t* ctrl = in;
	const uint8_t* data = in + (length + 3) / 4;
	for (uint32_t i = 0; i < length; ++i)
	{
		const unsigned code = (ctrl[i >> 2] >> ((i & 3) * 2)) & 3;
		const unsigned nbytes = v0124 ? (code == 3 ? 4 : code) : code + 1;
		uint32_t v = 0;
		for (unsigned b = 0; b < nbytes; ++b) v |= (uint32_t)(*data++) << (8 * b);
		out[i] = v;
	}
	return (size_t)(data - in);
}

static inline size_t streamvbyte_encode(const uint32_t* in, uint32_t length, uint8_t* out) { return svb_shim_encode(in, length, out, false); }
static inline size_t streamvbyte_decode(const uint8_t* in, uint32_t* out, uint32_t length) { return svb_shim_decode(in, out, length, false); }
static inline size_t streamvbyte_encode_0124(const uint32_t* in, uint32_t length, uint8_t* out) { return svb_shim_encode(in, length, out, true); }
static inline size_t streamvbyte_decode_0124(const uint8_t* in, uint32_t* out, uint32_t length) { return svb_shim_decode(in, out, length, true); }
