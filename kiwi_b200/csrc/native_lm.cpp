// kiwi_b200: native readers of the reference's language-model files - the first pieces of a model loader that does not link the
// reference.  They produce the Knlm / SkipBigram SECTIONS of the model image (include/kiwi_b200_image.h) from the files
// the reference ships, byte for byte what oracle/ref_build/tools/flatten_model.cpp dumps from the reference's own in-memory model
// (tests/test_native_lm.py).  The rest of the image (morphemes, forms, the frozen form trie, the combining-rule engine's output) still
// comes from flatten_model.
//   sj.knlm          KnLangModel<...>::KnLangModel(MemoryObject&&)      /root/reference/src/Knlm.hpp:1003-1167, header include/kiwi/Knlm.h:10-16
//                    node-size codec QCode<0, 2, 8, 16>                  src/QEncoder.hpp:13-47,135-176,215-250; 8-bit tables src/Knlm.hpp:427-460
//   skipbigram.mdl   SkipBigramModel<...>::SkipBigramModel(...)          src/SkipBigramModel.hpp:40-105, header include/kiwi/SkipBigramModel.h:9-13
//   cong.mdl         CoNgramModel<..., windowSize 0, quantized>::CoNgramModel   src/CoNgramModel.cpp:425-790 (8-bit rows only), header
//                    include/kiwi/CoNgramModel.h:18-33; integer codec: streamvbyte (fast-pack/streamvbyte @7c472d7d, absent from the
//                    reference snapshot - its published format is restated below: 2-bit length codes first, then the little-endian bytes)
// Host code only; no CUDA.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <deque>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/kiwi_b200.h"
#include "../../include/kiwi_b200_image.h"

namespace
{
	struct KnHeader      // KnLangModelHeader
	{
		uint64_t num_nodes, node_offset, key_offset, ll_offset, gamma_offset, qtable_offset, htx_offset;
		uint64_t unk_id, bos_id, eos_id, vocab_size;
		uint8_t order, key_size, diff_size, quantized;
		uint32_t extra_buf_size;
	};
	static_assert(sizeof(KnHeader) == 96, "KnLangModelHeader");

	struct SbHeader { uint64_t vocabSize; uint8_t keySize, windowSize, compressed, quantize, rsv[4]; };
	static_assert(sizeof(SbHeader) == 16, "SkipBigramModelHeader");

	std::vector<char> readFile(const char* path)
	{
		std::ifstream f{ path, std::ios::binary };
		if (!f) throw std::runtime_error(std::string{ "cannot open " } + path);
		f.seekg(0, std::ios::end); const std::streamoff n = f.tellg(); f.seekg(0);
		std::vector<char> b((size_t)n);
		f.read(b.data(), n);
		if (!f) throw std::runtime_error(std::string{ "cannot read " } + path);
		return b;
	}

	uint64_t keyAt(const char* p, unsigned keySize, size_t i)
	{
		if (keySize == 1) return (uint8_t)p[i];
		if (keySize == 2) { uint16_t v; std::memcpy(&v, p + 2 * i, 2); return v; }
		if (keySize == 4) { uint32_t v; std::memcpy(&v, p + 4 * i, 4); return v; }
		uint64_t v; std::memcpy(&v, p + 8 * i, 8); return v;
	}

	// QCode<0, 2, 8, 16>::decode: a 2-bit class per value (4 per header byte), the payloads as one little-endian bit stream over
	// 64-bit words: class 0 = the value 0, class 1 = 2 bits + 1, class 2 = 8 bits + 5, class 3 = 16 bits + 261
	std::vector<uint32_t> decodeNodeSizes(const uint8_t* header, const uint64_t* body, size_t n, const char* end)
	{
		static const unsigned bits[4] = { 0, 2, 8, 16 };
		static const uint32_t bias[4] = { 0, 1, 5, 261 };
		std::vector<uint32_t> out(n);
		size_t u = 0; unsigned b = 0;
		for (size_t i = 0; i < n; ++i)
		{
			const unsigned q = (header[i / 4] >> ((i % 4) * 2)) & 3;
			uint64_t e = 0;
			if (bits[q])
			{
				if (reinterpret_cast<const char*>(body + u + 1) > end) throw std::runtime_error("sj.knlm: node-size stream runs past the section");
				if (b + bits[q] <= 64) e = (body[u] >> b) & ((1ull << bits[q]) - 1);
				else
				{
					if (reinterpret_cast<const char*>(body + u + 2) > end) throw std::runtime_error("sj.knlm: node-size stream runs past the section");
					e = body[u] >> b;
					e |= (body[u + 1] & ((1ull << (bits[q] + b - 64)) - 1)) << (64 - b);
				}
				b += bits[q];
				if (b >= 64) { b -= 64; ++u; }
			}
			out[i] = (uint32_t)e + bias[q];
		}
		return out;
	}

	struct Knlm
	{
		std::vector<kb2_kn_node> nodes; std::vector<uint32_t> keys; std::vector<int32_t> values; std::vector<int32_t> root; std::vector<uint32_t> htx;
		uint32_t order = 0, vocab = 0, htxVocab = 0; int32_t bosNode = 0; float unkLl = 0;

		// exact lookup among a node's children (nst::search; the file keeps them ascending)
		bool search(uint32_t node, uint32_t key, int32_t& v) const
		{
			const kb2_kn_node& n = nodes[node];
			size_t lo = n.next_offset, hi = lo + n.num_nexts;
			while (lo < hi) { const size_t mid = (lo + hi) / 2; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
			if (lo < (size_t)n.next_offset + n.num_nexts && keys[lo] == key) { v = values[lo]; return true; }
			return false;
		}
		static float asFloat(int32_t v) { float f; std::memcpy(&f, &v, 4); return f; }

		// KnLangModel::getLL (src/Knlm.cpp:10-43)
		float getLL(int64_t node, uint32_t next) const
		{
			for (int guard = 0; guard < 64; ++guard)
			{
				int32_t v;
				float acc = 0;
				if (node == 0)
				{
					v = next < root.size() ? root[next] : 0;
					if (v == 0) return unkLl;
				}
				else if (!search((uint32_t)node, next, v))
				{
					// gamma + getLL(lower): written as a loop; the partial sums are added in the reference's order (outermost gamma first)
					std::vector<float> gammas;
					int64_t cur = node;
					while (true)
					{
						gammas.push_back(nodes[cur].gamma);
						if (nodes[cur].lower == 0) throw std::runtime_error("sj.knlm: getLL before the suffix links exist would not terminate");
						cur += nodes[cur].lower;
						if (cur == 0) { v = next < root.size() ? root[next] : 0; if (v == 0) { acc = unkLl; goto fold; } break; }
						if (search((uint32_t)cur, next, v)) break;
					}
					acc = v > 0 ? nodes[cur + v].ll : asFloat(v);
				fold:
					for (size_t i = gammas.size(); i-- > 0;) acc = gammas[i] + acc;
					return acc;
				}
				return v > 0 ? nodes[node + v].ll : asFloat(v);
			}
			return 0;
		}

		// KnLangModel::progress (src/Knlm.cpp:45-130), the state transition only
		void progress(int64_t& node, uint32_t next) const
		{
			while (true)
			{
				int32_t v = 0;
				if (node == 0)
				{
					v = next < root.size() ? root[next] : 0;
					if (v == 0)
					{
						if (!htx.empty()) { int32_t lv; node = search(0, htx[next], lv) ? lv : 0; }
						return;
					}
				}
				else if (!search((uint32_t)node, next, v)) { node += nodes[node].lower; continue; }
				if (v > 0) { node += v; return; }
				// leaf: the deepest suffix state that has `next` as an inner child
				while (nodes[node].lower)
				{
					node += nodes[node].lower;
					int32_t lv;
					if (node == 0)
					{
						lv = next < root.size() ? root[next] : 0;
						if (lv > 0) { node += lv; return; }
					}
					else if (search((uint32_t)node, next, lv) && lv > 0) { node += lv; return; }
				}
				if (!htx.empty()) { int32_t lv; node = search(0, htx[next], lv) ? lv : 0; }
				else node = 0;
				return;
			}
		}
	};

	Knlm parseKnlm(const std::vector<char>& file)
	{
		if (file.size() < sizeof(KnHeader)) throw std::runtime_error("sj.knlm: too small");
		KnHeader h; std::memcpy(&h, file.data(), sizeof(h));
		const char* p = file.data(); const char* end = p + file.size();
		const unsigned quantized = h.quantized & 0x1F; const bool compressed = (h.quantized & 0x80) != 0;
		if (h.diff_size != 4) throw std::runtime_error("sj.knlm: diff size must be 4");
		if (h.key_size != 1 && h.key_size != 2 && h.key_size != 4 && h.key_size != 8) throw std::runtime_error("sj.knlm: key size");
		if (quantized != 0 && quantized != 8) throw std::runtime_error("sj.knlm: only unquantized and 8-bit quantized files are read natively");
		if (h.node_offset > file.size() || h.key_offset > file.size() || h.ll_offset > file.size() || h.gamma_offset > file.size() || h.qtable_offset > file.size() || h.htx_offset > file.size())
			throw std::runtime_error("sj.knlm: section offset beyond the file");
		if (h.num_nodes == 0 || h.num_nodes > (1ull << 31)) throw std::runtime_error("sj.knlm: node count");

		// node sizes (children per node; 0 = leaf), in depth-first order
		std::vector<uint32_t> sizes;
		if (compressed)
		{
			// (the reference decodes 16-bit values into a KeyType array, Knlm.hpp:1018-1025: meaningful for 2-byte keys only)
			if (h.key_size != 2) throw std::runtime_error("sj.knlm: compressed node sizes are defined for 2-byte keys");
			const uint8_t* qh = reinterpret_cast<const uint8_t*>(p + h.node_offset);
			const size_t hb = (h.num_nodes + 3) / 4;
			if (p + h.node_offset + hb > end) throw std::runtime_error("sj.knlm: node-size header beyond the file");
			sizes = decodeNodeSizes(qh, reinterpret_cast<const uint64_t*>(qh + hb), h.num_nodes, p + h.key_offset);
		}
		else
		{
			if (h.node_offset + h.num_nodes * h.key_size > file.size()) throw std::runtime_error("sj.knlm: node sizes beyond the file");
			sizes.resize(h.num_nodes);
			for (size_t i = 0; i < h.num_nodes; ++i) sizes[i] = (uint32_t)keyAt(p + h.node_offset, h.key_size, i);
		}
		size_t nonLeaf = 0, leaf = 0;
		for (uint32_t s : sizes) (s ? nonLeaf : leaf)++;
		if (!sizes[0]) throw std::runtime_error("sj.knlm: the root has no children");

		Knlm m;
		m.order = h.order; m.vocab = (uint32_t)h.vocab_size;
		const size_t nEdges = h.num_nodes - 1;
		if (h.key_offset + nEdges * h.key_size > file.size()) throw std::runtime_error("sj.knlm: keys beyond the file");
		m.keys.resize(nEdges);
		for (size_t i = 0; i < nEdges; ++i) m.keys[i] = (uint32_t)keyAt(p + h.key_offset, h.key_size, i);

		// log-likelihoods (non-leaf nodes first, then leaves) and back-off weights
		std::vector<float> ll(nonLeaf), gamma(nonLeaf), leafLl(leaf);
		if (quantized)
		{
			if (h.ll_offset + nonLeaf + leaf > file.size() || h.gamma_offset + nonLeaf > file.size() || h.qtable_offset + 512 * 4 > file.size()) throw std::runtime_error("sj.knlm: quantized sections beyond the file");
			const uint8_t* lq = reinterpret_cast<const uint8_t*>(p + h.ll_offset); const uint8_t* gq = reinterpret_cast<const uint8_t*>(p + h.gamma_offset);
			std::vector<float> table(512); std::memcpy(table.data(), p + h.qtable_offset, 512 * 4);
			for (size_t i = 0; i < nonLeaf; ++i) ll[i] = table[lq[i]];
			for (size_t i = 0; i < leaf; ++i) leafLl[i] = table[lq[nonLeaf + i]];
			for (size_t i = 0; i < nonLeaf; ++i) gamma[i] = table[256 + gq[i]];
		}
		else
		{
			if (h.ll_offset + (nonLeaf + leaf) * 4 > file.size() || h.gamma_offset + nonLeaf * 4 > file.size()) throw std::runtime_error("sj.knlm: float sections beyond the file");
			std::memcpy(ll.data(), p + h.ll_offset, nonLeaf * 4); std::memcpy(leafLl.data(), p + h.ll_offset + nonLeaf * 4, leaf * 4);
			std::memcpy(gamma.data(), p + h.gamma_offset, nonLeaf * 4);
		}
		if (h.htx_offset)
		{
			if (h.htx_offset + h.vocab_size * h.key_size > file.size()) throw std::runtime_error("sj.knlm: history transform beyond the file");
			m.htx.resize(h.vocab_size);
			uint32_t mx = 0;
			for (size_t i = 0; i < h.vocab_size; ++i) { m.htx[i] = (uint32_t)keyAt(p + h.htx_offset, h.key_size, i); mx = std::max(mx, m.htx[i]); }
			m.htxVocab = mx + 1;
		}
		else m.htxVocab = (uint32_t)h.vocab_size;

		// nodes: the depth-first size list unrolled with a stack of open key ranges (Knlm.hpp:1091-1126)
		m.nodes.resize(nonLeaf); m.values.assign(nEdges, 0);
		struct Range { size_t node, cur, end; };
		std::vector<Range> open;
		size_t ni = 0, li = 0, nextOff = 0;
		for (size_t i = 0; i < h.num_nodes; ++i)
		{
			if (sizes[i])
			{
				if (ni >= nonLeaf) throw std::runtime_error("sj.knlm: inconsistent node sizes");
				if (!open.empty()) m.values[open.back().cur] = (int32_t)(ni - open.back().node);
				kb2_kn_node& n = m.nodes[ni];
				n.num_nexts = sizes[i]; n.lower = 0; n.next_offset = (uint32_t)nextOff; n.ll = ll[ni]; n.gamma = gamma[ni];
				nextOff += sizes[i];
				if (nextOff > nEdges) throw std::runtime_error("sj.knlm: more children than keys");
				open.push_back(Range{ ni, n.next_offset, (size_t)n.next_offset + n.num_nexts });
				++ni;
			}
			else
			{
				if (open.empty()) throw std::runtime_error("sj.knlm: a leaf without a parent");
				int32_t bits; std::memcpy(&bits, &leafLl[li], 4);
				m.values[open.back().cur] = bits;
				open.back().cur++;
				while (open.back().cur == open.back().end)
				{
					open.pop_back();
					if (open.empty()) break;
					open.back().cur++;
				}
				++li;
			}
		}
		// direct table of the root's children
		m.root.assign(m.htxVocab, 0);
		for (uint32_t i = 0; i < m.nodes[0].num_nexts; ++i) { if (m.keys[i] >= m.root.size()) throw std::runtime_error("sj.knlm: root key beyond the vocabulary"); m.root[m.keys[i]] = m.values[i]; }
		for (size_t n = 0; n < nonLeaf; ++n) for (uint32_t j = 1; j < m.nodes[n].num_nexts; ++j)
			if (!(m.keys[m.nodes[n].next_offset + j - 1] < m.keys[m.nodes[n].next_offset + j])) throw std::runtime_error("sj.knlm: a node's keys are not ascending");

		// unk_ll and the BOS state are computed BEFORE the suffix links exist, exactly as the reference's constructor does (1128-1147)
		if (!m.htx.empty())
		{
			int64_t node = 0;
			m.progress(node, (uint32_t)h.bos_id);
			m.unkLl = m.getLL(node, (uint32_t)h.unk_id);
			int64_t b = 0; m.progress(b, m.htx[h.bos_id]); m.bosNode = (int32_t)b;
		}
		else
		{
			m.unkLl = m.getLL(0, (uint32_t)h.unk_id);
			int64_t b = 0; m.progress(b, (uint32_t)h.bos_id); m.bosNode = (int32_t)b;
		}
		// suffix links, breadth first (1149-1166): lower(child of p by key k) = the node reached by k from p's suffix chain
		std::deque<uint32_t> dq;
		for (dq.push_back(0); !dq.empty(); dq.pop_front())
		{
			const uint32_t pi = dq.front();
			const kb2_kn_node pn = m.nodes[pi];
			for (uint32_t i = 0; i < pn.num_nexts; ++i)
			{
				const uint32_t k = m.keys[pn.next_offset + i];
				const int32_t v = m.values[pn.next_offset + i];
				if (v <= 0) continue;
				const uint32_t child = pi + (uint32_t)v;
				// findLowerNode(p, k), Knlm.hpp:38-63
				int64_t node = pi; uint32_t key = k;
				while (m.nodes[node].lower)
				{
					const int64_t low = node + m.nodes[node].lower;
					if (low == 0 && !m.htx.empty()) key = m.htx[key];
					int32_t found;
					if (m.search((uint32_t)low, key, found)) { node = low + found; goto linked; }
					node = low;
				}
			linked:
				m.nodes[child].lower = (int32_t)(node - (int64_t)child);
				dq.push_back(child);
			}
		}
		return m;
	}

	struct Sbg { std::vector<uint32_t> ptrs, keys; std::vector<float> comps, discnts; std::vector<uint8_t> valid; uint32_t vocab = 0, window = 0; };

	Sbg parseSbg(const std::vector<char>& file)
	{
		if (file.size() < sizeof(SbHeader)) throw std::runtime_error("skipbigram.mdl: too small");
		SbHeader h; std::memcpy(&h, file.data(), sizeof(h));
		if (h.keySize != 1 && h.keySize != 2 && h.keySize != 4 && h.keySize != 8) throw std::runtime_error("skipbigram.mdl: key size");
		if (h.vocabSize == 0 || h.vocabSize > (1ull << 31)) throw std::runtime_error("skipbigram.mdl: vocabulary size");
		const char* p = file.data() + sizeof(SbHeader); const char* end = file.data() + file.size();
		auto need = [&](size_t n) { if ((size_t)(end - p) < n) throw std::runtime_error("skipbigram.mdl: truncated"); };
		Sbg m; m.vocab = (uint32_t)h.vocabSize; m.window = h.windowSize;
		need(h.vocabSize * h.keySize);
		m.ptrs.assign(h.vocabSize + 1, 0);
		for (size_t i = 0; i < h.vocabSize; ++i) m.ptrs[i + 1] = m.ptrs[i] + (uint32_t)keyAt(p, h.keySize, i);
		p += h.vocabSize * h.keySize;
		const size_t total = m.ptrs[h.vocabSize];
		need(total * h.keySize);
		m.keys.resize(total);
		for (size_t i = 0; i < total; ++i) m.keys[i] = (uint32_t)keyAt(p, h.keySize, i);
		p += total * h.keySize;
		m.comps.resize(total); m.discnts.resize(h.vocabSize); m.valid.resize(h.vocabSize);
		if (h.quantize)
		{
			need(h.vocabSize + total + h.vocabSize + 512 * 4);
			const uint8_t* dq = reinterpret_cast<const uint8_t*>(p); const uint8_t* cq = dq + h.vocabSize; const uint8_t* vv = cq + total;
			std::vector<float> table(512); std::memcpy(table.data(), vv + h.vocabSize, 512 * 4);
			for (size_t i = 0; i < h.vocabSize; ++i) { m.discnts[i] = table[dq[i]]; m.valid[i] = vv[i]; }
			for (size_t i = 0; i < total; ++i) m.comps[i] = table[256 + cq[i]];
		}
		else
		{
			need(h.vocabSize * 4 + total * 4 + h.vocabSize);
			std::memcpy(m.discnts.data(), p, h.vocabSize * 4); std::memcpy(m.comps.data(), p + h.vocabSize * 4, total * 4);
			std::memcpy(m.valid.data(), p + h.vocabSize * 4 + total * 4, h.vocabSize);
		}
		for (size_t i = 0; i < h.vocabSize; ++i) for (size_t j = m.ptrs[i] + 1; j < m.ptrs[i + 1]; ++j)
			if (!(m.keys[j - 1] < m.keys[j])) throw std::runtime_error("skipbigram.mdl: a target's history keys are not ascending");
		return m;
	}

	struct CgHeader      // CoNgramModelHeader
	{
		uint64_t vocabSize, contextSize;
		uint16_t dim, flags;
		uint8_t keySize, windowSize, qbit, qgroup;
		uint64_t numNodes, nodeOffset, keyOffset, valueOffset, embOffset;
	};
	static_assert(sizeof(CgHeader) == 64, "CoNgramModelHeader");

	// streamvbyte: ceil(n / 4) control bytes (2 bits per value, first value in the low bits), then the data bytes.  The classic code maps
	// the control value c to c + 1 bytes; the "0124" code maps 0, 1, 2, 3 to 0, 1, 2, 4 bytes.  Returns the bytes consumed.
	size_t svbDecode(const uint8_t* in, const uint8_t* end, uint32_t* out, size_t n, bool code0124)
	{
		const size_t nc = (n + 3) / 4;
		if (in + nc > end) throw std::runtime_error("cong.mdl: integer stream beyond the file");
		const uint8_t* data = in + nc;
		for (size_t i = 0; i < n; ++i)
		{
			const unsigned c = (in[i / 4] >> ((i % 4) * 2)) & 3;
			const unsigned len = code0124 ? (c == 3 ? 4u : c) : c + 1;
			if (data + len > end) throw std::runtime_error("cong.mdl: integer stream beyond the file");
			uint32_t v = 0;
			for (unsigned k = 0; k < len; ++k) v |= (uint32_t)data[k] << (8 * k);
			out[i] = v; data += len;
		}
		return (size_t)(data - in);
	}

	float halfToFloat(uint16_t hv)
	{
		const uint32_t sign = (uint32_t)(hv >> 15) << 31, ex = (hv >> 10) & 31, man = hv & 1023;
		uint32_t bits;
		if (ex == 0)
		{
			if (!man) bits = sign;
			else { int e = -1; uint32_t m = man; do { ++e; m <<= 1; } while (!(m & 1024)); bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((m & 1023) << 13); }
		}
		else if (ex == 31) bits = sign | 0x7F800000u | (man << 13);
		else bits = sign | ((ex - 15 + 127) << 23) | (man << 13);
		float f; std::memcpy(&f, &bits, 4); return f;
	}

	struct Cong
	{
		std::vector<kb2_cg_node> nodes; std::vector<uint32_t> keys; std::vector<int32_t> values, root;
		std::vector<uint8_t> ctxEmb, outEmb; std::vector<uint32_t> invVocab; std::vector<float> outBias;
		CgHeader h{};
		// child of `node` by key: > 0 inner child diff, < 0 leaf (-contextIdx), 0 = none (nst::searchKV)
		int32_t search(int64_t node, uint32_t key) const
		{
			const kb2_cg_node& n = nodes[node];
			size_t lo = n.next_offset, hi = lo + n.num_nexts;
			while (lo < hi) { const size_t mid = (lo + hi) / 2; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
			return (lo < (size_t)n.next_offset + n.num_nexts && keys[lo] == key) ? values[lo] : 0;
		}
	};

	Cong parseCong(const std::vector<char>& file)
	{
		if (file.size() < sizeof(CgHeader)) throw std::runtime_error("cong.mdl: too small");
		Cong m; std::memcpy(&m.h, file.data(), sizeof(CgHeader));
		const CgHeader& h = m.h;
		const uint8_t* p = reinterpret_cast<const uint8_t*>(file.data()); const uint8_t* end = p + file.size();
		if (h.qbit != 8) throw std::runtime_error("cong.mdl: only 8-bit embedding rows are read natively");
		if (h.keySize != 2 && h.keySize != 3 && h.keySize != 4) throw std::runtime_error("cong.mdl: key size");      // (1-byte VL keys: not met)
		if (h.dim == 0 || h.numNodes < 2 || h.numNodes > (1ull << 31) || h.vocabSize > (1ull << 31) || h.contextSize > (1ull << 31)) throw std::runtime_error("cong.mdl: header");
		if (h.nodeOffset > file.size() || h.keyOffset > file.size() || h.valueOffset > file.size() || h.embOffset > file.size()) throw std::runtime_error("cong.mdl: section offset beyond the file");
		if (h.flags & 4) throw std::runtime_error("cong.mdl: trie frequencies are not read natively");
		if (h.flags && h.windowSize) throw std::runtime_error("cong.mdl: optional sections behind window data are not read natively");

		std::vector<uint32_t> sizes(h.numNodes), keyData(h.numNodes - 1), vals(h.numNodes);
		svbDecode(p + h.nodeOffset, end, sizes.data(), h.numNodes, true);
		svbDecode(p + h.keyOffset, end, keyData.data(), h.numNodes - 1, false);
		svbDecode(p + h.valueOffset, end, vals.data(), h.numNodes, true);
		size_t nonLeaf = 0;
		for (uint32_t s : sizes) if (s) ++nonLeaf;
		if (!sizes[0]) throw std::runtime_error("cong.mdl: the root has no children");

		// nodes and (key, value) pairs, the pairs of every node ascending by key (the image's order)
		m.nodes.resize(nonLeaf); m.keys.resize(h.numNodes - 1); m.values.assign(h.numNodes - 1, 0);
		struct Range { size_t node, cur, end; };
		std::vector<Range> open;
		size_t ni = 0, nextOff = 0;
		for (size_t i = 0; i < h.numNodes; ++i)
		{
			if (sizes[i])
			{
				if (!open.empty()) m.values[open.back().cur] = (int32_t)(ni - open.back().node);
				kb2_cg_node& n = m.nodes[ni];
				n.lower = 0; n.value = vals[i]; n.next_offset = (uint32_t)nextOff; n.num_nexts = sizes[i];
				open.push_back(Range{ ni, nextOff, nextOff + sizes[i] });
				nextOff += sizes[i];
				if (nextOff > h.numNodes - 1) throw std::runtime_error("cong.mdl: more children than keys");
				++ni;
			}
			else
			{
				if (open.empty()) throw std::runtime_error("cong.mdl: a leaf without a parent");
				m.values[open.back().cur] = -(int32_t)vals[i];
				open.back().cur++;
				while (open.back().cur == open.back().end) { open.pop_back(); if (open.empty()) break; open.back().cur++; }
			}
		}
		for (size_t i = 0; i < h.numNodes - 1; ++i) m.keys[i] = keyData[i];
		for (size_t n = 0; n < nonLeaf; ++n)
		{
			const kb2_cg_node& nd = m.nodes[n];
			std::vector<std::pair<uint32_t, int32_t>> kv(nd.num_nexts);
			for (uint32_t j = 0; j < nd.num_nexts; ++j) kv[j] = { m.keys[nd.next_offset + j], m.values[nd.next_offset + j] };
			std::sort(kv.begin(), kv.end());
			for (uint32_t j = 0; j < nd.num_nexts; ++j) { m.keys[nd.next_offset + j] = kv[j].first; m.values[nd.next_offset + j] = kv[j].second; }
		}
		m.root.assign(h.vocabSize, 0);
		for (uint32_t i = 0; i < m.nodes[0].num_nexts; ++i) { if (m.keys[i] >= m.root.size()) throw std::runtime_error("cong.mdl: root key beyond the vocabulary"); m.root[m.keys[i]] = m.values[i]; }

		// suffix links and inherited context ids, breadth first (CoNgramModel.cpp:547-570)
		std::deque<uint32_t> dq;
		for (dq.push_back(0); !dq.empty(); dq.pop_front())
		{
			const uint32_t pi = dq.front();
			const kb2_cg_node pn = m.nodes[pi];
			for (uint32_t i = 0; i < pn.num_nexts; ++i)
			{
				const uint32_t k = m.keys[pn.next_offset + i]; const int32_t v = m.values[pn.next_offset + i];
				if (v <= 0) continue;
				const uint32_t child = pi + (uint32_t)v;
				int64_t node = pi;                                   // findLowerNode (CoNgramModel.hpp:186-203)
				while (m.nodes[node].lower)
				{
					const int64_t low = node + m.nodes[node].lower;
					const int32_t found = m.search(low, k);
					if (found > 0) { node = low + found; goto linked; }
					node = low;
				}
			linked:
				m.nodes[child].lower = (int32_t)(node - (int64_t)child);
				if (m.nodes[child].value == 0)                        // findLowerValue (205-229)
				{
					int64_t q = pi; uint32_t val = 0; bool got = false;
					while (m.nodes[q].lower)
					{
						const int64_t low = q + m.nodes[q].lower;
						const int32_t found = m.search(low, k);
						if (found != 0) { val = found > 0 ? m.nodes[low + found].value : (uint32_t)(-found); got = true; break; }
						q = low;
					}
					if (!got) val = m.nodes[q].value;
					m.nodes[child].value = val;
				}
				dq.push_back(child);
			}
		}

		// embedding rows (598-681): context rows u8 (= s8 + 128) + scale + bias, output rows s8 + scale + 128 * sum
		const size_t stride = (size_t)h.dim + 8;
		m.ctxEmb.assign(h.contextSize * stride, 0); m.outEmb.assign(h.vocabSize * stride, 0);
		const uint8_t* e = p + h.embOffset;
		auto need = [&](size_t n) { if ((size_t)(end - e) < n) throw std::runtime_error("cong.mdl: embeddings beyond the file"); };
		for (size_t i = 0; i < h.contextSize; ++i)
		{
			need((size_t)h.dim + 4 + (h.windowSize ? 4 : 0));
			uint8_t* o = m.ctxEmb.data() + i * stride;
			for (size_t d = 0; d < h.dim; ++d) o[d] = (uint8_t)((int8_t)e[d] + 128);
			uint16_t hs, hb; std::memcpy(&hs, e + h.dim, 2); std::memcpy(&hb, e + h.dim + 2, 2);
			const float scale = halfToFloat(hs), bias = -halfToFloat(hb);
			std::memcpy(o + h.dim, &scale, 4); std::memcpy(o + h.dim + 4, &bias, 4);
			e += (size_t)h.dim + 4 + (h.windowSize ? 4 : 0);
		}
		for (size_t i = 0; i < h.vocabSize; ++i)
		{
			need((size_t)h.dim + 2);
			uint8_t* o = m.outEmb.data() + i * stride;
			int32_t sum = 0;
			for (size_t d = 0; d < h.dim; ++d) { o[d] = e[d]; sum += (int8_t)e[d]; }
			uint16_t hs; std::memcpy(&hs, e + h.dim, 2);
			const float scale = halfToFloat(hs); const int32_t hsum = sum * 128;
			std::memcpy(o + h.dim, &scale, 4); std::memcpy(o + h.dim + 4, &hsum, 4);
			e += (size_t)h.dim + 2;
		}
		if (h.flags & 1)      // output bias: one byte per token, dequantised against the minimum stored behind them (765-774)
		{
			need(h.vocabSize + 2);
			uint16_t hm; std::memcpy(&hm, e + h.vocabSize, 2);
			const float minVal = halfToFloat(hm);
			m.outBias.resize(h.vocabSize);
			for (size_t i = 0; i < h.vocabSize; ++i) m.outBias[i] = (float)e[i] * (-minVal) / 255.f + minVal;
			e += h.vocabSize + 2;
		}
		if (h.flags & 2)      // reordered vocabulary (776-780): KeyType entries (2 bytes for key sizes 2 / 3, else 4)
		{
			const size_t ks = h.keySize == 4 ? 4 : 2;
			need(h.vocabSize * ks);
			m.invVocab.resize(h.vocabSize);
			for (size_t i = 0; i < h.vocabSize; ++i) m.invVocab[i] = (uint32_t)keyAt(reinterpret_cast<const char*>(e), (unsigned)ks, i);
		}
		return m;
	}

	template<class T> void put(std::vector<char>& blob, uint64_t& off, uint64_t& bytes, const std::vector<T>& v)
	{
		while (blob.size() % 16) blob.push_back(0);
		off = blob.size(); bytes = v.size() * sizeof(T);
		const char* p = reinterpret_cast<const char*>(v.data());
		blob.insert(blob.end(), p, p + bytes);
	}

	thread_local std::string g_nativeError;
}

extern "C" {
#pragma GCC visibility push(default)

const char* kiwi_b200_native_error(void) { return g_nativeError.c_str(); }

int kiwi_b200_native_knlm(const char* sj_knlm_path, void** out_bytes, uint64_t* out_size)
{
	try
	{
		const Knlm m = parseKnlm(readFile(sj_knlm_path));
		kiwi_b200_native_knlm_t hd{};
		hd.num_nodes = (uint32_t)m.nodes.size(); hd.num_edges = (uint32_t)m.keys.size(); hd.htx_vocab = m.htxVocab; hd.has_htx = m.htx.empty() ? 0 : 1;
		hd.order = m.order; hd.vocab_size = m.vocab; hd.bos_node = m.bosNode; hd.unk_ll = m.unkLl;
		std::vector<char> blob(sizeof(hd), 0);
		put(blob, hd.nodes_off, hd.nodes_bytes, m.nodes); put(blob, hd.keys_off, hd.keys_bytes, m.keys); put(blob, hd.values_off, hd.values_bytes, m.values);
		put(blob, hd.root_off, hd.root_bytes, m.root); put(blob, hd.htx_off, hd.htx_bytes, m.htx);
		std::memcpy(blob.data(), &hd, sizeof(hd));
		void* p = std::malloc(blob.size());
		if (!p) throw std::bad_alloc();
		std::memcpy(p, blob.data(), blob.size());
		*out_bytes = p; *out_size = blob.size();
		return 0;
	}
	catch (const std::exception& e) { g_nativeError = e.what(); return -1; }
}

int kiwi_b200_native_sbg(const char* skipbigram_mdl_path, void** out_bytes, uint64_t* out_size)
{
	try
	{
		const Sbg m = parseSbg(readFile(skipbigram_mdl_path));
		kiwi_b200_native_sbg_t hd{};
		hd.vocab_size = m.vocab; hd.window_size = m.window; hd.num_pairs = (uint32_t)m.keys.size();
		std::vector<char> blob(sizeof(hd), 0);
		put(blob, hd.ptrs_off, hd.ptrs_bytes, m.ptrs); put(blob, hd.keys_off, hd.keys_bytes, m.keys); put(blob, hd.comps_off, hd.comps_bytes, m.comps);
		put(blob, hd.discnts_off, hd.discnts_bytes, m.discnts); put(blob, hd.valid_off, hd.valid_bytes, m.valid);
		std::memcpy(blob.data(), &hd, sizeof(hd));
		void* p = std::malloc(blob.size());
		if (!p) throw std::bad_alloc();
		std::memcpy(p, blob.data(), blob.size());
		*out_bytes = p; *out_size = blob.size();
		return 0;
	}
	catch (const std::exception& e) { g_nativeError = e.what(); return -1; }
}

int kiwi_b200_native_cong(const char* cong_mdl_path, void** out_bytes, uint64_t* out_size)
{
	try
	{
		const Cong m = parseCong(readFile(cong_mdl_path));
		kiwi_b200_native_cong_t hd{};
		hd.num_nodes = (uint32_t)m.nodes.size(); hd.num_edges = (uint32_t)m.keys.size(); hd.root_size = (uint32_t)m.h.vocabSize; hd.dim = m.h.dim;
		hd.context_size = (uint32_t)m.h.contextSize; hd.key_size = m.h.keySize; hd.flags = m.h.flags; hd.vocab_size = (uint32_t)m.h.vocabSize;
		std::vector<char> blob(sizeof(hd), 0);
		put(blob, hd.nodes_off, hd.nodes_bytes, m.nodes); put(blob, hd.keys_off, hd.keys_bytes, m.keys); put(blob, hd.values_off, hd.values_bytes, m.values);
		put(blob, hd.root_off, hd.root_bytes, m.root); put(blob, hd.ctx_emb_off, hd.ctx_emb_bytes, m.ctxEmb); put(blob, hd.out_emb_off, hd.out_emb_bytes, m.outEmb);
		put(blob, hd.inv_vocab_off, hd.inv_vocab_bytes, m.invVocab); put(blob, hd.out_bias_off, hd.out_bias_bytes, m.outBias);
		std::memcpy(blob.data(), &hd, sizeof(hd));
		void* p = std::malloc(blob.size());
		if (!p) throw std::bad_alloc();
		std::memcpy(p, blob.data(), blob.size());
		*out_bytes = p; *out_size = blob.size();
		return 0;
	}
	catch (const std::exception& e) { g_nativeError = e.what(); return -1; }
}

#pragma GCC visibility pop
}
