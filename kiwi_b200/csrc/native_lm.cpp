// kiwi_b200: native readers of the reference's language-model files - the first pieces of a model loader that does not link the
// reference.  They produce the Knlm / SkipBigram SECTIONS of the model image (include/kiwi_b200_image.h) from the files
// the reference ships, byte for byte what oracle/ref_build/tools/flatten_model.cpp dumps from the reference's own in-memory model
// (tests/test_native_lm.py).  The rest of the image (morphemes, forms, the frozen form trie, the combining-rule engine's output) still
// comes from flatten_model.
//   sj.knlm          KnLangModel<...>::KnLangModel(MemoryObject&&)      /root/reference/src/Knlm.hpp:1003-1167, header include/kiwi/Knlm.h:10-16
//                    node-size codec QCode<0, 2, 8, 16>                  src/QEncoder.hpp:13-47,135-176,215-250; 8-bit tables src/Knlm.hpp:427-460
//   skipbigram.mdl   SkipBigramModel<...>::SkipBigramModel(...)          src/SkipBigramModel.hpp:40-105, header include/kiwi/SkipBigramModel.h:9-13
// Host code only; no CUDA.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

#pragma GCC visibility pop
}
