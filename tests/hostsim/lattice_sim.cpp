// HOST SIMULATION of the lattice kernel (test infrastructure, tests/ only; see shim/cuda_runtime.h): the SAME sources as the
// device build (kiwi_b200/csrc/lattice.cu, model.cu) compiled as C++ with one lane per sentence, behind a small C API so that
// tests/test_hostsim_lattice.py can compare the kernel's logic with the reference's golden lattices without a GPU.
#include <cuda_runtime.h>
#include "../../kiwi_b200/csrc/model.cu"
#include "../../kiwi_b200/csrc/lattice.cu"
#include <memory>

namespace
{
	struct Sim
	{
		kb::Model model;
		std::vector<char> typo;          // flat typo image (include/kiwi_b200_typo.h), empty = none
		float typoThreshold = 2.5f;
		uint32_t graphPerUnit = 16, statesPerUnit = 16;
		uint32_t maxGraph = 0, maxStates = 0, maxUnits = 0;      // high-water marks per W unit, in 1/100
	};
	template<class T> std::vector<T> buf(size_t n) { return std::vector<T>(n); }
}

extern "C" {

void* hs_open(const char* imagePath)
{
	try
	{
		auto blob = kb::readImageFile(imagePath);
		auto* s = new Sim;
		s->model.load(blob.data(), blob.size());
		return s;
	}
	catch (...) { return nullptr; }
}
void hs_close(void* p) { delete reinterpret_cast<Sim*>(p); }

// analyse with the typo lattice of a flat typo image from now on (path == nullptr: off)
int hs_set_typo(void* p, const char* path, float threshold, uint32_t graphPerUnit, uint32_t statesPerUnit)
{
	Sim& s = *reinterpret_cast<Sim*>(p);
	s.typo.clear();
	if (!path) return 0;
	FILE* f = std::fopen(path, "rb");
	if (!f) return -1;
	std::fseek(f, 0, SEEK_END); const long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
	s.typo.resize((size_t)n);
	const bool ok = std::fread(s.typo.data(), 1, (size_t)n, f) == (size_t)n;
	std::fclose(f);
	if (!ok || reinterpret_cast<const kb2_typo_header*>(s.typo.data())->magic != KB2_TYPO_MAGIC) { s.typo.clear(); return -1; }
	s.typoThreshold = threshold; s.graphPerUnit = graphPerUnit; s.statesPerUnit = statesPerUnit;
	return 0;
}

// rows of 9 int32 as kiwi_b200_debug_lattice / orc_lattice; returns the node count, -status on a kernel status, -100 on error
int hs_lattice(void* p, const uint16_t* text, int len, uint32_t matchOptions, int32_t* rows, int maxRows)
{
	try
	{
		Sim& s = *reinterpret_cast<Sim*>(p);
		using namespace kb;
		const uint32_t off[2] = { 0, (uint32_t)len }, order[1] = { 0 };
		const size_t U = 2 * (size_t)len + 4, npu = KB_DEFAULT_NODES_PER_UNIT;
		BatchView bv{};
		bv.n_sent = 1; bv.text = text; bv.text_off = off; bv.match_options = matchOptions; bv.nodes_per_unit = (uint32_t)npu; bv.order = order;
		auto norm = buf<uint16_t>(U + 32); auto normLen = buf<uint32_t>(1); auto posTable = buf<uint32_t>(len + 2);
		auto nsToPos = buf<uint32_t>(U), posToNs = buf<uint32_t>(U), ctr = buf<uint32_t>(U); auto endPosMap = buf<uint2>(U);
		auto pats = buf<DPattern>(U); auto build = buf<DNode>(U * npu), nodes = buf<DNode>(U * npu); auto newIndex = buf<uint32_t>(U * npu);
		auto chunks = buf<DChunk>(U / 4 + 10); auto nChunks = buf<uint32_t>(1), status = buf<uint32_t>(1), debug = buf<uint32_t>(64);
		bv.norm = norm.data(); bv.norm_len = normLen.data(); bv.pos_table = posTable.data(); bv.ns_to_pos = nsToPos.data(); bv.pos_to_ns = posToNs.data();
		bv.end_pos_map = endPosMap.data(); bv.ctr = ctr.data(); bv.patterns = pats.data(); bv.build_nodes = build.data(); bv.nodes = nodes.data();
		bv.new_index = newIndex.data(); bv.chunks = chunks.data(); bv.n_chunks = nChunks.data(); bv.status = status.data(); bv.debug = debug.data();
		std::vector<DTypoNode> tgTmp, tg; std::vector<uint32_t> tgRemap; std::vector<uint2> tgRange; std::vector<DTypoState> tgStates; std::vector<DTypoMatch> tgMatches;
		if (!s.typo.empty())
		{
			const auto* th = reinterpret_cast<const kb2_typo_header*>(s.typo.data());
			size_t o = sizeof(kb2_typo_header);
			auto take = [&](size_t bytes) { o = (o + 15) / 16 * 16; const char* q = s.typo.data() + o; o += bytes; return q; };
			TypoView& tv = bv.typo;
			tv.nodes = reinterpret_cast<const kb2_typo_node*>(take(sizeof(kb2_typo_node) * th->n_nodes));
			tv.keys = reinterpret_cast<const uint16_t*>(take(2 * (size_t)th->n_edges));
			tv.diffs = reinterpret_cast<const int32_t*>(take(4 * (size_t)th->n_edges));
			tv.pats = reinterpret_cast<const kb2_typo_pat*>(take(sizeof(kb2_typo_pat) * th->n_pats));
			tv.repls = reinterpret_cast<const kb2_typo_repl*>(take(sizeof(kb2_typo_repl) * th->n_repls));
			tv.pool = reinterpret_cast<const uint16_t*>(take(2 * (size_t)th->n_pool));
			tv.threshold = s.typoThreshold; tv.continual_threshold = th->continual_typo_threshold;
			tv.graph_per_unit = s.graphPerUnit; tv.states_per_unit = s.statesPerUnit;
			tgTmp.resize(U * s.graphPerUnit); tg.resize(U * s.graphPerUnit); tgRemap.resize(U * s.graphPerUnit); tgRange.resize(U * s.graphPerUnit);
			tgMatches.resize(U * s.graphPerUnit); tgStates.resize(U * s.statesPerUnit);
			tv.tmp = tgTmp.data(); tv.graph = tg.data(); tv.remap = tgRemap.data(); tv.state_range = tgRange.data(); tv.matches = tgMatches.data(); tv.states = tgStates.data();
		}
		set_model_lattice(s.model.dev);
		launch_lattice(s.model.dev, bv, nullptr);
		if (std::getenv("HS_DUMP"))
		{
			std::fprintf(stderr, "[hs] status %u normLen %u nChunks %u norm:", status[0], normLen[0], nChunks[0]);
			for (uint32_t i = 0; i < normLen[0]; ++i) std::fprintf(stderr, " %04x", norm[i]);
			std::fprintf(stderr, "\n[hs] posTable:");
			for (int i = 0; i <= len; ++i) std::fprintf(stderr, " %u", posTable[i]);
			std::fprintf(stderr, "\n[hs] nsToPos:");
			for (uint32_t i = 0; i < normLen[0]; ++i) std::fprintf(stderr, " %u", nsToPos[i]);
			std::fprintf(stderr, "\n[hs] pats[0]: end %u len %u tag %u\n", pats[0].end, pats[0].len, pats[0].tag);
			for (uint32_t i = 0; i < 24; ++i) std::fprintf(stderr, "[hs] build %u: form %d u(%u,%u) pos [%u,%u) prev %u sib %u newIndex %d\n", i, build[i].form, build[i].uform_off, build[i].uform_len,
				build[i].start_pos, build[i].end_pos, build[i].prev, build[i].sibling, (int)newIndex[i]);
			for (uint32_t c = 0; c < nChunks[0]; ++c) std::fprintf(stderr, "[hs] chunk %u: [%u,%u) nodes %u\n", c, chunks[c].start, chunks[c].end, chunks[c].n_nodes);
		}
		if (status[0]) return -(int)status[0];
		int total = 0;
		for (uint32_t c = 0; c < nChunks[0]; ++c)
		{
			for (uint32_t i = 0; i < chunks[c].n_nodes; ++i)
			{
				const DNode& nd = nodes[chunks[c].node_off + i];
				if (total >= maxRows) return -100;
				int32_t* r = rows + 9 * (size_t)total++;
				r[0] = nd.form; r[1] = nd.uform_len ? (int32_t)nd.uform_off : -1; r[2] = (int32_t)nd.uform_len; r[3] = nd.prev; r[4] = nd.sibling;
				r[5] = (int32_t)nd.start_pos; r[6] = (int32_t)nd.end_pos; r[7] = nd.space_errors; r[8] = (int32_t)c;
			}
		}
		return total;
	}
	catch (...) { return -100; }
}

}
