// HOST SIMULATION of the whole device pipeline with 32-lane warps (test infrastructure, tests/ only; see shim32/cuda_runtime.h):
// lattice.cu, viterbi.cu (Knlm build), emit.cu and model.cu — the device SOURCES — compiled as C++, each kernel run by 32 OS
// threads that meet at a barrier for every warp collective.  One sentence per call, behind a small C API for
// tests/test_hostsim_pipeline.py, which compares tokens and scores with the reference's golden vectors without a GPU.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <mutex>
#include <string>
#include <vector>
#include "../../kiwi_b200/csrc/engine.h"

namespace kb
{
	cudaError_t launch_lattice(const DevModel& m, const BatchView& bv, cudaStream_t stream);
	cudaError_t launch_viterbi(const DevModel& m, const BatchView& bv, const VitView& vv, cudaStream_t stream);
	cudaError_t launch_emit(const DevModel& m, const BatchView& bv, const VitView& vv, cudaStream_t stream);
	cudaError_t set_model_lattice(const DevModel& m);
	cudaError_t set_model_viterbi(const DevModel& m);
	cudaError_t set_model_emit(const DevModel& m);
	cudaError_t launch_viterbi_cong(const DevModel& m, const BatchView& bv, const VitView& vv, cudaStream_t stream);
	cudaError_t set_model_viterbi_cong(const DevModel& m);
	cudaError_t launch_viterbi_sbg(const DevModel& m, const BatchView& bv, const VitView& vv, cudaStream_t stream);
	cudaError_t set_model_viterbi_sbg(const DevModel& m);
}

namespace
{
	struct Sim
	{
		kb::Model model;
		std::vector<char> typo;
		float typoThreshold = 2.5f;
		uint32_t graphPerUnit = 6, statesPerUnit = 6;
	};
	std::mutex g_mtx;      // the simulated warp and the constant-memory model views are process-wide
	template<class T> std::vector<T> buf(size_t n) { return std::vector<T>(n); }
}

static void hs32SegvHandler(int sig) { void* bt[64]; const int n = backtrace(bt, 64); backtrace_symbols_fd(bt, n, 2); _exit(139); }

extern "C" {

void* hs32_open(const char* imagePath)
{
	try
	{
		auto blob = kb::readImageFile(imagePath);
		auto* s = new Sim;
		s->model.load(blob.data(), blob.size());
		if (s->model.dev.model_type != 2 && s->model.dev.model_type != 3 && s->model.dev.model_type != 4) { delete s; return nullptr; }      // Knlm, SkipBigram and CoNg builds of viterbi.cu
		return s;
	}
	catch (...) { return nullptr; }
}
void hs32_close(void* p) { delete reinterpret_cast<Sim*>(p); }

int hs32_set_typo(void* p, const char* path, float threshold)
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
	s.typoThreshold = threshold;
	return 0;
}

// one sentence through lattice -> viterbi -> emit; tokens as (morph, tag, position, length, score); returns the token count,
// -status on a kernel status, -100 on an error.  nodes = number of lattice nodes (all chunks)
// AnalyzeOption::blocklist for the simulated kernels: the model's candidate table ("device" memory is host memory here) is rewritten in place
// with Model::blockedCands - the table the engine uploads for a call with a blocklist; n == 0 restores the model's own table
int hs32_set_blocklist(void* p, const uint32_t* ids, int n)
{
	try
	{
		std::lock_guard<std::mutex> lk(g_mtx);
		Sim& s = *reinterpret_cast<Sim*>(p);
		std::vector<uint32_t> v(ids, ids + n);
		std::sort(v.begin(), v.end());
		const std::vector<kb::DCand> rows = s.model.blockedCands(v);
		std::memcpy(const_cast<kb::DCand*>(s.model.dev.cands), rows.data(), rows.size() * sizeof(kb::DCand));
		return 0;
	}
	catch (...) { return -1; }
}

int hs32_analyze(void* p, const uint16_t* text, int len, uint32_t matchOptions, uint32_t* morph, uint8_t* tag, uint32_t* pos, uint16_t* length, float* score,
	int maxTokens, float* sentScore, int* nNodes, uint8_t* flags)
{
	try
	{
		std::lock_guard<std::mutex> lk(g_mtx);
		Sim& s = *reinterpret_cast<Sim*>(p);
		using namespace kb;
		const uint32_t off[2] = { 0, (uint32_t)len }, order[1] = { 0 };
		// (SkipBigram paths rarely merge: the arena of the engine's first retry round, 8 x, from the start)
		const size_t capMul = s.model.dev.model_type == 3 ? 8 : 1;
		const size_t U = 2 * (size_t)len + 4, npu = KB_DEFAULT_NODES_PER_UNIT, ppu = 128 * capMul, pc = 8192 * capMul;
		BatchView bv{}; VitView vv{};
		bv.n_sent = 1; bv.text = text; bv.text_off = off; bv.match_options = matchOptions; bv.nodes_per_unit = (uint32_t)npu; bv.order = order;
		auto norm = buf<uint16_t>(U + 32); auto normLen = buf<uint32_t>(1); auto posTable = buf<uint32_t>(len + 2);
		auto nsToPos = buf<uint32_t>(U), posToNs = buf<uint32_t>(U), ctr = buf<uint32_t>(U); auto endPosMap = buf<uint2>(U);
		auto pats = buf<DPattern>(U); auto build = buf<DNode>(U * npu), nodes = buf<DNode>(U * npu); auto newIndex = buf<uint32_t>(U * npu);
		const size_t chunkSlots = U / 4 + 10;
		auto chunks = buf<DChunk>(chunkSlots); auto nChunks = buf<uint32_t>(1), status = buf<uint32_t>(1), debug = buf<uint32_t>(64);
		bv.norm = norm.data(); bv.norm_len = normLen.data(); bv.pos_table = posTable.data(); bv.ns_to_pos = nsToPos.data(); bv.pos_to_ns = posToNs.data();
		bv.end_pos_map = endPosMap.data(); bv.ctr = ctr.data(); bv.patterns = pats.data(); bv.build_nodes = build.data(); bv.nodes = nodes.data();
		bv.new_index = newIndex.data(); bv.chunks = chunks.data(); bv.n_chunks = nChunks.data(); bv.status = status.data(); bv.debug = debug.data();
		const size_t pathStride = s.model.dev.model_type == 3 ? 96 : sizeof(DPath);      // (SkipBigram records carry the history, viterbi.cu PathS)
		auto paths = buf<DPath>((ppu * U + pc) * pathStride / sizeof(DPath)); auto npOff = buf<uint32_t>(U * npu), npCnt = buf<uint32_t>(U * npu); auto nodeCand = buf<uint2>(U * npu); auto reach = buf<uint8_t>(U * npu);
		auto recs = buf<DRec>(2 * chunkSlots); auto toks = buf<DToken>(U); auto nTok = buf<uint32_t>(2); auto bestRec = buf<int32_t>(1); auto sc = buf<float>(1);
		auto timing = buf<unsigned long long>(2);
		auto workCounter = buf<uint32_t>(4);
		vv.work_counter = workCounter.data(); vv.solo_blocks = 0; vv.solo_warps = 1;
		vv.n_team = std::getenv("HS32_TEAM") ? 1u : 0u;      // HS32_TEAM=1: the sentence is analysed by a team of warps (viterbi.cu team mode)
		vv.paths_per_unit = (uint32_t)ppu; vv.paths_const = (uint32_t)pc; vv.path_stride = (uint32_t)pathStride; vv.paths = paths.data(); vv.node_path_off = npOff.data(); vv.node_path_cnt = npCnt.data(); vv.node_cand = nodeCand.data();
		vv.reachable = reach.data(); vv.recs = recs.data(); vv.tokens = toks.data(); vv.n_tokens = nTok.data(); vv.best_rec = bestRec.data(); vv.score = sc.data(); vv.timing = timing.data();
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
		const bool cong = s.model.dev.model_type == 4;
		set_model_lattice(s.model.dev); set_model_emit(s.model.dev);
		const bool sbgModel = s.model.dev.model_type == 3;
		if (cong) set_model_viterbi_cong(s.model.dev); else if (sbgModel) set_model_viterbi_sbg(s.model.dev); else set_model_viterbi(s.model.dev);
		const bool trace = std::getenv("HS32_TRACE") != nullptr;
		if (std::getenv("HS32_BT")) signal(SIGSEGV, hs32SegvHandler);      // debugging aid: symbolised backtrace of a faulting lane
		if (trace) std::fprintf(stderr, "[hs32] lattice\n");
		if (launch_lattice(s.model.dev, bv, nullptr)) return -100;
		if (trace) std::fprintf(stderr, "[hs32] lattice done: status %u chunks %u\n", status[0], nChunks[0]);
		if (status[0]) return -(int)status[0];
		*nNodes = 0;
		for (uint32_t c = 0; c < nChunks[0]; ++c) *nNodes += (int)chunks[c].n_nodes;
		if (cong ? launch_viterbi_cong(s.model.dev, bv, vv, nullptr) : sbgModel ? launch_viterbi_sbg(s.model.dev, bv, vv, nullptr) : launch_viterbi(s.model.dev, bv, vv, nullptr)) return -100;
		if (const char* dn = std::getenv("HS32_DUMP_NODE"))
		{
			const int i = std::atoi(dn);
			for (uint32_t k = 0; k < npCnt[i]; ++k)
			{
				const char* rec = reinterpret_cast<const char*>(paths.data()) + (size_t)(npOff[i] + k) * pathStride;
				const DPath& q = *reinterpret_cast<const DPath*>(rec);
				std::fprintf(stderr, "[path] morph %d lm %d acc %a root %u sp %u", q.morpheme, q.lm_state, q.acc_score, q.root_id, q.sp_state);
				if (pathStride == 96) { const uint32_t* hx = reinterpret_cast<const uint32_t*>(rec + 48); std::fprintf(stderr, " hist %u %u %u %u %u %u %u %u pos %u", hx[0], hx[1], hx[2], hx[3], hx[4], hx[5], hx[6], hx[7], hx[8]); }
				std::fprintf(stderr, "\n");
			}
		}
		if (trace) { std::fprintf(stderr, "[hs32] paths per node:"); for (int i = 0; i < *nNodes; ++i) std::fprintf(stderr, " %u", npCnt[i]); std::fprintf(stderr, "\n"); }
		if (trace && bestRec[0] >= 0)
		{
			const DRec r = recs[bestRec[0]];
			std::fprintf(stderr, "[hs32] best rec %d: parent_rec %d end_parent %u chunk %u score %f\n", bestRec[0], r.parent_rec, r.end_parent, r.chunk, r.score);
			uint32_t pi = r.end_parent;
			for (int k = 0; k < 64 && pi != 0xFFFFFFFFu; ++k)
			{
				const DPath& q = paths[pi];
				std::fprintf(stderr, "[hs32]   path %u: node %u morph %d acc %f parent %u wid %u lm %d root %u sp %u\n", pi, q.node, q.morpheme, q.acc_score, q.parent, q.wid, q.lm_state, q.root_id, q.sp_state);
				pi = q.parent;
			}
			for (int i = 1; i < *nNodes; ++i)
			{
				std::fprintf(stderr, "[hs32] node %d: paths [%u, +%u) holes:", i, npOff[i], npCnt[i]);
				for (uint32_t k = 0; k < npCnt[i]; ++k) { const DPath& q = paths[npOff[i] + k]; if (q.node != (uint32_t)i) std::fprintf(stderr, " %u(node %u)", npOff[i] + k, q.node); }
				std::fprintf(stderr, "\n");
			}
		}
		if (trace) std::fprintf(stderr, "[hs32] viterbi done: status %u best_rec %d score %f paths@node1 %u\n", status[0], bestRec[0], sc[0], npCnt[1]);
		if (status[0]) return -(int)status[0];
		if (s.model.dev.debug[0] || debug[0]) return -101;
		if (launch_emit(s.model.dev, bv, vv, nullptr)) return -100;
		if (status[0]) return -(int)status[0];
		const int n = (int)nTok[0];
		if (n > maxTokens) return -100;
		for (int i = 0; i < n; ++i)
		{
			morph[i] = toks[i].morph; tag[i] = toks[i].tag; pos[i] = toks[i].position; length[i] = toks[i].length; score[i] = toks[i].score;
			if (flags) flags[i] = toks[i].flags;
		}
		*sentScore = sc[0];
		return n;
	}
	catch (...) { return -100; }
}

}
