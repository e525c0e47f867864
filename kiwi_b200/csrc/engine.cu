// kiwi_b200: host engine.  Owns the device-resident model, the per-batch scratch arena and the stream; turns
// one batch of raw UTF-16 sentences into flat token arrays with exactly four compute launches
// (lattice_kernel, viterbi_kernel, emit_kernel, pack_kernel) plus one cub scan.
//
// Host role (the reference does all of this per sentence on CPU threads, src/Kiwi.cpp:1014-1158 and
// include/kiwi/Kiwi.h:402-454): here the host only copies the text blob + offsets in and the packed
// tokens out; normalisation, chunking, lattice, Viterbi, stitching and position mapping run on the GPU.
// Sentences whose scratch demand exceeds the arithmetic capacity (rare, e.g. 50 x the same syllable) are
// re-run in a second, larger-capacity pass; a sentence that still does not fit is a hard error.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include "engine.h"

namespace kb
{
	cudaError_t launch_lattice(const DevModel& m, const BatchView& bv, cudaStream_t stream);
	cudaError_t launch_viterbi(const DevModel& m, const BatchView& bv, const VitView& vv, cudaStream_t stream);
	cudaError_t launch_emit(const DevModel& m, const BatchView& bv, const VitView& vv, cudaStream_t stream);
	cudaError_t set_model_lattice(const DevModel& m);
	cudaError_t set_model_viterbi(const DevModel& m);
	cudaError_t launch_viterbi_cong(const DevModel& m, const BatchView& bv, const VitView& vv, cudaStream_t stream);
	cudaError_t set_model_viterbi_cong(const DevModel& m);
	cudaError_t launch_cong_debug(uint32_t n, const uint32_t* ctx, const uint32_t* wid, const int32_t* node,
		int32_t* outDot, float* outEps, int32_t* outNode, uint32_t* outCtx, int32_t* outTile, cudaStream_t stream);
	cudaError_t set_model_emit(const DevModel& m);

	static void ck(cudaError_t e, const char* what)
	{
		if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
	}

	static constexpr uint32_t DEFAULT_PATHS_PER_UNIT = 128, DEFAULT_PATHS_CONST = 8192;
	// typo graph nodes / search states per normalised-unit slot (W_s = 2 n + 4 slots per sentence): the basic typo set needs < 4
	// on the reference's evaluation texts (tests/test_hostsim_lattice.py); overflow -> ST_TYPO_OVERFLOW -> retry arena
	static constexpr uint32_t DEFAULT_TYPO_GRAPH_PER_UNIT = 6, DEFAULT_TYPO_STATES_PER_UNIT = 6;

	__global__ void length_kernel(uint32_t nSent, const uint32_t* __restrict__ textOff, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx)
	{
		const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
		if (s >= nSent) return;
		keys[s] = textOff[s + 1] - textOff[s];
		idx[s] = s;
	}

	__global__ void pack_kernel(uint32_t nSent, const uint32_t* __restrict__ textOff, const uint32_t* __restrict__ nTokens,
		const uint32_t* __restrict__ tokOff, const DToken* __restrict__ tokens, DToken* __restrict__ packed)
	{
		const uint32_t lane = threadIdx.x & 31;
		const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
		if (s >= nSent) return;
		const size_t wbase = 2 * (size_t)textOff[s] + 4 * (size_t)s;
		const uint32_t n = nTokens[s], o = tokOff[s];
		for (uint32_t i = lane; i < n; i += 32) packed[o + i] = tokens[wbase + i];
	}

	Engine::Engine(const void* imageBytes, size_t size)
	{
		model.load(imageBytes, size);
		ck(set_model_lattice(model.dev), "constant upload"); ck(model.dev.model_type == 4 ? set_model_viterbi_cong(model.dev) : set_model_viterbi(model.dev), "constant upload"); ck(set_model_emit(model.dev), "constant upload");
		ck(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
		for (auto& e : ev) ck(cudaEventCreate(&e), "cudaEventCreate");
	}

	Engine::~Engine()
	{
		freeScratch(main_); freeScratch(retry_);
		if (hPinText) cudaFreeHost(hPinText);
		if (hPinOff) cudaFreeHost(hPinOff);
		if (hPinOut) cudaFreeHost(hPinOut);
		for (auto& e : ev) cudaEventDestroy(e);
		if (stream) cudaStreamDestroy(stream);
	}

	void Engine::freeScratch(Scratch& sc)
	{
		for (void* p : sc.bufs) cudaFree(p);
		for (void* p : sc.typoBufs) cudaFree(p);
		sc = Scratch{};
	}

	// ---- typo lattice: the device-resident transformer and the per-sentence graph / state scratch ----------------
	void TypoDev::load(const void* bytes, size_t size)
	{
		if (size < sizeof(kb2_typo_header)) throw std::runtime_error("typo image: too small");
		blob.assign(reinterpret_cast<const char*>(bytes), reinterpret_cast<const char*>(bytes) + size);
		const auto* h = reinterpret_cast<const kb2_typo_header*>(blob.data());
		if (h->magic != KB2_TYPO_MAGIC) throw std::runtime_error("typo image: bad magic");
		size_t o = sizeof(kb2_typo_header);
		auto take = [&](size_t bytesOf) { o = (o + 15) / 16 * 16; const size_t at = o; o += bytesOf; return at; };
		const size_t oNodes = take(sizeof(kb2_typo_node) * h->n_nodes), oKeys = take(2 * (size_t)h->n_edges), oDiffs = take(4 * (size_t)h->n_edges);
		const size_t oPats = take(sizeof(kb2_typo_pat) * h->n_pats), oRepls = take(sizeof(kb2_typo_repl) * h->n_repls), oPool = take(2 * (size_t)h->n_pool);
		if (o > size) throw std::runtime_error("typo image: truncated");
		if (std::isfinite(h->lengthening_typo_threshold)) throw std::runtime_error("typo image: lengthening typo sets are outside the kiwi_b200 hot path");
		ck(cudaMalloc(&dBlob, size), "cudaMalloc(typo image)");
		ck(cudaMemcpy(dBlob, blob.data(), size, cudaMemcpyHostToDevice), "cudaMemcpy(typo image)");
		const char* d = reinterpret_cast<const char*>(dBlob);
		view = TypoView{};
		view.nodes = reinterpret_cast<const kb2_typo_node*>(d + oNodes); view.keys = reinterpret_cast<const uint16_t*>(d + oKeys);
		view.diffs = reinterpret_cast<const int32_t*>(d + oDiffs); view.pats = reinterpret_cast<const kb2_typo_pat*>(d + oPats);
		view.repls = reinterpret_cast<const kb2_typo_repl*>(d + oRepls); view.pool = reinterpret_cast<const uint16_t*>(d + oPool);
		view.continual_threshold = h->continual_typo_threshold;
	}

	TypoDev::~TypoDev() { if (dBlob) cudaFree(dBlob); }

	void Engine::ensureTypoScratch(Scratch& sc, uint32_t gpu, uint32_t spu)
	{
		if (sc.typoCapUnits >= sc.capUnits && sc.typoGraphPerUnit == gpu && sc.typoStatesPerUnit == spu) return;
		for (void* p : sc.typoBufs) cudaFree(p);
		sc.typoBufs.clear();
		auto alloc = [&](size_t bytes) { void* p = nullptr; ck(cudaMalloc(&p, std::max<size_t>(bytes, 256)), "cudaMalloc(typo scratch)"); sc.typoBufs.push_back(p); return p; };
		const size_t G = sc.capUnits * gpu, S = sc.capUnits * spu;
		TypoView& t = sc.typoScratch;
		t = TypoView{};
		t.graph_per_unit = gpu; t.states_per_unit = spu;
		t.tmp = (DTypoNode*)alloc(G * sizeof(DTypoNode)); t.graph = (DTypoNode*)alloc(G * sizeof(DTypoNode));
		t.remap = (uint32_t*)alloc(G * 4); t.state_range = (uint2*)alloc(G * 8); t.matches = (DTypoMatch*)alloc(G * sizeof(DTypoMatch));
		t.states = (DTypoState*)alloc(S * sizeof(DTypoState));
		sc.typoCapUnits = sc.capUnits; sc.typoGraphPerUnit = gpu; sc.typoStatesPerUnit = spu;
	}

	void Engine::ensureScratch(Scratch& sc, size_t U, size_t B, uint32_t ppu, uint32_t pc, uint32_t npu)
	{
		const size_t T = (U - 4 * B) / 2 + 1;
		if (U <= sc.capUnits && B <= sc.capSent && ppu == sc.pathsPerUnit && pc == sc.pathsConst && npu == sc.nodesPerUnit && T <= sc.capText) return;
		const size_t capU = std::max(U, sc.capUnits), capB = std::max(B, sc.capSent), capT = std::max(T, sc.capText);
		freeScratch(sc);
		auto alloc = [&](size_t bytes) { void* p = nullptr; ck(cudaMalloc(&p, std::max<size_t>(bytes, 256)), "cudaMalloc(scratch)"); sc.bufs.push_back(p); return p; };
		BatchView& bv = sc.bv; VitView& vv = sc.vv;
		bv.nodes_per_unit = npu;
		bv.norm = (uint16_t*)alloc(capU * 2 + 64);
		bv.norm_len = (uint32_t*)alloc(capB * 4);
		bv.pos_table = (uint32_t*)alloc((capT + capB + 1) * 4);
		bv.ns_to_pos = (uint32_t*)alloc(capU * 4);
		bv.pos_to_ns = (uint32_t*)alloc(capU * 4);
		bv.end_pos_map = (uint2*)alloc(capU * 8);
		bv.ctr = (uint32_t*)alloc(capU * 4);
		bv.patterns = (DPattern*)alloc(capU * sizeof(DPattern));
		bv.build_nodes = (DNode*)alloc(capU * npu * sizeof(DNode));
		bv.nodes = (DNode*)alloc(capU * npu * sizeof(DNode));
		bv.new_index = (uint32_t*)alloc(capU * npu * 4);
		const size_t chunkSlots = capU / 4 + 2 * capB + 8;
		bv.chunks = (DChunk*)alloc(chunkSlots * sizeof(DChunk));
		bv.n_chunks = (uint32_t*)alloc(capB * 4);
		bv.status = (uint32_t*)alloc(capB * 4);
		bv.debug = (uint32_t*)alloc(64 * 4);
		ck(cudaMemset(bv.debug, 0, 64 * 4), "memset");
		vv.paths_per_unit = ppu; vv.paths_const = pc;
		vv.paths = (DPath*)alloc(((size_t)ppu * capU + (size_t)pc * capB) * sizeof(DPath));
		vv.node_path_off = (uint32_t*)alloc(capU * npu * 4);
		vv.node_path_cnt = (uint32_t*)alloc(capU * npu * 4);
		vv.reachable = (uint8_t*)alloc(capU * npu);
		vv.recs = (DRec*)alloc(2 * chunkSlots * sizeof(DRec));
		vv.tokens = (DToken*)alloc(capU * sizeof(DToken));
		vv.n_tokens = (uint32_t*)alloc((capB + 1) * 4);
		vv.best_rec = (int32_t*)alloc(capB * 4);
		vv.score = (float*)alloc(capB * 4);
		vv.timing = (unsigned long long*)alloc(capB * 16);
		sc.tokOff = (uint32_t*)alloc((capB + 1) * 4);
		sc.packed = (DToken*)alloc(capU * sizeof(DToken));
		sc.dText = (uint16_t*)alloc(capT * 2 + 64);
		sc.dOff = (uint32_t*)alloc((capB + 1) * 4);
		size_t tb = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, tb, vv.n_tokens, sc.tokOff, (int)(capB + 1), stream);
		sc.cubTempBytes = tb; sc.cubTemp = alloc(tb);
		sc.lenKeys = (uint32_t*)alloc(capB * 4); sc.lenKeysOut = (uint32_t*)alloc(capB * 4); sc.idxIn = (uint32_t*)alloc(capB * 4); sc.order = (uint32_t*)alloc(capB * 4);
		size_t sb = 0;
		cub::DeviceRadixSort::SortPairsDescending(nullptr, sb, sc.lenKeys, sc.lenKeysOut, sc.idxIn, sc.order, (int)capB, 0, 32, stream);
		sc.sortTempBytes = sb; sc.sortTemp = alloc(sb);
		sc.capUnits = capU; sc.capSent = capB; sc.capText = capT; sc.pathsPerUnit = ppu; sc.pathsConst = pc; sc.nodesPerUnit = npu;
	}

	void Engine::bind(Scratch& sc, const uint16_t* dText, const uint32_t* dOffsets, uint32_t n, uint32_t matchOptions)
	{
		sc.bv.n_sent = n; sc.bv.text = dText; sc.bv.text_off = dOffsets; sc.bv.match_options = matchOptions;
		sc.bv.typo = TypoView{};
		if (typo_)
		{
			// the retry arena carries 4 x the graph / state capacity, like its node capacity
			const uint32_t mul = &sc == &retry_ ? 4 : 1;
			ensureTypoScratch(sc, DEFAULT_TYPO_GRAPH_PER_UNIT * mul, DEFAULT_TYPO_STATES_PER_UNIT * mul);
			TypoView v = sc.typoScratch;
			v.nodes = typo_->view.nodes; v.keys = typo_->view.keys; v.diffs = typo_->view.diffs; v.pats = typo_->view.pats; v.repls = typo_->view.repls; v.pool = typo_->view.pool;
			v.continual_threshold = typo_->view.continual_threshold; v.threshold = typoThreshold_;
			sc.bv.typo = v;
		}
	}

	static const Model* g_constantsOwner = nullptr;     // the constant-memory model view belongs to one engine at a time

	void Engine::launchAll(Scratch& sc, uint32_t n)
	{
		if (g_constantsOwner != &model)
		{
			ck(set_model_lattice(model.dev), "constant upload"); ck(model.dev.model_type == 4 ? set_model_viterbi_cong(model.dev) : set_model_viterbi(model.dev), "constant upload"); ck(set_model_emit(model.dev), "constant upload");
			g_constantsOwner = &model;
		}
		ck(cudaEventRecord(ev[1], stream), "event");
		// longest-processing-time-first launch order (sentence cost grows with its length)
		length_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, sc.bv.text_off, sc.lenKeys, sc.idxIn);
		ck(cudaGetLastError(), "length_kernel launch");
		size_t sb = sc.sortTempBytes;
		ck(cub::DeviceRadixSort::SortPairsDescending(sc.sortTemp, sb, sc.lenKeys, sc.lenKeysOut, sc.idxIn, sc.order, (int)n, 0, 32, stream), "cub sort");
		sc.bv.order = sc.order;
		ck(launch_lattice(model.dev, sc.bv, stream), "lattice_kernel launch");
		ck(cudaEventRecord(ev[2], stream), "event");
		ck(model.dev.model_type == 4 ? launch_viterbi_cong(model.dev, sc.bv, sc.vv, stream) : launch_viterbi(model.dev, sc.bv, sc.vv, stream), "viterbi_kernel launch");
		ck(cudaEventRecord(ev[3], stream), "event");
		ck(launch_emit(model.dev, sc.bv, sc.vv, stream), "emit_kernel launch");
		ck(cudaMemsetAsync(sc.vv.n_tokens + n, 0, 4, stream), "memset");
		size_t tb = sc.cubTempBytes;
		ck(cub::DeviceScan::ExclusiveSum(sc.cubTemp, tb, sc.vv.n_tokens, sc.tokOff, (int)(n + 1), stream), "cub scan");
		const uint32_t threads = 256, blocks = (n * 32 + threads - 1) / threads;
		pack_kernel<<<blocks, threads, 0, stream>>>(n, sc.bv.text_off, sc.vv.n_tokens, sc.tokOff, sc.vv.tokens, sc.packed);
		ck(cudaGetLastError(), "pack_kernel launch");
		ck(cudaEventRecord(ev[4], stream), "event");
	}

	static void growPinned(void** p, size_t* cap, size_t bytes)
	{
		if (bytes <= *cap) return;
		if (*p) cudaFreeHost(*p);
		const size_t nb = std::max(bytes, *cap * 2);
		ck(cudaMallocHost(p, nb), "cudaMallocHost");
		*cap = nb;
	}

	// one pass: H2D text + offsets, the four kernels, D2H of offsets / scores / status and of exactly the packed tokens
	void Engine::runHostPass(Scratch& sc, const uint16_t* ptext, const uint32_t* poff, uint32_t pn, uint32_t matchOptions, uint32_t ppu, uint32_t pc, uint32_t npu, PassResult& r, BatchOutput& out)
	{
		const size_t pT = poff[pn];
		const size_t U = 2 * pT + 4 * (size_t)pn;
		ensureScratch(sc, U, pn, ppu, pc, npu);
		growPinned((void**)&hPinText, &pinTextCap, pT * 2 + 64);
		growPinned((void**)&hPinOff, &pinOffCap, ((size_t)pn + 1) * 4);
		std::memcpy(hPinText, ptext, pT * 2);
		std::memcpy(hPinOff, poff, ((size_t)pn + 1) * 4);
		ck(cudaEventRecord(ev[0], stream), "event");
		ck(cudaMemcpyAsync(sc.dText, hPinText, pT * 2, cudaMemcpyHostToDevice, stream), "H2D text");
		ck(cudaMemcpyAsync(sc.dOff, hPinOff, ((size_t)pn + 1) * 4, cudaMemcpyHostToDevice, stream), "H2D offsets");
		bind(sc, sc.dText, sc.dOff, pn, matchOptions);
		launchAll(sc, pn);
		const size_t headBytes = ((size_t)pn + 1) * 4 + (size_t)pn * 4 * 2;
		growPinned(&hPinOut, &pinOutCap, std::max(headBytes, (size_t)U * sizeof(DToken) / 4));
		uint32_t* hTokOff = (uint32_t*)hPinOut; float* hScore = (float*)(hTokOff + pn + 1); uint32_t* hStatus = (uint32_t*)(hScore + pn);
		ck(cudaMemcpyAsync(hTokOff, sc.tokOff, ((size_t)pn + 1) * 4, cudaMemcpyDeviceToHost, stream), "D2H offsets");
		ck(cudaMemcpyAsync(hScore, sc.vv.score, (size_t)pn * 4, cudaMemcpyDeviceToHost, stream), "D2H scores");
		ck(cudaMemcpyAsync(hStatus, sc.bv.status, (size_t)pn * 4, cudaMemcpyDeviceToHost, stream), "D2H status");
		ck(cudaStreamSynchronize(stream), "sync (a kernel fault surfaces here)");
		{
			uint32_t dbg[16];
			ck(cudaMemcpy(dbg, sc.bv.debug, sizeof(dbg), cudaMemcpyDeviceToHost), "D2H debug");
			if (!dbg[0]) ck(cudaMemcpy(dbg, model.dev.debug, sizeof(dbg), cudaMemcpyDeviceToHost), "D2H debug");
			if (dbg[0])
			{
				std::string msg = "internal consistency failure in viterbi_kernel:";
				for (int i = 1; i < 16; ++i) msg += " " + std::to_string(dbg[i]);
				throw std::runtime_error(msg);
			}
		}
		const uint32_t total = hTokOff[pn];
		r.tokOff.assign(hTokOff, hTokOff + pn + 1);
		r.scores.assign(hScore, hScore + pn);
		r.status.assign(hStatus, hStatus + pn);
		r.toks.resize(total);
		if (total)
		{
			growPinned(&hPinOut, &pinOutCap, (size_t)total * sizeof(DToken));
			ck(cudaMemcpyAsync(hPinOut, sc.packed, (size_t)total * sizeof(DToken), cudaMemcpyDeviceToHost, stream), "D2H tokens");
		}
		ck(cudaEventRecord(ev[5], stream), "event");
		ck(cudaStreamSynchronize(stream), "sync");
		if (total) std::memcpy(r.toks.data(), hPinOut, (size_t)total * sizeof(DToken));
		float ms;
		cudaEventElapsedTime(&ms, ev[0], ev[1]); out.msH2D += ms;
		cudaEventElapsedTime(&ms, ev[1], ev[2]); out.msLattice += ms;
		cudaEventElapsedTime(&ms, ev[2], ev[3]); out.msViterbi += ms;
		cudaEventElapsedTime(&ms, ev[3], ev[4]); out.msPack += ms;
		cudaEventElapsedTime(&ms, ev[4], ev[5]); out.msD2H += ms;
		cudaEventElapsedTime(&ms, ev[0], ev[5]); out.msTotal += ms;
		last.h2dBytes += pT * 2 + ((size_t)pn + 1) * 4;
		last.d2hBytes += headBytes + (size_t)total * sizeof(DToken);
		last.kernelLaunches += 5;
	}

	void Engine::analyzeOne(const uint16_t* text, const uint32_t* offsets, uint32_t n, uint32_t matchOptions, BatchOutput& out)
	{
		out = BatchOutput{};
		out.tokOff.assign(n + 1, 0);
		out.scores.assign(n, 0.f);
		out.status.assign(n, 0);
		if (n == 0) return;
		const size_t T = offsets[n];
		if (T >= (1ull << 31)) throw std::runtime_error("batch too large (>= 2^31 UTF-16 units); split it");

		PassResult r0;
		runHostPass(main_, text, offsets, n, matchOptions, DEFAULT_PATHS_PER_UNIT, DEFAULT_PATHS_CONST, KB_DEFAULT_NODES_PER_UNIT, r0, out);
		std::vector<uint32_t> failed;
		for (uint32_t i = 0; i < n; ++i) if (r0.status[i]) failed.push_back(i);
		last.retried += failed.size();
		if (failed.empty())
		{
			out.tokens = std::move(r0.toks); out.tokOff = std::move(r0.tokOff); out.scores = std::move(r0.scores); out.status = std::move(r0.status);
		}
		else
		{
			// second pass for the overflowed sentences only: 16 x path capacity, 4 x node capacity, in its own arena
			std::vector<uint16_t> subText; std::vector<uint32_t> subOff{ 0 };
			for (uint32_t id : failed)
			{
				subText.insert(subText.end(), text + offsets[id], text + offsets[id + 1]);
				subOff.push_back((uint32_t)subText.size());
			}
			PassResult r1;
			runHostPass(retry_, subText.data(), subOff.data(), (uint32_t)failed.size(), matchOptions, DEFAULT_PATHS_PER_UNIT * 8, DEFAULT_PATHS_CONST * 8, KB_DEFAULT_NODES_PER_UNIT * 4, r1, out);
			out.tokens.reserve(r0.toks.size() + r1.toks.size());
			out.tokOff.assign(n + 1, 0);
			out.scores = std::move(r0.scores); out.status = std::move(r0.status);
			size_t fi = 0;
			for (uint32_t i = 0; i < n; ++i)
			{
				out.tokOff[i] = (uint32_t)out.tokens.size();
				if (fi < failed.size() && failed[fi] == i)
				{
					out.tokens.insert(out.tokens.end(), r1.toks.begin() + r1.tokOff[fi], r1.toks.begin() + r1.tokOff[fi + 1]);
					out.scores[i] = r1.scores[fi]; out.status[i] = r1.status[fi];
					++fi;
				}
				else out.tokens.insert(out.tokens.end(), r0.toks.begin() + r0.tokOff[i], r0.toks.begin() + r0.tokOff[i + 1]);
			}
			out.tokOff[n] = (uint32_t)out.tokens.size();
		}
		for (uint32_t i = 0; i < n; ++i)
		{
			if (out.status[i])
			{
				throw std::runtime_error("sentence " + std::to_string(i) + " exceeded the device scratch capacity (status " + std::to_string(out.status[i]) + ")");
			}
		}
		last.nSentences = n; last.rawUnits = T; last.tokens = out.tokens.size();
		last.msLattice = out.msLattice; last.msViterbi = out.msViterbi; last.msPack = out.msPack;
	}

	// Public entry: batches of any size.  The scratch arena is sized per pass, so a large batch (65536 / 1 M sentences)
	// is cut into passes of at most MAX_UNITS_PER_PASS normalised units / MAX_SENT_PER_PASS sentences, run back to back.
	static constexpr size_t MAX_UNITS_PER_PASS = 2u << 20, MAX_SENT_PER_PASS = 16384;

	void Engine::analyze(const uint16_t* text, const uint32_t* offsets, uint32_t n, uint32_t matchOptions, BatchOutput& out)
	{
		last = Stats{};
		if (n && offsets[0] != 0) throw std::runtime_error("offsets[0] must be 0");
		for (uint32_t i = 0; i < n; ++i) if (offsets[i + 1] < offsets[i]) throw std::runtime_error("offsets must be non-decreasing");
		const size_t totalUnits = n ? 2 * (size_t)offsets[n] + 4 * (size_t)n : 0;
		if (totalUnits <= MAX_UNITS_PER_PASS && n <= MAX_SENT_PER_PASS) { analyzeOne(text, offsets, n, matchOptions, out); return; }
		out = BatchOutput{};
		out.tokOff.assign(1, 0);
		uint32_t i0 = 0;
		std::vector<uint32_t> subOff;
		while (i0 < n)
		{
			uint32_t i1 = i0; size_t units = 0;
			while (i1 < n && i1 - i0 < MAX_SENT_PER_PASS)
			{
				const size_t u = 2 * (size_t)(offsets[i1 + 1] - offsets[i1]) + 4;
				if (i1 > i0 && units + u > MAX_UNITS_PER_PASS) break;
				units += u; ++i1;
			}
			subOff.resize(i1 - i0 + 1);
			for (uint32_t k = 0; k <= i1 - i0; ++k) subOff[k] = offsets[i0 + k] - offsets[i0];
			BatchOutput part;
			analyzeOne(text + offsets[i0], subOff.data(), i1 - i0, matchOptions, part);
			const uint32_t base = (uint32_t)out.tokens.size();
			out.tokens.insert(out.tokens.end(), part.tokens.begin(), part.tokens.end());
			for (uint32_t k = 1; k <= i1 - i0; ++k) out.tokOff.push_back(base + part.tokOff[k]);
			out.scores.insert(out.scores.end(), part.scores.begin(), part.scores.end());
			out.status.insert(out.status.end(), part.status.begin(), part.status.end());
			out.msH2D += part.msH2D; out.msLattice += part.msLattice; out.msViterbi += part.msViterbi; out.msPack += part.msPack; out.msD2H += part.msD2H; out.msTotal += part.msTotal;
			i0 = i1;
		}
		last.nSentences = n; last.rawUnits = offsets[n]; last.tokens = out.tokens.size();
		last.msLattice = out.msLattice; last.msViterbi = out.msViterbi; last.msPack = out.msPack;
	}

	float Engine::analyzeDevice(const uint16_t* dText, const uint32_t* dOffsets, uint32_t n, uint64_t totalUnits, uint32_t matchOptions, uint64_t* nTokens)
	{
		const size_t U = 2 * (size_t)totalUnits + 4 * (size_t)n;
		if (U > 4 * MAX_UNITS_PER_PASS) throw std::runtime_error("kiwi_b200_analyze_device: batch too large for one device pass; split it (kiwi_b200_analyze_batch splits automatically)");
		Scratch& sc = main_;
		ensureScratch(sc, U, n, DEFAULT_PATHS_PER_UNIT, DEFAULT_PATHS_CONST, KB_DEFAULT_NODES_PER_UNIT);
		bind(sc, dText, dOffsets, n, matchOptions);
		launchAll(sc, n);
		growPinned(&hPinOut, &pinOutCap, (size_t)n * 4 + 64);
		uint32_t* hStatus = (uint32_t*)hPinOut;
		uint32_t total = 0;
		ck(cudaMemcpyAsync(&total, sc.tokOff + n, 4, cudaMemcpyDeviceToHost, stream), "D2H total");
		ck(cudaMemcpyAsync(hStatus, sc.bv.status, (size_t)n * 4, cudaMemcpyDeviceToHost, stream), "D2H status");
		ck(cudaStreamSynchronize(stream), "sync (a kernel fault surfaces here)");
		float ms = 0, a = 0;
		cudaEventElapsedTime(&ms, ev[1], ev[4]);
		cudaEventElapsedTime(&a, ev[1], ev[2]); last.msLattice = a;
		cudaEventElapsedTime(&a, ev[2], ev[3]); last.msViterbi = a;
		cudaEventElapsedTime(&a, ev[3], ev[4]); last.msPack = a;
		last.nSentences = n; last.rawUnits = totalUnits; last.tokens = total; last.kernelLaunches = 5; last.h2dBytes = 0; last.d2hBytes = 4 + (size_t)n * 4; last.retried = 0;
		// overflowed sentences (rare) are re-run through the larger arena; their text comes back from the device
		std::vector<uint32_t> failed;
		for (uint32_t i = 0; i < n; ++i) if (hStatus[i]) failed.push_back(i);
		if (!failed.empty())
		{
			std::vector<uint32_t> off(n + 1);
			ck(cudaMemcpy(off.data(), dOffsets, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost), "D2H offsets");
			std::vector<uint16_t> subText; std::vector<uint32_t> subOff{ 0 };
			for (uint32_t id : failed)
			{
				const size_t len = off[id + 1] - off[id], at = subText.size();
				subText.resize(at + len);
				if (len) ck(cudaMemcpy(subText.data() + at, dText + off[id], len * 2, cudaMemcpyDeviceToHost), "D2H text");
				subOff.push_back((uint32_t)subText.size());
			}
			BatchOutput tmp; PassResult r1;
			runHostPass(retry_, subText.data(), subOff.data(), (uint32_t)failed.size(), matchOptions, DEFAULT_PATHS_PER_UNIT * 8, DEFAULT_PATHS_CONST * 8, KB_DEFAULT_NODES_PER_UNIT * 4, r1, tmp);
			for (uint32_t s : r1.status) if (s) throw std::runtime_error("a sentence exceeded the device scratch capacity even in the retry arena (status " + std::to_string(s) + ")");
			ms += tmp.msTotal;
			total += (uint32_t)r1.toks.size();
			last.retried = failed.size(); last.tokens = total;
		}
		if (nTokens) *nTokens = total;
		return ms;
	}

	void Engine::debugCong(uint32_t n, const uint32_t* ctx, const uint32_t* wid, const int32_t* node,
		int32_t* outDot, float* outEps, int32_t* outNode, uint32_t* outCtx, int32_t* outTile)
	{
		if (model.dev.model_type != 4) throw std::runtime_error("debugCong needs a CoNg model image");
		if (n == 0) return;
		for (uint32_t i = 0; i < n; ++i)
		{
			if (ctx[i] >= model.dev.cg_context_size || wid[i] >= model.dev.lang_vocab_size || node[i] < 0 || (uint32_t)node[i] >= model.header.cg_num_nodes)
				throw std::runtime_error("debugCong: index out of range");
		}
		if (g_constantsOwner != &model) { ck(set_model_lattice(model.dev), "constant upload"); ck(set_model_viterbi_cong(model.dev), "constant upload"); ck(set_model_emit(model.dev), "constant upload"); g_constantsOwner = &model; }
		const uint32_t nU = std::min(n, 64u), nW = std::min(n, 32u);
		uint32_t* d = nullptr;
		const size_t words = (size_t)n * 9 + (size_t)nU * nW;
		ck(cudaMalloc(&d, words * 4), "cudaMalloc(debugCong)");
		uint32_t* dCtx = d; uint32_t* dWid = d + n; int32_t* dNode = reinterpret_cast<int32_t*>(d + 2 * n);
		int32_t* dDot = reinterpret_cast<int32_t*>(d + 3 * n); float* dEps = reinterpret_cast<float*>(d + 4 * n);
		int32_t* dNodeOut = reinterpret_cast<int32_t*>(d + 7 * n); uint32_t* dCtxOut = d + 8 * n; int32_t* dTile = reinterpret_cast<int32_t*>(d + 9 * n);
		cudaError_t e = cudaMemcpyAsync(dCtx, ctx, n * 4, cudaMemcpyHostToDevice, stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(dWid, wid, n * 4, cudaMemcpyHostToDevice, stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(dNode, node, n * 4, cudaMemcpyHostToDevice, stream);
		if (e == cudaSuccess) e = launch_cong_debug(n, dCtx, dWid, dNode, dDot, dEps, dNodeOut, dCtxOut, dTile, stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(outDot, dDot, n * 4, cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(outEps, dEps, n * 12, cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(outNode, dNodeOut, n * 4, cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(outCtx, dCtxOut, n * 4, cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(outTile, dTile, (size_t)nU * nW * 4, cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
		cudaFree(d);
		ck(e, "debugCong");
	}

	void Engine::setConfig(const kb2_config& cfg)
	{
		model.dev.cfg = cfg;
		model.header.config = cfg;
		if (g_constantsOwner == &model) g_constantsOwner = nullptr;
	}

	void Engine::debugTiming(uint32_t n, unsigned long long* out)
	{
		if (!main_.vv.timing || n > main_.capSent) throw std::runtime_error("debugTiming: no launch of that size yet");
		ck(cudaMemcpy(out, main_.vv.timing, (size_t)n * 16, cudaMemcpyDeviceToHost), "debugTiming");
	}

	int Engine::debugLattice(const uint16_t* text, uint32_t len, uint32_t matchOptions, std::vector<int32_t>& rows)
	{
		const uint32_t off[2] = { 0, len };
		const size_t U = 2 * (size_t)len + 4;
		Scratch& sc = main_;
		ensureScratch(sc, U, 1, DEFAULT_PATHS_PER_UNIT, DEFAULT_PATHS_CONST, KB_DEFAULT_NODES_PER_UNIT);
		ck(cudaMemcpyAsync(sc.dText, text, (size_t)len * 2, cudaMemcpyHostToDevice, stream), "H2D");
		ck(cudaMemcpyAsync(sc.dOff, off, 8, cudaMemcpyHostToDevice, stream), "H2D");
		bind(sc, sc.dText, sc.dOff, 1, matchOptions);
		ck(cudaMemsetAsync(sc.order, 0, 4, stream), "memset");
		sc.bv.order = sc.order;
		if (g_constantsOwner != &model) { ck(set_model_lattice(model.dev), "constant upload"); ck(model.dev.model_type == 4 ? set_model_viterbi_cong(model.dev) : set_model_viterbi(model.dev), "constant upload"); ck(set_model_emit(model.dev), "constant upload"); g_constantsOwner = &model; }
		ck(launch_lattice(model.dev, sc.bv, stream), "lattice launch");
		uint32_t nChunks = 0, status = 0;
		ck(cudaMemcpyAsync(&nChunks, sc.bv.n_chunks, 4, cudaMemcpyDeviceToHost, stream), "D2H");
		ck(cudaMemcpyAsync(&status, sc.bv.status, 4, cudaMemcpyDeviceToHost, stream), "D2H");
		ck(cudaStreamSynchronize(stream), "sync");
		if (status) return -(int)status;
		std::vector<DChunk> chunks(nChunks);
		if (nChunks) ck(cudaMemcpy(chunks.data(), sc.bv.chunks, nChunks * sizeof(DChunk), cudaMemcpyDeviceToHost), "D2H chunks");
		rows.clear();
		int total = 0;
		for (uint32_t c = 0; c < nChunks; ++c)
		{
			std::vector<DNode> nodes(chunks[c].n_nodes);
			ck(cudaMemcpy(nodes.data(), sc.bv.nodes + chunks[c].node_off, nodes.size() * sizeof(DNode), cudaMemcpyDeviceToHost), "D2H nodes");
			for (auto& nd : nodes)
			{
				const int32_t r[9] = { nd.form, nd.uform_len ? (int32_t)nd.uform_off : -1, (int32_t)nd.uform_len, nd.prev, nd.sibling,
					(int32_t)nd.start_pos, (int32_t)nd.end_pos, nd.space_errors, (int32_t)c };
				rows.insert(rows.end(), r, r + 9);
				++total;
			}
		}
		return total;
	}
}
