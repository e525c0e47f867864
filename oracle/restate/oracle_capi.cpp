// ORACLE C API (test infrastructure; loaded only by tests/, __graft_entry__.smoke() and bench.py's cpu legs).
// Thin extern "C" wrapper over the CPU restatement so Python can compare it with the CUDA path through ctypes.
#include <algorithm>
#include "analyze.hpp"
#include "typo.hpp"

struct OrcHandle { orc::Image im; orc::Analyzer* an = nullptr; orc::Counters cnt; orc::EvalStats es; };

extern "C" {

void* orc_open(const char* imagePath)
{
	try
	{
		auto* h = new OrcHandle;
		h->im.load(imagePath);
		h->an = new orc::Analyzer{ h->im };
		h->an->viterbi.cnt = &h->cnt;
		h->an->enableWorkCounters();
		return h;
	}
	catch (...) { return nullptr; }
}

void orc_close(void* p)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	if (!h) return;
	delete h->an;
	delete h;
}

// returns the token count (or -1 on error / -2 when max_tokens is too small)
int orc_analyze(void* p, const uint16_t* text, int len, uint32_t* morph, uint8_t* tag, uint32_t* pos, uint16_t* length, float* score, int maxTokens, float* sentScore)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	try
	{
		h->an->keepChunks = false;
		auto res = h->an->analyze(text, (size_t)len);
		if ((int)res.tokens.size() > maxTokens) return -2;
		for (size_t i = 0; i < res.tokens.size(); ++i)
		{
			morph[i] = res.tokens[i].morph; tag[i] = res.tokens[i].tag; pos[i] = res.tokens[i].position; length[i] = res.tokens[i].length; score[i] = res.tokens[i].score;
		}
		*sentScore = res.score;
		return (int)res.tokens.size();
	}
	catch (...) { return -1; }
}

// lattice rows {form, uform_off|-1, uform_len, prev, sibling, start, end, space_errors, chunk}; chunks with <= 2 nodes are skipped
int orc_lattice(void* p, const uint16_t* text, int len, int32_t* rows, int maxRows)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	try
	{
		h->an->keepChunks = true;
		auto res = h->an->analyze(text, (size_t)len);
		int n = 0, c = 0;
		for (auto& ch : res.chunks)
		{
			if (ch.nodes.size() <= 2) continue;
			for (auto& nd : ch.nodes)
			{
				if (n >= maxRows) return -2;
				int32_t* r = rows + 9 * n++;
				r[0] = nd.form; r[1] = nd.uformLen ? nd.uformOff : -1; r[2] = (int32_t)nd.uformLen; r[3] = (int32_t)nd.prev; r[4] = (int32_t)nd.sibling;
				r[5] = (int32_t)nd.startPos; r[6] = (int32_t)nd.endPos; r[7] = (int32_t)nd.spaceErrors; r[8] = c;
			}
			++c;
		}
		return n;
	}
	catch (...) { return -1; }
}

// byte-model counters accumulated since open, as 18 uint64 in declaration order of orc::WorkCounters
void orc_work_counters(void* p, uint64_t* out)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	static_assert(sizeof(orc::WorkCounters) == 20 * sizeof(uint64_t), "WorkCounters layout");
	std::memcpy(out, &h->an->work, 18 * sizeof(uint64_t));
}

// per-evaluate() shape rows {P, nCands, lmSteps, flags, pathsBeforePrune}: enable, then fetch (returns the row count; copies at most maxRows)
void orc_eval_stats_enable(void* p) { auto* h = reinterpret_cast<OrcHandle*>(p); h->an->viterbi.es = &h->es; }
uint64_t orc_eval_stats(void* p, uint32_t* out, uint64_t maxRows)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	const uint64_t n = h->es.rows.size() / 5;
	std::memcpy(out, h->es.rows.data(), std::min(n, maxRows) * 5 * sizeof(uint32_t));
	return n;
}

// forget the growth history of the `top1` container (= what a fresh reference process starts with)
void orc_reset_history(void* p)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	h->an->viterbi.resetHistory();
}
void orc_reset_history_buckets(void* p, uint64_t buckets)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	h->an->viterbi.resetHistory((size_t)buckets);
}

// CoNg byte-model counters: {unique rows gathered (contexts + outputs), int8 MACs of the per-node gather GEMMs}
void orc_cong_counters(void* p, uint64_t* out)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	out[0] = h->an->work.cgRows; out[1] = h->an->work.cgMacs;
}

// exact integer part and the three float epilogues of one (context row, output row) pair: {acc - hsum, E_scalar, E_small, E_gemv}
int orc_cong_pair(void* p, uint32_t ctx, uint32_t wid, int32_t* accOut, float* eps)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	const auto& cg = h->an->viterbi.cg;
	if (!cg.nodes || ctx >= h->im.h->cg_context_size || wid >= h->im.h->lang_vocab_size) return -1;
	*accOut = cg.dotMinusHsum(ctx, wid);
	eps[0] = cg.finish(ctx, wid, orc::Cong::E_scalar); eps[1] = cg.finish(ctx, wid, orc::Cong::E_small); eps[2] = cg.finish(ctx, wid, orc::Cong::E_gemv);
	return 0;
}

// which epilogue the reference's kernel dispatch uses for m unique contexts x n unique outputs (0 scalar is never returned)
int orc_cong_epilogue(int m, int n) { return (int)orc::Cong::epilogueOf((size_t)m, (size_t)n); }

// one context-trie transition: returns the new contextIdx, *node is updated
uint32_t orc_cong_step(void* p, int32_t* node, uint32_t wid)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	return h->an->viterbi.cg.step(*node, wid);
}

// work counters accumulated since open: {lmSteps, pairs, inserts, pathsOut, candEvals, evalCalls}; orc_counters2 adds {top1Mode, bucketFull, mediumMode, maxIncoming}
void orc_counters(void* p, uint64_t* out)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	out[0] = h->cnt.lmSteps; out[1] = h->cnt.pairs; out[2] = h->cnt.inserts; out[3] = h->cnt.pathsOut; out[4] = h->cnt.candEvals; out[5] = h->cnt.evalCalls;
}
void orc_counters2(void* p, uint64_t* out)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	out[0] = h->cnt.top1Mode; out[1] = h->cnt.bucketFull; out[2] = h->cnt.mediumMode; out[3] = h->cnt.maxIncoming;
}


// ---- typo graph: restated PreparedTypoTransformer::generateGraph over a flat typo image
struct OrcTypo { orc::TypoImage im; };

void* orc_typo_open(const char* path)
{
	try { auto* t = new OrcTypo; t->im.load(path); return t; }
	catch (...) { return nullptr; }
}
void orc_typo_close(void* p) { delete reinterpret_cast<OrcTypo*>(p); }

// AnalyzeOption::withTypoTransformer (include/kiwi/Kiwi.h:127-133): analyse with a typo lattice from now on (typo == nullptr: off).
// The typo handle must outlive the analyzer's use of it.
void orc_set_blocklist(void* p, const uint32_t* ids, int n) { auto& b = reinterpret_cast<OrcHandle*>(p)->an->viterbi.blocklist; b.assign(ids, ids + n); std::sort(b.begin(), b.end()); }

void orc_set_open_ending(void* p, int on) { reinterpret_cast<OrcHandle*>(p)->an->openEnding = on != 0; }

void orc_set_typo(void* p, void* typo, float threshold)
{
	auto* h = reinterpret_cast<OrcHandle*>(p);
	h->an->splitter.typoImg = typo ? &reinterpret_cast<OrcTypo*>(typo)->im : nullptr;
	h->an->splitter.typoThreshold = threshold;
}

// raw UTF-16 text -> normalizeHangul -> graph; rows of 9 int32 {endPos, typoCost bits, prevOffset, siblingOffset, continualTypoIdx, dialect,
// fromPool, off, len}; returns the node count (or -1 / -2), *normLen = length of the normalised string
int orc_typo_graph(void* p, const uint16_t* text, int len, int32_t* rows, int maxRows, int* normLen)
{
	try
	{
		auto* t = reinterpret_cast<OrcTypo*>(p);
		const std::u16string norm = orc::normalizeHangulPlain(std::u16string(reinterpret_cast<const char16_t*>(text), (size_t)len));
		*normLen = (int)norm.size();
		orc::TypoGraph tg{ t->im };
		const auto g = tg.generate(norm);
		if ((int)g.size() > maxRows) return -2;
		for (size_t i = 0; i < g.size(); ++i)
		{
			int32_t* r = rows + 9 * i;
			int32_t bits; std::memcpy(&bits, &g[i].typoCost, 4);
			r[0] = (int32_t)g[i].endPos; r[1] = bits; r[2] = (int32_t)g[i].prevOffset; r[3] = (int32_t)g[i].siblingOffset; r[4] = g[i].continualTypoIdx;
			r[5] = g[i].dialect; r[6] = g[i].fromPool ? 1 : 0; r[7] = (int32_t)g[i].off; r[8] = (int32_t)g[i].len;
		}
		return (int)g.size();
	}
	catch (...) { return -1; }
}

}
