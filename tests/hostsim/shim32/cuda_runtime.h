// HOST SIMULATION shim, 32-lane variant (test infrastructure, tests/ only).  The kernel SOURCES of kiwi_b200/csrc compile
// as plain C++; a warp is 32 fibers (one per lane, one OS thread) that run the kernel body and hand over at every warp
// collective (__ballot_sync, __shfl_sync, __match_any_sync, __syncwarp, ...), so lane arithmetic, ballots, shuffles,
// shared memory and the __syncwarp-ordered lane-0 stores behave as on the device (full-mask collectives by all lanes, as the
// kernels use them).  What it cannot show: timing, instruction-cache effects, real memory-model races between collectives,
// multi-warp blocks (one warp per block here; the lockstep barrier is compiled out), tensor-core PTX (Knlm build only).
// Two context switches per lane and collective: a few hundred sentences per minute.  Never linked into the product.
#pragma once
// (every standard header the kernel sources and harnesses use comes first: libstdc++ spells __attribute__((__noinline__)) itself)
#include <algorithm>
#include <cstdio>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include <atomic>
#include <condition_variable>
#include <dlfcn.h>
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <type_traits>
#include <vector>
#if !defined(KB_HOSTSIM) || KB_HOSTSIM != 32
#error "this shim is for -DKB_HOSTSIM=32 builds"
#endif
#define __device__
#define __host__
#define __global__
#define __constant__ static
#define __shared__ static
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) alignas(n)

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(8) int2 { int32_t x, y; };
struct alignas(8) float2 { float x, y; };
inline int2 make_int2(int32_t x, int32_t y) { return int2{ x, y }; }
inline float2 make_float2(float x, float y) { return float2{ x, y }; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{ x, y }; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{ x, y, z, w }; }

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize, cudaFuncAttributePreferredSharedMemoryCarveout };
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? 0 : 2; }
enum { cudaHostAllocPortable = 1 };
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "host simulation"; }
template<class T> inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) { std::memcpy(&sym, src, n); return 0; }
template<class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

namespace simt
{
	constexpr uint32_t W = 32;          // lanes of a warp
	constexpr uint32_t TMAX = 512;      // threads of a block (multi-warp blocks: collectives are per warp, barriers per block)
	struct Dim { uint32_t x = 0, y = 0, z = 0; };
	// one fiber per lane inside ONE OS thread: the per-lane "registers" below are swapped by the scheduler at every switch
	inline uint32_t lane = 0;           // lane inside the warp
	inline uint32_t tid = 0;            // thread inside the block (fiber index)
	inline Dim threadIdx, blockIdx;
	inline Dim blockDim, gridDim;

	// Cooperative scheduler: exactly one lane runs at a time, always the HIGHEST runnable lane, and a lane runs until its next
	// collective.  Between two collectives the lanes therefore execute in the order 31, 30, ..., 0: lane 0 — the lane that issues
	// the kernels' single-lane stores — runs last, so the other lanes' loads of that segment come first, as on a converged warp
	// (loads of an earlier instruction precede a later instruction's store).  Deterministic; a collective that not all lanes
	// reach (or that they reach from different call sites) is reported with the call sites instead of hanging.
	// Lanes are fibers (round 2; round 1 used 32 OS threads and condition variables, ~100x slower per collective).
	// minimal x86-64 System V context switch (callee-saved registers + stack pointer; no signal-mask system call as in swapcontext)
	struct Ctx { void* sp = nullptr; };
	extern "C" void kb_simt_switch(void** saveSp, void* loadSp);
	asm(R"(
	.text
	.weak kb_simt_switch
	.type kb_simt_switch,@function
kb_simt_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	subq $8, %rsp
	stmxcsr (%rsp)
	fnstcw 4(%rsp)
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	ldmxcsr (%rsp)
	fldcw 4(%rsp)
	addq $8, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
	.size kb_simt_switch,.-kb_simt_switch
)");
	enum St : uint8_t { RUN, WAIT_FULL, WAIT_PART, WAIT_BAR, DONE };
	struct Sched
	{
		Ctx mainCtx; Ctx ctx[TMAX];
		std::vector<char> stacks;
		uint32_t nLanes = W;                 // threads of the block
		St st[TMAX];
		void* site[TMAX];
		uint64_t fullVal[TMAX], fullRes[TMAX];
		uint64_t partTag[TMAX]; unsigned partGrp[TMAX]; uint32_t partVal[TMAX]; uint32_t partSnap[TMAX][W];
		uint32_t barId[TMAX], barCount[TMAX];      // bar.sync id, count (count 0 = every thread that has not exited: __syncthreads)
		uint64_t collectives = 0; uint64_t ep[TMAX];
		uint64_t epochOf[TMAX], subEpochOf[TMAX], subCountOf[TMAX];
	};
	inline Sched* sched = nullptr;
	inline const bool ascending = std::getenv("HS32_ASCENDING") != nullptr;      // experiment: lowest runnable lane first
	inline uint64_t epoch = 0, subEpoch = 0, subCount = 0;

	[[noreturn]] inline void die(Sched& sc, const char* what)
	{
		std::fprintf(stderr, "[simt] %s; lane:state@site (addr2line the sites):", what);
		for (uint32_t l = 0; l < sc.nLanes; ++l)
		{
			Dl_info di; uintptr_t off = (uintptr_t)sc.site[l];
			if (sc.site[l] && dladdr(sc.site[l], &di) && di.dli_fbase) off -= (uintptr_t)di.dli_fbase;      // offset inside the shared object
			std::fprintf(stderr, " %u:%d@0x%zx#%llu", l, (int)sc.st[l], (size_t)off, (unsigned long long)sc.ep[l]);
		}
		std::fprintf(stderr, "\n");
		std::abort();
	}

	// main context.  Releases complete collectives, then returns the highest runnable lane (-1: all lanes done).
	// completion checks for the things thread `me` may have just completed by arriving (or by exiting)
	inline void onArrive(Sched& sc, uint32_t me)
	{
		const uint32_t wb = me & ~(W - 1), we = std::min(wb + W, sc.nLanes);
		// full-warp collective of my warp: every lane of the warp that has not exited waits in one
		{
			bool allFull = true, any = false;
			for (uint32_t l = wb; l < we; ++l) if (sc.st[l] != DONE) { any = true; if (sc.st[l] != WAIT_FULL) allFull = false; }
			if (any && allFull)
			{
				// (call sites are not compared: the compiler duplicates one source-level collective into several branches)
				for (uint32_t l = wb; l < we; ++l) sc.fullRes[l] = sc.fullVal[l];
				for (uint32_t l = wb; l < we; ++l) if (sc.st[l] == WAIT_FULL) sc.st[l] = RUN;
				++sc.collectives;
			}
		}
		if (sc.st[me] == WAIT_PART)
		{
			const uint32_t l = me;
			bool ready = true;
			for (uint32_t k = 0; k < W; ++k) if ((sc.partGrp[l] >> k & 1) && !(wb + k < sc.nLanes && sc.st[wb + k] == WAIT_PART && sc.partTag[wb + k] == sc.partTag[l])) ready = false;
			if (ready)
			{
				const unsigned grp = sc.partGrp[l];
				for (uint32_t k = 0; k < W; ++k) if (grp >> k & 1) for (uint32_t j = 0; j < W; ++j) if (grp >> j & 1) sc.partSnap[wb + k][j] = sc.partVal[wb + j];
				for (uint32_t k = 0; k < W; ++k) if (grp >> k & 1) sc.st[wb + k] = RUN;
			}
		}
		// block barriers (bar.sync id, count): complete when `count` threads wait on the id (count 0: all threads that have not exited)
		if (sc.st[me] == WAIT_BAR || sc.st[me] == DONE)
		{
			for (uint32_t pass = 0; pass < 1; ++pass)
			{
				uint32_t alive = 0;
				for (uint32_t k = 0; k < sc.nLanes; ++k) if (sc.st[k] != DONE) ++alive;
				// ids that may complete: mine (arrival), or any id-0 style barrier when a thread exited
				for (uint32_t l = 0; l < sc.nLanes; ++l)
				{
					if (sc.st[l] != WAIT_BAR) continue;
					if (sc.st[me] == WAIT_BAR && sc.barId[l] != sc.barId[me]) continue;
					const uint32_t id = sc.barId[l];
					uint32_t waiting = 0;
					for (uint32_t k = 0; k < sc.nLanes; ++k) if (sc.st[k] == WAIT_BAR && sc.barId[k] == id) ++waiting;
					const uint32_t need = sc.barCount[l] ? sc.barCount[l] : alive;
					if (waiting >= need) for (uint32_t k = 0; k < sc.nLanes; ++k) if (sc.st[k] == WAIT_BAR && sc.barId[k] == id) sc.st[k] = RUN;
					if (sc.st[me] == WAIT_BAR || sc.st[me] == RUN) break;      // my id has been checked
				}
			}
		}
	}

	// main context.  Returns the highest runnable lane (-1: all lanes done).
	inline int reschedule(Sched& sc)
	{
		int next = -1;
		if (ascending) { for (int l = 0; l < (int)sc.nLanes; ++l) if (sc.st[l] == RUN) { next = l; break; } }
		else for (int l = (int)sc.nLanes - 1; l >= 0; --l) if (sc.st[l] == RUN) { next = l; break; }
		if (next < 0)
		{
			bool allDone = true;
			for (uint32_t l = 0; l < sc.nLanes; ++l) if (sc.st[l] != DONE) allDone = false;
			if (!allDone) die(sc, "deadlock: no lane can run (a collective that some lanes never reach)");
		}
		return next;
	}

	// lane context: park this lane in `state` and give the processor back to the scheduler
	inline void blockAs(St state, void* site)
	{
		Sched& sc = *sched;
		const uint32_t me = tid;
		sc.st[me] = state; sc.site[me] = site; sc.ep[me] = epoch;
		sc.epochOf[me] = epoch; sc.subEpochOf[me] = subEpoch; sc.subCountOf[me] = subCount;
		onArrive(sc, me);
		if (sc.st[me] == RUN && !ascending)
		{
			// I completed the collective: the highest runnable lane goes next - keep running when that is me
			bool higher = false;
			for (uint32_t l = sc.nLanes; l-- > me + 1;) if (sc.st[l] == RUN) { higher = true; break; }
			if (!higher) return;
		}
		kb_simt_switch(&sc.ctx[me].sp, sc.mainCtx.sp);
	}

	template<class T> __attribute__((noinline)) void exchange(T v, T* out)
	{
		static_assert(sizeof(T) <= 8, "exchange of <= 8-byte values");
		uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
		sched->fullVal[tid] = raw;
		++epoch;
		blockAs(WAIT_FULL, __builtin_return_address(0));
		const uint32_t wb = tid & ~(W - 1);
		for (uint32_t i = 0; i < W; ++i) std::memcpy(&out[i], &sched->fullRes[wb + i], sizeof(T));
	}

	__attribute__((noinline)) inline void sync()
	{
		sched->fullVal[tid] = 0;
		++epoch;
		blockAs(WAIT_FULL, __builtin_return_address(0));
	}

	// bar.sync id, count (count in threads; 0 = __syncthreads over the threads that have not exited)
	__attribute__((noinline)) inline void bar_sync(uint32_t id, uint32_t count)
	{
		sched->barId[tid] = id; sched->barCount[tid] = count;
		blockAs(WAIT_BAR, __builtin_return_address(0));
	}

	// partial-mask collectives: a lane calls with the mask of ITS group; lanes outside any multi-member group may skip the call.
	// The members of a group execute the same sequence of partial collectives between two full-warp collectives.
	template<class Fn> inline uint32_t groupReduce(unsigned grp, uint32_t v, uint32_t init, Fn&& fn)
	{
		if (grp == 0xFFFFFFFFu)
		{
			uint32_t o[W]; exchange<uint32_t>(v, o);
			uint32_t r = init;
			for (uint32_t l = 0; l < W; ++l) r = fn(r, o[l]);
			return r;
		}
		if (subEpoch != epoch) { subEpoch = epoch; subCount = 0; }
		sched->partTag[tid] = ((epoch + 1) << 16) | ++subCount;
		sched->partGrp[tid] = grp; sched->partVal[tid] = v;
		blockAs(WAIT_PART, __builtin_return_address(0));
		uint32_t r = init;
		for (uint32_t l = 0; l < W; ++l) if (grp >> l & 1) r = fn(r, sched->partSnap[tid][l]);
		return r;
	}

	// mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 (D = A * B + C) from its fragment layout (g = lane / 4, t = lane % 4):
	// A 16x32 u8 row-major: a0 = (row g, k 4t..4t+3), a1 = (row g+8, same k), a2 = (row g, k 16+4t..), a3 = (row g+8, k 16+4t..);
	// B 32x8 s8 col-major: b0 = (k 4t..4t+3, col g), b1 = (k 16+4t.., col g); C/D: c0 = (g, 2t), c1 = (g, 2t+1), c2 = (g+8, 2t), c3 = (g+8, 2t+1)
	inline void mma_m16n8k32_u8s8(int32_t& c0, int32_t& c1, int32_t& c2, int32_t& c3, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
	{
		uint32_t A0[W], A1[W], A2[W], A3[W], B0[W], B1[W];
		exchange<uint32_t>(a0, A0); exchange<uint32_t>(a1, A1); exchange<uint32_t>(a2, A2); exchange<uint32_t>(a3, A3);
		exchange<uint32_t>(b0, B0); exchange<uint32_t>(b1, B1);
		auto dot = [](uint32_t ua, uint32_t sb) { int32_t s = 0; for (int j = 0; j < 4; ++j) s += (int32_t)(uint8_t)(ua >> (8 * j)) * (int32_t)(int8_t)(sb >> (8 * j)); return s; };
		const uint32_t g = lane >> 2, t = lane & 3;
		for (uint32_t tt = 0; tt < 4; ++tt)
		{
			const uint32_t la = g * 4 + tt, lb0 = (2 * t) * 4 + tt, lb1 = (2 * t + 1) * 4 + tt;
			c0 += dot(A0[la], B0[lb0]) + dot(A2[la], B1[lb0]);
			c1 += dot(A0[la], B0[lb1]) + dot(A2[la], B1[lb1]);
			c2 += dot(A1[la], B0[lb0]) + dot(A3[la], B1[lb0]);
			c3 += dot(A1[la], B0[lb1]) + dot(A3[la], B1[lb1]);
		}
	}

	// debugging aid for kernels: all lanes must hold the same value here (warp-uniform state)
	inline void check_uniform(uint32_t v, int line)
	{
		uint32_t o[W]; exchange<uint32_t>(v, o);
		for (uint32_t i = 0; i < W; ++i) if (o[i] != v && lane == 0)
		{
			std::fprintf(stderr, "[simt] non-uniform value at line %d: lane 0 has %u, lane %u has %u\n", line, v, i, o[i]);
			std::abort();
		}
	}

	// one warp per block; blockDimX <= 32 lanes run (thread-per-item kernels pass 1)
	inline void (*fiberBody)(void*) = nullptr;
	inline void* fiberArg = nullptr;
	inline void fiberEntry()
	{
		fiberBody(fiberArg);
		sched->st[tid] = DONE; sched->site[tid] = nullptr;
		onArrive(*sched, tid);
		kb_simt_switch(&sched->ctx[tid].sp, sched->mainCtx.sp);
		std::abort();      // a finished lane is never resumed
	}
	template<class F> inline void launch(uint32_t blocks, uint32_t blockDimX, F&& body)
	{
		const uint32_t lanes = std::min(blockDimX, TMAX);
		blockDim.x = blockDimX; gridDim.x = blocks;
		constexpr size_t STACK = 1u << 20;
		static Sched* scp = new Sched;          // lane stacks are allocated once per process and reused by every launch
		Sched& sc = *scp;
		if (sc.stacks.size() < STACK * lanes) sc.stacks.resize(STACK * lanes);
		using Fn = std::remove_reference_t<F>;
		fiberBody = [](void* a) { (*reinterpret_cast<Fn*>(a))(); };
		fiberArg = (void*)&body;
		for (uint32_t b = 0; b < blocks; ++b)
		{
			sc.nLanes = lanes; sc.collectives = 0;
			for (uint32_t l = 0; l < TMAX; ++l) { sc.st[l] = l < lanes ? RUN : DONE; sc.site[l] = nullptr; sc.partTag[l] = 0; sc.epochOf[l] = sc.subEpochOf[l] = sc.subCountOf[l] = 0; }
			for (uint32_t l = 0; l < lanes; ++l)
			{
				// initial frame: [mxcsr/fpcw][r15 r14 r13 r12 rbx rbp][return address = fiberEntry][alignment slot]
				uintptr_t top = (uintptr_t)(sc.stacks.data() + STACK * (l + 1)); top &= ~(uintptr_t)15;
				uint64_t* sp = reinterpret_cast<uint64_t*>(top) - 11;
				uint32_t csr[2] = { 0x1F80u, 0x037Fu };
				std::memcpy(&sp[0], csr, 8);
				for (int k = 1; k <= 6; ++k) sp[k] = 0;
				sp[7] = (uint64_t)(uintptr_t)&fiberEntry; sp[8] = 0; sp[9] = 0; sp[10] = 0;      // after `ret`, rsp = &sp[8] = top - 24: 8 mod 16 as at a call boundary
				sc.ctx[l].sp = sp;
			}
			sched = &sc;
			while (true)
			{
				const int next = reschedule(sc);
				if (next < 0) break;
				tid = (uint32_t)next; lane = tid & (W - 1); threadIdx.x = tid; blockIdx.x = b;
				epoch = sc.epochOf[next]; subEpoch = sc.subEpochOf[next]; subCount = sc.subCountOf[next];
				kb_simt_switch(&sc.mainCtx.sp, sc.ctx[next].sp);
			}
			if (std::getenv("HS32_TRACE")) std::fprintf(stderr, "[simt] block %u: %llu full-warp collectives\n", b, (unsigned long long)sc.collectives);
			sched = nullptr;
		}
	}
}
using simt::threadIdx; using simt::blockIdx; using simt::blockDim; using simt::gridDim;

inline unsigned __ballot_sync(unsigned, bool p) { uint32_t o[32]; simt::exchange<uint32_t>(p ? 1u : 0u, o); unsigned m = 0; for (int i = 0; i < 32; ++i) m |= o[i] << i; return m; }
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xFFFFFFFFu; }
template<class T> inline T __shfl_sync(unsigned, T v, int src) { T o[32]; simt::exchange<T>(v, o); return o[src & 31]; }
template<class T> inline T __shfl_up_sync(unsigned, T v, int d) { T o[32]; simt::exchange<T>(v, o); return (int)simt::lane >= d ? o[simt::lane - d] : v; }
template<class T> inline T __shfl_down_sync(unsigned, T v, int d) { T o[32]; simt::exchange<T>(v, o); return simt::lane + d < 32 ? o[simt::lane + d] : v; }
template<class T> inline T __shfl_xor_sync(unsigned, T v, int d) { T o[32]; simt::exchange<T>(v, o); return o[(simt::lane ^ d) & 31]; }
template<class T> inline unsigned __match_any_sync(unsigned, T key) { T o[32]; simt::exchange<T>(key, o); unsigned m = 0; for (int i = 0; i < 32; ++i) if (o[i] == key) m |= 1u << i; return m; }
inline uint32_t __reduce_max_sync(unsigned grp, uint32_t v) { return simt::groupReduce(grp, v, 0u, [](uint32_t a, uint32_t b) { return std::max(a, b); }); }
inline uint32_t __reduce_min_sync(unsigned grp, uint32_t v) { return simt::groupReduce(grp, v, 0xFFFFFFFFu, [](uint32_t a, uint32_t b) { return std::min(a, b); }); }
inline uint32_t __reduce_or_sync(unsigned grp, uint32_t v) { return simt::groupReduce(grp, v, 0u, [](uint32_t a, uint32_t b) { return a | b; }); }
inline uint32_t __reduce_add_sync(unsigned grp, uint32_t v) { return simt::groupReduce(grp, v, 0u, [](uint32_t a, uint32_t b) { return a + b; }); }
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { simt::sync(); }
inline void __syncthreads() { simt::bar_sync(0, 0); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __int2float_rn(int v) { return (float)v; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline int32_t __float_as_int(float f) { int32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __dp4a(int a, int b, int c) { for (int i = 0; i < 4; ++i) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i)); return c; }
inline int __dp4a(unsigned a, int b, int c) { for (int i = 0; i < 4; ++i) c += (int)(uint8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i)); return c; }
template<class T> inline T atomicAdd(T* p, T v) { return __sync_fetch_and_add(p, v); }
template<class T> inline T atomicSub(T* p, T v) { return __sync_fetch_and_sub(p, v); }
template<class T> inline T atomicCAS(T* p, T cmp, T val) { return __sync_val_compare_and_swap(p, cmp, val); }
template<class T> inline T atomicOr(T* p, T v) { return __sync_fetch_and_or(p, v); }
template<class T> inline T atomicMax(T* p, T v) { T o = *p; while (o < v && !__sync_bool_compare_and_swap(p, o, v)) o = *p; return o; }
template<class T> inline T __ldg(const T* p) { return *p; }
template<class A, class B> inline std::common_type_t<A, B> min(A a, B b) { return a < b ? a : b; }
template<class A, class B> inline std::common_type_t<A, B> max(A a, B b) { return a < b ? b : a; }
