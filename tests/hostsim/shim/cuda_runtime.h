// HOST SIMULATION shim (test infrastructure, tests/ only): lets the kernel SOURCES of kiwi_b200/csrc compile as plain C++ so
// that their logic can be checked against the golden vectors on a machine without a GPU.  A "warp" is ONE lane here (the
// kernels take their lane count from KB_W, 1 under KB_HOSTSIM): every lane-strided loop degenerates to a sequential loop and
// every warp collective to the identity, which preserves the algorithm and nothing of the parallel execution (races, missing
// __syncwarp, 32-lane arithmetic are NOT exercised: the -m gpu parity tests stay the proof for the device build).
// Never linked into the product: libkiwi_b200.so is built by nvcc from the same sources with the real CUDA runtime.
#pragma once
// (every standard header the kernel sources and harnesses use comes first: libstdc++ spells __attribute__((__noinline__)) itself)
#include <algorithm>
#include <cstdio>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#ifndef KB_HOSTSIM
#error "the shim is only for -DKB_HOSTSIM builds"
#endif
#define __device__
#define __host__
#define __global__
#define __constant__ static
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
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
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? 0 : 2; }
enum { cudaHostAllocPortable = 1 };
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "host simulation"; }
template<class T> inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) { std::memcpy(&sym, src, n); return 0; }

// one-lane warp collectives
inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
inline bool __any_sync(unsigned, bool p) { return p; }
inline bool __all_sync(unsigned, bool p) { return p; }
inline uint32_t __reduce_or_sync(unsigned, uint32_t v) { return v; }
inline uint32_t __reduce_add_sync(unsigned, uint32_t v) { return v; }
inline uint32_t __reduce_min_sync(unsigned, uint32_t v) { return v; }
inline uint32_t __reduce_max_sync(unsigned, uint32_t v) { return v; }
template<class T> inline T __shfl_sync(unsigned, T v, int) { return v; }
template<class T> inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template<class T> inline T __shfl_down_sync(unsigned, T v, int) { return v; }
template<class T> inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
inline unsigned __match_any_sync(unsigned, uint32_t) { return 1u; }
inline void __syncwarp(unsigned = 0xFFFFFFFFu) {}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
template<class T> inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
template<class T> inline T __ldg(const T* p) { return *p; }
using std::min;
using std::max;
