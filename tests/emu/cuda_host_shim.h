// Host shim for running CUDA kernel TEXT on CPU threads (test infrastructure; see tests/test_kernel_models.py).
// A kernel region extracted from a .cu file is compiled by g++ with the CUDA keywords mapped as below and launched with
// emu_launch(): one std::thread per CUDA thread of ONE CTA at a time, __syncthreads() = std::barrier over the CTA,
// warp shuffles / match.any = per-warp barrier + exchange buffer, atomics = std::atomic_ref, directed-rounding adds via
// fesetround. Compile with -std=c++20 -pthread -ffp-contract=off -frounding-math (no FMA contraction: the kernels pin
// their rounding with __dmul_rn / __dadd_rn).
#pragma once
#include <atomic>
#include <barrier>
#include <cfenv>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct double2 { double x, y; };
struct int2 { int x, y; };
struct uint2 { unsigned int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline double2 make_double2(double x, double y) { return double2{ x, y }; }
static inline int2 make_int2(int x, int y) { return int2{ x, y }; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{ x, y, z, w }; }

#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__
#define __align__(n) alignas(n)

struct EmuIdx { int x = 0, y = 0, z = 0; };
static thread_local EmuIdx threadIdx, blockIdx;
static EmuIdx blockDim, gridDim;

struct EmuWarp {
	std::barrier<> bar{ 32 };
	unsigned long long buf[32];
};
static thread_local std::barrier<>* g_ctaBarrier = nullptr;   // (thread-local: the cluster emulation runs several CTAs at once)
static thread_local EmuWarp* g_warp = nullptr;
static thread_local int g_lane = 0;
static inline void __syncthreads() { g_ctaBarrier->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { g_warp->bar.arrive_and_wait(); }
static inline void __threadfence_system() {}
static inline void __threadfence_block() {}

template <class T> static inline T emu_exchange(T x, int src)
{
	static_assert(sizeof(T) <= 8, "shuffle payload");
	unsigned long long w = 0;
	std::memcpy(&w, &x, sizeof(T));
	g_warp->buf[g_lane] = w;
	g_warp->bar.arrive_and_wait();
	const unsigned long long r = g_warp->buf[src & 31];
	g_warp->bar.arrive_and_wait();
	T out;
	std::memcpy(&out, &r, sizeof(T));
	return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T x, int o) { return emu_exchange(x, g_lane ^ o); }
template <class T> static inline T __shfl_sync(unsigned, T x, int src) { return emu_exchange(x, src); }
template <class T> static inline T __shfl_up_sync(unsigned, T x, int d) { return emu_exchange(x, g_lane >= d ? g_lane - d : g_lane); }
static inline unsigned int __match_any_sync(unsigned, int key)
{
	g_warp->buf[g_lane] = (unsigned long long)(unsigned int)key;
	g_warp->bar.arrive_and_wait();
	unsigned int m = 0;
	for (int l = 0; l < 32; ++l) m |= (g_warp->buf[l] == (unsigned long long)(unsigned int)key ? 1u : 0u) << l;
	g_warp->bar.arrive_and_wait();
	return m;
}
static inline unsigned int __ballot_sync(unsigned, bool pred)
{
	g_warp->buf[g_lane] = pred ? 1ULL : 0ULL;
	g_warp->bar.arrive_and_wait();
	unsigned int m = 0;
	for (int l = 0; l < 32; ++l) m |= (unsigned int)g_warp->buf[l] << l;
	g_warp->bar.arrive_and_wait();
	return m;
}
static inline bool __all_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }
// redux.sync over the lanes named in `mask`. On the device the lanes of a warp may name different (disjoint) masks in one
// instruction - the groups then execute one after the other (WARPSYNC.EXCLUSIVE); here all 32 lanes arrive together and each
// takes the sum / maximum over its own mask, checking that its partners named the same one.
static inline unsigned int emu_redux(unsigned int mask, unsigned int x, bool isMax)
{
	if (!((mask >> g_lane) & 1u)) { std::fprintf(stderr, "redux: calling lane not in its mask\n"); std::abort(); }
	g_warp->buf[g_lane] = ((unsigned long long)mask << 32) | x;
	g_warp->bar.arrive_and_wait();
	unsigned int r = 0;
	for (int l = 0; l < 32; ++l)
		if ((mask >> l) & 1u) {
			if ((unsigned int)(g_warp->buf[l] >> 32) != mask) { std::fprintf(stderr, "redux: lanes of one group name different masks\n"); std::abort(); }
			const unsigned int v = (unsigned int)g_warp->buf[l];
			r = isMax ? (v > r ? v : r) : r + v;
		}
	g_warp->bar.arrive_and_wait();
	return r;
}
static inline unsigned int __reduce_add_sync(unsigned int mask, unsigned int x) { return emu_redux(mask, x, false); }
static inline unsigned int __reduce_max_sync(unsigned int mask, unsigned int x) { return emu_redux(mask, x, true); }
static inline unsigned int __brev(unsigned int x)
{
	unsigned int r = 0;
	for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
	return r;
}
static inline int __ffs(unsigned int x) { return x ? __builtin_ctz(x) + 1 : 0; }
static inline int __popc(unsigned int x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned int)x) : 32; }

template <class T> static inline T atomicAdd(T* p, T v) { return std::atomic_ref<T>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline double atomicAdd(double* p, double v)
{
	std::atomic_ref<double> a(*p);
	double old = a.load(std::memory_order_relaxed);
	while (!a.compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {}
	return old;
}
template <class T> static inline T atomicMax(T* p, T v)
{
	std::atomic_ref<T> a(*p);
	T old = a.load(std::memory_order_relaxed);
	while (old < v && !a.compare_exchange_weak(old, v, std::memory_order_relaxed)) {}
	return old;
}
template <class T> static inline T atomicMin(T* p, T v)
{
	std::atomic_ref<T> a(*p);
	T old = a.load(std::memory_order_relaxed);
	while (old > v && !a.compare_exchange_weak(old, v, std::memory_order_relaxed)) {}
	return old;
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void ptp_pdl_launch_dependents() {}
static inline void ptp_pdl_wait() {}
template <class T> static inline T atomicAdd_system(T* p, T v) { return atomicAdd(p, v); }
template <class T> static inline T atomicMax_system(T* p, T v) { return atomicMax(p, v); }

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dsqrt_rn(double a) { volatile double r = std::sqrt(a); return r; }
static inline double __dadd_rd(double a, double b)
{
	std::fesetround(FE_DOWNWARD);
	volatile double x = a, y = b;
	volatile double r = x + y;
	std::fesetround(FE_TONEAREST);
	return r;
}
static inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
static inline int __double2loint(double d) { return (int)(unsigned int)((unsigned long long)__double_as_longlong(d) & 0xffffffffULL); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
using std::floor;
using std::log;
using std::fabs;

// cp.async helpers of ptp_solve_wide.cu (synchronous on the host)
static inline void cpa8(void* dst, const void* src, bool valid) { *(double*)dst = valid ? *(const double*)src : 0.0; }
static inline void cpa_commit() {}
template <int N> static inline void cpa_wait() {}

// Run body() as the threads of one CTA (blockDim.x threads, a multiple of 32), for every block index of a gridX x gridY grid.
static inline void emu_launch(int gridX, int threads, const std::function<void()>& body, int gridY = 1)
{
	blockDim.x = threads; gridDim.x = gridX; gridDim.y = gridY;
	for (int by = 0; by < gridY; ++by)
		for (int b = 0; b < gridX; ++b) {
			std::barrier<> bar(threads);
			std::vector<std::unique_ptr<EmuWarp>> warps;
			for (int w = 0; w < threads / 32; ++w) warps.emplace_back(new EmuWarp);
			std::vector<std::thread> th;
			for (int t = 0; t < threads; ++t)
				th.emplace_back([&, t, b, by] {
					threadIdx.x = t; blockIdx.x = b; blockIdx.y = by;
					g_ctaBarrier = &bar;
					g_warp = warps[t / 32].get(); g_lane = t & 31;
					body();
				});
			for (auto& x : th) x.join();
		}
}
