// K1 / K2: fused gather + push + loss + deposit (and deposit alone) for one species.
//
// Replaces, per ring, Plasma::moveRings (reference Source/Plasma.cpp:100-120) with
// PenningTrap::getEField(int,double) (Source/PenningTrap.cpp:326-334) inlined, followed by the ring's
// contribution to Plasma::updateRHS (Source/Plasma.cpp:77-94) at its NEW position - so a ring is read
// once and written once per step (32 B of HBM traffic).
//
// Layout: rings are SoA (z[], v[]), bucketed by radial row (posR never changes, Source/Plasma.hpp:22-24),
// each bucket padded with NaN slots to PTP_ROW_ALIGN. A CTA owns "segments" = runs of tiles of one row.
// Deposition: every thread owns a private column of W axial-cell bins in shared memory
// (bins[cell - k0][thread]) and accumulates (count, sum of weights) there with plain read-modify-write -
// no atomics, no bank conflicts (column index = thread index), independent of how many rings share a
// cell (the default plasma puts ~5e5 rings of a 100 M load into each of ~400 cells). At the end of a
// segment the columns are tree-reduced and flushed with one global atomic per touched node. Rings whose
// cell falls outside the window [k0, k0+W) use global atomics directly (rare; the window is re-centred on
// the segment's measured cell range every step). 64-bit shared atomics are CAS loops on sm_100a
// (ATOMS.CAST.SPIN.64), which is why the accumulation is privatised instead.
//
// Node weights: with c_k = #rings in cell k and s_k = sum of their w, node k receives (c_k - s_k) + s_{k-1}
// in units of one ring; the -macroChargeDensity/epsilon0 factor (Source/Plasma.cpp:91-92) is applied when
// the Poisson solver reads the grid. Fixed-point mode packs (count:12 | sum:52) into one 64-bit word per
// bin with w quantised to 2^-F: all sums are exact integers, so the result is independent of summation
// order, CTA count and GPU count.
//
// Latency: the per-ring arithmetic is a ~40-deep dependent fp64 chain and only 8-16 warps fit next to the
// bins, so a thread handles 8 rings per tile in lock-step stages (branch-free fast path, rare cases
// deferred to fix-up loops) to give the scheduler 8 independent chains, and the loads of the next tile are
// issued before the current one is processed.
#include "ptp_internal.h"

#include <limits.h>

#include <algorithm>
#include <cstdlib>

namespace {

// [emu-begin] (tests/emu/emu_push.sh compiles the text between these markers for the host: tests/emu/emu_push.cpp)
constexpr double kMagic = 6755399441055744.0;            // 2^52 + 2^51: floor via add.rm, integer in the low word
constexpr unsigned long long kPackBias = 0x4320000000000000ULL; // bits(2^52 + x) - bias = (1 << 52) | x
constexpr unsigned long long kSumMask = (1ULL << 52) - 1;

struct PushArgs {
	int Nz, W, fixedBits, WE;   // W: cells of the thread-private deposit window, WE: cells of the (wider) field window
	double hz, invHz, eps, epsHi, length;
	double dt, charge, mass, invMass;
	double fixedScale;          // 2^fixedBits
	const double* eNodes;       // [G] node field of the pre-step potentials
	double* z;
	double* v;
	const PtpSegment* segs;
	const int* ctaSegBegin;
	int4* segBounds;            // per segment: (min cell, max cell, mean cell, -) of its live rings, updated every step
	void* rho[8];               // [G] double weights or int64 fixed point: this rank's grid, or every rank's (peer-memory mode)
	int nRho, pad1;
	int scatter, pad2;          // SCATTER variant of the kernel: per-warp bins over a wide window (hot species; see push_deposit_body)
	long long bndOffset;        // (uint2*)((double*)rho[r] + bndOffset) = this species' touched-node range per row (encoded maxima)
	// The deposit grids are double-buffered by step parity. While this step's sums go into one parity, the kernel zeroes this
	// species' part of the OTHER one (last step's sums, consumed by last step's solve), ready for the next step's deposits:
	// the populated rows of its grid (no ring, hence no deposit, ever lands above them) and its touched-node ranges.
	double* clearGrid;
	long long clearGridWords;
	double* clearBounds;
	int clearBoundsWords, pad3;
	unsigned long long* lost;   // [0] rings lost since upload, [1] deposits that missed the private window (re-sort trigger)
	// Loss log (nullptr: none): [0] entries written, [1] number of push launches of this species so far = the step tag, [2] CTA
	// ticket of the current launch, [3] unused, then (ring id, step tag) pairs. The reference removes lost rings step by step
	// (swap-with-back, Source/Plasma.cpp:114-118); the host classes replay that order from this log.
	unsigned long long* lossLog;
	const long long* id;        // [cap] ring ids (read for lost rings only)
	long long lossCap;
	double invFixedScale;       // 2^-fixedBits (SCATTER variant, fp64 deposit mode)
};

// Axial cell of a position: bit-exact (int)floor(z / hz) (Source/Plasma.cpp:87, Source/PenningTrap.cpp:328).
// Fast path: q = z * (1/hz), floor through a round-down add of 2^52+2^51. |q - z/hz| < eps/2, so whenever
// frac(q) is at least eps away from 0 and 1 both floors agree (ok = true); otherwise (probability ~2*eps per
// ring) the caller falls back to cell_exact.
__device__ __forceinline__ void cell_fast(double z, const PushArgs& a, int& k, double& kd, bool& ok)
{
	const double q = __dmul_rn(z, a.invHz);
	const double m = __dadd_rd(q, kMagic);
	k = __double2loint(m);
	kd = __dsub_rn(m, kMagic);
	const double f = __dsub_rn(q, kd);
	ok = (f >= a.eps) && (f <= a.epsHi);
}

__device__ __forceinline__ void cell_exact(double z, const PushArgs& a, int& k, double& kd)
{
	kd = floor(__ddiv_rn(z, a.hz));
	k = (int)kd;
}

// Cells and weights of R positions. w = (z - k*hz) / hz is the reference's weightFactor (Source/Plasma.cpp:89-90,
// Source/PenningTrap.cpp:331-332); in FAST arithmetic the division is a reciprocal multiply.
template <int R, bool EXACT>
__device__ __forceinline__ void cells_of(const double (&z)[R], const bool (&live)[R], const PushArgs& a, int (&k)[R], double (&w)[R])
{
	double kd[R];
	bool ok[R];
	bool bad = false;
	if (!EXACT) {
#pragma unroll
		for (int i = 0; i < R; ++i) {
			cell_fast(z[i], a, k[i], kd[i], ok[i]);
			bad |= live[i] && !ok[i];
		}
	}
	else {
#pragma unroll
		for (int i = 0; i < R; ++i) { k[i] = 0; kd[i] = 0.0; ok[i] = false; }
	}
	if (EXACT || bad) {
#pragma unroll
		for (int i = 0; i < R; ++i)
			if (live[i] && (EXACT || !ok[i])) cell_exact(z[i], a, k[i], kd[i]);
	}
#pragma unroll
	for (int i = 0; i < R; ++i) {
		// z < length always holds for a live ring, but z/hz may still round up to Nz; the reference would read
		// node Nz+1 there (its own warning at Source/PenningTrap.cpp:326). Deliberate divergence: stay in the last cell.
		if (k[i] > a.Nz - 1) { k[i] = a.Nz - 1; kd[i] = (double)k[i]; }
		const double dz = __dsub_rn(z[i], __dmul_rn(kd[i], a.hz));
		w[i] = EXACT ? __ddiv_rn(dz, a.hz) : __dmul_rn(dz, a.invHz);
	}
}

// Ring data is touched exactly once per step: evict-first loads / stores keep it from displacing the grids and
// solver constants that the next solve needs from L2.
#ifndef PTP_NO_STREAM
__device__ __forceinline__ double2 ld_ring(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_ring(double2* p, double2 v) { __stcs(p, v); }
#else
__device__ __forceinline__ double2 ld_ring(const double2* p) { return *p; }
__device__ __forceinline__ void st_ring(double2* p, double2 v) { *p = v; }
#endif

// Tiles further ahead than the register prefetch are pulled into L2, so that the register loads of the next tile
// see L2 latency instead of DRAM latency (the kernel is latency-bound: ~32 KB of loads in flight per SM).
#ifndef PTP_L2_PREFETCH_TILES
#define PTP_L2_PREFETCH_TILES 2
#endif

__device__ __forceinline__ void prefetch_l2(const void* p)
{
#ifndef PTP_HOST_EMU
	asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
#else
	(void)p;
#endif
}

template <typename T> __device__ __forceinline__ T warp_sum(T x)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
	return x;
}

// Grid adds of the flush. In peer-memory mode (nRho > 1) several GPUs add to the same node of the same grid over NVLink:
// the operation must be atomic at system scope (red.relaxed.sys) - a .gpu-scope atomic is only defined among the threads of
// one device. On one GPU the narrower scope is kept.
template <typename V> __device__ __forceinline__ void grid_add(V* p, V val, bool sys)
{
	if (sys) atomicAdd_system(p, val);
	else atomicAdd(p, val);
}
__device__ __forceinline__ void grid_max(unsigned int* p, unsigned int val, bool sys)
{
	if (sys) atomicMax_system(p, val);
	else atomicMax(p, val);
}

// SCATTER form of the deposit (hot species): the rings of the 32 lanes of a warp (cell io, packed word = count 1 | weight) go
// into the warp's own bins by plain read-modify-write, so lanes that share a cell must be found and their words added up first.
// The warp SORTS its (cell, lane) keys with a bitonic network of 15 shuffle steps, fetches each ring's packed word to its sorted
// position, adds up the runs of equal cells with a segmented scan - the last lane of every run then holds (rings in the cell |
// sum of their weights) - and those lanes write, one plain read-modify-write per distinct cell, conflict-free by construction.
// The cost hardly depends on the order of the rings.
// A key is 16 bits (cell < 2047 in 11 bits | lane in 5), so the keys of two rings share a register and one network sorts both:
// one shuffle per step, the per-halfword minimum / maximum (VIMNMX.U16x2) does the two compare-exchanges; the network is in the
// form where the lower lane of a pair always keeps the smaller key (the first step of every merge pairs lane l with its mirror
// image in the block of k lanes, l ^ (k - 1), the others with l ^ j). The segmented scan stops at the longest run of equal cells
// among the rings of the tile - the heads of the runs are a vote, the same in every lane, so the test is warp-uniform: two steps
// instead of five when the rings are mixed.
// Measured on the B200 and dropped (profiles/r02_hot_species.txt, 50 M electrons on the 4096 x 1024 grid, ms per step; this
// form: 0.63): one network per ring on 32-bit keys with all five scan steps (0.77; ptxas runs the four networks of a tile one
// after the other); match.any + turns by rank in the group (0.77 for rings in load order, 1.16 after a re-sort, 3.5 for 100 M
// electrons on the default grid: ~500 cycles until the result of match.any arrives with 16 warps of an SM asking, and as many
// turns as the largest group); one tag byte per bin, "try and see" (0.81); one vote per bit of the cell index + gather by the
// first lane of small groups, sort as fall-back (0.86).
#ifdef PTP_HOST_EMU
static inline unsigned int __vminu2(unsigned int a, unsigned int b)
{
	const unsigned int lo = (a & 0xffffu) < (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu), hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
	return (hi << 16) | lo;
}
static inline unsigned int __vmaxu2(unsigned int a, unsigned int b)
{
	const unsigned int lo = (a & 0xffffu) > (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu), hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
	return (hi << 16) | lo;
}
#endif
template <int R>
__device__ __forceinline__ void scatter_group_sorted(const bool (&in)[R], const unsigned int (&io)[R], const unsigned long long (&word)[R], int lane,
	unsigned int (&cellOut)[R], unsigned long long (&sumOut)[R], bool (&writeOut)[R])
{
	static_assert(R % 2 == 0, "rings are sorted in pairs");
	const unsigned int full = 0xffffffffu;
	constexpr unsigned int kNone = 0x7ffu;                                // lanes without a deposit sort to the end
	unsigned int x[R / 2];
#pragma unroll
	for (int h = 0; h < R / 2; ++h)
		x[h] = ((((in[2 * h] ? io[2 * h] : kNone) << 5) | (unsigned int)lane) << 16) | ((in[2 * h + 1] ? io[2 * h + 1] : kNone) << 5) | (unsigned int)lane;
#pragma unroll
	for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			const bool upper = (lane & j) != 0;                            // this lane keeps the larger key of each pair
#pragma unroll
			for (int h = 0; h < R / 2; ++h) {
				const unsigned int y = __shfl_xor_sync(full, x[h], j == (k >> 1) ? k - 1 : j);
				x[h] = upper ? __vmaxu2(x[h], y) : __vminu2(x[h], y);
			}
		}
	}
	unsigned int heads[R];
	int first[R];
	unsigned int longer[5] = { 0u, 0u, 0u, 0u, 0u };                    // [s]: some ring of the tile has a run of more than 2^s equal cells
#pragma unroll
	for (int h = 0; h < R / 2; ++h) {
		const unsigned int prev = __shfl_up_sync(full, x[h], 1);
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			const int i = 2 * h + q;
			const unsigned int key = q == 0 ? x[h] >> 16 : x[h] & 0xffffu, keyPrev = q == 0 ? prev >> 16 : prev & 0xffffu;
			cellOut[i] = key >> 5;
			sumOut[i] = __shfl_sync(full, word[i], (int)(key & 31u));
			if (cellOut[i] == kNone) sumOut[i] = 0ULL;
			// (lanes without a deposit are runs of their own: a tile's padding must not look like one long run)
			heads[i] = __ballot_sync(full, lane == 0 || cellOut[i] != (keyPrev >> 5) || cellOut[i] == kNone);
			first[i] = 31 - __clz((int)(heads[i] & (full >> (31 - lane))));  // first lane of this lane's run
			writeOut[i] = cellOut[i] != kNone && (lane == 31 || ((heads[i] >> (lane + 1)) & 1u));
			unsigned int c = ~heads[i];                                    // lanes that continue a run; c & (c >> 1): two in a row, ...
			longer[0] |= c;
			c &= c >> 1; longer[1] |= c;
			c &= c >> 2; longer[2] |= c;
			c &= c >> 4; longer[3] |= c;
			c &= c >> 8; longer[4] |= c;
		}
	}
#pragma unroll
	for (int st = 0; st < 5; ++st) {
		if (longer[st] == 0u) break;                                     // (warp-uniform: votes)
		const int d = 1 << st;
		unsigned long long up[R];
#pragma unroll
		for (int i = 0; i < R; ++i) up[i] = __shfl_up_sync(full, sumOut[i], d);
#pragma unroll
		for (int i = 0; i < R; ++i)
			if (lane - d >= first[i]) sumOut[i] += up[i];
	}
}

// The deposit of the SCATTER form for the R rings of a thread.
template <int R>
__device__ __forceinline__ void scatter_deposit(unsigned long long* wb, const bool (&in)[R], const unsigned int (&io)[R], const unsigned long long (&word)[R], int lane)
{
	unsigned int cellS[R];
	unsigned long long sumS[R];
	bool writeS[R];
	scatter_group_sorted<R>(in, io, word, lane, cellS, sumS, writeS);
#pragma unroll
	for (int i = 0; i < R; ++i) {
		if (writeS[i]) wb[cellS[i]] += sumS[i];
		__syncwarp();                                                    // the same cell may be written by another lane for the next ring
	}
}

// The kernel body for CTA `bid` of `nb` CTAs working on one species (k_push_deposit: the launch's own grid;
// k_push_deposit_multi: a sub-range of a launch that covers several species).
template <int T, int R, bool PUSH, bool FIXED, bool EXACT, bool SCATTER>
__device__ __forceinline__ void push_deposit_body(const PushArgs& a, const int bid, const int nb)
{
	constexpr int NV = R / 2;
	extern __shared__ __align__(16) unsigned char smem[];
	const int W = a.W, WE = a.WE;
	double2* eTile = reinterpret_cast<double2*>(smem);                       // [WE] (E[k], E[k+1]) of cell kE0+i
	unsigned long long* redC = reinterpret_cast<unsigned long long*>(eTile + WE); // [W]
	unsigned long long* redS = redC + W;                                     // [W] u64 (fixed) or double bits
	unsigned long long* bins = redS + W;                                     // [W][T] packed words / double sums
	unsigned short* cnts = reinterpret_cast<unsigned short*>(bins + (size_t)W * T); // [W][T] fp64 mode only (a thread sees < 4096 rings per segment)
	// SCATTER variant - for species whose rings mix over the whole plasma length within a few steps (electrons on a fine
	// grid: 1.6 cells per step, a bounce every ~35 steps), so that no cell sort survives and thread-private windows of 44
	// cells cannot hold them. Bins are private to a WARP instead ([T/32][W] packed words: 8 B per cell and warp, so the window
	// is ~30 x wider and serves as field window too: W == WE), and the 32 rings a warp handles in one instruction are put into
	// the warp's bins by plain read-modify-write, one lane per distinct cell (scatter_group_sorted / scatter_deposit) - no
	// atomics (shared 64-bit atomics are CAS loops, 32-bit ones cost ~2 cycles per lane), no re-sorts, whatever the order of the
	// rings. The sums are kept in fixed point in BOTH deposit modes (exact integers: the fixed-point mode gets bitwise the sums of the thread-private kernel;
	// the fp64 mode converts when the segment is flushed).
	__shared__ int sKmin, sKmax;
	__shared__ unsigned int sLost, sFar;
	__shared__ long long sKsum[T / 32];
	__shared__ unsigned int sNdep[T / 32];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	unsigned long long* wbins = bins + (size_t)warp * W;                     // SCATTER: this warp's [W] packed words (count:12 | sum:52)
	const int n1 = a.Nz + 1;
	const bool sys = a.nRho > 1;                 // remote grids are among the targets: system-scope atomics
	const double qNaN = __longlong_as_double(0x7ff8000000000000LL);
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();                              // the node field of the last solve, the rings of the last push
	const unsigned long long stepTag = (PUSH && a.lossLog) ? a.lossLog[1] : 0ULL;   // (advanced by the last CTA of this launch to finish)
	if (PUSH) {
		for (long long i = (long long)bid * T + tid; i < a.clearGridWords; i += (long long)nb * T) a.clearGrid[i] = 0.0;
		for (int i = bid * T + tid; i < a.clearBoundsWords; i += nb * T) a.clearBounds[i] = 0.0;
	}

	for (int s = a.ctaSegBegin[bid]; s < a.ctaSegBegin[bid + 1]; ++s) {
		const PtpSegment seg = a.segs[s];
		const int4 bounds = a.segBounds[s];
		if (bounds.x > bounds.y) continue;           // no live ring in this segment (uniform per CTA)
		const long long rowBase = (long long)seg.row * n1;
		// window: centred on the cell range when it fits; otherwise on the mean cell, so that a few far-away rings (the
		// sparse tails of a distribution, a stray fast ring) cannot drag the window off the bulk - they take the slow path
		const int span = bounds.y - bounds.x + 1;
		int k0 = span <= W ? bounds.x - ((W - span) >> 1) : bounds.z - (W >> 1);
		k0 = max(0, min(k0, a.Nz - W));
		// the field window is wider than the deposit window (16 B per cell instead of 10 B per cell and thread): rings that
		// have drifted out of the deposit window still gather from shared memory - a global load there would stall the warp
		int kE0 = max(0, min(k0 + (W >> 1) - (WE >> 1), a.Nz - WE));
		// SCATTER: one window for field and deposit, and only the part of it that this step can reach is loaded, cleared and
		// reduced - the cell range of the segment's rings and a margin on either side for the step's drift; a ring that flies
		// further in one step takes the global path
		int Wuse = W;
		if constexpr (SCATTER) {
			constexpr int kScatterMargin = 48;
			k0 = max(0, bounds.x - kScatterMargin);
			Wuse = min(a.Nz, bounds.y + kScatterMargin + 1) - k0;
			if (Wuse > W) { Wuse = W; k0 = max(0, min(bounds.z - (W >> 1), a.Nz - W)); }
			kE0 = k0;
		}
		const unsigned int weUse = SCATTER ? (unsigned int)Wuse : (unsigned int)WE;
		// The segment's first global loads go out before anything else - the first tile of rings and this thread's entry of
		// the field window (WE <= 256 <= T) - so that their latency overlaps the clearing of the bins.
		const double2* z2 = reinterpret_cast<const double2*>(a.z);
		const double2* v2 = reinterpret_cast<const double2*>(a.v);
		double2* z2w = reinterpret_cast<double2*>(a.z);
		double2* v2w = reinterpret_cast<double2*>(a.v);
		const long long tile = (long long)R * T;
		double2 zzN[NV], vvN[NV];
		{
			const long long p0 = (seg.begin >> 1) + tid;
#pragma unroll
			for (int j = 0; j < NV; ++j) zzN[j] = ld_ring(z2 + p0 + (long long)j * T);
			if (PUSH) {
#pragma unroll
				for (int j = 0; j < NV; ++j) vvN[j] = ld_ring(v2 + p0 + (long long)j * T);
			}
		}
		double eL0 = 0.0, eR0 = 0.0;
		if (!SCATTER && PUSH && tid < WE) {
			const int node = kE0 + tid;
			if (node <= a.Nz) eL0 = a.eNodes[rowBase + node];
			if (node + 1 <= a.Nz) eR0 = a.eNodes[rowBase + node + 1];
		}
		if constexpr (!SCATTER) {
			for (int i = 0; i < W; ++i) {
				bins[(size_t)i * T + tid] = 0ULL;
				if (!FIXED) cnts[(size_t)i * T + tid] = 0;
			}
			if (PUSH && tid < WE) eTile[tid] = make_double2(eL0, eR0);
		}
		else {
			for (int c = lane; c < Wuse; c += 32) wbins[c] = 0ULL;
			if (PUSH)
				for (int i = tid; i < Wuse; i += T) {                    // (k0 + i <= Nz - 1: both nodes of the cell exist)
					const long long node = rowBase + k0 + i;
					eTile[i] = make_double2(a.eNodes[node], a.eNodes[node + 1]);
				}
		}
		if (tid == 0) { sKmin = INT_MAX; sKmax = INT_MIN; sLost = 0u; sFar = 0u; }
		__syncthreads();

		int kMin = INT_MAX, kMax = INT_MIN;
		unsigned int lost = 0, nDep = 0, nFar = 0;
		long long kSum = 0;

		for (long long t0 = seg.begin; t0 < seg.end; t0 += tile) {
			const long long p0 = (t0 >> 1) + tid;
			double z[R], v[R];
#pragma unroll
			for (int j = 0; j < NV; ++j) {
				z[2 * j] = zzN[j].x; z[2 * j + 1] = zzN[j].y;
				if (PUSH) { v[2 * j] = vvN[j].x; v[2 * j + 1] = vvN[j].y; }
			}
			if (PTP_L2_PREFETCH_TILES > 0 && t0 + PTP_L2_PREFETCH_TILES * tile < seg.end) {
				const long long pf = p0 + PTP_L2_PREFETCH_TILES * (tile >> 1);
#pragma unroll
				for (int j = 0; j < NV; ++j) {
					prefetch_l2(z2 + pf + (long long)j * T);
					if (PUSH) prefetch_l2(v2 + pf + (long long)j * T);
				}
			}
			if (t0 + tile < seg.end) {               // next tile's loads are in flight while this one is processed
				const long long pn = p0 + (tile >> 1);
#pragma unroll
				for (int j = 0; j < NV; ++j) zzN[j] = ld_ring(z2 + pn + (long long)j * T);
				if (PUSH) {
#pragma unroll
					for (int j = 0; j < NV; ++j) vvN[j] = ld_ring(v2 + pn + (long long)j * T);
				}
			}

			bool live[R];
			bool liveIn[NV];
#pragma unroll
			for (int i = 0; i < R; ++i) live[i] = (z[i] == z[i]);   // NaN = empty slot / ring lost earlier
#pragma unroll
			for (int j = 0; j < NV; ++j) liveIn[j] = live[2 * j] || live[2 * j + 1];

			int k[R];
			double w[R];
			if (PUSH) {
				// ---- gather + kick + drift: Plasma::moveRings body (Source/Plasma.cpp:105-118) -------------
				cells_of<R, EXACT>(z, live, a, k, w);
				double eL[R], eR[R];
				bool far = false;
				unsigned int goneMask = 0;
#pragma unroll
				for (int i = 0; i < R; ++i) {
					const unsigned int io = (unsigned int)(k[i] - kE0);
					const bool in = io < weUse;
					far |= live[i] && !in;
					const double2 e = eTile[in ? io : 0u];
					eL[i] = e.x; eR[i] = e.y;
				}
				if (far) {
#pragma unroll
					for (int i = 0; i < R; ++i)
						if (live[i] && (unsigned int)(k[i] - kE0) >= weUse) {
							eL[i] = a.eNodes[rowBase + k[i]];
							eR[i] = a.eNodes[rowBase + k[i] + 1];
						}
				}
#pragma unroll
				for (int i = 0; i < R; ++i) {
					// (1 - w) * fieldLeft + w * fieldRight            Source/PenningTrap.cpp:333
					const double e = __dadd_rn(__dmul_rn(__dsub_rn(1.0, w[i]), eL[i]), __dmul_rn(w[i], eR[i]));
					// deltaT * E * charge / mass + speed              Source/Plasma.cpp:105
					double kick = __dmul_rn(__dmul_rn(a.dt, e), a.charge);
					kick = EXACT ? __ddiv_rn(kick, a.mass) : __dmul_rn(kick, a.invMass);
					const double vN = __dadd_rn(kick, v[i]);
					// deltaT * vNew + z                               Source/Plasma.cpp:106
					const double zN = __dadd_rn(__dmul_rn(a.dt, vN), z[i]);
					// zNew < length && zNew > 0 keeps the ring (Source/Plasma.cpp:108; NaN -> removed, as in the reference);
					// a removed ring becomes a NaN tombstone instead of swap-with-back + pop (:116-117)
					const bool keep = live[i] && (zN < a.length) && (zN > 0.0);
					lost += (live[i] && !keep) ? 1u : 0u;
					goneMask |= (live[i] && !keep) ? (1u << i) : 0u;
					z[i] = keep ? zN : qNaN;
					v[i] = keep ? vN : v[i];
					live[i] = keep;
				}
				if (goneMask && a.lossLog) {                             // rare: which ring left the trap, and in which step
#pragma unroll
					for (int i = 0; i < R; ++i)
						if ((goneMask >> i) & 1u) {
							const long long slot = 2 * (p0 + (long long)(i >> 1) * T) + (i & 1);
							const unsigned long long pos = atomicAdd(a.lossLog, 1ULL);
							if ((long long)pos < a.lossCap) {
								a.lossLog[4 + 2 * pos] = (unsigned long long)a.id[slot];
								a.lossLog[5 + 2 * pos] = stepTag;
							}
						}
				}
			}

			// ---- deposit at the (new) position: Plasma::updateRHS body (Source/Plasma.cpp:86-92) -----------
			cells_of<R, EXACT>(z, live, a, k, w);
			if constexpr (SCATTER) {
				// the rings go back to memory before the deposit: z and v are dead from here on, which leaves registers for the
				// sorting networks of the deposit to run side by side
				if (PUSH) {
#pragma unroll
					for (int j = 0; j < NV; ++j)
						if (liveIn[j]) {
							st_ring(z2w + p0 + (long long)j * T, make_double2(z[2 * j], z[2 * j + 1]));
							st_ring(v2w + p0 + (long long)j * T, make_double2(v[2 * j], v[2 * j + 1]));
						}
				}
			}
			bool farD = false;
			if constexpr (!SCATTER) {
#pragma unroll
				for (int i = 0; i < R; ++i) {
					const unsigned int io = (unsigned int)(k[i] - k0);
					const bool in = live[i] && io < (unsigned int)W;
					farD |= live[i] && io >= (unsigned int)W;
					nFar += (live[i] && io >= (unsigned int)W) ? 1u : 0u;
					if (live[i]) { kMin = min(kMin, k[i]); kMax = max(kMax, k[i]); kSum += k[i]; ++nDep; }
					if (in) {
						if (FIXED) {
							// round(w * 2^F) lands in the mantissa of (2^52 + x); subtracting the bias leaves (1 << 52) | x
							const double t = __fma_rn(w[i], a.fixedScale, 4503599627370496.0);
							bins[(size_t)io * T + tid] += (unsigned long long)__double_as_longlong(t) - kPackBias;
						}
						else {
							double* b = reinterpret_cast<double*>(bins) + (size_t)io * T + tid;
							*b = __dadd_rn(*b, w[i]);
							cnts[(size_t)io * T + tid] += 1;
						}
					}
				}
			}
			else {
				// SCATTER: the warp's 32 rings of this stage into the warp's bins
				unsigned int ioS[R];
				unsigned long long wordS[R];
				bool inS[R];
#pragma unroll
				for (int i = 0; i < R; ++i) {
					ioS[i] = (unsigned int)(k[i] - k0);
					inS[i] = live[i] && ioS[i] < (unsigned int)Wuse;
					farD |= live[i] && ioS[i] >= (unsigned int)Wuse;
					nFar += (live[i] && ioS[i] >= (unsigned int)Wuse) ? 1u : 0u;
					if (live[i]) { kMin = min(kMin, k[i]); kMax = max(kMax, k[i]); kSum += k[i]; ++nDep; }
					// the packed word of the thread-private path: (1 << 52) | round(w * 2^F)
					const double t = __fma_rn(w[i], a.fixedScale, 4503599627370496.0);
					wordS[i] = (unsigned long long)__double_as_longlong(t) - kPackBias;
				}
				scatter_deposit<R>(wbins, inS, ioS, wordS, lane);
			}
			if (farD) {
				// outside the private window: straight to the global grid (REDG.E.ADD.F64 / .64)
#pragma unroll
				for (int i = 0; i < R; ++i)
					if (live[i] && (unsigned int)(k[i] - k0) >= (unsigned int)Wuse) {
						// (the row's touched node range is widened to the segment's whole cell range in the epilogue)
						if (FIXED) {
							const double t = __fma_rn(w[i], a.fixedScale, 4503599627370496.0);
							const unsigned long long wq = (unsigned long long)__double_as_longlong(t) & kSumMask;
							for (int pr = 0; pr < a.nRho; ++pr) {
								unsigned long long* g = reinterpret_cast<unsigned long long*>(a.rho[pr]) + rowBase + k[i];
								grid_add(g, (1ULL << a.fixedBits) - wq, sys);
								grid_add(g + 1, wq, sys);
							}
						}
						else {
							for (int pr = 0; pr < a.nRho; ++pr) {
								double* g = reinterpret_cast<double*>(a.rho[pr]) + rowBase + k[i];
								grid_add(g, __dsub_rn(1.0, w[i]), sys);
								grid_add(g + 1, w[i], sys);
							}
						}
					}
			}
			if (PUSH && !SCATTER) {
#pragma unroll
				for (int j = 0; j < NV; ++j)
					if (liveIn[j]) {
						st_ring(z2w + p0 + (long long)j * T, make_double2(z[2 * j], z[2 * j + 1]));
						st_ring(v2w + p0 + (long long)j * T, make_double2(v[2 * j], v[2 * j + 1]));
					}
			}
		}

		// the CTA's next segment starts with cold loads: pull its first tiles into L2 while this segment's epilogue runs
		if (PTP_L2_PREFETCH_TILES > 0 && s + 1 < a.ctaSegBegin[bid + 1]) {
			const PtpSegment nx = a.segs[s + 1];
			for (int u = 0; u < PTP_L2_PREFETCH_TILES && nx.begin + u * tile < nx.end; ++u) {
				const long long pf = ((nx.begin + u * tile) >> 1) + tid;
#pragma unroll
				for (int j = 0; j < NV; ++j) {
					prefetch_l2(z2 + pf + (long long)j * T);
					if (PUSH) prefetch_l2(v2 + pf + (long long)j * T);
				}
			}
		}

		// segment epilogue: cell range, column reduction, flush
		for (int o = 16; o > 0; o >>= 1) {
			kMin = min(kMin, __shfl_xor_sync(0xffffffffu, kMin, o));
			kMax = max(kMax, __shfl_xor_sync(0xffffffffu, kMax, o));
		}
		lost = warp_sum(lost);
		nFar = warp_sum(nFar);
		kSum = warp_sum(kSum);
		nDep = warp_sum(nDep);
		if (lane == 0) {
			if (kMin <= kMax) { atomicMin(&sKmin, kMin); atomicMax(&sKmax, kMax); }
			if (lost) atomicAdd(&sLost, lost);
			if (nFar) atomicAdd(&sFar, nFar);
			sKsum[warp] = kSum;
			sNdep[warp] = nDep;
		}
		__syncthreads();
		const int gMin = sKmin, gMax = sKmax;
		if (gMin <= gMax) {
			const int lo = max(gMin, k0) - k0, hi = min(gMax, k0 + Wuse - 1) - k0;
			if constexpr (SCATTER) {
				for (int b = lo + tid; b <= hi; b += T) {               // one thread per cell: the T/32 warps' words
					unsigned long long c = 0, sm = 0;
					for (int wq = 0; wq < T / 32; ++wq) {
						const unsigned long long word = bins[(size_t)wq * W + b];
						c += word >> 52;
						sm += word & kSumMask;
					}
					redC[b] = c;
					// fp64 mode: the exact integer sum (< 2^16 rings x 2^F per warp, 2^20 x 2^F per CTA and segment) back to units of one ring
					redS[b] = FIXED ? sm : (unsigned long long)__double_as_longlong(__dmul_rn((double)sm, a.invFixedScale));
				}
			}
			else
			for (int b = lo + warp; b <= hi; b += T / 32) {
				if (FIXED) {
					unsigned long long c = 0, sm = 0;
					for (int t = lane; t < T; t += 32) {
						const unsigned long long word = bins[(size_t)b * T + t];
						c += word >> 52;
						sm += word & kSumMask;
					}
					c = warp_sum(c);
					sm = warp_sum(sm);
					if (lane == 0) { redC[b] = c; redS[b] = sm; }
				}
				else {
					unsigned int c = 0;
					double sm = 0.0;
					for (int t = lane; t < T; t += 32) {
						c += cnts[(size_t)b * T + t];
						sm += reinterpret_cast<const double*>(bins)[(size_t)b * T + t];
					}
					c = warp_sum(c);
					sm = warp_sum(sm);
					if (lane == 0) { redC[b] = c; redS[b] = (unsigned long long)__double_as_longlong(sm); }
				}
			}
			__syncthreads();
			for (int i = lo + tid; i <= hi + 1; i += T) {
				if (FIXED) {
					unsigned long long val = 0;
					if (i <= hi) val += (redC[i] << a.fixedBits) - redS[i];
					if (i > lo) val += redS[i - 1];
					if (val)
						for (int pr = 0; pr < a.nRho; ++pr) grid_add(reinterpret_cast<unsigned long long*>(a.rho[pr]) + rowBase + k0 + i, val, sys);
				}
				else {
					double val = 0.0;
					if (i <= hi) val = (double)redC[i] - __longlong_as_double((long long)redS[i]);
					if (i > lo) val += __longlong_as_double((long long)redS[i - 1]);
					if (val != 0.0)
						for (int pr = 0; pr < a.nRho; ++pr) grid_add(reinterpret_cast<double*>(a.rho[pr]) + rowBase + k0 + i, val, sys);
				}
			}
			// nodes gMin .. gMax+1 of this row were touched - by the flush above or, for rings outside the window, by their own
			// global adds: keep the row's range for the solver's forward transform, one pair of atomics per segment
			// (both ends stored as maxima so that a memset(0) resets them: Nz+2-kmin and kmax+2)
			if (tid == 0)
				for (int pr = 0; pr < a.nRho; ++pr) {
					unsigned int* bd = reinterpret_cast<unsigned int*>(reinterpret_cast<double*>(a.rho[pr]) + a.bndOffset) + 2 * seg.row;
					grid_max(bd, (unsigned int)(a.Nz + 2 - gMin), sys);
					grid_max(bd + 1, (unsigned int)(gMax + 2), sys);
				}
		}
		if (a.nRho > 1 && a.pad1) {                      // remote adds performed before the grid can be declared complete:
			__syncthreads();                             // one system fence per CTA (cumulative over the CTA's flush)
			if (tid == 0) __threadfence_system();
		}
		if (PUSH && tid == 0) {
			long long ks = 0;
			unsigned int nd = 0;
			for (int w = 0; w < T / 32; ++w) { ks += sKsum[w]; nd += sNdep[w]; }
			a.segBounds[s] = make_int4(gMin, gMax, nd ? (int)(ks / nd) : 0, 0);   // next step's window (this CTA owns the segment)
			if (sLost) atomicAdd(a.lost, (unsigned long long)sLost);
			if (sFar) atomicAdd(a.lost + 1, (unsigned long long)sFar);
		}
		__syncthreads();
	}
	if (PUSH && a.lossLog && tid == 0) {
		// every CTA of the launch has read the step tag before the last one to finish advances it
		__threadfence();
		const unsigned long long ticket = atomicAdd(a.lossLog + 2, 1ULL);
		if (ticket == (unsigned long long)nb - 1) {
			a.lossLog[2] = 0ULL;
			a.lossLog[1] = stepTag + 1;
		}
	}
}

template <int T, int R, bool PUSH, bool FIXED, bool EXACT, bool SCATTER = false>
__global__ void __launch_bounds__(T, 1) k_push_deposit(const PushArgs a)
{
	push_deposit_body<T, R, PUSH, FIXED, EXACT, SCATTER>(a, (int)blockIdx.x, (int)gridDim.x);
}

// [emu-end]

// All species of a step in ONE launch: CTAs [ctaBegin[s], ctaBegin[s + 1]) work on species s. The species of the reference's
// default configuration hold a few thousand rings each, the co-trapped 10 M-ring case 5 M each: one launch per species leaves
// most SMs idle for most of a short kernel, and every launch pays its own ramp and tail. (The species are pushed with the
// same pre-step field in the reference too, Source/PenningTrap.cpp:354-357, so their order is immaterial.)
constexpr int PTP_MULTI_MAX = 4;
struct PushArgsMulti {
	int n, ctaBegin[PTP_MULTI_MAX + 1], pad[2];
	PushArgs sp[PTP_MULTI_MAX];
};

template <int T, int R, bool FIXED, bool EXACT>
__global__ void __launch_bounds__(T, 1) k_push_deposit_multi(const __grid_constant__ PushArgsMulti m)
{
	int si = 0;
	while (si + 1 < m.n && (int)blockIdx.x >= m.ctaBegin[si + 1]) ++si;
	push_deposit_body<T, R, true, FIXED, EXACT, false>(m.sp[si], (int)blockIdx.x - m.ctaBegin[si], m.ctaBegin[si + 1] - m.ctaBegin[si]);
}

// Axial cell range and live count of every tile (window planning at upload / after a sort) and validation of the
// positions. One CTA per tile; tiles[] lists them as one-tile segments.
__global__ void __launch_bounds__(256) k_tile_bounds(const PushArgs a, const PtpSegment* __restrict__ tiles, int nTiles,
	int2* __restrict__ tileBounds, unsigned long long* nLive, int* invalid)
{
	__shared__ int sKmin, sKmax;
	for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
		const PtpSegment seg = tiles[tIdx];
		if (threadIdx.x == 0) { sKmin = INT_MAX; sKmax = INT_MIN; }
		__syncthreads();
		int kMin = INT_MAX, kMax = INT_MIN;
		unsigned int live = 0;
		for (long long i = seg.begin + threadIdx.x; i < seg.end; i += blockDim.x) {
			const double z = a.z[i];
			if (!(z == z)) continue;
			if (!(z > 0.0 && z < a.length)) { *invalid = 1; continue; }
			int k;
			double kd;
			bool ok;
			cell_fast(z, a, k, kd, ok);
			if (!ok) cell_exact(z, a, k, kd);
			k = min(k, a.Nz - 1);
			kMin = min(kMin, k);
			kMax = max(kMax, k);
			++live;
		}
		for (int o = 16; o > 0; o >>= 1) {
			kMin = min(kMin, __shfl_xor_sync(0xffffffffu, kMin, o));
			kMax = max(kMax, __shfl_xor_sync(0xffffffffu, kMax, o));
		}
		live = warp_sum(live);
		if ((threadIdx.x & 31) == 0 && live) {
			atomicMin(&sKmin, kMin);
			atomicMax(&sKmax, kMax);
			atomicAdd(nLive, (unsigned long long)live);
		}
		__syncthreads();
		if (threadIdx.x == 0) tileBounds[tIdx] = make_int2(sKmin, sKmax);
		__syncthreads();
	}
}

template <int T, int R, bool PUSH> struct Launcher {
	template <bool FIXED, bool EXACT> static cudaError_t go(const PushArgs& a, int grid, size_t smem, cudaStream_t st, bool pdl)
	{
		if (a.scatter) {                                             // per-warp bins (hot species); default tuning only
			auto kernS = k_push_deposit<512, 4, PUSH, FIXED, EXACT, true>;
			cudaError_t eS = cudaFuncSetAttribute(kernS, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (eS != cudaSuccess) return eS;
			return ptp_launch(kernS, dim3(grid), dim3(512), smem, st, pdl, a);
		}
		auto kern = k_push_deposit<T, R, PUSH, FIXED, EXACT>;
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
		return ptp_launch(kern, dim3(grid), dim3(T), smem, st, pdl, a);
	}
	static cudaError_t dispatch(bool fixed, bool exact, const PushArgs& a, int grid, size_t smem, cudaStream_t st, bool pdl)
	{
		if (fixed) return exact ? go<true, true>(a, grid, smem, st, pdl) : go<true, false>(a, grid, smem, st, pdl);
		return exact ? go<false, true>(a, grid, smem, st, pdl) : go<false, false>(a, grid, smem, st, pdl);
	}
};

template <int T, int R> cudaError_t launch_tr(bool push, bool fixed, bool exact, const PushArgs& a, int grid, size_t smem, cudaStream_t st, bool pdl)
{
	return push ? Launcher<T, R, true>::dispatch(fixed, exact, a, grid, smem, st, pdl)
	            : Launcher<T, R, false>::dispatch(fixed, exact, a, grid, smem, st, pdl);
}

PushArgs make_args(ptp_trap* t, ptp_plasma* p, double dt)
{
	PushArgs a{};
	a.Nz = t->Nz;
	a.W = t->window < t->Nz ? t->window : t->Nz;
	a.WE = ptp_push_field_window(t);
	a.fixedBits = t->fixedBits;
	a.scatter = p->scatter ? 1 : 0;
	if (p->scatter) {
		a.W = a.WE = ptp_push_scatter_window(t);
		if (t->depositMode != PTP_DEPOSIT_FIXED64) a.fixedBits = 40;     // the warps' bins hold fixed-point sums in fp64 mode too
	}
	a.hz = t->hz;
	a.invHz = 1.0 / t->hz;
	a.eps = (t->Nz + 2) * 1e-15;
	a.epsHi = 1.0 - a.eps;
	a.length = t->length;
	a.dt = dt;
	a.charge = p->charge;
	a.mass = p->mass;
	a.invMass = 1.0 / p->mass;
	a.fixedScale = (double)(1ULL << a.fixedBits);
	a.invFixedScale = 1.0 / a.fixedScale;
	a.eNodes = t->eNodes;
	a.z = p->z;
	a.v = p->v;
	a.segs = p->dSegs;
	a.ctaSegBegin = p->dCtaSegBegin;
	a.segBounds = p->dSegBounds;
	a.nRho = 1;
	{
		// Peer-memory mode: the remote adds of this kernel are complete at kernel end (stream order is a system-scope
		// release), and the flag barrier is a separate kernel behind it, so no in-kernel fence is needed. PTP_PEER_FENCE=1
		// adds one system fence per CTA after the flush anyway (measured cost: ~12 us per step at 4 GPUs).
		static const int fence = std::getenv("PTP_PEER_FENCE") ? 1 : 0;
		a.pad1 = fence;
	}
	a.rho[0] = t->rhoAll + (size_t)p->index * t->G;
	a.bndOffset = (long long)((size_t)t->capS * t->G - (size_t)p->index * t->G + (size_t)p->index * t->Nr);
	a.lost = p->dLost;
	a.lossLog = p->dLossLog;
	a.id = p->id;
	a.lossCap = p->lossCap;
	return a;
}

} // namespace

// Cells of the field window: as many as fit beside the deposit bins, at most 256 and never fewer than the deposit window.
int ptp_push_field_window(const ptp_trap* t)
{
	const size_t w = (size_t)(t->window < t->Nz ? t->window : t->Nz);
	const size_t perBin = t->depositMode == PTP_DEPOSIT_FIXED64 ? 8 : 10;
	const size_t fixedPart = w * 16 + w * (size_t)t->threads * perBin + 1024;   // reduction rows + bins + static shared memory
	size_t we = t->smemMax > fixedPart ? (t->smemMax - fixedPart) / 16 : 0;
	if (we > 256) we = 256;
	if (we > (size_t)t->Nz) we = (size_t)t->Nz;
	if (we < w) we = w;
	return (int)we;
}

// SCATTER variant (hot species): cells of the one window that serves as field window and deposit window - 16 B of field,
// 16 B of reduction rows and one 8-byte word per warp and cell (512 threads).
constexpr size_t kScatterBytesPerCell = 16 + 16 + 8 * (512 / 32);
int ptp_push_scatter_window(const ptp_trap* t)
{
	const size_t perCell = kScatterBytesPerCell;
	size_t w = t->smemMax > 2048 ? (t->smemMax - 2048) / perCell : 0;
	if (w > (size_t)t->Nz) w = (size_t)t->Nz;
	if (w > 2047) w = 2047;
	return (int)w;
}

bool ptp_push_scatter_usable(const ptp_trap* t)
{
	const int w = ptp_push_scatter_window(t);
	return t->threads == 512 && t->ringsPerThread == 4 && (w >= 128 || w == t->Nz);
}

size_t ptp_push_scatter_smem_bytes(const ptp_trap* t) { return (size_t)ptp_push_scatter_window(t) * kScatterBytesPerCell; }

size_t ptp_push_smem_bytes(const ptp_trap* t, int threads, int window)
{
	size_t w = (size_t)(window < t->Nz ? window : t->Nz);
	size_t perBin = t->depositMode == PTP_DEPOSIT_FIXED64 ? 8 : 10;
	return (size_t)ptp_push_field_window(t) * 16 + w * 16 + w * (size_t)threads * perBin;
}

int ptp_push_configure(ptp_trap* t)
{
	if (t->threads != 256 && t->threads != 512) { ptp_set_error("threads per CTA must be 256 or 512"); return PTP_EINVAL; }
	if (t->ringsPerThread != 4 && t->ringsPerThread != 8) { ptp_set_error("rings per thread must be 4 or 8"); return PTP_EINVAL; }
	if (t->window < 4) { ptp_set_error("window must be at least 4 cells"); return PTP_EINVAL; }
	if (ptp_push_smem_bytes(t, t->threads, t->window) > t->smemMax) {
		ptp_set_error("threads x window does not fit in shared memory");
		return PTP_EINVAL;
	}
	return PTP_OK;
}

// Launch K1 (push = true) or K2 (push = false) for one species on the trap's stream.
int ptp_push_launch(ptp_trap* t, ptp_plasma* p, double dt, bool push)
{
	if (p->cap == 0 || p->nCta == 0) return PTP_OK;
	PushArgs a = make_args(t, p, dt);
	if (push && ptp_peer_fused(t)) ptp_peer_targets(t, t->rhoParity, (size_t)p->index * t->G, a.rho, &a.nRho);
	if (push) {
		double* other = t->rhoStore + (size_t)(t->rhoParity ^ 1) * t->spanDoubles;
		a.clearGrid = other + (size_t)p->index * t->G;
		a.clearGridWords = (long long)std::min(t->rowExtent, t->Nr) * (t->Nz + 1);
		a.clearBounds = other + (size_t)t->capS * t->G + (size_t)p->index * t->Nr;
		a.clearBoundsWords = t->Nr;
	}
	const size_t smem = p->scatter ? ptp_push_scatter_smem_bytes(t) : ptp_push_smem_bytes(t, t->threads, t->window);
	const bool fixed = t->depositMode == PTP_DEPOSIT_FIXED64, exact = t->arithMode == PTP_ARITH_EXACT;
	const bool pdl = t->usePdl;
	cudaError_t e;
	if (t->threads == 512)
		e = t->ringsPerThread == 8 ? launch_tr<512, 8>(push, fixed, exact, a, p->nCta, smem, t->stream, pdl)
		                           : launch_tr<512, 4>(push, fixed, exact, a, p->nCta, smem, t->stream, pdl);
	else
		e = t->ringsPerThread == 8 ? launch_tr<256, 8>(push, fixed, exact, a, p->nCta, smem, t->stream, pdl)
		                           : launch_tr<256, 4>(push, fixed, exact, a, p->nCta, smem, t->stream, pdl);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_push_deposit launch", __FILE__, __LINE__);
	t->lastLaunches++;
	return PTP_OK;
}

// K1 for several species in one launch (same tuning and modes for all; at most PTP_MULTI_MAX species per launch).
int ptp_push_launch_multi(ptp_trap* t, ptp_plasma* const* ps, int n, double dt)
{
	if (n < 1 || n > PTP_MULTI_MAX) { ptp_set_error("ptp_push_launch_multi: bad species count"); return PTP_EINVAL; }
	PushArgsMulti m{};
	m.n = n;
	int total = 0;
	for (int i = 0; i < n; ++i) {
		ptp_plasma* p = ps[i];
		PushArgs a = make_args(t, p, dt);
		if (ptp_peer_fused(t)) ptp_peer_targets(t, t->rhoParity, (size_t)p->index * t->G, a.rho, &a.nRho);
		double* other = t->rhoStore + (size_t)(t->rhoParity ^ 1) * t->spanDoubles;
		a.clearGrid = other + (size_t)p->index * t->G;
		a.clearGridWords = (long long)std::min(t->rowExtent, t->Nr) * (t->Nz + 1);
		a.clearBounds = other + (size_t)t->capS * t->G + (size_t)p->index * t->Nr;
		a.clearBoundsWords = t->Nr;
		m.sp[i] = a;
		m.ctaBegin[i] = total;
		total += p->nCta;
	}
	for (int i = n; i <= PTP_MULTI_MAX; ++i) m.ctaBegin[i] = total;
	const size_t smem = ptp_push_smem_bytes(t, t->threads, t->window);
	const bool fixed = t->depositMode == PTP_DEPOSIT_FIXED64, exact = t->arithMode == PTP_ARITH_EXACT;
	auto go = [&](auto kern) -> cudaError_t {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
		return ptp_launch(kern, dim3(total), dim3(512), smem, t->stream, t->usePdl, m);
	};
	cudaError_t e;
	if (fixed) e = exact ? go(k_push_deposit_multi<512, 4, true, true>) : go(k_push_deposit_multi<512, 4, true, false>);
	else e = exact ? go(k_push_deposit_multi<512, 4, false, true>) : go(k_push_deposit_multi<512, 4, false, false>);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_push_deposit_multi launch", __FILE__, __LINE__);
	t->lastLaunches++;
	return PTP_OK;
}

// Cell range of every tile in `tiles` (host list) -> tileBounds (host), live ring count; validates 0 < z < length.
int ptp_tile_bounds(ptp_trap* t, ptp_plasma* p, const std::vector<PtpSegment>& tiles, std::vector<int2>& tileBounds, int64_t* nLive)
{
	tileBounds.assign(tiles.size(), make_int2(INT_MAX, INT_MIN));
	*nLive = 0;
	if (tiles.empty()) return PTP_OK;
	const PushArgs a = make_args(t, p, 0.0);
	const int n = (int)tiles.size();
	const size_t need = (size_t)n * (sizeof(PtpSegment) + sizeof(int2)) + 64;
	if (p->planScratchBytes < need) {
		cudaFree(p->planScratch);
		p->planScratch = nullptr; p->planScratchBytes = 0;
		PTP_CUDA(cudaMalloc(&p->planScratch, need + need / 2));
		p->planScratchBytes = need + need / 2;
	}
	unsigned long long* dLive = static_cast<unsigned long long*>(p->planScratch);   // counter + two flags in the first 64 bytes
	int* dInvalid = reinterpret_cast<int*>(dLive + 1);
	PtpSegment* dTiles = reinterpret_cast<PtpSegment*>(static_cast<char*>(p->planScratch) + 64);
	int2* dB = reinterpret_cast<int2*>(dTiles + n);
	PTP_CUDA(cudaMemcpyAsync(dTiles, tiles.data(), (size_t)n * sizeof(PtpSegment), cudaMemcpyHostToDevice, t->stream));
	PTP_CUDA(cudaMemsetAsync(dLive, 0, sizeof(unsigned long long) + 2 * sizeof(int), t->stream));
	const int grid = n < t->smCount * 32 ? n : t->smCount * 32;
	k_tile_bounds<<<grid, 256, 0, t->stream>>>(a, dTiles, n, dB, dLive, dInvalid);
	cudaError_t e = cudaGetLastError();
	unsigned long long live = 0;
	int invalid = 0;
	if (e == cudaSuccess) e = cudaMemcpyAsync(tileBounds.data(), dB, (size_t)n * sizeof(int2), cudaMemcpyDeviceToHost, t->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(&live, dLive, sizeof(live), cudaMemcpyDeviceToHost, t->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(&invalid, dInvalid, sizeof(int), cudaMemcpyDeviceToHost, t->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_tile_bounds", __FILE__, __LINE__);
	t->lastLaunches++;
	if (invalid) { ptp_set_error("ring position outside (0, trap length)"); return PTP_EINVAL; }
	*nLive = (int64_t)live;
	return PTP_OK;
}

// Segment tables and windows are planned together (ptp_particles.cu).
int ptp_bounds_launch(ptp_trap* t, ptp_plasma* p) { return ptp_build_segments(t, p); }
