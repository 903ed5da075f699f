// Ring storage: row-bucketed SoA, segment tables for the push kernel, download/compaction, the integer
// keys of the step, K5 (per-row counting sort by axial cell + compaction) and the diagnostics reductions.
#include "ptp_internal.h"

#include <limits.h>

#include <algorithm>
#include <cmath>
#include <thread>

namespace {

__device__ __forceinline__ int exact_cell(double z, double hz, int Nz)
{
	int k = (int)floor(__ddiv_rn(z, hz));                 // Source/Plasma.cpp:87
	return k > Nz - 1 ? Nz - 1 : k;
}

__global__ void k_cell_index(const double* __restrict__ z, long long cap, double hz, int Nz, int* __restrict__ k)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cap) return;
	const double zz = z[i];
	k[i] = (zz == zz) ? exact_cell(zz, hz, Nz) : -1;
}

__global__ void k_iota_rows(long long* __restrict__ id, const long long* __restrict__ rowOff, const long long* __restrict__ rowSrc, int Nr)
{
	const int r = blockIdx.y;
	if (r >= Nr) return;
	const long long n = rowSrc[r + 1] - rowSrc[r];
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
		id[rowOff[r] + i] = rowSrc[r] + i;
}

// Plasma::getPotentialEnergy (Source/Plasma.cpp:244-252) with getTotalPhi(int,double) (Source/PenningTrap.cpp:335-351).
__global__ void k_potential_energy(const double* __restrict__ z, const PtpSegment* __restrict__ segs, int nSegs,
	const double* __restrict__ phiTrap, const double* __restrict__ phiSelf, int nS, long long G, int Nz, double hz,
	double chargeMacro, double* __restrict__ out)
{
	double acc = 0.0;
	for (int s = blockIdx.x; s < nSegs; s += gridDim.x) {
		const PtpSegment seg = segs[s];
		const long long rowBase = (long long)seg.row * (Nz + 1);
		const double q = seg.row == 0 ? chargeMacro : (double)(seg.row * 8) * chargeMacro;
		for (long long i = seg.begin + threadIdx.x; i < seg.end; i += blockDim.x) {
			const double zz = z[i];
			if (!(zz == zz)) continue;
			const int k = exact_cell(zz, hz, Nz);
			const double w = __ddiv_rn(__dsub_rn(zz, __dmul_rn((double)k, hz)), hz);
			double pl = phiTrap[rowBase + k], pr = phiTrap[rowBase + k + 1];
			for (int sp = 0; sp < nS; ++sp) { pl += phiSelf[(size_t)sp * G + rowBase + k]; pr += phiSelf[(size_t)sp * G + rowBase + k + 1]; }
			acc += ((1 - w) * pl + w * pr) * q;
		}
	}
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(out, acc);
}

// Plasma::getNumMacroCentralWell (Source/Plasma.cpp:151-162).
__global__ void k_central_well(const double* __restrict__ z, const PtpSegment* __restrict__ segs, int nSegs,
	const int* __restrict__ left, const int* __restrict__ right, double hz, unsigned long long* __restrict__ out)
{
	unsigned int n = 0;
	for (int s = blockIdx.x; s < nSegs; s += gridDim.x) {
		const PtpSegment seg = segs[s];
		const double lo = __dmul_rn((double)left[seg.row], hz), hi = __dmul_rn((double)right[seg.row], hz);
		for (long long i = seg.begin + threadIdx.x; i < seg.end; i += blockDim.x) {
			const double zz = z[i];
			if (zz >= lo && zz <= hi) ++n;
		}
	}
	for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
	if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, (unsigned long long)n);
}

// Sums for Plasma::getTemperature / getAverageTemperature / getstdDeviation (Source/Plasma.cpp:163-228): over the live rings,
// weight w = 1 on the axis, 8 r elsewhere (ring mass = w massMacro, Source/Plasma.cpp:169-170), speed = mean of the speeds at
// the last two save points (positions and speeds are staggered by dt/2, :180,224) - or the present speed when there is no
// earlier save point. out[0] += sum w, out[1] += sum w speed^2; T = mass out[1] / (KB out[0]). With mark, the present speed
// becomes the save point.
__global__ void __launch_bounds__(256) k_kinetic_sums(const double* __restrict__ z, const double* __restrict__ v, double* __restrict__ vSaved, bool paired, bool mark,
	const PtpSegment* __restrict__ segs, int nSegs, double* __restrict__ out)
{
	double sw = 0.0, ss = 0.0;
	for (int s = blockIdx.x; s < nSegs; s += gridDim.x) {
		const PtpSegment seg = segs[s];
		const double w = seg.row == 0 ? 1.0 : (double)(8 * seg.row);
		for (long long i = seg.begin + threadIdx.x; i < seg.end; i += blockDim.x) {
			const double zz = z[i];
			if (!(zz == zz)) continue;
			const double now = v[i];
			const double speed = paired ? (vSaved[i] + now) / 2 : now;
			sw += w;
			ss += w * speed * speed;
			if (mark) vSaved[i] = now;
		}
	}
	for (int o = 16; o > 0; o >>= 1) { sw += __shfl_xor_sync(0xffffffffu, sw, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
	if ((threadIdx.x & 31) == 0 && sw != 0.0) { atomicAdd(out, sw); atomicAdd(out + 1, ss); }
}

// ---- K5: per-row counting sort by axial cell ---------------------------------------------------------
// [emu-begin] (tests/emu/emu_sort.cpp runs the text between these markers on host threads)
// The live prefix of every row bucket is cut into chunks of SORT_CHUNK slots (host table); one CTA per chunk. Rings arrive
// nearly sorted (the loaders emit them in z order and a re-sort happens long before they mix), so most of a warp shares
// one cell: ranks are formed per warp with match.any - one shared-memory atomic per distinct cell and warp instruction
// instead of 32 colliding ones.
constexpr int SORT_CHUNK = 32768;
struct SortChunk { int row, pad; long long begin, end; };

// rank of this lane among the lanes of the warp with the same key + the counter's value before the warp's update
__device__ __forceinline__ unsigned int warp_claim(unsigned int* hist, int k, bool valid, int lane)
{
	const unsigned int peers = __match_any_sync(0xffffffffu, valid ? k : -1);
	const int leader = __ffs(peers) - 1;
	unsigned int base = 0;
	if (valid && lane == leader) base = atomicAdd(&hist[k], (unsigned int)__popc(peers));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + (unsigned int)__popc(peers & ((1u << lane) - 1u));
}

// pass 1: histogram of live rings per (row, cell)
__global__ void __launch_bounds__(256) k_sort_count(const double* __restrict__ z, const SortChunk* __restrict__ chunks,
	int Nz, double hz, unsigned int* __restrict__ counts)
{
	extern __shared__ unsigned int hist[];                 // [Nz]
	const SortChunk c = chunks[blockIdx.x];
	const int lane = threadIdx.x & 31;
	for (int i = threadIdx.x; i < Nz; i += blockDim.x) hist[i] = 0;
	__syncthreads();
	for (long long i0 = c.begin; i0 < c.end; i0 += blockDim.x) {
		const long long i = i0 + threadIdx.x;
		const double zz = i < c.end ? z[i] : __longlong_as_double(0x7ff8000000000000LL);
		const bool valid = zz == zz;
		warp_claim(hist, valid ? exact_cell(zz, hz, Nz) : 0, valid, lane);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < Nz; i += blockDim.x)
		if (hist[i]) atomicAdd(&counts[(size_t)c.row * Nz + i], hist[i]);
}

// pass 2: exclusive scan of each row's cell counts -> cursor[row][cell] (slot offsets inside the bucket), live[row].
__global__ void __launch_bounds__(256) k_sort_scan(const unsigned int* __restrict__ counts, int Nz,
	unsigned long long* __restrict__ cursor, unsigned long long* __restrict__ live)
{
	__shared__ unsigned long long part[256];
	const int r = blockIdx.x, tid = threadIdx.x;
	const int per = (Nz + 255) / 256;
	const int b = tid * per, e = min(Nz, b + per);
	unsigned long long s = 0;
	for (int i = b; i < e; ++i) s += counts[(size_t)r * Nz + i];
	part[tid] = s;
	__syncthreads();
	if (tid == 0) {
		unsigned long long run = 0;
		for (int i = 0; i < 256; ++i) { const unsigned long long c = part[i]; part[i] = run; run += c; }
		live[r] = run;
	}
	__syncthreads();
	unsigned long long run = part[tid];
	for (int i = b; i < e; ++i) { cursor[(size_t)r * Nz + i] = run; run += counts[(size_t)r * Nz + i]; }
}

// pass 3: scatter. Ranks inside the chunk from the shared histogram, one global atomic per (chunk, non-empty cell) to
// reserve the destination range.
__global__ void __launch_bounds__(256) k_sort_scatter(const double* __restrict__ z, const double* __restrict__ v,
	const long long* __restrict__ id, double* __restrict__ zOut, double* __restrict__ vOut, long long* __restrict__ idOut,
	const long long* __restrict__ rowOff, const SortChunk* __restrict__ chunks, int Nz, double hz, unsigned long long* __restrict__ cursor,
	const double* __restrict__ vs, double* __restrict__ vsOut)
{
	extern __shared__ unsigned int sh[];                   // hist[Nz] then base (as 2 x u32 per cell)
	unsigned int* hist = sh;
	unsigned long long* base = reinterpret_cast<unsigned long long*>(sh + ((Nz + 1) & ~1));
	const SortChunk c = chunks[blockIdx.x];
	const int lane = threadIdx.x & 31;
	const long long b = rowOff[c.row];
	for (int i = threadIdx.x; i < Nz; i += blockDim.x) hist[i] = 0;
	__syncthreads();
	for (long long i0 = c.begin; i0 < c.end; i0 += blockDim.x) {
		const long long i = i0 + threadIdx.x;
		const double zz = i < c.end ? z[i] : __longlong_as_double(0x7ff8000000000000LL);
		const bool valid = zz == zz;
		warp_claim(hist, valid ? exact_cell(zz, hz, Nz) : 0, valid, lane);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < Nz; i += blockDim.x) {
		const unsigned int n = hist[i];
		base[i] = n ? atomicAdd(&cursor[(size_t)c.row * Nz + i], (unsigned long long)n) : 0ULL;
		hist[i] = 0;
	}
	__syncthreads();
	for (long long i0 = c.begin; i0 < c.end; i0 += blockDim.x) {
		const long long i = i0 + threadIdx.x;
		const double zz = i < c.end ? z[i] : __longlong_as_double(0x7ff8000000000000LL);
		const bool valid = zz == zz;
		const int k = valid ? exact_cell(zz, hz, Nz) : 0;
		const unsigned int rank = warp_claim(hist, k, valid, lane);
		if (valid) {
			const long long dst = b + (long long)base[k] + rank;
			zOut[dst] = zz;
			vOut[dst] = v[i];
			idOut[dst] = id[i];
			if (vs) vsOut[dst] = vs[i];                          // speeds at the last save point travel with their rings
		}
	}
}

// Slots of every bucket behind its (new) live prefix get the empty-slot pattern: NaN position, zero speed, id -1.
__global__ void __launch_bounds__(256) k_sort_pad(double* __restrict__ z, double* __restrict__ v, long long* __restrict__ id,
	const long long* __restrict__ rowOff, const unsigned long long* __restrict__ live, const long long* __restrict__ oldLive)
{
	const int r = blockIdx.y;
	const long long b = rowOff[r] + (long long)live[r], e = rowOff[r] + oldLive[r];   // slots beyond the mark already hold the pattern
	const double nan = __longlong_as_double(-1LL);
	for (long long i = b + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (long long)gridDim.x * blockDim.x) {
		z[i] = nan; v[i] = 0.0; id[i] = -1LL;
	}
}
// [emu-end]

} // namespace

// Plan the push kernel's work: the live part of every row bucket is cut into tiles, tiles are dealt to CTAs in
// contiguous ranges of equal length, and inside a CTA's range consecutive tiles of one row are merged into a segment
// as long as their combined axial cell range still fits the thread-private deposit window (with some slack for the drift
// until the next re-plan) - so a z-ordered load of a long plasma (config 5: ~240 cells) gets many narrow segments while
// the default plasma (37 cells) keeps one segment per CTA and row. Also yields the live ring count and validates z.
int ptp_build_segments(ptp_trap* t, ptp_plasma* p)
{
	const long long tile = (long long)t->ringsPerThread * t->threads;
	// which variant of K1 pushes this species (ptp_plasma::hot; the automatic policy sets it in maintain_order)
	p->scatter = p->hot == 1 && ptp_push_scatter_usable(t);
	// 12-bit count field of the packed bins: per thread, or - SCATTER variant - per warp
	const long long maxSegTiles = p->scatter ? 4095 / (32 * t->ringsPerThread) : 4095 / t->ringsPerThread;
	std::vector<PtpSegment> tiles;
	for (int r = 0; r < t->Nr; ++r)
		for (long long b = 0; b < p->rowLive[r]; b += tile) {
			PtpSegment s;
			s.row = r; s.pad = 0;
			s.begin = p->rowOff[r] + b;
			s.end = s.begin + tile;
			tiles.push_back(s);
		}
	std::vector<int2> tb;
	int64_t live = 0;
	PTP_TRY(ptp_tile_bounds(t, p, tiles, tb, &live));
	p->nAlive = live;
	const long long totalTiles = (long long)tiles.size();
	int nCta = t->ctas > 0 ? t->ctas : t->smCount;
	if (p->ctaShare < 1.0) nCta = std::max(1, (int)(nCta * p->ctaShare + 0.5));   // all species in one launch: SMs by live rings
	if (totalTiles < nCta) nCta = (int)totalTiles;
	const int W = t->window < t->Nz ? t->window : t->Nz;
	// A row whose rings fit the window with a little room to spare is never split by cell range: wherever its rings drift
	// inside that range, the window covers them (the reference's default plasma: 37 cells in a 44-cell window). Wider rows are
	// cut into segments that use only part of the window, so that rings can drift for many steps before any of them leaves
	// its segment's window - an out-of-window deposit costs a global atomic on a contended node, ~10^3 x an in-window one.
	const int fitLimit = std::max(1, W - std::max(2, W / 8));
	const int wideLimit = std::max(1, W - (t->planSlack >= 0 ? t->planSlack : W / 2));
	std::vector<int> rowLimit(t->Nr, fitLimit);
	{
		std::vector<int> rlo(t->Nr, INT_MAX), rhi(t->Nr, INT_MIN);
		for (size_t q = 0; q < tiles.size(); ++q)
			if (tb[q].x <= tb[q].y) { rlo[tiles[q].row] = std::min(rlo[tiles[q].row], tb[q].x); rhi[tiles[q].row] = std::max(rhi[tiles[q].row], tb[q].y); }
		for (int r = 0; r < t->Nr; ++r)
			if (rlo[r] <= rhi[r] && rhi[r] - rlo[r] + 1 > fitLimit) rowLimit[r] = std::min(fitLimit, wideLimit);
	}
	const int limit = fitLimit;                                 // cells a tile may span before it counts as wide
	p->segs.clear();
	p->ctaSegBegin.assign(1, 0);
	p->nCta = nCta;
	std::vector<int4> segBounds;
	if (nCta > 0) {
		// pass 1: runs of tiles of one row whose combined cell range fits the window. A tile that is wider than the window
		// on its own (the sparse tails of a z-ordered load, or rings not ordered at tile granularity at all) cannot be helped
		// by splitting: in an ordered load such tiles are rare and become runs of their own so that they do not widen their
		// neighbours' windows; when most tiles are wide (unordered rings - a sort fixes that) rows are not split at all.
		auto isWide = [&](long long q) { return tb[q].x <= tb[q].y && tb[q].y - tb[q].x + 1 > limit; };
		long long wide = 0;
		for (long long q = 0; q < totalTiles; ++q) wide += isWide(q) ? 1 : 0;
		const bool mostlyWide = p->scatter || 2 * wide > totalTiles;   // (SCATTER variant: the window covers the whole plasma, rows are not split by cell range)
		std::vector<std::pair<long long, long long>> runs;      // [first tile, end tile)
		for (long long i = 0; i < totalTiles;) {
			int lo = tb[i].x, hi = tb[i].y;
			long long n = 1;
			const bool alone = !mostlyWide && isWide(i);
			while (!alone && i + n < totalTiles && tiles[i + n].row == tiles[i].row) {
				const int nlo = std::min(lo, tb[i + n].x), nhi = std::max(hi, tb[i + n].y);
				if (!mostlyWide && nlo <= nhi && nhi - nlo + 1 > rowLimit[tiles[i].row]) break;
				lo = nlo; hi = nhi;
				++n;
			}
			runs.emplace_back(i, i + n);
			i += n;
		}
		// pass 2: deal the runs to CTAs in order, balancing cost = tiles + a fixed charge per segment (zeroing and reducing
		// the bins, the flush, the pipeline restart); runs are cut at tile granularity where a CTA's budget ends
		const double segCharge = 4.0;
		long long remTiles = totalTiles;
		int cta = 0;
		double mine = -segCharge;                               // cost dealt to the current CTA (its first segment is free:
		                                                        // every CTA pays that one anyway)
		auto target = [&](size_t runsLeft) { return ((double)remTiles + segCharge * (double)(runsLeft > 0 ? runsLeft - 1 : 0)) / (double)(nCta - cta); };
		double tgt = target(runs.size());
		auto emit = [&](long long a0, long long b0) {
			PtpSegment sg = tiles[a0];
			sg.end = tiles[b0 - 1].end;
			int lo = INT_MAX, hi = INT_MIN;
			for (long long q = a0; q < b0; ++q) { lo = std::min(lo, tb[q].x); hi = std::max(hi, tb[q].y); }
			while ((int)p->ctaSegBegin.size() < cta + 1) p->ctaSegBegin.push_back((int)p->segs.size());
			p->segs.push_back(sg);
			segBounds.push_back(make_int4(lo, hi, lo <= hi ? (lo + hi) / 2 : 0, 0));
			mine += (double)(b0 - a0) + segCharge;
			remTiles -= b0 - a0;
		};
		for (size_t ri = 0; ri < runs.size(); ++ri) {
			long long a0 = runs[ri].first;
			while (a0 < runs[ri].second) {
				if (cta < nCta - 1 && mine >= tgt - 0.5) {          // this CTA is full: re-balance what is left over the rest
					++cta;
					mine = -segCharge;
					tgt = target(runs.size() - ri);
				}
				const double room = cta == nCta - 1 ? 1e300 : tgt - mine;
				long long take = std::min<long long>(runs[ri].second - a0, maxSegTiles);
				if ((double)take + segCharge > room) take = std::max<long long>(1, (long long)(room - segCharge + 0.5));
				take = std::min<long long>(take, runs[ri].second - a0);
				emit(a0, a0 + take);
				a0 += take;
			}
		}
		while ((int)p->ctaSegBegin.size() < nCta + 1) p->ctaSegBegin.push_back((int)p->segs.size());
	}
	if (p->segs.size() > p->segCap) {
		cudaFree(p->dSegs); cudaFree(p->dSegBounds);
		p->dSegs = nullptr; p->dSegBounds = nullptr; p->segCap = 0;
		const size_t cap = p->segs.size() + p->segs.size() / 2 + 64;
		PTP_CUDA(cudaMalloc(&p->dSegs, cap * sizeof(PtpSegment)));
		PTP_CUDA(cudaMalloc(&p->dSegBounds, cap * sizeof(int4)));
		p->segCap = cap;
	}
	if (p->ctaSegBegin.size() > p->ctaCap) {
		cudaFree(p->dCtaSegBegin);
		p->dCtaSegBegin = nullptr; p->ctaCap = 0;
		const size_t cap = p->ctaSegBegin.size() + 64;
		PTP_CUDA(cudaMalloc(&p->dCtaSegBegin, cap * sizeof(int)));
		p->ctaCap = cap;
	}
	if (!p->segs.empty()) {
		PTP_CUDA(cudaMemcpyAsync(p->dSegs, p->segs.data(), p->segs.size() * sizeof(PtpSegment), cudaMemcpyHostToDevice, t->stream));
		PTP_CUDA(cudaMemcpyAsync(p->dCtaSegBegin, p->ctaSegBegin.data(), p->ctaSegBegin.size() * sizeof(int), cudaMemcpyHostToDevice, t->stream));
		PTP_CUDA(cudaMemcpyAsync(p->dSegBounds, segBounds.data(), segBounds.size() * sizeof(int4), cudaMemcpyHostToDevice, t->stream));
		PTP_CUDA(cudaStreamSynchronize(t->stream));
	}
	++t->cfgEpoch;
	p->boundsValid = true;
	return PTP_OK;
}

// K5. Rings of each row are re-ordered by axial cell into the alternate buffers (lost rings dropped), the
// buffers are swapped and the segment tables rebuilt on the shrunken live ranges. Scratch (alternate ring buffers, the
// per-(row, cell) counters) is allocated on the first sort and kept.
int ptp_sort_plasma(ptp_trap* t, ptp_plasma* p)
{
	if (p->cap == 0) return PTP_OK;
	++t->cfgEpoch;
	const int Nr = t->Nr, Nz = t->Nz;
	const bool freshAlt = !p->zAlt;
	if (freshAlt) {
		PTP_CUDA(cudaMalloc(&p->zAlt, p->cap * sizeof(double)));
		PTP_CUDA(cudaMalloc(&p->vAlt, p->cap * sizeof(double)));
		PTP_CUDA(cudaMalloc(&p->idAlt, p->cap * sizeof(long long)));
		// padding pattern everywhere once; later sorts only rewrite the slots between the new and the old live prefix
		PTP_CUDA(cudaMemsetAsync(p->zAlt, 0xFF, p->cap * sizeof(double), t->stream));
		PTP_CUDA(cudaMemsetAsync(p->vAlt, 0, p->cap * sizeof(double), t->stream));
		PTP_CUDA(cudaMemsetAsync(p->idAlt, 0xFF, p->cap * sizeof(long long), t->stream));
		p->altDirty.assign(Nr, 0);
	}
	std::vector<SortChunk> chunks;
	for (int r = 0; r < Nr; ++r)
		for (long long b = 0; b < p->rowLive[r]; b += SORT_CHUNK) {
			SortChunk c;
			c.row = r; c.pad = 0;
			c.begin = p->rowOff[r] + b;
			c.end = p->rowOff[r] + std::min<long long>(p->rowLive[r], b + SORT_CHUNK);
			chunks.push_back(c);
		}
	// scratch: counts [Nr][Nz] u32 | cursor [Nr][Nz] u64 | live [Nr] u64 | oldLive [Nr] i64 | chunk table
	const size_t nCells = (size_t)Nr * Nz;
	const size_t need = nCells * sizeof(unsigned int) + (nCells + 2 * (size_t)Nr) * sizeof(unsigned long long) + (chunks.size() + 1) * sizeof(SortChunk) + 64;
	if (p->sortScratchBytes < need) {
		cudaFree(p->sortScratch);
		p->sortScratch = nullptr; p->sortScratchBytes = 0;
		PTP_CUDA(cudaMalloc(&p->sortScratch, need));
		p->sortScratchBytes = need;
	}
	unsigned long long* dCursor = reinterpret_cast<unsigned long long*>(p->sortScratch);
	unsigned long long* dLive = dCursor + nCells;
	long long* dOldLive = reinterpret_cast<long long*>(dLive + Nr);
	SortChunk* dChunks = reinterpret_cast<SortChunk*>(dOldLive + Nr);
	unsigned int* dCounts = reinterpret_cast<unsigned int*>(dChunks + chunks.size() + 1);
	PTP_CUDA(cudaMemsetAsync(dCounts, 0, nCells * sizeof(unsigned int), t->stream));
	PTP_CUDA(cudaMemcpyAsync(dOldLive, p->altDirty.data(), (size_t)Nr * sizeof(long long), cudaMemcpyHostToDevice, t->stream));
	if (!chunks.empty()) PTP_CUDA(cudaMemcpyAsync(dChunks, chunks.data(), chunks.size() * sizeof(SortChunk), cudaMemcpyHostToDevice, t->stream));
	const size_t smCount = (size_t)Nz * sizeof(unsigned int);
	const size_t smScatter = (size_t)((Nz + 1) & ~1) * sizeof(unsigned int) + (size_t)Nz * sizeof(unsigned long long);
	if (smScatter > t->smemMax) { ptp_set_error("ptp_trap_sort: Nz too large for the sort kernels of this build"); return PTP_EINVAL; }
	PTP_CUDA(cudaFuncSetAttribute(k_sort_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smCount));
	PTP_CUDA(cudaFuncSetAttribute(k_sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smScatter));
	const unsigned int nChunks = (unsigned int)chunks.size();
	if (nChunks) k_sort_count<<<nChunks, 256, smCount, t->stream>>>(p->z, dChunks, Nz, t->hz, dCounts);
	k_sort_scan<<<Nr, 256, 0, t->stream>>>(dCounts, Nz, dCursor, dLive);
	if (p->vSaved && !p->vSavedAlt) PTP_CUDA(cudaMalloc(&p->vSavedAlt, p->cap * sizeof(double)));
	if (nChunks) k_sort_scatter<<<nChunks, 256, smScatter, t->stream>>>(p->z, p->v, p->id, p->zAlt, p->vAlt, p->idAlt, p->dRowOff, dChunks, Nz, t->hz, dCursor,
		p->vSaved, p->vSavedAlt);
	// the alternate buffers hold the empty-slot pattern beyond altDirty (what they held when they were last the primary
	// ones): only the slots between the new live prefix and that mark need it again
	k_sort_pad<<<dim3(32, Nr), 256, 0, t->stream>>>(p->zAlt, p->vAlt, p->idAlt, p->dRowOff, dLive, dOldLive);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return ptp_cuda_fail(e, "sort launch", __FILE__, __LINE__);
	t->lastLaunches += 4;
	std::vector<unsigned long long> live(Nr);
	PTP_CUDA(cudaMemcpyAsync(live.data(), dLive, Nr * sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	std::swap(p->z, p->zAlt); std::swap(p->v, p->vAlt); std::swap(p->id, p->idAlt);
	if (p->vSaved) std::swap(p->vSaved, p->vSavedAlt);
	long long total = 0;
	for (int r = 0; r < Nr; ++r) { p->altDirty[r] = p->rowLive[r]; p->rowLive[r] = (long long)live[r]; total += (long long)live[r]; }
	p->nAlive = total;
	PTP_CUDA(cudaMemsetAsync(p->dLost, 0, 2 * sizeof(unsigned long long), t->stream));
	p->nUploaded = total;                                  // loss counter restarts from the compacted population
	p->farBaseline = -1.0;
	return ptp_build_segments(t, p);
}

// Row-bucketed storage for count[j] rings in row j (n in total): bucket offsets, (re)allocation, the empty-slot pattern in
// every bucket's padding, the species' RHS factor, a cleared loss counter and the fixed-point scale. The caller fills the
// live prefix of every bucket (z, v, id) and then builds the segments.
int ptp_plasma_set_layout(ptp_plasma* p, const std::vector<long long>& count, int64_t n, double macroChargeDensity)
{
	ptp_trap* t = p->trap;
	const int Nr = t->Nr;
	++t->cfgEpoch;                                               // ring buffers / segment tables change: cached step graph is stale
	++t->layoutEpoch;                                            // ... and so is the outermost populated row
	p->rowOff.assign(Nr + 1, 0);
	p->rowLive.assign(Nr, 0);
	for (int j = 0; j < Nr; ++j) {
		p->rowLive[j] = count[j];
		p->rowOff[j + 1] = p->rowOff[j] + (count[j] + PTP_ROW_ALIGN - 1) / PTP_ROW_ALIGN * PTP_ROW_ALIGN;
	}
	const long long newCap = p->rowOff[Nr];
	if (newCap != p->cap) {                                    // same-size reloads keep their buffers
		cudaFree(p->z); cudaFree(p->v); cudaFree(p->id);
		p->z = p->v = nullptr; p->id = nullptr;
	}
	cudaFree(p->zAlt); cudaFree(p->vAlt); cudaFree(p->idAlt); cudaFree(p->dRowOff); cudaFree(p->vSaved); cudaFree(p->vSavedAlt);
	p->zAlt = p->vAlt = nullptr; p->idAlt = nullptr; p->dRowOff = nullptr; p->vSaved = p->vSavedAlt = nullptr; p->vSavedValid = false;
	if (!p->dLossLog) {
		p->lossCap = 1 << 16;
		PTP_CUDA(cudaMalloc(&p->dLossLog, (4 + 2 * (size_t)p->lossCap) * sizeof(unsigned long long)));
	}
	PTP_CUDA(cudaMemsetAsync(p->dLossLog, 0, 4 * sizeof(unsigned long long), t->stream));
	p->cap = newCap;
	p->farBaseline = -1.0;
	p->lastSortStep = -1;
	p->quickSorts = 0;
	if (p->hotAuto) { p->hot = -1; p->hotAuto = false; }         // new rings: the policy decides again
	t->stepsSinceCheck = 0;
	t->nextCheckSteps = 4;
	p->nUploaded = n;
	p->nAlive = n;
	p->macroChargeDensity = macroChargeDensity;
	const double scale = -macroChargeDensity / 8.8541878128e-12;     // Source/Plasma.cpp:91-92, Source/Constants.hpp:12
	PTP_CUDA(cudaMemcpyAsync(t->dScale + p->index, &scale, sizeof(double), cudaMemcpyHostToDevice, t->stream));
	PTP_CUDA(cudaMemsetAsync(p->dLost, 0, 2 * sizeof(unsigned long long), t->stream));
	PTP_CUDA(cudaMalloc(&p->dRowOff, (Nr + 1) * sizeof(long long)));
	PTP_CUDA(cudaMemcpyAsync(p->dRowOff, p->rowOff.data(), (Nr + 1) * sizeof(long long), cudaMemcpyHostToDevice, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));                  // `scale` is a stack variable
	if (p->cap > 0) {
		if (!p->z) {
			PTP_CUDA(cudaMalloc(&p->z, p->cap * sizeof(double)));
			PTP_CUDA(cudaMalloc(&p->v, p->cap * sizeof(double)));
			PTP_CUDA(cudaMalloc(&p->id, p->cap * sizeof(long long)));
		}
		// only the padding at the end of every bucket needs the empty-slot pattern (all-ones = NaN / id -1)
		for (int j = 0; j < Nr; ++j) {
			const long long padBegin = p->rowOff[j] + count[j], padLen = p->rowOff[j + 1] - padBegin;
			if (padLen <= 0) continue;
			PTP_CUDA(cudaMemsetAsync(p->z + padBegin, 0xFF, padLen * sizeof(double), t->stream));
			PTP_CUDA(cudaMemsetAsync(p->v + padBegin, 0, padLen * sizeof(double), t->stream));
			PTP_CUDA(cudaMemsetAsync(p->id + padBegin, 0xFF, padLen * sizeof(long long), t->stream));
		}
	}
	// (the fixed-point scale follows from the global ring count: ptp_layout_sync, first use after this load)
	p->encValid = false;
	return PTP_OK;
}

// ---- C ABI: particle-side entry points ---------------------------------------------------------------
extern "C" {

int ptp_plasma_upload(ptp_plasma* p, int64_t n, const int32_t* r, const double* z, const double* v, double macroChargeDensity)
{
	if (!p || n < 0 || (n > 0 && (!r || !z || !v))) { ptp_set_error("ptp_plasma_upload: bad arguments"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	++t->cfgEpoch;                                               // ring buffers / segment tables change: cached step graph is stale
	const int Nr = t->Nr;
	// row histogram + "already bucketed?" test, split over host threads (one pass over r is the only O(n) host work
	// of a sorted upload; the loaders emit rows in ascending order)
	std::vector<long long> count(Nr, 0);
	bool sorted = true, bad = false;
	{
		const int nThreads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())), n / 1000000));
		std::vector<std::vector<long long>> part(nThreads, std::vector<long long>(Nr, 0));
		std::vector<char> partSorted(nThreads, 1), partBad(nThreads, 0);
		auto scan = [&](int th) {
			const int64_t b = n * th / nThreads, e = n * (th + 1) / nThreads;
			std::vector<long long>& c = part[th];
			int prev = b > 0 ? r[b - 1] : INT_MIN;
			for (int64_t i = b; i < e; ++i) {
				const int ri = r[i];
				if (ri < 0 || ri >= Nr) { partBad[th] = 1; return; }
				++c[ri];
				if (ri < prev) partSorted[th] = 0;
				prev = ri;
			}
		};
		std::vector<std::thread> pool;
		for (int th = 1; th < nThreads; ++th) pool.emplace_back(scan, th);
		scan(0);
		for (std::thread& th : pool) th.join();
		for (int th = 0; th < nThreads; ++th) {
			bad |= partBad[th] != 0;
			sorted &= partSorted[th] != 0;
			for (int j = 0; j < Nr; ++j) count[j] += part[th][j];
		}
	}
	if (bad) { ptp_set_error("ptp_plasma_upload: radial index outside [0, Nr)"); return PTP_EINVAL; }
	std::vector<long long> rowSrc(Nr + 1, 0);
	for (int j = 0; j < Nr; ++j) rowSrc[j + 1] = rowSrc[j] + count[j];
	PTP_TRY(ptp_plasma_set_layout(p, count, n, macroChargeDensity));
	if (p->cap > 0) {
		if (sorted) {
			// loaders emit rows in ascending order (Source/Plasma.cpp:510-526): each bucket is one contiguous copy
			for (int j = 0; j < Nr; ++j) {
				if (!count[j]) continue;
				PTP_CUDA(cudaMemcpyAsync(p->z + p->rowOff[j], z + rowSrc[j], count[j] * sizeof(double), cudaMemcpyHostToDevice, t->stream));
				PTP_CUDA(cudaMemcpyAsync(p->v + p->rowOff[j], v + rowSrc[j], count[j] * sizeof(double), cudaMemcpyHostToDevice, t->stream));
			}
			long long* dRowSrc = nullptr;
			PTP_CUDA(cudaMalloc(&dRowSrc, (Nr + 1) * sizeof(long long)));
			PTP_CUDA(cudaMemcpyAsync(dRowSrc, rowSrc.data(), (Nr + 1) * sizeof(long long), cudaMemcpyHostToDevice, t->stream));
			k_iota_rows<<<dim3(64, Nr), 256, 0, t->stream>>>(p->id, p->dRowOff, dRowSrc, Nr);
			PTP_CUDA(cudaStreamSynchronize(t->stream));
			cudaFree(dRowSrc);
		}
		else {
			// general order (e.g. after the reference's swap-with-back removals): stable counting sort on the host
			std::vector<double> zs(p->cap), vs(p->cap);
			std::vector<long long> ids(p->cap, -1);
			std::vector<long long> cur(p->rowOff.begin(), p->rowOff.end() - 1);
			const double nan = std::nan("");
			std::fill(zs.begin(), zs.end(), nan);
			for (int64_t i = 0; i < n; ++i) {
				const long long d = cur[r[i]]++;
				zs[d] = z[i]; vs[d] = v[i]; ids[d] = i;
			}
			PTP_CUDA(cudaMemcpyAsync(p->z, zs.data(), p->cap * sizeof(double), cudaMemcpyHostToDevice, t->stream));
			PTP_CUDA(cudaMemcpyAsync(p->v, vs.data(), p->cap * sizeof(double), cudaMemcpyHostToDevice, t->stream));
			PTP_CUDA(cudaMemcpyAsync(p->id, ids.data(), p->cap * sizeof(long long), cudaMemcpyHostToDevice, t->stream));
			PTP_CUDA(cudaStreamSynchronize(t->stream));
		}
	}
	return ptp_build_segments(t, p);
}

int ptp_plasma_count(ptp_plasma* p, int64_t* nAlive)
{
	if (!p || !nAlive) { ptp_set_error("ptp_plasma_count: null argument"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	unsigned long long lost = 0;
	PTP_CUDA(cudaMemcpyAsync(&lost, p->dLost, sizeof(lost), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	p->nAlive = p->nUploaded - (int64_t)lost;
	*nAlive = p->nAlive;
	return PTP_OK;
}

static int download_impl(ptp_plasma* p, int32_t* r, double* z, double* v, int64_t* id, int32_t* kOut, int32_t* idxOut)
{
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	if (p->cap == 0) return PTP_OK;
	std::vector<double> zs(p->cap), vs;
	std::vector<long long> ids;
	std::vector<int> ks;
	PTP_CUDA(cudaMemcpyAsync(zs.data(), p->z, p->cap * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
	if (v) { vs.resize(p->cap); PTP_CUDA(cudaMemcpyAsync(vs.data(), p->v, p->cap * sizeof(double), cudaMemcpyDeviceToHost, t->stream)); }
	if (id) { ids.resize(p->cap); PTP_CUDA(cudaMemcpyAsync(ids.data(), p->id, p->cap * sizeof(long long), cudaMemcpyDeviceToHost, t->stream)); }
	if (kOut || idxOut) {
		int* dK = nullptr;
		PTP_CUDA(cudaMalloc(&dK, p->cap * sizeof(int)));
		k_cell_index<<<(unsigned)((p->cap + 255) / 256), 256, 0, t->stream>>>(p->z, p->cap, t->hz, t->Nz, dK);
		ks.resize(p->cap);
		PTP_CUDA(cudaMemcpyAsync(ks.data(), dK, p->cap * sizeof(int), cudaMemcpyDeviceToHost, t->stream));
		PTP_CUDA(cudaStreamSynchronize(t->stream));
		cudaFree(dK);
	}
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	int64_t o = 0;
	for (int j = 0; j < t->Nr; ++j) {
		for (long long i = p->rowOff[j]; i < p->rowOff[j] + p->rowLive[j]; ++i) {
			if (!(zs[i] == zs[i])) continue;
			if (r) r[o] = j;
			if (z) z[o] = zs[i];
			if (v) v[o] = vs[i];
			if (id) id[o] = ids[i];
			if (kOut) kOut[o] = ks[i];
			if (idxOut) idxOut[o] = (t->Nz + 1) * j + ks[i];
			++o;
		}
	}
	p->nAlive = o;
	return PTP_OK;
}

int ptp_plasma_download(ptp_plasma* p, int32_t* r, double* z, double* v, int64_t* id)
{
	if (!p) { ptp_set_error("ptp_plasma_download: null plasma"); return PTP_EINVAL; }
	return download_impl(p, r, z, v, id, nullptr, nullptr);
}

int ptp_plasma_cell_index(ptp_plasma* p, int32_t* k, int32_t* idx)
{
	if (!p) { ptp_set_error("ptp_plasma_cell_index: null plasma"); return PTP_EINVAL; }
	return download_impl(p, nullptr, nullptr, nullptr, nullptr, k, idx);
}

int ptp_plasma_potential_energy(ptp_plasma* p, double chargeMacro, double* pe)
{
	if (!p || !pe) { ptp_set_error("ptp_plasma_potential_energy: null argument"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	*pe = 0.0;
	if (p->segs.empty()) return PTP_OK;
	for (int j = t->Nr - 1; j >= t->phiRows; --j)               // rings in rows the last step's solve left out (loaded since)
		if (p->rowLive[j] > 0) { PTP_TRY(ptp_materialize_fields(t)); break; }
	double* dOut = nullptr;
	PTP_CUDA(cudaMalloc(&dOut, sizeof(double)));
	PTP_CUDA(cudaMemsetAsync(dOut, 0, sizeof(double), t->stream));
	int grid = std::min<int>((int)p->segs.size(), t->smCount * 4);
	k_potential_energy<<<grid, 256, 0, t->stream>>>(p->z, p->dSegs, (int)p->segs.size(), t->phiTrap, t->phiSelfAll,
		(int)t->plasmas.size(), t->G, t->Nz, t->hz, chargeMacro, dOut);
	double h = 0;
	PTP_CUDA(cudaMemcpyAsync(&h, dOut, sizeof(double), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	cudaFree(dOut);
	*pe = h / 2;                                            // Source/Plasma.cpp:251
	return PTP_OK;
}

int ptp_plasma_kinetic_sums(ptp_plasma* p, int markSavePoint, double* sumW, double* sumWS2, int* paired)
{
	if (!p || !sumW || !sumWS2) { ptp_set_error("ptp_plasma_kinetic_sums: null argument"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	*sumW = *sumWS2 = 0.0;
	if (paired) *paired = p->vSavedValid ? 1 : 0;
	if (p->segs.empty()) return PTP_OK;
	if (markSavePoint && !p->vSaved) PTP_CUDA(cudaMalloc(&p->vSaved, p->cap * sizeof(double)));
	double* dOut = nullptr;
	PTP_CUDA(cudaMalloc(&dOut, 2 * sizeof(double)));
	PTP_CUDA(cudaMemsetAsync(dOut, 0, 2 * sizeof(double), t->stream));
	const int grid = std::min<int>((int)p->segs.size(), t->smCount * 4);
	k_kinetic_sums<<<grid, 256, 0, t->stream>>>(p->z, p->v, p->vSaved, p->vSavedValid, markSavePoint != 0, p->dSegs, (int)p->segs.size(), dOut);
	double h[2] = { 0, 0 };
	cudaError_t e = cudaMemcpyAsync(h, dOut, sizeof(h), cudaMemcpyDeviceToHost, t->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
	cudaFree(dOut);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_kinetic_sums", __FILE__, __LINE__);
	if (markSavePoint) p->vSavedValid = true;
	*sumW = h[0];
	*sumWS2 = h[1];
	return PTP_OK;
}

int ptp_plasma_download_row(ptp_plasma* p, int row, int64_t nMax, double* z, double* v, int64_t* id, int64_t* n)
{
	if (!p || !n || row < 0 || row >= p->trap->Nr || nMax < 0) { ptp_set_error("ptp_plasma_download_row: bad arguments"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	*n = 0;
	const long long live = p->cap ? p->rowLive[row] : 0;           // slots of the bucket that may hold rings: one contiguous slice
	if (live == 0) return PTP_OK;
	const long long b = p->rowOff[row];
	std::vector<double> zs((size_t)live), vs(v ? (size_t)live : 0);
	std::vector<long long> ids(id ? (size_t)live : 0);
	PTP_CUDA(cudaMemcpyAsync(zs.data(), p->z + b, (size_t)live * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
	if (v) PTP_CUDA(cudaMemcpyAsync(vs.data(), p->v + b, (size_t)live * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
	if (id) PTP_CUDA(cudaMemcpyAsync(ids.data(), p->id + b, (size_t)live * sizeof(long long), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	int64_t o = 0;
	for (long long i = 0; i < live; ++i) {
		if (!(zs[(size_t)i] == zs[(size_t)i])) continue;          // lost ring / empty slot
		if (o < nMax) {
			if (z) z[o] = zs[(size_t)i];
			if (v) v[o] = vs[(size_t)i];
			if (id) id[o] = ids[(size_t)i];
		}
		++o;
	}
	*n = o;
	if (o > nMax) { ptp_set_error("ptp_plasma_download_row: buffers too small"); return PTP_EINVAL; }
	return PTP_OK;
}

int ptp_plasma_loss_log(ptp_plasma* p, int64_t first, int64_t nMax, int64_t* ids, int64_t* steps, int64_t* total, int* overflowed)
{
	if (!p || !total || first < 0 || nMax < 0) { ptp_set_error("ptp_plasma_loss_log: bad arguments"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	*total = 0;
	if (overflowed) *overflowed = 0;
	if (!p->dLossLog) return PTP_OK;
	unsigned long long count = 0;
	PTP_CUDA(cudaMemcpyAsync(&count, p->dLossLog, sizeof(count), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	*total = (int64_t)count;
	if (overflowed && (long long)count > p->lossCap) *overflowed = 1;
	const int64_t have = std::min<int64_t>((int64_t)count, p->lossCap);
	const int64_t take = std::max<int64_t>(0, std::min<int64_t>(nMax, have - first));
	if (take == 0 || !ids || !steps) return PTP_OK;
	std::vector<unsigned long long> h(2 * (size_t)take);
	PTP_CUDA(cudaMemcpyAsync(h.data(), p->dLossLog + 4 + 2 * first, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	for (int64_t i = 0; i < take; ++i) { ids[i] = (int64_t)h[2 * (size_t)i]; steps[i] = (int64_t)h[2 * (size_t)i + 1]; }
	return PTP_OK;
}

int ptp_plasma_count_central_well(ptp_plasma* p, const int32_t* limitLeft, const int32_t* limitRight, int64_t* n)
{
	if (!p || !limitLeft || !limitRight || !n) { ptp_set_error("ptp_plasma_count_central_well: null argument"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	*n = 0;
	if (p->segs.empty()) return PTP_OK;
	int* dLim = nullptr;
	unsigned long long* dOut = nullptr;
	PTP_CUDA(cudaMalloc(&dLim, 2 * t->Nr * sizeof(int)));
	PTP_CUDA(cudaMalloc(&dOut, sizeof(unsigned long long)));
	PTP_CUDA(cudaMemcpyAsync(dLim, limitLeft, t->Nr * sizeof(int), cudaMemcpyHostToDevice, t->stream));
	PTP_CUDA(cudaMemcpyAsync(dLim + t->Nr, limitRight, t->Nr * sizeof(int), cudaMemcpyHostToDevice, t->stream));
	PTP_CUDA(cudaMemsetAsync(dOut, 0, sizeof(unsigned long long), t->stream));
	int grid = std::min<int>((int)p->segs.size(), t->smCount * 4);
	k_central_well<<<grid, 256, 0, t->stream>>>(p->z, p->dSegs, (int)p->segs.size(), dLim, dLim + t->Nr, t->hz, dOut);
	unsigned long long h = 0;
	PTP_CUDA(cudaMemcpyAsync(&h, dOut, sizeof(h), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	cudaFree(dLim); cudaFree(dOut);
	*n = (int64_t)h;
	return PTP_OK;
}

} // extern "C"
