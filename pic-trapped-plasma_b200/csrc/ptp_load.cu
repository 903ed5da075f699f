// Ring placement of the reference's loaders on the device ("next" row f-1 of SURVEY 8).
//
// Plasma::loadProfile and Plasma::loadDensityFile end with the same code (reference Source/Plasma.cpp:464-526 and
// :558-620): per radial row the cumulative charge along z, the number of rings of the row, equally spaced charge
// quantiles inverted by linear interpolation, and one Maxwellian speed per ring drawn from a freshly seeded
// std::default_random_engine through std::normal_distribution. At 1e8 rings that serial loop plus the 2 GB upload
// dominates start-up, so here
//   * the O(G) part (cumulative sums, chargeMacro, rings per row) stays on the host in the reference's own serial
//     order - it fixes integers (ring counts) and must be bit-exact;
//   * k_place inverts every quantile independently (binary search instead of the reference's running index - the same
//     node for a density of one sign, which is what a single species has) with the reference's expression order and
//     IEEE divisions: positions are bit-identical to the reference;
//   * the speeds reproduce the reference's deviate stream IN PARALLEL: libstdc++'s default_random_engine is
//     minstd_rand0, x -> 16807 x mod (2^31 - 1), which can jump ahead by modular exponentiation; generate_canonical
//     takes two engine calls per uniform; normal_distribution is Marsaglia's polar method, so attempt a consumes calls
//     4a+1 .. 4a+4 whether or not it is accepted, and accepted attempt number n yields ring 2n (y * mult) and ring
//     2n+1 (x * mult). k_rng_count / k_rng_emit evaluate all attempts independently and compact the accepted ones with
//     a prefix sum. Uniforms, acceptance and therefore the assignment of deviates to rings are bit-exact; the deviate
//     itself goes through log(), where CUDA and glibc may differ in the last bit (speeds agree to ~2 ulp).
// Rings of shard s of S are rings i = s (mod S) of every row (the multi-GPU partition of SURVEY 8e); every shard evaluates
// the whole deviate stream and keeps its own part.
#include "ptp_internal.h"

#include <cmath>
#include <vector>

namespace {

// [emu-begin] (tests/emu/emu_rng.cpp runs the deviate-stream kernels on host threads against libstdc++'s own engine)
constexpr unsigned long long kM = 2147483647ULL;            // 2^31 - 1
constexpr unsigned long long kA = 16807ULL;
constexpr int RNG_CH = 16;                                  // attempts per thread

__device__ __forceinline__ unsigned long long lcg_next(unsigned long long x) { return (x * kA) % kM; }

__device__ unsigned long long lcg_pow(unsigned long long e)    // 16807^e mod (2^31 - 1)
{
	unsigned long long r = 1, b = kA;
	while (e) {
		if (e & 1ULL) r = (r * b) % kM;
		b = (b * b) % kM;
		e >>= 1;
	}
	return r;
}

// std::generate_canonical<double, 53>(minstd_rand0): two calls, sum = (u1 - 1) + (u2 - 1) * r, r = 2^31 - 2, over r * r
// (bits/random.tcc:3349-3381), every operation rounded to double.
__device__ __forceinline__ double canonical(unsigned long long u1, unsigned long long u2)
{
	const double r = 2147483646.0;
	const double sum = __dadd_rn((double)(u1 - 1ULL), __dmul_rn((double)(u2 - 1ULL), r));
	const double ret = __ddiv_rn(sum, __dmul_rn(r, r));
	return ret >= 1.0 ? 0.99999999999999988898 : ret;
}

// One attempt of the polar method (bits/random.tcc:1827-1836). Advances the engine by four calls.
__device__ __forceinline__ bool polar_attempt(unsigned long long& state, double& x, double& y, double& r2)
{
	const unsigned long long u1 = lcg_next(state), u2 = lcg_next(u1), u3 = lcg_next(u2), u4 = lcg_next(u3);
	state = u4;
	x = __dsub_rn(__dmul_rn(2.0, canonical(u1, u2)), 1.0);
	y = __dsub_rn(__dmul_rn(2.0, canonical(u3, u4)), 1.0);
	r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
	return !(r2 > 1.0 || r2 == 0.0);
}

__global__ void __launch_bounds__(256) k_rng_count(long long nAttempts, unsigned int* __restrict__ blockCount)
{
	const long long a0 = ((long long)blockIdx.x * 256 + threadIdx.x) * RNG_CH;
	unsigned int c = 0;
	if (a0 < nAttempts) {
		unsigned long long state = lcg_pow(4ULL * (unsigned long long)a0);
		const int n = (int)min((long long)RNG_CH, nAttempts - a0);
		for (int i = 0; i < n; ++i) {
			double x, y, r2;
			c += polar_attempt(state, x, y, r2) ? 1u : 0u;
		}
	}
	__shared__ unsigned int sC[8];
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0) sC[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned int s = 0;
		for (int w = 0; w < 8; ++w) s += sC[w];
		blockCount[blockIdx.x] = s;
	}
}

// exclusive scan of the block counts (one CTA; n is a few ten thousand); total -> out[n]
__global__ void __launch_bounds__(1024) k_rng_scan(const unsigned int* __restrict__ in, unsigned long long* __restrict__ out, int n)
{
	__shared__ unsigned long long sW[32];
	__shared__ unsigned long long sCarry;
	if (threadIdx.x == 0) sCarry = 0;
	__syncthreads();
	for (int base = 0; base < n; base += 1024) {
		const int i = base + threadIdx.x;
		const unsigned long long v = i < n ? in[i] : 0ULL;
		unsigned long long incl = v;
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
			if ((threadIdx.x & 31) >= o) incl += t;
		}
		if ((threadIdx.x & 31) == 31) sW[threadIdx.x >> 5] = incl;
		__syncthreads();
		if (threadIdx.x < 32) {
			unsigned long long w = sW[threadIdx.x], wi = w;
			for (int o = 1; o < 32; o <<= 1) {
				const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
				if (threadIdx.x >= o) wi += t;
			}
			sW[threadIdx.x] = wi - w;                               // exclusive warp offsets
		}
		__syncthreads();
		const unsigned long long carry = sCarry;
		if (i < n) out[i] = carry + sW[threadIdx.x >> 5] + incl - v;
		__syncthreads();
		if (threadIdx.x == 1023) sCarry = carry + sW[31] + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) out[n] = sCarry;
}

// normals[2n] = y * mult, normals[2n + 1] = x * mult of accepted attempt n (the order std::normal_distribution returns them)
__global__ void __launch_bounds__(256) k_rng_emit(long long nAttempts, const unsigned long long* __restrict__ blockOffset, double* __restrict__ normals,
	long long nNormals)
{
	const long long a0 = ((long long)blockIdx.x * 256 + threadIdx.x) * RNG_CH;
	double xs[RNG_CH], ys[RNG_CH], rs[RNG_CH];
	unsigned int mask = 0;
	if (a0 < nAttempts) {
		unsigned long long state = lcg_pow(4ULL * (unsigned long long)a0);
		const int n = (int)min((long long)RNG_CH, nAttempts - a0);
#pragma unroll
		for (int i = 0; i < RNG_CH; ++i)
			if (i < n && polar_attempt(state, xs[i], ys[i], rs[i])) mask |= 1u << i;
	}
	// exclusive scan of the per-thread counts inside the CTA
	const unsigned int c = __popc(mask);
	unsigned int incl = c;
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
		if ((threadIdx.x & 31) >= o) incl += t;
	}
	__shared__ unsigned int sW[8];
	if ((threadIdx.x & 31) == 31) sW[threadIdx.x >> 5] = incl;
	__syncthreads();
	unsigned int before = 0;
	for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += sW[w];
	long long slot = (long long)blockOffset[blockIdx.x] + before + incl - c;
#pragma unroll
	for (int i = 0; i < RNG_CH; ++i) {
		if (!((mask >> i) & 1u)) continue;
		if (2 * slot < nNormals) {
			// sqrt(-2 * log(r2) / r2)   bits/random.tcc:1836
			const double mult = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, log(rs[i])), rs[i]));
			normals[2 * slot] = __dmul_rn(ys[i], mult);
			if (2 * slot + 1 < nNormals) normals[2 * slot + 1] = __dmul_rn(xs[i], mult);
		}
		++slot;
	}
}

// [emu-end]

// Device temporaries of one load, released on every exit path.
struct LoadScratch {
	double* normals = nullptr;
	unsigned int* blockCount = nullptr;
	unsigned long long* blockOffset = nullptr;
	double* cum = nullptr;
	void* rows = nullptr;
	void dropStream() { cudaFree(normals); cudaFree(blockCount); cudaFree(blockOffset); normals = nullptr; blockCount = nullptr; blockOffset = nullptr; }
	~LoadScratch() { dropStream(); cudaFree(cum); cudaFree(rows); }
};

// [place-begin] (tests/emu/emu_place.cpp: the placement kernel and, further down, the host arithmetic in front of it)
struct PlaceRow {
	int row, pad;
	long long perRow;       // rings of the row over all shards (numAtR)
	long long local;        // rings of this shard
	long long slot0;        // first slot of the row bucket
	long long id0;          // id of the shard's first ring of the row
	long long stream0;      // index of the row's first ring in the deviate stream
	double quantum;         // deltaQ = cumulative.back() / (numAtR + 1)
};

// Source/Plasma.cpp:512-524 for ring i = shard + q * nShards of row rows[blockIdx.y].row.
__global__ void __launch_bounds__(256) k_place(const PlaceRow* __restrict__ rows, const double* __restrict__ cum, const double* __restrict__ normals,
	int n1, double hz, double sigma, int shard, int nShards, double* __restrict__ z, double* __restrict__ v, long long* __restrict__ id)
{
	const PlaceRow pr = rows[blockIdx.y];
	const double* c = cum + (size_t)pr.row * n1;
	for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < pr.local; q += (long long)gridDim.x * 256) {
		const long long i = shard + q * nShards;
		const double target = __dmul_rn(pr.quantum, (double)(i + 1));          // deltaQ * (i + 1)
		const double at = fabs(target);
		int lo = 1, hi = n1 - 1;                                                // first node with |cumulative| >= |target|
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (fabs(c[mid]) < at) lo = mid + 1;
			else hi = mid;
		}
		const double c0 = c[lo - 1], c1 = c[lo];
		// (currentIndex - 1) * hz + hz / 2 + hz * (currentInvert - cum[ci - 1]) / (cum[ci] - cum[ci - 1])
		const double head = __dadd_rn(__dmul_rn((double)(lo - 1), hz), __ddiv_rn(hz, 2.0));
		const double tail = __ddiv_rn(__dmul_rn(hz, __dsub_rn(target, c0)), __dsub_rn(c1, c0));
		z[pr.slot0 + q] = __dadd_rn(head, tail);
		v[pr.slot0 + q] = __dadd_rn(__dmul_rn(normals[pr.stream0 + i], sigma), 0.0);   // ret * stddev + mean
		id[pr.slot0 + q] = pr.id0 + q;
	}
}
// [place-end]

} // namespace

extern "C" int ptp_plasma_load_density(ptp_plasma* p, const double* density, double temperature, int64_t numMacro, int shard, int nShards,
	double* chargeMacroOut, int64_t* nAtRow, int64_t* nLoaded)
{
	if (!p || !density || !(temperature > 0) || numMacro <= 0 || nShards < 1 || shard < 0 || shard >= nShards) {
		ptp_set_error("ptp_plasma_load_density: bad arguments");
		return PTP_EINVAL;
	}
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	// [counts-begin] (pure host arithmetic up to [counts-end])
	const int Nr = t->Nr, n1 = t->Nz + 1;
	const double hz = t->hz, hr = t->hr;
	const double PI = 3.141592653589793238463, KB = 1.380649e-23;          // Source/Constants.hpp:15-16
	// cumulative charge per row, chargeMacro, rings per row: Source/Plasma.cpp:466-500, same serial order
	std::vector<double> cum((size_t)Nr * n1);
	for (int j = 0; j < Nr; ++j) {
		const double volume = j == 0 ? PI * hz * hr * hr / 4 : hz * hr * 2 * PI * j * hr;
		double running = 0;
		for (int k = 0; k < n1; ++k) {
			running += volume * density[(size_t)n1 * j + k];
			cum[(size_t)n1 * j + k] = running;
		}
	}
	double axisCharge = cum[(size_t)n1 - 1];
	for (int j = 1; j < Nr; ++j) axisCharge += cum[(size_t)n1 * j + n1 - 1] / (8 * j);
	const double chargeMacro = axisCharge / (double)numMacro;
	if (!(chargeMacro != 0.0) || !std::isfinite(chargeMacro)) { ptp_set_error("ptp_plasma_load_density: the density holds no charge"); return PTP_EINVAL; }
	const double mcd = 4 * chargeMacro / (PI * hz * hr * hr);              // :494
	std::vector<long long> perRow(Nr), count(Nr);
	std::vector<PlaceRow> rows;
	long long total = 0, local = 0;
	for (int j = 0; j < Nr; ++j) {
		const double last = cum[(size_t)n1 * j + n1 - 1];
		const long long n = (long long)std::round(j == 0 ? last / chargeMacro : last / (8 * j * chargeMacro));   // :496,499
		perRow[j] = n > 0 ? n : 0;
		count[j] = perRow[j] > shard ? (perRow[j] - shard + nShards - 1) / nShards : 0;
		if (nAtRow) nAtRow[j] = perRow[j];
		if (count[j] > 0) {
			PlaceRow pr;
			pr.row = j; pr.pad = 0;
			pr.perRow = perRow[j];
			pr.local = count[j];
			pr.slot0 = 0;                                                   // filled in below, once the layout is known
			pr.id0 = local;
			pr.stream0 = total;
			pr.quantum = last / (double)(perRow[j] + 1);                    // :512
			rows.push_back(pr);
		}
		total += perRow[j];
		local += count[j];
	}
	// [counts-end]
	if (chargeMacroOut) *chargeMacroOut = chargeMacro;
	if (nLoaded) *nLoaded = local;
	PTP_TRY(ptp_plasma_set_layout(p, count, local, mcd));
	if (local == 0) return ptp_build_segments(t, p);
	for (PlaceRow& pr : rows) pr.slot0 = p->rowOff[pr.row];

	// deviate stream: accepted pairs needed = ceil(total / 2); acceptance probability pi / 4
	const long long pairs = (total + 1) / 2;
	long long nAttempts = (long long)std::ceil((double)pairs * 1.2732395447351628 * 1.002) + 4096;
	LoadScratch tmp;
	for (int tries = 0;; ++tries) {
		const long long nThreads = (nAttempts + RNG_CH - 1) / RNG_CH;
		const int nBlocks = (int)((nThreads + 255) / 256);
		PTP_CUDA(cudaMalloc(&tmp.blockCount, (size_t)nBlocks * sizeof(unsigned int)));
		PTP_CUDA(cudaMalloc(&tmp.blockOffset, ((size_t)nBlocks + 1) * sizeof(unsigned long long)));
		k_rng_count<<<nBlocks, 256, 0, t->stream>>>(nAttempts, tmp.blockCount);
		k_rng_scan<<<1, 1024, 0, t->stream>>>(tmp.blockCount, tmp.blockOffset, nBlocks);
		unsigned long long accepted = 0;
		PTP_CUDA(cudaMemcpyAsync(&accepted, tmp.blockOffset + nBlocks, sizeof(accepted), cudaMemcpyDeviceToHost, t->stream));
		PTP_CUDA(cudaStreamSynchronize(t->stream));
		if ((long long)accepted >= pairs) {
			PTP_CUDA(cudaMalloc(&tmp.normals, (size_t)total * sizeof(double)));
			k_rng_emit<<<nBlocks, 256, 0, t->stream>>>(nAttempts, tmp.blockOffset, tmp.normals, total);
			break;
		}
		tmp.dropStream();
		if (tries > 8) { ptp_set_error("ptp_plasma_load_density: deviate stream came up short"); return PTP_ESTATE; }
		nAttempts = nAttempts + nAttempts / 16 + 4096;
	}
	PTP_CUDA(cudaMalloc(&tmp.cum, cum.size() * sizeof(double)));
	PTP_CUDA(cudaMalloc(&tmp.rows, rows.size() * sizeof(PlaceRow)));
	PTP_CUDA(cudaMemcpyAsync(tmp.cum, cum.data(), cum.size() * sizeof(double), cudaMemcpyHostToDevice, t->stream));
	PTP_CUDA(cudaMemcpyAsync(tmp.rows, rows.data(), rows.size() * sizeof(PlaceRow), cudaMemcpyHostToDevice, t->stream));
	long long maxLocal = 0;
	for (const PlaceRow& pr : rows) maxLocal = std::max(maxLocal, pr.local);
	const double sigma = std::sqrt(KB * temperature / p->mass);              // :509
	const dim3 grid((unsigned)std::min<long long>((maxLocal + 255) / 256, 4096), (unsigned)rows.size());
	k_place<<<grid, 256, 0, t->stream>>>(static_cast<const PlaceRow*>(tmp.rows), tmp.cum, tmp.normals, n1, hz, sigma, shard, nShards, p->z, p->v, p->id);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return ptp_cuda_fail(e, "loader launch", __FILE__, __LINE__);
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	t->lastLaunches = 4;
	return ptp_build_segments(t, p);
}
