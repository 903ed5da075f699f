// K3 for large grids (config 5: Nz = 4096, Nr = 1024 -> 4.2 M unknowns, 33.6 MB per fp64 grid).
//
// Same direct solve as ptp_solve.cu - phi = DCT^-1 . Thomas_r . DCT (scale * rho), the matrix of
// PenningTrap::generateSparse (Source/PenningTrap.cpp:94-162) solved as Plasma::solvePoisson does with
// solver.solve (Source/Plasma.cpp:95-99) - but organised for grids whose radial extent no longer fits the fused
// shared-memory kernel of the default grid:
//   k_fwd_dct      forward DCT-I restricted to the touched rows and their touched axial range, as a tiled product
//                  against the forward matrix (64 modes x 64 rows per CTA, 2 x 8 register tile per thread);
//   k_thomas_wide  the radial solves, one warp per 32 axial modes, every coefficient row streamed through a deep
//                  cp.async ring so that the only serial cost is one dependent FMA per row. The deposit is zero above
//                  the plasma's outermost row J, so the rows above J are folded into one precomputed pivot
//                  ("burn at both ends"): a forward sweep over rows J0..J-1 only, x_J from the folded pivot, the usual
//                  back-substitution below J and a pure product chain x_j = r_j x_{j-1} above J. Identical to the plain
//                  Thomas solve in exact arithmetic; 16 B instead of 40 B of traffic per node above J;
//   k_idct_fft_field  inverse DCT-I of a whole grid row through ONE complex FFT of length Nz (the even extension of
//                  the row is real, so its 2 Nz-point transform is obtained from an Nz-point complex transform of the
//                  even/odd samples), all species of the row in one CTA, fused with the node field of
//                  PenningTrap::getEField(int,int) (Source/PenningTrap.cpp:208-236).
#include "ptp_internal.h"

#include <limits.h>

#include <algorithm>
#include <cstdlib>

namespace {

__device__ __forceinline__ void cpa8(void* smemDst, const void* gmemSrc, bool valid)
{
	const unsigned int d = (unsigned int)__cvta_generic_to_shared(smemDst);
	const int bytes = valid ? 8 : 0;                            // src-size 0 -> zero fill
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmemSrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// [wide-begin] (tests/emu/emu_wide.cpp runs the kernels from here to [r16-end] on host threads)
__device__ __forceinline__ int2 row_bounds_of(const int2* bounds, const uint2* enc, int idx, int n1)
{
	if (enc) {                                                  // maxima written by the push kernel's flush: (Nz+2-kmin, kmax+1), 0 = untouched
		const uint2 e = enc[idx];
		return e.y ? make_int2(n1 + 1 - (int)e.x, (int)e.y - 1) : make_int2(INT_MAX, INT_MIN);
	}
	return bounds[idx];
}

// ---- forward DCT-I of the touched rows -------------------------------------------------------------------------
constexpr int FD_M = 64;        // modes per CTA (lane -> modes 2 lane, 2 lane + 1)
constexpr int FD_R = 32;        // rows per CTA (warp -> rows 4 warp .. 4 warp + 3)
constexpr int FD_K = 64;        // axial nodes staged per chunk
constexpr int FD_RP = FD_R + 2; // padded row count of the transposed deposit tile (keeps 16-byte alignment)

// beta[j][m] = scale * sum_k rho[j][k] FT[k][m] over the union of the touched axial ranges of the tile's rows. Chunks of
// 64 nodes are double-buffered through cp.async; per node a lane reads its two matrix entries with one 16-byte load and
// the warp's four deposit values with two 16-byte broadcasts for 8 FMAs.
template <bool A_FIXED>
__global__ void __launch_bounds__(256, 2) k_fwd_dct(const double* __restrict__ rho, const int2* __restrict__ bounds, const uint2* __restrict__ encBounds,
	const double* __restrict__ FT, const double* __restrict__ rowScale, double fixedInv, double* __restrict__ spec, int Nr, int n1)
{
	extern __shared__ __align__(16) double smw[];
	constexpr int BUF = FD_K * FD_M + FD_K * FD_RP;             // doubles per buffer: sFT [FD_K][FD_M] | sRho [FD_K][FD_RP]
	__shared__ int2 sBd[FD_R];
	__shared__ int sLo, sHi;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int mBase = blockIdx.x * FD_M, jBase = blockIdx.y * FD_R, s = blockIdx.z;
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();
	if (tid == 0) { sLo = INT_MAX; sHi = INT_MIN; }
	__syncthreads();
	if (tid < FD_R) {
		int2 bd = make_int2(INT_MAX, INT_MIN);
		if (jBase + tid < Nr) bd = row_bounds_of(bounds, encBounds, s * Nr + jBase + tid, n1);
		sBd[tid] = bd;
		if (bd.x <= bd.y) { atomicMin(&sLo, bd.x); atomicMax(&sHi, bd.y); }
	}
	__syncthreads();
	const int kLo = sLo, kHi = sHi;
	if (kLo > kHi) return;                                      // nothing deposited in these rows
	unsigned int rowsActive = 0;                                // bit jj: row jBase + jj holds a deposit
	for (int jj = 0; jj < FD_R; ++jj) rowsActive |= (sBd[jj].x <= sBd[jj].y ? 1u : 0u) << jj;
	const bool warpActive = ((rowsActive >> (4 * warp)) & 15u) != 0u;
	const double* b = rho + (size_t)s * Nr * n1;
	const int nC = (kHi - kLo + FD_K) / FD_K;
	auto issue = [&](int c) {
		double* sFT = smw + (size_t)(c & 1) * BUF;
		double* sRho = sFT + FD_K * FD_M;
		const int k0 = kLo + c * FD_K, kn = min(FD_K, kHi - k0 + 1);
		for (int e = tid; e < FD_K * FD_M; e += 256) {
			const int kk = e / FD_M, mm = e % FD_M;
			const bool ok = kk < kn && mBase + mm < n1;
			cpa8(&sFT[e], FT + (ok ? (size_t)(k0 + kk) * n1 + mBase + mm : 0), ok);
		}
		// deposit tile, transposed to [node][row]: values outside a row's own range are exact zeros (the grid is cleared
		// every step), rows without a deposit and nodes past the end of the range are zero-filled
		for (int e = tid; e < FD_K * FD_R; e += 256) {
			const int jj = e / FD_K, kk = e % FD_K;
			const bool ok = kk < kn && ((rowsActive >> jj) & 1u);
			cpa8(&sRho[kk * FD_RP + jj], b + (ok ? (size_t)(jBase + jj) * n1 + k0 + kk : 0), ok);
		}
		cpa_commit();
	};
	double acc[4][2];
#pragma unroll
	for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = 0.0;
	issue(0);
	for (int c = 0; c < nC; ++c) {
		if (c + 1 < nC) { issue(c + 1); cpa_wait<1>(); }
		else cpa_wait<0>();
		double* sFT = smw + (size_t)(c & 1) * BUF;
		double* sRho = sFT + FD_K * FD_M;
		if (A_FIXED) {                                          // int64 accumulators -> double, each thread its own copies
			for (int e = tid; e < FD_K * FD_R; e += 256) {
				double* q = &sRho[(e % FD_K) * FD_RP + e / FD_K];
				*q = (double)__double_as_longlong(*q);
			}
		}
		__syncthreads();
		if (warpActive) {
#pragma unroll 4
			for (int kk = 0; kk < FD_K; ++kk) {
				const double2 f = *reinterpret_cast<const double2*>(&sFT[kk * FD_M + 2 * lane]);
				const double2 r01 = *reinterpret_cast<const double2*>(&sRho[kk * FD_RP + 4 * warp]);
				const double2 r23 = *reinterpret_cast<const double2*>(&sRho[kk * FD_RP + 4 * warp + 2]);
				acc[0][0] = fma(r01.x, f.x, acc[0][0]); acc[0][1] = fma(r01.x, f.y, acc[0][1]);
				acc[1][0] = fma(r01.y, f.x, acc[1][0]); acc[1][1] = fma(r01.y, f.y, acc[1][1]);
				acc[2][0] = fma(r23.x, f.x, acc[2][0]); acc[2][1] = fma(r23.x, f.y, acc[2][1]);
				acc[3][0] = fma(r23.y, f.x, acc[3][0]); acc[3][1] = fma(r23.y, f.y, acc[3][1]);
			}
		}
		__syncthreads();                                        // this buffer is refilled by the next iteration's issue
	}
	const double scale = (rowScale ? rowScale[s] : 1.0) * (A_FIXED ? fixedInv : 1.0);
	double* out = spec + (size_t)s * Nr * n1;
#pragma unroll
	for (int u = 0; u < 4; ++u) {
		const int jj = 4 * warp + u;
		if (!((rowsActive >> jj) & 1u)) continue;
		const size_t row = (size_t)(jBase + jj) * n1;
		if (mBase + 2 * lane < n1) out[row + mBase + 2 * lane] = acc[u][0] * scale;
		if (mBase + 2 * lane + 1 < n1) out[row + mBase + 2 * lane + 1] = acc[u][1] * scale;
	}
}

// ---- radial solves -----------------------------------------------------------------------------------------------
constexpr int TW_RS = 8;        // rows per ring stage
constexpr int TW_ST = 16;       // stages: 15 x 4 KB in flight per warp
constexpr int TW_BLK = PTP_THOMAS_BLOCK;   // rows per block of the product tables (thP)

// Visit rows jFirst, jFirst + DIR, ... (count rows) of one or two [Nr][n1] arrays for the 32 modes of this warp, each lane
// copying and later reading only its own mode (no cross-lane hazards, so no barriers). a1 is read only where flag1 says
// so (zero elsewhere). use(j, v0, v1) is called in row order - that is where the serial recurrence lives.
template <int DIR, bool TWO, class Use>
__device__ __forceinline__ void stream_rows(double* ring, int jFirst, int count, const double* __restrict__ a0, const double* __restrict__ a1,
	const unsigned char* flag1, int n1, int m, bool mOk, int lane, Use use)
{
	const int nG = (count + TW_RS - 1) / TW_RS;
	auto issue = [&](int g) {
		if (g < nG) {
			double* dst = ring + (size_t)(g % TW_ST) * 2 * TW_RS * 32 + lane;
#pragma unroll
			for (int r = 0; r < TW_RS; ++r) {
				const int i = g * TW_RS + r;
				if (i < count) {
					const int j = jFirst + DIR * i;
					const size_t off = (size_t)j * n1 + (mOk ? m : 0);
					cpa8(dst + r * 32, a0 + off, mOk);
					if (TWO) cpa8(dst + (TW_RS + r) * 32, a1 + off, mOk && (!flag1 || flag1[j]));
				}
			}
		}
		cpa_commit();
	};
#pragma unroll 1
	for (int g = 0; g < TW_ST - 1; ++g) issue(g);
#pragma unroll 1
	for (int g = 0; g < nG; ++g) {
		issue(g + TW_ST - 1);
		cpa_wait<TW_ST - 1>();
		const double* src = ring + (size_t)(g % TW_ST) * 2 * TW_RS * 32 + lane;
		double v0[TW_RS], v1[TW_RS];
#pragma unroll
		for (int r = 0; r < TW_RS; ++r) {
			v0[r] = src[r * 32];
			v1[r] = TWO ? src[(TW_RS + r) * 32] : 0.0;
		}
#pragma unroll
		for (int r = 0; r < TW_RS; ++r) {
			const int i = g * TW_RS + r;
			if (i < count) use(jFirst + DIR * i, v0[r], v1[r]);
		}
	}
	cpa_wait<0>();
}

// spec holds beta (forward-transformed deposit) on the touched rows on entry and alpha = (T_r + lambda_m)^-1 beta on
// all rows on exit (rows above the block of the outermost deposit row J are left to k_thomas_expand). 32 modes per CTA
// (lane = mode), blockIdx.y = species.
//   compact path (rows 0 .. end of J's block fit in shared memory - the usual case, a plasma near the axis): the four warps
//   bring every coefficient of those rows in with one wave of cp.async, warp 0 runs the three short recurrences out of
//   shared memory (7 instructions per row instead of ~40 for the streamed form), all warps write the rows back;
//   streamed path (deposit reaching far out, e.g. the wall RHS of solveLaplace): warp 0 walks the rows through the ring.
constexpr int TW_RCAP = 160;    // rows of the compact path: 3 x 160 x 256 B = 120 KB

__global__ void __launch_bounds__(128) k_thomas_wide(double* __restrict__ specAll, const int2* __restrict__ bounds, const uint2* __restrict__ encBounds,
	const double* __restrict__ thInv, const double* __restrict__ thCp, const double* __restrict__ thR, const double* __restrict__ thQ,
	const double* __restrict__ thP, const double* __restrict__ thLower, double* __restrict__ xbAll, int* __restrict__ wideJ, int Nr, int n1, int rowsOut)
{
	extern __shared__ __align__(16) double smw[];
	double* ring = smw;                                         // streamed: [TW_ST][2][TW_RS][32]; compact: sInv | sB | sC, [TW_RCAP][32] each
	double* sLower = smw + (size_t)3 * TW_RCAP * 32;            // [Nr]
	unsigned char* sTouched = reinterpret_cast<unsigned char*>(sLower + Nr); // [Nr]
	__shared__ int sJ0, sJ;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, s = blockIdx.y;
	const int m = blockIdx.x * 32 + lane;
	const bool mOk = m < n1;
	double* spec = specAll + (size_t)s * Nr * n1;
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();
	if (tid == 0) { sJ0 = INT_MAX; sJ = INT_MIN; }
	__syncthreads();
	{
		int lo = INT_MAX, hi = INT_MIN;
		for (int j = tid; j < Nr; j += 128) {
			const int2 bd = row_bounds_of(bounds, encBounds, s * Nr + j, n1);
			const bool touched = bd.x <= bd.y;
			sTouched[j] = touched ? 1 : 0;
			sLower[j] = thLower[j];
			if (touched) { lo = min(lo, j); hi = max(hi, j); }
		}
		if (lo <= hi) { atomicMin(&sJ0, lo); atomicMax(&sJ, hi); }
	}
	__syncthreads();
	const int J0 = sJ0, J = sJ;
	if (J < 0) {                                                // empty deposit: the potential is zero
		if (mOk) for (int j = warp; j < Nr; j += 4) spec[(size_t)j * n1 + m] = 0.0;
		if (blockIdx.x == 0 && tid == 0) wideJ[s] = -1;
		return;
	}
	// rows above J: homogeneous recurrence towards the wall, x_j = r_j x_{j-1}. Only the rest of J's block of TW_BLK rows is
	// walked here; for the blocks above, the value entering each block is propagated with the precomputed block products
	// (thP at a block's last row) and k_thomas_expand fills the rows in parallel from thP's in-block prefix products.
	const int jEnd = min(Nr - 1, (J / TW_BLK) * TW_BLK + TW_BLK - 1);
	double x = 0.0;                                             // warp 0: alpha at row jEnd
	if (jEnd < TW_RCAP) {
		double* sInv = smw;
		double* sB = smw + TW_RCAP * 32;
		double* sC = smw + 2 * TW_RCAP * 32;
		const double q = (warp == 0 && mOk) ? thQ[(size_t)J * n1 + m] : 0.0;
		for (int j = warp; j <= jEnd; j += 4) {
			const size_t off = (size_t)j * n1 + (mOk ? m : 0);
			if (j <= J) {
				cpa8(&sB[j * 32 + lane], spec + off, mOk && sTouched[j]);
				if (j >= J0 && j < J) cpa8(&sInv[j * 32 + lane], thInv + off, mOk);
				if (j < J) cpa8(&sC[j * 32 + lane], thCp + off, mOk);
			}
			else cpa8(&sC[j * 32 + lane], thR + off, mOk);
		}
		cpa_commit();
		cpa_wait<0>();
		__syncthreads();
		if (warp == 0) {
			// forward sweep over rows J0 .. J-1:  y_j = (beta_j - l_j y_{j-1}) / pivot_j
			double y = 0.0;
#pragma unroll 4
			for (int j = J0; j < J; ++j) {
				const double inv = sInv[j * 32 + lane];
				const double g = sB[j * 32 + lane] * inv, c = -(sLower[j] * inv);
				y = fma(c, y, g);
				sB[j * 32 + lane] = y;
			}
			// row J closes the system: the rows above it carry no deposit and are folded into the pivot 1 / thQ
			const double xJ = (sB[J * 32 + lane] - sLower[J] * y) * q;
			sB[J * 32 + lane] = xJ;
			// back-substitution below J:  x_j = y_j - cp_j x_{j+1}   (y_j = 0 below the first touched row: sB was zero-filled)
			x = xJ;
#pragma unroll 4
			for (int j = J - 1; j >= 0; --j) {
				x = fma(-sC[j * 32 + lane], x, sB[j * 32 + lane]);
				sB[j * 32 + lane] = x;
			}
			x = xJ;
#pragma unroll 4
			for (int j = J + 1; j <= jEnd; ++j) {
				x = sC[j * 32 + lane] * x;
				sB[j * 32 + lane] = x;
			}
		}
		__syncthreads();
		if (mOk) for (int j = warp; j <= jEnd; j += 4) spec[(size_t)j * n1 + m] = sB[j * 32 + lane];
	}
	else if (warp == 0) {
		double y = 0.0;
		stream_rows<1, true>(ring, J0, J - J0, thInv, spec, sTouched, n1, m, mOk, lane, [&](int j, double inv, double beta) {
			const double g = beta * inv, c = -(sLower[j] * inv);
			y = fma(c, y, g);
			if (mOk) spec[(size_t)j * n1 + m] = y;
		});
		double xJ = 0.0;
		if (mOk) {
			const size_t o = (size_t)J * n1 + m;
			xJ = (spec[o] - sLower[J] * y) * thQ[o];
			spec[o] = xJ;
		}
		__threadfence_block();                                  // this lane's y values are read back through cp.async below
		x = xJ;
		__syncwarp();                                           // every lane is done reading the flags in their first meaning
		for (int j = J0 + lane; j < J; j += 32) sTouched[j] = 1;    // from here on the flag means "row >= J0": y was stored there
		__syncwarp();
		stream_rows<-1, true>(ring, J - 1, J, thCp, spec, sTouched, n1, m, mOk, lane, [&](int j, double cp, double yj) {
			x = fma(-cp, x, yj);
			if (mOk) spec[(size_t)j * n1 + m] = x;
		});
		x = xJ;
		stream_rows<1, false>(ring, J + 1, jEnd - J, thR, nullptr, nullptr, n1, m, mOk, lane, [&](int j, double r, double) {
			x = r * x;
			if (mOk) spec[(size_t)j * n1 + m] = x;
		});
	}
	if (warp != 0) return;
	const int nB = (Nr + TW_BLK - 1) / TW_BLK, bFirst = J / TW_BLK + 1;
	double* xb = xbAll + (size_t)s * nB * n1;
	if (blockIdx.x == 0 && lane == 0) wideJ[s] = J;
	const int nBOut = min(nB, (rowsOut + TW_BLK - 1) / TW_BLK);  // blocks the caller wants (the step: the populated rows only)
	for (int b0 = bFirst; b0 < nBOut; b0 += 16) {
		double pb[16];
#pragma unroll
		for (int u = 0; u < 16; ++u) {
			const int b = b0 + u;
			pb[u] = (b < nBOut && mOk) ? thP[(size_t)min(Nr - 1, b * TW_BLK + TW_BLK - 1) * n1 + m] : 0.0;
		}
#pragma unroll
		for (int u = 0; u < 16; ++u) {
			const int b = b0 + u;
			if (b < nBOut && mOk) xb[(size_t)b * n1 + m] = x;
			x = pb[u] * x;
		}
	}
}

// Rows above the block of the outermost deposit row: alpha[j][m] = (value entering the block) * (in-block prefix product).
__global__ void __launch_bounds__(256) k_thomas_expand(double* __restrict__ specAll, const double* __restrict__ xbAll, const int* __restrict__ wideJ,
	const double* __restrict__ thP, int Nr, int n1)
{
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();
	const int s = blockIdx.z, J = wideJ[s];
	const int j0 = blockIdx.y * 8;
	if (J < 0 || j0 <= min(Nr - 1, (J / TW_BLK) * TW_BLK + TW_BLK - 1)) return;   // TW_BLK is a multiple of 8: the whole group is on one side
	const int m = blockIdx.x * 256 + threadIdx.x;
	if (m >= n1) return;
	const int nB = (Nr + TW_BLK - 1) / TW_BLK;
	const double xin = xbAll[((size_t)s * nB + j0 / TW_BLK) * n1 + m];
	double* spec = specAll + (size_t)s * Nr * n1;
	double p[8];
#pragma unroll
	for (int u = 0; u < 8; ++u) p[u] = j0 + u < Nr ? thP[(size_t)(j0 + u) * n1 + m] : 0.0;
#pragma unroll
	for (int u = 0; u < 8; ++u)
		if (j0 + u < Nr) spec[(size_t)(j0 + u) * n1 + m] = xin * p[u];
}

// ---- inverse DCT-I through an Nz-point complex FFT + node field ---------------------------------------------
// phi_k = sum_{m=0}^{N} a_m cos(pi m k / N) = X_k / 2 + (a_0 + (-1)^k a_N) / 2, where X is the 2N-point DFT of the even
// extension e_n = a_n (n <= N), a_{2N-n} (n > N). With z_n = e_{2n} + i e_{2n+1} and Z = DFT_N(z):
//   X_k = (Z_k + conj Z_{N-k}) / 2 + W_{2N}^k (Z_k - conj Z_{N-k}) / (2 i),   W_{2N} = exp(-i pi / N),
// real for a symmetric e. The N-point transform runs in shared memory as radix-2 decimation-in-frequency passes fused
// in pairs (four points per thread and pair of passes, same arithmetic as two radix-2 passes, half the barriers and half
// the shared-memory traffic); Z_k is read from the bit-reversed slot. One CTA per grid row loops over the species.
template <bool FIELD>
__global__ void __launch_bounds__(256, 2) k_idct_fft_field(const double* __restrict__ alphaAll, double* __restrict__ phiAll, const double2* __restrict__ tw,
	const double* __restrict__ phiTrap, double* __restrict__ eNodes, int nS, int Nr, int N, int bits /* log2 N */, double hz)
{
	extern __shared__ double2 fbw[];                            // [N], swizzled
	double* tot = reinterpret_cast<double*>(fbw + N);           // [N+1] running total potential of the row (FIELD)
	const int tid = threadIdx.x, T = blockDim.x, n1 = N + 1;
	const int row = blockIdx.x;
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();
	// Element p lives at p ^ f(p >> 3). 16-byte accesses are served a quarter-warp at a time, so the 8 lanes of a quarter
	// must hit the 8 distinct 16-byte bank groups (p & 7). The late passes touch q-strided points with q < 8, which moves the
	// lane index into bits 3..5 of p, and the bit-reversed read-out moves it into the top three bits: both fields are folded
	// into the group (bits 3, 4 twice so that every small stride stays a bijection on the quarter-warp).
	const int sh = bits >= 9 ? bits - 3 : 31;
	auto SW = [sh](int p) { const int x = p >> 3; return p ^ ((x ^ ((x & 3) << 1) ^ (p >> sh)) & 7); };
	if (FIELD) for (int k = tid; k <= N; k += T) tot[k] = phiTrap[(size_t)row * n1 + k];
	for (int sp = 0; sp < nS; ++sp) {
		const double* a = alphaAll + ((size_t)sp * Nr + row) * n1;
		__syncthreads();                                        // previous species is done with fbw
		for (int n = tid; n < N; n += T) {
			const int i0 = 2 * n, i1 = 2 * n + 1;
			fbw[SW(n)] = make_double2(a[i0 <= N ? i0 : 2 * N - i0], a[i1 <= N ? i1 : 2 * N - i1]);
		}
		const double a0 = a[0], aN = a[N];
		__syncthreads();
		int h = N >> 1;
		for (; h >= 2; h >>= 2) {                               // passes with half-sizes h and h/2, fused
			const int q = h >> 1, sA = N / h;                   // twiddle steps: W_{2h}^j = tw[j * N / h]
			for (int i0 = tid; i0 < (N >> 2); i0 += 2 * T) {    // two items per thread in flight
				int p[2][4];
				double2 x[2][4], wa[2], wb[2];
#pragma unroll
				for (int u = 0; u < 2; ++u) {
					const int i = min(i0 + u * T, (N >> 2) - 1);
					const int j0 = i & (q - 1);
					const int b = ((i - j0) << 2) + j0;
					p[u][0] = SW(b); p[u][1] = SW(b + q); p[u][2] = SW(b + h); p[u][3] = SW(b + h + q);
					// W_{2h}^{j0 + h/2} = -i W_{2h}^{j0}; the second pass' twiddle W_{h}^{j0} comes from the table (exactly rounded)
					wa[u] = __ldg(&tw[j0 * sA]);
					wb[u] = __ldg(&tw[j0 * 2 * sA]);
#pragma unroll
					for (int c = 0; c < 4; ++c) x[u][c] = (i0 + u * T < (N >> 2)) ? fbw[p[u][c]] : make_double2(0.0, 0.0);   // (a clamped item is another thread's: not even read)
				}
#pragma unroll
				for (int u = 0; u < 2; ++u) {
					if (i0 + u * T >= (N >> 2)) continue;
					const double2 x0 = x[u][0], x1 = x[u][1], x2 = x[u][2], x3 = x[u][3];
					const double2 u0 = make_double2(x0.x + x2.x, x0.y + x2.y), u1 = make_double2(x1.x + x3.x, x1.y + x3.y);
					const double d2x = x0.x - x2.x, d2y = x0.y - x2.y, d3x = x1.x - x3.x, d3y = x1.y - x3.y;
					const double2 u2 = make_double2(d2x * wa[u].x - d2y * wa[u].y, d2x * wa[u].y + d2y * wa[u].x);
					const double2 u3 = make_double2(d3x * wa[u].y + d3y * wa[u].x, d3y * wa[u].y - d3x * wa[u].x);   // (d3x + i d3y) (wa.y - i wa.x)
					fbw[p[u][0]] = make_double2(u0.x + u1.x, u0.y + u1.y);
					const double e1x = u0.x - u1.x, e1y = u0.y - u1.y;
					fbw[p[u][1]] = make_double2(e1x * wb[u].x - e1y * wb[u].y, e1x * wb[u].y + e1y * wb[u].x);
					fbw[p[u][2]] = make_double2(u2.x + u3.x, u2.y + u3.y);
					const double e3x = u2.x - u3.x, e3y = u2.y - u3.y;
					fbw[p[u][3]] = make_double2(e3x * wb[u].x - e3y * wb[u].y, e3x * wb[u].y + e3y * wb[u].x);
				}
			}
			__syncthreads();
		}
		if (h == 1) {                                           // odd number of passes: the last one on its own (twiddle 1)
			for (int i = tid; i < (N >> 1); i += T) {
				const int p0 = SW(2 * i), p1 = SW(2 * i + 1);
				const double2 u = fbw[p0], v = fbw[p1];
				fbw[p0] = make_double2(u.x + v.x, u.y + v.y);
				fbw[p1] = make_double2(u.x - v.x, u.y - v.y);
			}
			__syncthreads();
		}
		double* out = phiAll + ((size_t)sp * Nr + row) * n1;
		for (int k = tid; k <= N; k += T) {
			const unsigned int ka = (unsigned int)(k & (N - 1)), kb = (unsigned int)((N - k) & (N - 1));
			const double2 A = fbw[SW((int)(__brev(ka) >> (32 - bits)))], B = fbw[SW((int)(__brev(kb) >> (32 - bits)))];   // Z_k, Z_{N-k}
			const double2 w = k < N ? __ldg(&tw[k]) : make_double2(-1.0, 0.0);
			const double dx = A.x - B.x, dy = A.y + B.y;            // Z_k - conj Z_{N-k}
			const double X = 0.5 * (A.x + B.x) + 0.5 * (dy * w.x + dx * w.y);
			const double v = 0.5 * X + 0.5 * (a0 + ((k & 1) ? -aN : aN));
			out[k] = v;
			if (FIELD) tot[k] = __dadd_rn(tot[k], v);
		}
	}
	if (FIELD) {
		__syncthreads();
		for (int k = tid; k <= N; k += T) {
			double e = 0.0;
			if (k > 0 && k < N) e = __ddiv_rn(__dsub_rn(tot[k - 1], tot[k + 1]), __dmul_rn(2.0, hz));
			eNodes[(size_t)row * n1 + k] = e;
		}
	}
}


// ---- the same transform for N = 4096 in three register-resident radix-16 stages ------------------------------------
// 4096 = 16^3: with n = 256 n2 + 16 n1 + n0 and k = k0 + 16 k1 + 256 k2 the N-point transform is three rounds of 256
// independent 16-point transforms (over n2, n1, n0), each done by one thread in registers, with a twiddle W_N^{(16 n1 + n0) k0}
// after the first round and W_256^{n0 k1} after the second. The rounds exchange their data through a [16][257] buffer
// (row stride 257 sixteen-byte elements: every access of every round is conflict-free for the quarter-warps that serve
// 16-byte accesses). Shared-memory traffic per row and species: 6 x 64 KB (three writes, three reads) against 15 x 64 KB for
// the radix-2 pass pairs, and five barriers against eight. The read-out forms phi_k and phi_{N-k} from the same pair
// (Z_k, Z_{N-k}). tools/fft16_model.py is a thread-level model of this kernel (index maps, twiddles, bank groups).
// [r16-begin] (tests/emu/emu_r16.sh compiles the text between these markers for the host and runs it on 256 CPU threads)
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// 4-point transform, exp(-2 pi i j k / 4), in place
__device__ __forceinline__ void fft4(double2& a0, double2& a1, double2& a2, double2& a3)
{
	const double2 s0 = make_double2(a0.x + a2.x, a0.y + a2.y), s1 = make_double2(a0.x - a2.x, a0.y - a2.y);
	const double2 s2 = make_double2(a1.x + a3.x, a1.y + a3.y), s3 = make_double2(a1.x - a3.x, a1.y - a3.y);
	a0 = make_double2(s0.x + s2.x, s0.y + s2.y);
	a2 = make_double2(s0.x - s2.x, s0.y - s2.y);
	a1 = make_double2(s1.x + s3.y, s1.y - s3.x);                // s1 - i s3
	a3 = make_double2(s1.x - s3.y, s1.y + s3.x);                // s1 + i s3
}

// y[k] = sum_j x[j] exp(-2 pi i j k / 16), natural order in and out: j = 4 j1 + j0, k = k0 + 4 k1
__device__ __forceinline__ void fft16(double2 (&x)[16])
{
	constexpr double C8 = 0.92387953251128675613, S8 = 0.38268343236508977173, R2 = 0.70710678118654752440;
	double2 t[16];                                              // t[4 j0 + k0]
#pragma unroll
	for (int j0 = 0; j0 < 4; ++j0) {
		t[4 * j0] = x[j0]; t[4 * j0 + 1] = x[4 + j0]; t[4 * j0 + 2] = x[8 + j0]; t[4 * j0 + 3] = x[12 + j0];
		fft4(t[4 * j0], t[4 * j0 + 1], t[4 * j0 + 2], t[4 * j0 + 3]);
	}
	const double2 W1 = make_double2(C8, -S8), W3 = make_double2(S8, -C8);
	t[5] = cmul(t[5], W1);                                                      // W_16^{j0 k0}
	t[6] = make_double2(R2 * (t[6].x + t[6].y), R2 * (t[6].y - t[6].x));        // W^2 = (1 - i) / sqrt 2
	t[7] = cmul(t[7], W3);
	t[9] = make_double2(R2 * (t[9].x + t[9].y), R2 * (t[9].y - t[9].x));        // W^2
	t[10] = make_double2(t[10].y, -t[10].x);                                    // W^4 = -i
	t[11] = make_double2(R2 * (t[11].y - t[11].x), -R2 * (t[11].x + t[11].y));  // W^6 = -(1 + i) / sqrt 2
	t[13] = cmul(t[13], W3);
	t[14] = make_double2(R2 * (t[14].y - t[14].x), -R2 * (t[14].x + t[14].y));  // W^6
	{ const double2 m = cmul(t[15], W1); t[15] = make_double2(-m.x, -m.y); }    // W^9 = -W^1
#pragma unroll
	for (int k0 = 0; k0 < 4; ++k0) {
		fft4(t[k0], t[4 + k0], t[8 + k0], t[12 + k0]);
		x[k0] = t[k0]; x[k0 + 4] = t[4 + k0]; x[k0 + 8] = t[8 + k0]; x[k0 + 12] = t[12 + k0];
	}
}

// x[k] *= w1^k, the powers by binary splitting (at most four products deep)
__device__ __forceinline__ void twiddle16(double2 (&x)[16], double2 w1)
{
	double2 w[8];
	w[1] = w1;
	w[2] = cmul(w1, w1);
	w[3] = cmul(w[2], w1);
	w[4] = cmul(w[2], w[2]);
	w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]);
	const double2 w8 = cmul(w[4], w[4]);
#pragma unroll
	for (int k = 1; k < 8; ++k) x[k] = cmul(x[k], w[k]);
	x[8] = cmul(x[8], w8);
#pragma unroll
	for (int k = 1; k < 8; ++k) x[8 + k] = cmul(x[8 + k], cmul(w8, w[k]));
}

constexpr int R16_N = 4096, R16_RS = 257;

// Rows above the block of the outermost deposit row can be formed on the fly instead of being read from alphaAll (xbAll != nullptr):
// alpha[j][m] = (value entering the row's block, xbAll) * (in-block prefix product, thP) - what k_thomas_expand would have written.
template <bool FIELD>
__global__ void __launch_bounds__(256, 2) k_idct_r16_field(const double* __restrict__ alphaAll, double* __restrict__ phiAll, const double2* __restrict__ tw,
	const double* __restrict__ phiTrap, double* __restrict__ eNodes, int nS, int Nr, double hz,
	const double* __restrict__ xbAll, const int* __restrict__ wideJ, const double* __restrict__ thP, int blockRows)
{
	constexpr int N = R16_N, n1 = N + 1, RS = R16_RS;
	extern __shared__ double2 fbw[];                            // [16][RS] exchange buffer; natural-order Z after the third round
	double* tot = reinterpret_cast<double*>(fbw + 16 * RS);     // [N+1] running total potential of the row (FIELD)
	const int t = threadIdx.x, row = blockIdx.x;
	ptp_pdl_launch_dependents();
	const double2 wA = __ldg(&tw[2 * t]);                       // W_N^t      (tw[j] = exp(-i pi j / N))   - constants, before the wait
	const double2 wB = __ldg(&tw[32 * (t & 15)]);               // W_256^n0
	ptp_pdl_wait();
	if (FIELD) {                                                // asynchronous: needed only at the first read-out
		for (int k = t; k <= N; k += 256) cpa8(&tot[k], phiTrap + (size_t)row * n1 + k, true);
		cpa_commit();
	}
	for (int sp = 0; sp < nS; ++sp) {
		const double* a = alphaAll + ((size_t)sp * Nr + row) * n1;
		double2 x[16];
		double a0, aN;
		bool formed = false;                                    // uniform over the CTA
		if (xbAll) {
			const int J = wideJ[sp];
			formed = J >= 0 && row > min(Nr - 1, (J / blockRows) * blockRows + blockRows - 1);
		}
		// round 1 (over n2): thread t = 16 n1 + n0 owns the samples n = t + 256 j of z_n = e_2n + i e_2n+1 (even extension of a)
		if (!formed) {
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				const int i0 = 2 * (t + 256 * j), i1 = i0 + 1;
				x[j] = make_double2(a[i0 <= N ? i0 : 2 * N - i0], a[i1 <= N ? i1 : 2 * N - i1]);
			}
			a0 = a[0]; aN = a[N];
		}
		else {
			const int nB = (Nr + blockRows - 1) / blockRows;
			const double* xin = xbAll + ((size_t)sp * nB + row / blockRows) * n1;
			const double* pp = thP + (size_t)row * n1;
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				int i0 = 2 * (t + 256 * j), i1 = i0 + 1;
				i0 = i0 <= N ? i0 : 2 * N - i0;
				i1 = i1 <= N ? i1 : 2 * N - i1;
				x[j] = make_double2(xin[i0] * pp[i0], xin[i1] * pp[i1]);
			}
			a0 = xin[0] * pp[0]; aN = xin[N] * pp[N];
		}
		fft16(x);
		twiddle16(x, wA);
		if (FIELD && sp == 0) cpa_wait<0>();                    // this thread's part of tot has landed; the barriers below publish it
		__syncthreads();                                        // the previous species' read-out is done with fbw
#pragma unroll
		for (int k0 = 0; k0 < 16; ++k0) fbw[(t >> 4) * RS + k0 * 16 + (t & 15)] = x[k0];
		__syncthreads();
		// round 2 (over n1): thread t = 16 k0 + n0
#pragma unroll
		for (int j = 0; j < 16; ++j) x[j] = fbw[j * RS + t];
		fft16(x);
		twiddle16(x, wB);
		__syncthreads();                                        // every read of this round precedes its writes (other layout)
#pragma unroll
		for (int k1 = 0; k1 < 16; ++k1) fbw[(t & 15) * RS + (t >> 4) + 16 * k1] = x[k1];
		__syncthreads();
		// round 3 (over n0): thread t = k0 + 16 k1 ends up with Z[t + 256 k2]
#pragma unroll
		for (int j = 0; j < 16; ++j) x[j] = fbw[j * RS + t];
		fft16(x);
		__syncthreads();
#pragma unroll
		for (int k2 = 0; k2 < 16; ++k2) fbw[t + 256 * k2] = x[k2];
		__syncthreads();
		// read-out: phi_k and phi_{N-k} from the pair (Z_k, Z_{N-k});  W_2N^{N-k} = -conj W_2N^k
		double* out = phiAll + ((size_t)sp * Nr + row) * n1;
		const double half0 = 0.5 * (a0 + aN), half1 = 0.5 * (a0 - aN);
		auto emit = [&](int k, double2 A, double2 B, double2 w) {
			const double dx = A.x - B.x, dy = A.y + B.y;            // Z_k - conj Z_{N-k}
			const double X = 0.5 * (A.x + B.x) + 0.5 * (dy * w.x + dx * w.y);
			const double v = 0.5 * X + ((k & 1) ? half1 : half0);
			out[k] = v;
			if (FIELD) tot[k] = __dadd_rn(tot[k], v);
		};
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int k = t + 256 * j;
			const double2 A = fbw[k], B = fbw[(N - k) & (N - 1)];
			const double2 w = __ldg(&tw[k]);
			emit(k, A, B, w);
			emit(N - k, B, A, make_double2(-w.x, w.y));
		}
		if (t == 0) {
			const double2 A = fbw[N / 2];
			emit(N / 2, A, A, __ldg(&tw[N / 2]));
		}
	}
	if (FIELD) {
		__syncthreads();
		for (int k = t; k <= N; k += 256) {
			double e = 0.0;
			if (k > 0 && k < N) e = __ddiv_rn(__dsub_rn(tot[k - 1], tot[k + 1]), __dmul_rn(2.0, hz));
			eNodes[(size_t)row * n1 + k] = e;
		}
	}
}
// [r16-end]

} // namespace

bool ptp_solver_fft_fits(const ptp_trap* t)
{
	return t->fftTw && (size_t)t->Nz * sizeof(double2) + (size_t)(t->Nz + 1) * sizeof(double) <= t->smemMax;
}

// beta -> alpha for nS grids: forward transform of the touched rows + radial solves. bounds / encBounds as in ptp_solver_run.
int ptp_solver_forward_wide(ptp_trap* t, const double* rho, bool rhoIsFixed, const double* dScale, int nS, double* spec, const uint2* encBounds, bool expand, int rowLimit, int rowsOut)
{
	const int n1 = t->Nz + 1, Nr = t->Nr;
	if (rowsOut <= 0 || rowsOut > Nr) rowsOut = Nr;
	const int rowsIn = rowLimit < 0 ? Nr : std::max(1, std::min(rowLimit, Nr));   // rows that can hold a deposit
	const double fixedInv = 1.0 / (double)(1ULL << t->fixedBits);
	const size_t smDct = (size_t)2 * (FD_K * FD_M + FD_K * FD_RP) * sizeof(double);
	const dim3 gridDct((n1 + FD_M - 1) / FD_M, (rowsIn + FD_R - 1) / FD_R, nS);
	if (rhoIsFixed) {
		PTP_CUDA(cudaFuncSetAttribute(k_fwd_dct<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smDct));
		const cudaError_t ed = ptp_launch(k_fwd_dct<true>, gridDct, dim3(256), smDct, t->stream, t->usePdl, rho, t->rowBounds, encBounds, t->dctFwd, dScale, fixedInv, spec, Nr, n1);
		if (ed != cudaSuccess) return ptp_cuda_fail(ed, "k_fwd_dct launch", __FILE__, __LINE__);
	}
	else {
		PTP_CUDA(cudaFuncSetAttribute(k_fwd_dct<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smDct));
		const cudaError_t ed = ptp_launch(k_fwd_dct<false>, gridDct, dim3(256), smDct, t->stream, t->usePdl, rho, t->rowBounds, encBounds, t->dctFwd, dScale, 1.0, spec, Nr, n1);
		if (ed != cudaSuccess) return ptp_cuda_fail(ed, "k_fwd_dct launch", __FILE__, __LINE__);
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_fwd_dct launch", __FILE__, __LINE__);
	static_assert(3 * TW_RCAP >= TW_ST * 2 * TW_RS, "the ring of the streamed path lives in the compact path's tiles");
	const size_t smTh = (size_t)3 * TW_RCAP * 32 * sizeof(double) + (size_t)Nr * sizeof(double) + (size_t)Nr;
	if (smTh > t->smemMax) { ptp_set_error("direct solver: Nr too large for the radial-solve kernel of this build"); return PTP_EINVAL; }
	PTP_CUDA(cudaFuncSetAttribute(k_thomas_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smTh));
	e = ptp_launch(k_thomas_wide, dim3((n1 + 31) / 32, nS), dim3(128), smTh, t->stream, t->usePdl, spec, t->rowBounds, encBounds, t->thInv, t->thCp, t->thR, t->thQ, t->thP, t->thLower,
		t->wideXb, t->wideJ, Nr, n1, rowsOut);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_thomas_wide launch", __FILE__, __LINE__);
	t->lastLaunches += 2;
	if (!expand) return PTP_OK;                                 // the inverse transform forms the rows above the deposit itself
	e = ptp_launch(k_thomas_expand, dim3((n1 + 255) / 256, (rowsOut + 7) / 8, nS), dim3(256), 0, t->stream, t->usePdl, spec, t->wideXb, t->wideJ, t->thP, Nr, n1);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_thomas_expand launch", __FILE__, __LINE__);
	t->lastLaunches += 1;
	return PTP_OK;
}

// alpha -> phi for nS grids (+ node field when withField: nS covers all species and phi = phiSelfAll).
// Does the inverse transform of this trap form the rows above the deposit itself (no k_thomas_expand needed)?
bool ptp_solver_inverse_forms_rows(const ptp_trap* t)
{
	return t->Nz == R16_N && t->fftR16 && t->fftFormRows;
}

int ptp_solver_inverse_fft(ptp_trap* t, const double* spec, double* phi, int nS, bool withField, bool rowsFormed, int rowsOut)
{
	const int N = t->Nz, Nr = t->Nr;
	if (rowsOut <= 0 || rowsOut > Nr) rowsOut = Nr;
	int bits = 0;
	while ((1 << bits) < N) ++bits;
	// N = 4096: three radix-16 rounds in registers (PTP_FFT_R16=0 at trap creation selects the radix-2 pass pairs instead)
	if (N == R16_N && t->fftR16) {
		const size_t sm16 = (size_t)16 * R16_RS * sizeof(double2) + (size_t)(N + 1) * sizeof(double);
		if (withField) {
			PTP_CUDA(cudaFuncSetAttribute(k_idct_r16_field<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm16));
			const cudaError_t el = ptp_launch(k_idct_r16_field<true>, dim3(rowsOut), dim3(256), sm16, t->stream, t->usePdl, spec, phi, t->fftTw, t->phiTrap, t->eNodes, nS, Nr, t->hz,
				rowsFormed ? nullptr : t->wideXb, t->wideJ, t->thP, TW_BLK);
			if (el != cudaSuccess) return ptp_cuda_fail(el, "k_idct_r16_field launch", __FILE__, __LINE__);
		}
		else {
			PTP_CUDA(cudaFuncSetAttribute(k_idct_r16_field<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm16));
			const cudaError_t el = ptp_launch(k_idct_r16_field<false>, dim3(rowsOut), dim3(256), sm16, t->stream, t->usePdl, spec, phi, t->fftTw, (const double*)nullptr, (double*)nullptr, nS, Nr, t->hz,
				rowsFormed ? nullptr : t->wideXb, t->wideJ, t->thP, TW_BLK);
			if (el != cudaSuccess) return ptp_cuda_fail(el, "k_idct_r16_field launch", __FILE__, __LINE__);
		}
		const cudaError_t e16 = cudaGetLastError();
		if (e16 != cudaSuccess) return ptp_cuda_fail(e16, "k_idct_r16_field launch", __FILE__, __LINE__);
		t->lastLaunches += 1;
		return PTP_OK;
	}
	const size_t sm = (size_t)N * sizeof(double2) + (size_t)(N + 1) * sizeof(double);
	const int threads = N >= 1024 ? 256 : 128;
	if (withField) {
		PTP_CUDA(cudaFuncSetAttribute(k_idct_fft_field<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
		const cudaError_t el = ptp_launch(k_idct_fft_field<true>, dim3(rowsOut), dim3(threads), sm, t->stream, t->usePdl, spec, phi, t->fftTw, t->phiTrap, t->eNodes, nS, Nr, N, bits, t->hz);
		if (el != cudaSuccess) return ptp_cuda_fail(el, "k_idct_fft_field launch", __FILE__, __LINE__);
	}
	else {
		PTP_CUDA(cudaFuncSetAttribute(k_idct_fft_field<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
		const cudaError_t el = ptp_launch(k_idct_fft_field<false>, dim3(rowsOut), dim3(threads), sm, t->stream, t->usePdl, spec, phi, t->fftTw, (const double*)nullptr, (double*)nullptr, nS, Nr, N, bits, t->hz);
		if (el != cudaSuccess) return ptp_cuda_fail(el, "k_idct_fft_field launch", __FILE__, __LINE__);
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_idct_fft_field launch", __FILE__, __LINE__);
	t->lastLaunches += 1;
	return PTP_OK;
}
