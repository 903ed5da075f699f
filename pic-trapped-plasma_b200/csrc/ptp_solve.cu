// K3 / K4: the electrostatic solve and the node field.
//
// The reference assembles the 5-point cylindrical finite-difference matrix in
// PenningTrap::generateSparse (Source/PenningTrap.cpp:94-162), LU-factorises it once with Eigen::SparseLU
// (:56-57) and calls solver.solve(b) for the trap potential (:202) and for every species' self potential
// (Source/Plasma.cpp:98). The operator is separable, A = T_r (x) I_z + I_r (x) T_z:
//   T_z: -2/hz^2 on the diagonal, 1/hz^2 off it, 2/hz^2 to the single neighbour at k = 0 and k = Nz
//        (Neumann mirror, :102,113,126,130,147,159)  ->  diagonalised exactly by the DCT-I of length Nz+1,
//        eigenvalues -2/hz^2 + 2/hz^2 cos(pi m / Nz);
//   T_r: tridiagonal in the radial index with the coefficients of :103,121-122,139,142-144.
// So A^-1 b = DCT-I^-1 . [Thomas solve in r per axial mode m] . DCT-I b, a direct solve of the same
// matrix (agreement with sparse LU ~1e-12 rel-L2, the rounding level of the LU itself).
//
// Kernels: k_row_bounds + k_fwd_thomas (forward DCT-I restricted to each row's non-zero deposit range, fused
// with the Thomas solve per axial mode, factors precomputed in extended precision on the host), k_inv_gemm
// (dense inverse DCT-I as an fp64 GEMM against a precomputed cosine matrix - works for any Nz; the grid
// sizes of this code are not powers of two), k_node_field = PenningTrap::getEField(int,int)
// (Source/PenningTrap.cpp:208-236) for all nodes, k_apply (A x), k_wall_rhs (Source/PenningTrap.cpp:163-198).
#include "ptp_internal.h"

#include <limits.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace {

// [solve-begin] (tests/emu/emu_solve.cpp runs the kernels from here to [solve-end] on host threads)
constexpr int INV_TM = 16;      // output rows (radial nodes) per CTA of the inverse kernel
constexpr int INV_TN = 40;      // output columns (axial nodes) per CTA
constexpr int INV_KC = 1024;    // modes staged in shared memory per chunk

// Per radial row: first / last axial node with a non-zero deposit (lo > hi: empty row). Deposits are exact
// zeros outside the plasma, so skipping them in the forward transform changes nothing bitwise.
__global__ void __launch_bounds__(256) k_row_bounds(const double* __restrict__ rho, int rows, int n1, int2* __restrict__ bounds, int Nr, int rowLimit)
{
	const int lane = threadIdx.x & 31;
	const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();
	if (row >= rows) return;
	if (row % Nr >= rowLimit) {                                 // no ring lives in this radial row: nothing to scan
		if (lane == 0) bounds[row] = make_int2(INT_MAX, INT_MIN);
		return;
	}
	const unsigned long long* w = reinterpret_cast<const unsigned long long*>(rho) + (size_t)row * n1;
	int lo = INT_MAX, hi = INT_MIN;
	constexpr int UB = 20;                                      // one wave of loads covers a row of up to 640 nodes
	for (int k0 = lane; k0 < n1; k0 += 32 * UB) {
		unsigned long long word[UB];
#pragma unroll
		for (int u = 0; u < UB; ++u) word[u] = w[min(k0 + 32 * u, n1 - 1)];
#pragma unroll
		for (int u = 0; u < UB; ++u) {
			const int k = k0 + 32 * u;
			const bool nz = (k < n1) && ((word[u] << 1) != 0ULL);   // any bit but the sign: non-zero as double and as int64
			lo = nz ? min(lo, k) : lo;
			hi = nz ? max(hi, k) : hi;
		}
	}
	for (int o = 16; o > 0; o >>= 1) {
		lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
		hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
	}
	if (lane == 0) bounds[row] = make_int2(lo, hi);
}

__device__ __forceinline__ void cp_async8(void* smemDst, const void* gmemSrc, bool valid)
{
	const unsigned int d = (unsigned int)__cvta_generic_to_shared(smemDst);
	const int bytes = valid ? 8 : 0;                            // src-size 0 -> the 8 bytes are zero-filled
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmemSrc), "r"(bytes));
}

// Forward half of the solve for FWD_MB axial modes per CTA and one species per blockIdx.y:
//   beta[j][m] = sum_k (scale * rho[j][k]) * FT[k][m]      (DCT-I with its weights, only over each row's non-zero range)
//   alpha[.][m] = (T_r + lambda_m I)^-1 beta[.][m]          (Thomas in r, factors precomputed)
// alpha is written to spec[s][j][m].
constexpr int FWD_KB = 64;      // axial nodes of the deposit staged per chunk
constexpr int FWD_RP = 128;     // deposit rows handled per pass
template <bool A_FIXED, int FWD_MB>
__global__ void __launch_bounds__(256) k_fwd_thomas(const double* __restrict__ rho, const int2* __restrict__ bounds, const uint2* __restrict__ encBounds,
	const double* __restrict__ FT, const double* __restrict__ rowScale, double fixedInv,
	const double* __restrict__ thInv, const double* __restrict__ thCp, const double* __restrict__ thR, const double* __restrict__ thQ,
	const double* __restrict__ thLower, double* __restrict__ spec, int Nr, int n1, int Jf, int rowsOut)
{
	// Jf: a row at or above the outermost deposit row (host-known: rings never change their row, so no deposit can land above
	// the outermost populated row; Nr - 1 when nothing is known). The rows above Jf carry no deposit and are folded into the
	// pivot of row Jf (thQ, "burn at both ends"): forward sweep over the touched rows below Jf, x_Jf from the folded pivot,
	// back-substitution below, and above Jf the homogeneous recurrence x_j = thR_j x_{j-1} - only as far as rowsOut, the
	// number of rows the caller wants (the push needs the potential in the populated rows only). With Jf = Nr - 1 this is the
	// plain Thomas solve.
	// Everything global is brought in with cp.async in a few large waves (the solve runs right after the push kernel
	// has streamed gigabytes through L2, so every serial round trip costs a DRAM latency).
	extern __shared__ double sB[];                              // [Nr][FWD_MB] beta -> alpha
	double* sInv = sB + (size_t)Nr * FWD_MB;                    // [Nr][FWD_MB] 1/pivot
	double* sCp = sInv + (size_t)Nr * FWD_MB;                   // [Nr][FWD_MB] upper/pivot
	double* sFT = sCp + (size_t)Nr * FWD_MB;                    // [FWD_KB][FWD_MB] chunk of the forward matrix
	double* sRho = sFT + FWD_KB * FWD_MB;                       // [FWD_RP][FWD_KB] chunk of the deposit rows of this pass
	int2* sBd = reinterpret_cast<int2*>(sRho + (size_t)FWD_RP * FWD_KB); // [Nr] non-zero range per row
	__shared__ int sLo, sHi, sJ0;
	const int tid = threadIdx.x, mi = tid % FWD_MB, slot = tid / FWD_MB;
	const int mBase = blockIdx.x * FWD_MB;
	const int m = mBase + mi;
	const int s = blockIdx.y;
	const bool mOk = m < n1;
	const double scale = (rowScale ? rowScale[s] : 1.0) * (A_FIXED ? fixedInv : 1.0);
	const double* b = rho + (size_t)s * Nr * n1;
	double* sLower = reinterpret_cast<double*>(sBd + Nr);       // [Nr] sub-diagonal of T_r
	const int lane = tid & 31, warp = tid >> 5;
	if (tid == 0) { sLo = INT_MAX; sHi = INT_MIN; sJ0 = Jf; }
	const double q = (tid < FWD_MB && mOk) ? thQ[(size_t)Jf * n1 + m] : 0.0;
	for (int j = slot; j < rowsOut || j <= Jf; j += 256 / FWD_MB) {
		const size_t off = (size_t)j * n1 + (mOk ? m : 0);
		if (j < Jf) {
			cp_async8(&sInv[j * FWD_MB + mi], thInv + off, mOk);
			cp_async8(&sCp[j * FWD_MB + mi], thCp + off, mOk);
		}
		else if (j > Jf) cp_async8(&sCp[j * FWD_MB + mi], thR + off, mOk);
		sB[j * FWD_MB + mi] = 0.0;
	}
	for (int j = tid; j < Nr; j += 256) cp_async8(&sLower[j], thLower + j, true);
	asm volatile("cp.async.commit_group;\n" ::);
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();                                             // the deposit (and its touched-node ranges) of this step; the tables above are constants
	__syncthreads();
	for (int j = tid; j < Nr; j += 256) {
		int2 bd;
		if (encBounds) {                                        // maxima written by the push kernel's flush: (Nz+2-kmin, kmax+1), 0 = untouched
			const uint2 e = encBounds[s * Nr + j];
			bd = e.y ? make_int2(n1 + 1 - (int)e.x, (int)e.y - 1) : make_int2(INT_MAX, INT_MIN);
		}
		else bd = bounds[s * Nr + j];
		if (j > Jf) bd = make_int2(INT_MAX, INT_MIN);           // (cannot hold a deposit)
		sBd[j] = bd;
		if (bd.x <= bd.y) { atomicMin(&sLo, bd.x); atomicMax(&sHi, bd.y); atomicMin(&sJ0, j); }
	}
	__syncthreads();
	const int kLo = sLo, kHi = sHi;
	constexpr int SLOTS = 256 / FWD_MB, U = FWD_RP / SLOTS;     // this thread's rows of a pass: jp + slot + u * SLOTS
	double acc[U];
	for (int jp = 0; jp <= Jf; jp += FWD_RP) {
#pragma unroll
		for (int u = 0; u < U; ++u) acc[u] = 0.0;
		for (int k0 = kLo; k0 <= kHi; k0 += FWD_KB) {
			const int kn = min(FWD_KB, kHi - k0 + 1);
			__syncthreads();
			for (int e = tid; e < kn * FWD_MB; e += 256) {
				const int kk = e / FWD_MB, mm = e % FWD_MB;
				const bool ok = mBase + mm < n1;
				cp_async8(&sFT[e], FT + (size_t)(k0 + kk) * n1 + (ok ? mBase + mm : 0), ok);
			}
			// deposit rows of this pass: one warp per row, only rows that have data in this chunk
			for (int jj = warp; jj < FWD_RP && jp + jj <= Jf; jj += 8) {
				const int j = jp + jj;
				const int2 bd = sBd[j];
				if (bd.x > bd.y || bd.y < k0 || bd.x >= k0 + kn) continue;      // uniform per warp
				for (int kk = lane; kk < kn; kk += 32) cp_async8(&sRho[(size_t)jj * FWD_KB + kk], b + (size_t)j * n1 + k0 + kk, true);
			}
			asm volatile("cp.async.commit_group;\n" ::);
			asm volatile("cp.async.wait_group 0;\n" ::);
			__syncthreads();
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int jj = slot + u * SLOTS, j = jp + jj;
				if (j > Jf) continue;
				const int2 bd = sBd[j];
				if (bd.x > bd.y) continue;                              // row without a deposit (the arithmetic below would overflow)
				const int a0 = max(bd.x, k0) - k0, a1 = min(bd.y, k0 + kn - 1) - k0;
				double t = acc[u];
				for (int kk = a0; kk <= a1; ++kk) {
					const double val = A_FIXED ? (double)reinterpret_cast<const long long*>(sRho)[(size_t)jj * FWD_KB + kk] : sRho[(size_t)jj * FWD_KB + kk];
					t = fma(val, sFT[kk * FWD_MB + mi], t);
				}
				acc[u] = t;
			}
		}
#pragma unroll
		for (int u = 0; u < U; ++u) {
			const int j = jp + slot + u * SLOTS;
			if (j <= Jf) sB[j * FWD_MB + mi] = acc[u] * scale;
		}
	}
	asm volatile("cp.async.wait_group 0;\n" ::);
	__syncthreads();
	if (tid < FWD_MB && mOk) {
		// forward sweep over rows J0 .. Jf-1 (J0 = first touched row; below it y = 0):  y_j = g_j - (l_j inv_j) y_{j-1}; the
		// coefficients of 8 rows are prepared off the chain, so the serial part is one dependent FMA per row
		const int J0 = sJ0;
		double y = 0.0;
		for (int j0 = J0; j0 < Jf; j0 += 8) {
			double c[8], g[8];
#pragma unroll
			for (int u = 0; u < 8; ++u) {
				const int j = min(j0 + u, Jf - 1);
				const double inv = sInv[j * FWD_MB + mi];
				g[u] = sB[j * FWD_MB + mi] * inv;
				c[u] = -(sLower[j] * inv);
			}
#pragma unroll
			for (int u = 0; u < 8; ++u)
				if (j0 + u < Jf) { y = fma(c[u], y, g[u]); g[u] = y; }
#pragma unroll
			for (int u = 0; u < 8; ++u)
				if (j0 + u < Jf) sB[(j0 + u) * FWD_MB + mi] = g[u];
		}
		// row Jf closes the system: the rows above it are folded into the pivot 1 / thQ
		const double xJ = (sB[Jf * FWD_MB + mi] - sLower[Jf] * y) * q;
		sB[Jf * FWD_MB + mi] = xJ;
		// back-substitution below Jf:  x_j = y_j - cp_j x_{j+1}
		y = xJ;
		for (int j0 = Jf - 1; j0 >= 0; j0 -= 8) {
			double c[8], g[8];
#pragma unroll
			for (int u = 0; u < 8; ++u) {
				const int j = max(j0 - u, 0);
				c[u] = -sCp[j * FWD_MB + mi];
				g[u] = sB[j * FWD_MB + mi];
			}
#pragma unroll
			for (int u = 0; u < 8; ++u)
				if (j0 - u >= 0) { y = fma(c[u], y, g[u]); g[u] = y; }
#pragma unroll
			for (int u = 0; u < 8; ++u)
				if (j0 - u >= 0) sB[(j0 - u) * FWD_MB + mi] = g[u];
		}
		// above Jf:  x_j = r_j x_{j-1}
		y = xJ;
		for (int j0 = Jf + 1; j0 < rowsOut; j0 += 8) {
			double c[8];
#pragma unroll
			for (int u = 0; u < 8; ++u) c[u] = sCp[min(j0 + u, rowsOut - 1) * FWD_MB + mi];
#pragma unroll
			for (int u = 0; u < 8; ++u)
				if (j0 + u < rowsOut) { y = c[u] * y; sB[(j0 + u) * FWD_MB + mi] = y; }
		}
	}
	__syncthreads();
	double* out = spec + (size_t)s * Nr * n1;
	for (int j = slot; j < rowsOut; j += 256 / FWD_MB)
		if (mOk) out[(size_t)j * n1 + m] = sB[j * FWD_MB + mi];
}

// Inverse DCT-I as a dense fp64 GEMM: C[M][N] = A[M][K] * B[K][N] (A = alpha, B = cosine matrix).
// The output is small (M*N ~ 75 k values) and K long, so a CTA owns a 16 x 40 output tile and its 8 warps
// split K between them (warp w takes k = w, w+8, ...): every lane keeps a 4 x 5 register tile (9 operand
// loads per 20 FMAs), the A rows are staged once in shared memory, and each warp streams its rows of B
// through a private 4-stage cp.async ring (8 rows = 2.5 KB per stage), so ~60 KB per SM are in flight and
// the kernel is bound by L2 bandwidth / fp64 issue rather than by load latency. The 8 partial tiles are
// summed through shared memory at the end.
constexpr int INV_KS = 8;       // B rows per pipeline stage and warp
constexpr int INV_ST = 4;       // stages

__device__ __forceinline__ void cp_async16(void* smemDst, const void* gmemSrc, bool valid)
{
	const unsigned int d = (unsigned int)__cvta_generic_to_shared(smemDst);
	const int bytes = valid ? 16 : 0;
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmemSrc), "r"(bytes));
}

// VEC: N is even, so every B row starts 16-byte aligned and the ring is filled with 16-byte copies.
template <bool VEC>
__global__ void __launch_bounds__(256) k_inv_gemm(const double* __restrict__ A, const double* __restrict__ B,
	double* __restrict__ C, int M, int N, int K)
{
	extern __shared__ double sm[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int la = lane >> 3, lb = lane & 7;                    // 4 row groups x 8 column groups
	const int m0 = blockIdx.y * INV_TM, n0 = blockIdx.x * INV_TN;
	const int kc = K < INV_KC ? K : INV_KC;
	const int lda = kc | 1;                                     // odd leading dimension: the 4 row groups hit distinct banks
	double* sA = sm;                                            // [16][lda]
	double* ring = sm + (size_t)INV_TM * lda + (size_t)warp * INV_ST * INV_KS * INV_TN; // [ST][KS][TN] of this warp
	double acc[4][5];
#pragma unroll
	for (int i = 0; i < 4; ++i)
#pragma unroll
		for (int j = 0; j < 5; ++j) acc[i][j] = 0.0;

	// per-lane copy slots of one stage (KS rows x TN columns), fixed for the whole kernel
	constexpr int EPL = VEC ? INV_KS * INV_TN / 2 / 32 : INV_KS * INV_TN / 32;   // copies per lane and stage
	int cpRow[EPL], cpCol[EPL];
	bool cpOk[EPL];
#pragma unroll
	for (int c = 0; c < EPL; ++c) {
		const int e = lane + 32 * c;
		cpRow[c] = VEC ? e / (INV_TN / 2) : e / INV_TN;
		cpCol[c] = VEC ? 2 * (e % (INV_TN / 2)) : e % INV_TN;
		cpOk[c] = n0 + cpCol[c] < N;
	}

	for (int k0 = 0; k0 < K; k0 += kc) {
		const int kn = min(kc, K - k0);
		const int rowsMine = kn > warp ? (kn - warp + 7) / 8 : 0;    // k = warp + 8 i, i < rowsMine
		const int nStages = (rowsMine + INV_KS - 1) / INV_KS;
		const double* bBase = B + (size_t)(k0 + warp) * N + n0;
		auto issue = [&](int st) {
			if (st < nStages) {
				double* dst = ring + (size_t)(st % INV_ST) * INV_KS * INV_TN;
#pragma unroll
				for (int c = 0; c < EPL; ++c) {
					const int i = st * INV_KS + cpRow[c];
					const bool valid = cpOk[c] && i < rowsMine;
					const double* src = bBase + (valid ? (size_t)8 * i * N + cpCol[c] : 0);
					if (VEC) cp_async16(dst + cpRow[c] * INV_TN + cpCol[c], src, valid);
					else cp_async8(dst + cpRow[c] * INV_TN + cpCol[c], src, valid);
				}
			}
			asm volatile("cp.async.commit_group;\n" ::);
		};
#pragma unroll
		for (int st = 0; st < INV_ST - 1; ++st) issue(st);          // B is in flight while A is staged
		__syncthreads();
		for (int r = 0; r < INV_TM; ++r) {                          // A tile: all copies in flight at once
			const bool ok = m0 + r < M;
			const double* src = A + (size_t)(ok ? m0 + r : 0) * K + k0;
			for (int c = tid; c < kn; c += 256) cp_async8(&sA[r * lda + c], src + c, ok);
		}
		asm volatile("cp.async.commit_group;\n" ::);
		asm volatile("cp.async.wait_group 0;\n" ::);
		__syncthreads();
		const double* a = sA + (4 * la) * lda;
		for (int st = 0; st < nStages; ++st) {
			issue(st + INV_ST - 1);
			asm volatile("cp.async.wait_group %0;\n" ::"n"(INV_ST - 1));
			__syncwarp();
			const double* bs = ring + (size_t)(st % INV_ST) * INV_KS * INV_TN + 5 * lb;
			const int kFirst = warp + 8 * st * INV_KS;
#pragma unroll
			for (int rr = 0; rr < INV_KS; ++rr) {
				const int k = min(kFirst + 8 * rr, kn - 1);             // rows past the end were zero-filled
				double bv[5], av[4];
#pragma unroll
				for (int j = 0; j < 5; ++j) bv[j] = bs[rr * INV_TN + j];
#pragma unroll
				for (int i = 0; i < 4; ++i) av[i] = a[i * lda + k];
#pragma unroll
				for (int i = 0; i < 4; ++i)
#pragma unroll
					for (int j = 0; j < 5; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
			}
			__syncwarp();
		}
		asm volatile("cp.async.wait_group 0;\n" ::);
	}
	__syncthreads();
	double* red = sm;                                           // [8][16][40]
#pragma unroll
	for (int i = 0; i < 4; ++i)
#pragma unroll
		for (int j = 0; j < 5; ++j) red[(warp * INV_TM + 4 * la + i) * INV_TN + 5 * lb + j] = acc[i][j];
	__syncthreads();
	for (int o = tid; o < INV_TM * INV_TN; o += 256) {
		double v = 0.0;
#pragma unroll
		for (int w = 0; w < 8; ++w) v += red[w * INV_TM * INV_TN + o];
		const int r = o / INV_TN, c = o - r * INV_TN;
		if (m0 + r < M && n0 + c < N) C[(size_t)(m0 + r) * N + n0 + c] = v;
	}
}

// Inverse DCT-I for all species of a 16-row strip + (optionally) the node field, in one kernel.
//   * cos(pi (Nz-m) k / Nz) = (-1)^k cos(pi m k / Nz): modes are paired, phi[k] = sum_{m < K2} (alpha_m +- alpha_{Nz-m}) C[m][k],
//     '+' for even k, '-' for odd k. The A tile is transformed in place (alpha_m + alpha_{Nz-m} at [m], the difference at
//     [Nz-m]) and every lane reads the variant that matches the parity of its columns: half the FMAs and half the B rows.
//   * FIELD: column tiles overlap by one node on each side (stride 38, width 40); the species' potentials of the tile are
//     summed onto phi_trap in registration order and E = (Phi[k-1] - Phi[k+1]) / (2 hz) is written for the 38 inner nodes -
//     the same arithmetic as k_node_field (Source/PenningTrap.cpp:226-233), so the result is bit-identical.
template <bool VEC, bool FIELD>
__global__ void __launch_bounds__(256) k_inv_field(const double* __restrict__ alpha, const double* __restrict__ B,
	double* __restrict__ phiSelf, const double* __restrict__ phiTrap, double* __restrict__ eNodes, int nS, int Nr, int n1, double hz, int ST)
{
	// ST = stages of the per-warp B ring. When the whole K-slice of a warp fits (ST >= its stage count; the default grid
	// needs 5 x 2.5 KB) everything is requested up front together with the A tile and the main loop never waits;
	// otherwise the ring is refilled one stage per iteration with INV_ST - 1 stages in flight.
	extern __shared__ double sm[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int la = lane >> 3, lb = lane & 7;                    // 4 row groups x 8 column groups
	const int Nz = n1 - 1, K2 = (n1 + 1) / 2;
	const int j0 = blockIdx.y * INV_TM;
	const int c0 = FIELD ? (int)blockIdx.x * (INV_TN - 2) - 2 : (int)blockIdx.x * INV_TN;   // first column of the tile (even; may be -2)
	const int lda = n1 | 1;
	double* sA = sm;                                            // [16][lda]
	double* ringAll = sm + (size_t)INV_TM * lda;               // 8 warps x [ST][KS][TN]; reused for the cross-warp reduction
	double* ring = ringAll + (size_t)warp * ST * INV_KS * INV_TN;
	double* sTot = ringAll + (size_t)8 * ST * INV_KS * INV_TN; // [16][40] total potential of the tile (FIELD)

	// per-lane copy slots of one stage (KS rows x TN columns). VEC needs 16-byte aligned sources: even n1 and even c0.
	constexpr int EPL = VEC ? INV_KS * INV_TN / 2 / 32 : INV_KS * INV_TN / 32;
	int cpRow[EPL], cpCol[EPL];
	bool cpOk[EPL];
#pragma unroll
	for (int c = 0; c < EPL; ++c) {
		const int e = lane + 32 * c;
		cpRow[c] = VEC ? e / (INV_TN / 2) : e / INV_TN;
		cpCol[c] = VEC ? 2 * (e % (INV_TN / 2)) : e % INV_TN;
		cpOk[c] = c0 + cpCol[c] >= 0 && c0 + cpCol[c] + (VEC ? 1 : 0) < n1;
	}
	// parity of this lane's first column decides which variant (sum / difference) its even-j columns use
	const bool firstEven = ((c0 + 5 * lb) & 1) == 0;
	const int offE = firstEven ? 0 : Nz, sgnE = firstEven ? 1 : -1;  // index of the value for columns j = 0, 2, 4: offE + sgnE * m
	const int offO = firstEven ? Nz : 0, sgnO = -sgnE;               // ... and for j = 1, 3
	const int rowsMine = K2 > warp ? (K2 - warp + 7) / 8 : 0;        // this warp's modes: m = warp + 8 i
	const int nStages = (rowsMine + INV_KS - 1) / INV_KS;
	const bool upFront = ST >= (K2 + 8 * INV_KS - 1) / (8 * INV_KS);   // uniform over the CTA (warp 0 has the most rows)
	const double* bBase = B + (size_t)warp * n1 + c0;

	ptp_pdl_launch_dependents();
	// The first species' slice of the cosine matrix (a constant) is requested before the wait, so that it streams in while
	// the forward kernel is still finishing; alpha and the trap potential are read after it.
	bool early = false;
	if (upFront) {
		for (int st = 0; st < nStages; ++st) {
			double* dst = ring + (size_t)(st % ST) * INV_KS * INV_TN;
#pragma unroll
			for (int c = 0; c < EPL; ++c) {
				const int i = st * INV_KS + cpRow[c];
				const bool valid = cpOk[c] && i < rowsMine;
				const double* src = valid ? bBase + (size_t)8 * i * n1 + cpCol[c] : B;
				if (VEC) cp_async16(dst + cpRow[c] * INV_TN + cpCol[c], src, valid);
				else cp_async8(dst + cpRow[c] * INV_TN + cpCol[c], src, valid);
			}
			asm volatile("cp.async.commit_group;\n" ::);
		}
		early = true;
	}
	ptp_pdl_wait();
	if (FIELD) {
		for (int o = tid; o < INV_TM * INV_TN; o += 256) {
			const int r = o / INV_TN, c = o - r * INV_TN, col = c0 + c;
			sTot[o] = (j0 + r < Nr && col >= 0 && col < n1) ? phiTrap[(size_t)(j0 + r) * n1 + col] : 0.0;
		}
	}
	for (int sp = 0; sp < nS; ++sp) {
		double acc[4][5];
#pragma unroll
		for (int i = 0; i < 4; ++i)
#pragma unroll
			for (int j = 0; j < 5; ++j) acc[i][j] = 0.0;
		auto issue = [&](int st) {
			if (st < nStages) {
				double* dst = ring + (size_t)(st % ST) * INV_KS * INV_TN;
#pragma unroll
				for (int c = 0; c < EPL; ++c) {
					const int i = st * INV_KS + cpRow[c];
					const bool valid = cpOk[c] && i < rowsMine;
					const double* src = valid ? bBase + (size_t)8 * i * n1 + cpCol[c] : B;
					if (VEC) cp_async16(dst + cpRow[c] * INV_TN + cpCol[c], src, valid);
					else cp_async8(dst + cpRow[c] * INV_TN + cpCol[c], src, valid);
				}
			}
			asm volatile("cp.async.commit_group;\n" ::);
		};
		__syncthreads();                                            // previous species' reduction buffers are free again
		if (upFront) {
			if (!(early && sp == 0)) for (int st = 0; st < nStages; ++st) issue(st);
		}
		else {
#pragma unroll
			for (int st = 0; st < INV_ST - 1; ++st) issue(st);      // B is in flight while A is staged
		}
		const double* aSrc = alpha + ((size_t)sp * Nr + j0) * n1;
		for (int r = 0; r < INV_TM; ++r) {
			const bool ok = j0 + r < Nr;
			for (int c = tid; c < n1; c += 256) cp_async8(&sA[r * lda + c], aSrc + (ok ? (size_t)r * n1 + c : 0), ok);
		}
		asm volatile("cp.async.commit_group;\n" ::);
		asm volatile("cp.async.wait_group 0;\n" ::);
		__syncthreads();
		for (int e = tid; e < INV_TM * K2; e += 256) {              // pair the modes: [m] <- a_m + a_{Nz-m}, [Nz-m] <- a_m - a_{Nz-m}
			const int r = e / K2, m = e - r * K2;
			if (m != Nz - m) {
				const double p = sA[r * lda + m], q = sA[r * lda + Nz - m];
				sA[r * lda + m] = p + q;
				sA[r * lda + Nz - m] = p - q;
			}
		}
		__syncthreads();
		const double* a = sA + (4 * la) * lda;
		for (int st = 0; st < nStages; ++st) {
			if (!upFront) {
				issue(st + INV_ST - 1);
				asm volatile("cp.async.wait_group %0;\n" ::"n"(INV_ST - 1));
				__syncwarp();
			}
			const double* bs = ring + (size_t)(st % ST) * INV_KS * INV_TN + 5 * lb;
			const int mFirst = warp + 8 * st * INV_KS;
#pragma unroll
			for (int rr = 0; rr < INV_KS; ++rr) {
				const int m = min(mFirst + 8 * rr, K2 - 1);             // rows past the end were zero-filled
				const int iE = offE + sgnE * m, iO = offO + sgnO * m;
				double bv[5], ae[4], ao[4];
#pragma unroll
				for (int j = 0; j < 5; ++j) bv[j] = bs[rr * INV_TN + j];
#pragma unroll
				for (int i = 0; i < 4; ++i) { ae[i] = a[i * lda + iE]; ao[i] = a[i * lda + iO]; }
#pragma unroll
				for (int i = 0; i < 4; ++i)
#pragma unroll
					for (int j = 0; j < 5; ++j) acc[i][j] = fma((j & 1) ? ao[i] : ae[i], bv[j], acc[i][j]);
			}
			if (!upFront) __syncwarp();
		}
		asm volatile("cp.async.wait_group 0;\n" ::);
		__syncthreads();                                            // every warp is done with its ring: reuse as [8][16][40]
#pragma unroll
		for (int i = 0; i < 4; ++i)
#pragma unroll
			for (int j = 0; j < 5; ++j) ringAll[(warp * INV_TM + 4 * la + i) * INV_TN + 5 * lb + j] = acc[i][j];
		__syncthreads();
		double* out = phiSelf + (size_t)sp * Nr * n1;
		for (int o = tid; o < INV_TM * INV_TN; o += 256) {
			double v = 0.0;
#pragma unroll
			for (int w = 0; w < 8; ++w) v += ringAll[w * INV_TM * INV_TN + o];
			const int r = o / INV_TN, c = o - r * INV_TN, col = c0 + c;
			if (j0 + r < Nr && col >= 0 && col < n1) {
				out[(size_t)(j0 + r) * n1 + col] = v;             // overlapping columns get the identical value from both tiles
				if (FIELD) sTot[o] = __dadd_rn(sTot[o], v);
			}
		}
	}
	if (FIELD) {
		__syncthreads();
		for (int o = tid; o < INV_TM * (INV_TN - 2); o += 256) {
			const int r = o / (INV_TN - 2), c = 1 + o % (INV_TN - 2), col = c0 + c;
			if (j0 + r >= Nr || col < 0 || col >= n1) continue;     // the first tile starts at column -2
			double e = 0.0;
			if (col > 0 && col < n1 - 1)
				e = __ddiv_rn(__dsub_rn(sTot[r * INV_TN + c - 1], sTot[r * INV_TN + c + 1]), __dmul_rn(2.0, hz));
			eNodes[(size_t)(j0 + r) * n1 + col] = e;
		}
	}
}

// ---- bulk-async (TMA) staging -----------------------------------------------------------------------------------
// cp.async.bulk moves a whole contiguous run with ONE instruction issued by one thread; completion is signalled on an
// mbarrier through its transaction count. Source, destination and size must be multiples of 16 bytes.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned int)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned int)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@!p bra WAIT_%=;\n"
		"}\n" ::"r"((unsigned int)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smemDst, const void* gmemSrc, unsigned int bytes, unsigned long long* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"((unsigned int)__cvta_generic_to_shared(smemDst)),
		"l"(gmemSrc), "r"(bytes), "r"((unsigned int)__cvta_generic_to_shared(bar)) : "memory");
}
// generic-proxy accesses to shared memory (ordered by the preceding barrier) before the async proxy writes there again
__device__ __forceinline__ void fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// k_inv_field for even row lengths (every grid row and every row of the cosine matrix then starts 16-byte aligned), staged
// with bulk-async copies: the CTA's whole slice of the cosine matrix - K2 rows of 40 columns, 320 B each - and the 16 alpha
// rows of the strip arrive through ~300 copy instructions instead of ~15 000 eight- and sixteen-byte cp.async, the cosine
// slice once for all species and, being a constant, before the programmatic-launch wait (it streams in while the forward
// kernel is still finishing). Same tiling, same paired modes, same arithmetic order per output value as k_inv_field, so
// potentials and node field are bit-identical to it.
// Shared memory: sA [16 rows] (row r at r * LDA + 2 (r >> 2): 16-byte aligned and the four rows a quarter-warp group reads
// together fall into distinct banks) | sB [K2][40] | sTot [16][40]; the cross-warp reduction reuses sB.
template <bool FIELD>
__global__ void __launch_bounds__(256) k_inv_field_bulk(const double* __restrict__ alpha, const double* __restrict__ B,
	double* __restrict__ phiSelf, const double* __restrict__ phiTrap, double* __restrict__ eNodes, int nS, int Nr, int n1, double hz, int LDA)
{
	extern __shared__ __align__(16) double sm[];
	__shared__ unsigned long long bar[2];                           // [0] alpha tile (one phase per species), [1] cosine slice
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int la = lane >> 3, lb = lane & 7;                    // 4 row groups x 8 column groups
	const int Nz = n1 - 1, K2 = (n1 + 1) / 2;
	const int j0 = blockIdx.y * INV_TM;
	const int c0 = FIELD ? (int)blockIdx.x * (INV_TN - 2) - 2 : (int)blockIdx.x * INV_TN;   // first column of the tile (even; may be -2)
	double* sA = sm;
	double* sB = sm + (size_t)INV_TM * LDA + 8;
	double* sTot = sB + (size_t)max(K2, 8 * INV_TM) * INV_TN;   // (the reduction needs [8][16][40] there, also on very short rows)
	auto rowOff = [LDA](int r) { return r * LDA + 2 * (r >> 2); };
	const int cFirst = max(c0, 0), cEnd = min(c0 + INV_TN, n1);  // valid columns of the tile: [cFirst, cEnd), both even
	const unsigned int rowBytes = (unsigned int)(cEnd - cFirst) * 8u;
	const int rowsValid = min(INV_TM, Nr - j0);
	if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
	// columns outside the grid and rows below the strip's end are never stored; they get zeros once so that nothing undefined is computed
	if (cEnd - cFirst < INV_TN)
		for (int e = tid; e < K2 * INV_TN; e += 256) sB[e] = 0.0;
	for (int e = tid; e < INV_TM * LDA + 8; e += 256) sA[e] = 0.0;
	__syncthreads();
	fence_proxy_async();
	ptp_pdl_launch_dependents();
	if (warp == 0) {
		if (lane == 0) mbar_expect_tx(&bar[1], rowBytes * (unsigned int)K2);
		__syncwarp();
		for (int m = lane; m < K2; m += 32) bulk_g2s(sB + (size_t)m * INV_TN + (cFirst - c0), B + (size_t)m * n1 + cFirst, rowBytes, &bar[1]);
	}
	ptp_pdl_wait();                                             // alpha, the trap potential (constant in a step, but not a table)
	if (FIELD) {
		for (int o = tid; o < INV_TM * INV_TN; o += 256) {
			const int r = o / INV_TN, c = o - r * INV_TN, col = c0 + c;
			sTot[o] = (j0 + r < Nr && col >= 0 && col < n1) ? phiTrap[(size_t)(j0 + r) * n1 + col] : 0.0;
		}
	}
	// parity of this lane's first column decides which variant (sum / difference) its even-j columns use
	const bool firstEven = ((c0 + 5 * lb) & 1) == 0;
	const int offE = firstEven ? 0 : Nz, sgnE = firstEven ? 1 : -1;  // index of the value for columns j = 0, 2, 4: offE + sgnE * m
	const int offO = firstEven ? Nz : 0, sgnO = -sgnE;               // ... and for j = 1, 3
	const int rowsMine = K2 > warp ? (K2 - warp + 7) / 8 : 0;        // this warp's modes: m = warp + 8 i
	const double* a0 = sA + rowOff(4 * la), *a1 = sA + rowOff(4 * la + 1), *a2 = sA + rowOff(4 * la + 2), *a3 = sA + rowOff(4 * la + 3);
	for (int sp = 0; sp < nS; ++sp) {
		if (sp > 0) { __syncthreads(); fence_proxy_async(); }       // the last species' reads of sA and of the reduction buffer are done
		if (warp == 1) {
			const double* aSrc = alpha + ((size_t)sp * Nr + j0) * n1;
			if (lane == 0) mbar_expect_tx(&bar[0], (unsigned int)rowsValid * (unsigned int)n1 * 8u);
			__syncwarp();
			if (lane < rowsValid) bulk_g2s(sA + rowOff(lane), aSrc + (size_t)lane * n1, (unsigned int)n1 * 8u, &bar[0]);
		}
		if (sp > 0 && warp == 0) {                                  // the reduction overwrote the head of the cosine slice: fetch those rows again
			const int rowsDirty = min(K2, (8 * INV_TM * INV_TN + INV_TN - 1) / INV_TN);
			if (lane == 0) mbar_expect_tx(&bar[1], rowBytes * (unsigned int)rowsDirty);
			__syncwarp();
			for (int m = lane; m < rowsDirty; m += 32) bulk_g2s(sB + (size_t)m * INV_TN + (cFirst - c0), B + (size_t)m * n1 + cFirst, rowBytes, &bar[1]);
		}
		mbar_wait(&bar[0], (unsigned int)(sp & 1));
		for (int e = tid; e < INV_TM * K2; e += 256) {              // pair the modes: [m] <- a_m + a_{Nz-m}, [Nz-m] <- a_m - a_{Nz-m}
			const int r = e / K2, m = e - r * K2;
			if (m != Nz - m) {
				double* row = sA + rowOff(r);
				const double p = row[m], q = row[Nz - m];
				row[m] = p + q;
				row[Nz - m] = p - q;
			}
		}
		mbar_wait(&bar[1], (unsigned int)(sp & 1));
		__syncthreads();
		double acc[4][5];
#pragma unroll
		for (int i = 0; i < 4; ++i)
#pragma unroll
			for (int j = 0; j < 5; ++j) acc[i][j] = 0.0;
		// the same order of accumulation per output value as k_inv_field: modes warp, warp + 8, ... in stages of INV_KS
		for (int i0 = 0; i0 < rowsMine; i0 += INV_KS) {
#pragma unroll
			for (int rr = 0; rr < INV_KS; ++rr) {
				const int i = i0 + rr;
				const int m = min(warp + 8 * i, K2 - 1);
				const double* bs = sB + (size_t)m * INV_TN + 5 * lb;
				const int iE = offE + sgnE * m, iO = offO + sgnO * m;
				double bv[5], ae[4], ao[4];
#pragma unroll
				for (int j = 0; j < 5; ++j) bv[j] = i < rowsMine ? bs[j] : 0.0;
				ae[0] = a0[iE]; ae[1] = a1[iE]; ae[2] = a2[iE]; ae[3] = a3[iE];
				ao[0] = a0[iO]; ao[1] = a1[iO]; ao[2] = a2[iO]; ao[3] = a3[iO];
#pragma unroll
				for (int ii = 0; ii < 4; ++ii)
#pragma unroll
					for (int j = 0; j < 5; ++j) acc[ii][j] = fma((j & 1) ? ao[ii] : ae[ii], bv[j], acc[ii][j]);
			}
		}
		__syncthreads();                                            // every warp is done with the cosine slice: its head becomes [8][16][40]
#pragma unroll
		for (int i = 0; i < 4; ++i)
#pragma unroll
			for (int j = 0; j < 5; ++j) sB[(warp * INV_TM + 4 * la + i) * INV_TN + 5 * lb + j] = acc[i][j];
		__syncthreads();
		double* out = phiSelf + (size_t)sp * Nr * n1;
		for (int o = tid; o < INV_TM * INV_TN; o += 256) {
			double v = 0.0;
#pragma unroll
			for (int w = 0; w < 8; ++w) v += sB[w * INV_TM * INV_TN + o];
			const int r = o / INV_TN, c = o - r * INV_TN, col = c0 + c;
			if (j0 + r < Nr && col >= 0 && col < n1) {
				out[(size_t)(j0 + r) * n1 + col] = v;             // overlapping columns get the identical value from both tiles
				if (FIELD) sTot[o] = __dadd_rn(sTot[o], v);
			}
		}
	}
	if (FIELD) {
		__syncthreads();
		for (int o = tid; o < INV_TM * (INV_TN - 2); o += 256) {
			const int r = o / (INV_TN - 2), c = 1 + o % (INV_TN - 2), col = c0 + c;
			if (j0 + r >= Nr || col < 0 || col >= n1) continue;     // the first tile starts at column -2
			double e = 0.0;
			if (col > 0 && col < n1 - 1)
				e = __ddiv_rn(__dsub_rn(sTot[r * INV_TN + c - 1], sTot[r * INV_TN + c + 1]), __dmul_rn(2.0, hz));
			eNodes[(size_t)(j0 + r) * n1 + col] = e;
		}
	}
}

// PenningTrap::getEField(int,int), Source/PenningTrap.cpp:208-236: E = (sumPhi[idx-1] - sumPhi[idx+1]) / (2 hz),
// zero at both axial ends, species added in registration order.
__global__ void k_node_field(const double* __restrict__ phiTrap, const double* __restrict__ phiSelf, int nS,
	long long G, int n1, double hz, double* __restrict__ eNodes)
{
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= G) return;
	const int k = (int)(idx % n1);
	if (k == 0 || k == n1 - 1) { eNodes[idx] = 0.0; return; }
	double left = phiTrap[idx - 1], right = phiTrap[idx + 1];
	for (int s = 0; s < nS; ++s) {
		left = __dadd_rn(left, phiSelf[(size_t)s * G + idx - 1]);
		right = __dadd_rn(right, phiSelf[(size_t)s * G + idx + 1]);
	}
	eNodes[idx] = __ddiv_rn(__dsub_rn(left, right), __dmul_rn(2.0, hz));
}

// y = A x with the stencil of generateSparse (Source/PenningTrap.cpp:94-162).
__global__ void k_apply(const double* __restrict__ x, double* __restrict__ y, int Nr, int n1, double diag, double hz2,
	const double* __restrict__ lower, const double* __restrict__ upper)
{
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (long long)Nr * n1) return;
	const int j = (int)(idx / n1), k = (int)(idx % n1);
	double s = diag * x[idx];
	if (k > 0) s += (k == n1 - 1 ? 2.0 * hz2 : hz2) * x[idx - 1];
	if (k < n1 - 1) s += (k == 0 ? 2.0 * hz2 : hz2) * x[idx + 1];
	if (j > 0) s += lower[j] * x[idx - n1];
	if (j < Nr - 1) s += upper[j] * x[idx + n1];
	y[idx] = s;
}

// PenningTrap::updateRHS, Source/PenningTrap.cpp:177,185,195: RHS(last row) = -1 * matrixFactor * boundary.
__global__ void k_wall_rhs(const double* __restrict__ wall, double* __restrict__ rhs, long long G, int n1, double factor)
{
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= G) return;
	const long long first = G - n1;
	rhs[idx] = idx >= first ? __dmul_rn(__dmul_rn(-1.0, factor), wall[idx - first]) : 0.0;
}

// [solve-end]

// Red-black SOR sweep (cross-check solver): one colour of  A phi = b.
template <bool A_FIXED>
__global__ void k_sor_sweep(double* __restrict__ phi, const double* __restrict__ rho, const double* __restrict__ scale,
	int species, double fixedInv, int Nr, int n1, double diag, double hz2, const double* __restrict__ lower,
	const double* __restrict__ upper, double omega, int colour, double* __restrict__ resid2)
{
	const long long half = ((long long)Nr * n1 + 1) / 2;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	double r2 = 0.0;
	if (t < half) {
		// enumerate nodes of this colour: (j + k) % 2 == colour
		const int perRow = n1;
		long long idx = 2 * t;
		int j = (int)(idx / perRow), k = (int)(idx % perRow);
		if (((j + k) & 1) != colour) { ++idx; j = (int)(idx / perRow); k = (int)(idx % perRow); }
		if (idx < (long long)Nr * n1 && ((j + k) & 1) == colour) {
			double b;
			if (A_FIXED) b = (double)reinterpret_cast<const long long*>(rho)[idx] * fixedInv;
			else b = rho[idx];
			if (scale) b *= scale[species];
			double s = 0.0;
			if (k > 0) s += (k == n1 - 1 ? 2.0 * hz2 : hz2) * phi[idx - 1];
			if (k < n1 - 1) s += (k == 0 ? 2.0 * hz2 : hz2) * phi[idx + 1];
			if (j > 0) s += lower[j] * phi[idx - n1];
			if (j < Nr - 1) s += upper[j] * phi[idx + n1];
			const double res = b - s - diag * phi[idx];
			phi[idx] += omega * res / diag;
			r2 = res * res;
		}
	}
	if (resid2) {
		for (int o = 16; o > 0; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
		if ((threadIdx.x & 31) == 0 && r2 != 0.0) atomicAdd(resid2, r2);
	}
}

__global__ void k_norm2(const double* __restrict__ rho, const double* __restrict__ scale, int species, bool fixed,
	double fixedInv, long long G, double* __restrict__ out)
{
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	double v = 0.0;
	if (idx < G) {
		v = fixed ? (double)reinterpret_cast<const long long*>(rho)[idx] * fixedInv : rho[idx];
		if (scale) v *= scale[species];
		v *= v;
	}
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(out, v);
}

} // namespace

// Host precompute of the operator: stencil coefficients exactly as the reference evaluates them, DCT matrices and
// Thomas factors in extended precision.
int ptp_solver_build(ptp_trap* t)
{
	// [tables-begin] (pure host arithmetic up to [tables-end]: tests/emu/emu_wide.cpp includes this text to feed the kernels it emulates)
	const int Nz = t->Nz, Nr = t->Nr, n1 = Nz + 1;
	const double hz = t->hz, hr = t->hr;
	const double hz2 = std::pow(hz, -2);                                   // Source/PenningTrap.cpp:97
	const double hr2 = std::pow(hr, -2);                                   // :98
	const double diag = -2 * (hr2 + hz2);                                  // :101
	std::vector<double> lower(Nr, 0.0), upper(Nr, 0.0);
	upper[0] = 2 * hr2;                                                    // :103
	for (int j = 1; j < Nr - 1; ++j) {
		const double radius = std::floor((double)j) * hr;                  // :121
		lower[j] = hr2 - std::pow(2 * radius * hr, -1);                    // :122
		upper[j] = hr2 + std::pow(2 * radius * hr, -1);                    // :139
	}
	if (Nr > 1) {
		const double radius = t->radius - hr;                              // :142
		lower[Nr - 1] = hr2 - std::pow(2 * radius * hr, -1);               // :144
		upper[Nr - 1] = 0.0;                                               // Dirichlet: wall term lives in the RHS
	}
	t->stDiag = diag;
	t->stHz2 = hz2;
	t->wallFactor = hr2 + std::pow(2 * (t->radius - hr) * hr, -1);         // :165-167

	const long double pi = 3.141592653589793238462643383279502884L;
	std::vector<double> fwd((size_t)n1 * n1), inv((size_t)n1 * n1);
	for (int k = 0; k < n1; ++k) {
		const long double wk = (k == 0 || k == Nz) ? 0.5L : 1.0L;
		for (int m = 0; m < n1; ++m) {
			const long double wm = (m == 0 || m == Nz) ? 0.5L : 1.0L;
			const long long red = ((long long)k * m) % (2LL * Nz);         // exact argument reduction
			const long double c = cosl(pi * (long double)red / (long double)Nz);
			fwd[(size_t)k * n1 + m] = (double)((2.0L / Nz) * wk * wm * c);
			inv[(size_t)m * n1 + k] = (double)c;
		}
	}
	// thR / thQ: the same factorisation started from the wall (rows above the plasma carry no deposit):
	//   r_j = -l_j / (d_m + u_j r_{j+1})  propagates x_j = r_j x_{j-1} above the outermost deposit row J,
	//   q_J = 1 / (pivot_J + u_J r_{J+1}) closes the downward sweep at row J (k_thomas_wide, ptp_solve_wide.cu).
	std::vector<double> thInv((size_t)Nr * n1), thCp((size_t)Nr * n1), thR((size_t)Nr * n1), thQ((size_t)Nr * n1);
	std::vector<double> thP((size_t)Nr * n1);
	std::vector<long double> piv((size_t)Nr), rr((size_t)Nr);
	for (int m = 0; m < n1; ++m) {
		const long double dm = (long double)diag + 2.0L * (long double)hz2 * cosl(pi * (long double)m / (long double)Nz);
		long double cpPrev = 0.0L;
		for (int j = 0; j < Nr; ++j) {
			const long double pivot = dm - (long double)lower[j] * cpPrev;
			const long double pinv = 1.0L / pivot;
			cpPrev = (long double)upper[j] * pinv;
			piv[j] = pivot;
			thInv[(size_t)j * n1 + m] = (double)pinv;
			thCp[(size_t)j * n1 + m] = (double)cpPrev;
		}
		long double rNext = 0.0L;
		for (int j = Nr - 1; j >= 0; --j) {
			thQ[(size_t)j * n1 + m] = (double)(1.0L / (piv[j] + (long double)upper[j] * rNext));
			rNext = -(long double)lower[j] / (dm + (long double)upper[j] * rNext);
			thR[(size_t)j * n1 + m] = (double)rNext;
			rr[j] = rNext;
		}
		long double prod = 1.0L;
		for (int j = 0; j < Nr; ++j) {                          // in-block prefix products of r (k_thomas_expand)
			if (j % PTP_THOMAS_BLOCK == 0) prod = 1.0L;
			prod *= rr[j];
			thP[(size_t)j * n1 + m] = (double)prod;
		}
	}
	// [tables-end]
	if (Nz >= 8 && (Nz & (Nz - 1)) == 0) {                      // power-of-two cell count: FFT path for the inverse transform
		std::vector<double2> tw((size_t)Nz);
		for (int j = 0; j < Nz; ++j) {
			const long double ang = -pi * (long double)j / (long double)Nz;     // exp(-2 pi i j / (2 Nz))
			tw[j] = make_double2((double)cosl(ang), (double)sinl(ang));
		}
		PTP_CUDA(cudaMalloc(&t->fftTw, (size_t)Nz * sizeof(double2)));
		PTP_CUDA(cudaMemcpy(t->fftTw, tw.data(), (size_t)Nz * sizeof(double2), cudaMemcpyHostToDevice));
	}
	const size_t nn = (size_t)n1 * n1 * sizeof(double), gg = (size_t)Nr * n1 * sizeof(double);
	// one allocation for all solver constants, so that a single L2 access-policy window can keep them resident
	// while the push kernel streams the ring arrays through L2 between two solves
	const size_t constBytes = 2 * nn + 2 * gg;
	PTP_CUDA(cudaMalloc(&t->solverConst, constBytes));
	t->dctInv = t->solverConst;
	t->dctFwd = t->dctInv + (size_t)n1 * n1;
	t->thInv = t->dctFwd + (size_t)n1 * n1;
	t->thCp = t->thInv + (size_t)Nr * n1;
	{
		cudaDeviceProp prop;
		PTP_CUDA(cudaGetDeviceProperties(&prop, t->device));
		// only when the whole set fits: on large grids (config 5: 335 MB of tables) a partial carve-out would just take most
		// of L2 away from the push kernel's prefetch stream without making the solve any faster
		if (prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0 && constBytes <= (size_t)prop.persistingL2CacheMaxSize / 2 &&
			!getenv("PTP_NO_L2_PERSIST")) {
			const size_t carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, constBytes);
			size_t current = 0;
			cudaDeviceGetLimit(&current, cudaLimitPersistingL2CacheSize);
			if (current < carve) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
			cudaStreamAttrValue attr{};
			attr.accessPolicyWindow.base_ptr = t->solverConst;
			attr.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)prop.accessPolicyMaxWindowSize, constBytes);
			attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)attr.accessPolicyWindow.num_bytes);
			attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
			attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
			if (cudaStreamSetAttribute(t->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
		}
	}
	PTP_CUDA(cudaMalloc(&t->thR, gg));
	PTP_CUDA(cudaMalloc(&t->thQ, gg));
	PTP_CUDA(cudaMemcpy(t->thR, thR.data(), gg, cudaMemcpyHostToDevice));
	PTP_CUDA(cudaMemcpy(t->thQ, thQ.data(), gg, cudaMemcpyHostToDevice));
	PTP_CUDA(cudaMalloc(&t->thP, gg));
	PTP_CUDA(cudaMemcpy(t->thP, thP.data(), gg, cudaMemcpyHostToDevice));
	PTP_CUDA(cudaMalloc(&t->thLower, Nr * sizeof(double)));
	PTP_CUDA(cudaMalloc(&t->stLower, Nr * sizeof(double)));
	PTP_CUDA(cudaMalloc(&t->stUpper, Nr * sizeof(double)));
	PTP_CUDA(cudaMemcpy(t->dctFwd, fwd.data(), nn, cudaMemcpyHostToDevice));
	PTP_CUDA(cudaMemcpy(t->dctInv, inv.data(), nn, cudaMemcpyHostToDevice));
	PTP_CUDA(cudaMemcpy(t->thInv, thInv.data(), gg, cudaMemcpyHostToDevice));
	PTP_CUDA(cudaMemcpy(t->thCp, thCp.data(), gg, cudaMemcpyHostToDevice));
	PTP_CUDA(cudaMemcpy(t->thLower, lower.data(), Nr * sizeof(double), cudaMemcpyHostToDevice));
	PTP_CUDA(cudaMemcpy(t->stLower, lower.data(), Nr * sizeof(double), cudaMemcpyHostToDevice));
	PTP_CUDA(cudaMemcpy(t->stUpper, upper.data(), Nr * sizeof(double), cudaMemcpyHostToDevice));
	return PTP_OK;
}

void ptp_solver_free(ptp_trap* t)
{
	cudaFree(t->solverConst); cudaFree(t->rowBounds); cudaFree(t->fftTw);
	cudaFree(t->thLower); cudaFree(t->stLower); cudaFree(t->stUpper); cudaFree(t->thR); cudaFree(t->thQ); cudaFree(t->thP); cudaFree(t->wideXb); cudaFree(t->wideJ);
}

int ptp_solver_reserve(ptp_trap* t, int nS)
{
	const int M = nS * t->Nr;
	if ((int)t->rowBoundsCap < M) {
		cudaFree(t->rowBounds);
		t->rowBounds = nullptr;
		PTP_CUDA(cudaMalloc(&t->rowBounds, (size_t)M * sizeof(int2)));
		t->rowBoundsCap = M;
	}
	if (t->wideCap < nS) {                                      // scratch of the large-grid radial solves
		cudaFree(t->wideXb); cudaFree(t->wideJ);
		t->wideXb = nullptr; t->wideJ = nullptr;
		const size_t nB = (size_t)(t->Nr + PTP_THOMAS_BLOCK - 1) / PTP_THOMAS_BLOCK;
		PTP_CUDA(cudaMalloc(&t->wideXb, (size_t)nS * nB * (t->Nz + 1) * sizeof(double)));
		PTP_CUDA(cudaMalloc(&t->wideJ, (size_t)nS * sizeof(int)));
		t->wideCap = nS;
	}
	return PTP_OK;
}

int ptp_solver_run(ptp_trap* t, const double* rho, bool rhoIsFixed, const double* dScale, int nS, double* spec, double* phi, bool withField, const uint2* encBounds, int rowLimit, int rowsWanted, int* rowsDone)
{
	if (rowsDone) *rowsDone = t->Nr;
	if (nS <= 0) return PTP_OK;
	const int n1 = t->Nz + 1, Nr = t->Nr, M = nS * Nr;
	const double fixedInv = 1.0 / (double)(1ULL << t->fixedBits);
	PTP_TRY(ptp_solver_reserve(t, nS));
	if (!encBounds) {
		const cudaError_t eb = ptp_launch(k_row_bounds, dim3((M + 7) / 8), dim3(256), 0, t->stream, t->usePdl, rho, M, n1, t->rowBounds, Nr, rowLimit < 0 ? Nr : rowLimit);
		if (eb != cudaSuccess) return ptp_cuda_fail(eb, "k_row_bounds launch", __FILE__, __LINE__);
		t->lastLaunches++;
	}
	// The step's solve for a plasma that occupies few radial rows: one cluster kernel (ptp_solve_cluster.cu)
	if (withField && rowsWanted > 0 && rowLimit >= 0 && t->solver == PTP_SOLVER_DIRECT) {
		const int rows16 = std::min(Nr, (std::max(rowsWanted, rowLimit) + 15) & ~15);
		int PM, NC, KWc, CW;
		size_t smc;
		if (rows16 < Nr && ptp_solver_cluster_plan(t, nS, rowLimit, rows16, &PM, &NC, &KWc, &CW, &smc)) {
			PTP_TRY(ptp_solver_cluster_run(t, rho, rhoIsFixed, dScale, nS, phi, encBounds, rowLimit, rows16, PM, NC, KWc, CW, smc));
			if (rowsDone) *rowsDone = rows16;
			t->eNodesValid = true;
			return PTP_OK;
		}
	}
	// Rows to produce: all of them, or - for the step, where only the populated rows are ever read by the push - the first
	// rowsWanted, rounded up to whole blocks of the radial tables (a multiple of the inverse kernels' strip height too).
	int rowsOut = Nr;
	if (rowsWanted > 0 && rowsWanted < Nr) rowsOut = std::min(Nr, (rowsWanted + PTP_THOMAS_BLOCK - 1) / PTP_THOMAS_BLOCK * PTP_THOMAS_BLOCK);
	if (rowLimit >= 0 && rowsOut < rowLimit) rowsOut = Nr;         // (a caller asking for fewer rows than may hold a deposit gets all)
	auto smFwdBytes = [&](int mb) {
		return ((size_t)3 * Nr * mb + (size_t)FWD_KB * mb + (size_t)FWD_RP * FWD_KB + (size_t)Nr) * sizeof(double) + (size_t)Nr * sizeof(int2);
	};
	const int mb = smFwdBytes(16) <= t->smemMax ? 16 : 4;       // fewer modes per CTA when the radial tiles get large
	const size_t smFwd = smFwdBytes(mb);
	// inverse transform (+ node field): paired-mode kernel when its tiles fit in shared memory, FFT for long power-of-two
	// rows, chunked GEMM otherwise
	const int stagesAll = ((n1 + 1) / 2 + 8 * INV_KS - 1) / (8 * INV_KS);    // stages that hold a warp's whole K-slice
	auto smFieldBytes = [&](int st) { return ((size_t)INV_TM * (n1 | 1) + (size_t)8 * st * INV_KS * INV_TN + (size_t)INV_TM * INV_TN) * sizeof(double); };
	const int ringStages = smFieldBytes(std::max(stagesAll, 2)) <= t->smemMax ? std::max(stagesAll, 2) : INV_ST;
	const size_t smField = smFieldBytes(ringStages);
	// even row length: every row of the grids and of the cosine matrix starts 16-byte aligned - bulk-async (TMA) staged form
	const int ldaBulk = (n1 + 15) & ~15;
	const size_t smBulk = ((size_t)INV_TM * ldaBulk + 8 + (size_t)std::max((n1 + 1) / 2, 8 * INV_TM) * INV_TN + (size_t)INV_TM * INV_TN) * sizeof(double);
	const bool bulkOk = n1 % 2 == 0 && t->invBulk && smBulk <= t->smemMax && (size_t)n1 * 8 * INV_TM < (1u << 20);
	const bool fusedFits = bulkOk || smField <= t->smemMax;
	const bool useFft = ptp_solver_fft_fits(t) && (t->solver == PTP_SOLVER_DIRECT_FFT || !fusedFits);
	if (!useFft && !fusedFits) rowsOut = Nr;                    // chunked inverse GEMM + separate node field: whole grids only
	if (rowsDone) *rowsDone = rowsOut;
	// large grids (many radial nodes or long rows): separate forward transform + streamed radial solves (ptp_solve_wide.cu)
	const bool wide = mb == 4 || useFft || smFwd > t->smemMax;
	if (wide) {
		const bool formInInverse = useFft && ptp_solver_inverse_forms_rows(t);
		PTP_TRY(ptp_solver_forward_wide(t, rho, rhoIsFixed, dScale, nS, spec, encBounds, !formInInverse, rowLimit, rowsOut));
		if (useFft) {
			PTP_TRY(ptp_solver_inverse_fft(t, spec, phi, nS, withField, !formInInverse, rowsOut));
			if (withField) t->eNodesValid = true;
			return PTP_OK;
		}
	}
	else {
		const dim3 gridFwd((n1 + mb - 1) / mb, nS);
		// the fold row of the radial solves: the outermost row that can hold a deposit
		const int Jf = rowLimit < 0 ? Nr - 1 : std::max(0, std::min(rowLimit, Nr) - 1);
		auto launchFwd = [&](auto kern, double fInv) -> cudaError_t {
			cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smFwd);
			if (e != cudaSuccess) return e;
			return ptp_launch(kern, gridFwd, dim3(256), smFwd, t->stream, t->usePdl, rho, t->rowBounds, encBounds, t->dctFwd, dScale, fInv, t->thInv, t->thCp, t->thR, t->thQ,
				t->thLower, spec, Nr, n1, Jf, rowsOut);
		};
		const cudaError_t ef = rhoIsFixed ? launchFwd(k_fwd_thomas<true, 16>, fixedInv) : launchFwd(k_fwd_thomas<false, 16>, 1.0);
		if (ef != cudaSuccess) return ptp_cuda_fail(ef, "k_fwd_thomas launch", __FILE__, __LINE__);
		t->lastLaunches++;
	}
	bool fieldDone = false;
	if (bulkOk) {
		auto launchBulk = [&](auto kern, bool field) -> cudaError_t {
			cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smBulk);
			if (e != cudaSuccess) return e;
			const dim3 grid(field ? (n1 + 1 + INV_TN - 3) / (INV_TN - 2) : (n1 + INV_TN - 1) / INV_TN, (rowsOut + INV_TM - 1) / INV_TM);
			return ptp_launch(kern, grid, dim3(256), smBulk, t->stream, t->usePdl, spec, t->dctInv, phi, t->phiTrap, t->eNodes, nS, Nr, n1, t->hz, ldaBulk);
		};
		const cudaError_t eb = withField ? launchBulk(k_inv_field_bulk<true>, true) : launchBulk(k_inv_field_bulk<false>, false);
		if (eb != cudaSuccess) return ptp_cuda_fail(eb, "k_inv_field_bulk launch", __FILE__, __LINE__);
		fieldDone = withField;
	}
	else if (smField <= t->smemMax) {
		auto launchInv = [&](auto kern, bool field) -> cudaError_t {
			cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smField);
			if (e != cudaSuccess) return e;
			const dim3 grid(field ? (n1 + 1 + INV_TN - 3) / (INV_TN - 2) : (n1 + INV_TN - 1) / INV_TN, (rowsOut + INV_TM - 1) / INV_TM);
			return ptp_launch(kern, grid, dim3(256), smField, t->stream, t->usePdl, spec, t->dctInv, phi, t->phiTrap, t->eNodes, nS, Nr, n1, t->hz, ringStages);
		};
		const bool vec = n1 % 2 == 0;
		cudaError_t ei;
		if (withField) ei = vec ? launchInv(k_inv_field<true, true>, true) : launchInv(k_inv_field<false, true>, true);
		else ei = vec ? launchInv(k_inv_field<true, false>, false) : launchInv(k_inv_field<false, false>, false);
		if (ei != cudaSuccess) return ptp_cuda_fail(ei, "k_inv_field launch", __FILE__, __LINE__);
		fieldDone = withField;
	}
	else {
		// (rows that are not a power of two and too long for the fused kernel: all rows, one GEMM over the species)
		const int kc = n1 < INV_KC ? n1 : INV_KC;
		size_t smInv = ((size_t)INV_TM * (kc | 1) + (size_t)8 * INV_ST * INV_KS * INV_TN) * sizeof(double);
		if (smInv < (size_t)8 * INV_TM * INV_TN * sizeof(double)) smInv = (size_t)8 * INV_TM * INV_TN * sizeof(double);
		const dim3 gridInv((n1 + INV_TN - 1) / INV_TN, (M + INV_TM - 1) / INV_TM);
		if (n1 % 2 == 0) {
			PTP_CUDA(cudaFuncSetAttribute(k_inv_gemm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smInv));
			k_inv_gemm<true><<<gridInv, 256, smInv, t->stream>>>(spec, t->dctInv, phi, M, n1, n1);
		}
		else {
			PTP_CUDA(cudaFuncSetAttribute(k_inv_gemm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smInv));
			k_inv_gemm<false><<<gridInv, 256, smInv, t->stream>>>(spec, t->dctInv, phi, M, n1, n1);
		}
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return ptp_cuda_fail(e, "solver launch", __FILE__, __LINE__);
	t->lastLaunches += 1;
	if (withField) {
		if (fieldDone) t->eNodesValid = true;
		else return ptp_node_field(t);
	}
	return PTP_OK;
}

int ptp_solver_apply(ptp_trap* t, const double* x, double* y)
{
	const int n1 = t->Nz + 1;
	k_apply<<<(unsigned)((t->G + 255) / 256), 256, 0, t->stream>>>(x, y, t->Nr, n1, t->stDiag, t->stHz2, t->stLower, t->stUpper);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_apply launch", __FILE__, __LINE__);
	return PTP_OK;
}

int ptp_node_field(ptp_trap* t)
{
	const int n1 = t->Nz + 1;
	k_node_field<<<(unsigned)((t->G + 255) / 256), 256, 0, t->stream>>>(t->phiTrap, t->phiSelfAll, (int)t->plasmas.size(), t->G, n1, t->hz, t->eNodes);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_node_field launch", __FILE__, __LINE__);
	t->lastLaunches++;
	t->eNodesValid = true;
	return PTP_OK;
}

int ptp_wall_rhs(ptp_trap* t, const double* dWall, double* dRhs)
{
	k_wall_rhs<<<(unsigned)((t->G + 255) / 256), 256, 0, t->stream>>>(dWall, dRhs, t->G, t->Nz + 1, t->wallFactor);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_wall_rhs launch", __FILE__, __LINE__);
	return PTP_OK;
}

// Red-black SOR on the same operator (north_star names it; here a cross-check of the direct solver: it needs
// ~1e-12 residual for 1e-10 parity, i.e. thousands of sweeps, so it is not the default).
int ptp_sor_run(ptp_trap* t, const double* rho, bool rhoIsFixed, const double* dScale, int nS, double* phi)
{
	const int n1 = t->Nz + 1;
	const double fixedInv = 1.0 / (double)(1ULL << t->fixedBits);
	// optimal-ish omega for a (Nr x Nz) Laplacian
	const double pi = 3.14159265358979323846;
	const double rhoJ = 0.5 * (std::cos(pi / (t->Nr + 1)) + std::cos(pi / (t->Nz + 1)));
	const double omega = 2.0 / (1.0 + std::sqrt(1.0 - rhoJ * rhoJ));
	const long long half = (t->G + 1) / 2;
	const unsigned blocks = (unsigned)((half + 255) / 256);
	double* dRes = nullptr;
	PTP_CUDA(cudaMalloc(&dRes, 2 * sizeof(double)));
	for (int s = 0; s < nS; ++s) {
		const double* b = rho + (size_t)s * t->G;
		double* x = phi + (size_t)s * t->G;
		PTP_CUDA(cudaMemsetAsync(dRes, 0, 2 * sizeof(double), t->stream));
		k_norm2<<<(unsigned)((t->G + 255) / 256), 256, 0, t->stream>>>(b, dScale, s, rhoIsFixed, fixedInv, t->G, dRes + 1);
		double h[2];
		PTP_CUDA(cudaMemcpyAsync(h, dRes, 2 * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
		PTP_CUDA(cudaStreamSynchronize(t->stream));
		const double bnorm2 = h[1] > 0 ? h[1] : 1.0;
		for (int it = 0; it < t->sorMaxIter; ++it) {
			const bool check = (it % 50) == 49;
			if (check) PTP_CUDA(cudaMemsetAsync(dRes, 0, sizeof(double), t->stream));
			for (int colour = 0; colour < 2; ++colour) {
				if (rhoIsFixed)
					k_sor_sweep<true><<<blocks, 256, 0, t->stream>>>(x, b, dScale, s, fixedInv, t->Nr, n1, t->stDiag, t->stHz2, t->stLower, t->stUpper, omega, colour, check ? dRes : nullptr);
				else
					k_sor_sweep<false><<<blocks, 256, 0, t->stream>>>(x, b, dScale, s, 1.0, t->Nr, n1, t->stDiag, t->stHz2, t->stLower, t->stUpper, omega, colour, check ? dRes : nullptr);
				t->lastLaunches++;
			}
			if (check) {
				PTP_CUDA(cudaMemcpyAsync(h, dRes, sizeof(double), cudaMemcpyDeviceToHost, t->stream));
				PTP_CUDA(cudaStreamSynchronize(t->stream));
				if (std::sqrt(h[0] / bnorm2) < t->sorTol) break;
			}
		}
	}
	cudaFree(dRes);
	return PTP_OK;
}
