// K3 + K4 of a step in ONE kernel for plasmas that occupy few radial rows (the reference's default: 12 of 128):
// Plasma::solvePoisson's solver.solve (Source/Plasma.cpp:95-99) for every species and the node field of
// PenningTrap::getEField(int,int) (Source/PenningTrap.cpp:208-236), restricted to the populated rows (rings never change
// their row, Source/Plasma.hpp:22-24, so the push reads the field nowhere else; ptp_materialize_fields produces whole grids
// on demand). Same direct solve as ptp_solve.cu - forward DCT-I of the touched nodes, radial Thomas solves with the rows
// above the plasma folded into one pivot, inverse DCT-I with paired modes - organised around thread-block clusters:
//
//   * a cluster of 16 CTAs owns a range of axial nodes; CTA c of the cluster owns mode pairs (p, Nz - p), p in [c PM, (c+1) PM);
//   * every CTA runs the forward transform and the radial solves for ITS modes (a few hundred FMAs per thread; every cluster
//     repeats this, which is cheaper than a second kernel), pairs them, and forms the PARTIAL inverse transform of its modes
//     for all nodes of the cluster's range: partial[j][k] = sum over own pairs of (alpha_p +- alpha_{Nz-p}) cos(pi p k / Nz);
//   * after one cluster barrier CTA c sums the 16 partials of its share of the range straight out of the peers' shared memory
//     (distributed shared memory, fixed order: deterministic), adds the species onto the trap potential in registration
//     order and writes potentials and node field - the centred difference needs one halo node on either side, which the CTA
//     reduces itself from the same partials in the same order, so it is bit-identical to the neighbour's value.
//
// Nothing but the touched deposit nodes, ~40 KB of solver tables per CTA and the output crosses L2; the only global round
// trips on the critical path are touched-node ranges -> deposit chunk. The tables (constants) are requested before the
// programmatic-launch wait, i.e. while the push kernel is still finishing.
#include "ptp_internal.h"

#include <cooperative_groups.h>
#include <limits.h>

#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace {

// [cluster-begin] (tests/emu/emu_cluster.sh compiles the text up to [cluster-end] for the host: 16 CTAs of a cluster run
// concurrently on CPU threads, cluster.sync is a barrier over all of them, map_shared_rank a pointer translation)
constexpr int CL = 16;          // CTAs per cluster (non-portable size: needs cudaFuncAttributeNonPortableClusterSizeAllowed)
constexpr int CS_KB = 64;       // axial nodes of the deposit staged per chunk
constexpr int CS_T = 256;
constexpr int CS_MAXROWS = 64;  // rows the fast path handles (deposit rows and output rows)

struct ClusterSolveArgs {
	const double* rho;          // [nS][Nr][n1] deposit accumulators (double weights or int64 fixed point)
	const int2* bounds;         // touched node range per (species, row), or
	const uint2* encBounds;     // ... the push kernel's encoded maxima
	const double* FT;           // [n1][n1] forward DCT-I with weights
	const double* C;            // [n1][n1] cos(pi m k / Nz)
	const double* rowScale;     // [nS] rho -> RHS factor
	const double* thInv, *thCp, *thR, *thQ, *thLower;
	double* phiSelf;            // [nS][Nr][n1]
	const double* phiTrap;
	double* eNodes;
	double fixedInv, hz;
	int nS, Nr, n1, Jf, rowsOut, PM, KWc, CW;
};

__device__ __forceinline__ void cs_cp8(void* smemDst, const void* gmemSrc, bool valid)
{
	const unsigned int d = (unsigned int)__cvta_generic_to_shared(smemDst);
	const int bytes = valid ? 8 : 0;                            // src-size 0 -> zero fill
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmemSrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cs_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cs_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <bool A_FIXED>
__global__ void __launch_bounds__(CS_T, 1) k_solve_cluster(const ClusterSolveArgs a)
{
	cg::cluster_group cluster = cg::this_cluster();
	extern __shared__ __align__(16) double smc[];
	const int tid = threadIdx.x;
	const int c = (int)cluster.block_rank();
	const int cid = blockIdx.x / CL;
	const int n1 = a.n1, Nz = n1 - 1, K2 = (n1 + 1) / 2, PM = a.PM, NM = 2 * PM;
	const int Jf = a.Jf, rowsIn = Jf + 1, rowsOut = a.rowsOut, rowsT = max(rowsIn, rowsOut);
	// axial nodes of this cluster: core [kc0, kc1), with one halo node on either side [kA, kB)
	const int kc0 = cid * a.KWc, kc1 = min(kc0 + a.KWc, n1);
	const int kA = max(kc0 - 1, 0), kB = min(kc1 + 1, n1), KW = kB - kA, KWp = a.KWc + 2;
	const int CWp = a.CW + 2;

	// All species go through every phase together (species x row = one longer row index), so that the phases - each a short
	// dependent chain behind a barrier - are paid once per step, not once per species.
	const int nS = a.nS;
	double* sInv = smc;                                         // [rowsIn][NM] 1 / pivot          (rows < Jf)
	double* sCp = sInv + (size_t)rowsIn * NM;                   // [rowsT][NM]  upper / pivot (rows < Jf), thR (rows > Jf)
	double* sB = sCp + (size_t)rowsT * NM;                      // [nS][rowsT][NM]  beta -> alpha
	double* sS = sB + (size_t)nS * rowsT * NM;                  // [nS][rowsOut][NM] paired modes: sums at [s], differences at [PM + s]
	double* sFT = sS + (size_t)nS * rowsOut * NM;               // [CS_KB][NM]  chunk of the forward matrix (own modes)
	double* sRho = sFT + (size_t)CS_KB * NM;                    // [nS][rowsIn][CS_KB] chunk of the deposit rows
	double* sC = sRho + (size_t)nS * rowsIn * CS_KB;            // [PM][KWp]    cos(pi p k / Nz), own pairs x the cluster's nodes
	double* sPart = sC + (size_t)PM * KWp;                      // [nS][rowsOut][KWp] partial inverse transform of the own modes
	double* sTot = sPart + (size_t)nS * rowsOut * KWp;          // [rowsOut][CWp] total potential of this CTA's share (+ halo)
	double* sLower = sTot + (size_t)rowsOut * CWp;              // [rowsT]
	double* sScale = sLower + rowsT;                            // [nS]
	int2* sBd = reinterpret_cast<int2*>(sScale + nS);           // [nS][rowsIn]
	int* sJ0 = reinterpret_cast<int*>(sBd + (size_t)nS * rowsIn); // [nS] first touched row
	int& sLo = sJ0[nS];                                         // union of the touched node ranges of all species and rows
	int& sHi = sJ0[nS + 1];

	// mode of slot s: pair p = c PM + (s mod PM); slots [0, PM) hold mode p, slots [PM, 2 PM) hold mode Nz - p
	const int slot = tid % NM, lane6 = tid / NM, nLanes = CS_T / NM;      // forward transform: thread -> (slot, rows lane6 + i nLanes)
	auto modeOf = [&](int s, bool& ok) {
		const int p = c * PM + (s < PM ? s : s - PM);
		const int m = s < PM ? p : Nz - p;
		ok = p < K2 && !(s >= PM && Nz - p == p);                // (the self-paired middle mode of an even Nz is held once)
		return ok ? m : 0;
	};
	bool mOk;
	const int m = modeOf(slot, mOk);
	const bool worker = tid < NM * nLanes;

	// ---- constants, requested before the wait -------------------------------------------------------------------
	if (worker) {
		for (int j = lane6; j < rowsT; j += nLanes) {
			const size_t off = (size_t)j * n1 + m;
			if (j < Jf) {
				cs_cp8(&sInv[j * NM + slot], a.thInv + off, mOk);
				cs_cp8(&sCp[j * NM + slot], a.thCp + off, mOk);
			}
			else if (j > Jf) cs_cp8(&sCp[j * NM + slot], a.thR + off, mOk);
		}
	}
	for (int j = tid; j < rowsT; j += CS_T) cs_cp8(&sLower[j], a.thLower + j, true);
	for (int e = tid; e < PM * KW; e += CS_T) {
		const int s = e / KW, kk = e - s * KW;
		const int p = c * PM + s;
		cs_cp8(&sC[s * KWp + kk], a.C + (size_t)(p < K2 ? p : 0) * n1 + kA + kk, p < K2);
	}
	cs_commit();
	const double q = mOk ? a.thQ[(size_t)Jf * n1 + m] : 0.0;    // (by slot: every thread that may run a radial solve has it)
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();                                             // this step's deposit and its touched-node ranges

	// this CTA's share of the cluster's nodes: core [myK0, myK1), one halo node on either side
	const int myK0 = kc0 + c * a.CW, myK1 = min(myK0 + a.CW, kc1);          // (may be empty for the last CTAs)
	if (myK0 < myK1)
		for (int o = tid; o < rowsOut * CWp; o += CS_T) {
			const int j = o / CWp, x = o - j * CWp, k = myK0 - 1 + x;
			sTot[o] = (k >= 0 && k < n1 && x < myK1 - myK0 + 2) ? a.phiTrap[(size_t)j * n1 + k] : 0.0;
		}
	if (tid == 0) { sLo = INT_MAX; sHi = INT_MIN; }
	if (tid < nS) {
		sJ0[tid] = Jf;
		sScale[tid] = (a.rowScale ? a.rowScale[tid] : 1.0) * (A_FIXED ? a.fixedInv : 1.0);
	}
	__syncthreads();
	for (int v = tid; v < nS * rowsIn; v += CS_T) {
		const int sp = v / rowsIn, j = v - sp * rowsIn;
		int2 bd;
		if (a.encBounds) {                                      // maxima written by the push kernel's flush: (Nz+2-kmin, kmax+1), 0 = untouched
			const uint2 e = a.encBounds[sp * a.Nr + j];
			bd = e.y ? make_int2(n1 + 1 - (int)e.x, (int)e.y - 1) : make_int2(INT_MAX, INT_MIN);
		}
		else bd = a.bounds[sp * a.Nr + j];
		sBd[v] = bd;
		if (bd.x <= bd.y) { atomicMin(&sLo, bd.x); atomicMax(&sHi, bd.y); atomicMin(&sJ0[sp], j); }
	}
	__syncthreads();
	const int kLo = sLo, kHi = sHi;
	// ---- forward DCT-I of the touched nodes for the own modes: beta[sp][j][slot] ----------------------------------
	// (Every loop of this kernel is kept rolled: the code runs once per SM and step, so it is instruction-fetch bound -
	// ncu shows "no instruction" as the top stall of every solve kernel - and short loops are what the fetch unit can keep up with.)
	if (worker)
		for (int v = lane6; v < nS * rowsT; v += nLanes) sB[v * NM + slot] = 0.0;
	for (int k0 = kLo; k0 <= kHi; k0 += CS_KB) {
		const int kn = min(CS_KB, kHi - k0 + 1);
		__syncthreads();                                        // the previous chunk has been consumed
#pragma unroll 1
		for (int e = tid; e < kn * NM; e += CS_T) {
			const int kk = e / NM, s = e - kk * NM;
			bool ok;
			const int mm = modeOf(s, ok);
			cs_cp8(&sFT[e], a.FT + (size_t)(k0 + kk) * n1 + mm, ok);
		}
#pragma unroll 1
		for (int e = tid; e < nS * rowsIn * kn; e += CS_T) {
			const int v = e / kn, kk = e - v * kn;
			const int sp = v / rowsIn, j = v - sp * rowsIn;
			const int2 bd = sBd[v];
			const bool ok = bd.x <= bd.y && k0 + kk >= bd.x && k0 + kk <= bd.y;   // exact zeros elsewhere
			cs_cp8(&sRho[v * CS_KB + kk], a.rho + ((size_t)sp * a.Nr + j) * n1 + k0 + kk, ok);
		}
		cs_commit();
		cs_wait_all();
		__syncthreads();
		if (worker) {
#pragma unroll 1
			for (int v = lane6; v < nS * rowsIn; v += nLanes) {
				const int sp = v / rowsIn, j = v - sp * rowsIn;
				const int2 bd = sBd[v];
				if (bd.x > bd.y) continue;                              // row without a deposit (the arithmetic below would overflow)
				const int a0 = max(bd.x, k0) - k0, a1 = min(bd.y, k0 + kn - 1) - k0;
				double t = 0.0;
#pragma unroll 2
				for (int kk = a0; kk <= a1; ++kk) {
					const double val = A_FIXED ? (double)reinterpret_cast<const long long*>(sRho)[v * CS_KB + kk] : sRho[v * CS_KB + kk];
					t = fma(val, sFT[kk * NM + slot], t);
				}
				sB[(sp * rowsT + j) * NM + slot] += t * sScale[sp];  // (own element)
			}
		}
	}
	cs_wait_all();                                              // (also the tables requested before the wait)
	__syncthreads();
	// ---- radial solves: forward sweep over the touched rows below Jf, folded pivot at Jf, back-substitution, rows above ----
	if (worker && mOk) {
#pragma unroll 1
		for (int sp = lane6; sp < nS; sp += nLanes) {
			double* bb = sB + (size_t)sp * rowsT * NM + slot;
			const int J0 = sJ0[sp];
			double y = 0.0;
#pragma unroll 1
			for (int j = J0; j < Jf; ++j) {
				const double inv = sInv[j * NM + slot];
				y = fma(-(sLower[j] * inv), y, bb[j * NM] * inv);
				bb[j * NM] = y;
			}
			const double xJ = (bb[Jf * NM] - sLower[Jf] * y) * q;
			bb[Jf * NM] = xJ;
			y = xJ;
#pragma unroll 1
			for (int j = Jf - 1; j >= 0; --j) {
				y = fma(-sCp[j * NM + slot], y, bb[j * NM]);
				bb[j * NM] = y;
			}
			y = xJ;
#pragma unroll 1
			for (int j = Jf + 1; j < rowsOut; ++j) {
				y = sCp[j * NM + slot] * y;
				bb[j * NM] = y;
			}
		}
	}
	__syncthreads();
	// ---- pair the modes: cos(pi (Nz - p) k / Nz) = (-1)^k cos(pi p k / Nz) -----------------------------------
	for (int e = tid; e < nS * rowsOut * PM; e += CS_T) {
		const int v = e / PM, s = e - v * PM, p = c * PM + s;   // v = sp * rowsOut + j
		const int sp = v / rowsOut, j = v - sp * rowsOut;
		const double* bb = sB + ((size_t)sp * rowsT + j) * NM;
		double plus = 0.0, minus = 0.0;
		if (p < K2) {
			const double ap = bb[s];
			if (Nz - p == p) plus = minus = ap;
			else { const double aq = bb[PM + s]; plus = ap + aq; minus = ap - aq; }
		}
		sS[v * NM + s] = plus;
		sS[v * NM + PM + s] = minus;
	}
	__syncthreads();
	// ---- partial inverse transform of the own pairs for every node of the cluster's range ---------------------
	{
		const int ng = CS_T / KW;                               // row groups: thread -> (node kk, rows g, g + ng, ...)
		const int g = tid / KW, kk = tid - g * KW;
		if (g < ng) {
			const double* cv = sC + kk;
			const int sel = ((kA + kk) & 1) ? PM : 0;
#pragma unroll 1
			for (int v = g; v < nS * rowsOut; v += ng) {
				const double* sv = sS + v * NM + sel;
				double t = 0.0;
#pragma unroll 2
				for (int s2 = 0; s2 < PM; ++s2) t = fma(sv[s2], cv[s2 * KWp], t);
				sPart[v * KWp + kk] = t;
			}
		}
	}
	cluster.sync();                                             // every CTA's partials are in place
	// ---- sum the 16 partials of this CTA's share (+ halo) out of the peers' shared memory --------------------
	if (myK0 < myK1) {
		const int nx = myK1 - myK0 + 2;
		for (int o = tid; o < rowsOut * nx; o += CS_T) {
			const int j = o / nx, x = o - j * nx, k = myK0 - 1 + x;
			if (k < 0 || k >= n1) continue;
			const int kk = k - kA;
			double tot = sTot[j * CWp + x];
#pragma unroll 1
			for (int sp = 0; sp < nS; ++sp) {                       // species added in registration order (Source/PenningTrap.cpp:226-232)
				const int off = (sp * rowsOut + j) * KWp + kk;
				double v = 0.0;
#pragma unroll 4
				for (int r = 0; r < CL; ++r) v += cluster.map_shared_rank(sPart, r)[off];
				if (x >= 1 && x <= myK1 - myK0) a.phiSelf[((size_t)sp * a.Nr + j) * n1 + k] = v;
				tot = __dadd_rn(tot, v);
			}
			sTot[j * CWp + x] = tot;
		}
	}
	cluster.sync();                                             // the peers are done reading this CTA's partials
	// ---- node field: E = (Phi[k-1] - Phi[k+1]) / (2 hz), zero at both ends (Source/PenningTrap.cpp:218-233) ----------
	if (myK0 < myK1) {
		const int nc = myK1 - myK0;
		for (int o = tid; o < rowsOut * nc; o += CS_T) {
			const int j = o / nc, x = 1 + (o - j * nc), k = myK0 - 1 + x;
			double e = 0.0;
			if (k > 0 && k < n1 - 1) e = __ddiv_rn(__dsub_rn(sTot[j * CWp + x - 1], sTot[j * CWp + x + 1]), __dmul_rn(2.0, a.hz));
			a.eNodes[(size_t)j * n1 + k] = e;
		}
	}
}

// [cluster-end]

size_t cluster_smem_bytes(int nS, int Jf, int rowsOut, int PM, int KWc, int CW)
{
	const size_t S = (size_t)nS, NM = 2 * (size_t)PM, rowsIn = (size_t)Jf + 1, rowsT = std::max<size_t>(rowsIn, rowsOut);
	const size_t doubles = rowsIn * NM + rowsT * NM + S * rowsT * NM + S * (size_t)rowsOut * NM + (size_t)CS_KB * NM + S * rowsIn * CS_KB + (size_t)PM * (KWc + 2) +
		S * (size_t)rowsOut * (KWc + 2) + (size_t)rowsOut * (CW + 2) + rowsT + S;
	return doubles * sizeof(double) + S * rowsIn * sizeof(int2) + (S + 2) * sizeof(int) + 64;
}

} // namespace

// Can (and should) this solve go through the cluster kernel? All species, node field wanted, few rows, tables fit.
bool ptp_solver_cluster_plan(const ptp_trap* t, int nS, int rowLimit, int rowsOut, int* PMout, int* NCout, int* KWcOut, int* CWout, size_t* smemOut)
{
	if (!t->clusterSolve || rowLimit < 0 || rowLimit > CS_MAXROWS || rowsOut > CS_MAXROWS || rowsOut < 1) return false;
	const int n1 = t->Nz + 1, K2 = (n1 + 1) / 2;
	const int PM = (K2 + CL - 1) / CL;
	if (2 * PM > 85) return false;                              // (forward transform: at least 3 row lanes of 2 PM threads)
	int NC = std::min(8, std::max(1, (n1 + 63) / 64));          // clusters over the axial nodes: >= ~64 nodes each, at most 8 (one per GPC)
	const int KWc = (n1 + NC - 1) / NC;
	NC = (n1 + KWc - 1) / KWc;
	const int CW = (KWc + CL - 1) / CL;
	if (KWc + 2 > CS_T) return false;                           // (partial inverse: one thread per node of the cluster's range)
	const int Jf = std::max(0, std::min(rowLimit, t->Nr) - 1);
	const size_t smem = cluster_smem_bytes(nS, Jf, rowsOut, PM, KWc, CW);
	if (smem > t->smemMax) return false;
	*PMout = PM; *NCout = NC; *KWcOut = KWc; *CWout = CW; *smemOut = smem;
	return true;
}

int ptp_solver_cluster_run(ptp_trap* t, const double* rho, bool rhoIsFixed, const double* dScale, int nS, double* phi, const uint2* encBounds, int rowLimit, int rowsOut,
	int PM, int NC, int KWc, int CW, size_t smem)
{
	ClusterSolveArgs a{};
	a.rho = rho; a.bounds = t->rowBounds; a.encBounds = encBounds;
	a.FT = t->dctFwd; a.C = t->dctInv; a.rowScale = dScale;
	a.thInv = t->thInv; a.thCp = t->thCp; a.thR = t->thR; a.thQ = t->thQ; a.thLower = t->thLower;
	a.phiSelf = phi; a.phiTrap = t->phiTrap; a.eNodes = t->eNodes;
	a.fixedInv = 1.0 / (double)(1ULL << t->fixedBits); a.hz = t->hz;
	a.nS = nS; a.Nr = t->Nr; a.n1 = t->Nz + 1; a.Jf = std::max(0, std::min(rowLimit, t->Nr) - 1); a.rowsOut = rowsOut;
	a.PM = PM; a.KWc = KWc; a.CW = CW;
	auto kern = rhoIsFixed ? k_solve_cluster<true> : k_solve_cluster<false>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_solve_cluster attributes", __FILE__, __LINE__);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(CL * NC);
	cfg.blockDim = dim3(CS_T);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = t->stream;
	cudaLaunchAttribute attr[2];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[1].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = t->usePdl ? 2 : 1;
	e = cudaLaunchKernelEx(&cfg, kern, a);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_solve_cluster launch", __FILE__, __LINE__);
	t->lastLaunches++;
	return PTP_OK;
}
