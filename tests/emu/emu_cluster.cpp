// Host emulation of the one-kernel step solve (pic-trapped-plasma_b200/csrc/ptp_solve_cluster.cu, [cluster-begin]..[cluster-end]):
// the 16 CTAs of a thread-block cluster run CONCURRENTLY on CPU threads (16 x 256 std::threads), cluster.sync() is a barrier
// over all of them and cluster.map_shared_rank() a pointer translation between the CTAs' shared-memory buffers; clusters run
// one after the other. Launch arithmetic as ptp_solver_cluster_plan / ptp_solver_cluster_run.
//   usage: emu_cluster <case.bin> <out.bin> <rowLimit> [nS]
//   case: Nz Nr (int32) hz hr radius (f64) rho[nS][G] phiTrap[G] ; out: phi[nS][G] eNodes[G] (rows >= rowsOut untouched: phi 0, eNodes -1)
#include "cuda_host_shim.h"

#define PTP_THOMAS_BLOCK 32
static thread_local unsigned char* g_smem;
static inline void cs_cp8(void* dst, const void* src, bool valid) { if (valid) std::memcpy(dst, src, 8); else std::memset(dst, 0, 8); }
static inline void cs_commit() {}
static inline void cs_wait_all() {}

struct EmuClusterShared {
	std::barrier<> bar;
	unsigned char* base[16];
	explicit EmuClusterShared(int n) : bar(n) {}
};
static thread_local EmuClusterShared* g_cluster = nullptr;
static thread_local int g_clusterRank = 0;
struct EmuCluster {
	unsigned int block_rank() const { return (unsigned int)g_clusterRank; }
	void sync() const { g_cluster->bar.arrive_and_wait(); }
	template <class T> T* map_shared_rank(T* p, int r) const
	{
		return reinterpret_cast<T*>(g_cluster->base[r] + (reinterpret_cast<unsigned char*>(p) - g_smem));
	}
};
#include "cluster_snippet.inc"

struct FakeTrap { int Nz, Nr; double hz, hr, radius, stDiag, stHz2, wallFactor; };

int main(int argc, char** argv)
{
	if (argc < 4) return 2;
	const int rowLimit = std::atoi(argv[3]), nS = argc > 4 ? std::atoi(argv[4]) : 1;
	FILE* f = std::fopen(argv[1], "rb");
	if (!f) return 3;
	int dims[2];
	double geo[3];
	if (std::fread(dims, 4, 2, f) != 2 || std::fread(geo, 8, 3, f) != 3) return 4;
	FakeTrap trap{ dims[0], dims[1], geo[0], geo[1], geo[2], 0, 0, 0 };
	FakeTrap* t = &trap;
	const long long G = (long long)(t->Nz + 1) * t->Nr;
	std::vector<double> rho((size_t)nS * G), phiTrap(G);
	if (std::fread(rho.data(), 8, rho.size(), f) != rho.size() || std::fread(phiTrap.data(), 8, G, f) != (size_t)G) return 5;
	std::fclose(f);

#include "tables_snippet.inc"
	(void)thP; (void)hr; (void)hr2; (void)upper;
	// touched node range per (species, row), as k_row_bounds finds it
	std::vector<int2> bounds((size_t)nS * Nr);
	for (int v = 0; v < nS * Nr; ++v) {
		int lo = INT_MAX, hi = INT_MIN;
		for (int k = 0; k < n1; ++k)
			if (rho[(size_t)v * n1 + k] != 0.0) { lo = std::min(lo, k); hi = std::max(hi, k); }
		bounds[v] = make_int2(lo, hi);
	}
	// ---- as ptp_solver_cluster_plan ----
	const int rows16 = std::min(Nr, (rowLimit + 15) & ~15);
	const int K2 = (n1 + 1) / 2, PM = (K2 + CL - 1) / CL;
	int NC = std::min(8, std::max(1, (n1 + 63) / 64));
	const int KWc = (n1 + NC - 1) / NC;
	NC = (n1 + KWc - 1) / KWc;
	const int CW = (KWc + CL - 1) / CL;
	const int Jf = std::max(0, std::min(rowLimit, Nr) - 1);
	if (rowLimit > CS_MAXROWS || rows16 > CS_MAXROWS || 2 * PM > 85 || KWc + 2 > CS_T) { std::printf("emu_cluster: not a case for the cluster kernel\n"); return 6; }
	const size_t S = (size_t)nS, NM = 2 * (size_t)PM, rowsIn = (size_t)Jf + 1, rowsT = std::max<size_t>(rowsIn, rows16);
	const size_t smem = (rowsIn * NM + rowsT * NM + S * rowsT * NM + S * (size_t)rows16 * NM + (size_t)CS_KB * NM + S * rowsIn * CS_KB + (size_t)PM * (KWc + 2) +
		S * (size_t)rows16 * (KWc + 2) + (size_t)rows16 * (CW + 2) + rowsT + S) * sizeof(double) + S * rowsIn * sizeof(int2) + (S + 2) * sizeof(int) + 64;
	std::vector<double> scale(nS, 1.0), phi((size_t)nS * G, 0.0), eN(G, -1.0);
	ClusterSolveArgs a{};
	a.rho = rho.data(); a.bounds = bounds.data(); a.encBounds = nullptr;
	a.FT = fwd.data(); a.C = inv.data(); a.rowScale = scale.data();
	a.thInv = thInv.data(); a.thCp = thCp.data(); a.thR = thR.data(); a.thQ = thQ.data(); a.thLower = lower.data();
	a.phiSelf = phi.data(); a.phiTrap = phiTrap.data(); a.eNodes = eN.data();
	a.fixedInv = 1.0; a.hz = hz;
	a.nS = nS; a.Nr = Nr; a.n1 = n1; a.Jf = Jf; a.rowsOut = rows16; a.PM = PM; a.KWc = KWc; a.CW = CW;
	blockDim.x = CS_T; gridDim.x = CL * NC; gridDim.y = 1;
	for (int cid = 0; cid < NC; ++cid) {
		EmuClusterShared shared(CL * CS_T);
		std::vector<std::vector<unsigned char>> bufs(CL, std::vector<unsigned char>(smem + 16, 0));   // exactly the bytes the launch requests (AddressSanitizer)
		std::vector<std::unique_ptr<std::barrier<>>> bars;
		std::vector<std::unique_ptr<EmuWarp>> warps;
		for (int c = 0; c < CL; ++c) {
			shared.base[c] = reinterpret_cast<unsigned char*>(((uintptr_t)bufs[c].data() + 15) & ~(uintptr_t)15);
			bars.emplace_back(new std::barrier<>(CS_T));
			for (int w = 0; w < CS_T / 32; ++w) warps.emplace_back(new EmuWarp);
		}
		std::vector<std::thread> th;
		for (int c = 0; c < CL; ++c)
			for (int tt = 0; tt < CS_T; ++tt)
				th.emplace_back([&, c, tt, cid] {
					threadIdx.x = tt; blockIdx.x = cid * CL + c; blockIdx.y = 0;
					g_ctaBarrier = bars[c].get();
					g_warp = warps[c * (CS_T / 32) + tt / 32].get(); g_lane = tt & 31;
					g_smem = shared.base[c]; g_cluster = &shared; g_clusterRank = c;
					k_solve_cluster<false>(a);
				});
		for (auto& x : th) x.join();
	}
	f = std::fopen(argv[2], "wb");
	if (!f) return 7;
	std::fwrite(phi.data(), 8, phi.size(), f);
	std::fwrite(eN.data(), 8, G, f);
	std::fclose(f);
	std::printf("emu_cluster: %d x %d grid, %d species, %d clusters of %d CTAs, %d mode pairs per CTA, fold row %d, %d rows produced, %zu bytes of shared memory per CTA\n",
		Nz, Nr, nS, NC, CL, PM, Jf, rows16, smem);
	return 0;
}
