// Host emulation of the default-grid Poisson solve: the table builder of ptp_solver_build ([tables-begin]..[tables-end]) and
// the kernels of pic-trapped-plasma_b200/csrc/ptp_solve.cu between [solve-begin] and [solve-end] - k_row_bounds,
// k_fwd_thomas, k_inv_field (or k_inv_gemm + k_node_field when the fused tiles do not fit), k_apply, k_wall_rhs - compiled
// unchanged for the CPU and launched with the grid / shared-memory arithmetic of ptp_solver_run.
//   usage: emu_solve <case.bin> <out.bin>   case: Nz Nr (int32) hz hr radius (f64) rho[G] phiTrap[G] wall[Nz+1]
//                                           out: phi[G] eNodes[G] Aphi[G] wallRhs[G]
#include "cuda_host_shim.h"

#define PTP_THOMAS_BLOCK 32
static unsigned char* g_smem;
static inline void cp_async8(void* dst, const void* src, bool valid) { if (valid) std::memcpy(dst, src, 8); else std::memset(dst, 0, 8); }
static inline void cp_async16(void* dst, const void* src, bool valid) { if (valid) std::memcpy(dst, src, 16); else std::memset(dst, 0, 16); }
// bulk-async copies complete at once here; the alignment rules of cp.async.bulk are checked
static inline void mbar_init(unsigned long long*, int) {}
static inline void mbar_fence_init() {}
static inline void mbar_expect_tx(unsigned long long*, unsigned int) {}
static inline void mbar_wait(unsigned long long*, unsigned int) { __syncthreads(); }   // every thread of the CTA waits on the same phase: a CTA barrier orders it behind the (synchronous) copies
static inline void fence_proxy_async() {}
static inline void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long*)
{
	if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (bytes & 15) || bytes == 0) { std::fprintf(stderr, "bulk_g2s: misaligned copy\n"); std::abort(); }
	std::memcpy(dst, src, bytes);
}
#include "solve_snippet.inc"

struct FakeTrap { int Nz, Nr; double hz, hr, radius, stDiag, stHz2, wallFactor; };

int main(int argc, char** argv)
{
	if (argc < 3) return 2;
	FILE* f = std::fopen(argv[1], "rb");
	if (!f) return 3;
	int dims[2];
	double geo[3];
	if (std::fread(dims, 4, 2, f) != 2 || std::fread(geo, 8, 3, f) != 3) return 4;
	FakeTrap trap{ dims[0], dims[1], geo[0], geo[1], geo[2], 0, 0, 0 };
	FakeTrap* t = &trap;
	const long long G = (long long)(t->Nz + 1) * t->Nr;
	std::vector<double> rho(G), phiTrap(G), wall(t->Nz + 1);
	if (std::fread(rho.data(), 8, G, f) != (size_t)G || std::fread(phiTrap.data(), 8, G, f) != (size_t)G ||
	    std::fread(wall.data(), 8, wall.size(), f) != wall.size()) return 5;
	std::fclose(f);

#include "tables_snippet.inc"
	(void)thP; (void)hr; (void)hr2;
	const size_t smemMax = 232448;
	std::vector<unsigned char> smem;                            // exactly the dynamic shared memory each launch requests (AddressSanitizer)
	auto dynSmem = [&](size_t bytes) { smem.assign(bytes, 0); g_smem = smem.data(); };
	const int nS = 1, M = nS * Nr;
	// optional: rows that may hold a deposit (the outermost populated row + 1) and rows the caller wants, as ptp_solver_run gets them
	const int rowLimit = argc > 3 ? std::atoi(argv[3]) : -1, rowsWanted = argc > 4 ? std::atoi(argv[4]) : 0;
	int rowsOut = Nr;
	if (rowsWanted > 0 && rowsWanted < Nr) rowsOut = std::min(Nr, (rowsWanted + PTP_THOMAS_BLOCK - 1) / PTP_THOMAS_BLOCK * PTP_THOMAS_BLOCK);
	if (rowLimit >= 0 && rowsOut < rowLimit) rowsOut = Nr;
	const int Jf = rowLimit < 0 ? Nr - 1 : std::max(0, std::min(rowLimit, Nr) - 1);

	// ---- as ptp_solver_run ----
	std::vector<int2> bounds(M);
	emu_launch((M + 7) / 8, 256, [&] { k_row_bounds(rho.data(), M, n1, bounds.data(), Nr, Nr); });
	auto smFwdBytes = [&](int mb) { return ((size_t)3 * Nr * mb + (size_t)FWD_KB * mb + (size_t)FWD_RP * FWD_KB + (size_t)Nr) * sizeof(double) + (size_t)Nr * sizeof(int2); };
	if (smFwdBytes(16) > smemMax) { std::printf("emu_solve: grid belongs to the large-grid path (emu_wide)\n"); return 6; }
	std::vector<double> spec(G, 0.0), phi(G, 0.0), eN(G, -1.0);
	dynSmem(smFwdBytes(16));
	emu_launch((n1 + 15) / 16, 256, [&] {
		blockIdx.y = 0;
		k_fwd_thomas<false, 16>(rho.data(), bounds.data(), nullptr, fwd.data(), nullptr, 1.0, thInv.data(), thCp.data(), thR.data(), thQ.data(), lower.data(), spec.data(), Nr, n1, Jf, rowsOut);
	});
	const int stagesAll = ((n1 + 1) / 2 + 8 * INV_KS - 1) / (8 * INV_KS);
	auto smFieldBytes = [&](int st) { return ((size_t)INV_TM * (n1 | 1) + (size_t)8 * st * INV_KS * INV_TN + (size_t)INV_TM * INV_TN) * sizeof(double); };
	const int ringStages = smFieldBytes(std::max(stagesAll, 2)) <= smemMax ? std::max(stagesAll, 2) : INV_ST;
	const bool vec = n1 % 2 == 0;
	const int ldaBulk = (n1 + 15) & ~15;
	const size_t smBulk = ((size_t)INV_TM * ldaBulk + 8 + (size_t)std::max((n1 + 1) / 2, 8 * INV_TM) * INV_TN + (size_t)INV_TM * INV_TN) * sizeof(double);
	const bool noBulk = std::getenv("PTP_INV_BULK") && std::atoi(std::getenv("PTP_INV_BULK")) == 0;
	const bool bulkOk = vec && !noBulk && smBulk <= smemMax && (size_t)n1 * 8 * INV_TM < (1u << 20);
	const char* path;
	if (!bulkOk && smFieldBytes(ringStages) > smemMax && rowsOut < Nr) { std::printf("emu_solve: the chunked inverse produces whole grids only\n"); return 8; }
	if (bulkOk) {
		path = "k_inv_field_bulk";
		dynSmem(smBulk + 16);
		// (the kernel's extern array is 16-byte aligned on the device; same here)
		g_smem = reinterpret_cast<unsigned char*>(((uintptr_t)g_smem + 15) & ~(uintptr_t)15);
		const int gx = (n1 + 1 + INV_TN - 3) / (INV_TN - 2), gy = (rowsOut + INV_TM - 1) / INV_TM;
		for (int by = 0; by < gy; ++by)
			emu_launch(gx, 256, [&] {
				blockIdx.y = by;
				k_inv_field_bulk<true>(spec.data(), inv.data(), phi.data(), phiTrap.data(), eN.data(), nS, Nr, n1, hz, ldaBulk);
			});
	}
	else if (smFieldBytes(ringStages) <= smemMax) {
		path = "k_inv_field";
		dynSmem(smFieldBytes(ringStages));
		const int gx = (n1 + 1 + INV_TN - 3) / (INV_TN - 2), gy = (rowsOut + INV_TM - 1) / INV_TM;
		for (int by = 0; by < gy; ++by)
			emu_launch(gx, 256, [&] {
				blockIdx.y = by;
				if (vec) k_inv_field<true, true>(spec.data(), inv.data(), phi.data(), phiTrap.data(), eN.data(), nS, Nr, n1, hz, ringStages);
				else k_inv_field<false, true>(spec.data(), inv.data(), phi.data(), phiTrap.data(), eN.data(), nS, Nr, n1, hz, ringStages);
			});
	}
	else {
		path = "k_inv_gemm + k_node_field";
		{
			const int kc = n1 < INV_KC ? n1 : INV_KC;
			size_t smInv = ((size_t)INV_TM * (kc | 1) + (size_t)8 * INV_ST * INV_KS * INV_TN) * sizeof(double);
			if (smInv < (size_t)8 * INV_TM * INV_TN * sizeof(double)) smInv = (size_t)8 * INV_TM * INV_TN * sizeof(double);
			dynSmem(smInv);
		}
		const int gx = (n1 + INV_TN - 1) / INV_TN, gy = (M + INV_TM - 1) / INV_TM;
		for (int by = 0; by < gy; ++by)
			emu_launch(gx, 256, [&] {
				blockIdx.y = by;
				if (vec) k_inv_gemm<true>(spec.data(), inv.data(), phi.data(), M, n1, n1);
				else k_inv_gemm<false>(spec.data(), inv.data(), phi.data(), M, n1, n1);
			});
		emu_launch((int)((G + 255) / 256), 256, [&] { k_node_field(phiTrap.data(), phi.data(), nS, G, n1, hz, eN.data()); });
	}
	// operator and wall right-hand side
	std::vector<double> Aphi(G, 0.0), wallRhs(G, 0.0);
	emu_launch((int)((G + 255) / 256), 256, [&] { k_apply(phi.data(), Aphi.data(), Nr, n1, t->stDiag, t->stHz2, lower.data(), upper.data()); });
	emu_launch((int)((G + 255) / 256), 256, [&] { k_wall_rhs(wall.data(), wallRhs.data(), G, n1, t->wallFactor); });

	f = std::fopen(argv[2], "wb");
	if (!f) return 7;
	std::fwrite(phi.data(), 8, G, f);
	std::fwrite(eN.data(), 8, G, f);
	std::fwrite(Aphi.data(), 8, G, f);
	std::fwrite(wallRhs.data(), 8, G, f);
	std::fclose(f);
	std::printf("emu_solve: %d x %d grid, inverse through %s (%d ring stages), fold row %d, %d rows produced\n", Nz, Nr, path, ringStages, Jf, rowsOut);
	return 0;
}
