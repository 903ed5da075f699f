// Host emulation of the large-grid Poisson solve: the table builder of ptp_solver_build (pic-trapped-plasma_b200/csrc/
// ptp_solve.cu, [tables-begin]..[tables-end]) and the kernels of ptp_solve_wide.cu ([wide-begin]..[r16-end]) - k_fwd_dct,
// k_thomas_wide, k_thomas_expand, k_idct_fft_field / k_idct_r16_field - compiled unchanged for the CPU and run CTA by CTA
// on host threads (tests/emu/cuda_host_shim.h), in the order ptp_solver_run launches them.
//   usage: emu_wide <case.bin> <out.bin>     case: Nz Nr (int32) hz hr radius (f64) rho[G] phiTrap[G]; out: phi[G] eNodes[G] phiFormed[G]
#include "cuda_host_shim.h"

#define PTP_THOMAS_BLOCK 32
static unsigned char* g_smem;
static double2* fbw;
#include "wide_snippet.inc"

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
	std::vector<double> rho(G), phiTrap(G);
	if (std::fread(rho.data(), 8, G, f) != (size_t)G || std::fread(phiTrap.data(), 8, G, f) != (size_t)G) return 5;
	std::fclose(f);

#include "tables_snippet.inc"
	// from here on: Nz, Nr, n1, hz, lower, fwd, inv, thInv, thCp, thR, thQ, thP are the shipped builder's own values
	(void)inv; (void)hr; (void)hr2;
	const int N = Nz;
	int bits = 0;
	while ((1 << bits) < N) ++bits;
	if ((1 << bits) != N) { std::printf("emu_wide: Nz must be a power of two\n"); return 6; }
	std::vector<double2> tw((size_t)Nz);
	for (int j = 0; j < Nz; ++j) {
		const long double ang = -pi * (long double)j / (long double)Nz;
		tw[j] = make_double2((double)cosl(ang), (double)sinl(ang));
	}
	// touched node range per row, as k_row_bounds finds it
	std::vector<int2> bounds(Nr);
	for (int j = 0; j < Nr; ++j) {
		int lo = INT_MAX, hi = INT_MIN;
		for (int k = 0; k < n1; ++k)
			if (rho[(size_t)j * n1 + k] != 0.0) { lo = std::min(lo, k); hi = std::max(hi, k); }
		bounds[j] = make_int2(lo, hi);
	}
	// dynamic shared memory: exactly the bytes the launch code requests for each kernel (heap blocks, so that
	// AddressSanitizer sees an access past the end)
	std::vector<unsigned char> smem;
	auto dynSmem = [&](size_t bytes) { smem.assign(bytes, 0); g_smem = smem.data(); fbw = reinterpret_cast<double2*>(g_smem); };
	std::vector<double> spec(G, 0.0);
	// optional: rows that may hold a deposit and rows the caller wants, as ptp_solver_run gets them
	const int rowLimit = argc > 3 ? std::atoi(argv[3]) : -1, rowsWanted = argc > 4 ? std::atoi(argv[4]) : 0;
	int rowsOut = Nr;
	if (rowsWanted > 0 && rowsWanted < Nr) rowsOut = std::min(Nr, (rowsWanted + TW_BLK - 1) / TW_BLK * TW_BLK);
	if (rowLimit >= 0 && rowsOut < rowLimit) rowsOut = Nr;
	const int rowsIn = rowLimit < 0 ? Nr : std::max(1, std::min(rowLimit, Nr));
	const int nB = (Nr + TW_BLK - 1) / TW_BLK;
	std::vector<double> xb((size_t)nB * n1, 0.0);
	std::vector<int> wideJ(1, -7);

	// forward transform of the touched rows: grid (modes / 64, rows / 32, species)
	dynSmem((size_t)2 * (FD_K * FD_M + FD_K * FD_RP) * sizeof(double));
	for (int by = 0; by < (rowsIn + FD_R - 1) / FD_R; ++by)
		emu_launch((n1 + FD_M - 1) / FD_M, 256, [&] {
			blockIdx.y = by;
			k_fwd_dct<false>(rho.data(), bounds.data(), nullptr, fwd.data(), nullptr, 1.0, spec.data(), Nr, n1);
		});
	// radial solves: grid (modes / 32, species), 128 threads
	dynSmem((size_t)3 * TW_RCAP * 32 * sizeof(double) + (size_t)Nr * sizeof(double) + (size_t)Nr);
	emu_launch((n1 + 31) / 32, 128, [&] {
		blockIdx.y = 0;
		k_thomas_wide(spec.data(), bounds.data(), nullptr, thInv.data(), thCp.data(), thR.data(), thQ.data(), thP.data(), lower.data(),
			xb.data(), wideJ.data(), Nr, n1, rowsOut);
	});
	std::vector<double> specLazy = spec;                        // rows above the deposit's block still unset here
	for (int by = 0; by < (rowsOut + 7) / 8; ++by)
		emu_launch((n1 + 255) / 256, 256, [&] {
			blockIdx.y = by;
			k_thomas_expand(spec.data(), xb.data(), wideJ.data(), thP.data(), Nr, n1);
		});
	// inverse transform + node field, one CTA per row
	std::vector<double> phi(G, 0.0), eN(G, 0.0), phiFormed(G, 0.0), eN2(G, 0.0);
	const int threads = N >= 1024 ? 256 : 128;
	if (N == R16_N) {
		dynSmem((size_t)16 * R16_RS * sizeof(double2) + (size_t)(N + 1) * sizeof(double));
		emu_launch(rowsOut, 256, [&] { k_idct_r16_field<true>(spec.data(), phi.data(), tw.data(), phiTrap.data(), eN.data(), 1, Nr, hz, nullptr, nullptr, nullptr, TW_BLK); });
		for (int j = 0; j < Nr; ++j)                            // rows the inverse must form itself: poison what expand would have written
			if (wideJ[0] >= 0 && j > std::min(Nr - 1, (wideJ[0] / TW_BLK) * TW_BLK + TW_BLK - 1))
				for (int k = 0; k < n1; ++k) specLazy[(size_t)j * n1 + k] = 1e300;
		emu_launch(rowsOut, 256, [&] { k_idct_r16_field<true>(specLazy.data(), phiFormed.data(), tw.data(), phiTrap.data(), eN2.data(), 1, Nr, hz, xb.data(), wideJ.data(), thP.data(), TW_BLK); });
	}
	else {
		dynSmem((size_t)N * sizeof(double2) + (size_t)(N + 1) * sizeof(double));
		emu_launch(rowsOut, threads, [&] { k_idct_fft_field<true>(spec.data(), phi.data(), tw.data(), phiTrap.data(), eN.data(), 1, Nr, N, bits, hz); });
		phiFormed = phi;
	}
	f = std::fopen(argv[2], "wb");
	if (!f) return 7;
	std::fwrite(phi.data(), 8, G, f);
	std::fwrite(eN.data(), 8, G, f);
	std::fwrite(phiFormed.data(), 8, G, f);
	std::fclose(f);
	std::printf("emu_wide: %d x %d grid, outermost deposit row %d, %s inverse, %d rows produced\n", Nz, Nr, wideJ[0], N == R16_N ? "radix-16" : "radix-2", rowsOut);
	return 0;
}
