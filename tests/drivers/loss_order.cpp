// Test program over the host class surface (tests/test_gpu_drivers.py): rings lost in SEVERAL steps of ONE multi-step
// movePlasmas call must leave the plasma's ring order - hence the row order of the history files - exactly as the reference's
// swap-with-back removal (Source/Plasma.cpp:108-118) leaves it when it is applied step by step.
//   usage: loss_order <rings.bin> <out prefix> <steps> <electrode 1 potential>
//   rings.bin: int64 n, double chargeMacro, int32 r[n], double z[n], double v[n]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "Constants.hpp"
#include "PenningTrap.hpp"
#include "Plasma.hpp"

int main(int argc, char** argv)
{
	if (argc < 5) return 2;
	std::FILE* f = std::fopen(argv[1], "rb");
	if (!f) return 3;
	std::int64_t n = 0;
	double chargeMacro = 0;
	if (std::fread(&n, 8, 1, f) != 1 || std::fread(&chargeMacro, 8, 1, f) != 1) return 4;
	std::vector<int> r((std::size_t)n);
	std::vector<double> z((std::size_t)n), v((std::size_t)n);
	if (std::fread(r.data(), 4, (std::size_t)n, f) != (std::size_t)n || std::fread(z.data(), 8, (std::size_t)n, f) != (std::size_t)n ||
	    std::fread(v.data(), 8, (std::size_t)n, f) != (std::size_t)n) return 5;
	std::fclose(f);
	// the trap of the reference's drivers (Diagnostics/A) Grid Size and Plasma Period.txt:57-69)
	std::vector<Electrode> electrodes;
	const double potentials[5] = { 0, -70, -15, -70, 0 };
	for (double p : potentials) electrodes.push_back(Electrode(0.01322, p));
	PenningTrap trap(0.01488, electrodes, std::vector<double>(4, 0.0005), 585, 128);
	Plasma electrons(trap, "Electrons", massE, -ePos);
	electrons.loadRings(r, z, v, chargeMacro, 150.0);
	trap.setPotential(1, std::atof(argv[4]));                   // lowered barrier: the fast rings leave, a few per step
	trap.movePlasmas(2e-8 / 35, std::atoi(argv[3]));            // ONE call, many steps
	trap.saveStates(0.0);
	trap.extractPlasmasHistories(argv[2]);
	std::printf("%d\n", electrons.getNumMacro());
	return 0;
}
