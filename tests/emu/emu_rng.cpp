// Host emulation of the loader's deviate stream (k_rng_count / k_rng_scan / k_rng_emit, pic-trapped-plasma_b200/csrc/
// ptp_load.cu): the parallel reproduction of std::default_random_engine + std::normal_distribution<double> (what
// Plasma::loadProfile / loadDensityFile draw their speeds from, reference Source/Plasma.cpp:508-509,602-603) against
// libstdc++'s own objects. On the host both sides use the same log(), so the streams must agree bit for bit.
#include "cuda_host_shim.h"
#include "rng_snippet.inc"

#include <random>

int main()
{
	int rc = 0;
	for (long long total : { 1LL, 2LL, 4097LL, 100001LL }) {
		const long long pairs = (total + 1) / 2;
		const long long nAttempts = (long long)std::ceil((double)pairs * 1.2732395447351628 * 1.002) + 4096;   // as ptp_plasma_load_density
		const long long nThreads = (nAttempts + RNG_CH - 1) / RNG_CH;
		const int nBlocks = (int)((nThreads + 255) / 256);
		std::vector<unsigned int> blockCount(nBlocks, 0);
		std::vector<unsigned long long> blockOffset(nBlocks + 1, 0);
		std::vector<double> normals(total, -777.0);
		emu_launch(nBlocks, 256, [&] { k_rng_count(nAttempts, blockCount.data()); });
		emu_launch(1, 1024, [&] { k_rng_scan(blockCount.data(), blockOffset.data(), nBlocks); });
		if ((long long)blockOffset[nBlocks] < pairs) { std::printf("total %lld: only %llu accepted pairs\n", total, blockOffset[nBlocks]); rc = 1; continue; }
		emu_launch(nBlocks, 256, [&] { k_rng_emit(nAttempts, blockOffset.data(), normals.data(), total); });
		std::default_random_engine eng;                         // minstd_rand0, seed 1: a fresh engine per load, as the reference
		std::normal_distribution<double> dist(0.0, 1.0);
		long long bad = 0;
		for (long long i = 0; i < total; ++i) {
			const double want = dist(eng);
			if (normals[i] != want) { if (!bad) std::printf("total %lld: deviate %lld is %.17g, libstdc++ gives %.17g\n", total, i, normals[i], want); ++bad; }
		}
		if (bad) rc = 1;
		else std::printf("total %lld: %lld deviates identical to std::normal_distribution over minstd_rand0 (%d CTAs, %llu accepted pairs)\n", total, total, nBlocks, blockOffset[nBlocks]);
	}
	return rc;
}
