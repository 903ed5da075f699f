// Host emulation of K5 (k_sort_count / k_sort_scan / k_sort_scatter / k_sort_pad): the kernel text between the
// [emu-begin] / [emu-end] markers of pic-trapped-plasma_b200/csrc/ptp_particles.cu on CPU threads, driven the way
// ptp_sort_plasma drives it (chunk table over the live prefixes, re-sort into already used alternate buffers).
// Self-checking: exits 0 when every row comes out ordered by axial cell with its ring multiset intact, the live counts
// are right and every slot behind the live prefix holds the empty-slot pattern.
#include "cuda_host_shim.h"

static unsigned char* g_smem;
static inline int exact_cell(double z, double hz, int Nz)
{
	int k = (int)floor(__ddiv_rn(z, hz));
	return k > Nz - 1 ? Nz - 1 : k;
}
#include "sort_snippet.inc"

#include <algorithm>
#include <map>
#include <random>

int main()
{
	const int Nz = 300, Nr = 5;
	const double hz = 1e-4;
	const long long bucket = 3 * 4096;
	std::vector<long long> rowOff(Nr + 1), rowLive(Nr), altDirty(Nr, 0);
	for (int r = 0; r <= Nr; ++r) rowOff[r] = r * bucket;
	const long long cap = rowOff[Nr];
	const double nan = __longlong_as_double(-1LL);
	std::vector<double> z(cap, nan), v(cap, 0.0), zA(cap, nan), vA(cap, 0.0);
	std::vector<long long> id(cap, -1), idA(cap, -1);
	std::mt19937_64 rng(5);
	std::uniform_real_distribution<double> U(0.0, 1.0);
	const long long live0[Nr] = { 9000, 0, 4097, 12288, 33 };
	long long next = 0;
	for (int r = 0; r < Nr; ++r) {
		rowLive[r] = live0[r];
		for (long long i = 0; i < live0[r]; ++i) {
			const long long s = rowOff[r] + i;
			z[s] = hz * (100 + 60 * U(rng) + 0.001 * i);         // nearly ordered, as the loaders leave them, + scatter
			if (z[s] >= Nz * hz) z[s] = Nz * hz * 0.999;
			v[s] = 1000.0 * U(rng); id[s] = next++;
		}
	}
	std::vector<unsigned char> smem(((Nz + 1) & ~1) * 4 + Nz * 8 + 64);
	g_smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
	int rc = 0;
	for (int round = 0; round < 3 && rc == 0; ++round) {
		// losses between sorts: tombstones inside the live prefix
		for (int r = 0; r < Nr; ++r)
			for (long long i = 0; i < rowLive[r]; ++i)
				if (U(rng) < 0.07) z[rowOff[r] + i] = nan;
		std::map<long long, std::pair<double, double>> want;       // id -> (z, v) of the survivors
		std::vector<long long> wantLive(Nr, 0);
		for (int r = 0; r < Nr; ++r)
			for (long long i = 0; i < rowLive[r]; ++i) {
				const long long s = rowOff[r] + i;
				if (z[s] == z[s]) { want[id[s]] = { z[s], v[s] }; ++wantLive[r]; }
			}
		std::vector<SortChunk> chunks;
		for (int r = 0; r < Nr; ++r)
			for (long long b = 0; b < rowLive[r]; b += SORT_CHUNK)
				chunks.push_back(SortChunk{ r, 0, rowOff[r] + b, rowOff[r] + std::min<long long>(rowLive[r], b + SORT_CHUNK) });
		std::vector<unsigned int> counts((size_t)Nr * Nz, 0);
		std::vector<unsigned long long> cursor((size_t)Nr * Nz, 0), live(Nr, 0);
		emu_launch((int)chunks.size(), 256, [&] { k_sort_count(z.data(), chunks.data(), Nz, hz, counts.data()); });
		emu_launch(Nr, 256, [&] { k_sort_scan(counts.data(), Nz, cursor.data(), live.data()); });
		// a fourth per-ring array (the speeds at the last save point) must travel with its ring: here twice the speed
		std::vector<double> vs(v.size()), vsA(v.size(), -7.0);
		for (size_t s = 0; s < v.size(); ++s) vs[s] = 2.0 * v[s];
		emu_launch((int)chunks.size(), 256, [&] { k_sort_scatter(z.data(), v.data(), id.data(), zA.data(), vA.data(), idA.data(), rowOff.data(), chunks.data(), Nz, hz, cursor.data(), vs.data(), vsA.data()); });
		emu_launch(4, 256, [&] { k_sort_pad(zA.data(), vA.data(), idA.data(), rowOff.data(), live.data(), altDirty.data()); }, Nr);
		std::swap(z, zA); std::swap(v, vA); std::swap(id, idA);
		for (int r = 0; r < Nr; ++r) { altDirty[r] = rowLive[r]; rowLive[r] = (long long)live[r]; }
		for (int r = 0; r < Nr && rc == 0; ++r)
			for (long long i = 0; i < (long long)live[r]; ++i)
				if (vsA[rowOff[r] + i] != 2.0 * v[rowOff[r] + i]) { std::printf("round %d row %d slot %lld: saved speed did not travel with its ring\n", round, r, i); rc = 1; break; }
		size_t seen = 0;
		for (int r = 0; r < Nr && rc == 0; ++r) {
			if ((long long)live[r] != wantLive[r]) { std::printf("round %d row %d: live %llu, expected %lld\n", round, r, live[r], wantLive[r]); rc = 1; }
			int prev = -1;
			for (long long i = 0; i < bucket && rc == 0; ++i) {
				const long long s = rowOff[r] + i;
				if (i < (long long)live[r]) {
					const int k = exact_cell(z[s], hz, Nz);
					auto it = want.find(id[s]);
					if (!(z[s] == z[s]) || k < prev || it == want.end() || it->second.first != z[s] || it->second.second != v[s]) { std::printf("round %d row %d slot %lld: wrong ring\n", round, r, i); rc = 1; }
					prev = k; ++seen;
				}
				else if (z[s] == z[s] || v[s] != 0.0 || id[s] != -1) { std::printf("round %d row %d slot %lld: padding not clean\n", round, r, i); rc = 1; }
			}
		}
		if (rc == 0 && seen != want.size()) { std::printf("round %d: %zu rings out, %zu in\n", round, seen, want.size()); rc = 1; }
		if (rc == 0) std::printf("round %d: %zu rings ordered by cell in %d rows, padding clean\n", round, seen, Nr);
	}
	return rc;
}
