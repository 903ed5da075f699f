// Host emulation of the device-side loader (ptp_plasma_load_density, pic-trapped-plasma_b200/csrc/ptp_load.cu): its host
// arithmetic ([counts-begin]..[counts-end]: cumulative charge per row, chargeMacro, rings per row, quantum) and the placement
// kernel k_place ([place-begin]..[place-end]) on host threads, fed with the deviate stream of libstdc++'s own
// std::normal_distribution (k_rng_* are checked against it in emu_rng.cpp). Writes r, z, v in load order for comparison with
// Plasma::loadDensityFile of the compiled reference (Source/Plasma.cpp:558-622).
//   usage: emu_place <case.bin> <out.bin>   case: Nz Nr shard nShards (int32) numMacro (int64) hz hr temperature mass (f64) density[G]
#include "cuda_host_shim.h"
#include <random>

#define PTP_EINVAL 1
static void ptp_set_error(const char*) {}
#include "place_snippet.inc"

struct FakeTrap { int Nz, Nr; double hz, hr; };

int main(int argc, char** argv)
{
	if (argc < 3) return 2;
	FILE* f = std::fopen(argv[1], "rb");
	if (!f) return 3;
	int dims[4];
	long long numMacro;
	double geo[4];
	if (std::fread(dims, 4, 4, f) != 4 || std::fread(&numMacro, 8, 1, f) != 1 || std::fread(geo, 8, 4, f) != 4) return 4;
	FakeTrap trap{ dims[0], dims[1], geo[0], geo[1] };
	FakeTrap* t = &trap;
	const int shard = dims[2], nShards = dims[3];
	const double temperature = geo[2], mass = geo[3];
	const long long G = (long long)(t->Nz + 1) * t->Nr;
	std::vector<double> dens(G);
	if (std::fread(dens.data(), 8, G, f) != (size_t)G) return 5;
	std::fclose(f);
	const double* density = dens.data();
	long long* nAtRow = nullptr;

#include "counts_snippet.inc"
	(void)mcd;
	// layout as ptp_plasma_set_layout: buckets padded to 4096 slots
	std::vector<long long> rowOff(Nr + 1, 0);
	for (int j = 0; j < Nr; ++j) rowOff[j + 1] = rowOff[j] + (count[j] + 4095) / 4096 * 4096;
	for (PlaceRow& pr : rows) pr.slot0 = rowOff[pr.row];
	const long long cap = rowOff[Nr];
	std::vector<double> z(cap, std::nan("")), v(cap, 0.0), normals(total);
	std::vector<long long> id(cap, -1);
	{
		std::default_random_engine eng;                         // Source/Plasma.cpp:602-603: a fresh engine per load
		std::normal_distribution<double> dist(0.0, 1.0);
		for (long long i = 0; i < total; ++i) normals[i] = dist(eng);
	}
	const double sigma = std::sqrt(KB * temperature / mass);  // :509
	long long maxLocal = 0;
	for (const PlaceRow& pr : rows) maxLocal = std::max(maxLocal, pr.local);
	const int gx = (int)std::min<long long>((maxLocal + 255) / 256, 64);
	if (!rows.empty())
		emu_launch(gx, 256, [&] { k_place(rows.data(), cum.data(), normals.data(), n1, hz, sigma, shard, nShards, z.data(), v.data(), id.data()); }, (int)rows.size());

	f = std::fopen(argv[2], "wb");
	if (!f) return 6;
	std::fwrite(&local, 8, 1, f);
	std::fwrite(&chargeMacro, 8, 1, f);
	std::fwrite(perRow.data(), 8, Nr, f);
	for (int j = 0; j < Nr; ++j)
		for (long long q = 0; q < count[j]; ++q) { const int r = j; std::fwrite(&r, 4, 1, f); }
	for (int j = 0; j < Nr; ++j) std::fwrite(z.data() + rowOff[j], 8, count[j], f);
	for (int j = 0; j < Nr; ++j) std::fwrite(v.data() + rowOff[j], 8, count[j], f);
	for (int j = 0; j < Nr; ++j) std::fwrite(id.data() + rowOff[j], 8, count[j], f);
	std::fclose(f);
	std::printf("emu_place: %lld of %lld rings (shard %d of %d) in %zu rows\n", local, total, shard, nShards, rows.size());
	return 0;
}
