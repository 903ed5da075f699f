// Check of the hot-species form of the push kernel through the C ABI, without Python (a few seconds of GPU time):
//  * validation: 20 M electrons on the c5 grid (4096 x 1024), fixed-point deposits, 40 steps - the per-warp-bin form against
//    the thread-private form without re-sorts: deposit grid bit for bit, rings (id, z, v) bit for bit (order-independent
//    checksum), alive counts;
//  * timing: 50 M electrons, fp64 deposits, hot from the load, 100 steps after 60 of warm-up.
// (profiles/r02_hot_ab_harness.txt: the run that decided between two sorting forms, selected by PTP_SCATTER_FORM then.)
// Inputs: build/hot_ab/c5.bin (tools/hot_ab_prepare.py: wall potentials and the non-zero nodes of the expected density).
// Build: g++ -O2 -std=c++17 tools/hot_ab.cpp -Iinclude -Lpic-trapped-plasma_b200 -lptp_b200 -Wl,-rpath,'$ORIGIN/../../pic-trapped-plasma_b200' -o build/hot_ab/hot_ab
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ptp.h"

#define CK(call)                                                                         \
	do {                                                                                 \
		if ((call) != PTP_OK) { std::printf("FAILED %s: %s\n", #call, ptp_last_error()); std::exit(1); } \
	} while (0)

struct Inputs {
	int64_t Nz, Nr, nnz;
	double hz, hr, length, radius, dt, temperature, mass, charge;
	std::vector<double> wall, dens;
};

static Inputs load(const char* path)
{
	Inputs in;
	FILE* f = std::fopen(path, "rb");
	if (!f) { std::printf("cannot open %s\n", path); std::exit(2); }
	int64_t h[3];
	double d[8];
	if (std::fread(h, 8, 3, f) != 3 || std::fread(d, 8, 8, f) != 8) std::exit(3);
	in.Nz = h[0]; in.Nr = h[1]; in.nnz = h[2];
	in.hz = d[0]; in.hr = d[1]; in.length = d[2]; in.radius = d[3]; in.dt = d[4]; in.temperature = d[5]; in.mass = d[6]; in.charge = d[7];
	in.wall.resize(in.Nz + 1);
	std::vector<int64_t> idx(in.nnz);
	std::vector<double> val(in.nnz);
	if (std::fread(in.wall.data(), 8, in.Nz + 1, f) != (size_t)(in.Nz + 1) || std::fread(idx.data(), 8, in.nnz, f) != (size_t)in.nnz ||
	    std::fread(val.data(), 8, in.nnz, f) != (size_t)in.nnz) std::exit(4);
	std::fclose(f);
	in.dens.assign((size_t)(in.Nz + 1) * in.Nr, 0.0);
	for (int64_t i = 0; i < in.nnz; ++i) in.dens[idx[i]] = val[i];
	return in;
}

static uint64_t mix(uint64_t x)
{
	x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
	return x;
}

struct Result {
	std::vector<double> rhs;
	uint64_t ringSum = 0;
	int64_t alive = 0, sorts = 0;
	int hot = 0;
	double msPerStep = 0;
};

static Result run(const Inputs& in, int hot, int mode, int64_t rings, int warm, int steps, bool check)
{
	ptp_trap* t = nullptr;
	CK(ptp_trap_create(&t, (int)in.Nz, (int)in.Nr, in.hz, in.hr, in.length, in.radius, 0));
	CK(ptp_trap_set_wall(t, in.wall.data()));
	CK(ptp_trap_set_deposit_mode(t, mode));
	if (hot == 0) CK(ptp_trap_set_sort_interval(t, 0));
	ptp_plasma* p = nullptr;
	CK(ptp_plasma_create(t, &p, in.mass, in.charge));
	CK(ptp_plasma_set_hot(p, hot));
	double chargeMacro = 0;
	int64_t loaded = 0;
	CK(ptp_plasma_load_density(p, in.dens.data(), in.temperature, rings, 0, 1, &chargeMacro, nullptr, &loaded));
	CK(ptp_plasma_deposit_solve(p));
	if (warm) CK(ptp_trap_step(t, in.dt, warm));
	CK(ptp_trap_step(t, in.dt, steps));
	CK(ptp_trap_sync(t));
	double ms[4];
	CK(ptp_trap_last_times(t, ms));
	Result r;
	r.msPerStep = ms[0] / steps;
	r.hot = ptp_plasma_is_hot(p);
	r.sorts = ptp_trap_sorts_done(t);
	CK(ptp_plasma_count(p, &r.alive));
	if (check) {
		r.rhs.resize(in.dens.size());
		CK(ptp_plasma_get_rhs(p, r.rhs.data()));
		std::vector<int32_t> rr(r.alive);
		std::vector<double> z(r.alive), v(r.alive);
		std::vector<int64_t> id(r.alive);
		CK(ptp_plasma_download(p, rr.data(), z.data(), v.data(), id.data()));
		for (int64_t i = 0; i < r.alive; ++i) {
			uint64_t zb, vb;
			std::memcpy(&zb, &z[i], 8); std::memcpy(&vb, &v[i], 8);
			r.ringSum += mix((uint64_t)id[i] * 0x9e3779b97f4a7c15ULL ^ mix(zb) ^ mix(vb + 0x1234567ULL) ^ (uint64_t)rr[i] << 48);
		}
	}
	std::printf("  hot %d (in use %d) mode %s rings %lld alive %lld sorts %lld: %.4f ms per step over %d steps\n", hot, r.hot, mode ? "fixed" : "fp64",
		(long long)loaded, (long long)r.alive, (long long)r.sorts, r.msPerStep, steps);
	std::fflush(stdout);
	CK(ptp_plasma_destroy(p));
	CK(ptp_trap_destroy(t));
	return r;
}

int main(int argc, char** argv)
{
	const Inputs in = load(argc > 1 ? argv[1] : "build/hot_ab/c5.bin");
	std::printf("hot_ab: grid %lld x %lld, dt %.3e\n", (long long)in.Nz, (long long)in.Nr, in.dt);
	std::printf("validation (20 M electrons, fixed point, 40 steps):\n");
	const Result w = run(in, 0, PTP_DEPOSIT_FIXED64, 20000000, 0, 40, true);
	const Result a = run(in, 1, PTP_DEPOSIT_FIXED64, 20000000, 0, 40, true);
	const bool same = a.alive == w.alive && a.ringSum == w.ringSum && a.rhs.size() == w.rhs.size() && std::memcmp(a.rhs.data(), w.rhs.data(), w.rhs.size() * 8) == 0;
	std::printf("  hot form == thread-private form, bit for bit (grid, rings, counts): %s\n", same ? "yes" : "NO");
	std::printf("timing (50 M electrons, fp64 deposits, hot from the load, 60 warm-up + 100 timed steps):\n");
	const Result t1 = run(in, 1, PTP_DEPOSIT_FP64, 50000000, 60, 100, false);
	std::printf("the same left to the re-sort policy (ptp_plasma_set_hot(-1)):\n");
	const Result t2 = run(in, -1, PTP_DEPOSIT_FP64, 50000000, 60, 100, false);
	std::printf("RESULT hot form %.4f ms/step, by policy %.4f ms/step (%lld re-sorts, hot form in use %d); valid %d\n", t1.msPerStep, t2.msPerStep, (long long)t2.sorts, t2.hot, (int)same);
	return same ? 0 : 5;
}
