// Host emulation of K1 (k_push_deposit): the kernel text between the [emu-begin] / [emu-end] markers of
// pic-trapped-plasma_b200/csrc/ptp_push.cu compiled unchanged for the CPU (tests/emu/cuda_host_shim.h) and run CTA by CTA
// on 512 std::threads. Reads a case file written by tests/test_kernel_models.py, writes the kernel's outputs next to it.
//   usage: emu_push <case.bin> <out.bin>
#include "cuda_host_shim.h"

// must match ptp_internal.h
struct PtpSegment {
	int row;
	int pad;
	long long begin, end;
};
static_assert(sizeof(PtpSegment) == 24, "PtpSegment layout");

static unsigned char* g_smem;
#define PTP_HOST_EMU 1
#include "push_snippet.inc"

struct CaseHeader {
	int Nz, Nr, W, WE, fixed, exact, fixedBits, segTiles, nCta, scatter;
	long long n;
	double hz, length, dt, charge, mass;
};

template <bool FIXED, bool EXACT> static void run(const PushArgs& a, int nCta)
{
	if (a.scatter) emu_launch(nCta, 512, [&] { k_push_deposit<512, 4, true, FIXED, EXACT, true>(a); });
	else emu_launch(nCta, 512, [&] { k_push_deposit<512, 4, true, FIXED, EXACT>(a); });
}

int main(int argc, char** argv)
{
	if (argc < 3) return 2;
	FILE* f = std::fopen(argv[1], "rb");
	if (!f) return 3;
	CaseHeader h;
	if (std::fread(&h, sizeof(h), 1, f) != 1) return 4;
	const int n1 = h.Nz + 1;
	const long long G = (long long)n1 * h.Nr;
	std::vector<double> eNodes(G), z(h.n), v(h.n);
	std::vector<int> r(h.n);
	if (std::fread(eNodes.data(), 8, G, f) != (size_t)G || std::fread(r.data(), 4, h.n, f) != (size_t)h.n ||
	    std::fread(z.data(), 8, h.n, f) != (size_t)h.n || std::fread(v.data(), 8, h.n, f) != (size_t)h.n) return 5;
	std::fclose(f);

	// row buckets padded with NaN slots to whole tiles (2048 slots)
	const long long tile = 4 * 512;
	std::vector<long long> count(h.Nr, 0), rowOff(h.Nr + 1, 0);
	for (long long i = 0; i < h.n; ++i) ++count[r[i]];
	for (int j = 0; j < h.Nr; ++j) rowOff[j + 1] = rowOff[j] + (count[j] + tile - 1) / tile * tile;
	const long long cap = rowOff[h.Nr];
	const double nan = std::nan("");
	std::vector<double> bz(cap, nan), bv(cap, 0.0);
	std::vector<long long> slotOf(h.n), cur(rowOff.begin(), rowOff.end() - 1);
	for (long long i = 0; i < h.n; ++i) { const long long d = cur[r[i]]++; bz[d] = z[i]; bv[d] = v[i]; slotOf[i] = d; }

	// segments of at most segTiles tiles, dealt to the CTAs in contiguous ranges; windows from the true cell range
	std::vector<PtpSegment> segs;
	std::vector<int4> bounds;
	for (int j = 0; j < h.Nr; ++j)
		for (long long b = rowOff[j]; b < rowOff[j] + count[j]; b += h.segTiles * tile) {
			PtpSegment s{ j, 0, b, std::min(rowOff[j + 1], b + h.segTiles * tile) };
			int lo = INT_MAX, hi = INT_MIN, nl = 0;
			long long sum = 0;
			for (long long i = s.begin; i < s.end; ++i)
				if (bz[i] == bz[i]) {
					int k = (int)std::floor(bz[i] / h.hz);
					if (k > h.Nz - 1) k = h.Nz - 1;
					lo = std::min(lo, k); hi = std::max(hi, k); sum += k; ++nl;
				}
			segs.push_back(s);
			bounds.push_back(make_int4(lo, hi, nl ? (int)(sum / nl) : 0, 0));
		}
	const int nCta = std::max(1, std::min<int>(h.nCta, (int)segs.size()));
	std::vector<int> ctaSegBegin(nCta + 1);
	for (int c = 0; c <= nCta; ++c) ctaSegBegin[c] = (int)((long long)segs.size() * c / nCta);

	std::vector<double> rho(G + h.Nr, 0.0);                  // grid + per-row touched-node range (two u32 per row)
	unsigned long long lost[2] = { 0, 0 };
	const size_t perBin = h.fixed ? 8 : 10;
	// = ptp_push_smem_bytes / ptp_push_scatter_smem_bytes, to the byte (AddressSanitizer)
	std::vector<unsigned char> smem(h.scatter ? (size_t)h.W * (16 + 16 + 8 * (512 / 32)) : (size_t)h.WE * 16 + (size_t)h.W * 16 + (size_t)h.W * 512 * perBin);
	g_smem = smem.data();

	PushArgs a{};
	a.Nz = h.Nz; a.W = h.W; a.fixedBits = h.fixedBits; a.WE = h.WE;
	a.hz = h.hz; a.invHz = 1.0 / h.hz; a.eps = (h.Nz + 2) * 1e-15; a.epsHi = 1.0 - a.eps; a.length = h.length;
	a.dt = h.dt; a.charge = h.charge; a.mass = h.mass; a.invMass = 1.0 / h.mass;
	a.fixedScale = (double)(1ULL << h.fixedBits);
	a.invFixedScale = 1.0 / a.fixedScale;
	a.eNodes = eNodes.data(); a.z = bz.data(); a.v = bv.data();
	a.segs = segs.data(); a.ctaSegBegin = ctaSegBegin.data(); a.segBounds = bounds.data();
	a.rho[0] = rho.data(); a.nRho = 1; a.pad1 = 0; a.scatter = h.scatter; a.bndOffset = G; a.lost = lost;
	if (h.fixed) { if (h.exact) run<true, true>(a, nCta); else run<true, false>(a, nCta); }
	else { if (h.exact) run<false, true>(a, nCta); else run<false, false>(a, nCta); }

	std::vector<double> zo(h.n), vo(h.n);
	for (long long i = 0; i < h.n; ++i) { zo[i] = bz[slotOf[i]]; vo[i] = bv[slotOf[i]]; }
	f = std::fopen(argv[2], "wb");
	if (!f) return 6;
	const long long nSeg = (long long)segs.size();
	std::fwrite(&nSeg, 8, 1, f);
	std::fwrite(lost, 8, 2, f);
	std::fwrite(zo.data(), 8, h.n, f);
	std::fwrite(vo.data(), 8, h.n, f);
	std::fwrite(rho.data(), 8, G + h.Nr, f);
	std::fwrite(bounds.data(), sizeof(int4), segs.size(), f);
	std::fclose(f);
	std::printf("emu_push: %lld rings, %lld segments on %d CTAs, lost %llu, out-of-window deposits %llu\n", h.n, nSeg, nCta, lost[0], lost[1]);
	return 0;
}
