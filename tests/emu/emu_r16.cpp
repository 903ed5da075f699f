// Host emulation of k_idct_r16_field: the kernel text between the [r16-begin] / [r16-end] markers of
// pic-trapped-plasma_b200/csrc/ptp_solve_wide.cu is compiled unchanged for the CPU (CUDA keywords shimmed below) and run
// as one CTA of 256 std::threads with a std::barrier for __syncthreads(). Checks phi against the DCT-I sum in long double
// and the node field against its definition. Build + run: tests/emu/emu_r16.sh
#include "cuda_host_shim.h"

static double2* fbw;   // "extern __shared__ double2 fbw[];" in the kernel becomes a use of this pointer

#include "r16_snippet.inc"

int main()
{
	const int N = 4096, n1 = N + 1, Nr = 3, nS = 2, row = 1;
	std::vector<double2> tw(N);
	const long double pi = 3.141592653589793238462643383279502884L;
	for (int j = 0; j < N; ++j) { const long double a = -pi * j / N; tw[j] = make_double2((double)cosl(a), (double)sinl(a)); }
	std::vector<double> alpha((size_t)nS * Nr * n1), phi((size_t)nS * Nr * n1, 0.0), phiTrap((size_t)Nr * n1), eN((size_t)Nr * n1, -1.0);
	srand(7);
	for (auto& v : alpha) v = (rand() / (double)RAND_MAX - 0.5) * 2;
	for (auto& v : phiTrap) v = (rand() / (double)RAND_MAX - 0.5) * 50;
	std::vector<double2> smem(16 * 257 + (N + 1) / 2 + 8);
	fbw = smem.data();
	const double hz = 1.6625e-5;
	// second species: its row is not read from alpha but formed as (value entering the row's block) x (in-block prefix
	// product), the way the kernel does for rows above the outermost deposit row; alpha keeps the product for the reference
	const int blockRows = 1, nB = Nr;
	std::vector<double> xb((size_t)nS * nB * n1, 0.0), thP((size_t)Nr * n1, 0.0);
	std::vector<int> wideJ = { Nr - 1, 0 };                     // species 0: every row read; species 1: rows > 0 formed
	for (int k = 0; k <= N; ++k) {
		thP[(size_t)row * n1 + k] = 0.5 + rand() / (double)RAND_MAX;
		xb[((size_t)1 * nB + row) * n1 + k] = (rand() / (double)RAND_MAX - 0.5) * 2;
		alpha[((size_t)1 * Nr + row) * n1 + k] = xb[((size_t)1 * nB + row) * n1 + k] * thP[(size_t)row * n1 + k];
	}
	std::vector<double> alphaSeen = alpha;
	for (int k = 0; k <= N; ++k) alphaSeen[((size_t)1 * Nr + row) * n1 + k] = 1e300;   // must not be read
	emu_launch(1, 256, [&] {
		blockIdx.x = row;
		k_idct_r16_field<true>(alphaSeen.data(), phi.data(), tw.data(), phiTrap.data(), eN.data(), nS, Nr, hz, xb.data(), wideJ.data(), thP.data(), blockRows);
	});
	// reference: phi_k = sum_m a_m cos(pi m k / N) in long double with exact argument reduction (m k mod 2N)
	double worst = 0, norm = 0, err = 0;
	std::vector<long double> tot(n1);
	for (int k = 0; k <= N; ++k) tot[k] = phiTrap[(size_t)row * n1 + k];
	for (int sp = 0; sp < nS; ++sp) {
		const double* a = &alpha[((size_t)sp * Nr + row) * n1];
		for (int k = 0; k <= N; ++k) {
			long double s = 0;
			for (int m = 0; m <= N; ++m) s += a[m] * cosl(pi * (long double)(((long long)m * k) % (2 * N)) / N);
			const double got = phi[((size_t)sp * Nr + row) * n1 + k];
			err += (double)((got - s) * (got - s)); norm += (double)(s * s);
			if (fabs((double)(got - s)) > worst) worst = fabs((double)(got - s));
			tot[k] += got;
		}
	}
	double eerr = 0;
	for (int k = 0; k <= N; ++k) {
		const double want = (k > 0 && k < N) ? (double)((tot[k - 1] - tot[k + 1]) / (2 * hz)) : 0.0;
		const double got = eN[(size_t)row * n1 + k];
		eerr = fmax(eerr, fabs(got - want) / (1 + fabs(want)));
	}
	// rows other than `row` must be untouched
	bool clean = true;
	for (int r = 0; r < Nr; ++r) if (r != row) for (int k = 0; k <= N; ++k) clean &= eN[(size_t)r * n1 + k] == -1.0 && phi[(size_t)r * n1 + k] == 0.0;
	printf("phi rel-L2 %.3e  max abs %.3e  field rel %.3e  other rows untouched: %s\n", sqrt(err / norm), worst, eerr, clean ? "yes" : "NO");
	return (sqrt(err / norm) < 1e-14 && eerr < 1e-9 && clean) ? 0 : 1;
}
