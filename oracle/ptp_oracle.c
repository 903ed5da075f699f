/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's PIC step.
 *
 * CPU oracle for the hot path PenningTrap::movePlasmas(dt) of
 * Daniel32Duque/PIC-Trapped-Plasma: deposit -> Poisson solve -> node E -> gather
 * -> push/loss. Every function cites the reference lines (paths relative to
 * /root/reference) whose arithmetic it restates, expression by expression and
 * in the same evaluation order, in IEEE double (build with -ffp-contract=off).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * call this. The product library (pic-trapped-plasma_b200/csrc) never does.
 *
 * Third-party arithmetic: the reference solves A x = b with
 * Eigen::SparseLU<SparseMatrix<double>, COLAMDOrdering<int>> (Source/PenningTrap.hpp:50;
 * call sites Source/PenningTrap.cpp:56-57,202 and Source/Plasma.cpp:98,389,412).
 * Eigen is absent from /root/reference (not vendored, not pinned; the sources are
 * dated Mar-Jul 2020, when Eigen 3.3.7 was current) and from this image. SparseLU
 * is a direct sparse LU factorisation; restated here as a direct banded LU of
 * the same matrix in z-major ordering (bandwidth Nr), no pivoting (-A is a row
 * diagonally dominant M-matrix).
 *
 * Pinning: validated (tests/test_oracle.py) against (i) oracle/_ref = the
 * reference's own .cpp files compiled unmodified, phase by phase, bit-exact for
 * deposit/push/gather/node-E and <=1e-11 rel-L2 for solves; (ii) the reference's
 * only data file Diagnostics/Charge_Density-0.txt (tests/golden/charge_density_0.txt)
 * to 2.7e-14; (iii) the golden vectors in tests/golden generated from (i).
 * The Eigen boundary itself is "parity unpinned": the reference holds no test
 * that pins solver output.
 */
#include "ptp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Source/Constants.hpp:12,15 */
static const double epsilon = 8.8541878128e-12;

struct ptpo_trap {
	double trapRadius;
	int nElectrodes, nGaps;
	double *electrodeLength, *electrodePotential, *gaps;
	int Nz, Nr;
	double hr, hz, lengthTrap;
	/* banded LU of the matrix in z-major ordering p = k*Nr + j; row p stores columns p-Nr .. p+Nr */
	double* band;
};

#define G_OF(t) ((t)->Nz * (t)->Nr + (t)->Nr)

/* One row of the matrix built by PenningTrap::generateSparse (Source/PenningTrap.cpp:94-162) for grid
 * node (j = radial index, k = axial index), flat index i = (Nz+1)*j + k. Outputs the five stencil
 * coefficients; an absent neighbour gets 0. */
static void stencil(const ptpo_trap* t, int j, int k, double* diag, double* zLeft, double* zRight,
	double* rLower, double* rUpper)
{
	int Nz = t->Nz, Nr = t->Nr;
	double hz2 = pow(t->hz, -2);                       /* :97 */
	double hr2 = pow(t->hr, -2);                       /* :98 */
	*diag = -2 * (hr2 + hz2);                          /* :101,108,114,138,146,154,160 */
	if (k == 0) { *zLeft = 0; *zRight = 2 * hz2; }     /* :102,126,147 */
	else if (k == Nz) { *zLeft = 2 * hz2; *zRight = 0; } /* :113,130,159 */
	else { *zLeft = hz2; *zRight = hz2; }              /* :107,109,135-136,153,155 */
	if (j == 0) {                                      /* :103,110,116 */
		*rLower = 0;
		*rUpper = 2 * hr2;
	}
	else if (j < Nr - 1) {                             /* :119-140 */
		int i = (Nz + 1) * j + k;
		double radius = floor(i / (Nz + 1)) * t->hr;   /* :121 (integer division, then floor) */
		*rLower = hr2 - pow(2 * radius * t->hr, -1);   /* :122 */
		*rUpper = hr2 + pow(2 * radius * t->hr, -1);   /* :139 */
	}
	else {                                             /* :142-160, Dirichlet row */
		double radius = t->trapRadius - t->hr;         /* :142 */
		*rLower = hr2 - pow(2 * radius * t->hr, -1);   /* :144,152,158 */
		*rUpper = 0;                                   /* wall value goes to the RHS (:163-198) */
	}
}

/* y = A x, A as assembled by generateSparse (used like coefficients * potentialsVector, :250). */
void ptpo_matrix_apply(const ptpo_trap* t, const double* x, double* y)
{
	int Nz = t->Nz, Nr = t->Nr;
	for (int j = 0; j < Nr; ++j)
		for (int k = 0; k <= Nz; ++k) {
			double d, zl, zr, rl, ru;
			int i = (Nz + 1) * j + k;
			stencil(t, j, k, &d, &zl, &zr, &rl, &ru);
			double s = d * x[i];
			if (k > 0) s += zl * x[i - 1];
			if (k < Nz) s += zr * x[i + 1];
			if (j > 0) s += rl * x[i - Nz - 1];
			if (j < Nr - 1) s += ru * x[i + Nz + 1];
			y[i] = s;
		}
}

static void factorize(ptpo_trap* t)
{
	int Nz = t->Nz, Nr = t->Nr, bw = Nr, w = 2 * bw + 1;
	long n = G_OF(t);
	t->band = (double*)calloc((size_t)n * w, sizeof(double));
#define BAND(p, q) t->band[(size_t)(p) * w + ((q) - (p) + bw)]
	for (int k = 0; k <= Nz; ++k)
		for (int j = 0; j < Nr; ++j) {
			double d, zl, zr, rl, ru;
			long p = (long)k * Nr + j;
			stencil(t, j, k, &d, &zl, &zr, &rl, &ru);
			BAND(p, p) = d;
			if (k > 0) BAND(p, p - Nr) = zl;
			if (k < Nz) BAND(p, p + Nr) = zr;
			if (j > 0) BAND(p, p - 1) = rl;
			if (j < Nr - 1) BAND(p, p + 1) = ru;
		}
	for (long c = 0; c < n; ++c) {
		double pivot = BAND(c, c);
		long iEnd = c + bw < n - 1 ? c + bw : n - 1;
		for (long i = c + 1; i <= iEnd; ++i) {
			double m = BAND(i, c);
			if (m == 0.0) continue;
			m /= pivot;
			BAND(i, c) = m;
			double* rowI = &BAND(i, c + 1);
			const double* rowC = &BAND(c, c + 1);
			long len = iEnd - c;
			for (long q = 0; q < len; ++q) rowI[q] -= m * rowC[q];
		}
	}
}

/* solver.solve(b) (Source/PenningTrap.cpp:202, Source/Plasma.cpp:98): x = A^-1 b, indices r-major. */
void ptpo_solve(const ptpo_trap* t, const double* rhs, double* phi)
{
	int Nz = t->Nz, Nr = t->Nr, bw = Nr, w = 2 * bw + 1;
	long n = G_OF(t);
	double* y = (double*)malloc((size_t)n * sizeof(double));
	for (int k = 0; k <= Nz; ++k)
		for (int j = 0; j < Nr; ++j) y[(long)k * Nr + j] = rhs[(Nz + 1) * j + k];
	for (long i = 0; i < n; ++i) {
		long q0 = i - bw > 0 ? i - bw : 0;
		double s = y[i];
		for (long q = q0; q < i; ++q) s -= BAND(i, q) * y[q];
		y[i] = s;
	}
	for (long i = n - 1; i >= 0; --i) {
		long q1 = i + bw < n - 1 ? i + bw : n - 1;
		double s = y[i];
		for (long q = i + 1; q <= q1; ++q) s -= BAND(i, q) * y[q];
		y[i] = s / BAND(i, i);
	}
	for (int k = 0; k <= Nz; ++k)
		for (int j = 0; j < Nr; ++j) phi[(Nz + 1) * j + k] = y[(long)k * Nr + j];
	free(y);
#undef BAND
}

/* PenningTrap::PenningTrap (Source/PenningTrap.cpp:36-58) up to and including the factorisation. */
ptpo_trap* ptpo_trap_create(double radius, int nElectrodes, const double* lengths, const double* potentials,
	int nGaps, const double* gaps, int Nz, int Nr)
{
	if (nElectrodes != nGaps + 1) return NULL;         /* :39-42 */
	ptpo_trap* t = (ptpo_trap*)calloc(1, sizeof(ptpo_trap));
	t->trapRadius = radius; t->nElectrodes = nElectrodes; t->nGaps = nGaps; t->Nz = Nz; t->Nr = Nr;
	t->electrodeLength = (double*)malloc(sizeof(double) * nElectrodes);
	t->electrodePotential = (double*)malloc(sizeof(double) * nElectrodes);
	t->gaps = (double*)malloc(sizeof(double) * (nGaps > 0 ? nGaps : 1));
	memcpy(t->electrodeLength, lengths, sizeof(double) * nElectrodes);
	memcpy(t->electrodePotential, potentials, sizeof(double) * nElectrodes);
	memcpy(t->gaps, gaps, sizeof(double) * nGaps);
	t->lengthTrap = 0;                                 /* zero-initialised storage, see ref_harness.cpp */
	for (int i = 0; i < nElectrodes; ++i) {            /* :43-50 */
		t->lengthTrap += lengths[i];
		if (i < nGaps) t->lengthTrap += gaps[i];
	}
	t->hz = t->lengthTrap / Nz;                        /* :51 */
	t->hr = t->trapRadius / Nr;                        /* :52 */
	factorize(t);                                      /* :55-57 */
	return t;
}

void ptpo_trap_destroy(ptpo_trap* t)
{
	if (!t) return;
	free(t->band); free(t->electrodeLength); free(t->electrodePotential); free(t->gaps); free(t);
}

void ptpo_trap_info(const ptpo_trap* t, int* Nz, int* Nr, double* hz, double* hr, double* length, double* radius)
{
	*Nz = t->Nz; *Nr = t->Nr; *hz = t->hz; *hr = t->hr; *length = t->lengthTrap; *radius = t->trapRadius;
}

/* Electrode::setPotential via PenningTrap::setPotential (Source/PenningTrap.cpp:313-317), without the solve. */
void ptpo_set_electrode(ptpo_trap* t, int index, double potential) { t->electrodePotential[index] = potential; }

/* Wall potential per axial node: the `boundary` values of PenningTrap::updateRHS (Source/PenningTrap.cpp:169-197). */
void ptpo_wall_potential(const ptpo_trap* t, double* vWall)
{
	int point = 0;
	double totalLength = 0;
	double boundary;
	for (int i = 0; i < t->nElectrodes; ++i) {
		boundary = t->electrodePotential[i];                                            /* :174 */
		while (point * t->hz <= t->electrodeLength[i] + totalLength) {                  /* :175 */
			if (point <= t->Nz) vWall[point] = boundary;   /* the reference would write past the row here */
			point++;
		}
		if (i < t->nGaps) {
			while (point * t->hz < t->electrodeLength[i] + t->gaps[i] + totalLength) {  /* :182 */
				boundary = (point * t->hz - t->electrodeLength[i] - totalLength)
					* (t->electrodePotential[i + 1] - t->electrodePotential[i]) / t->gaps[i]
					+ t->electrodePotential[i];                                         /* :184 */
				if (point <= t->Nz) vWall[point] = boundary;
				point++;
			}
			totalLength += t->electrodeLength[i] + t->gaps[i];                          /* :188 */
		}
	}
	while (point < t->Nz + 1) {                                                         /* :193-197 */
		vWall[point] = t->electrodePotential[t->nElectrodes - 1];
		point++;
	}
}

/* PenningTrap::updateRHS (Source/PenningTrap.cpp:163-198): dense copy of the sparse RHS. */
void ptpo_wall_rhs(const ptpo_trap* t, double* rhs)
{
	double hr2 = pow(t->hr, -2);                                    /* :165 */
	double radius = t->trapRadius - t->hr;                          /* :166 */
	double matrixFactor = hr2 + pow(2 * radius * t->hr, -1);        /* :167 */
	int N = t->Nz * t->Nr + t->Nr - t->Nz - 1;                      /* :168 */
	double* vWall = (double*)malloc(sizeof(double) * (t->Nz + 1));
	ptpo_wall_potential(t, vWall);
	memset(rhs, 0, sizeof(double) * G_OF(t));
	for (int point = 0; point <= t->Nz; ++point) rhs[point + N] = -1 * matrixFactor * vWall[point]; /* :177,185,195 */
	free(vWall);
}

/* Well limits, tail of the constructor (Source/PenningTrap.cpp:63-90). */
void ptpo_well_limits(const ptpo_trap* t, const double* phi, int* left, int* right)
{
	int Nz = t->Nz, Nr = t->Nr;
	int indexZ = (int)floor(t->lengthTrap / (2 * t->hz));           /* :63 */
	for (int indexR = 0; indexR < Nr; ++indexR) {                   /* :66-77 */
		int limit = indexZ + 1;
		double change = phi[(Nz + 1) * indexR + limit + 1] - phi[(Nz + 1) * indexR + limit];
		double changeTwo;
		do {
			++limit;
			changeTwo = phi[(Nz + 1) * indexR + limit + 1] - phi[(Nz + 1) * indexR + limit];
		} while (change * changeTwo > 0 && limit + 1 < Nz);
		right[indexR] = limit;
	}
	for (int indexR = 0; indexR < Nr; ++indexR) {                   /* :79-90 */
		int limit = indexZ;
		double change = phi[(Nz + 1) * indexR + limit - 1] - phi[(Nz + 1) * indexR + limit];
		double changeTwo;
		do {
			--limit;
			changeTwo = phi[(Nz + 1) * indexR + limit - 1] - phi[(Nz + 1) * indexR + limit];
		} while (change * changeTwo > 0 && limit - 1 > 0);
		left[indexR] = limit;
	}
}

/* PenningTrap::getEField(int,int) (Source/PenningTrap.cpp:208-236) for every node. phis[0] is the trap
 * potential, phis[1..] the plasmas' self potentials in registration order (the summation order of :226-232). */
void ptpo_node_efield(const ptpo_trap* t, int nPhi, const double* const* phis, double* eNodes)
{
	int Nz = t->Nz, Nr = t->Nr;
	for (int indexR = 0; indexR < Nr; ++indexR)
		for (int indexZ = 0; indexZ <= Nz; ++indexZ) {
			int index = (Nz + 1) * indexR + indexZ;                 /* :210 */
			if (indexZ == 0 || indexZ == Nz) { eNodes[index] = 0; continue; } /* :218-222 */
			double valueLeft = phis[0][index - 1];                  /* :226 */
			double valueRight = phis[0][index + 1];                 /* :227 */
			for (int s = 1; s < nPhi; ++s) {                        /* :228-232 */
				valueLeft += phis[s][index - 1];
				valueRight += phis[s][index + 1];
			}
			eNodes[index] = (valueLeft - valueRight) / (2 * t->hz); /* :233 */
		}
}

/* PenningTrap::getEField(int,double) (Source/PenningTrap.cpp:326-334). */
double ptpo_gather(const ptpo_trap* t, const double* eNodes, int r, double z)
{
	int indexZ = (int)floor(z / t->hz);                             /* :328 */
	double fieldLeft = eNodes[(t->Nz + 1) * r + indexZ];            /* :329 */
	double fieldRight = eNodes[(t->Nz + 1) * r + indexZ + 1];       /* :330 */
	double dz = z - indexZ * t->hz;                                 /* :331 */
	double weightFactor = dz / t->hz;                               /* :332 */
	return ((1 - weightFactor) * fieldLeft + weightFactor * fieldRight); /* :333 */
}

/* Plasma::moveRings (Source/Plasma.cpp:100-120) on SoA copies of the rings; returns the new ring count.
 * Removal = swap with the last ring and pop, index not advanced (:114-118). */
long ptpo_move_rings(const ptpo_trap* t, const double* eNodes, long n, int* r, double* z, double* v,
	double dt, double charge, double mass)
{
	for (long i = 0; i < n; ) {
		double vNew = dt * ptpo_gather(t, eNodes, r[i], z[i]) * charge / mass + v[i]; /* :105 */
		double zNew = dt * vNew + z[i];                                               /* :106 */
		if (zNew < t->lengthTrap && zNew > 0) {                                        /* :108 */
			z[i] = zNew;                                                               /* :110 */
			v[i] = vNew;                                                               /* :111 */
			++i;
		}
		else {
			int tr = r[i]; double tz = z[i], tv = v[i];                                /* :116 std::swap */
			r[i] = r[n - 1]; z[i] = z[n - 1]; v[i] = v[n - 1];
			r[n - 1] = tr; z[n - 1] = tz; v[n - 1] = tv;
			--n;                                                                       /* :117 pop_back */
		}
	}
	return n;
}

/* Plasma::updateRHS (Source/Plasma.cpp:77-94): rhs = -rho/epsilon0 by first-order weighting in z. */
void ptpo_deposit(const ptpo_trap* t, long n, const int* r, const double* z, double macroChargeDensity, double* rhs)
{
	memset(rhs, 0, sizeof(double) * G_OF(t));                       /* :81 */
	double hz = t->hz;
	for (long i = 0; i < n; ++i) {
		int indexR = r[i];                                          /* :86 */
		int indexZ = (int)floor(z[i] / hz);                         /* :87 */
		int indexRHS = (t->Nz + 1) * indexR + indexZ;               /* :88 */
		double zz = z[i] - indexZ * hz;                             /* :89 */
		double weightFactor = zz / hz;                              /* :90 */
		rhs[indexRHS] += -macroChargeDensity * (1 - weightFactor) / epsilon;     /* :91 */
		rhs[indexRHS + 1] += -macroChargeDensity * weightFactor / epsilon;       /* :92 */
	}
}

/* The integer keys of the step: axial cell k = (int)floor(z/hz) and flat index (Nz+1)*r + k
 * (Source/Plasma.cpp:87-88, Source/PenningTrap.cpp:328). */
void ptpo_cell_index(const ptpo_trap* t, long n, const int* r, const double* z, int* k, int* idx)
{
	for (long i = 0; i < n; ++i) {
		k[i] = (int)floor(z[i] / t->hz);
		idx[i] = (t->Nz + 1) * r[i] + k[i];
	}
}

/* Plasma::getPotentialEnergy (Source/Plasma.cpp:244-252) with getTotalPhi(int,double)
 * (Source/PenningTrap.cpp:335-351). */
double ptpo_potential_energy(const ptpo_trap* t, int nPhi, const double* const* phis, long n, const int* r,
	const double* z, double chargeMacro)
{
	double potentialEnergy = 0;
	for (long i = 0; i < n; ++i) {
		int indexZ = (int)floor(z[i] / t->hz);                      /* :347 */
		double dz = z[i] - indexZ * t->hz;                          /* :348 */
		double weightFactor = dz / t->hz;                           /* :349 */
		int index = (t->Nz + 1) * r[i] + indexZ;                    /* :337 */
		double phiL = phis[0][index], phiR = phis[0][index + 1];    /* :338 */
		for (int s = 1; s < nPhi; ++s) { phiL += phis[s][index]; phiR += phis[s][index + 1]; } /* :339-342 */
		double phi = ((1 - weightFactor) * phiL + weightFactor * phiR);                       /* :350 */
		potentialEnergy += phi * (r[i] == 0 ? chargeMacro : r[i] * 8 * chargeMacro);          /* :249 */
	}
	return potentialEnergy / 2;                                     /* :251 */
}
