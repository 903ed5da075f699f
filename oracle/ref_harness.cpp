// TEST INFRASTRUCTURE ONLY -- C handle API around the UNMODIFIED reference sources.
//
// Built by oracle/Makefile into oracle/_ref/libptp_ref.so together with
// /root/reference/Source/PenningTrap.cpp and Plasma.cpp (compiled where they
// lie; never copied) against oracle/eigen_standin/Eigen/SparseLU.
// -fno-access-control lets this file reach the reference's private step
// methods (Plasma::moveRings, Plasma::updateRHS, PenningTrap::getEField ...),
// so every phase of PenningTrap::movePlasmas (Source/PenningTrap.cpp:352-363)
// can be driven and observed separately by tests/ and by bench.py's CPU legs.
// The product library never links or loads this.
#include "PenningTrap.hpp"
#include "Plasma.hpp"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

namespace {
thread_local std::string g_error;
struct TrapBox { PenningTrap* trap; void* storage; };
struct PlasmaBox { Plasma* plasma; void* storage; };
inline int G(const PenningTrap& t) { return t.Nz * t.Nr + t.Nr; }
double nowSeconds()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

// PenningTrap::PenningTrap (Source/PenningTrap.cpp:36-91). The object is placed in
// zeroed storage because lengthTrap is accumulated with += without ever being
// initialised (Source/PenningTrap.hpp:45, Source/PenningTrap.cpp:45).
void* ref_trap_create(double radius, int nElectrodes, const double* lengths, const double* potentials,
	int nGaps, const double* gaps, int Nz, int Nr)
{
	try {
		std::vector<Electrode> electrodes;
		for (int i = 0; i < nElectrodes; ++i) electrodes.push_back(Electrode(lengths[i], potentials[i]));
		std::vector<double> theGaps(gaps, gaps + nGaps);
		void* storage = std::calloc(1, sizeof(PenningTrap));
		PenningTrap* trap = new (storage) PenningTrap(radius, electrodes, theGaps, Nz, Nr);
		return new TrapBox{ trap, storage };
	}
	catch (const std::exception& e) { g_error = e.what(); return nullptr; }
}
void ref_trap_destroy(void* h)
{
	TrapBox* box = (TrapBox*)h;
	box->trap->~PenningTrap();
	std::free(box->storage);
	delete box;
}
#define TRAP(h) (*((TrapBox*)(h))->trap)
#define PLASMA(h) (*((PlasmaBox*)(h))->plasma)

void ref_trap_info(void* h, int* Nz, int* Nr, double* hz, double* hr, double* length, double* radius)
{
	PenningTrap& t = TRAP(h);
	*Nz = t.Nz; *Nr = t.Nr; *hz = t.hz; *hr = t.hr; *length = t.lengthTrap; *radius = t.trapRadius;
}
void ref_trap_get_phi(void* h, double* out)
{
	PenningTrap& t = TRAP(h);
	for (int i = 0; i < G(t); ++i) out[i] = t.potentialsVector.coeff(i);
}
void ref_trap_get_wall_rhs(void* h, double* out)
{
	PenningTrap& t = TRAP(h);
	for (int i = 0; i < G(t); ++i) out[i] = t.RHS.coeff(i);
}
void ref_trap_limits(void* h, int* left, int* right)
{
	PenningTrap& t = TRAP(h);
	for (int i = 0; i < t.Nr; ++i) { left[i] = t.limitLeft[i]; right[i] = t.limitRight[i]; }
}
void ref_trap_set_potential(void* h, int indexElectrode, double v) { TRAP(h).setPotential(indexElectrode, v); }
void ref_trap_solve(void* h, const double* rhs, double* out)
{
	PenningTrap& t = TRAP(h);
	Eigen::VectorXd b(G(t));
	for (int i = 0; i < G(t); ++i) b.coeffRef(i) = rhs[i];
	Eigen::VectorXd x = t.solver.solve(b);
	for (int i = 0; i < G(t); ++i) out[i] = x.coeff(i);
}
void ref_trap_apply(void* h, const double* x, double* y)
{
	PenningTrap& t = TRAP(h);
	Eigen::VectorXd v(G(t));
	for (int i = 0; i < G(t); ++i) v.coeffRef(i) = x[i];
	Eigen::VectorXd r = t.coefficients * v;
	for (int i = 0; i < G(t); ++i) y[i] = r.coeff(i);
}
long ref_trap_matrix_nnz(void* h) { return (long)TRAP(h).coefficients.nonZeros().size(); }
void ref_trap_matrix(void* h, int* rows, int* cols, double* vals)
{
	long i = 0;
	for (const auto& e : TRAP(h).coefficients.nonZeros()) { rows[i] = e.row; cols[i] = e.col; vals[i] = e.value; ++i; }
}
// PenningTrap::getEField(int,int) (Source/PenningTrap.cpp:208-236) on every node; cache dropped afterwards.
void ref_trap_get_enodes(void* h, double* out)
{
	PenningTrap& t = TRAP(h);
	for (int r = 0; r < t.Nr; ++r)
		for (int k = 0; k <= t.Nz; ++k) out[(t.Nz + 1) * r + k] = t.getEField(r, k);
	t.eFields.clear();
}
double ref_trap_efield(void* h, int r, double z)
{
	PenningTrap& t = TRAP(h);
	double e = t.getEField(r, z);
	t.eFields.clear();
	return e;
}
double ref_trap_total_phi(void* h, int r, double z) { return TRAP(h).getTotalPhi(r, z); }
void ref_trap_move_plasmas(void* h, double dt, int nSteps)
{
	for (int i = 0; i < nSteps; ++i) TRAP(h).movePlasmas(dt);
}
void ref_trap_save_states(void* h, double t) { TRAP(h).saveStates(t); }
void ref_trap_save_states_r(void* h, double t, int r) { TRAP(h).saveStates(t, r); }
double ref_trap_last_potential_energy(void* h) { return TRAP(h).potentialEnergiesHistory.back(); }
void ref_trap_extract_trap_potential(void* h, const char* f) { TRAP(h).extractTrapPotential(f); }
void ref_trap_extract_trap_laplacian(void* h, const char* f) { TRAP(h).extractTrapLaplacian(f); }
void ref_trap_extract_trap_parameters(void* h, const char* f) { TRAP(h).extractTrapParameters(f); }
void ref_trap_extract_histories(void* h, const char* prefix) { TRAP(h).extractPlasmasHistories(prefix); }

// Plasma::Plasma (Source/Plasma.cpp:68-72). Never destroyed before its trap (the
// reference never deregisters, Source/Plasma.cpp:73-76).
void* ref_plasma_create(void* trap, const char* name, double mass, double charge)
{
	void* storage = std::calloc(1, sizeof(Plasma));
	Plasma* p = new (storage) Plasma(TRAP(trap), name, mass, charge);
	return new PlasmaBox{ p, storage };
}
void ref_plasma_destroy(void* h)
{
	PlasmaBox* box = (PlasmaBox*)h;
	box->plasma->~Plasma();
	std::free(box->storage);
	delete box;
}
int ref_plasma_load_profile(void* h, double T, double totalCharge, double shape, double scale, int numMacro, double ks)
{
	try { PLASMA(h).loadProfile(T, totalCharge, shape, scale, numMacro, ks); return 0; }
	catch (const std::exception& e) { g_error = e.what(); return 1; }
}
int ref_plasma_load_density_file(void* h, const char* file, double T, int numMacro)
{
	try { PLASMA(h).loadDensityFile(file, T, numMacro); return 0; }
	catch (const std::exception& e) { g_error = e.what(); return 1; }
}
// Inject an explicit set of rings; macro quantities derived as both loaders do
// (Source/Plasma.cpp:492-494 / 586-588). No solve.
void ref_plasma_set_rings(void* h, long n, const int* r, const double* z, const double* v, double chargeMacro)
{
	Plasma& p = PLASMA(h);
	p.chargeMacro = chargeMacro;
	p.massMacro = chargeMacro * p.mass / p.charge;
	p.macroChargeDensity = 4 * chargeMacro / (PI * p.refTrap.hz * p.refTrap.hr * p.refTrap.hr);
	p.rings.clear();
	p.rings.reserve((size_t)n);
	for (long i = 0; i < n; ++i) p.rings.push_back(MacroRing(r[i], z[i], v[i]));
}
long ref_plasma_count(void* h) { return (long)PLASMA(h).rings.size(); }
int ref_plasma_num_central_well(void* h) { return PLASMA(h).getNumMacroCentralWell(); }
void ref_plasma_get_rings(void* h, int* r, double* z, double* v)
{
	Plasma& p = PLASMA(h);
	for (size_t i = 0; i < p.rings.size(); ++i) { r[i] = p.rings[i].posR; z[i] = p.rings[i].posZ; v[i] = p.rings[i].speed; }
}
void ref_plasma_params(void* h, double* chargeMacro, double* macroChargeDensity, double* massMacro, double* temperature)
{
	Plasma& p = PLASMA(h);
	*chargeMacro = p.chargeMacro; *macroChargeDensity = p.macroChargeDensity; *massMacro = p.massMacro; *temperature = p.temperature;
}
void ref_plasma_get_rhs(void* h, double* out)
{
	Plasma& p = PLASMA(h);
	for (int i = 0; i < G(p.refTrap); ++i) out[i] = p.RHS.coeff(i);
}
void ref_plasma_get_self_potential(void* h, double* out)
{
	Plasma& p = PLASMA(h);
	for (int i = 0; i < G(p.refTrap); ++i) out[i] = p.selfPotential.coeff(i);
}
void ref_plasma_set_self_potential(void* h, const double* in)
{
	Plasma& p = PLASMA(h);
	for (int i = 0; i < G(p.refTrap); ++i) p.selfPotential.coeffRef(i) = in[i];
}
void ref_plasma_get_initial_density(void* h, double* out)
{
	Plasma& p = PLASMA(h);
	for (int i = 0; i < G(p.refTrap); ++i) out[i] = p.initialDensity.coeff(i);
}
void ref_plasma_update_rhs(void* h) { PLASMA(h).updateRHS(); }
void ref_plasma_solve_poisson(void* h) { PLASMA(h).solvePoisson(); }
void ref_plasma_move_rings(void* h, double dt) { PLASMA(h).moveRings(dt); PLASMA(h).refTrap.eFields.clear(); }
double ref_plasma_temperature(void* h) { return PLASMA(h).getTemperature(); }
double ref_plasma_potential_energy(void* h) { return PLASMA(h).getPotentialEnergy(); }
void ref_plasma_extract_self_potential(void* h, const char* f) { PLASMA(h).extractSelfPotential(f); }
void ref_plasma_extract_parameters(void* h, const char* f) { PLASMA(h).extractPlasmaParameters(f); }
void ref_plasma_extract_initial_density(void* h, const char* f) { PLASMA(h).extractInitialDensity(f); }

// CPU baseline: nSteps of PenningTrap::movePlasmas with the three phases timed
// apart (same order as Source/PenningTrap.cpp:352-363). seconds[0]=moveRings,
// [1]=updateRHS, [2]=solver.solve, [3]=whole. Returns ring-steps processed
// (rings alive at the start of each step, summed).
double ref_trap_timed_steps(void* h, double dt, int nSteps, double* seconds)
{
	PenningTrap& t = TRAP(h);
	double ringSteps = 0;
	seconds[0] = seconds[1] = seconds[2] = seconds[3] = 0;
	for (int s = 0; s < nSteps; ++s) {
		double t0 = nowSeconds();
		for (Plasma& p : t.plasmas) { ringSteps += (double)p.rings.size(); p.moveRings(dt); }
		double t1 = nowSeconds();
		double tRhs = 0, tSolve = 0;
		for (Plasma& p : t.plasmas) {
			double a = nowSeconds();
			p.updateRHS();
			double b = nowSeconds();
			p.selfPotential = t.solver.solve(p.RHS);
			double c = nowSeconds();
			tRhs += b - a; tSolve += c - b;
		}
		t.eFields.clear();
		double t2 = nowSeconds();
		seconds[0] += t1 - t0; seconds[1] += tRhs; seconds[2] += tSolve; seconds[3] += t2 - t0;
	}
	return ringSteps;
}

} // extern "C"
