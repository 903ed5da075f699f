// Host side of Plasma (public surface of reference Source/Plasma.hpp:139-197) over the C ABI of libptp_b200.
// The step itself (moveRings / updateRHS / solvePoisson, Source/Plasma.cpp:77-120) runs on the GPU. The
// loaders are the reference's one-off initial-condition builders (Source/Plasma.cpp:366-623): their
// equilibrium iteration calls the GPU solver through ptp_trap_solve, the ring placement and the Maxwellian
// speeds are evaluated here and uploaded. Saved histories and the diagnostics that read them live here too.
#include "Plasma.hpp"

#include "ptp.h"

#include <algorithm>
#include <cstdlib>
#include <numeric>

namespace {

void check(int rc)
{
	if (rc == PTP_OK) return;
	if (rc == PTP_EINVAL) throw std::logic_error(ptp_last_error());
	throw std::runtime_error(ptp_last_error());
}

void writeColumn(const std::string& fileName, const std::vector<double>& values)
{
	std::ofstream out(fileName);
	out << std::setprecision(std::numeric_limits<double>::digits10);
	for (std::size_t i = 0; i < values.size(); ++i) {
		if (i) out << "\n";
		out << values[i];
	}
}

// Volume of the grid cell that belongs to radial index j (a disc of radius hr/2 on the axis, a ring elsewhere).
double cellVolume(int j, double hz, double hr)
{
	if (j == 0) return PI * hz * hr * hr / 4;
	return hz * hr * 2 * PI * j * hr;
}

} // namespace

Plasma::Plasma(PenningTrap& trap, std::string aName, double aMass, double aCharge)
	: refTrap(trap), name(aName), mass(aMass), charge(aCharge), chargeMacro(0), macroChargeDensity(0), massMacro(0),
	  temperature(0), device(nullptr), initialDensity((std::size_t)(trap.Nz + 1) * trap.Nr, 0.0)
{
	check(ptp_plasma_create(trap.device, &device, mass, charge));
	refTrap.addPlasma(*this);
}

Plasma::~Plasma()
{
	// The reference never deregisters a plasma from its trap (Source/Plasma.cpp:73-76): the trap must outlive
	// it. If the trap is already gone it has taken the device twin with it (device == nullptr).
	if (device) {
		ptp_plasma_destroy(device);
		auto& list = refTrap.plasmas;
		for (std::size_t i = 0; i < list.size(); ++i)
			if (&list[i].get() == this) { list.erase(list.begin() + i); break; }
	}
}

std::vector<double> Plasma::selfPotential() const
{
	std::vector<double> phi(initialDensity.size());
	check(ptp_plasma_get_self_potential(device, phi.data()));
	return phi;
}

void Plasma::solvePoisson() { check(ptp_plasma_deposit_solve(device)); }

void Plasma::extractSelfPotential(std::string fileName) const { writeColumn(fileName, selfPotential()); }

void Plasma::extractPlasmaParameters(std::string fileName) const
{
	std::ofstream out(fileName);
	out << std::setprecision(std::numeric_limits<double>::digits10);
	out << mass << '\n' << charge << '\n' << chargeMacro;
}

void Plasma::extractInitialDensity(std::string fileName) const { writeColumn(fileName, initialDensity); }

int Plasma::getNumMacro() const
{
	int64_t n = 0;
	check(ptp_plasma_count(device, &n));
	return (int)n;
}

int Plasma::getNumMacroCentralWell() const
{
	int64_t n = 0;
	check(ptp_plasma_count_central_well(device, refTrap.limitLeft.data(), refTrap.limitRight.data(), &n));
	return (int)n;
}

// ---- ring bookkeeping --------------------------------------------------------------------------------

void Plasma::getRings(std::vector<int>& r, std::vector<double>& z, std::vector<double>& v, std::vector<std::int64_t>& id) const
{
	const std::size_t n = (std::size_t)getNumMacro();
	r.resize(n); z.resize(n); v.resize(n); id.resize(n);
	static_assert(sizeof(int) == sizeof(int32_t), "int must be 32 bits");
	check(ptp_plasma_download(device, reinterpret_cast<int32_t*>(r.data()), z.data(), v.data(), reinterpret_cast<int64_t*>(id.data())));
}

// Bring `order` (ids of the live rings in the reference's ring order) up to date. The reference removes a ring at once,
// in the step in which it leaves the trap, by swapping it with the last one and shrinking the vector without advancing the
// index (Source/Plasma.cpp:114-118). The push kernel logs (ring id, step) for every lost ring; replaying the reference's
// sweep once per logged step reproduces its ring order - hence the row order of the history files - also when rings were
// lost in several steps since the last refresh (multi-step movePlasmas calls, electrode programmes). One sweep over the
// rings lost in one step only moves rings from the back into the holes, so it is replayed from the position table in
// O(lost) instead of O(live).
void Plasma::refreshAlive()
{
	if (!device) return;
	const std::size_t alive = (std::size_t)getNumMacro();
	if (alive == order.size()) return;
	std::int64_t total = 0;
	int overflowed = 0;
	check(ptp_plasma_loss_log(device, lossSeen, 0, nullptr, nullptr, &total, &overflowed));
	const std::int64_t expected = (std::int64_t)order.size() - (std::int64_t)alive;
	if (!overflowed && total - lossSeen == expected && posOf.size() == ringAlive.size()) {
		std::vector<std::int64_t> ids((std::size_t)expected), steps((std::size_t)expected);
		check(ptp_plasma_loss_log(device, lossSeen, expected, ids.data(), steps.data(), &total, &overflowed));
		lossSeen += expected;
		// entries arrive in step order (kernels of consecutive steps are stream-ordered); group by step
		std::vector<std::size_t> idx((std::size_t)expected);
		std::iota(idx.begin(), idx.end(), (std::size_t)0);
		std::stable_sort(idx.begin(), idx.end(), [&](std::size_t a, std::size_t b) { return steps[a] < steps[b]; });
		std::size_t g0 = 0;
		std::vector<std::int64_t> holes;
		while (g0 < idx.size()) {
			std::size_t g1 = g0;
			holes.clear();
			while (g1 < idx.size() && steps[idx[g1]] == steps[idx[g0]]) {
				const std::int64_t id = ids[idx[g1]];
				ringAlive[(std::size_t)id] = 0;
				holes.push_back(posOf[(std::size_t)id]);
				++g1;
			}
			std::sort(holes.begin(), holes.end());
			std::size_t n = order.size();
			for (std::int64_t h : holes) {                        // the reference's sweep: ascending index, re-examine after a swap
				const std::size_t i = (std::size_t)h;
				while (i < n && !ringAlive[(std::size_t)order[i]]) {
					posOf[(std::size_t)order[i]] = -1;
					order[i] = order[n - 1];
					if (i != n - 1) posOf[(std::size_t)order[i]] = (std::int64_t)i;
					--n;
				}
			}
			order.resize(n);
			g0 = g1;
		}
		return;
	}
	// log not usable (more rings lost than it holds): one sweep over everything lost since the last refresh - the reference's
	// order only if all of them left in the same step
	lossSeen = total;
	std::vector<int> r;
	std::vector<double> z, v;
	std::vector<std::int64_t> id;
	getRings(r, z, v, id);
	std::fill(ringAlive.begin(), ringAlive.end(), 0);
	for (std::int64_t i : id) ringAlive[(std::size_t)i] = 1;
	for (std::size_t i = 0; i < order.size();) {
		if (ringAlive[(std::size_t)order[i]]) ++i;
		else {
			std::swap(order[i], order.back());
			order.pop_back();
		}
	}
	posOf.assign(ringAlive.size(), -1);
	for (std::size_t i = 0; i < order.size(); ++i) posOf[(std::size_t)order[i]] = (std::int64_t)i;
}

// PenningTrap::saveStates -> Plasma::saveState (Source/Plasma.cpp:330-346). Every save point also feeds the device-side
// temperature sums (mean of the speeds at two consecutive save points, Source/Plasma.cpp:180,224). With histories on the host
// (the default, what extractPlasmasHistories needs) the rings are downloaded - all of them, or for saveState(indexR) just
// the contiguous slice of that radial row.
void Plasma::saveSelected(int indexR, bool all)
{
	if (all) {
		double sw = 0, s2 = 0;
		int paired = 0;
		check(ptp_plasma_kinetic_sums(device, 1, &sw, &s2, &paired));
		if (paired && sw > 0) pairTemperature.push_back(mass * s2 / (KB * sw));
	}
	if (!hostHistories) return;
	refreshAlive();
	if (all) {
		std::vector<int> r;
		std::vector<double> z, v;
		std::vector<std::int64_t> id;
		getRings(r, z, v, id);
		for (std::size_t i = 0; i < id.size(); ++i) {
			historyZ[(std::size_t)id[i]].push_back(z[i]);
			historySpeed[(std::size_t)id[i]].push_back(v[i]);
		}
		return;
	}
	if (indexR < 0 || indexR >= refTrap.Nr) return;
	std::size_t cap = 0;
	for (std::int64_t i : order) cap += ringR[(std::size_t)i] == indexR ? 1 : 0;
	std::vector<double> z(cap), v(cap);
	std::vector<std::int64_t> id(cap);
	std::int64_t n = 0;
	check(ptp_plasma_download_row(device, indexR, (std::int64_t)cap, z.data(), v.data(), reinterpret_cast<int64_t*>(id.data()), &n));
	for (std::int64_t i = 0; i < n; ++i) {
		historyZ[(std::size_t)id[(std::size_t)i]].push_back(z[(std::size_t)i]);
		historySpeed[(std::size_t)id[(std::size_t)i]].push_back(v[(std::size_t)i]);
	}
}

void Plasma::saveState() { saveSelected(0, true); }
void Plasma::saveState(int indexR) { saveSelected(indexR, false); }

void Plasma::reserve(int desired)
{
	for (std::int64_t i : order) {
		historyZ[(std::size_t)i].reserve(desired);
		historySpeed[(std::size_t)i].reserve(desired);
	}
}

void Plasma::extractHistory(std::string preName) const
{
	std::ofstream positions(preName + "Positions" + name + ".csv"), speeds(preName + "Speeds" + name + ".csv");
	positions << std::setprecision(std::numeric_limits<double>::digits10);
	speeds << std::setprecision(std::numeric_limits<double>::digits10);
	for (std::int64_t i : order) {
		const std::vector<double>& hz = historyZ[(std::size_t)i];
		const std::vector<double>& hv = historySpeed[(std::size_t)i];
		if (!hz.empty()) {
			positions << ringR[(std::size_t)i] << ",";
			for (std::size_t t = 0; t + 1 < hz.size(); ++t) positions << hz[t] << ",";
			positions << hz.back() << "\n";
		}
		for (std::size_t t = 0; t < hv.size(); ++t) {
			speeds << hv[t];
			if (t + 1 < hv.size()) speeds << ","; else speeds << "\n";
		}
	}
}

// ---- diagnostics on saved histories (host loops, as in the reference) -------------------------------------

double Plasma::getPotentialEnergy() const
{
	double pe = 0;
	check(ptp_plasma_potential_energy(device, chargeMacro, &pe));
	return pe;
}

double Plasma::getTemperature() const
{
	if (!hostHistories) {                                       // from the device sums of the last two save points
		if (pairTemperature.empty()) throw std::logic_error("getTemperature needs two saved states");
		return pairTemperature.back();
	}
	double numParticles = 0, KE = 0;
	for (std::int64_t i : order) {
		const int r = ringR[(std::size_t)i];
		const double aMass = r == 0 ? massMacro : 8 * r * massMacro;
		numParticles += aMass / mass;
	}
	for (std::int64_t i : order) {
		const int r = ringR[(std::size_t)i];
		const double aMass = r == 0 ? massMacro : 8 * r * massMacro;
		const std::vector<double>& hv = historySpeed[(std::size_t)i];
		const double speed = (hv.end()[-2] + hv.back()) / 2; // positions and speeds are staggered by dt/2
		KE += 0.5 * aMass * speed * speed;
	}
	return 2 * KE / (KB * numParticles);
}

double Plasma::getAverageTemperature() const
{
	if (!hostHistories) {
		if (pairTemperature.empty()) throw std::logic_error("getAverageTemperature needs two saved states");
		double T = 0;
		for (double x : pairTemperature) T += x;
		return T / (double)pairTemperature.size();
	}
	const int times = (int)historySpeed[(std::size_t)order[0]].size() - 1;
	double numParticles = 0;
	for (std::int64_t i : order) {
		const int r = ringR[(std::size_t)i];
		numParticles += (r == 0 ? massMacro : 8 * r * massMacro) / mass;
	}
	double T = 0;
	for (int t = 0; t < times; ++t) {
		double KE = 0;
		for (std::int64_t i : order) {
			const int r = ringR[(std::size_t)i];
			const double aMass = r == 0 ? massMacro : 8 * r * massMacro;
			const std::vector<double>& hv = historySpeed[(std::size_t)i];
			const double speed = (hv[t] + hv[t + 1]) / 2;
			KE += 0.5 * aMass * speed * speed;
		}
		T += 2 * KE / (KB * numParticles);
	}
	return T / times;
}

double Plasma::getstdDeviation() const
{
	if (!hostHistories) {
		const double m = getAverageTemperature();
		double acc = 0;
		for (double x : pairTemperature) acc += pow(x - m, 2);
		return sqrt(acc / ((double)pairTemperature.size() - 1));
	}
	const double mean = getAverageTemperature();
	const int times = (int)historySpeed[(std::size_t)order[0]].size() - 1;
	double numParticles = 0;
	for (std::int64_t i : order) {
		const int r = ringR[(std::size_t)i];
		numParticles += (r == 0 ? massMacro : 8 * r * massMacro) / mass;
	}
	double acc = 0;
	for (int t = 0; t < times; ++t) {
		double KE = 0;
		for (std::int64_t i : order) {
			const int r = ringR[(std::size_t)i];
			const double aMass = r == 0 ? massMacro : 8 * r * massMacro;
			const std::vector<double>& hv = historySpeed[(std::size_t)i];
			const double speed = (hv[t] + hv[t + 1]) / 2;
			KE += 0.5 * aMass * speed * speed;
		}
		acc += pow((2 * KE / (KB * numParticles)) - mean, 2);
	}
	return sqrt(acc / (times - 1));
}

double Plasma::getCentralDensity() const
{
	// The reference tests `pointsZ - 1 % 2 == 0` (Source/Plasma.cpp:357), which parses as pointsZ - (1 % 2) == 0
	// and is never true for a real grid, so it always averages the two nodes around pointsZ / 2. Kept as is.
	const int pointsZ = refTrap.Nz + 1;
	return (initialDensity[(std::size_t)(pointsZ / 2)] + initialDensity[(std::size_t)(pointsZ / 2) - 1]) / (2 * charge);
}

// ---- loaders ---------------------------------------------------------------------------------------------

void Plasma::loadRings(const std::vector<int>& r, const std::vector<double>& z, const std::vector<double>& v, double aChargeMacro, double aTemperature)
{
	if (r.size() != z.size() || r.size() != v.size()) throw std::logic_error("loadRings: r, z and v must have the same length");
	temperature = aTemperature;
	chargeMacro = aChargeMacro;
	massMacro = chargeMacro * mass / charge;
	macroChargeDensity = 4 * chargeMacro / (PI * refTrap.hz * refTrap.hr * refTrap.hr);
	check(ptp_plasma_upload(device, (int64_t)r.size(), reinterpret_cast<const int32_t*>(r.data()), z.data(), v.data(), macroChargeDensity));
	ringR = r;
	ringAlive.assign(r.size(), 1);
	historyZ.assign(r.size(), std::vector<double>());
	historySpeed.assign(r.size(), std::vector<double>());
	order.resize(r.size());
	std::iota(order.begin(), order.end(), (std::int64_t)0);
	posOf = order;
	lossSeen = 0;
	pairTemperature.clear();
	solvePoisson();
}

// Turn the expected density grid into rings: per radial row the cumulative charge along z, a ring count
// proportional to the row's charge (a ring at index i carries 8 i chargeMacro), equally spaced charge
// quantiles inverted by linear interpolation, Maxwellian speeds from the standard library's default engine
// (freshly seeded for every load, so every species sees the same deviate sequence scaled by its sigma).
void Plasma::placeRings(int numMacro)
{
	// Large loads are placed on the device (ptp_plasma_load_density: same counts, bit-identical positions, the same deviate
	// stream to the last bits of log()); small ones stay here so that their speeds are bit-identical too.
	// PTP_DEVICE_LOADER=1 / 0 forces one or the other.
	const char* force = std::getenv("PTP_DEVICE_LOADER");
	if (force ? force[0] == '1' : numMacro >= 2000000) {
		std::vector<std::int64_t> perRowDevice((std::size_t)refTrap.Nr);
		std::int64_t loaded = 0;
		double newChargeMacro = 0;
		check(ptp_plasma_load_density(device, initialDensity.data(), temperature, numMacro, 0, 1, &newChargeMacro, perRowDevice.data(), &loaded));
		chargeMacro = newChargeMacro;
		massMacro = chargeMacro * mass / charge;
		macroChargeDensity = 4 * chargeMacro / (PI * refTrap.hz * refTrap.hr * refTrap.hr);
		ringR.clear();
		ringR.reserve((std::size_t)loaded);
		for (int j = 0; j < refTrap.Nr; ++j) ringR.insert(ringR.end(), (std::size_t)perRowDevice[(std::size_t)j], j);
		ringAlive.assign(ringR.size(), 1);
		historyZ.assign(ringR.size(), std::vector<double>());
		historySpeed.assign(ringR.size(), std::vector<double>());
		order.resize(ringR.size());
		std::iota(order.begin(), order.end(), (std::int64_t)0);
		posOf = order;
		lossSeen = 0;
		pairTemperature.clear();
		std::cout << "Loading " << ringR.size() << " macro-particles from which " << perRowDevice[0] << " are at r=0.\n";
		solvePoisson();
		return;
	}
	const int Nz = refTrap.Nz, Nr = refTrap.Nr, n1 = Nz + 1;
	const double hz = refTrap.hz, hr = refTrap.hr;
	std::vector<std::vector<double>> cumulative((std::size_t)Nr, std::vector<double>((std::size_t)n1));
	for (int j = 0; j < Nr; ++j) {
		const double volume = cellVolume(j, hz, hr);
		double running = 0;
		for (int k = 0; k < n1; ++k) {
			running += volume * initialDensity[(std::size_t)n1 * j + k];
			cumulative[j][k] = running;
		}
	}
	double axisEquivalentCharge = cumulative[0].back();
	for (int j = 1; j < Nr; ++j) axisEquivalentCharge += cumulative[j].back() / (8 * j);
	const double newChargeMacro = axisEquivalentCharge / numMacro;
	std::vector<int> perRow((std::size_t)Nr);
	perRow[0] = (int)round(cumulative[0].back() / newChargeMacro);
	for (int j = 1; j < Nr; ++j) perRow[j] = (int)round(cumulative[j].back() / (8 * j * newChargeMacro));

	std::vector<int> r;
	std::vector<double> z, v;
	r.reserve(numMacro); z.reserve(numMacro); v.reserve(numMacro);
	std::default_random_engine generator;
	std::normal_distribution<double> maxwellian(0, sqrt(KB * temperature / mass));
	for (int j = 0; j < Nr; ++j) {
		const std::vector<double>& cum = cumulative[j];
		const double quantum = cum.back() / (perRow[j] + 1);
		int node = 0;
		for (int i = 0; i < perRow[j]; ++i) {
			const double target = quantum * (i + 1);
			while (std::abs(cum[node]) < std::abs(target)) ++node;
			const double position = (node - 1) * hz + hz / 2 + hz * (target - cum[node - 1]) / (cum[node] - cum[node - 1]);
			r.push_back(j);
			z.push_back(position);
			v.push_back(maxwellian(generator));
		}
	}
	std::cout << "Loading " << r.size() << " macro-particles from which " << perRow[0] << " are at r=0.\n";
	loadRings(r, z, v, newChargeMacro, temperature);
}

void Plasma::loadDensityFile(std::string fileName, double aTemperature, int numMacro)
{
	if (aTemperature <= 0) throw std::logic_error("Temperature needs to be positive");
	if (numMacro <= 0) throw std::logic_error("Number of macro-particles has to be a positive integer");
	temperature = aTemperature;
	std::fill(initialDensity.begin(), initialDensity.end(), 0.0);
	std::ifstream in(fileName);
	std::size_t count = 0;
	double value;
	while (in >> value) {
		if (count < initialDensity.size()) initialDensity[count] = value;
		++count;
	}
	if (count != initialDensity.size()) throw std::logic_error("The number of grid points in the file do not match this trap.");
	placeRings(numMacro);
}

// Local-thermal-equilibrium estimate along every radial row: n(z) proportional to exp(-q (phi - phi_centre) / kT)
// inside the central well, zero outside.
void Plasma::estimateDensityProportions(const std::vector<double>& totalPhi)
{
	const int n1 = refTrap.Nz + 1;
	for (int j = 0; j < refTrap.Nr; ++j) {
		const double phiCentre = refTrap.getTotalPhi(totalPhi, j, refTrap.lengthTrap / 2);
		for (int k = 0; k < n1; ++k) {
			double& n = initialDensity[(std::size_t)n1 * j + k];
			if (k < refTrap.limitLeft[j] || k > refTrap.limitRight[j]) n = 0;
			else n = exp(-(charge / (KB * temperature)) * (refTrap.getTotalPhi(totalPhi, j, k) - phiCentre));
		}
	}
}

// Scale every row so that its z-integrated density follows exp(-(r / scale)^shape), r in millimetres.
void Plasma::fitDensityProportionToProfile(double shape, double scale)
{
	const int n1 = refTrap.Nz + 1;
	for (int j = 0; j < refTrap.Nr; ++j) {
		double lineDensity = 0;
		for (int k = 0; k < n1; ++k) lineDensity += initialDensity[(std::size_t)n1 * j + k] * refTrap.hz;
		const double factor = exp(-pow((j * refTrap.hr * 1000 / scale), shape)) / lineDensity;
		for (int k = 0; k < n1; ++k) initialDensity[(std::size_t)n1 * j + k] *= factor;
	}
}

void Plasma::normalizeDensityToTotalCharge(double totalCharge)
{
	const int n1 = refTrap.Nz + 1;
	double currentCharge = 0;
	for (int j = 0; j < refTrap.Nr; ++j) {
		const double volume = cellVolume(j, refTrap.hz, refTrap.hr);
		for (int k = 0; k < n1; ++k) currentCharge += initialDensity[(std::size_t)n1 * j + k] * volume;
	}
	const double correction = totalCharge / currentCharge;
	for (double& n : initialDensity) n *= correction;
}

// Damped fixed-point iteration for the thermal-equilibrium density with a prescribed radial profile:
// density -> Poisson solve (GPU) -> Boltzmann factor along z -> refit radial profile -> renormalise charge,
// mixed with the previous iterate with a weight set by the first Kolmogorov-Smirnov-like distance, until that
// distance between successive axial distributions drops below the threshold.
void Plasma::loadProfile(double aTemperature, double aTotalCharge, double shape, double scale, int numMacro, double KSThreshold)
{
	if (aTemperature <= 0 || shape <= 0 || scale <= 0) throw std::logic_error("Temperature, shape, and scale all need to be positive");
	if (numMacro <= 0) throw std::logic_error("Number of macro-particles has to be a positive integer");
	if (aTotalCharge * charge < 0) throw std::logic_error("Total charge and the plasma type charge must have the same sign");
	if (KSThreshold <= 0 || KSThreshold >= 1) throw std::logic_error("The Kolmogorov Smirnov distance threshold needs to be a number between 0 and 1");
	temperature = aTemperature;
	const int Nz = refTrap.Nz, Nr = refTrap.Nr, n1 = Nz + 1;
	const std::size_t G = initialDensity.size();

	// potential of everything except this plasma (trap + species loaded earlier)
	std::vector<double> background = refTrap.trapPotential();
	{
		std::vector<double> self(G);
		for (const Plasma& other : refTrap.plasmas) {
			if (&other == this || !other.device) continue;
			check(ptp_plasma_get_self_potential(other.device, self.data()));
			for (std::size_t i = 0; i < G; ++i) background[i] += self[i];
		}
	}
	auto refit = [&](const std::vector<double>& ownPhi) {
		std::vector<double> total(G);
		for (std::size_t i = 0; i < G; ++i) total[i] = background[i] + ownPhi[i];
		estimateDensityProportions(total);
		fitDensityProportionToProfile(shape, scale);
		normalizeDensityToTotalCharge(aTotalCharge);
	};
	std::fill(initialDensity.begin(), initialDensity.end(), 0.0);
	refit(std::vector<double>(G, 0.0)); // first guess: no space charge

	std::vector<double> volume((std::size_t)Nr);
	for (int j = 0; j < Nr; ++j) volume[j] = cellVolume(j, refTrap.hz, refTrap.hr);
	double damping = 0;
	bool first = true;
	std::vector<double> rhs(G);
	for (;;) {
		const std::vector<double> previous = initialDensity;
		for (std::size_t i = 0; i < G; ++i) rhs[i] = -initialDensity[i] / epsilon;
		refit(refTrap.solve(rhs));
		// largest gap between the cumulative axial charge distributions of the two iterates
		double newSum = 0, oldSum = 0, maxDistance = 0;
		for (int k = 0; k < n1; ++k) {
			for (int j = 0; j < Nr; ++j) {
				newSum += volume[j] * initialDensity[(std::size_t)n1 * j + k];
				oldSum += volume[j] * previous[(std::size_t)n1 * j + k];
			}
			const double distance = std::abs(newSum / aTotalCharge - oldSum / aTotalCharge);
			if (distance > maxDistance) maxDistance = distance;
		}
		if (first) {
			first = false;
			damping = 2 * maxDistance; // symmetric trap: only half of the distribution matters
		}
		std::cout << maxDistance << '\n';
		if (maxDistance < KSThreshold) break;
		for (std::size_t i = 0; i < G; ++i) initialDensity[i] = damping * previous[i] + (1 - damping) * initialDensity[i];
	}
	placeRings(numMacro);
}
