// Host side of Electrode / PenningTrap (public surface of reference Source/PenningTrap.hpp:38-98) over the
// C ABI of libptp_b200 (include/ptp.h). Everything numerical on the grid happens on the GPU; what stays here
// is the electrode bookkeeping, the electrode -> wall-node mapping (it decides, through floating-point
// comparisons, which node belongs to which electrode, so it is evaluated on the host exactly as the
// reference does, Source/PenningTrap.cpp:169-197), the well limits and the text writers.
#include "PenningTrap.hpp"

#include <cstdlib>
#include "Plasma.hpp"

#include "ptp.h"

namespace {

int g_selectedDevice = -1;

int chosenDevice()
{
	if (g_selectedDevice >= 0) return g_selectedDevice;
	const char* env = std::getenv("PTP_DEVICE");
	return env ? std::atoi(env) : 0;
}

// Bad arguments surface as std::logic_error like the reference's own checks; everything else (no GPU, CUDA
// or NCCL failure) as std::runtime_error - there is no CPU fallback to fall back to.
void check(int rc)
{
	if (rc == PTP_OK) return;
	if (rc == PTP_EINVAL) throw std::logic_error(ptp_last_error());
	throw std::runtime_error(ptp_last_error());
}

// One value per line, 15 significant digits, no newline after the last value: the layout of the reference's
// grid dumps (Eigen FullPrecision + "\n" row separator, Source/PenningTrap.cpp:237-244).
void writeColumn(const std::string& fileName, const std::vector<double>& values)
{
	std::ofstream out(fileName);
	out << std::setprecision(std::numeric_limits<double>::digits10);
	for (std::size_t i = 0; i < values.size(); ++i) {
		if (i) out << "\n";
		out << values[i];
	}
}

void writeCommaLine(std::ofstream& out, const std::vector<double>& values)
{
	for (std::size_t i = 0; i < values.size(); ++i) {
		out << values[i];
		if (i + 1 < values.size()) out << ",";
	}
}

} // namespace

Electrode::Electrode(double aLength, double aPotential) : length(aLength), potential(aPotential) {}
Electrode::Electrode(const Electrode& copiable) : length(copiable.length), potential(copiable.potential) {}
Electrode::~Electrode() {}
double Electrode::getLength() const { return length; }
double Electrode::getPotential() const { return potential; }
void Electrode::setPotential(double finalPotential) { potential = finalPotential; }

void PenningTrap::selectDevice(int cudaOrdinal) { g_selectedDevice = cudaOrdinal; }

PenningTrap::PenningTrap(double radius, const std::vector<Electrode>& theElectrodes, const std::vector<double>& theGaps, int NumCellsZ, int NumCellsR)
	: trapRadius(radius), electrodes(theElectrodes), gaps(theGaps), Nz(NumCellsZ), Nr(NumCellsR), hr(0), hz(0), lengthTrap(0), device(nullptr)
{
	if (electrodes.size() != gaps.size() + 1)
		throw std::logic_error("Error number of gaps and electrodes; No. electrods should match No. gaps + 1");
	for (std::size_t i = 0; i < electrodes.size(); ++i) {
		lengthTrap += electrodes[i].getLength();
		if (i < gaps.size()) lengthTrap += gaps[i];
	}
	hz = lengthTrap / Nz;
	hr = trapRadius / Nr;
	check(ptp_trap_create(&device, Nz, Nr, hz, hr, lengthTrap, trapRadius, chosenDevice()));
	solveLaplace();
	findWellLimits();
}

PenningTrap::~PenningTrap()
{
	for (Plasma& p : plasmas) p.device = nullptr; // the device twins die with the trap
	ptp_trap_destroy(device);
}

void PenningTrap::addPlasma(Plasma& aPlasma) { plasmas.push_back(aPlasma); }

// Wall potential at every axial node: an electrode's potential up to and including its end, a linear ramp
// strictly inside a gap, and the last electrode for any node the comparisons left over.
std::vector<double> PenningTrap::wallPotential() const
{
	std::vector<double> wall(Nz + 1, 0.0);
	int node = 0;
	double start = 0; // axial position where the current electrode begins
	for (std::size_t i = 0; i < electrodes.size(); ++i) {
		const double V = electrodes[i].getPotential();
		const double electrodeEnd = electrodes[i].getLength() + start;
		while (node * hz <= electrodeEnd) {
			if (node <= Nz) wall[node] = V;
			++node;
		}
		if (i < gaps.size()) {
			const double gapEnd = electrodes[i].getLength() + gaps[i] + start;
			while (node * hz < gapEnd) {
				const double ramp = (node * hz - electrodes[i].getLength() - start) * (electrodes[i + 1].getPotential() - V) / gaps[i] + V;
				if (node <= Nz) wall[node] = ramp;
				++node;
			}
			start += electrodes[i].getLength() + gaps[i];
		}
	}
	for (; node <= Nz; ++node) wall[node] = electrodes.back().getPotential();
	return wall;
}

void PenningTrap::solveLaplace()
{
	const std::vector<double> wall = wallPotential();
	check(ptp_trap_set_wall(device, wall.data()));
}

std::vector<double> PenningTrap::trapPotential() const
{
	std::vector<double> phi((std::size_t)(Nz + 1) * Nr);
	check(ptp_trap_get_phi(device, phi.data()));
	return phi;
}

std::vector<double> PenningTrap::solve(const std::vector<double>& rhs) const
{
	std::vector<double> phi(rhs.size());
	check(ptp_trap_solve(device, rhs.data(), phi.data()));
	return phi;
}

// Index range of the central well of the vacuum potential on every radial row: starting next to the trap
// centre, walk outwards while the potential keeps changing in the same direction.
void PenningTrap::findWellLimits()
{
	const std::vector<double> phi = trapPotential();
	const int n1 = Nz + 1;
	const int centre = (int)floor(lengthTrap / (2 * hz));
	limitLeft.assign(Nr, 0);
	limitRight.assign(Nr, 0);
	for (int j = 0; j < Nr; ++j) {
		const double* row = phi.data() + (std::size_t)n1 * j;
		int k = centre + 1;
		const double firstStep = row[k + 1] - row[k];
		double step;
		do {
			++k;
			step = row[k + 1] - row[k];
		} while (firstStep * step > 0 && k + 1 < Nz);
		limitRight[j] = k;
		k = centre;
		const double firstStepLeft = row[k - 1] - row[k];
		do {
			--k;
			step = row[k - 1] - row[k];
		} while (firstStepLeft * step > 0 && k - 1 > 0);
		limitLeft[j] = k;
	}
}

std::vector<double> PenningTrap::totalPotential() const
{
	std::vector<double> total = trapPotential();
	std::vector<double> self(total.size());
	for (const Plasma& p : plasmas) {
		if (!p.device) continue;
		check(ptp_plasma_get_self_potential(p.device, self.data()));
		for (std::size_t i = 0; i < total.size(); ++i) total[i] += self[i];
	}
	return total;
}

double PenningTrap::getTotalPhi(const std::vector<double>& total, int r, int z) const
{
	return total[(std::size_t)(Nz + 1) * r + z];
}

double PenningTrap::getTotalPhi(const std::vector<double>& total, int r, double z) const
{
	const int k = (int)floor(z / hz);
	const double w = (z - k * hz) / hz;
	return (1 - w) * getTotalPhi(total, r, k) + w * getTotalPhi(total, r, k + 1);
}

void PenningTrap::extractTrapPotential(std::string fileName) const { writeColumn(fileName, trapPotential()); }

void PenningTrap::extractTrapLaplacian(std::string fileName) const
{
	const std::vector<double> phi = trapPotential();
	std::vector<double> residual(phi.size());
	check(ptp_trap_apply(device, phi.data(), residual.data()));
	// subtract the wall term that the Dirichlet row carries on the right-hand side
	const std::vector<double> wall = wallPotential();
	const double factor = std::pow(hr, -2) + std::pow(2 * (trapRadius - hr) * hr, -1);
	const std::size_t lastRow = (std::size_t)Nz * Nr + Nr - Nz - 1;
	for (int k = 0; k <= Nz; ++k) residual[lastRow + k] -= -1 * factor * wall[k];
	writeColumn(fileName, residual);
}

void PenningTrap::extractPlasmasHistories(std::string pathAndPreName) const
{
	std::ofstream out(pathAndPreName + "Times.csv");
	out << std::setprecision(std::numeric_limits<double>::digits10);
	writeCommaLine(out, timesSaved);
	out.close();
	out.open(pathAndPreName + "PotentialEnergies.csv");
	out << std::setprecision(std::numeric_limits<double>::digits10);
	writeCommaLine(out, potentialEnergiesHistory);
	out.close();
	for (const Plasma& p : plasmas) p.extractHistory(pathAndPreName);
}

void PenningTrap::extractTrapParameters(std::string filename) const
{
	std::ofstream out(filename);
	out << std::setprecision(std::numeric_limits<double>::digits10);
	out << trapRadius << '\n';
	std::vector<double> lengths, potentials;
	for (const Electrode& e : electrodes) {
		lengths.push_back(e.getLength());
		potentials.push_back(e.getPotential());
	}
	// every list is followed by a newline only when it is non-empty, as in the reference's writer
	writeCommaLine(out, lengths);
	if (!lengths.empty()) out << '\n';
	writeCommaLine(out, potentials);
	if (!potentials.empty()) out << '\n';
	writeCommaLine(out, gaps);
	if (!gaps.empty()) out << '\n';
	out << Nz << ',' << Nr << '\n';
	out << hz << ',' << hr << '\n';
	out << lengthTrap;
}

void PenningTrap::setPotential(int indexElectrode, double newPotential)
{
	electrodes[indexElectrode].setPotential(newPotential);
	if (!electrodeBasis) {
		const char* env = std::getenv("PTP_ELECTRODE_BASIS");
		if (env && env[0] == '1') useElectrodeBasis();
	}
	if (electrodeBasis) {
		std::vector<double> weights;
		for (const Electrode& e : electrodes) weights.push_back(e.getPotential());
		check(ptp_trap_set_wall_weights(device, weights.data()));
	}
	else solveLaplace();
}

void PenningTrap::useElectrodeBasis()
{
	// wall profile of "electrode i at 1 V, the others grounded" through the same node loop as every other wall
	std::vector<double> saved, walls;
	for (const Electrode& e : electrodes) saved.push_back(e.getPotential());
	for (std::size_t i = 0; i < electrodes.size(); ++i) {
		for (std::size_t j = 0; j < electrodes.size(); ++j) electrodes[j].setPotential(i == j ? 1.0 : 0.0);
		const std::vector<double> wall = wallPotential();
		walls.insert(walls.end(), wall.begin(), wall.end());
	}
	for (std::size_t j = 0; j < electrodes.size(); ++j) electrodes[j].setPotential(saved[j]);
	check(ptp_trap_set_wall_basis(device, (int)electrodes.size(), walls.data()));
	check(ptp_trap_set_wall_weights(device, saved.data()));
	electrodeBasis = true;
}

void PenningTrap::movePlasmas(double deltaT, const std::vector<std::vector<double>>& potentials)
{
	if (potentials.empty()) return;
	if (!electrodeBasis) useElectrodeBasis();
	std::vector<double> flat;
	for (const std::vector<double>& step : potentials) {
		if (step.size() != electrodes.size()) throw std::logic_error("movePlasmas: one potential per electrode and step");
		flat.insert(flat.end(), step.begin(), step.end());
	}
	check(ptp_trap_step_programme(device, deltaT, (int)potentials.size(), flat.data()));
	for (std::size_t j = 0; j < electrodes.size(); ++j) electrodes[j].setPotential(potentials.back()[j]);
	for (Plasma& p : plasmas) p.refreshAlive();
}

double PenningTrap::getLength() const { return lengthTrap; }
double PenningTrap::getRadius() const { return trapRadius; }

void PenningTrap::movePlasmas(double deltaT)
{
	check(ptp_trap_step(device, deltaT, 1));
	for (Plasma& p : plasmas) p.refreshAlive(); // keeps the ring order of saved histories in step with removals
}

void PenningTrap::movePlasmas(double deltaT, int numSteps)
{
	check(ptp_trap_step(device, deltaT, numSteps));
	for (Plasma& p : plasmas) p.refreshAlive();
}

void PenningTrap::keepHistories(bool on)
{
	for (Plasma& p : plasmas) p.hostHistories = on;
}

void PenningTrap::saveStates(double aTime)
{
	timesSaved.push_back(aTime);
	double potentialEnergy = 0;
	for (Plasma& p : plasmas) {
		p.saveState();
		potentialEnergy += p.getPotentialEnergy();
	}
	potentialEnergiesHistory.push_back(potentialEnergy);
}

void PenningTrap::saveStates(double aTime, int indexR)
{
	timesSaved.push_back(aTime);
	double potentialEnergy = 0;
	for (Plasma& p : plasmas) {
		p.saveState(indexR);
		potentialEnergy += p.getPotentialEnergy();
	}
	potentialEnergiesHistory.push_back(potentialEnergy);
}

void PenningTrap::reserve(int desired)
{
	timesSaved.reserve(desired);
	for (Plasma& p : plasmas) p.reserve(desired);
	potentialEnergiesHistory.reserve(desired);
}
