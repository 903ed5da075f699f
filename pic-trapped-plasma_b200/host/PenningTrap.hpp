// Host-side Electrode / PenningTrap with the public surface of the reference
// (reference Source/PenningTrap.hpp:38-98), so that existing driver programs compile unchanged.
// All grid work (operator, Laplace/Poisson solves, node field, the PIC step) runs on the GPU through
// the C ABI in include/ptp.h; this class keeps only what the reference keeps on the host side of the
// hot path: the electrode list, the electrode -> wall-node mapping, the well limits, saved histories
// and the text writers whose formats the Diagnostics scripts read.
#ifndef PENNINGTRAP_HPP
#define PENNINGTRAP_HPP

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

struct ptp_trap;
struct ptp_plasma;
class Plasma;

class Electrode
{
private:
	double length;
	double potential;

public:
	Electrode(double aLength, double aPotential);
	Electrode(const Electrode& copiable);
	~Electrode();
	double getLength() const;
	double getPotential() const;
	void setPotential(double finalPotential);
};

class PenningTrap
{
private:
	double trapRadius;
	std::vector<Electrode> electrodes;
	std::vector<double> gaps;
	int Nz; // cells along z (Nz + 1 nodes)
	int Nr; // nodes along r; the wall r = R is not stored
	double hr, hz;
	double lengthTrap;
	std::vector<double> potentialEnergiesHistory;
	std::vector<double> timesSaved;
	std::vector<int> limitLeft;
	std::vector<int> limitRight;
	std::vector<std::reference_wrapper<Plasma>> plasmas;
	ptp_trap* device; // GPU twin (operator, phi_trap, node field, per-species grids)
	bool electrodeBasis = false;

	void addPlasma(Plasma&);
	void solveLaplace();                       // wall potential -> device Laplace solve
	void findWellLimits();                     // central-well index range per radial row
	std::vector<double> wallPotential() const; // electrode / gap potential at every axial node
	std::vector<double> trapPotential() const; // phi_trap copied from the device
	std::vector<double> solve(const std::vector<double>& rhs) const; // A^-1 rhs on the device
	double getTotalPhi(const std::vector<double>& total, int r, int z) const;
	double getTotalPhi(const std::vector<double>& total, int r, double z) const;
	std::vector<double> totalPotential() const; // phi_trap + sum of the plasmas' self potentials
	PenningTrap(const PenningTrap&) = delete;
	PenningTrap& operator=(const PenningTrap&) = delete;

public:
	PenningTrap(double radius, const std::vector<Electrode>& theElectrodes, const std::vector<double>& theGaps, int NumCellsZ, int NumCellsR);
	~PenningTrap();
	friend class Plasma;
	void extractTrapPotential(std::string fileName) const;
	void extractTrapLaplacian(std::string fileName) const; // A*phi - RHS, expected to be all zeros
	void extractPlasmasHistories(std::string pathAndPreName) const;
	void extractTrapParameters(std::string filename) const;
	void setPotential(int indexElectrode, double newPotential);
	double getLength() const;
	double getRadius() const;
	void movePlasmas(double deltaT);
	void saveStates(double aTime);
	void saveStates(double aTime, int indexR);
	void reserve(int desired);

	// Extensions (not in the reference): several steps per call, device selection, raw handle.
	void movePlasmas(double deltaT, int numSteps);
	// keepHistories(false): saveStates keeps only what the temperature diagnostics need (sums formed on the device at every
	// save point) and the potential energy - no ring crosses PCIe, getTemperature / getAverageTemperature / getstdDeviation
	// work as before, extractPlasmasHistories has nothing to write. For loads whose histories would not fit on the host.
	void keepHistories(bool on);
	// Electrode programmes: keep one Laplace solution per electrode on the device; setPotential then costs one axpy
	// kernel instead of a solve (also switched on by PTP_ELECTRODE_BASIS=1), and a schedule potentials[step][electrode]
	// runs without returning to the host between steps.
	void useElectrodeBasis();
	void movePlasmas(double deltaT, const std::vector<std::vector<double>>& potentials);
	ptp_trap* deviceHandle() const { return device; }
	static void selectDevice(int cudaOrdinal); // device used by traps constructed afterwards (default 0 / $PTP_DEVICE)
};
#endif
