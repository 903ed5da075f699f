// Host-side Plasma with the public surface of the reference (reference Source/Plasma.hpp:139-197).
// The rings live on the GPU (row-bucketed SoA behind ptp_plasma); the host keeps the species constants,
// the expected initial density, and - only when the user calls saveStates - the saved (z, v) histories
// keyed by ring id, which is what the reference stores inside each MacroRing (Source/Plasma.hpp:118-137).
#ifndef PLASMA_HPP
#define PLASMA_HPP

#include "Constants.hpp"
#include "PenningTrap.hpp"

#include <array>
#include <cstdint>
#include <iostream>
#include <random>
#include <string>
#include <vector>

class Plasma
{
private:
	PenningTrap& refTrap;
	const std::string name;
	const double mass;   // of one particle of the species [kg]
	const double charge; // of one particle of the species [C]
	double chargeMacro;  // ring charge at r = 0; a ring at radial index i carries 8 i chargeMacro
	double macroChargeDensity;
	double massMacro;
	double temperature;
	ptp_plasma* device;
	std::vector<double> initialDensity; // expected (not deposited) charge density the plasma was loaded from

	// Rings as last uploaded / downloaded, in the reference's ring order (index = ring id at upload).
	std::vector<int> ringR;
	std::vector<char> ringAlive;
	std::vector<std::vector<double>> historyZ, historySpeed; // per ring id, appended by saveState
	std::vector<std::int64_t> order;                          // ids of the live rings in output order
	std::vector<std::int64_t> posOf;                          // position of every ring id in `order` (-1: removed)
	std::int64_t lossSeen = 0;                                // entries of the device's loss log already replayed
	bool hostHistories = true;                                // false: save points keep only the temperature sums (PenningTrap::keepHistories)
	std::vector<double> pairTemperature;                      // temperature of every pair of consecutive save points, from device sums

	void placeRings(int numMacro);            // inverse-CDF placement + Maxwellian speeds, upload, first solve
	void solvePoisson();
	void saveState();
	void saveState(int indexR);
	void saveSelected(int indexR, bool all);
	void reserve(int desired);
	void extractHistory(std::string preName) const;
	double getPotentialEnergy() const;
	void refreshAlive();                      // pull the live set from the device
	std::vector<double> selfPotential() const;
	void estimateDensityProportions(const std::vector<double>& totalPhi);
	void fitDensityProportionToProfile(double shape, double scale);
	void normalizeDensityToTotalCharge(double totalCharge);
	Plasma(const Plasma&) = delete;
	Plasma& operator=(const Plasma&) = delete;

public:
	Plasma(PenningTrap& trap, std::string name, double mass, double charge);
	~Plasma();
	friend class PenningTrap;
	void extractSelfPotential(std::string fileName) const;
	void extractPlasmaParameters(std::string filename) const;
	void extractInitialDensity(std::string filename) const;
	int getNumMacro() const;
	int getNumMacroCentralWell() const;
	double getAverageTemperature() const;
	double getstdDeviation() const;
	double getTemperature() const;
	double getCentralDensity() const;
	void loadProfile(double aTemperature, double totalCharge, double shape, double scale, int numMacro, double KSThreshold);
	void loadDensityFile(std::string fileName, double aTemperature, int numMacro);

	// Extensions (not in the reference): explicit ring upload / download for large synthetic loads.
	void loadRings(const std::vector<int>& r, const std::vector<double>& z, const std::vector<double>& v, double aChargeMacro, double aTemperature);
	void getRings(std::vector<int>& r, std::vector<double>& z, std::vector<double>& v, std::vector<std::int64_t>& id) const;
	ptp_plasma* deviceHandle() const { return device; }
};

#endif
