// Physical constants under the names the reference's user programs expect
// (reference Source/Constants.hpp:11-16; CODATA 2018 values).
#ifndef CONSTANTS_HPP
#define CONSTANTS_HPP

const double ePos{ 1.602176634e-19 };      // elementary charge [C]
const double epsilon{ 8.8541878128e-12 };  // vacuum permittivity [F/m]
const double massE{ 9.1093837015e-31 };    // electron mass [kg]
const double massP{ 1.67262192369e-27 };   // proton (antiproton) mass [kg]
const double PI{ 3.141592653589793238463 };
const double KB{ 1.380649e-23 };           // Boltzmann constant [J/K]

#endif
