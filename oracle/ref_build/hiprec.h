// Force-included (-include) when compiling the reference's MakeDensity.cpp for the oracle:
// widens the 8-significant-digit moment columns (MakeDensity.cpp:2437-2500) to 17 digits so that
// golden vectors carry full double precision. The reference source itself is not modified.
#include <iomanip>
#include <iostream>
#include <fstream>
#include <sstream>
#define setprecision(n) setprecision(((n) == 8) ? 17 : (n))
#define setw(n) setw(((n) == 16) ? 26 : (n))
