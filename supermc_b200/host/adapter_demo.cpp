// adapter_demo -- MakeDensity-style event loop on top of MCnuclB200 (host/MCnuclB200.h): the calls of the reference's
// generateEccTable (src/MakeDensity.cpp:2143-2200) and the grid loops of its dumpEccentricities (:2273-2298: total, centre of
// mass, <r^2>) written against the adapter, printed next to the columns the engine computed on the device.
//   usage: adapter_demo nev [name=value ...]      (reads parameters.dat from the working directory)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include "MCnuclB200.h"
#include "ParameterReader.h"

typedef MCnuclB200T<ParameterReader> MCnuclB200;

int main(int argc, char* argv[]) {
  if (argc < 2) { std::cerr << "usage: adapter_demo nev [name=value ...]" << std::endl; return 2; }
  const int nev = std::atoi(argv[1]);
  ParameterReader rdr;
  try {
    rdr.readFromFile("parameters.dat"); rdr.readFromArguments(argc, argv, "#", 2);
    MCnuclB200 mc_obj(&rdr, SMC_RUN_KEEP_RHO | SMC_RUN_THICKNESS, 64);
    MCnuclB200* mc = &mc_obj;
    const int Maxx = mc->getMaxx(), Maxy = mc->getMaxy();
    const double dx = rdr.getVal("dx"), dy = rdr.getVal("dy"), Xmin = -rdr.getVal("maxx"), Ymin = -rdr.getVal("maxy"), ff = rdr.getVal("finalFactor");
    for (int event = 1; event <= nev; event++) {
      int binary = 0;
      while (binary == 0 || mc->CentralityCut() == 0) {                  // MakeDensity.cpp:2147-2162
        mc->generateNuclei(0.0);
        binary = mc->getBinaryCollision();
        if (binary == 0 || mc->CentralityCut() == 0) mc->deleteNucleus();
      }
      mc->calculateThickness(); mc->setDensity(0, -1);
      // the reference's own grid loops, reading the lattice through getRho (MakeDensity.cpp:2273-2298)
      double total = 0, xc = 0, yc = 0, ta = 0;
      for (int i = 0; i < Maxx; i++) for (int j = 0; j < Maxy; j++) {
        const double d = mc->getRho(0, i, j) * ff;
        total += d; xc += (Xmin + i * dx) * d; yc += (Ymin + j * dy) * d; ta += mc->getTA1(i, j);
      }
      xc /= total; yc /= total;
      double r2 = 0;
      for (int i = 0; i < Maxx; i++) for (int j = 0; j < Maxy; j++) {
        const double x = Xmin + i * dx - xc, y = Ymin + j * dy - yc;
        r2 += (x * x + y * y) * mc->getRho(0, i, j) * ff;
      }
      std::printf("%d %d %d %d %.15e %.15e %.15e %.15e %.15e %.15e\n", event, mc->getNpart1(), mc->getNpart2(), mc->getNcoll(),
                  total * dx * dy, mc->totalEntropy(), r2 / total, mc->moments(2)[4], ta * dx * dy, mc->lastB());
    }
  } catch (std::exception& e) { std::cerr << "adapter_demo: " << e.what() << std::endl; return 255; }
  return 0;
}
