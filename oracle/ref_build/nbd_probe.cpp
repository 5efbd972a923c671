// nbd_probe -- test infrastructure: draws from the UNMODIFIED reference's NBD class (src/NBD.cpp, RandomVariable.cpp,
// TableFunction.cpp, Table.cpp, arsenal.cpp) so that the oracle's restatement of NBD::rand can be pinned sample for sample.
//   nbd_probe <seed> <n> <p> <r> [<p> <r> ...]     prints, per (p, r): a header line "p r" and n samples
// drand48 is seeded once with srand48(seed); MCnucl::fluctuateCurrentDensity (src/MCnucl.cpp:868-905) calls exactly
// nbd->rand(p, r) per cell.
#include <cstdio>
#include <cstdlib>
#include "NBD.h"
int main(int argc, char** argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: nbd_probe seed n p r [p r ...]\n"); return 2; }
  const long seed = std::atol(argv[1]); const long n = std::atol(argv[2]);
  srand48(seed);
  NBD nbd;
  for (int a = 3; a + 1 < argc; a += 2) {
    const double p = std::atof(argv[a]), r = std::atof(argv[a + 1]);
    std::printf("%.17g %.17g\n", p, r);
    for (long i = 0; i < n; i++) std::printf("%ld\n", nbd.rand(p, r));
  }
  return 0;
}
