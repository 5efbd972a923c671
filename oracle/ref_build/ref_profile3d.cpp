// ref_profile3d: harness around the UNMODIFIED 3-D extension of the reference (compiled from
// /root/reference/scripts/generate_3d_profiles/{profile_3d,Regge96}.cpp where they lie).  TEST INFRASTRUCTURE ONLY.
// It replaces only main.cpp (same list readers, :25-46) so that the lattice, ecm and random_flag can be chosen and the
// rapidities / widths the object drew (private members; time-seeded mt19937) can be written next to the lattice.
// usage: ref_profile3d participants.dat binary.dat nx ny neta dx dy deta ecm random_flag out.bin
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <cmath>
#define private public
#include "profile_3d.h"
#undef private
using namespace std;

int main(int argc, char* argv[]) {
  if (argc != 12) { fprintf(stderr, "usage: ref_profile3d part.dat binary.dat nx ny neta dx dy deta ecm random_flag out.bin\n"); return 2; }
  vector<participant_info> part_list, binary_list;
  { ifstream f(argv[1]); while (!f.eof()) { participant_info t; f >> t.x >> t.y >> t.id; part_list.push_back(t); } part_list.pop_back(); }
  { ifstream f(argv[2]); while (!f.eof()) { participant_info t; f >> t.x >> t.y; binary_list.push_back(t); } binary_list.pop_back(); }
  const int nx = atoi(argv[3]), ny = atoi(argv[4]), neta = atoi(argv[5]);
  profile_3d p(part_list, binary_list, nx, ny, neta, atof(argv[6]), atof(argv[7]), atof(argv[8]), atof(argv[9]), atoi(argv[10]));
  p.generate_3d_profile();
  p.output_3d_rhob_profile("ref_rhob.dat");
  FILE* o = fopen(argv[11], "wb");
  long n = (long)p.participant_list.size(); fwrite(&n, 8, 1, o);
  for (long i = 0; i < n; i++) { double r[7] = {p.participant_list[i].x, p.participant_list[i].y, (double)p.participant_list[i].id, p.participant_list[i].eta,
                                                p.participant_list[i].sigma_x, p.participant_list[i].sigma_y, p.participant_list[i].sigma_eta}; fwrite(r, 8, 7, o); }
  for (int j = 0; j < neta; j++) for (int k = 0; k < nx; k++) fwrite(p.rho_part[j][k], 8, ny, o);
  fclose(o);
  return 0;
}
