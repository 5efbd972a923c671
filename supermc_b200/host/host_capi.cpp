// C hooks over the host layer so the parameter parser and the text writers can be checked from the test
// suite without a GPU (tests/test_host_layer.py).
#include <cstring>
#include <string>
#include "MakeDensity.h"

extern "C" {
// parse `text` (a parameters.dat body) then `overrides` (space-separated name=value); returns the value of `name`
int smc_host_param(const char* text, const char* overrides, const char* name, double* value) {
  try {
    ParameterReader r;
    std::string t(text), line;
    size_t a = 0;
    while (a <= t.size()) { size_t b = t.find('\n', a); if (b == std::string::npos) b = t.size(); r.phraseOneLine(t.substr(a, b - a)); a = b + 1; }
    std::string o(overrides ? overrides : "");
    a = 0;
    while (a < o.size()) { size_t b = o.find(' ', a); if (b == std::string::npos) b = o.size(); if (b > a) r.phraseOneLine(o.substr(a, b - a)); a = b + 1; }
    *value = r.getVal(name);
    return 0;
  } catch (std::exception&) { return 1; }
}
static int copy_out(const std::string& s, char* out, int cap) { if ((int)s.size() + 1 > cap) return -(int)s.size() - 1; std::memcpy(out, s.c_str(), s.size() + 1); return (int)s.size(); }
int smc_host_format_ecc_row(const smc_event_out* ev, int order, int deformed, char* out, int cap) {
  return copy_out(order == 10 ? MakeDensity::formatEccRowAll(*ev, deformed != 0) : MakeDensity::formatEccRow(*ev, order, deformed != 0), out, cap);
}
int smc_host_format_list(int kind, const double* rows, int n, int stride, char* out, int cap) {
  return copy_out(kind == 0 ? smc_fmt_xy(rows, n, stride) : kind == 1 ? smc_fmt_participants(rows, n) : smc_fmt_spectators(rows, n), out, cap);
}
int smc_host_format_block(const double* g, int Maxx, int Maxy, char* out, int cap) { std::string s; MakeDensity::formatDensityBlock(g, Maxx, Maxy, s); return copy_out(s, out, cap); }
int smc_host_format_4col(const double* g, int Maxx, int Maxy, double Xmin, double Ymin, double dx, double dy, double rap, double npart, char* out, int cap) {
  std::string s; MakeDensity::formatDensity4Col(g, Maxx, Maxy, Xmin, Ymin, dx, dy, rap, npart, s); return copy_out(s, out, cap);
}
}
