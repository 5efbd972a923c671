// C hooks over the host layer so the parameter parser and the text writers can be checked from the test
// suite without a GPU (tests/test_host_layer.py).
#include <cstring>
#include <string>
#include "MakeDensity.h"

extern "C" {
// ---- driver handle for the multi-GPU launcher (supermc_b200/launch.py): one per rank ----
struct smc_host { ParameterReader rdr; MakeDensity* md; std::string err; };
smc_host* smc_host_create(const char* parameter_file, int argc, char** argv, int device, int rank, int world, const char* data_dir) {
  smc_host* h = new smc_host(); h->md = nullptr;
  try { h->rdr.readFromFile(parameter_file); h->rdr.readFromArguments(argc, argv, "#", 0); }
  catch (std::exception& e) { h->err = e.what(); return h; }
  h->md = new MakeDensity(&h->rdr, device, smc_shard{rank, world}, data_dir);
  if (!h->md->ok()) { h->err = h->md->error(); delete h->md; h->md = nullptr; }
  return h;
}
const char* smc_host_error(smc_host* h) { return h ? (h->md && !h->md->error().empty() ? h->md->error().c_str() : h->err.c_str()) : "null"; }
void smc_host_destroy(smc_host* h) { if (h) { delete h->md; delete h; } }
smc_ctx* smc_host_context(smc_host* h) { return (h && h->md) ? h->md->context() : nullptr; }
int smc_host_run(smc_host* h) {            // operations 1, 2, 9 (and 3 on one GPU)
  if (!h || !h->md) return 1;
  try { return h->md->run((int)h->rdr.getVal("operation"), (int)h->rdr.getVal("nev")); } catch (std::exception& e) { h->err = e.what(); return 1; }
}
int smc_host_average_accumulate(smc_host* h) { if (!h || !h->md) return 1; try { return h->md->average_accumulate((int)h->rdr.getVal("nev")); } catch (std::exception& e) { h->err = e.what(); return 1; } }
int smc_host_average_write(smc_host* h) { if (!h || !h->md) return 1; try { return h->md->average_write(); } catch (std::exception& e) { h->err = e.what(); return 1; } }
int smc_host_get(smc_host* h, const char* name, double* v) { if (!h) return 1; try { *v = h->rdr.getVal(name); return 0; } catch (std::exception&) { return 1; } }
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
  return copy_out(kind == 0 ? smc_fmt_xy(rows, n, stride) : kind == 1 ? smc_fmt_participants(rows, n) : kind == 3 ? smc_fmt_quarks(rows, n) : smc_fmt_spectators(rows, n), out, cap);
}
int smc_host_format_block(const double* g, int Maxx, int Maxy, char* out, int cap) { std::string s; MakeDensity::formatDensityBlock(g, Maxx, Maxy, s); return copy_out(s, out, cap); }
int smc_host_format_4col(const double* g, int Maxx, int Maxy, double Xmin, double Ymin, double dx, double dy, double rap, double npart, char* out, int cap) {
  std::string s; MakeDensity::formatDensity4Col(g, Maxx, Maxy, Xmin, Ymin, dx, dy, rap, npart, s); return copy_out(s, out, cap);
}
}
