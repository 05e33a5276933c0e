// nmp_tables.cpp — host-only readers of the Noah-MP parameter tables (no GPU needed):
//   read_mp_veg_parameters   phys/module_sf_noahmplsm.F90:274-404   MPTABLE.TBL   (Fortran NAMELIST)
//   SOIL_VEG_GEN_PARM        phys/module_sf_noahmpdrv.F90:1528-1821 VEGPARM.TBL, SOILPARM.TBL, GENPARM.TBL
//                                                                   (list-directed records)
// The file semantics are the contract: '!' comments, blank/comma separated values,
// values continuing over lines, arrays pre-set to -1.E36 and filled in Fortran element order, and the
// reshape fix-up of the 2-D arrays when NVEG < MVT.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "noahmp_b200.h"

namespace {

std::string g_err;

std::string lower(std::string s) {
  for (auto& ch : s) ch = (char)tolower((unsigned char)ch);
  return s;
}

float to_f32(std::string t) {
  for (auto& ch : t)
    if (ch == 'd' || ch == 'D') ch = 'e';
  return (float)strtod(t.c_str(), nullptr);
}

std::vector<std::string> read_lines(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<std::string> lines;
  std::string l;
  while (std::getline(f, l)) {
    if (!l.empty() && l.back() == '\r') l.pop_back();
    lines.push_back(l);
  }
  return lines;
}

// ---- Fortran NAMELIST subset ------------------------------------------------------------------------
std::string strip_comment(const std::string& line) {
  std::string out;
  char q = 0;
  for (char ch : line) {
    if (q) {
      out.push_back(ch);
      if (ch == q) q = 0;
    } else if (ch == '"' || ch == '\'') {
      q = ch;
      out.push_back(ch);
    } else if (ch == '!') {
      break;
    } else {
      out.push_back(ch);
    }
  }
  return out;
}

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

std::map<std::string, std::string> namelist_groups(const std::vector<std::string>& lines) {
  std::map<std::string, std::string> groups;
  bool in = false;
  std::string name, cur;
  for (const auto& raw : lines) {
    std::string line = trim(strip_comment(raw));
    if (line.empty()) continue;
    if (!in) {
      if (line[0] == '&') {
        size_t p = 1;
        while (p < line.size() && isspace((unsigned char)line[p])) ++p;
        size_t e = p;
        while (e < line.size() && (isalnum((unsigned char)line[e]) || line[e] == '_')) ++e;
        name = lower(line.substr(p, e - p));
        cur = trim(line.substr(e));
        in = true;
      }
      continue;
    }
    if (line[0] == '/') {
      groups[name] = cur;
      in = false;
      continue;
    }
    cur += " " + line;
  }
  return groups;
}

// tokens: quoted strings, '=', or runs of non-space/non-comma/non-'=' characters
std::vector<std::string> nl_tokens(const std::string& body) {
  std::vector<std::string> t;
  size_t i = 0;
  while (i < body.size()) {
    char ch = body[i];
    if (isspace((unsigned char)ch) || ch == ',') { ++i; continue; }
    if (ch == '=') { t.push_back("="); ++i; continue; }
    if (ch == '"' || ch == '\'') {
      size_t e = body.find(ch, i + 1);
      if (e == std::string::npos) e = body.size() - 1;
      t.push_back(body.substr(i, e - i + 1));
      i = e + 1;
      continue;
    }
    size_t e = i;
    while (e < body.size() && !isspace((unsigned char)body[e]) && body[e] != ',' && body[e] != '=') ++e;
    t.push_back(body.substr(i, e - i));
    i = e;
  }
  return t;
}

std::map<std::string, std::vector<std::string>> namelist_assignments(const std::string& body) {
  std::map<std::string, std::vector<std::string>> out;
  auto toks = nl_tokens(body);
  std::string cur;
  for (size_t i = 0; i < toks.size();) {
    if (i + 1 < toks.size() && toks[i + 1] == "=") {
      cur = lower(toks[i]);
      out[cur];
      i += 2;
      continue;
    }
    if (cur.empty()) throw std::runtime_error("namelist value before any name");
    const std::string& t = toks[i];
    size_t star = t.find('*');
    bool rep = star != std::string::npos && star > 0;
    for (size_t k = 0; rep && k < star; ++k)
      if (!isdigit((unsigned char)t[k])) rep = false;
    if (rep) {
      int n = atoi(t.substr(0, star).c_str());
      for (int k = 0; k < n; ++k) out[cur].push_back(t.substr(star + 1));
    } else {
      out[cur].push_back(t);
    }
    ++i;
  }
  return out;
}

const float UNDEF = -1.0e36f;

void fill_1d(const std::map<std::string, std::vector<std::string>>& par, const char* name, float* dst) {
  for (int k = 0; k < NOAHMP_MVT; ++k) dst[k] = UNDEF;
  auto it = par.find(name);
  if (it == par.end()) return;
  for (size_t k = 0; k < it->second.size() && k < (size_t)NOAHMP_MVT; ++k) dst[k] = to_f32(it->second[k]);
}

// X(MVT,ncol) in Fortran element order; dst[ncol][MVT]
void fill_2d(const std::map<std::string, std::vector<std::string>>& par, const char* name, int ncol, int nveg,
             float* dst) {
  std::vector<float> flat((size_t)NOAHMP_MVT * ncol, UNDEF);
  auto it = par.find(name);
  if (it != par.end())
    for (size_t k = 0; k < it->second.size() && k < flat.size(); ++k) flat[k] = to_f32(it->second[k]);
  if (NOAHMP_MVT > nveg) {  // reshape fix-up, noahmplsm.F90:373-402
    for (int c = 0; c < ncol; ++c)
      for (int v = 0; v < NOAHMP_MVT; ++v) dst[c * NOAHMP_MVT + v] = v < nveg ? flat[(size_t)c * nveg + v] : UNDEF;
  } else {
    for (size_t k = 0; k < flat.size(); ++k) dst[k] = flat[k];
  }
}

void read_mptable(const std::string& path, const std::string& dataset, noahmp_tables* T) {
  std::string gcat, gpar;
  if (dataset == "USGS") { gcat = "noah_mp_usgs_veg_categories"; gpar = "noah_mp_usgs_parameters"; }
  else if (dataset == "MODIFIED_IGBP_MODIS_NOAH") { gcat = "noah_mp_modis_veg_categories"; gpar = "noah_mp_modis_parameters"; }
  else throw std::runtime_error("Unrecognized DATASET_IDENTIFIER in subroutine READ_MP_VEG_PARAMETERS: " + dataset);
  auto groups = namelist_groups(read_lines(path));
  if (!groups.count(gcat) || !groups.count(gpar)) throw std::runtime_error("namelist group missing in " + path);
  auto cat = namelist_assignments(groups[gcat]);
  auto par = namelist_assignments(groups[gpar]);
  auto geti = [&](std::map<std::string, std::vector<std::string>>& m, const char* k) {
    auto it = m.find(k);
    if (it == m.end() || it->second.empty()) throw std::runtime_error(std::string("missing ") + k);
    return atoi(it->second[0].c_str());
  };
  const int nveg = geti(cat, "nveg");
  T->nveg = nveg;
  T->isurban_mp = geti(par, "isurban");
  T->iswater = geti(par, "iswater");
  T->isbarren = geti(par, "isbarren");
  T->issnow = geti(par, "issnow");
  T->eblforest = geti(par, "eblforest");
#define F1(n) fill_1d(par, #n, T->n)
  F1(ch2op); F1(dleaf); F1(z0mvt); F1(hvt); F1(hvb); F1(den); F1(rc);
  F1(xl); F1(cwpvt); F1(c3psn); F1(kc25); F1(akc); F1(ko25); F1(ako); F1(avcmx); F1(aqe); F1(ltovrc); F1(dilefc);
  F1(dilefw); F1(rmf25); F1(sla); F1(fragr); F1(tmin); F1(vcmx25); F1(tdlef); F1(bp); F1(mp); F1(qe25); F1(rms25);
  F1(rmr25); F1(arm); F1(folnmx); F1(wdpool); F1(wrrat); F1(mrp); F1(slarea);
#undef F1
  fill_2d(par, "rhol", 2, nveg, &T->rhol[0][0]);
  fill_2d(par, "rhos", 2, nveg, &T->rhos[0][0]);
  fill_2d(par, "taul", 2, nveg, &T->taul[0][0]);
  fill_2d(par, "taus", 2, nveg, &T->taus[0][0]);
  fill_2d(par, "saim", 12, nveg, &T->saim[0][0]);
  fill_2d(par, "laim", 12, nveg, &T->laim[0][0]);
  fill_2d(par, "eps", 5, nveg, &T->eps[0][0]);
}

// ---- list-directed records ----------------------------------------------------------------------------
struct Records {
  std::vector<std::string> lines;
  size_t pos = 0;
  explicit Records(const std::string& path) : lines(read_lines(path)) {}
  void skip(size_t n = 1) {
    pos += n;
    if (pos > lines.size()) throw std::out_of_range("eof");
  }
  // READ(unit,*) of n items: tokens from successive records until n are found
  std::vector<std::string> read(size_t n) {
    std::vector<std::string> items;
    while (items.size() < n) {
      if (pos >= lines.size()) throw std::out_of_range("eof");
      const std::string& l = lines[pos++];
      size_t i = 0;
      while (i < l.size()) {
        char ch = l[i];
        if (isspace((unsigned char)ch) || ch == ',') { ++i; continue; }
        if (ch == '\'' || ch == '"') {
          size_t e = l.find(ch, i + 1);
          if (e == std::string::npos) e = l.size() - 1;
          items.push_back(l.substr(i, e - i + 1));
          i = e + 1;
          continue;
        }
        size_t e = i;
        while (e < l.size() && !isspace((unsigned char)l[e]) && l[e] != ',') ++e;
        items.push_back(l.substr(i, e - i));
        i = e;
      }
    }
    items.resize(n);
    return items;
  }
};

void read_vegparm(const std::string& path, const std::string& mminlu, noahmp_tables* T) {
  Records r(path);
  int lucats = 0;
  for (;;) {
    std::string lutype;
    try {
      r.skip();
      lutype = r.read(1)[0];
      auto v = r.read(2);
      lucats = atoi(v[0].c_str());
    } catch (const std::out_of_range&) {
      throw std::runtime_error("Land Use Dataset '" + mminlu + "' not found in VEGPARM.TBL.");
    }
    if (lutype == mminlu) break;
    r.skip((size_t)lucats + 12);
  }
  if (lucats > NOAHMP_NLUS) throw std::runtime_error("Table sizes too small for value of LUCATS");
  T->lucats = lucats;
  float* cols[17] = {T->shdtbl, nullptr, T->rstbl, T->rgltbl, T->hstbl, T->snuptbl, T->maxalb, T->laimintbl,
                     T->laimaxtbl, T->emissmintbl, T->emissmaxtbl, T->albedomintbl, T->albedomaxtbl, T->z0mintbl,
                     T->z0maxtbl, T->ztopvtbl, T->zbotvtbl};
  for (int c = 0; c < 17; ++c)
    if (cols[c]) memset(cols[c], 0, sizeof(float) * NOAHMP_NLUS);
  memset(T->nrotbl, 0, sizeof(T->nrotbl));
  for (int lc = 0; lc < lucats; ++lc) {
    auto t = r.read(18);
    for (int c = 0; c < 17; ++c) {
      if (c == 1) T->nrotbl[lc] = (int)strtod(t[1 + c].c_str(), nullptr);
      else cols[c][lc] = to_f32(t[1 + c]);
    }
  }
  float* sc[4] = {&T->topt_data, &T->cmcmax_data, &T->cfactr_data, &T->rsmax_data};
  for (auto p : sc) { r.skip(); *p = to_f32(r.read(1)[0]); }
  r.skip(); T->bare = atoi(r.read(1)[0].c_str());
  r.skip(); T->natural = atoi(r.read(1)[0].c_str());
}

void read_soilparm(const std::string& path, const std::string& mminsl, noahmp_tables* T) {
  Records r(path);
  r.skip();
  if (r.pos >= r.lines.size()) throw std::runtime_error("INCONSISTENT OR MISSING SOILPARM FILE");
  std::string sltype = trim(r.lines[r.pos].substr(0, 4));  // FORMAT(A4)
  r.skip();
  auto v = r.read(2);
  int slcats = atoi(v[0].c_str());
  if (sltype != mminsl) throw std::runtime_error("INCONSISTENT OR MISSING SOILPARM FILE");
  if (slcats > NOAHMP_NSLTYPE) throw std::runtime_error("Table sizes too small for value of SLCATS");
  T->slcats = slcats;
  float* cols[10] = {T->bb, T->drysmc, T->f11, T->maxsmc, T->refsmc, T->satpsi, T->satdk, T->satdw, T->wltsmc, T->qtz};
  for (auto c : cols) memset(c, 0, sizeof(float) * NOAHMP_NSLTYPE);
  for (int lc = 0; lc < slcats; ++lc) {
    auto t = r.read(11);
    for (int c = 0; c < 10; ++c) cols[c][lc] = to_f32(t[1 + c]);
  }
}

void read_genparm(const std::string& path, noahmp_tables* T) {
  Records r(path);
  r.skip(2);
  int num_slope = atoi(r.read(1)[0].c_str());
  if (num_slope > NOAHMP_NSLOPE) throw std::runtime_error("NUM_SLOPE too large for slope_data array");
  T->slpcats = num_slope;
  memset(T->slope_data, 0, sizeof(T->slope_data));
  for (int lc = 0; lc < num_slope; ++lc) T->slope_data[lc] = to_f32(r.read(1)[0]);
  float* sc[12] = {&T->sbeta_data, &T->fxexp_data, &T->csoil_data, &T->salp_data, &T->refdk_data, &T->refkdt_data,
                   &T->frzk_data, &T->zbot_data, &T->czil_data, &T->smlow_data, &T->smhigh_data, &T->lvcoef_data};
  for (auto p : sc) { r.skip(); *p = to_f32(r.read(1)[0]); }
}

}  // namespace

extern "C" {

const char* noahmp_b200_tables_error(void) { return g_err.c_str(); }

int noahmp_b200_read_tables(const char* dir, const char* dataset, const char* soil, noahmp_tables* out) {
  if (!dir || !dataset || !soil || !out) return NOAHMP_ERR_ARG;
  try {
    std::string d(dir);
    if (!d.empty() && d.back() != '/') d += '/';
    memset(out, 0, sizeof(*out));
    read_mptable(d + "MPTABLE.TBL", dataset, out);
    read_vegparm(d + "VEGPARM.TBL", dataset, out);
    read_soilparm(d + "SOILPARM.TBL", soil, out);
    read_genparm(d + "GENPARM.TBL", out);
  } catch (const std::exception& e) {
    g_err = e.what();
    return NOAHMP_ERR_ARG;
  }
  return 0;
}

unsigned long long noahmp_b200_sizeof_tables(void) { return sizeof(noahmp_tables); }
unsigned long long noahmp_b200_sizeof_args(void) { return sizeof(noahmp_lsm_args); }
}
