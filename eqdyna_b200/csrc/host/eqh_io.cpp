// Input readers and Fortran-format output writers of the stand-in host.
//
// Readers follow src/readInputFiles.f90 (list-directed reads of the five
// b*.txt files written by scripts/case.setup:13-76) and the slot mapping of
// src/netcdf_io.f90:70-107.  The container has no netCDF library, so the 24
// on-fault fields come from `on_fault_vars_input.bin`:
//     char[8]  "EQDOFV1\0"
//     int32    nfx, nfz, nvar(=24), 0
//     float64  field[nvar][nfz][nfx]      (C order == Fortran (nfx,nfz,nvar))
// in the order of var_id(1..24) in netcdf_io.f90:41-64.
// (tools/gen_case_fixtures.py writes it from the arrays scripts/case.setup hands to netCDF4.)
// mode == 2 additionally reads `fault.r.bin`, the same container with the 12 fields of the
// restart file fault.r.nc (netcdf_io.f90:116-185).
//
// Writers reproduce src/library_output.f90:16-205 (faultst*.txt, body*.txt,
// frt.txt<me>) so that scripts/plotRuptureDynamics keeps working.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "eqh_state.h"

namespace eqh {

namespace {

// Minimal emulation of Fortran list-directed READ on a formatted file:
// `skip()` == `read(u,*)` with an empty list (consumes one record);
// `get(n)` consumes whole records until n values were found.
class ListReader {
 public:
  explicit ListReader(const std::string& path) : path_(path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error(path + " is required but missing");
    std::string line;
    while (std::getline(f, line)) lines_.push_back(line);
  }
  void skip() { ++pos_; }
  std::vector<double> get(int n) {
    std::vector<double> out;
    while ((int)out.size() < n) {
      if (pos_ >= lines_.size()) throw std::runtime_error("unexpected end of " + path_);
      std::string line = lines_[pos_++];
      for (char& c : line)
        if (c == ',' || c == '\t') c = ' ';
      std::istringstream is(line);
      std::string tok;
      while ((int)out.size() < n && (is >> tok)) {
        for (char& c : tok)
          if (c == 'd' || c == 'D') c = 'e';
        out.push_back(std::stod(tok));
      }
    }
    return out;
  }

 private:
  std::string path_;
  std::vector<std::string> lines_;
  size_t pos_ = 0;
};

inline int nint(double x) { return (int)std::lround(x); }

}  // namespace

void read_case(const std::string& dir, CaseInput& in) {
  in.dir = dir;
  in.pi = 4 * std::atan(1.0);  // globalvar.f90:15
  {
    // readglobal, readInputFiles.f90:28-61
    ListReader r(dir + "/bGlobal.txt");
    in.mode = (int)r.get(1)[0];
    in.C_elastic = (int)r.get(1)[0];
    in.C_nuclea = (int)r.get(1)[0];
    in.C_degen = r.get(1)[0];
    in.insertFaultType = (int)r.get(1)[0];
    in.friclaw = (int)r.get(1)[0];
    in.ntotft = (int)r.get(1)[0];
    in.nucfault = (int)r.get(1)[0];
    in.TPV = (int)r.get(1)[0];
    in.output_plastic = (int)r.get(1)[0];
    in.outputGroundMotion = (int)r.get(1)[0];
    in.outputFinalSurfDisp = (int)r.get(1)[0];
    r.skip();
    auto v = r.get(3);
    in.npx = (int)v[0]; in.npy = (int)v[1]; in.npz = (int)v[2];
    r.skip();
    in.totalSimuTime = r.get(1)[0];
    in.dt = r.get(1)[0];
    r.skip();
    v = r.get(2);
    in.nmat = (int)v[0]; in.n2mat = (int)v[1];
    v = r.get(3);
    in.roumax = v[0]; in.rhow = v[1]; in.gamar = v[2];
    v = r.get(2);
    in.rdampk = v[0]; in.vmaxPML = v[1];
    r.skip();
    v = r.get(3);
    in.xsource = v[0]; in.ysource = v[1]; in.zsource = v[2];
    v = r.get(4);
    in.nucR = v[0]; in.nucRuptVel = v[1]; in.nucdtau0 = v[2]; in.nucT = v[3];
    v = r.get(2);
    in.str1ToFaultAngle = v[0]; in.devStrToStrVertRatio = v[1];
    v = r.get(2);
    in.bulk = v[0]; in.coheplas = v[1];
    v = r.get(2);
    in.fstrike = v[0]; in.fdip = v[1];
    in.slipRateThres = r.get(1)[0];
    in.str1ToFaultAngle = in.str1ToFaultAngle * in.pi / 180.0;
  }
  {
    // readmodelgeometry, readInputFiles.f90:87-95
    ListReader r(dir + "/bModelGeometry.txt");
    auto v = r.get(2); in.xmin = v[0]; in.xmax = v[1];
    v = r.get(2); in.ymin = v[0]; in.ymax = v[1];
    v = r.get(2); in.zmin = v[0]; in.zmax = v[1];
    r.skip();
    v = r.get(2); in.dis4uniF = (int)v[0]; in.dis4uniB = (int)v[1];
    in.rat = r.get(1)[0];
    v = r.get(3); in.dx = v[0]; in.dy = v[1]; in.dz = v[2];
  }
  {
    // readfaultgeometry, readInputFiles.f90:124-146
    ListReader r(dir + "/bFaultGeometry.txt");
    int n = in.ntotft;
    in.fxmin.resize(n); in.fxmax.resize(n); in.fymin.resize(n);
    in.fymax.resize(n); in.fzmin.resize(n); in.fzmax.resize(n);
    in.fltxyz.assign(2 * 4 * n, 0.0);
    for (int i = 0; i < n; ++i) {
      r.skip();
      auto v = r.get(2); in.fxmin[i] = v[0]; in.fxmax[i] = v[1];
      v = r.get(2); in.fymin[i] = v[0]; in.fymax[i] = v[1];
      v = r.get(2); in.fzmin[i] = v[0]; in.fzmax[i] = v[1];
    }
    for (int i = 0; i < n; ++i) {
      double* f = &in.fltxyz[8 * i];  // f[(a-1)+2*(b-1)] = fltxyz(a,b,i)
      f[0] = in.fxmin[i]; f[1] = in.fxmax[i];
      f[2] = in.fymin[i]; f[3] = in.fymax[i];
      f[4] = in.fzmin[i]; f[5] = in.fzmax[i];
      f[6] = in.fstrike * in.pi / 180.0;
      if (in.C_degen > 3.0) f[7] = in.C_degen * in.pi / 180.0;
      else f[7] = 90.0 * in.pi / 180.0;
    }
  }
  {
    // readmaterial, readInputFiles.f90:176-186
    ListReader r(dir + "/bMaterial.txt");
    in.material.assign((size_t)in.nmat * in.n2mat, 0.0);
    for (int i = 0; i < in.nmat; ++i) {
      auto v = r.get(in.n2mat);
      for (int j = 0; j < in.n2mat; ++j) in.material[i + (size_t)in.nmat * j] = v[j];
    }
    in.ccosphi = in.coheplas * std::cos(std::atan(in.bulk));
    in.sinphi = std::sin(std::atan(in.bulk));
    in.nstep = nint(in.totalSimuTime / in.dt);
    in.rdampk = in.rdampk * in.dt;
    in.tv = 2.0 * in.dz / 3464.0;
  }
  {
    // readstations1/2, readInputFiles.f90:215-218,247-263
    ListReader r(dir + "/bStations.txt");
    in.totalNumOfOffSt = (int)r.get(1)[0];
    auto v = r.get(in.ntotft);
    in.nonfs.resize(in.ntotft);
    int mx = 0;
    for (int i = 0; i < in.ntotft; ++i) { in.nonfs[i] = (int)v[i]; mx = std::max(mx, in.nonfs[i]); }
    in.xonfs.assign((size_t)2 * std::max(mx, 1) * in.ntotft, 0.0);
    in.x4nds.assign((size_t)3 * std::max(in.totalNumOfOffSt, 1), 0.0);
    ListReader r2(dir + "/bStations.txt");
    r2.skip(); r2.skip(); r2.skip();
    for (int i = 0; i < in.ntotft; ++i)
      for (int j = 0; j < in.nonfs[i]; ++j) {
        auto c = r2.get(2);
        in.xonfs[0 + 2 * (j + (size_t)mx * i)] = c[0] * 1000.0;
        in.xonfs[1 + 2 * (j + (size_t)mx * i)] = c[1] * 1000.0;
      }
    r2.skip();
    for (int i = 0; i < in.totalNumOfOffSt; ++i) {
      auto c = r2.get(3);
      for (int k = 0; k < 3; ++k) in.x4nds[k + 3 * (size_t)i] = c[k] * 1000.0;
    }
  }
  if (in.insertFaultType > 0) {
    // read_fault_rough_geometry, readInputFiles.f90:290-311
    ListReader r(dir + "/bFault_Rough_Geometry.txt");
    auto v = r.get(2);
    in.nnx = nint(v[0]); in.nnz = nint(v[1]);
    r = ListReader(dir + "/bFault_Rough_Geometry.txt");
    r.skip();
    v = r.get(3);
    in.dxtmp = v[0]; in.rough_fx_min = v[1]; in.rough_fz_min = v[2];
    in.rough_fx_max = (in.nnx - 1) * in.dxtmp + in.rough_fx_min;
    in.rough_geo.assign((size_t)3 * in.nnx * in.nnz, 0.0);
    for (int i = 0; i < in.nnx * in.nnz; ++i) {
      auto c = r.get(3);
      for (int k = 0; k < 3; ++k) in.rough_geo[k + 3 * (size_t)i] = c[k];
    }
  }
  {
    // netcdf_read_on_fault_eqdyna, netcdf_io.f90:30-68 (raw dump, see header)
    in.fnx = nint((in.fxmax[0] - in.fxmin[0]) / in.dx) + 1;
    in.fnz = nint((in.fzmax[0] - in.fzmin[0]) / in.dz) + 1;
    std::string p = dir + "/on_fault_vars_input.bin";
    FILE* f = std::fopen(p.c_str(), "rb");
    if (!f) throw std::runtime_error(p + " is required but missing");
    char magic[8];
    int32_t hdr[4];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "EQDOFV1", 7) != 0 ||
        std::fread(hdr, 4, 4, f) != 4) {
      std::fclose(f);
      throw std::runtime_error(p + ": bad header");
    }
    if (hdr[0] != in.fnx || hdr[1] != in.fnz || hdr[2] != 24) {
      std::fclose(f);
      throw std::runtime_error(p + ": dimensions do not match bFaultGeometry/dx,dz");
    }
    size_t n = (size_t)in.fnx * in.fnz * 24;
    in.on_fault_vars.resize(n);
    size_t got = std::fread(in.on_fault_vars.data(), 8, n, f);
    std::fclose(f);
    if (got != n) throw std::runtime_error(p + ": truncated");
  }
  if (in.mode == 2) {
    // netcdf_read_on_fault_eqdyna_restart, netcdf_io.f90:116-150: fault.r.nc of the previous
    // earthquake cycle, here its raw dump fault.r.bin (same container as above, 12 fields in the
    // order shear_strike, shear_dip, effective_normal, slip_rate, state_variable, state_normal,
    // vxm, vym, vzm, vxs, vys, vzs).  Required in this mode, as the reference's nf90_open is.
    std::string p = dir + "/fault.r.bin";
    FILE* f = std::fopen(p.c_str(), "rb");
    if (!f) throw std::runtime_error(p + " is required for mode == 2 but missing");
    char magic[8];
    int32_t hdr[4];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "EQDOFV1", 7) != 0 || std::fread(hdr, 4, 4, f) != 4) {
      std::fclose(f);
      throw std::runtime_error(p + ": bad header");
    }
    if (hdr[0] != in.fnx || hdr[1] != in.fnz || hdr[2] != 12) {
      std::fclose(f);
      throw std::runtime_error(p + ": dimensions do not match bFaultGeometry/dx,dz (12 fields expected)");
    }
    size_t n = (size_t)in.fnx * in.fnz * 12;
    in.restart_vars.resize(n);
    size_t got = std::fread(in.restart_vars.data(), 8, n, f);
    std::fclose(f);
    if (got != n) throw std::runtime_error(p + ": truncated");
  }
}

// ---------------------------------------------------------------------------
// Fortran E editing: Ew.d (2-digit exponent, 'E' dropped for 3 digits) and
// Ew.dEe.  Built on a correctly rounded %.{d-1}e conversion.
namespace {
std::string fortran_e(double v, int w, int d, int e) {
  char buf[64];
  std::string out;
  if (std::isnan(v)) out = "NaN";
  else if (std::isinf(v)) out = v < 0 ? "-Infinity" : "Infinity";
  else {
    std::snprintf(buf, sizeof buf, "%.*e", d - 1, std::fabs(v));
    // buf = D.DDDDDDe+XX
    std::string digits;
    int ex = 0;
    const char* p = buf;
    for (; *p && *p != 'e'; ++p)
      if (*p >= '0' && *p <= '9') digits.push_back(*p);
    if (*p == 'e') ex = std::atoi(p + 1);
    if (v != 0.0) ex += 1;
    else ex = 0;
    std::string s = std::signbit(v) ? "-0." : "0.";
    s += digits;
    char eb[16];
    int ae = std::abs(ex);
    if (e > 0) {
      std::snprintf(eb, sizeof eb, "E%c%0*d", ex < 0 ? '-' : '+', e, ae);
    } else if (ae <= 99) {
      std::snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', ae);
    } else {
      std::snprintf(eb, sizeof eb, "%c%03d", ex < 0 ? '-' : '+', ae);
    }
    out = s + eb;
  }
  if ((int)out.size() > w) {
    // gfortran drops the optional leading zero before giving up
    if (out.size() >= 2 && out[0] == '0' && out[1] == '.' && (int)out.size() - 1 <= w) out = out.substr(1);
    else if (out.size() >= 3 && out[0] == '-' && out[1] == '0' && (int)out.size() - 1 <= w) out = "-" + out.substr(2);
    else return std::string(w, '*');
  }
  return std::string(w - out.size(), ' ') + out;
}
}  // namespace

// output_frt, library_output.f90:160-205, format (1x,22e18.7e4)
void write_frt(const RankState& s, const std::string& dir) {
  if (s.nftnd.empty() || s.nftnd[0] <= 0) return;
  std::string p = dir + "/frt.txt" + std::to_string(s.me);
  FILE* f = std::fopen(p.c_str(), "w");
  if (!f) throw std::runtime_error("cannot write " + p);
  const int cols[] = {71, 72, 73, 74, 75, 76, 47, 78, 79, 80, 31, 32, 33, 34, 35, 36, 20, 23};
  for (int i = 0; i < s.nftnd[0]; ++i) {
    std::string line = " ";
    int slave = s.nsmp[0 + 2 * (size_t)i];
    for (int j = 0; j < 3; ++j) line += fortran_e(s.meshCoor[j + 3 * (size_t)(slave - 1)], 18, 7, 4);
    line += fortran_e(s.fnft[i], 18, 7, 4);
    for (int c : cols) line += fortran_e(s.fric[(c - 1) + 100 * (size_t)i], 18, 7, 4);
    std::fprintf(f, "%s\n", line.c_str());
  }
  std::fclose(f);
}

namespace {
std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(' ');
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(' ');
  return s.substr(a, b - a + 1);
}
void write_header(FILE* f, const RankState& s, bool off) {
  const CaseInput& in = *s.in;
  std::fprintf(f, " # Project=San-Ti                        \n");
  std::fprintf(f, " # Author=Sophon                        \n");
  std::fprintf(f, " # date = (stand-in host)\n");
  std::fprintf(f, " # code = EQdyna\n");
  std::fprintf(f, " # element_size =%25.14f     \n", in.dx);
  if (off) {
    std::fprintf(f, "   # time_step=%8.4f s\n", in.dt);
    std::fprintf(f, "   # num_time_steps=%6d\n", in.nstep);
  } else {
    std::fprintf(f, " # time_step =%8.4f  s\n", in.dt);
    std::fprintf(f, " # num_time_steps =%6d\n", in.nstep);
  }
}
}  // namespace

// output_onfault_st, library_output.f90:16-96
void write_onfault_stations(const RankState& s, const std::string& dir) {
  const CaseInput& in = *s.in;
  if (s.numOfOnFaultStCount <= 0) return;
  int mx = 0;
  for (int v : in.nonfs) mx = std::max(mx, v);
  for (int i = 0; i < s.numOfOnFaultStCount; ++i) {
    int j = s.anonfs[2 + 3 * (size_t)i];
    if (j != 1) continue;
    int ist = s.anonfs[1 + 3 * (size_t)i];
    double xs = in.xonfs[0 + 2 * ((ist - 1) + (size_t)mx * (j - 1))];
    double zs = in.xonfs[1 + 2 * ((ist - 1) + (size_t)mx * (j - 1))];
    char st[16], dp[16];
    std::snprintf(st, sizeof st, "%03d", nint(xs / 100.0));
    std::snprintf(dp, sizeof dp, "%03d", nint(std::fabs(zs) / std::sin(in.fltxyz[7]) / 100.0));
    // Fortran i3.3 of a negative value prints '-' + digits only if it fits; mimic "***" fallback
    std::string sst = nint(xs / 100.0) < 0 ? (std::abs(nint(xs / 100.0)) > 99 ? "***" : st) : st;
    std::string p = dir + "/faultst" + trim(sst) + "dp" + trim(dp) + ".txt";
    FILE* f = std::fopen(p.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + p);
    write_header(f, s, false);
    std::fprintf(f, " # Time series in 11 columns in format E15.7\n");
    std::fprintf(f, " # Column #1 = Time (s)\n # Column #2 = horizontal slip (m)\n"
                    " # Column #3 = horizontal slip rate (m/s)\n # Column #4 = horizontal shear stress (MPa)\n"
                    " # Column #5 = down-dip slip (m)\n # Column #6 = down-dip slip rate (m/s)\n"
                    " # Column #7 = down-dip shear stress (MPa)\n # Column #8 = normal stress (MPa)\n");
    const double* q = &s.onFaultQuantHistSCECForm[(size_t)12 * in.nstep * i];
    if (in.friclaw >= 3) {
      std::fprintf(f, " # Column #9 = state variable psi (dimensionless)\n"
                      " # Column #10 = Temperature (degrees Kelvin)\n # Column #11 = Pore pressure (MPa)\n");
      std::fprintf(f, " # The line below lists the names of the data fields:\n");
      std::fprintf(f, " t h-slip h-slip-rate h-shear-stress v-slip v-slip-rate v-shear-stress n-stress psi temperature pressure\n");
    } else {
      std::fprintf(f, " # The line below lists the names of the data fields:\n");
      std::fprintf(f, " t h-slip h-slip-rate h-shear-stress v-slip v-slip-rate v-shear-stress n-stress\n");
    }
    for (int n = 0; n < in.nstep; ++n) {
      const double* r = q + (size_t)12 * n;  // r[k-1] = onFaultQuantHistSCECForm(k,n+1,i+1)
      std::string line = fortran_e(r[0], 21, 13, 0);
      line += fortran_e(r[4], 16, 7, 0);
      line += fortran_e(r[1], 16, 7, 0);
      line += fortran_e(r[7] / 1.0e6, 16, 7, 0);
      line += fortran_e(-r[5], 16, 7, 0);
      line += fortran_e(-r[2], 16, 7, 0);
      line += fortran_e(-r[8] / 1.0e6, 16, 7, 0);
      line += fortran_e(-r[9] / 1.0e6, 16, 7, 0);
      if (in.friclaw >= 3) {
        line += fortran_e(r[3], 16, 7, 0);
        line += fortran_e(r[11], 16, 7, 0);
        line += fortran_e(r[10] / 1.0e6, 16, 7, 0);
      }
      std::fprintf(f, "%s\n", line.c_str());
    }
    std::fclose(f);
  }
}

// output_offfault_st, library_output.f90:99-157
void write_offfault_stations(const RankState& s, const std::string& dir) {
  const CaseInput& in = *s.in;
  if (s.numOfOffFaultStCount <= 0) return;
  int rows = s.numOfOffFaultStCount * 6 + 1;
  for (int i = 0; i < s.numOfOffFaultStCount; ++i) {
    int ist = s.OffFaultStNodeIdIndex[0 + 2 * (size_t)i];
    const double* x = &in.x4nds[3 * (size_t)(ist - 1)];
    char b[16], st[16], dp[16];
    // '(i4.3)' of int(x/100): width 4, at least 3 digits
    auto i43 = [](char* out, int v) {
      char t[16];
      std::snprintf(t, sizeof t, "%s%03d", v < 0 ? "-" : "", std::abs(v));
      std::snprintf(out, 16, "%s", t);
    };
    i43(b, (int)(x[1] / 100.0));
    i43(st, (int)(x[0] / 100.0));
    i43(dp, (int)(std::fabs(x[2]) / 100.0));
    std::string p = dir + "/body" + trim(b) + "st" + trim(st) + "dp" + trim(dp) + ".txt";
    FILE* f = std::fopen(p.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + p);
    write_header(f, s, true);
    std::fprintf(f, " # Column #1 = Time (s)\n # Column #2 = horizontal displacement (m)\n"
                    " # Column #3 = horizontal displacement (m)\n # Column #3 = horizontal velocity (m/s)\n"
                    " # Column #4 = vertical displacement (m)\n # Column #5 = vertical velocity (m/s)\n"
                    " # Column #6 = normal displacement (m)\n # Column #7 = normal velocity (m/s)\n #\n"
                    " # The line below lists the names of the data fields:\n"
                    " t h-disp h-vel v-disp v-vel n-disp n-vel\n");
    for (int n = 0; n < in.nstep; ++n) {
      const double* r = &s.OffFaultStGramSCEC[(size_t)rows * n];  // r[k-1] = (k, n+1)
      std::string line = fortran_e(r[0], 21, 13, 0);
      line += fortran_e(r[i * 6 + 1], 16, 7, 0);
      line += fortran_e(r[i * 6 + 2], 16, 7, 0);
      line += fortran_e(-r[i * 6 + 5], 16, 7, 0);
      line += fortran_e(-r[i * 6 + 6], 16, 7, 0);
      line += fortran_e(r[i * 6 + 3], 16, 7, 0);
      line += fortran_e(r[i * 6 + 4], 16, 7, 0);
      std::fprintf(f, "%s\n", line.c_str());
    }
    std::fclose(f);
  }
}

// ---------------------------------------------------------------------------
// The remaining writers of library_output.f90.  The reference opens most of these
// with position='append' while it steps; here each file is written once, whole.
namespace {
FILE* open_rank_file(const RankState& s, const std::string& dir, const char* stem, const char* mode) {
  std::string p = dir + "/" + stem + std::to_string(s.me);   // 'name'//mm, mm = trimmed rank id (eqdyna3d.f90:95-96)
  FILE* f = std::fopen(p.c_str(), mode);
  if (!f) throw std::runtime_error("cannot write " + p);
  return f;
}
}  // namespace

// find_surfaceNodeIdArr, library_output.f90:247-264: (1x,3e18.7e4) per surface node
void write_surface_coor(const RankState& s, const std::string& dir) {
  const CaseInput& in = *s.in;
  if (!(in.outputGroundMotion == 1 || in.outputFinalSurfDisp == 1) || s.surface_nnode <= 0) return;
  FILE* f = open_rank_file(s, dir, "surface_coor.txt", "w");
  for (int i = 0; i < s.surface_nnode; ++i) {
    const double* c = &s.meshCoor[3 * (size_t)(s.surfaceNodeIdArr[i] - 1)];
    std::string line = " ";
    for (int j = 0; j < 3; ++j) line += fortran_e(c[j], 18, 7, 4);
    std::fprintf(f, "%s\n", line.c_str());
  }
  std::fclose(f);
}

// output_gm, library_output.f90:267-279: unformatted stream, velArr(1:3,node) of every
// surface node, one block per sampled step
void write_gm(const RankState& s, const std::string& dir) {
  if (s.in->outputGroundMotion != 1 || s.surface_nnode <= 0) return;
  FILE* f = open_rank_file(s, dir, "gm", "wb");
  const size_t n = (size_t)3 * s.surface_nnode * s.nGmSamples;
  if (n && std::fwrite(s.gmHist.data(), sizeof(double), n, f) != n) { std::fclose(f); throw std::runtime_error("short write: gm"); }
  std::fclose(f);
}

// output_src_evol, library_output.f90:297-312: fric(47,i,1), i = 1..nftnd(1), per sampled step
void write_src_evol(const RankState& s, const std::string& dir) {
  if (s.in->outputGroundMotion != 1 || s.nftnd.empty() || s.nftnd[0] <= 0) return;
  FILE* f = open_rank_file(s, dir, "src_evol", "wb");
  const size_t n = (size_t)s.nftnd[0] * s.nGmSamples;
  if (n && std::fwrite(s.srcEvolHist.data(), sizeof(double), n, f) != n) { std::fclose(f); throw std::runtime_error("short write: src_evol"); }
  std::fclose(f);
}

// output_finalSurfDisp, library_output.f90:282-295
void write_final_surf_disp(const RankState& s, const std::string& dir) {
  if (s.in->outputFinalSurfDisp != 1 || s.surface_nnode <= 0) return;
  FILE* f = open_rank_file(s, dir, "finalSurfDisp.txt", "w");
  for (int i = 0; i < s.surface_nnode; ++i) {
    const double* d = &s.dispArr[3 * (size_t)(s.surfaceNodeIdArr[i] - 1)];
    std::string line = " ";
    for (int j = 0; j < 3; ++j) line += fortran_e(d[j], 18, 7, 4);
    std::fprintf(f, "%s\n", line.c_str());
  }
  std::fclose(f);
}

// output_plastic_strain, library_output.f90:221-244: elements with pstrain > 1e-4 whose
// first node lies within |x| < 5 km, |y| < 2 km, |z| < 8 km: centroid, pstrain, 12 stress slots
void write_plastic_strain(const RankState& s, const std::string& dir) {
  if (s.in->output_plastic != 1) return;
  FILE* f = nullptr;
  for (int i = 0; i < s.totalNumOfElements; ++i) {
    const int* c = &s.nodeElemIdRelation[8 * (size_t)i];
    const double* x1 = &s.meshCoor[3 * (size_t)(c[0] - 1)];
    if (!(s.pstrain[i] > 1.0e-4 && std::fabs(x1[0]) < 5.0e3 && std::fabs(x1[1]) < 2.0e3 && std::fabs(x1[2]) < 8.0e3)) continue;
    if (!f) f = open_rank_file(s, dir, "pstr.txt", "w");   // the reference creates the file at the first hit
    double sc[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < 8; ++j)
      for (int k = 0; k < 3; ++k) sc[k] = sc[k] + s.meshCoor[k + 3 * (size_t)(c[j] - 1)];
    std::string line = " ";
    for (int k = 0; k < 3; ++k) line += fortran_e(sc[k] / 8.0, 18, 7, 4);
    line += fortran_e(s.pstrain[i], 18, 7, 4);
    for (int j = 1; j <= 12; ++j) line += fortran_e(s.stressArr[(size_t)s.stressCompIndexArr[i] + j - 1], 18, 7, 4);
    std::fprintf(f, "%s\n", line.c_str());
  }
  if (f) std::fclose(f);
}

// output_timeanalysis, library_output.f90:208-218: (1x,10e18.7e4,2i10)
void write_comp_time(const RankState& s, const std::string& dir) {
  FILE* f = open_rank_file(s, dir, "compTime", "w");
  std::string line = " ";
  for (int k = 0; k < 10; ++k) line += fortran_e(s.compTime[k], 18, 7, 4);
  std::fprintf(f, "%s%10d%10d\n", line.c_str(), s.totalNumOfElements, s.totalNumOfEquations);
  std::fclose(f);
}

}  // namespace eqh
