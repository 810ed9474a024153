// Writes the four ASCII tables with host/mcbrat_host.hpp::writeResults_ASCII from deterministic values;
// tests/test_write_results.py builds the same values and compares with the Python writer byte for byte.
#include <cstdio>
#include <string>

#include "../../host/mcbrat_host.hpp"

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  const int nx = 3, ny = 2, nz = 4, nd = 3;
  std::vector<double> x(nx + 1), y(ny + 1), z = {0.2, 0.45, 0.7, 1.7, 9.95};
  for (int i = 0; i <= nx; ++i) x[i] = 0.0625 * i;
  for (int j = 0; j <= ny; ++j) y[j] = 0.5 * j + 100.0;
  const size_t cols = (size_t)nx * ny;
  const size_t total = 2 * (3 + 3 * cols + nz + cols * nz + nd * cols);
  std::vector<double> v(total);
  long long s = 12345;
  for (size_t i = 0; i < total; ++i) {
    s = (s * 1103515245LL + 12345LL) % (1LL << 31);
    v[i] = ((double)s / (double)(1LL << 31) - 0.3) * std::pow(10.0, (double)((int)(i % 7) - 3));
  }
  size_t o = 0;
  auto take = [&](size_t n) { std::vector<double> a(v.begin() + o, v.begin() + o + n); o += n; return a; };
  mcbrat::Statistics st;
  const std::vector<double> mf = take(6);                       // (2, 3): means then errors
  for (int q = 0; q < 6; ++q) st.meanFlux[q] = mf[q];
  st.fluxUp = take(2 * cols); st.fluxDown = take(2 * cols); st.fluxAbsorbed = take(2 * cols);
  st.absorbedProfile = take(2 * nz); st.absorbedVolume = take(2 * cols * nz); st.radiance = take(2 * nd * cols);
  st.fluxUp[0] = 0.00005; st.fluxUp[cols] = -0.00001; st.fluxDown[0] = 99999.99996; st.fluxDown[cols] = 2.5e-5;
  mcbrat::RadianceOptions ro;
  ro.useRussianRouletteForIntensity = false; ro.zetaMin = 0.15f; ro.limitIntensityContributions = true;
  ro.maxIntensityContribution = 3.402823466e38f;
  mcbrat::Status status;
  mcbrat::writeResults_ASCII("a_rather_long_domain_file_name_that_exceeds_sixty_characters_by_a_good_margin.dom", 12345678901LL, 17,
                             false, true, true, 3.25f, 1367.0, 0.8660254f, 275.5f, 0.05, x, y, z, dir + "/flux.out", st,
                             dir + "/absprof.out", dir + "/absvol.out", dir + "/rad.out", {1.0f, -0.5f, 0.25f},
                             {0.0f, 90.0f, 359.99f}, ro, status);
  if (status.stateIsFailure()) { std::fprintf(stderr, "%s\n", status.message.c_str()); return 1; }
  return 0;
}
