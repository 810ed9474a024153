// mcbrat_host.hpp -- C++17 host-side mirror of the reference's module API for the photon path, written directly
// above the C ABI (include/mcbrat_cuda.h).  The reference's host is compiled code (Fortran 95); no Fortran compiler
// exists in this image, so this header is the compiled-language rendering of what the ISO_C_BINDING shim
// (fortran/mcbrat_cuda_mod.f90) does inside the reference's procedures -- same names, argument meaning and error
// texts: new_Domain / addOpticalComponent / getOpticalPropertiesByComponent (OPT:500-1072), new_PhaseFunction /
// new_PhaseFunctionTable (SPF:101-300), computeLobattoTerms / computeLegendrePolynomials (NUM:27-205),
// new_Integrator / specifyParameters / computeRadiativeTransfer / reportResults (INT:121-123),
// new_RandomNumberSequence, new_PhotonStream, and the driver's batch statistics (DRV:1023-1052, 1188-1228).
// Staging producers run on the host only as far as they are O(table); everything O(cells) or O(photons) is a
// C-ABI call.  There is no CPU fallback: without the CUDA library / a GPU every call fails with a Status.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/mcbrat_cuda.h"

namespace mcbrat {

// ---- ErrorMessages.f95: setStateToFailure / stateIsFailure, reduced to what the photon path reports ----
struct Status {
  bool failure = false;
  std::string message;
  void setStateToFailure(const std::string &m) { failure = true; message = m; }
  void setStateToSuccess() { failure = false; message.clear(); }
  bool stateIsFailure() const { return failure; }
};

// ---- RandomNumbersForMC: the host object carries the Philox key and the next unused photon id ----
struct randomNumberSequence { uint64_t seed = 0, nextPhotonId = 0; };
inline uint64_t mix64(uint64_t z) {                         // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline randomNumberSequence new_RandomNumberSequence(const std::vector<int64_t> &seed) {   // DRV:901 (/ iseed, thisProc, thisThread /)
  uint64_t key = 0;
  for (int64_t s : seed) key = mix64(key ^ (uint64_t)s);
  return randomNumberSequence{key, 0};
}

// ---- numericUtilities.f95 ----
inline float spacing32(float x) {
  x = std::fabs(x);
  if (x == 0.0f) return 1.17549435e-38f;
  const float s = std::nextafter(x, INFINITY) - x;
  return s < 1.17549435e-38f ? 1.17549435e-38f : s;
}
// P_0..P_maxL at mus, upward recursion in single precision (NUM:187-205); P[l][i]
inline std::vector<std::vector<float>> computeLegendrePolynomials(int maxL, const std::vector<float> &mus) {
  const int L = std::max(maxL, 1);
  std::vector<std::vector<float>> P(L + 1, std::vector<float>(mus.size()));
  for (size_t i = 0; i < mus.size(); ++i) { P[0][i] = 1.0f; P[1][i] = mus[i]; }
  for (int l = 1; l < maxL; ++l)
    for (size_t i = 0; i < mus.size(); ++i)
      P[l + 1][i] = (((float)(2 * l + 1) * mus[i]) * P[l][i] - (float)l * P[l - 1][i]) / (float)(l + 1);
  P.resize(maxL + 1 > 2 ? maxL + 1 : P.size());
  return P;
}
// Lobatto abscissas on [-1, 1] by Newton iteration (NUM:27-114); weights are not needed by the inversion
inline std::vector<float> computeLobattoMus(int n) {
  const float relativeAccuracy = 3.0f;
  const int maxIterations = 25;
  const float pi = (float)std::acos(-1.0);
  const int nTerms = n, midPoint = (nTerms + 1) / 2, m = midPoint - 1;
  std::vector<float> mus(n, 0.0f), trial(m), last(m), d1(m), d2(m);
  const float c1 = (nTerms % 2 == 1) ? 1.0f : 0.5f;
  for (int i = 1; i <= m; ++i)
    trial[i - 1] = (float)std::sin((double)(pi * ((float)i - c1) / (float)(nTerms - 1.0 + 0.5)));
  auto newton = [&](const std::vector<float> &t, std::vector<float> &a, std::vector<float> &b) {
    auto P = computeLegendrePolynomials(nTerms - 1, t);
    for (size_t i = 0; i < t.size(); ++i) {
      a[i] = (float)(nTerms - 1) * (t[i] * P[nTerms - 1][i] - P[nTerms - 2][i]) / (t[i] * t[i] - 1.0f);
      b[i] = (2.0f * t[i] * a[i] - (float)(nTerms * (nTerms - 1)) * P[nTerms - 1][i]) / (1.0f - t[i] * t[i]);
    }
  };
  if (m > 0) {
    newton(trial, d1, d2);
    last = trial;
    for (int i = 0; i < m; ++i) trial[i] = trial[i] - d1[i] / d2[i];
    int it = 0;
    for (;;) {
      std::vector<char> moving(m);
      bool any = false;
      for (int i = 0; i < m; ++i) { moving[i] = std::fabs(trial[i] - last[i]) > relativeAccuracy * spacing32(trial[i]); any = any || moving[i]; }
      if (!any) break;
      std::vector<float> a(m), b(m);
      newton(trial, a, b);
      for (int i = 0; i < m; ++i)
        if (moving[i]) { last[i] = trial[i]; trial[i] = trial[i] - a[i] / b[i]; }
      if (++it > maxIterations) break;
    }
  }
  mus[0] = -1.0f;
  for (int i = 0; i < m; ++i) mus[1 + i] = -trial[m - 1 - i];                 // mus(midPoint:2:-1) = -trialMus(:)
  if (nTerms % 2 == 0) for (int i = 0; i < midPoint; ++i) mus[midPoint + i] = -mus[midPoint - 1 - i];
  else for (int i = 0; i < midPoint; ++i) mus[midPoint - 1 + i] = -mus[midPoint - 1 - i];
  return mus;
}

// ---- scatteringPhaseFunctions.f95: Legendre-moment phase functions (chi_1.., P0 implied) ----
struct phaseFunction { std::vector<float> legendreCoefficients; std::string description; };
struct phaseFunctionTable { std::vector<phaseFunction> phaseFunctions; std::vector<float> key; std::string description; };
inline phaseFunction new_PhaseFunction(const std::vector<float> &legendreCoefficients, Status &status, const std::string &d = "") {
  if (legendreCoefficients.size() > 1 && (legendreCoefficients[0] > 1.0f || legendreCoefficients[0] < -1.0f))
    status.setStateToFailure("newPhaseFunction: Asymmetery parameter out of bounds.");
  return phaseFunction{legendreCoefficients, d};
}
inline phaseFunctionTable new_PhaseFunctionTable(const std::vector<phaseFunction> &pfs, const std::vector<float> &key, Status &status,
                                                 const std::string &d = "") {
  if (key.size() != pfs.size()) status.setStateToFailure("newPhaseFunctionTable: Number of phase functions and key values must match.");
  for (size_t i = 1; i < key.size(); ++i)
    if (!(key[i] > key[i - 1])) status.setStateToFailure("newPhaseFunctionTable: Key values must be unique, increasing.");
  return phaseFunctionTable{pfs, key, d};
}
// Henyey-Greenstein as the reference's generators build it: chi_l = g**l (Domain-Files/i3rcStepCloud.f95:55)
inline phaseFunction henyeyGreenstein(float g, int nLegendreCoefficients, Status &status) {
  std::vector<float> c(nLegendreCoefficients);
  for (int l = 1; l <= nLegendreCoefficients; ++l) c[l - 1] = std::pow(g, (float)l);
  return new_PhaseFunction(c, status, "Henyey-Greenstein");
}
// getPhaseFunctionValues SPF:448-498 at the given cosines: sum (2l+1) chi_l P_l in single precision
inline std::vector<float> getPhaseFunctionValuesAtMus(const phaseFunction &pf, const std::vector<float> &mus) {
  const int maxL = (int)pf.legendreCoefficients.size();
  std::vector<float> value(mus.size(), 0.0f);
  if (maxL == 0) { std::fill(value.begin(), value.end(), 0.5f); return value; }        // SPF:486-491 (quirk q14)
  auto P = computeLegendrePolynomials(maxL, mus);
  for (int l = 0; l <= maxL; ++l) {
    const float coef = (l == 0 ? 1.0f : pf.legendreCoefficients[l - 1]) * (float)(2 * l + 1);
    for (size_t i = 0; i < mus.size(); ++i) value[i] = value[i] + coef * P[l][i];
  }
  return value;
}

// ---- opticalProperties.f95: the domain ----
struct opticalComponent {
  std::string name; int zLevelBase; bool horizontallyUniform;
  std::vector<double> extinction, singleScatteringAlbedo; std::vector<int32_t> phaseFunctionIndex; int nLevels;
  phaseFunctionTable table;
};
struct Domain {
  std::vector<double> xPosition, yPosition, zPosition;
  double surfaceAlbedo = 0.0;
  std::vector<opticalComponent> components;
  // dense arrays, Fortran layout (x fastest), built by getOpticalPropertiesByComponent
  std::vector<double> totalExt, cumulativeExt, ssa; std::vector<int32_t> phaseFunctionIndex;
  int numX() const { return (int)xPosition.size() - 1; }
  int numY() const { return (int)yPosition.size() - 1; }
  int numZ() const { return (int)zPosition.size() - 1; }
};
inline Domain new_Domain(const std::vector<double> &x, const std::vector<double> &y, const std::vector<double> &z, double albedo,
                         Status &status) {                                         // OPT:500-555
  for (const auto *p : {&x, &y, &z})
    for (size_t i = 1; i < p->size(); ++i)
      if (!((*p)[i] > (*p)[i - 1])) status.setStateToFailure("new_Domain: Positions must be increasing, unique.");
  Domain d; d.xPosition = x; d.yPosition = y; d.zPosition = z; d.surfaceAlbedo = albedo;
  return d;
}
// 3-D component (nx,ny,nLevels) or horizontally uniform profile (nLevels), OPT:557-665
inline void addOpticalComponent(Domain &d, const std::string &name, const std::vector<double> &extinction,
                                const std::vector<double> &ssa, const std::vector<int32_t> &idx, const phaseFunctionTable &table,
                                int zLevelBase, Status &status) {
  const size_t cols = (size_t)d.numX() * d.numY();
  const bool uniform = extinction.size() % cols != 0 || extinction.size() < cols;
  const int nLev = (int)(uniform ? extinction.size() : extinction.size() / cols);
  if (ssa.size() != extinction.size() || idx.size() != extinction.size())
    return status.setStateToFailure("addOpticalComponent: optical property grids must be the same size.");
  if (zLevelBase < 1 || zLevelBase + nLev - 1 > d.numZ())
    return status.setStateToFailure("addOpticalComponent: arrays don't fit the vertical extent of the domain.");
  for (size_t i = 0; i < extinction.size(); ++i) {
    if (extinction[i] < 0) return status.setStateToFailure("addOpticalComponent: extinction must be >= 0.");
    if (ssa[i] < 0 || ssa[i] > 1) return status.setStateToFailure("addOpticalComponent: singleScatteringAlbedo must be between 0 and 1");
    if (idx[i] < 0 || idx[i] > (int)table.phaseFunctions.size())
      return status.setStateToFailure("addOpticalComponent: phase function index is out of bounds");
  }
  d.components.push_back(opticalComponent{name, zLevelBase, uniform, extinction, ssa, idx, nLev, table});
  d.totalExt.clear();
}
inline void getOpticalPropertiesByComponent(Domain &d, Status &status) {            // OPT:966-1072
  const int nc = (int)d.components.size();
  if (nc == 0) return status.setStateToFailure("getOpticalPropertiesByComponent: domain contains no optical components.");
  const size_t cols = (size_t)d.numX() * d.numY(), cells = cols * d.numZ();
  d.cumulativeExt.assign(cells * nc, 0.0); d.ssa.assign(cells * nc, 0.0); d.phaseFunctionIndex.assign(cells * nc, 0);
  for (int c = 0; c < nc; ++c) {
    const opticalComponent &q = d.components[c];
    for (int k = 0; k < q.nLevels; ++k) {
      const size_t lev = (size_t)(q.zLevelBase - 1 + k);
      for (size_t j = 0; j < cols; ++j) {
        const size_t src = q.horizontallyUniform ? (size_t)k : j + cols * k, dst = j + cols * lev + cells * c;
        d.cumulativeExt[dst] = q.extinction[src]; d.ssa[dst] = q.singleScatteringAlbedo[src];
        d.phaseFunctionIndex[dst] = q.phaseFunctionIndex[src];
      }
    }
  }
  for (int c = 1; c < nc; ++c)                                                     // OPT:1055-1057
    for (size_t i = 0; i < cells; ++i) d.cumulativeExt[i + cells * c] += d.cumulativeExt[i + cells * (c - 1)];
  d.totalExt.assign(d.cumulativeExt.begin() + cells * (nc - 1), d.cumulativeExt.begin() + cells * nc);
  for (int c = 0; c < nc; ++c)                                                     // OPT:1059-1061
    for (size_t i = 0; i < cells; ++i)
      if (d.totalExt[i] > 2.2250738585072014e-308) d.cumulativeExt[i + cells * c] /= d.totalExt[i];
}

// ---- monteCarloIllumination.f95: the stream is a description, photons are born on the device ----
struct photonStream { double solarMu = 1.0, solarAzimuth = 0.0; int64_t numberOfPhotons = 0, currentPhoton = 0; uint64_t firstPhotonId = 0; };
inline photonStream new_PhotonStream(float solarMu, float solarAzimuth, int64_t numberOfPhotons, randomNumberSequence &randoms,
                                     Status &status) {                              // ILL:62-101
  if (numberOfPhotons < 0) status.setStateToFailure("setIllumination: must ask for non-negative number of photons.");
  if (solarAzimuth < 0.0f || solarAzimuth > 360.0f) status.setStateToFailure("setIllumination: solarAzimuth out of bounds");
  if (std::fabs(solarMu) > 1.0f || std::fabs(solarMu) <= 1.17549435e-38f) status.setStateToFailure("setIllumination: solarMu out of bounds");
  photonStream p; p.solarMu = solarMu; p.solarAzimuth = solarAzimuth; p.numberOfPhotons = numberOfPhotons; p.currentPhoton = 1;
  p.firstPhotonId = randoms.nextPhotonId; randoms.nextPhotonId += (uint64_t)numberOfPhotons;
  return p;
}

// ---- monteCarloRadiativeTransfer.f95: type(integrator) owns one mcb_handle ----
struct integrator {
  mcb_handle *gpu = nullptr;
  mcb_options options;
  int numX = 0, numY = 0, numZ = 0, numComps = 0, numDirections = 0;
  int minInverseTableSize = 9001, minForwardTableSize = 9001;                       // INT:24-25
  const Domain *stagedDomain = nullptr;
  std::string lastMessage(const char *where) const {
    char buf[512] = "";
    if (gpu) mcb_last_error(gpu, buf, sizeof(buf));
    return std::string(where) + ": " + (buf[0] ? buf : "error");
  }
};
inline integrator new_Integrator(const Domain &atmosphere, Status &status, int device = 0) {      // INT:129-201
  integrator g;
  mcb_default_options(&g.options);
  const int rc = mcb_create(device, &g.gpu);
  if (rc != 0) { status.setStateToFailure("new_Integrator: mcb_create failed (no CUDA device? there is no CPU fallback)"); return g; }
  g.numX = atmosphere.numX(); g.numY = atmosphere.numY(); g.numZ = atmosphere.numZ();
  if (mcb_set_grid(g.gpu, g.numX, g.numY, g.numZ, atmosphere.xPosition.data(), atmosphere.yPosition.data(), atmosphere.zPosition.data()))
    status.setStateToFailure(g.lastMessage("new_Integrator"));
  return g;
}
inline void finalize_Integrator(integrator &g) { if (g.gpu) mcb_destroy(g.gpu); g.gpu = nullptr; }   // INT:1486
// specifyParameters INT:1046-1337: the options are fields of g.options; directions as (mu, phi in degrees)
inline void specifyParameters(integrator &g, Status &status, const std::vector<float> &intensityMus = {},
                              const std::vector<float> &intensityPhis = {}) {
  if (intensityMus.size() != intensityPhis.size())
    return status.setStateToFailure("specifyParameters: intensityMus, intensityPhis must be the same length.");
  std::vector<float> dirs;
  const float Pi = 3.14159265358979312f;
  for (size_t i = 0; i < intensityMus.size(); ++i) {                                 // INT:1245-1271, makeDirectionCosines INT:1876-1894
    const float mu = intensityMus[i], phi = intensityPhis[i] * Pi / 180.0f;
    if (std::fabs(mu) < 1.17549435e-38f) return status.setStateToFailure("specifyParameters: intensityMus can't be 0 (directly sideways)");
    const float s = std::sqrt(1.0f - mu * mu);
    dirs.push_back(s * (float)std::cos((double)phi)); dirs.push_back(s * (float)std::sin((double)phi)); dirs.push_back(mu);
  }
  g.numDirections = (int)intensityMus.size();
  if (mcb_set_views(g.gpu, g.numDirections, dirs.empty() ? nullptr : dirs.data()) || mcb_set_options(g.gpu, &g.options))
    status.setStateToFailure(g.lastMessage("specifyParameters"));
}
// the per-batch getInfo_Domain copies of INT:434-443 become one staging per domain; the tables of INT:280-285 are
// built in HBM
inline void stageDomain(integrator &g, Domain &d, Status &status) {
  if (g.stagedDomain == &d) return;
  if (d.totalExt.empty()) getOpticalPropertiesByComponent(d, status);
  if (status.stateIsFailure()) return;
  const int nc = (int)d.components.size();
  if (mcb_set_optics(g.gpu, nc, d.totalExt.data(), d.cumulativeExt.data(), d.ssa.data(), d.phaseFunctionIndex.data(), d.surfaceAlbedo))
    return status.setStateToFailure(g.lastMessage("computeRadiativeTransfer"));
  // Every table is built in HBM from the Legendre moments alone (round 2): the Lobatto abscissas, the phase function
  // there and the inversion (INV:66-174: mcb_build_inverse_table_legendre), and the equal-angle forward tables of
  // tabulateForwardPhaseFunctions OPT:1872-1934 (mcb_build_forward_table_general; no hybrid peak asked for here).
  // computeLobattoMus / getPhaseFunctionValuesAtMus above remain as the host-side statement of the same arithmetic.
  for (int c = 0; c < nc; ++c) {
    std::vector<int32_t> nCoef; std::vector<float> coefs;
    for (const phaseFunction &pf : d.components[c].table.phaseFunctions) {
      nCoef.push_back((int32_t)pf.legendreCoefficients.size());
      coefs.insert(coefs.end(), pf.legendreCoefficients.begin(), pf.legendreCoefficients.end());
    }
    if (mcb_build_inverse_table_legendre(g.gpu, c + 1, g.minInverseTableSize, (int)nCoef.size(), nCoef.data(),
                                         coefs.empty() ? nullptr : coefs.data()))
      return status.setStateToFailure(g.lastMessage("tabulateInversePhaseFunctions"));
    if (g.numDirections > 0 &&
        mcb_build_forward_table_general(g.gpu, c + 1, g.minForwardTableSize, (int)nCoef.size(), nCoef.data(),
                                        coefs.empty() ? nullptr : coefs.data(), nullptr, nullptr, nullptr, 0.0f))
      return status.setStateToFailure(g.lastMessage("tabulateForwardPhaseFunctions"));
  }
  g.numComps = nc; g.stagedDomain = &d;
}
inline void computeRadiativeTransfer(integrator &g, Domain &thisDomain, randomNumberSequence &randomNumbers, photonStream &incomingPhotons,
                                     int64_t numPhotonsPerBatch, int64_t &numPhotonsProcessed, Status &status) {   // INT:209-218
  numPhotonsProcessed = 0;
  stageDomain(g, thisDomain, status);
  if (status.stateIsFailure()) return;
  if (mcb_set_solar_source(g.gpu, (float)incomingPhotons.solarMu, (float)incomingPhotons.solarAzimuth))
    return status.setStateToFailure(g.lastMessage("new_PhotonStream"));
  const int64_t left = incomingPhotons.numberOfPhotons - (incomingPhotons.currentPhoton - 1);
  if (incomingPhotons.currentPhoton < 1 || left <= 0) return status.setStateToFailure("computeRadiativeTransfer: Didn't process any photons.");
  const int64_t n = std::min(numPhotonsPerBatch, left);
  if (mcb_run_batch(g.gpu, n, randomNumbers.seed, incomingPhotons.firstPhotonId + (uint64_t)(incomingPhotons.currentPhoton - 1), &numPhotonsProcessed))
    return status.setStateToFailure(g.lastMessage("computeRadiativeTransfer"));
  incomingPhotons.currentPhoton += n;
  status.setStateToSuccess();
}
// reportResults INT:845-1042: any output may be omitted (nullptr); means are sum()/numColumns in single precision
struct Results { std::vector<float> fluxUp, fluxDown, fluxAbsorbed, volumeAbsorption, intensity; float meanFluxUp = 0, meanFluxDown = 0, meanFluxAbsorbed = 0; };
inline void reportResults(integrator &g, Results &r, Status &status, bool volume = false, bool radiance = false) {
  const size_t cols = (size_t)g.numX * g.numY;
  r.fluxUp.resize(cols); r.fluxDown.resize(cols); r.fluxAbsorbed.resize(cols);
  if (volume) r.volumeAbsorption.resize(cols * g.numZ);
  if (radiance) r.intensity.resize(cols * g.numDirections);
  if (mcb_get_results(g.gpu, 0, r.fluxUp.data(), r.fluxDown.data(), r.fluxAbsorbed.data(), volume ? r.volumeAbsorption.data() : nullptr,
                      radiance ? r.intensity.data() : nullptr, nullptr))
    return status.setStateToFailure(g.lastMessage("reportResults"));
  auto mean = [&](const std::vector<float> &a) { float s = 0.0f; for (float v : a) s += v; return s / (float)cols; };   // INT:881-884
  r.meanFluxUp = mean(r.fluxUp); r.meanFluxDown = mean(r.fluxDown); r.meanFluxAbsorbed = mean(r.fluxAbsorbed);
}

// ---- the driver's batch loop and statistics on the device (DRV:949-1052, 1188-1228) ----
// every array is stats(..., 1:2) in the Fortran layout: the block of means, then the block of standard errors;
// meanFlux = (up, down, absorbed | their standard errors)
struct Statistics {
  double meanFlux[6] = {0, 0, 0, 0, 0, 0};
  std::vector<double> fluxUp, fluxDown, fluxAbsorbed, absorbedProfile, absorbedVolume, radiance;
  int64_t totalNumPhotons = 0, batchesCompleted = 0;
};
inline void runBatches(integrator &g, Domain &d, randomNumberSequence &randoms, float solarMu, float solarAzimuth, int64_t numBatches,
                       int64_t numPhotonsPerBatch, Status &status) {
  stageDomain(g, d, status);
  if (status.stateIsFailure()) return;
  int64_t done = 0;
  if (mcb_set_solar_source(g.gpu, solarMu, solarAzimuth) ||
      mcb_run_batches(g.gpu, numBatches, numPhotonsPerBatch, randoms.seed, randoms.nextPhotonId, &done))
    return status.setStateToFailure(g.lastMessage("computeRadiativeTransfer"));
  randoms.nextPhotonId += (uint64_t)done;
}
inline void reportStatistics(integrator &g, double solarFlux, Statistics &s, Status &status, bool pixels = false, bool volume = false) {
  const size_t cols = (size_t)g.numX * g.numY;
  s.absorbedProfile.assign(2 * (size_t)g.numZ, 0.0);
  s.radiance.assign(2 * cols * g.numDirections, 0.0);
  if (pixels) { s.fluxUp.assign(2 * cols, 0.0); s.fluxDown.assign(2 * cols, 0.0); s.fluxAbsorbed.assign(2 * cols, 0.0); }
  if (volume) s.absorbedVolume.assign(2 * cols * g.numZ, 0.0);
  if (mcb_get_statistics(g.gpu, solarFlux, s.meanFlux, pixels ? s.fluxUp.data() : nullptr, pixels ? s.fluxDown.data() : nullptr,
                         pixels ? s.fluxAbsorbed.data() : nullptr, s.absorbedProfile.data(), volume ? s.absorbedVolume.data() : nullptr,
                         g.numDirections ? s.radiance.data() : nullptr, &s.totalNumPhotons, &s.batchesCompleted))
    status.setStateToFailure(g.lastMessage("reportStatistics"));
}

// ---- multipleProcesses_mpi.f95 over NCCL (MPIW:29-251): one process per GPU ----
// The 128-byte id is created by rank 0 (mcb_comm_unique_id) and carried to the other ranks by the host's own means --
// MPI_BCAST in the Fortran host; the example driver uses a file.
inline void initializeProcesses(integrator &g, int numProcs, int thisProc, const void *id128, Status &status) {   // MPIW:29-52
  if (mcb_comm_init(g.gpu, numProcs, thisProc, id128)) status.setStateToFailure(g.lastMessage("initializeProcesses"));
}
inline void sumAcrossProcesses_tallies(integrator &g, Status &status, int root = 0) {                           // DRV:1151-1166
  if (mcb_reduce_tallies(g.gpu, root) || mcb_synchronize(g.gpu)) status.setStateToFailure(g.lastMessage("sumAcrossProcesses"));
}
inline void sumAcrossProcesses_statistics(integrator &g, Status &status, int root = 0) {
  if (mcb_reduce_statistics(g.gpu, root) || mcb_synchronize(g.gpu)) status.setStateToFailure(g.lastMessage("sumAcrossProcesses"));
}
inline void finalizeProcesses(integrator &g) { mcb_comm_destroy(g.gpu); }                                       // MPIW:62-68

// ---- writeResults_ASCII (DRV:1324-1495): the driver's four ASCII tables, format-exact ----
namespace fmt {
// Fw.d: right-justified, the optional leading zero dropped when the field is one character short, w asterisks on overflow
inline std::string F(double v, int w, int d) {
  char buf[512];
  std::snprintf(buf, sizeof(buf), "%.*f", d, v);           // glibc rounds the exact binary value, ties to even, like gfortran
  std::string t(buf);
  if ((int)t.size() > w) {
    if (t.compare(0, 2, "0.") == 0) t.erase(0, 1);
    else if (t.compare(0, 3, "-0.") == 0) t.erase(1, 1);
  }
  if ((int)t.size() > w) return std::string((size_t)w, '*');
  return std::string((size_t)w - t.size(), ' ') + t;
}
// Ew.d: 0.ddddddE+ee
inline std::string E(double v, int w, int d) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%.*e", d - 1, std::fabs(v));        // D.ddddde+XX with d significant digits
  std::string t(buf);
  const size_t e = t.find('e');
  std::string digits = t.substr(0, 1) + t.substr(2, e - 2);
  int ex = std::atoi(t.c_str() + e + 1) + 1;
  if (v == 0.0) ex = 0;
  char es[16];
  if (std::abs(ex) < 100) std::snprintf(es, sizeof(es), "E%+03d", ex); else std::snprintf(es, sizeof(es), "%+04d", ex);
  std::string r = std::string(v < 0 ? "-" : "") + "0." + digits + es;
  if ((int)r.size() > w) {
    if (r.compare(0, 2, "0.") == 0) r.erase(0, 1);
    else if (r.compare(0, 3, "-0.") == 0) r.erase(1, 1);
  }
  if ((int)r.size() > w) return std::string((size_t)w, '*');
  return std::string((size_t)w - r.size(), ' ') + r;
}
inline std::string I(long long v, int w) {
  const std::string t = std::to_string(v);
  return (int)t.size() > w ? std::string((size_t)w, '*') : std::string((size_t)w - t.size(), ' ') + t;
}
// A60 of the driver's character(len=256) variable: the leftmost 60 characters of the blank-padded name
inline std::string A60(const std::string &name) { std::string t = name; t.resize(256, ' '); return t.substr(0, 60); }
inline const char *L(bool b) { return b ? "T" : "F"; }
inline std::string pair(const std::vector<double> &stats, size_t i, size_t n) {      // 2(1X,F9.4) of stats(i, 1:2)
  return " " + F(stats[i], 9, 4) + " " + F(stats[n + i], 9, 4);
}
}  // namespace fmt

struct RadianceOptions {                  // module variables of the driver the radiance header prints (DRV:78-82)
  bool useRussianRouletteForIntensity = true; float zetaMin = 0.3f;
  bool limitIntensityContributions = false; float maxIntensityContribution = 77.0f;
};

inline void writeResults_ASCII(const std::string &domainFileName, int64_t totalNumPhotons, int /*numBatches*/, bool useRayTracing,
                               bool useRussianRoulette, bool useHybridPhaseFunsForIntenCalcs, float hybridPhaseFunWidth,
                               double solarFlux, float solarMu, float solarAzimuth, double surfaceAlbedo,
                               const std::vector<double> &xPosition, const std::vector<double> &yPosition,
                               const std::vector<double> &zPosition, const std::string &outputFluxFile, const Statistics &s,
                               const std::string &outputAbsProfFile, const std::string &outputAbsVolumeFile,
                               const std::string &outputRadFile, const std::vector<float> &intensityMus,
                               const std::vector<float> &intensityPhis, const RadianceOptions &ro, Status &status) {
  using namespace fmt;
  const int nx = (int)xPosition.size() - 1, ny = (int)yPosition.size() - 1, nz = (int)zPosition.size() - 1;
  const size_t cols = (size_t)nx * ny;
  auto trimmed = [](const std::string &f) { return f.find_first_not_of(' ') != std::string::npos; };    // len_trim(f) > 0
  auto xc = [&](int i) { return (xPosition[i] + xPosition[i + 1]) / 2.0; };
  auto yc = [&](int j) { return (yPosition[j] + yPosition[j + 1]) / 2.0; };
  auto header = [&](FILE *f, const char *kind, bool radiance) {
    std::fprintf(f, "!   I3RC Monte Carlo 3D Solar Radiative Transfer: %s\n", kind);
    std::fprintf(f, "!  Property_File=%s\n", A60(domainFileName).c_str());
    std::fprintf(f, "!  Num_Photons=%s\n", I(totalNumPhotons, 10).c_str());
    std::fprintf(f, "!  PhotonTracing=%s    Russian_Roulette=%s\n", L(useRayTracing), L(useRussianRoulette));
    std::fprintf(f, "!  Hybrid_Phase_Func_for_Radiance=%s   Gaussian_Phase_Func_Width_deg=%s\n", L(useHybridPhaseFunsForIntenCalcs),
                 F((double)hybridPhaseFunWidth, 5, 2).c_str());
    if (radiance) {
      std::fprintf(f, "!  Intensity_uses_Russian_Roulette=%s   Intensity_Russian_Roulette_zeta_min=%s\n",
                   L(ro.useRussianRouletteForIntensity), F((double)ro.zetaMin, 5, 2).c_str());
      std::fprintf(f, "!  limited_intensity_contributions=%s   max_intensity_contribution=%s\n",
                   L(ro.limitIntensityContributions), F((double)ro.maxIntensityContribution, 5, 2).c_str());
    }
    std::fprintf(f, "!  Solar_Flux=%s   Solar_Mu=%s   Solar_Phi=%s\n", E(solarFlux, 13, 6).c_str(), F((double)solarMu, 10, 7).c_str(),
                 F((double)solarAzimuth, 7, 3).c_str());
    std::fprintf(f, "!  Lambertian_Surface_Albedo=%s\n", F(surfaceAlbedo, 7, 4).c_str());
  };
  auto open = [&](const std::string &name) -> FILE * {
    const size_t a = name.find_first_not_of(' '), b = name.find_last_not_of(' ');
    FILE *f = std::fopen(name.substr(a, b - a + 1).c_str(), "w");
    if (!f) status.setStateToFailure("writeResults_ASCII: cannot open " + name);
    return f;
  };
  if (trimmed(outputFluxFile) && s.fluxUp.size() == 2 * cols) {                                      // DRV:1375-1403
    FILE *f = open(outputFluxFile); if (!f) return;
    header(f, "Flux", false);
    std::fprintf(f, "!  Output_Type= Pixel Flux\n");
    std::fprintf(f, "!  Upwelling_Level=%s   Downwelling_level=%s\n", F(zPosition[nz], 7, 3).c_str(), F(zPosition[0], 7, 3).c_str());
    std::fprintf(f, "!   X      Y           Flux_Up             Flux_Down            Flux_Absorbed \n");
    std::fprintf(f, "!                  Mean     StdErr       Mean     StdErr       Mean     StdErr\n");
    std::fprintf(f, "!  Average:   ");
    for (int q = 0; q < 3; ++q) std::fprintf(f, "  %s %s", F(s.meanFlux[q], 9, 4).c_str(), F(s.meanFlux[3 + q], 9, 4).c_str());
    std::fprintf(f, "\n");
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const size_t c = (size_t)i + (size_t)nx * j;
        std::fprintf(f, "%s%s %s %s %s\n", F(xc(i), 7, 3).c_str(), F(yc(j), 7, 3).c_str(), pair(s.fluxUp, c, cols).c_str(),
                     pair(s.fluxDown, c, cols).c_str(), pair(s.fluxAbsorbed, c, cols).c_str());
      }
    std::fclose(f);
  }
  if (trimmed(outputAbsProfFile) && s.absorbedProfile.size() == 2 * (size_t)nz) {                    // DRV:1410-1431
    FILE *f = open(outputAbsProfFile); if (!f) return;
    header(f, "Absorption Profile", false);
    std::fprintf(f, "!  Output_Type= Absorption Profile\n!   Z    Absorbed_Flux (flux/km) \n!          Mean     StdErr \n");
    for (int k = 0; k < nz; ++k)
      std::fprintf(f, "%s %s\n", F(0.5 * (zPosition[k] + zPosition[k + 1]), 7, 3).c_str(), pair(s.absorbedProfile, (size_t)k, (size_t)nz).c_str());
    std::fclose(f);
  }
  if (trimmed(outputAbsVolumeFile) && s.absorbedVolume.size() == 2 * cols * nz) {                    // DRV:1437-1463
    FILE *f = open(outputAbsVolumeFile); if (!f) return;
    header(f, "3D Absorption Field", false);
    std::fprintf(f, "!  Output_Type= Volume Absorption \n!    X       Y        Z       Absorbed_Flux (flux/km)\n"
                    "!                               Mean     StdErr \n");
    for (int i = 0; i < nx; ++i)
      for (int j = 0; j < ny; ++j)
        for (int k = 0; k < nz; ++k)
          std::fprintf(f, "%s %s %s %s\n", F(xc(i), 7, 3).c_str(), F(yc(j), 7, 3).c_str(),
                       F((zPosition[k] + zPosition[k + 1]) / 2.0, 7, 3).c_str(),
                       pair(s.absorbedVolume, (size_t)i + (size_t)nx * ((size_t)j + (size_t)ny * k), cols * nz).c_str());
    std::fclose(f);
  }
  if (trimmed(outputRadFile) && !s.radiance.empty()) {                                               // DRV:1468-1493
    FILE *f = open(outputRadFile); if (!f) return;
    int numRadDir = 0;
    for (float m : intensityMus) numRadDir += std::fabs(m) > 0.0f ? 1 : 0;
    const size_t n = s.radiance.size() / 2;
    header(f, "Radiance", true);
    std::fprintf(f, "!  Output_Type= Pixel Radiance\n");
    std::fprintf(f, "!  RADIANCE AT Z=%s   NXO=%s   NYO=%s   NDIR=%s\n", F(zPosition[nz], 7, 3).c_str(), I(nx, 4).c_str(), I(ny, 4).c_str(),
                 I(numRadDir, 4).c_str());
    std::fprintf(f, "!   X      Y         Radiance (Mean, StdErr)\n");
    for (int k = 0; k < numRadDir; ++k) {
      std::fprintf(f, "!  %s %s  <- (mu,phi)\n", F((double)intensityMus[k], 8, 5).c_str(), F((double)intensityPhis[k], 6, 2).c_str());
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
          const size_t c = (size_t)i + (size_t)nx * ((size_t)j + (size_t)ny * k);
          std::fprintf(f, "%s%s %s %s\n", F(xc(i), 7, 3).c_str(), F(yc(j), 7, 3).c_str(), F(s.radiance[c], 9, 4).c_str(),
                       F(s.radiance[n + c], 9, 4).c_str());
        }
    }
    std::fclose(f);
  }
}

}  // namespace mcbrat
