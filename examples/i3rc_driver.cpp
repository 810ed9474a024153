// i3rc_driver.cpp -- a compiled host for the photon path, in the shape of Drivers/monteCarloDriver.f95: build the
// domain (the generators Domain-Files/i3rcStepCloud.f95 and the homogeneous I3RC_mono_SWhomog slab), create the
// integrator (DRV:533), set its parameters (DRV:540-597), run numBatches batches (DRV:949-1052, here one C-ABI call
// with the moments kept on the device) and print the domain-mean results with their standard errors (DRV:1188-1228,
// 1324-1400).  Everything below the mcbrat:: calls is include/mcbrat_cuda.h; no Python, no torch.
//   usage: i3rc_driver <homog|stepcloud> [numBatches=32] [numPhotonsPerBatch=100000] [iseed=10] [views=0|1]
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../host/mcbrat_host.hpp"

using namespace mcbrat;

static void printStatus(const Status &status) {            // userInterface_Unix.f95:32-51
  if (status.stateIsFailure()) { std::fprintf(stderr, "%s\n", status.message.c_str()); std::exit(1); }
}

int main(int argc, char **argv) {
  const std::string deck = argc > 1 ? argv[1] : "homog";
  const int64_t numBatches = argc > 2 ? std::atoll(argv[2]) : 32, numPhotonsPerBatch = argc > 3 ? std::atoll(argv[3]) : 100000;
  const int64_t iseed = argc > 4 ? std::atoll(argv[4]) : 10;
  const bool views = argc > 5 && std::atoi(argv[5]) != 0;
  Status status;
  Domain domain;
  float solarMu = 0.5f, solarAzimuth = 0.0f;
  if (deck == "homog") {                                   // C1: 20^3 cells of 0.0625 km, tau = 10, ssa 0.99, HG g = 0.85 (64 terms), A = 0.2
    const int n = 20; const double delta = 0.0625;
    std::vector<double> e(n + 1);
    for (int i = 0; i <= n; ++i) e[i] = delta * i;
    domain = new_Domain(e, e, e, 0.2, status); printStatus(status);
    const size_t cells = (size_t)n * n * n;
    const phaseFunctionTable table = new_PhaseFunctionTable({henyeyGreenstein(0.85f, 64, status)}, {1.0f}, status);
    addOpticalComponent(domain, "cloud", std::vector<double>(cells, 10.0 / (n * delta)), std::vector<double>(cells, 0.99),
                        std::vector<int32_t>(cells, 1), table, 1, status);
  } else if (deck == "stepcloud") {                        // C2: Domain-Files/i3rcStepCloud.f95:27-84
    const int nColumns = 32, nLayers = 32;
    const float deltaX = 500.0f / nColumns, deltaZ = 250.0f / nLayers;
    std::vector<double> x(nColumns + 1), y = {0.0, 500.0}, z(nLayers + 1);
    for (int i = 0; i <= nColumns; ++i) x[i] = (double)(deltaX * (float)i);
    for (int i = 0; i <= nLayers; ++i) z[i] = (double)(deltaZ * (float)i);
    domain = new_Domain(x, y, z, 0.0, status); printStatus(status);
    std::vector<double> ext((size_t)nColumns * nLayers), ssa(ext.size(), 0.99);
    for (int k = 0; k < nLayers; ++k)
      for (int i = 0; i < nColumns; ++i) ext[i + (size_t)nColumns * k] = (double)((i < nColumns / 2 ? 2.0f : 18.0f) / 250.0f);
    const phaseFunctionTable table = new_PhaseFunctionTable({henyeyGreenstein(0.85f, 64, status)}, {1.0f}, status);
    addOpticalComponent(domain, "cloud", ext, ssa, std::vector<int32_t>(ext.size(), 1), table, 1, status);
  } else {
    std::fprintf(stderr, "unknown deck %s\n", deck.c_str()); return 2;
  }
  printStatus(status);
  getOpticalPropertiesByComponent(domain, status); printStatus(status);

  integrator mcIntegrator = new_Integrator(domain, status); printStatus(status);                          // DRV:533
  mcIntegrator.minInverseTableSize = mcIntegrator.minForwardTableSize = 10001;                            // nPhaseIntervals DRV:71
  if (views) {                                                                                            // the 5 I3RC view angles
    mcIntegrator.options.useRussianRouletteForIntensity = 1; mcIntegrator.options.zetaMin = 0.3f;           // DRV:78-79
    specifyParameters(mcIntegrator, status, {1.0f, 0.866f, 0.866f, 0.5f, 0.5f}, {0.0f, 0.0f, 180.0f, 0.0f, 180.0f});
  } else {
    specifyParameters(mcIntegrator, status);
  }
  printStatus(status);
  randomNumberSequence randoms = new_RandomNumberSequence({iseed, 1, 0});                                  // DRV:901

  // one batch through the module API as the reference's worker loop calls it (DRV:956-1011) ...
  photonStream incomingPhotons = new_PhotonStream(solarMu, solarAzimuth, numPhotonsPerBatch, randoms, status); printStatus(status);
  int64_t numPhotonsProcessed = 0;
  computeRadiativeTransfer(mcIntegrator, domain, randoms, incomingPhotons, numPhotonsPerBatch, numPhotonsProcessed, status); printStatus(status);
  Results one;
  reportResults(mcIntegrator, one, status); printStatus(status);
  std::printf("first batch: %lld photons  meanFluxUp %.6f meanFluxDown %.6f meanFluxAbsorbed %.6f\n", (long long)numPhotonsProcessed,
              one.meanFluxUp, one.meanFluxDown, one.meanFluxAbsorbed);
  // ... then the whole batch loop with the driver's moments kept on the device
  runBatches(mcIntegrator, domain, randoms, solarMu, solarAzimuth, numBatches, numPhotonsPerBatch, status); printStatus(status);
  Statistics stats;
  reportStatistics(mcIntegrator, 1.0, stats, status); printStatus(status);
  std::printf("batches %lld photons %lld\n", (long long)stats.batchesCompleted, (long long)stats.totalNumPhotons);
  std::printf("Flux Up    %.6f +- %.6f\nFlux Down  %.6f +- %.6f\nFlux Absorbed %.6f +- %.6f\n", stats.meanFlux[0], stats.meanFlux[3],
              stats.meanFlux[1], stats.meanFlux[4], stats.meanFlux[2], stats.meanFlux[5]);
  if (views) {
    const size_t cols = (size_t)mcIntegrator.numX * mcIntegrator.numY, n = cols * mcIntegrator.numDirections;
    for (int d = 0; d < mcIntegrator.numDirections; ++d) {
      double m = 0.0;
      for (size_t i = 0; i < cols; ++i) m += stats.radiance[i + cols * d];
      std::printf("Radiance view %d  %.6f\n", d + 1, m / cols);
    }
    (void)n;
  }
  finalize_Integrator(mcIntegrator);
  return 0;
}
