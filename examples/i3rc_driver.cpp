// i3rc_driver.cpp -- a compiled host for the photon path, in the shape of Drivers/monteCarloDriver.f95: build the
// domain (the generators Domain-Files/i3rcStepCloud.f95 and the homogeneous I3RC_mono_SWhomog slab), create the
// integrator (DRV:533), set its parameters (DRV:540-597), run numBatches batches (DRV:949-1052, here one C-ABI call
// with the moments kept on the device) and print the domain-mean results with their standard errors (DRV:1188-1228,
// 1324-1400).  Everything below the mcbrat:: calls is include/mcbrat_cuda.h; no Python, no torch.
//   usage: i3rc_driver <homog|stepcloud> [numBatches=32] [numPhotonsPerBatch=100000] [iseed=10] [views=0|1]
//                      [--ranks N] [--out PREFIX]
// --ranks N: one process per GPU (the reference: one MPI rank per core, DRV:441-444): the launcher starts N copies of
// itself, rank r on device r; the batches are dealt out in contiguous blocks of global photon ids, every rank keeps
// its moments on its device, and ONE ncclReduce (mcb_reduce_statistics, replacing the sumAcrossProcesses calls at
// DRV:1151-1166) brings them to rank 0, which reports.  The NCCL id travels through a file -- the job MPI_BCAST has
// in the Fortran host.  --out PREFIX writes the driver's ASCII tables (writeResults_ASCII, DRV:1324-1495).
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>

#include "../host/mcbrat_host.hpp"

using namespace mcbrat;

static void printStatus(const Status &status) {            // userInterface_Unix.f95:32-51
  if (status.stateIsFailure()) { std::fprintf(stderr, "%s\n", status.message.c_str()); std::exit(1); }
}

int main(int argc, char **argv) {
  // options first (they may appear anywhere), the positional arguments keep their order
  int numProcs = 1, thisProc = -1;
  std::string idFile, outPrefix;
  std::vector<std::string> pos;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--ranks" && i + 1 < argc) numProcs = std::atoi(argv[++i]);
    else if (a == "--rank" && i + 1 < argc) thisProc = std::atoi(argv[++i]);
    else if (a == "--id-file" && i + 1 < argc) idFile = argv[++i];
    else if (a == "--out" && i + 1 < argc) outPrefix = argv[++i];
    else pos.push_back(a);
  }
  const std::string deck = pos.size() > 0 ? pos[0] : "homog";
  const int64_t numBatches = pos.size() > 1 ? std::atoll(pos[1].c_str()) : 32, numPhotonsPerBatch = pos.size() > 2 ? std::atoll(pos[2].c_str()) : 100000;
  const int64_t iseed = pos.size() > 3 ? std::atoll(pos[3].c_str()) : 10;
  const bool views = pos.size() > 4 && std::atoi(pos[4].c_str()) != 0;
  if (numProcs > 1 && thisProc < 0) {                      // launcher: start one process per GPU and wait for them
    if (numBatches % numProcs != 0) { std::fprintf(stderr, "numBatches must be a multiple of --ranks\n"); return 2; }
    char tmpl[] = "/tmp/mcb_nccl_id_XXXXXX";
    const int fd = mkstemp(tmpl);
    if (fd < 0) { std::perror("mkstemp"); return 2; }
    close(fd); std::remove(tmpl);
    std::vector<pid_t> kids;
    for (int r = 0; r < numProcs; ++r) {
      const pid_t pid = fork();
      if (pid == 0) {
        std::vector<std::string> args(argv, argv + argc);
        args.push_back("--rank"); args.push_back(std::to_string(r));
        args.push_back("--id-file"); args.push_back(tmpl);
        std::vector<char *> cargs;
        for (auto &a : args) cargs.push_back(const_cast<char *>(a.c_str()));
        cargs.push_back(nullptr);
        execv("/proc/self/exe", cargs.data());
        std::perror("execv"); _exit(127);
      }
      kids.push_back(pid);
    }
    int rc = 0;
    for (pid_t k : kids) { int st = 0; waitpid(k, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1; }
    std::remove(tmpl);
    return rc;
  }
  if (thisProc < 0) thisProc = 0;
  const bool MasterProc = thisProc == 0;                   // MPIW:24
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

  integrator mcIntegrator = new_Integrator(domain, status, thisProc); printStatus(status);                // DRV:533; rank r on device r
  if (numProcs > 1) {                                                                                     // initializeProcesses, MPIW:29-52
    unsigned char id[128];
    if (MasterProc) {
      if (mcb_comm_unique_id(id)) { std::fprintf(stderr, "mcb_comm_unique_id failed: is libnccl.so.2 loadable?\n"); return 1; }
      FILE *f = std::fopen((idFile + ".tmp").c_str(), "wb");
      if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) { std::perror("id file"); return 1; }
      std::fclose(f);
      std::rename((idFile + ".tmp").c_str(), idFile.c_str());                                             // appears complete or not at all
    } else {
      FILE *f = nullptr;
      for (int tries = 0; tries < 3000 && !(f = std::fopen(idFile.c_str(), "rb")); ++tries)
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
      if (!f || std::fread(id, 1, sizeof(id), f) != sizeof(id)) { std::fprintf(stderr, "rank %d: no NCCL id\n", thisProc); return 1; }
      std::fclose(f);
    }
    initializeProcesses(mcIntegrator, numProcs, thisProc, id, status); printStatus(status);
  }
  mcIntegrator.minInverseTableSize = mcIntegrator.minForwardTableSize = 10001;                            // nPhaseIntervals DRV:71
  if (views) {                                                                                            // the 5 I3RC view angles
    mcIntegrator.options.useRussianRouletteForIntensity = 1; mcIntegrator.options.zetaMin = 0.3f;           // DRV:78-79
    specifyParameters(mcIntegrator, status, {1.0f, 0.866f, 0.866f, 0.5f, 0.5f}, {0.0f, 0.0f, 180.0f, 0.0f, 180.0f});
  } else {
    specifyParameters(mcIntegrator, status);
  }
  printStatus(status);
  randomNumberSequence randoms = new_RandomNumberSequence({iseed, 1, 0});                                  // DRV:901

  if (numProcs == 1) {
    // one batch through the module API as the reference's worker loop calls it (DRV:956-1011) ...
    photonStream incomingPhotons = new_PhotonStream(solarMu, solarAzimuth, numPhotonsPerBatch, randoms, status); printStatus(status);
    int64_t numPhotonsProcessed = 0;
    computeRadiativeTransfer(mcIntegrator, domain, randoms, incomingPhotons, numPhotonsPerBatch, numPhotonsProcessed, status); printStatus(status);
    Results one;
    reportResults(mcIntegrator, one, status); printStatus(status);
    std::printf("first batch: %lld photons  meanFluxUp %.6f meanFluxDown %.6f meanFluxAbsorbed %.6f\n", (long long)numPhotonsProcessed,
                one.meanFluxUp, one.meanFluxDown, one.meanFluxAbsorbed);
    randoms.nextPhotonId = 0;                              // the batch loop below starts again at photon 0 (same ids as a --ranks run)
  }
  // ... then the whole batch loop with the driver's moments kept on the device: this rank's block of batches
  const int64_t myBatches = numBatches / numProcs;
  randoms.nextPhotonId = (uint64_t)thisProc * (uint64_t)myBatches * (uint64_t)numPhotonsPerBatch;
  runBatches(mcIntegrator, domain, randoms, solarMu, solarAzimuth, myBatches, numPhotonsPerBatch, status); printStatus(status);
  if (numProcs > 1) { sumAcrossProcesses_statistics(mcIntegrator, status, 0); printStatus(status); }       // DRV:1151-1166
  if (!MasterProc) { finalizeProcesses(mcIntegrator); finalize_Integrator(mcIntegrator); return 0; }
  Statistics stats;
  reportStatistics(mcIntegrator, 1.0, stats, status, true, !outPrefix.empty()); printStatus(status);
  std::printf("ranks %d batches %lld photons %lld\n", numProcs, (long long)stats.batchesCompleted, (long long)stats.totalNumPhotons);
  std::printf("means %.15e %.15e %.15e  errors %.15e %.15e %.15e\n", stats.meanFlux[0], stats.meanFlux[1], stats.meanFlux[2],
              stats.meanFlux[3], stats.meanFlux[4], stats.meanFlux[5]);
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
  if (!outPrefix.empty()) {                                                                                // DRV:1240-1258
    const std::vector<float> mus = views ? std::vector<float>{1.0f, 0.866f, 0.866f, 0.5f, 0.5f} : std::vector<float>{};
    const std::vector<float> phis = views ? std::vector<float>{0.0f, 0.0f, 180.0f, 0.0f, 180.0f} : std::vector<float>{};
    RadianceOptions ro;
    ro.useRussianRouletteForIntensity = mcIntegrator.options.useRussianRouletteForIntensity != 0; ro.zetaMin = mcIntegrator.options.zetaMin;
    writeResults_ASCII(deck + ".dom", stats.totalNumPhotons, (int)stats.batchesCompleted, mcIntegrator.options.useRayTracing != 0,
                       mcIntegrator.options.useRussianRoulette != 0, false, 7.0f, 1.0, solarMu, solarAzimuth, domain.surfaceAlbedo,
                       domain.xPosition, domain.yPosition, domain.zPosition, outPrefix + "_flux.out", stats, outPrefix + "_absprof.out",
                       outPrefix + "_absvol.out", views ? outPrefix + "_rad.out" : std::string(), mus, phis, ro, status);
    printStatus(status);
  }
  if (numProcs > 1) finalizeProcesses(mcIntegrator);
  finalize_Integrator(mcIntegrator);
  return 0;
}
