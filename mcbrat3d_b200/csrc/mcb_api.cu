// mcb_api.cu -- the C ABI (include/mcbrat_cuda.h): handle, staging into HBM, launches,
// normalisation and read-back.  No photon arithmetic happens on the host.
#include <dlfcn.h>
#include <nccl.h>                      // types and prototypes only: libnccl is bound at run time (mcb_comm_init)

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "mcb_device.cuh"

// launchers implemented next to their kernels
void mcb_launch_reference_batch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
                                int numSMs, cudaStream_t stream);
void mcb_launch_trace(const DevDomain &P, long long nPhotons, const float *rn, long long rnStride,
                      mcb_event *events, int maxEventsPerPhoton, int *eventCount, cudaStream_t stream);
void mcb_launch_fast_batch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
                           int numSMs, unsigned long long *workCounter, cudaStream_t stream);
void mcb_launch_philox_kat(uint64_t seed, uint64_t photon, int n, uint32_t *out, cudaStream_t stream);
double mcb_run_gather_probe(size_t bytes, int inFlight, int blocksPerSM, int iterations, int numSMs, cudaStream_t stream);
// mcb_stage.cu: device-side packing / validation, normalisation, emission CDF
void mcb_launch_distance_map(const DevDomain &P, int cap, uint8_t *dist, uint8_t *scratch, int *leapCount, int numSMs, cudaStream_t stream);
void mcb_launch_pack_crop(const DevDomain &P, float *ext, int numSMs, cudaStream_t stream);
void mcb_launch_column_ranges(const DevDomain &P, uint32_t *range, int *count, int *offset, int *tileSum, int *sum, int numSMs,
                              cudaStream_t stream);
void mcb_launch_column_fill(const DevDomain &P, const uint32_t *range, const int *offset, uint32_t *recC,
                            uint32_t *cellC, uint2 *colTab, int numSMs, cudaStream_t stream);
void mcb_launch_pack_field(const DevDomain &P, int brick, float *ext, uint32_t *mask, float *layerExt, int *flags, const uint8_t *dist, int *leapCount,
                           int numSMs, cudaStream_t stream);
void mcb_launch_pack_records(const DevDomain &P, uint32_t *rec, int *flags, int numSMs, cudaStream_t stream);
bool mcb_fast_reads_bricks(const DevDomain &P);
void mcb_launch_gather_column_cdf(const double *voxelCDF, int nx, int ny, int nz, double *colCDF, int numSMs, cudaStream_t stream);
void mcb_launch_normalise(const DevDomain &P, float numPhotons, float *out, int numSMs, cudaStream_t stream);
long long mcb_emission_tiles(long long cells);
void mcb_launch_emission_cdf(const DevDomain &P, const double *temps, double a, double b, double lambda5, void *scratch,
                             double *cdf, int *flags, int numSMs, cudaStream_t stream);
void mcb_launch_emission_normalise(double *cdf, long long cells, const void *total, int numSMs, cudaStream_t stream);
void mcb_launch_assemble_optics(int nx, int ny, int nz, int nc, const int *kind, const int *physIndex, const int *nTable,
                                const int *zLevelBase, const float *const *key, const double *const *ext,
                                const double *const *ssa, const int32_t *const *idx, const float *kmin, const float *kmax,
                                int nPhys, const double *massConc, const double *Reff, const double *numConc, int setup,
                                double *totalExt, double *cumExt, double *ssaOut, int32_t *phaseIdx, int *flags,
                                int numSMs, cudaStream_t stream);
void mcb_launch_inverse_table(const int *offsets, const float *mus, const float *values, int nEntries, int nSteps, float *out,
                              float *cdfScratch, cudaStream_t stream);
void mcb_launch_lobatto_inputs(const int *coefOff, const float *coefs, const int *nodeOff, int nEntries, long long nNodes,
                               float *mus, float *values, int numSMs, cudaStream_t stream);
void mcb_launch_forward_tables(const int *coefOff, const float *coefs, const int *angOff, const float *angles, const float *values,
                               int nEntries, int nSteps, float hybridWidth, float *orig, float *fwd, float *scratch, int numSMs,
                               cudaStream_t stream);
void mcb_launch_forward_table(const int *offsets, const float *coefs, int nEntries, int nSteps, float *out, int numSMs,
                              cudaStream_t stream);
void mcb_launch_frequency_distribution(const double *cdf, int nLambda, long long totalPhotons, uint64_t seed,
                                       unsigned long long *counts, int numSMs, cudaStream_t stream);
long long mcb_stats_elements(const DevDomain &P);
void mcb_launch_stats_accumulate(const DevDomain &P, const float *results, double *stats, double weight, int numSMs,
                                 cudaStream_t stream);
void mcb_launch_stats_finalise(const double *stats, long long n, double solarFlux, double *out, int numSMs,
                               cudaStream_t stream);

struct mcb_handle {
  int device = 0;
  int numSMs = 148;
  cudaStream_t ownStream = nullptr, stream = nullptr;
  cudaEvent_t evStart = nullptr, evStop = nullptr;
  bool timed = false;
  std::string err;
  DevDomain P;
  bool haveGrid = false, haveOptics = false, haveSource = false;
  bool haveTemps = false;                        // dTemps holds the temperatures of the current grid
  bool packedLin = false, packedBrk = false;     // which layouts of the extinction field hold the current optics
  int maskKnob = 0;                              // mcb_options.tuneExtMask the packed field was set up with
  bool haveDist = false;                         // dDist holds the vacuum-distance map of the current optics
  bool haveInv[MCB_MAX_COMP] = {false}, haveFwd[MCB_MAX_COMP] = {false};
  std::vector<double> xE, yE, zE;
  std::map<void **, size_t> slotBytes;    // capacity of every re-stageable slot (reused while large enough)
  // re-stageable slots (freed on re-set)
  void *dXE = nullptr, *dYE = nullptr, *dZE = nullptr;
  void *dTotalExt = nullptr, *dCumExt = nullptr, *dSsa = nullptr, *dPhaseIdx = nullptr;
  void *dExt32 = nullptr, *dExtBrick = nullptr, *dRec = nullptr;
  void *dExtMask = nullptr, *dExtMaskBrick = nullptr, *dLayerExt = nullptr;     // occupancy bitmap of fields too large for L2
  void *dDist = nullptr, *dDistScratch = nullptr;   // vacuum-distance map (u8 per cell) the packed fields are encoded with
  void *dColRange = nullptr, *dColCount = nullptr, *dColOffset = nullptr, *dColTab = nullptr;   // column-compressed storage
  void *dColTiles = nullptr;
  std::vector<float> hLayer;                     // host copy of layerExt (read with the leap count: the crop decision needs it)
  void *dRecC = nullptr, *dCellC = nullptr, *dExtCrop = nullptr, *dTallyC = nullptr;
  void *dInv[MCB_MAX_COMP] = {nullptr}, *dFwd[MCB_MAX_COMP] = {nullptr}, *dFwdOrig[MCB_MAX_COMP] = {nullptr};
  int invE[MCB_MAX_COMP] = {0}, fwdE[MCB_MAX_COMP] = {0};
  void *dColCDF = nullptr, *dVoxelCDF = nullptr, *dTemps = nullptr, *dScratch = nullptr, *dResults = nullptr;
  int *dFlags = nullptr;
  void *dMassConc = nullptr, *dReff = nullptr, *dNumConc = nullptr, *dAsmTables = nullptr;   // physical state (commonDomain)
  int nPhys = 0; bool havePhysical = false, haveNumConc = false;
  void *dStats = nullptr, *dStatsOut = nullptr; long long nStats = 0;   // batch statistics: moments, finalised copy
  double *dTally = nullptr; long long nTally = 0;
  unsigned long long *dCounters = nullptr;     // CNT_N counters + 1 work counter
  double *hTally = nullptr; long long hTallyCap = 0;   // pinned
  ncclComm_t comm = nullptr; int commRanks = 1, commRank = 0;   // mcb_comm_init
};

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------
// The library has no link-time dependency on libnccl: a single-GPU run needs none, and inside a process that already
// carries an NCCL (PyTorch bundles its own) dlopen by SONAME returns THAT copy instead of loading a second one.
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  std::string err;
};
static NcclApi *nccl_api() {
  static NcclApi api;
  if (api.lib || !api.err.empty()) return &api;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
  if (!api.lib) { api.err = std::string("libnccl.so.2 not found: ") + dlerror(); return &api; }
  auto sym = [&](const char *n) { void *p = dlsym(api.lib, n); if (!p && api.err.empty()) api.err = std::string("missing NCCL symbol ") + n; return p; };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  if (!api.err.empty()) { dlclose(api.lib); api.lib = nullptr; }
  return &api;
}

#define FAIL(h, ...)                                                     \
  do {                                                                   \
    char _b[512]; snprintf(_b, sizeof(_b), __VA_ARGS__);                 \
    if (h) (h)->err = _b;                                                \
    return 1;                                                            \
  } while (0)
#define CK(h, call)                                                      \
  do {                                                                   \
    cudaError_t _e = (call);                                             \
    if (_e != cudaSuccess) FAIL(h, "%s: %s", #call, cudaGetErrorString(_e)); \
  } while (0)

// Make *slot a device buffer of at least `bytes` (kept while large enough: re-staging a domain of the
// same shape does no cudaMalloc / cudaFree, which would serialise the device).
static int reserve(mcb_handle *h, void **slot, size_t bytes) {
  CK(h, cudaSetDevice(h->device));
  if (bytes == 0) bytes = 16;
  auto it = h->slotBytes.find(slot);
  if (*slot && it != h->slotBytes.end() && it->second >= bytes) return 0;
  if (*slot) { cudaFree(*slot); *slot = nullptr; }
  CK(h, cudaMalloc(slot, bytes));
  h->slotBytes[slot] = bytes;
  return 0;
}

// Copy a host array into a slot, asynchronously on the handle's stream.  Every mcb_set_* ends with
// settle(): the library never retains the host pointer.
static int stage_async(mcb_handle *h, void **slot, const void *src, size_t bytes) {
  if (reserve(h, slot, bytes)) return 1;
  if (src && bytes) CK(h, cudaMemcpyAsync(*slot, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return 0;
}
static int settle(mcb_handle *h) {
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}
static int stage(mcb_handle *h, void **slot, const void *src, size_t bytes) {
  return stage_async(h, slot, src, bytes) || settle(h);
}

static double sp64h(double x) {
  x = std::fabs(x);
  if (x == 0.0) return 2.2250738585072014e-308;
  double s = std::nextafter(x, INFINITY) - x;
  return s < 2.2250738585072014e-308 ? 2.2250738585072014e-308 : s;
}

extern "C" {

int mcb_version(void) { return 100; }

void mcb_default_options(mcb_options *o) {          // INT:53-96
  memset(o, 0, sizeof(*o));
  o->useRayTracing = 1; o->useRussianRoulette = 1; o->russianRouletteW = 1.0f;
  o->useRussianRouletteForIntensity = 0; o->zetaMin = 0.3f;
  o->useHybridPhaseFunsForIntenCalcs = 0; o->numOrdersOrigPhaseFunIntenCalcs = 0;
  o->limitIntensityContributions = 0; o->maxIntensityContribution = 3.402823466e+38f;
  o->LW_flag = -1.0f; o->arithmetic = MCB_ARITH_FAST;
}

int mcb_create(int device, mcb_handle **out) {
  if (!out) return 1;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return 2;   // no CPU fallback
  if (device < 0 || device >= count) return 3;
  mcb_handle *h = new mcb_handle();
  h->device = device;
  memset(&h->P, 0, sizeof(h->P));
  mcb_default_options(&h->P.opt);
  if (cudaSetDevice(device) != cudaSuccess) { delete h; return 4; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->numSMs = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return 5; }
  h->stream = h->ownStream;
  cudaEventCreate(&h->evStart); cudaEventCreate(&h->evStop);
  if (cudaMalloc((void **)&h->dCounters, sizeof(unsigned long long) * 32) != cudaSuccess) { delete h; return 6; }
  cudaMemset(h->dCounters, 0, sizeof(unsigned long long) * 32);
  h->P.counters = h->dCounters;
  if (cudaMalloc((void **)&h->dFlags, sizeof(int) * 4) != cudaSuccess) { delete h; return 6; }
  *out = h;
  return 0;
}

int mcb_destroy(mcb_handle *h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  void *slots[] = {h->dXE, h->dYE, h->dZE, h->dTotalExt, h->dCumExt, h->dSsa, h->dPhaseIdx,
                   h->dExt32, h->dRec, h->dVoxelCDF, h->dTally, h->dCounters,
                   h->dTemps, h->dScratch, h->dResults, (void *)h->dFlags, h->dStats, h->dStatsOut,
                   h->dMassConc, h->dReff, h->dNumConc, h->dAsmTables, h->dExtMask, h->dLayerExt,
                   h->dExtBrick, h->dExtMaskBrick, h->dColCDF, h->dDist, h->dDistScratch,
                   h->dColRange, h->dColCount, h->dColOffset, h->dColTab, h->dColTiles, h->dRecC, h->dCellC, h->dExtCrop, h->dTallyC};
  for (void *p : slots) if (p) cudaFree(p);
  for (int c = 0; c < MCB_MAX_COMP; ++c) {
    if (h->dInv[c]) cudaFree(h->dInv[c]);
    if (h->dFwd[c]) cudaFree(h->dFwd[c]);
    if (h->dFwdOrig[c]) cudaFree(h->dFwdOrig[c]);
  }
  if (h->hTally) cudaFreeHost(h->hTally);
  if (h->comm && nccl_api()->lib) nccl_api()->CommDestroy(h->comm);
  cudaEventDestroy(h->evStart); cudaEventDestroy(h->evStop);
  cudaStreamDestroy(h->ownStream);
  delete h;
  return 0;
}

int mcb_last_error(const mcb_handle *h, char *buf, int len) {
  if (!buf || len <= 0) return 1;
  const char *s = h ? h->err.c_str() : "invalid handle";
  snprintf(buf, (size_t)len, "%s", s);
  return 0;
}

int mcb_set_stream(mcb_handle *h, void *cudaStream) {
  if (!h) return 1;
  h->stream = cudaStream ? (cudaStream_t)cudaStream : h->ownStream;
  return 0;
}

int mcb_synchronize(mcb_handle *h) {
  if (!h) return 1;
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// Division by a launch-invariant divisor d for dividends n < 2^31 (Granlund & Montgomery 1994):
// with l = ceil(log2 d) and M = ceil(2^(31+l) / d) < 2^32, floor(n / d) = (M * n) >> (31 + l).
static void magic_divisor(uint32_t d, uint32_t *M, int *S) {
  int l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned long long num = 1ull << (31 + l);
  *M = (uint32_t)((num + d - 1) / d);
  *S = 31 + l;
}

int mcb_set_grid(mcb_handle *h, int nx, int ny, int nz,
                 const double *xEdges, const double *yEdges, const double *zEdges) {
  if (!h) return 1;
  if (nx < 1 || ny < 1 || nz < 1 || !xEdges || !yEdges || !zEdges) FAIL(h, "mcb_set_grid: bad arguments");
  for (int i = 0; i < nx; ++i) if (!(xEdges[i + 1] > xEdges[i])) FAIL(h, "new_Domain: Positions must be increasing, unique.");
  for (int i = 0; i < ny; ++i) if (!(yEdges[i + 1] > yEdges[i])) FAIL(h, "new_Domain: Positions must be increasing, unique.");
  for (int i = 0; i < nz; ++i) if (!(zEdges[i + 1] > zEdges[i])) FAIL(h, "new_Domain: Positions must be increasing, unique.");
  h->xE.assign(xEdges, xEdges + nx + 1); h->yE.assign(yEdges, yEdges + ny + 1); h->zE.assign(zEdges, zEdges + nz + 1);
  if (stage(h, &h->dXE, xEdges, sizeof(double) * (nx + 1))) return 1;
  if (stage(h, &h->dYE, yEdges, sizeof(double) * (ny + 1))) return 1;
  if (stage(h, &h->dZE, zEdges, sizeof(double) * (nz + 1))) return 1;
  DevDomain &P = h->P;
  P.nx = nx; P.ny = ny; P.nz = nz;
  P.xE = (const double *)h->dXE; P.yE = (const double *)h->dYE; P.zE = (const double *)h->dZE;
  P.x0 = xEdges[0]; P.y0 = yEdges[0]; P.z0 = zEdges[0];
  P.xMax = xEdges[nx]; P.yMax = yEdges[ny]; P.zMax = zEdges[nz];
  // new_Integrator INT:163-181: spacing held in default-real locals (quirk q1)
  const float deltaX = (float)(xEdges[1] - xEdges[0]), deltaY = (float)(yEdges[1] - yEdges[0]),
              deltaZ = (float)(zEdges[1] - zEdges[0]);
  bool xyReg = true, zReg = true;
  for (int i = 0; i < nx; ++i)
    if (!(std::fabs((xEdges[i + 1] - xEdges[i]) - (double)deltaX) <= 2.0 * sp64h(xEdges[i + 1]))) xyReg = false;
  for (int i = 0; i < ny; ++i)
    if (!(std::fabs((yEdges[i + 1] - yEdges[i]) - (double)deltaY) <= 2.0 * sp64h(yEdges[i + 1]))) xyReg = false;
  for (int i = 0; i < nz; ++i)
    if (!(std::fabs((zEdges[i + 1] - zEdges[i]) - (double)deltaZ) <= sp64h(zEdges[i + 1]))) zReg = false;
  P.xyRegular = xyReg ? 1 : 0; P.zRegular = zReg ? 1 : 0;
  P.deltaX = xyReg ? (double)deltaX : 0.0; P.deltaY = xyReg ? (double)deltaY : 0.0;
  P.deltaZ = zReg ? (double)deltaZ : 0.0;
  P.fx0 = (float)P.x0; P.fy0 = (float)P.y0; P.fz0 = (float)P.z0;
  P.fLx = (float)(P.xMax - P.x0); P.fLy = (float)(P.yMax - P.y0); P.fLz = (float)(P.zMax - P.z0);
  // The throughput kernel is statistical, not bit-faithful: a grid whose spacings are uniform to 1e-6 is stepped
  // incrementally with its mean spacing whether or not quirk q1 calls it regular (the reference-arithmetic kernel and
  // the normalisation keep the reference's own flags).
  auto uniformAxis = [](const double *e, int n) {
    const double mean = (e[n] - e[0]) / n;
    for (int i = 0; i < n; ++i) if (!(std::fabs((e[i + 1] - e[i]) - mean) <= 1.0e-6 * mean)) return false;
    return true;
  };
  P.uniform = (uniformAxis(xEdges, nx) && uniformAxis(yEdges, ny) && uniformAxis(zEdges, nz)) ? 1 : 0;
  P.fhx = (float)((P.xMax - P.x0) / nx); P.fhy = (float)((P.yMax - P.y0) / ny); P.fhz = (float)((P.zMax - P.z0) / nz);
  P.finvLx = 1.0f / P.fLx; P.finvLy = 1.0f / P.fLy; P.fzMax = (float)P.zMax;
  P.finvhx = 1.0f / P.fhx; P.finvhy = 1.0f / P.fhy; P.finvhz = 1.0f / P.fhz;
  // padded extinction field: MCB_GHOST cells on every side (see mcb_set_optics)
  {                                                  // x-fastest layout
    DevDomain::ExtField &F = P.lin;
    F.nxp = nx + 2 * MCB_GHOST; F.nyp = ny + 2 * MCB_GHOST; F.cY = F.cZ = 0;
    F.padded = (long long)F.nxp * F.nyp * (nz + 2 * MCB_GHOST);
    F.origin = MCB_GHOST + F.nxp * (MCB_GHOST + F.nyp * MCB_GHOST);
    magic_divisor((uint32_t)F.nxp * (uint32_t)F.nyp, &F.divSliceM, &F.divSliceS);
    magic_divisor((uint32_t)F.nxp, &F.divRowM, &F.divRowS);
    magic_divisor((uint32_t)nx * (uint32_t)ny, &P.divColsM, &P.divColsS);       // unpadded cell index -> (ix, iy, iz)
    magic_divisor((uint32_t)nx, &P.divNxM, &P.divNxS);
  }
  {                                                  // 2x2x2 bricks: padded dimensions rounded up to even
    DevDomain::ExtField &F = P.brk;
    const int nxp = (nx + 2 * MCB_GHOST + 1) & ~1, nyp = (ny + 2 * MCB_GHOST + 1) & ~1, nzp = (nz + 2 * MCB_GHOST + 1) & ~1;
    if ((long long)nxp * nyp * nzp >= (1LL << 31)) FAIL(h, "mcb_set_grid: more than 2^31 cells");
    const int bx = nxp / 2, by = nyp / 2;
    F.nxp = nxp; F.nyp = nyp; F.cY = 4 * bx; F.cZ = 4 * bx * by;
    F.padded = (long long)nxp * nyp * nzp;
    F.origin = (int)mcb_brick_address(MCB_GHOST, MCB_GHOST, MCB_GHOST, bx, by);
    magic_divisor((uint32_t)bx * (uint32_t)by, &F.divSliceM, &F.divSliceS);      // brick index -> (bx, by, bz)
    magic_divisor((uint32_t)bx, &F.divRowM, &F.divRowS);
  }
  h->haveGrid = true; h->haveOptics = false; h->haveSource = false; h->havePhysical = false; h->haveTemps = false;
  return 0;
}

static int finish_optics(mcb_handle *h, int nc, double albedo, bool zeroFlags = true);
static int setup_packed_field(mcb_handle *h);
static int setup_compact_storage(mcb_handle *h);

// the thermal source's CDF is in HBM: derive the compact column weights and publish the pointers
static int finish_thermal_source(mcb_handle *h, double fracAtmsPower) {
  DevDomain &P = h->P;
  if (reserve(h, &h->dColCDF, sizeof(double) * (size_t)P.ny * P.nz)) return 1;
  mcb_launch_gather_column_cdf((const double *)h->dVoxelCDF, P.nx, P.ny, P.nz, (double *)h->dColCDF, h->numSMs, h->stream);
  CK(h, cudaGetLastError());
  P.source = 1; P.fracAtmsPower = fracAtmsPower;
  P.voxelCDF = (const double *)h->dVoxelCDF; P.colCDF = (const double *)h->dColCDF;
  return 0;
}

static int pack_field(mcb_handle *h, bool brick) {
  DevDomain &P = h->P;
  if (brick ? h->packedBrk : h->packedLin) return 0;
  // Vacuum-distance encoding (mcb_stage.cu): on grids the photon-pool kernels march (uniform, at least a ghost shell
  // wide, field read without the occupancy bitmap) a cell without extinction holds -D, D = its Chebyshev distance to
  // the nearest cell with extinction; every marcher clamps at 0, the pool kernels leap (march_leap, mcb_march.cuh) --
  // provided a fair share of the domain lies deep enough in vacuum to pay for the leap code (P.leap): else the field
  // stays as it is and the launchers pick the kernels without it.  Bitmap-marched fields: the same decision from the
  // number of layers that are clear throughout (layer tables, mcb_launch_pack_field).
#ifdef MCB_NO_LEAP
  const bool candidate = false;
#else
  const bool candidate = P.uniform && P.nx >= MCB_GHOST && P.ny >= MCB_GHOST;
#endif
  int *leapCount = h->dFlags + 1;
  if (candidate && !P.lin.mask && !h->haveDist) {
    const size_t cells = (size_t)P.nx * P.ny * P.nz;
    if (reserve(h, &h->dDist, cells) || reserve(h, &h->dDistScratch, cells)) return 1;
    int cap = MCB_LEAP_CAP;
    cap = cap < P.nx ? cap : P.nx; cap = cap < P.ny ? cap : P.ny;
    mcb_launch_distance_map(P, cap, (uint8_t *)h->dDist, (uint8_t *)h->dDistScratch, leapCount, h->numSMs, h->stream);
    CK(h, cudaGetLastError());
    int deep = 0;
    CK(h, cudaMemcpyAsync(&deep, leapCount, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (settle(h)) return 1;
    h->haveDist = true;
    P.leap = (long long)deep * 32 >= (long long)cells ? 1 : 0;
  }
  const bool encode = candidate && !P.lin.mask && P.leap;
  mcb_launch_pack_field(P, brick ? 1 : 0, (float *)(brick ? h->dExtBrick : h->dExt32),
                        (uint32_t *)(brick ? P.brk.mask : P.lin.mask), (float *)P.layerExt, h->dFlags,
                        encode ? (const uint8_t *)h->dDist : nullptr, leapCount, h->numSMs, h->stream);
  CK(h, cudaGetLastError());
  if (P.lin.mask) {                                   // layers clear throughout, far enough from the nearest cloudy one
    int layers = 0;
    h->hLayer.resize((size_t)P.nz + 2 * MCB_GHOST);
    CK(h, cudaMemcpyAsync(&layers, leapCount, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->hLayer.data(), P.layerExt, sizeof(float) * h->hLayer.size(), cudaMemcpyDeviceToHost, h->stream));
    if (settle(h)) return 1;
    P.leap = candidate && layers * 16 >= P.nz ? 1 : 0;
  }
  (brick ? h->packedBrk : h->packedLin) = true;
  return 0;
}

// Buffers of the packed single-precision extinction field (both layouts), the decision whether it is marched through
// an occupancy bitmap, and the packing of the layout the next launch reads (the other one on demand, run()).
// Occupancy bitmap: only for fields that do not stay L2-resident (default: padded field > 48 MB; the C3 field,
// 15.5 MB, is served by L2 and gains nothing); mcb_options.tuneExtMask forces it off / on (measurements, tests).
static int setup_packed_field(mcb_handle *h) {
  DevDomain &P = h->P;
  const size_t padded = (size_t)P.brk.padded;          // the larger of the two layouts
  if (reserve(h, &h->dExt32, sizeof(float) * (size_t)P.lin.padded)) return 1;
  if (reserve(h, &h->dExtBrick, sizeof(float) * (size_t)P.brk.padded)) return 1;
  P.lin.ext = (const float *)h->dExt32 + P.lin.origin;
  P.brk.ext = (const float *)h->dExtBrick + P.brk.origin;
  const int knob = P.opt.tuneExtMask;
  const bool useMask = knob != 0 ? knob > 0 : sizeof(float) * padded > ((size_t)48 << 20);
  P.lin.mask = P.brk.mask = nullptr; P.layerExt = P.layerLeap = P.layerCum = nullptr;
  if (useMask) {
    if (reserve(h, &h->dExtMask, sizeof(uint32_t) * ((padded + 31) / 32))) return 1;
    if (reserve(h, &h->dExtMaskBrick, sizeof(uint32_t) * ((padded + 31) / 32))) return 1;
    const size_t nLayer = (size_t)(P.nz + 2 * MCB_GHOST + 2);             // layerExt | layerLeap | layerCum
    if (reserve(h, &h->dLayerExt, sizeof(float) * 3 * nLayer)) return 1;
    P.lin.mask = (const uint32_t *)h->dExtMask; P.brk.mask = (const uint32_t *)h->dExtMaskBrick;
    P.layerExt = (const float *)h->dLayerExt; P.layerLeap = P.layerExt + nLayer; P.layerCum = P.layerLeap + nLayer;
  }
  h->maskKnob = knob;
  h->packedLin = h->packedBrk = false; h->haveDist = false; P.leap = 0;
  return pack_field(h, mcb_fast_reads_bricks(P));
}

int mcb_set_optics(mcb_handle *h, int nc, const double *totalExt, const double *cumExt,
                   const double *ssa, const int32_t *phaseIdx, double albedo) {
  if (!h) return 1;
  if (!h->haveGrid) FAIL(h, "mcb_set_optics: call mcb_set_grid first");
  if (nc < 1 || nc > MCB_MAX_COMP) FAIL(h, "mcb_set_optics: number of components must be 1..%d", MCB_MAX_COMP);
  if (!totalExt || !cumExt || !ssa || !phaseIdx) FAIL(h, "mcb_set_optics: null array");
  DevDomain &P = h->P;
  const size_t cells = (size_t)P.nx * P.ny * P.nz;
  h->haveOptics = false;
  // The four host arrays cross PCIe once, as they are; the packed single-precision copies the fast kernel reads
  // (ghost-shelled extinction: periodic replicas in x, y -- OPT:1782-1796 becomes data instead of per-cell tests --
  // and empty layers above and below, OPT:1801-1812) and the argument checks of addOpticalComponent are produced
  // in HBM by mcb_stage.cu.
  if (stage_async(h, &h->dTotalExt, totalExt, sizeof(double) * cells)) return 1;
  if (stage_async(h, &h->dCumExt, cumExt, sizeof(double) * cells * nc)) return 1;
  if (stage_async(h, &h->dSsa, ssa, sizeof(double) * cells * nc)) return 1;
  if (stage_async(h, &h->dPhaseIdx, phaseIdx, sizeof(int32_t) * cells * nc)) return 1;
  return finish_optics(h, nc, albedo);
}

// Column-compressed event data and the layer-cropped field for the pool flux kernel (mcb_device.cuh); needs the packed
// field's layer table (setup_packed_field) and the event records.
static int setup_compact_storage(mcb_handle *h) {
  DevDomain &P = h->P;
  // Column-compressed event data for the pool flux kernel (fields marched through the bitmap, i.e. too large for L2): for
  // the cells inside the per-column ranges the event record, the cell index and an absorption tally, densely, column by column.
  P.colTab = nullptr; P.recC = nullptr; P.cellC = nullptr; P.nCompact = 0; P.tallyC = nullptr;
  P.crp.ext = nullptr; P.cropLo = 0; P.cropN = 0;
  // (tuneExtMask = 1 keeps the run on the bitmap: nothing of this is built)
  if (P.lin.mask && P.opt.tuneExtMask != 1 && P.uniform && P.nx >= MCB_GHOST && P.ny >= MCB_GHOST && P.nz <= 65535) {
    const size_t cols = (size_t)P.nx * P.ny;
    if (reserve(h, &h->dColRange, sizeof(uint32_t) * cols) || reserve(h, &h->dColCount, sizeof(int) * cols) ||
        reserve(h, &h->dColOffset, sizeof(int) * cols) || reserve(h, &h->dColTab, sizeof(uint2) * (size_t)P.lin.nxp * P.lin.nyp) ||
        reserve(h, &h->dColTiles, sizeof(int) * ((cols + 1023) / 1024 + 1)))
      return 1;
    int *dSum = h->dFlags + 1;
    mcb_launch_column_ranges(P, (uint32_t *)h->dColRange, (int *)h->dColCount, (int *)h->dColOffset, (int *)h->dColTiles, dSum,
                             h->numSMs, h->stream);
    CK(h, cudaGetLastError());
    int nCompact = 0;
    CK(h, cudaMemcpyAsync(&nCompact, dSum, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (settle(h)) return 1;
    const size_t n = nCompact > 0 ? (size_t)nCompact : 1;
    if (reserve(h, &h->dRecC, sizeof(uint32_t) * (n << P.recShift)) ||
        reserve(h, &h->dCellC, sizeof(uint32_t) * n) || reserve(h, &h->dTallyC, sizeof(double) * n))
      return 1;
    CK(h, cudaMemsetAsync(h->dTallyC, 0, sizeof(double) * n, h->stream));
    mcb_launch_column_fill(P, (const uint32_t *)h->dColRange, (const int *)h->dColOffset, (uint32_t *)h->dRecC,
                           (uint32_t *)h->dCellC, (uint2 *)h->dColTab, h->numSMs, h->stream);
    CK(h, cudaGetLastError());
    P.colTab = (const uint2 *)h->dColTab; P.recC = (const uint32_t *)h->dRecC;
    P.cellC = (const uint32_t *)h->dCellC; P.nCompact = nCompact; P.tallyC = (double *)h->dTallyC;
    // The layer-cropped field: the band of layers that hold cloud somewhere (the sign bits of layerExt mark the layers
    // that are clear throughout), bricked, if it is small enough to stay in L2 (same 48 MB limit as the bitmap decision).
    const std::vector<float> &layer = h->hLayer;               // read back with the leap count (pack_field)
    int lo = P.nz, hi = 0;
    for (int k = 0; k < P.nz; ++k) if (!std::signbit(layer[(size_t)k + MCB_GHOST])) { lo = lo < k ? lo : k; hi = k + 1; }
    if (hi > lo) {
      lo &= ~1; hi = (hi + 1) & ~1;
      DevDomain::ExtField &F = P.crp;
      F = P.brk;                                               // same padded row / slice geometry and divisors
      F.mask = nullptr; F.ext = nullptr;
      F.padded = (long long)F.nxp * F.nyp * (hi - lo);
      F.origin = (int)mcb_brick_address(MCB_GHOST, MCB_GHOST, 0, F.nxp / 2, F.nyp / 2);
      if (sizeof(float) * (size_t)F.padded <= ((size_t)48 << 20)) {
        if (reserve(h, &h->dExtCrop, sizeof(float) * (size_t)F.padded)) return 1;
        P.cropLo = lo; P.cropN = hi - lo;
        mcb_launch_pack_crop(P, (float *)h->dExtCrop, h->numSMs, h->stream);
        CK(h, cudaGetLastError());
        F.ext = (const float *)h->dExtCrop + F.origin;
      }
    }
  }
  return 0;
}

// The dense f64 arrays are in HBM (uploaded by mcb_set_optics or built by mcb_assemble_optics): derive the packed
// single-precision copies, run the argument checks, find maxval(totalExt), publish the pointers.
static int finish_optics(mcb_handle *h, int nc, double albedo, bool zeroFlags) {
  DevDomain &P = h->P;
  const size_t cells = (size_t)P.nx * P.ny * P.nz;
  int recShift = 0;                                   // event record: (nc-1) + nc + ceil(nc/2) words, padded to 2^recShift
  while ((1 << recShift) < 2 * nc - 1 + (nc + 1) / 2) ++recShift;
  if (reserve(h, &h->dRec, sizeof(uint32_t) * (cells << recShift))) return 1;
  P.nc = nc; P.albedo = albedo;
  P.totalExt = (const double *)h->dTotalExt; P.cumExt = (const double *)h->dCumExt;
  P.ssa = (const double *)h->dSsa; P.phaseIdx = (const int32_t *)h->dPhaseIdx;
  P.rec = (const uint32_t *)h->dRec; P.recShift = recShift;
  if (zeroFlags) CK(h, cudaMemsetAsync(h->dFlags, 0, sizeof(int) * 4, h->stream));
  if (setup_packed_field(h)) return 1;
  mcb_launch_pack_records(P, (uint32_t *)h->dRec, h->dFlags, h->numSMs, h->stream);
  CK(h, cudaGetLastError());
  if (setup_compact_storage(h)) return 1;
  int flags4[4] = {0, 0, 0, 0};
  CK(h, cudaMemcpyAsync(flags4, h->dFlags, sizeof(flags4), cudaMemcpyDeviceToHost, h->stream));
  if (settle(h)) return 1;
  const int flags = flags4[0];
  {
    double emax; memcpy(&emax, &flags4[2], sizeof(double));
    P.maxExtinction = (float)emax;                                       // INT:448, default real
  }
  if (flags & 16) FAIL(h, "read_SSPTable: Effective radius outside of table range");
  if (flags & 1) FAIL(h, "addOpticalComponent: extinction must be >= 0.");
  if (flags & 2) FAIL(h, "addOpticalComponent: singleScatteringAlbedo must be between 0 and 1");
  if (flags & 4) FAIL(h, "addOpticalComponent: phase function index is out of bounds");
  for (int c = 0; c < MCB_MAX_COMP; ++c) { h->haveInv[c] = false; h->haveFwd[c] = false; }
  h->haveOptics = true;
  return 0;
}

// ---- per-wavelength assembly on the device (read_SSPTable OPT:147-343 + getOpticalPropertiesByComponent) ----
int mcb_set_physical(mcb_handle *h, int nPhys, const double *massConc, const double *Reff, const double *numConc) {
  if (!h) return 1;
  if (!h->haveGrid) FAIL(h, "mcb_set_physical: call mcb_set_grid first");
  if (nPhys < 0 || (nPhys > 0 && (!massConc || !Reff))) FAIL(h, "mcb_set_physical: bad arguments");
  const size_t cells = (size_t)h->P.nx * h->P.ny * h->P.nz;
  h->havePhysical = false;
  if (nPhys > 0) {
    if (stage_async(h, &h->dMassConc, massConc, sizeof(double) * cells * nPhys)) return 1;
    if (stage_async(h, &h->dReff, Reff, sizeof(double) * cells * nPhys)) return 1;
  }
  h->haveNumConc = numConc != nullptr;
  if (numConc && stage_async(h, &h->dNumConc, numConc, sizeof(double) * h->P.nz)) return 1;
  if (settle(h)) return 1;
  h->nPhys = nPhys; h->havePhysical = true;
  return 0;
}

int mcb_assemble_optics(mcb_handle *h, int nc, const mcb_component *comps, int setup, double albedo) {
  if (!h) return 1;
  if (!h->havePhysical) FAIL(h, "read_SSPTable: the common physical domain has not been staged (mcb_set_physical)");
  if (nc < 1 || nc > MCB_MAX_COMP || !comps) FAIL(h, "mcb_assemble_optics: number of components must be 1..%d", MCB_MAX_COMP);
  DevDomain &P = h->P;
  const size_t cells = (size_t)P.nx * P.ny * P.nz;
  h->haveOptics = false;
  // argument checks that need no data, then one contiguous upload of the small per-wavelength tables
  size_t bytes = 0;
  std::vector<size_t> offKey(nc), offExt(nc), offSsa(nc), offIdx(nc);
  int kind[MCB_MAX_COMP], phys[MCB_MAX_COMP], nTab[MCB_MAX_COMP], zBase[MCB_MAX_COMP];
  float kmin[MCB_MAX_COMP], kmax[MCB_MAX_COMP];
  for (int c = 0; c < nc; ++c) {
    const mcb_component &q = comps[c];
    kind[c] = q.kind; phys[c] = q.physIndex; nTab[c] = q.nTable; zBase[c] = q.zLevelBase; kmin[c] = kmax[c] = 0.0f;
    if (q.kind < 0 || q.kind > 2) FAIL(h, "read_SSPTable: unrecognizable extType");
    if (q.nTable < 1 || !q.ext) FAIL(h, "mcb_assemble_optics: component %d has no table", c + 1);
    const int nLev = q.kind == MCB_COMP_VOLEXT ? P.nz : q.nTable;
    if (q.zLevelBase < 1 || q.zLevelBase + nLev - 1 > P.nz)
      FAIL(h, "addOpticalComponent: arrays don't fit the vertical extent of the domain.");
    if (q.kind == MCB_COMP_VOLEXT) {
      if (q.physIndex < 1 || q.physIndex > h->nPhys || q.nTable < 2 || !q.key || !q.ssa)
        FAIL(h, "mcb_assemble_optics: component %d: bad volExt description", c + 1);
      kmin[c] = kmax[c] = q.key[0];
      for (int i = 0; i < q.nTable; ++i) { kmin[c] = std::fmin(kmin[c], q.key[i]); kmax[c] = std::fmax(kmax[c], q.key[i]); }
    }
    if (q.kind == MCB_COMP_ABSXSEC && !h->haveNumConc) FAIL(h, "read_SSPTable: absXsec component needs numConc");
    if (q.kind == MCB_COMP_PROFILE && (!q.ssa || !q.phaseIdx)) FAIL(h, "mcb_assemble_optics: component %d: bad profile", c + 1);
    auto take = [&](size_t n) { const size_t o = bytes; bytes += (n + 15) & ~(size_t)15; return o; };
    offExt[c] = take(sizeof(double) * q.nTable);
    offSsa[c] = q.ssa ? take(sizeof(double) * q.nTable) : 0;
    offKey[c] = q.key && q.kind == MCB_COMP_VOLEXT ? take(sizeof(float) * q.nTable) : 0;
    offIdx[c] = q.phaseIdx && q.kind == MCB_COMP_PROFILE ? take(sizeof(int32_t) * q.nTable) : 0;
  }
  std::vector<char> blob(bytes, 0);
  for (int c = 0; c < nc; ++c) {
    const mcb_component &q = comps[c];
    memcpy(blob.data() + offExt[c], q.ext, sizeof(double) * q.nTable);
    if (q.ssa) memcpy(blob.data() + offSsa[c], q.ssa, sizeof(double) * q.nTable);
    if (q.key && q.kind == MCB_COMP_VOLEXT) memcpy(blob.data() + offKey[c], q.key, sizeof(float) * q.nTable);
    if (q.phaseIdx && q.kind == MCB_COMP_PROFILE) memcpy(blob.data() + offIdx[c], q.phaseIdx, sizeof(int32_t) * q.nTable);
  }
  if (stage(h, &h->dAsmTables, blob.data(), bytes)) return 1;
  const char *base = (const char *)h->dAsmTables;
  const float *dKey[MCB_MAX_COMP]; const double *dExt[MCB_MAX_COMP], *dSsa[MCB_MAX_COMP]; const int32_t *dIdx[MCB_MAX_COMP];
  for (int c = 0; c < nc; ++c) {
    dExt[c] = (const double *)(base + offExt[c]); dSsa[c] = (const double *)(base + offSsa[c]);
    dKey[c] = (const float *)(base + offKey[c]); dIdx[c] = (const int32_t *)(base + offIdx[c]);
  }
  if (reserve(h, &h->dTotalExt, sizeof(double) * cells)) return 1;
  if (reserve(h, &h->dCumExt, sizeof(double) * cells * nc)) return 1;
  if (reserve(h, &h->dSsa, sizeof(double) * cells * nc)) return 1;
  if (reserve(h, &h->dPhaseIdx, sizeof(int32_t) * cells * nc)) return 1;
  CK(h, cudaMemsetAsync(h->dFlags, 0, sizeof(int) * 4, h->stream));
  mcb_launch_assemble_optics(P.nx, P.ny, P.nz, nc, kind, phys, nTab, zBase, dKey, dExt, dSsa, dIdx, kmin, kmax, h->nPhys,
                             (const double *)h->dMassConc, (const double *)h->dReff, (const double *)h->dNumConc, setup,
                             (double *)h->dTotalExt, (double *)h->dCumExt, (double *)h->dSsa, (int32_t *)h->dPhaseIdx,
                             h->dFlags, h->numSMs, h->stream);
  CK(h, cudaGetLastError());
  return finish_optics(h, nc, albedo, false);
}

// the dense arrays currently staged, in the domain's layout (any pointer may be NULL)
int mcb_get_optics(mcb_handle *h, double *totalExt, double *cumExt, double *ssa, int32_t *phaseIdx) {
  if (!h) return 1;
  if (!h->haveOptics) FAIL(h, "mcb_get_optics: no optical properties are staged");
  CK(h, cudaSetDevice(h->device));
  const size_t cells = (size_t)h->P.nx * h->P.ny * h->P.nz, nc = (size_t)h->P.nc;
  if (totalExt) CK(h, cudaMemcpyAsync(totalExt, h->dTotalExt, sizeof(double) * cells, cudaMemcpyDeviceToHost, h->stream));
  if (cumExt) CK(h, cudaMemcpyAsync(cumExt, h->dCumExt, sizeof(double) * cells * nc, cudaMemcpyDeviceToHost, h->stream));
  if (ssa) CK(h, cudaMemcpyAsync(ssa, h->dSsa, sizeof(double) * cells * nc, cudaMemcpyDeviceToHost, h->stream));
  if (phaseIdx) CK(h, cudaMemcpyAsync(phaseIdx, h->dPhaseIdx, sizeof(int32_t) * cells * nc, cudaMemcpyDeviceToHost, h->stream));
  return settle(h);
}

int mcb_set_inverse_table(mcb_handle *h, int comp, int nS, int nE, const float *T) {
  if (!h) return 1;
  if (!h->haveOptics) FAIL(h, "mcb_set_inverse_table: call mcb_set_optics first");
  if (comp < 1 || comp > h->P.nc || nS < 2 || nE < 1 || !T) FAIL(h, "mcb_set_inverse_table: bad arguments");
  const int c = comp - 1;
  if (stage(h, &h->dInv[c], T, sizeof(float) * (size_t)nS * nE)) return 1;
  h->P.inv[c] = (const float *)h->dInv[c]; h->P.invS[c] = nS; h->invE[c] = nE; h->P.invE[c] = nE; h->haveInv[c] = true;
  return 0;
}

// computeInversePhaseFuncTable INV:26-64 on the device: entry e is given at nAngles[e] points increasing in mu
// (mus / values concatenated); the table is built straight into the slot mcb_set_inverse_table would fill.
int mcb_build_inverse_table(mcb_handle *h, int comp, int nS, int nE, const int32_t *nAngles, const float *mus,
                            const float *values) {
  if (!h) return 1;
  if (!h->haveOptics) FAIL(h, "mcb_build_inverse_table: call mcb_set_optics first");
  if (comp < 1 || comp > h->P.nc || nS < 2 || nE < 1 || !nAngles || !mus || !values)
    FAIL(h, "mcb_build_inverse_table: bad arguments");
  std::vector<int> off(nE + 1, 0);
  for (int e = 0; e < nE; ++e) {
    if (nAngles[e] < 2) FAIL(h, "computeInversePhaseFunction: a phase function needs at least two angles");
    off[e + 1] = off[e] + nAngles[e];
  }
  const size_t total = (size_t)off[nE];
  const int c = comp - 1;
  // scratch layout: offsets | mus | values | cdf
  const size_t bOff = (sizeof(int) * (nE + 1) + 15) & ~(size_t)15, bF = sizeof(float) * total;
  if (reserve(h, &h->dScratch, bOff + 3 * bF)) return 1;
  char *base = (char *)h->dScratch;
  CK(h, cudaMemcpyAsync(base, off.data(), sizeof(int) * (nE + 1), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(base + bOff, mus, bF, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(base + bOff + bF, values, bF, cudaMemcpyHostToDevice, h->stream));
  if (reserve(h, &h->dInv[c], sizeof(float) * (size_t)nS * nE)) return 1;
  mcb_launch_inverse_table((const int *)base, (const float *)(base + bOff), (const float *)(base + bOff + bF), nE, nS,
                           (float *)h->dInv[c], (float *)(base + bOff + 2 * bF), h->stream);
  CK(h, cudaGetLastError());
  if (settle(h)) return 1;
  h->P.inv[c] = (const float *)h->dInv[c]; h->P.invS[c] = nS; h->invE[c] = nE; h->P.invE[c] = nE; h->haveInv[c] = true;
  return 0;
}

// computeInversePhaseFuncTable INV:66-174 for a table whose entries are stored as Legendre moments, without any host
// arithmetic: the Lobatto abscissas (NUM:27-114), the phase function there (SPF:480-498, NUM:187-205) and the inversion
// all run in HBM; only the moments (chi_1.. of entry e: nCoef[e] values, concatenated) cross PCIe.
int mcb_build_inverse_table_legendre(mcb_handle *h, int comp, int nS, int nE, const int32_t *nCoef, const float *coefs) {
  if (!h) return 1;
  if (!h->haveOptics) FAIL(h, "mcb_build_inverse_table_legendre: call mcb_set_optics first");
  if (comp < 1 || comp > h->P.nc || nS < 2 || nE < 1 || !nCoef) FAIL(h, "mcb_build_inverse_table_legendre: bad arguments");
  std::vector<int> off(2 * (nE + 1), 0);                              // coefficient offsets | node offsets
  int *coefOff = off.data(), *nodeOff = off.data() + nE + 1;
  for (int e = 0; e < nE; ++e) {
    if (nCoef[e] < 0) FAIL(h, "mcb_build_inverse_table_legendre: bad arguments");
    coefOff[e + 1] = coefOff[e] + nCoef[e];
    nodeOff[e + 1] = nodeOff[e] + (nCoef[e] > 2 ? nCoef[e] : 2);      // INV:103: nAngles = max(nMoments, 2)
  }
  const size_t nC = (size_t)coefOff[nE], nN = (size_t)nodeOff[nE];
  if (nC > 0 && !coefs) FAIL(h, "mcb_build_inverse_table_legendre: bad arguments");
  const int c = comp - 1;
  // scratch layout: offsets | coefficients | mus | values | cdf
  const size_t bOff = (sizeof(int) * off.size() + 15) & ~(size_t)15, bC = (sizeof(float) * (nC ? nC : 1) + 15) & ~(size_t)15,
               bN = sizeof(float) * nN;
  if (reserve(h, &h->dScratch, bOff + bC + 3 * bN)) return 1;
  char *base = (char *)h->dScratch;
  CK(h, cudaMemcpyAsync(base, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, h->stream));
  if (nC) CK(h, cudaMemcpyAsync(base + bOff, coefs, sizeof(float) * nC, cudaMemcpyHostToDevice, h->stream));
  float *mus = (float *)(base + bOff + bC), *values = mus + nN, *cdf = values + nN;
  const int *dCoefOff = (const int *)base, *dNodeOff = dCoefOff + nE + 1;
  mcb_launch_lobatto_inputs(dCoefOff, (const float *)(base + bOff), dNodeOff, nE, (long long)nN, mus, values, h->numSMs, h->stream);
  if (reserve(h, &h->dInv[c], sizeof(float) * (size_t)nS * nE)) return 1;
  mcb_launch_inverse_table(dNodeOff, mus, values, nE, nS, (float *)h->dInv[c], cdf, h->stream);
  CK(h, cudaGetLastError());
  if (settle(h)) return 1;
  h->P.inv[c] = (const float *)h->dInv[c]; h->P.invS[c] = nS; h->invE[c] = nE; h->P.invE[c] = nE; h->haveInv[c] = true;
  return 0;
}

// tabulateForwardPhaseFunctions OPT:1872-1934 on the device.  Entry e is stored as Legendre moments (nAngles == NULL or
// nAngles[e] == 0: chi_1.. = nCoef[e] values of coefs) or as nAngles[e] angle / value pairs (SPF:499-527); entries are
// concatenated in coefs, and in angles / values.  hybridWidthDeg > 0 builds the table with the Gaussian forward peak
// (computeHybridPhaseFunctions OPT:1936-2050) next to the original one; otherwise the two are the same.
int mcb_build_forward_table_general(mcb_handle *h, int comp, int nS, int nE, const int32_t *nCoef, const float *coefs,
                                    const int32_t *nAngles, const float *angles, const float *values, float hybridWidthDeg) {
  if (!h) return 1;
  if (!h->haveOptics) FAIL(h, "mcb_build_forward_table: call mcb_set_optics first");
  if (comp < 1 || comp > h->P.nc || nS < 2 || nE < 1 || !nCoef) FAIL(h, "mcb_build_forward_table: bad arguments");
  std::vector<int> off(2 * (nE + 1), 0);                              // coefficient offsets | angle offsets
  int *coefOff = off.data(), *angOff = off.data() + nE + 1;
  for (int e = 0; e < nE; ++e) {
    const int na = nAngles ? nAngles[e] : 0;
    if (nCoef[e] < 0 || na < 0 || na == 1 || (na > 0 && nCoef[e] > 0)) FAIL(h, "mcb_build_forward_table: bad arguments");
    coefOff[e + 1] = coefOff[e] + nCoef[e];
    angOff[e + 1] = angOff[e] + na;
  }
  const size_t nC = (size_t)coefOff[nE], nA = (size_t)angOff[nE];
  if ((nC > 0 && !coefs) || (nA > 0 && (!angles || !values))) FAIL(h, "mcb_build_forward_table: bad arguments");
  const int c = comp - 1;
  // scratch layout: offsets | coefficients | angles | values | 2 nS floats (hybrid: cosines, Gaussian)
  const size_t bOff = (sizeof(int) * off.size() + 15) & ~(size_t)15, bC = (sizeof(float) * (nC ? nC : 1) + 15) & ~(size_t)15,
               bA = (sizeof(float) * (nA ? nA : 1) + 15) & ~(size_t)15;
  if (reserve(h, &h->dScratch, bOff + bC + 2 * bA + sizeof(float) * 2 * (size_t)nS)) return 1;
  char *base = (char *)h->dScratch;
  CK(h, cudaMemcpyAsync(base, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, h->stream));
  if (nC) CK(h, cudaMemcpyAsync(base + bOff, coefs, sizeof(float) * nC, cudaMemcpyHostToDevice, h->stream));
  if (nA) {
    CK(h, cudaMemcpyAsync(base + bOff + bC, angles, sizeof(float) * nA, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(base + bOff + bC + bA, values, sizeof(float) * nA, cudaMemcpyHostToDevice, h->stream));
  }
  if (reserve(h, &h->dFwd[c], sizeof(float) * (size_t)nS * nE)) return 1;
  if (reserve(h, &h->dFwdOrig[c], sizeof(float) * (size_t)nS * nE)) return 1;
  mcb_launch_forward_tables((const int *)base, (const float *)(base + bOff), (const int *)base + nE + 1,
                            (const float *)(base + bOff + bC), (const float *)(base + bOff + bC + bA), nE, nS, hybridWidthDeg,
                            (float *)h->dFwdOrig[c], (float *)h->dFwd[c], (float *)(base + bOff + bC + 2 * bA), h->numSMs, h->stream);
  CK(h, cudaGetLastError());
  if (settle(h)) return 1;
  h->P.fwd[c] = (const float *)h->dFwd[c]; h->P.fwdOrig[c] = (const float *)h->dFwdOrig[c];
  h->P.fwdS[c] = nS; h->fwdE[c] = nE; h->P.fwdE[c] = nE; h->haveFwd[c] = true;
  h->P.fwdInvDTheta[c] = (float)(nS - 1) / 3.14159265358979312f;
  return 0;
}

// the Legendre-only form of round 1 (same tables, no hybrid peak)
int mcb_build_forward_table(mcb_handle *h, int comp, int nS, int nE, const int32_t *nCoef, const float *coefs) {
  return mcb_build_forward_table_general(h, comp, nS, nE, nCoef, coefs, nullptr, nullptr, nullptr, 0.0f);
}

int mcb_get_forward_table(mcb_handle *h, int comp, float *T, int64_t nFloats) {
  if (!h || !T) return 1;
  if (comp < 1 || comp > h->P.nc || !h->haveFwd[comp - 1]) FAIL(h, "mcb_get_forward_table: no table for this component");
  const int c = comp - 1;
  const int64_t n = (int64_t)h->P.fwdS[c] * h->fwdE[c];
  if (nFloats < n) FAIL(h, "mcb_get_forward_table: buffer too small (%lld needed)", (long long)n);
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaMemcpyAsync(T, h->dFwd[c], sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  return settle(h);
}

int mcb_get_inverse_table(mcb_handle *h, int comp, float *T, int64_t nFloats) {
  if (!h || !T) return 1;
  if (comp < 1 || comp > h->P.nc || !h->haveInv[comp - 1]) FAIL(h, "mcb_get_inverse_table: no table for this component");
  const int c = comp - 1;
  const int64_t n = (int64_t)h->P.invS[c] * h->invE[c];
  if (nFloats < n) FAIL(h, "mcb_get_inverse_table: buffer too small (%lld needed)", (long long)n);
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaMemcpyAsync(T, h->dInv[c], sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  return settle(h);
}

int mcb_set_forward_table(mcb_handle *h, int comp, int nS, int nE, const float *Pf, const float *Porig) {
  if (!h) return 1;
  if (!h->haveOptics) FAIL(h, "mcb_set_forward_table: call mcb_set_optics first");
  if (comp < 1 || comp > h->P.nc || nS < 2 || nE < 1 || !Pf) FAIL(h, "mcb_set_forward_table: bad arguments");
  const int c = comp - 1;
  if (stage(h, &h->dFwd[c], Pf, sizeof(float) * (size_t)nS * nE)) return 1;
  if (stage(h, &h->dFwdOrig[c], Porig ? Porig : Pf, sizeof(float) * (size_t)nS * nE)) return 1;
  h->P.fwd[c] = (const float *)h->dFwd[c]; h->P.fwdOrig[c] = (const float *)h->dFwdOrig[c];
  h->P.fwdS[c] = nS; h->fwdE[c] = nE; h->P.fwdE[c] = nE; h->haveFwd[c] = true;
  h->P.fwdInvDTheta[c] = (float)(nS - 1) / 3.14159265358979312f;
  return 0;
}

int mcb_set_views(mcb_handle *h, int nDir, const float *dirCos) {
  if (!h) return 1;
  if (nDir < 0 || nDir > MCB_MAX_DIR) FAIL(h, "specifyParameters: at most %d intensity directions", MCB_MAX_DIR);
  if (nDir > 0 && !dirCos) FAIL(h, "specifyParameters: Can't compute intensity without specifying directions.");
  for (int i = 0; i < nDir; ++i)
    if (std::fabs(dirCos[3 * i + 2]) < 1.17549435e-38f)
      FAIL(h, "specifyParameters: intensityMus can't be 0 (directly sideways)");
  h->P.nDir = nDir;
  for (int i = 0; i < 3 * nDir; ++i) h->P.viewDir[i] = dirCos[i];
  for (int i = 0; i < nDir; ++i) h->P.viewNorm[i] = 1.0f / (4.0f * 3.14159265358979312f * std::fabs(dirCos[3 * i + 2]));
  return 0;
}

int mcb_set_options(mcb_handle *h, const mcb_options *o) {
  if (!h || !o) return 1;
  if (o->zetaMin < 0.0f) FAIL(h, "specifyParameters: zetaMin must be >= 0.");
  if (o->arithmetic != MCB_ARITH_FAST && o->arithmetic != MCB_ARITH_REFERENCE) FAIL(h, "mcb_set_options: unknown arithmetic mode");
  h->P.opt = *o;
  if (h->haveOptics && o->tuneExtMask != h->maskKnob) {          // the knob changed after staging: pack again
    CK(h, cudaSetDevice(h->device));
    if (setup_packed_field(h) || setup_compact_storage(h) || settle(h)) return 1;
  }
  return 0;
}

int mcb_set_solar_source(mcb_handle *h, float solarMu, float solarAzimuthDeg) {   // ILL:62-101
  if (!h) return 1;
  if (solarAzimuthDeg < 0.0f || solarAzimuthDeg > 360.0f) FAIL(h, "setIllumination: solarAzimuth out of bounds");
  if (std::fabs(solarMu) > 1.0f || std::fabs(solarMu) <= 1.17549435e-38f) FAIL(h, "setIllumination: solarMu out of bounds");
  h->P.source = 0;
  h->P.solarMu = -std::fabs(solarMu);                                             // ILL:95
  const float pi32 = (float)std::acos(-1.0);
  volatile float t = solarAzimuthDeg * pi32;                                      // ILL:96, single precision
  h->P.solarPhi = t / 180.0f;
  {                                                                               // INT:1876-1894
    const float mu = h->P.solarMu, st = std::sqrt(std::fmax(1.0f - mu * mu, 0.0f));
    h->P.solarDir[0] = st * std::cos(h->P.solarPhi); h->P.solarDir[1] = st * std::sin(h->P.solarPhi); h->P.solarDir[2] = mu;
  }
  h->haveSource = true;
  return 0;
}

int mcb_set_thermal_source(mcb_handle *h, double fracAtmsPower, const double *voxelCDF) {  // ILL:431-522
  if (!h) return 1;
  if (!h->haveGrid) FAIL(h, "mcb_set_thermal_source: call mcb_set_grid first");
  if (!voxelCDF) FAIL(h, "mcb_set_thermal_source: null CDF");
  const size_t cells = (size_t)h->P.nx * h->P.ny * h->P.nz;
  if (stage(h, &h->dVoxelCDF, voxelCDF, sizeof(double) * cells)) return 1;
  if (finish_thermal_source(h, fracAtmsPower)) return 1;
  h->haveSource = true;
  return 0;
}

// emission_weightingNEW EMI:424-550 on the staged optics: Planck emission per cell and the prefix sum
// over cells run on the device (mcb_stage.cu); only the total comes back for the power bookkeeping.
int mcb_build_thermal_source(mcb_handle *h, const double *temps, double lambda_um,
                             double surfaceTemp, double *fracAtmsPowerOut, double *totalFluxOut) {
  if (!h) return 1;
  if (!h->haveOptics) FAIL(h, "emission_weighting: domain hasn't been initialized.");
  if (!temps && !h->haveTemps) FAIL(h, "emission_weighting: null temperature array");
  const DevDomain &P = h->P;
  const int nx = P.nx, ny = P.ny;
  const long long cells = (long long)nx * ny * P.nz;
  const double hP = 6.62606957e-34, cL = 2.99792458e+8, kB = 1.3806488e-23;
  const double a = 2.0 * hP * (cL * cL);
  const double Pi = 4.0 * std::atan(1.0);
  const double emiss = 1.0 - P.albedo;
  const double lambda = lambda_um / 1.0e6;
  const double b = hP * cL / (kB * lambda);
  const double lambda5 = std::pow(lambda, 5.0);
  const double areaX = h->xE[nx] - h->xE[0], areaY = h->yE[ny] - h->yE[0];
  double sfcPower = 0.0;
  if (!(emiss == 0.0 || surfaceTemp == 0.0)) {
    const double sfcPlanckRad = (a / (lambda5 * (std::exp(b / surfaceTemp) - 1.0))) / 1.0e6;
    sfcPower = Pi * emiss * sfcPlanckRad * areaX * areaY * (1000.0 * 1000.0);
  }
  const long long tiles = mcb_emission_tiles(cells);
  // temps == NULL: the temperatures staged by the previous call (they do not depend on the wavelength, DRV:326-345)
  if (temps) { if (stage_async(h, &h->dTemps, temps, sizeof(double) * cells)) return 1; h->haveTemps = true; }
  if (reserve(h, &h->dScratch, sizeof(double) * 2 * (size_t)(tiles + 1))) return 1;
  if (reserve(h, &h->dVoxelCDF, sizeof(double) * cells)) return 1;
  CK(h, cudaMemsetAsync(h->dFlags, 0, sizeof(int) * 4, h->stream));
  mcb_launch_emission_cdf(P, (const double *)h->dTemps, a, b, lambda5, h->dScratch, (double *)h->dVoxelCDF, h->dFlags,
                          h->numSMs, h->stream);
  CK(h, cudaGetLastError());
  int flags = 0;
  double total[2] = {0.0, 0.0};
  CK(h, cudaMemcpyAsync(&flags, h->dFlags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(total, (const double *)h->dScratch + 2 * tiles, sizeof(total), cudaMemcpyDeviceToHost, h->stream));
  if (settle(h)) return 1;
  double last = total[0];
  if (flags & 8) {                                   // a temperature <= 0: the atmosphere does not emit (EMI:498)
    last = 0.0;
    CK(h, cudaMemsetAsync(h->dVoxelCDF, 0, sizeof(double) * cells, h->stream));
  }
  double atmsPower = 0.0, frac = 0.0;
  if (last > 0.0) {                                                                       // EMI:512-521
    atmsPower = last * areaX * areaY * (1000.0 * 1000.0) / (double)(nx * ny);
    mcb_launch_emission_normalise((double *)h->dVoxelCDF, cells, (const double *)h->dScratch + 2 * tiles, h->numSMs, h->stream);
    CK(h, cudaGetLastError());
    frac = atmsPower / (atmsPower + sfcPower);
  }
  if (settle(h)) return 1;
  if (atmsPower + sfcPower == 0.0)
    FAIL(h, "emission_weightingNEW: Neither surface nor atmosphere will emitt photons since total power is 0. Not a valid solution");
  if (fracAtmsPowerOut) *fracAtmsPowerOut = frac;
  if (totalFluxOut) *totalFluxOut = (atmsPower + sfcPower) / (areaX * areaY * (1000.0 * 1000.0));
  if (finish_thermal_source(h, frac)) return 1;
  h->haveSource = true;
  return 0;
}

// the emission CDF currently staged (voxelWeights of type(Weights), EMI:56-57)
int mcb_get_thermal_source(mcb_handle *h, double *fracAtmsPower, double *voxelCDF, int64_t nDoubles) {
  if (!h) return 1;
  if (!h->haveSource || h->P.source != 1) FAIL(h, "mcb_get_thermal_source: no thermal source is staged");
  const long long cells = (long long)h->P.nx * h->P.ny * h->P.nz;
  if (fracAtmsPower) *fracAtmsPower = h->P.fracAtmsPower;
  if (voxelCDF) {
    if (nDoubles < cells) FAIL(h, "mcb_get_thermal_source: buffer too small (%lld needed)", cells);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaMemcpyAsync(voxelCDF, h->dVoxelCDF, sizeof(double) * cells, cudaMemcpyDeviceToHost, h->stream));
    if (settle(h)) return 1;
  }
  return 0;
}

static int ensure_tallies(mcb_handle *h) {
  DevDomain &P = h->P;
  const long long cols = (long long)P.nx * P.ny, cells = cols * P.nz;
  P.offFluxUp = 0; P.offFluxDown = cols; P.offFluxAbs = 2 * cols; P.offVolAbs = 3 * cols;
  P.offInt = P.offVolAbs + cells;
  P.offIntByComp = P.offInt + cols * P.nDir;
  P.offExcess = P.offIntByComp + cols * P.nDir * (P.nc + 1);
  P.offPhotons = P.offExcess + (long long)P.nDir * (P.nc + 1);
  const long long need = P.offPhotons + 1;
  if (need != h->nTally || !h->dTally) {
    // grow-only (reserve keeps the allocation while it is large enough): callers alias this buffer for the
    // cross-rank reduce (mcb_tally_buffer), so it must not move when views or components are switched off and on
    if (reserve(h, (void **)&h->dTally, sizeof(double) * (size_t)need)) return 1;
    CK(h, cudaMemsetAsync(h->dTally, 0, sizeof(double) * need, h->stream));
    h->nTally = need;
  }
  P.tally = h->dTally;
  return 0;
}

static int check_ready(mcb_handle *h) {
  if (!h->haveGrid || !h->haveOptics) FAIL(h, "computeRadiativeTransfer: problem not completely specified.");
  if (!h->haveSource) FAIL(h, "getNextPhoton: photons have not been initialized.");
  for (int c = 0; c < h->P.nc; ++c) {
    if (!h->haveInv[c]) FAIL(h, "computeRadiativeTransfer: no inverse phase function table for component %d", c + 1);
    if (h->P.nDir > 0 && !h->haveFwd[c]) FAIL(h, "computeRadiativeTransfer: no forward phase function table for component %d", c + 1);
  }
  // the local-estimation queue of the throughput kernel packs cell indices into 16-bit fields (mcb_fast.cu)
  if (h->P.nDir > 0 && (h->P.nx > 65535 || h->P.ny > 65535 || h->P.nz > 65535))
    FAIL(h, "computeRadiativeTransfer: intensity calculations support at most 65535 cells per axis");
  return 0;
}

static int run(mcb_handle *h, long long nPhotons, uint64_t seed, uint64_t firstPhotonId, bool zero, int64_t *nProcessed,
               bool zeroCounters = true, bool timeIt = true) {
  if (!h) return 1;
  if (nProcessed) *nProcessed = 0;
  if (check_ready(h)) return 1;
  if (nPhotons < 0) FAIL(h, "setIllumination: must ask for non-negative number of photons.");
  CK(h, cudaSetDevice(h->device));
  if (ensure_tallies(h)) return 1;
  DevDomain &P = h->P;
  if (zero) {                                                    // INT:247-272
    CK(h, cudaMemsetAsync(h->dTally, 0, sizeof(double) * h->nTally, h->stream));
    if (zeroCounters) CK(h, cudaMemsetAsync(h->dCounters, 0, sizeof(unsigned long long) * 32, h->stream));
  }
  if (!(zero && zeroCounters)) CK(h, cudaMemsetAsync(h->dCounters + CNT_N, 0, sizeof(unsigned long long), h->stream));
  if (nPhotons == 0) FAIL(h, "computeRadiativeTransfer: Didn't process any photons.");   // INT:835-836
  if (timeIt) CK(h, cudaEventRecord(h->evStart, h->stream));
  if (!P.opt.useRayTracing && !(P.maxExtinction > 0.0f))
    FAIL(h, "computeRadiativeTransfer: maximum cross-section needs a domain with extinction > 0");
  // the maximum cross-section branch (INT:564-571) exists in reference arithmetic only: as shipped it never
  // refreshes the photon's cell indices after a move, which is reproduced verbatim and is not a throughput path
  if (P.opt.arithmetic == MCB_ARITH_REFERENCE || !P.opt.useRayTracing)
    mcb_launch_reference_batch(P, nPhotons, seed, firstPhotonId, h->numSMs, h->stream);
  else if (pack_field(h, mcb_fast_reads_bricks(P)))            // no-op unless views were switched on/off since staging
    return 1;
  else
    mcb_launch_fast_batch(P, nPhotons, seed, firstPhotonId, h->numSMs, h->dCounters + CNT_N, h->stream);
  CK(h, cudaGetLastError());
  if (timeIt) { CK(h, cudaEventRecord(h->evStop, h->stream)); h->timed = true; }
  if (nProcessed) *nProcessed = nPhotons;
  return 0;
}

int mcb_run_batch(mcb_handle *h, int64_t nPhotons, uint64_t seed, uint64_t firstPhotonId, int64_t *nProcessed) {
  return run(h, nPhotons, seed, firstPhotonId, true, nProcessed);
}
int mcb_accumulate_batch(mcb_handle *h, int64_t nPhotons, uint64_t seed, uint64_t firstPhotonId, int64_t *nProcessed) {
  return run(h, nPhotons, seed, firstPhotonId, false, nProcessed);
}

// getFrequencyDistr EMI:552-573: how many of totalPhotons photons fall into each wavelength bin of the flux CDF
int mcb_frequency_distribution(mcb_handle *h, int nLambda, const double *cdf, int64_t totalPhotons, uint64_t seed,
                               int64_t *distribution) {
  if (!h) return 1;
  if (nLambda < 1 || nLambda > 8192 || !cdf || !distribution || totalPhotons < 0)
    FAIL(h, "getFrequencyDistr: bad arguments (1..8192 wavelength bins)");
  if (totalPhotons / 4 / ((long long)h->numSMs * 8) >= (1LL << 30)) FAIL(h, "getFrequencyDistr: too many photons for one call");
  for (int i = 1; i < nLambda; ++i) if (!(cdf[i] >= cdf[i - 1])) FAIL(h, "getFrequencyDistr: CDF must be non-decreasing");
  CK(h, cudaSetDevice(h->device));
  const size_t cdfBytes = sizeof(double) * nLambda;
  if (reserve(h, &h->dScratch, cdfBytes + sizeof(unsigned long long) * nLambda)) return 1;
  unsigned long long *dCounts = (unsigned long long *)((char *)h->dScratch + cdfBytes);
  CK(h, cudaMemcpyAsync(h->dScratch, cdf, cdfBytes, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemsetAsync(dCounts, 0, sizeof(unsigned long long) * nLambda, h->stream));
  if (totalPhotons > 0)
    mcb_launch_frequency_distribution((const double *)h->dScratch, nLambda, totalPhotons, seed, dCounts, h->numSMs, h->stream);
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(distribution, dCounts, sizeof(int64_t) * nLambda, cudaMemcpyDeviceToHost, h->stream));
  return settle(h);
}

// ---- the driver's batch loop with its statistics on the device (DRV:949-1052, 1188-1228) ----
static int ensure_stats(mcb_handle *h) {
  if (ensure_tallies(h)) return 1;
  const long long n = mcb_stats_elements(h->P);
  if (n != h->nStats || !h->dStats) {
    if (reserve(h, &h->dStats, sizeof(double) * (size_t)(2 * n + 2))) return 1;
    if (reserve(h, &h->dStatsOut, sizeof(double) * (size_t)(2 * n))) return 1;
    CK(h, cudaMemsetAsync(h->dStats, 0, sizeof(double) * (size_t)(2 * n + 2), h->stream));
    h->nStats = n;
  }
  return 0;
}

int mcb_stats_reset(mcb_handle *h) {
  if (!h) return 1;
  if (!h->haveGrid || !h->haveOptics) FAIL(h, "mcb_stats_reset: problem not completely specified.");
  CK(h, cudaSetDevice(h->device));
  if (ensure_stats(h)) return 1;
  CK(h, cudaMemsetAsync(h->dStats, 0, sizeof(double) * (size_t)(2 * h->nStats + 2), h->stream));
  return 0;
}

int mcb_run_batches(mcb_handle *h, int64_t numBatches, int64_t photonsPerBatch, uint64_t seed, uint64_t firstPhotonId,
                    int64_t *nProcessed) {
  if (!h) return 1;
  if (nProcessed) *nProcessed = 0;
  if (check_ready(h)) return 1;
  if (numBatches < 1 || photonsPerBatch < 1) FAIL(h, "mcb_run_batches: numBatches and photonsPerBatch must be positive");
  CK(h, cudaSetDevice(h->device));
  if (ensure_stats(h)) return 1;
  if (reserve(h, &h->dResults, sizeof(float) * (size_t)h->P.offExcess)) return 1;
  CK(h, cudaMemsetAsync(h->dCounters, 0, sizeof(unsigned long long) * 32, h->stream));
  CK(h, cudaEventRecord(h->evStart, h->stream));
  for (int64_t b = 0; b < numBatches; ++b) {
    // one batch = computeRadiativeTransfer + reportResults + the moment updates of DRV:1023-1052; nothing
    // leaves the device and the host does not wait between batches
    if (run(h, photonsPerBatch, seed, firstPhotonId + (uint64_t)b * (uint64_t)photonsPerBatch, true, nullptr, false, false))
      return 1;
    mcb_launch_normalise(h->P, (float)photonsPerBatch, (float *)h->dResults, h->numSMs, h->stream);
    mcb_launch_stats_accumulate(h->P, (const float *)h->dResults, (double *)h->dStats, (double)photonsPerBatch, h->numSMs,
                                h->stream);
    CK(h, cudaGetLastError());
  }
  CK(h, cudaEventRecord(h->evStop, h->stream));
  h->timed = true;
  if (nProcessed) *nProcessed = numBatches * photonsPerBatch;
  return 0;
}

int mcb_stats_buffer(mcb_handle *h, void **devicePtr, int64_t *nDoubles) {
  if (!h) return 1;
  if (!h->haveGrid || !h->haveOptics) FAIL(h, "mcb_stats_buffer: problem not completely specified.");
  CK(h, cudaSetDevice(h->device));
  if (ensure_stats(h)) return 1;
  if (devicePtr) *devicePtr = h->dStats;
  if (nDoubles) *nDoubles = 2 * h->nStats + 2;
  return 0;
}

int mcb_get_statistics(mcb_handle *h, double solarFlux, double *meanFluxStats, double *fluxUpStats, double *fluxDownStats,
                       double *fluxAbsorbedStats, double *absorbedProfileStats, double *absorbedVolumeStats,
                       double *radianceStats, int64_t *totalNumPhotons, int64_t *batchesCompleted) {
  if (!h) return 1;
  CK(h, cudaSetDevice(h->device));
  if (!h->dStats || h->nStats == 0) FAIL(h, "mcb_get_statistics: no batches have run");
  const DevDomain &P = h->P;
  const long long n = h->nStats;
  if (radianceStats && P.nDir == 0) FAIL(h, "reportResults: intensity information not available");
  double book[2] = {0.0, 0.0};
  CK(h, cudaMemcpyAsync(book, (const double *)h->dStats + 2 * n, sizeof(book), cudaMemcpyDeviceToHost, h->stream));
  if (settle(h)) return 1;
  if (totalNumPhotons) *totalNumPhotons = (int64_t)book[0];
  if (batchesCompleted) *batchesCompleted = (int64_t)book[1];
  if (!(book[1] >= 2.0)) FAIL(h, "mcb_get_statistics: at least two batches are needed for a standard error");
  double *out = (double *)h->dStatsOut;
  mcb_launch_stats_finalise((const double *)h->dStats, n, solarFlux, out, h->numSMs, h->stream);
  CK(h, cudaGetLastError());
  const size_t cols = (size_t)P.nx * P.ny, cells = cols * P.nz;
  // every output is stats(..., 1:2) in the Fortran layout: the block of means, then the block of standard errors
  struct { double *dst; long long off; size_t cnt; } parts[] = {
      {meanFluxStats, 0, 3}, {fluxUpStats, 3, cols}, {fluxDownStats, 3 + (long long)cols, cols},
      {fluxAbsorbedStats, 3 + 2 * (long long)cols, cols}, {absorbedProfileStats, 3 + 3 * (long long)cols, (size_t)P.nz},
      {absorbedVolumeStats, 3 + 3 * (long long)cols + P.nz, cells},
      {radianceStats, 3 + 3 * (long long)cols + P.nz + (long long)cells, cols * (size_t)P.nDir}};
  for (auto &q : parts)
    if (q.dst && q.cnt) {
      CK(h, cudaMemcpyAsync(q.dst, out + q.off, sizeof(double) * q.cnt, cudaMemcpyDeviceToHost, h->stream));
      CK(h, cudaMemcpyAsync(q.dst + q.cnt, out + n + q.off, sizeof(double) * q.cnt, cudaMemcpyDeviceToHost, h->stream));
    }
  return settle(h);
}

int mcb_last_batch_ms(mcb_handle *h, float *ms) {
  if (!h || !ms) return 1;
  if (!h->timed) FAIL(h, "mcb_last_batch_ms: no batch has run");
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaEventSynchronize(h->evStop));
  CK(h, cudaEventElapsedTime(ms, h->evStart, h->evStop));
  return 0;
}

int mcb_get_counters(mcb_handle *h, mcb_counters *c) {
  if (!h || !c) return 1;
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaStreamSynchronize(h->stream));
  unsigned long long v[CNT_END];
  CK(h, cudaMemcpy(v, h->dCounters, sizeof(v), cudaMemcpyDeviceToHost));
  memset(c, 0, sizeof(*c));
  c->photons = (int64_t)v[CNT_PHOTONS]; c->crossings = (int64_t)v[CNT_CROSSINGS];
  c->scatters = (int64_t)v[CNT_SCATTERS]; c->surfaceHits = (int64_t)v[CNT_SURFACE];
  c->topExits = (int64_t)v[CNT_TOP]; c->bad = (int64_t)v[CNT_BAD];
  c->leRays = (int64_t)v[CNT_LE_RAYS]; c->leCrossings = (int64_t)v[CNT_LE_CROSSINGS];
  c->rouletteKills = (int64_t)v[CNT_RR_KILLS];
  c->surfaceKills = (int64_t)v[CNT_SURFACE_KILLS];                     // photons absorbed by the surface (weight <= tiny)
  c->leaps = (int64_t)v[CNT_LEAPS]; c->leapCells = (int64_t)v[CNT_LEAP_CELLS];
  // the fast kernel does not count top exits while marching: every photon ends at the top, at the
  // surface (weight <= tiny), by roulette, or is dropped as bad
  if (h->P.opt.arithmetic == MCB_ARITH_FAST)
    c->topExits = c->photons - c->rouletteKills - c->surfaceKills - c->bad;
  return 0;
}

int mcb_tally_buffer(mcb_handle *h, void **devicePtr, int64_t *nDoubles) {
  if (!h) return 1;
  if (!h->haveGrid || !h->haveOptics) FAIL(h, "mcb_tally_buffer: problem not completely specified.");
  CK(h, cudaSetDevice(h->device));
  if (ensure_tallies(h)) return 1;
  if (devicePtr) *devicePtr = (void *)h->dTally;
  if (nDoubles) *nDoubles = h->nTally;
  return 0;
}

static int fetch_tallies(mcb_handle *h) {
  CK(h, cudaSetDevice(h->device));
  if (!h->dTally) FAIL(h, "reportResults: no results available");
  if (h->hTallyCap < h->nTally) {
    if (h->hTally) cudaFreeHost(h->hTally);
    h->hTally = nullptr;
    CK(h, cudaMallocHost((void **)&h->hTally, sizeof(double) * h->nTally));
    h->hTallyCap = h->nTally;
  }
  CK(h, cudaMemcpyAsync(h->hTally, h->dTally, sizeof(double) * h->nTally, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int mcb_get_raw_tallies(mcb_handle *h, double *out, int64_t nDoubles) {
  if (!h || !out) return 1;
  if (fetch_tallies(h)) return 1;
  if (nDoubles < h->nTally) FAIL(h, "mcb_get_raw_tallies: buffer too small (%lld needed)", h->nTally);
  memcpy(out, h->hTally, sizeof(double) * h->nTally);
  return 0;
}

int mcb_get_results(mcb_handle *h, int64_t nPhotonsNormalise,
                    float *fluxUp, float *fluxDown, float *fluxAbsorbed,
                    float *volumeAbsorption, float *intensity, float *intensityByComponent) {
  if (!h) return 1;
  CK(h, cudaSetDevice(h->device));
  if (!h->dTally) FAIL(h, "reportResults: no results available");
  const DevDomain &P = h->P;
  const int nDir = P.nDir, nc = P.nc;
  const size_t cols = (size_t)P.nx * P.ny, cells = cols * P.nz;
  if ((intensity || intensityByComponent) && nDir == 0) FAIL(h, "reportResults: intensity information not available");
  double numPhotonsProcessed = (double)nPhotonsNormalise;
  if (nPhotonsNormalise <= 0) {
    CK(h, cudaMemcpyAsync(&numPhotonsProcessed, h->dTally + P.offPhotons, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (settle(h)) return 1;
  }
  if (!(numPhotonsProcessed > 0)) FAIL(h, "computeRadiativeTransfer: Didn't process any photons.");
  // INT:294-388 on the device: the f64 tallies become the reference's single-precision, normalised arrays
  // in HBM; only the arrays asked for cross PCIe.
  if (reserve(h, &h->dResults, sizeof(float) * (size_t)P.offExcess)) return 1;
  float *out = (float *)h->dResults;
  mcb_launch_normalise(P, (float)numPhotonsProcessed, out, h->numSMs, h->stream);
  CK(h, cudaGetLastError());
  struct { float *dst; long long off; size_t n; } parts[] = {
      {fluxUp, P.offFluxUp, cols}, {fluxDown, P.offFluxDown, cols}, {fluxAbsorbed, P.offFluxAbs, cols},
      {volumeAbsorption, P.offVolAbs, cells}, {intensity, P.offInt, cols * nDir},
      {intensityByComponent, P.offIntByComp, cols * nDir * (size_t)(nc + 1)}};
  for (auto &q : parts)
    if (q.dst && q.n) CK(h, cudaMemcpyAsync(q.dst, out + q.off, sizeof(float) * q.n, cudaMemcpyDeviceToHost, h->stream));
  return settle(h);
}

int mcb_run_trace(mcb_handle *h, int64_t nPhotons, const float *rn, int64_t rnStride,
                  int32_t maxEventsPerPhoton, mcb_event *events, int64_t eventCap, int64_t *nEvents) {
  if (!h) return 1;
  if (nEvents) *nEvents = 0;
  if (check_ready(h)) return 1;
  if (nPhotons <= 0 || !rn || rnStride <= 0 || maxEventsPerPhoton <= 0 || !events) FAIL(h, "mcb_run_trace: bad arguments");
  CK(h, cudaSetDevice(h->device));
  if (ensure_tallies(h)) return 1;
  CK(h, cudaMemsetAsync(h->dTally, 0, sizeof(double) * h->nTally, h->stream));
  CK(h, cudaMemsetAsync(h->dCounters, 0, sizeof(unsigned long long) * 32, h->stream));
  struct DeviceScratch {                               // freed on every return path
    void *p = nullptr;
    ~DeviceScratch() { if (p) cudaFree(p); }
  } sRn, sEv, sCount;
  const size_t nRn = (size_t)nPhotons * rnStride, nEv = (size_t)nPhotons * maxEventsPerPhoton;
  CK(h, cudaMalloc(&sRn.p, sizeof(float) * nRn));
  CK(h, cudaMalloc(&sEv.p, sizeof(mcb_event) * nEv));
  CK(h, cudaMalloc(&sCount.p, sizeof(int) * nPhotons));
  float *dRn = (float *)sRn.p; mcb_event *dEv = (mcb_event *)sEv.p; int *dCount = (int *)sCount.p;
  CK(h, cudaMemcpyAsync(dRn, rn, sizeof(float) * nRn, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemsetAsync(dCount, 0, sizeof(int) * nPhotons, h->stream));
  mcb_launch_trace(h->P, nPhotons, dRn, rnStride, dEv, maxEventsPerPhoton, dCount, h->stream);
  CK(h, cudaGetLastError());
  std::vector<int> count((size_t)nPhotons);
  std::vector<mcb_event> ev(nEv);
  CK(h, cudaMemcpyAsync(count.data(), dCount, sizeof(int) * nPhotons, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(ev.data(), dEv, sizeof(mcb_event) * nEv, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  int64_t total = 0;
  for (int64_t p = 0; p < nPhotons; ++p) {
    const int n = count[p] < maxEventsPerPhoton ? count[p] : maxEventsPerPhoton;
    for (int k = 0; k < n; ++k) {
      if (total < eventCap) events[total] = ev[(size_t)p * maxEventsPerPhoton + k];
      total++;
    }
  }
  if (nEvents) *nEvents = total;
  {  // photons started, for mcb_get_results on traced batches
    const double np = (double)nPhotons;
    CK(h, cudaMemcpy(h->dTally + h->P.offPhotons, &np, sizeof(double), cudaMemcpyHostToDevice));
  }
  return 0;
}

// debugging aid used by the tests: the raw Philox4x32-10 stream of one photon
int mcb_debug_philox(mcb_handle *h, uint64_t seed, uint64_t photon, int n, uint32_t *out) {
  if (!h || !out || n <= 0) return 1;
  CK(h, cudaSetDevice(h->device));
  if (reserve(h, &h->dScratch, sizeof(uint32_t) * (size_t)n)) return 1;
  uint32_t *d = (uint32_t *)h->dScratch;
  mcb_launch_philox_kat(seed, photon, n, d, h->stream);
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(out, d, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// ---- multi-GPU: one process per GPU, the domain replicated, photons split by global id, ONE sum-reduce ----------
#define NCK(h, call)                                                                 \
  do {                                                                               \
    ncclResult_t _r = (call);                                                        \
    if (_r != ncclSuccess) FAIL(h, "%s: %s", #call, nccl_api()->GetErrorString(_r)); \
  } while (0)

int mcb_comm_unique_id(void *id128) {
  if (!id128) return 1;
  NcclApi *n = nccl_api();
  if (!n->lib) return 2;
  ncclUniqueId id;
  if (n->GetUniqueId(&id) != ncclSuccess) return 3;
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int mcb_comm_init(mcb_handle *h, int nranks, int rank, const void *id128) {
  if (!h) return 1;
  if (nranks < 1 || rank < 0 || rank >= nranks || !id128) FAIL(h, "initializeProcesses: bad rank / number of processes");
  NcclApi *n = nccl_api();
  if (!n->lib) FAIL(h, "initializeProcesses: %s", n->err.c_str());
  CK(h, cudaSetDevice(h->device));
  if (h->comm) { n->CommDestroy(h->comm); h->comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NCK(h, n->CommInitRank(&h->comm, nranks, id, rank));
  h->commRanks = nranks; h->commRank = rank;
  return 0;
}

int mcb_comm_info(mcb_handle *h, int *nranks, int *rank, int *ncclVersion) {
  if (!h) return 1;
  if (nranks) *nranks = h->comm ? h->commRanks : 1;
  if (rank) *rank = h->comm ? h->commRank : 0;
  if (ncclVersion) { *ncclVersion = 0; NcclApi *n = nccl_api(); if (n->lib) n->GetVersion(ncclVersion); }
  return 0;
}

int mcb_comm_destroy(mcb_handle *h) {
  if (!h) return 1;
  if (h->comm) {
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    nccl_api()->CommDestroy(h->comm);
    h->comm = nullptr; h->commRanks = 1; h->commRank = 0;
  }
  return 0;
}

// sum-reduce (root >= 0) or all-reduce (root < 0) of a device buffer of doubles, in place, on the handle's stream
static int reduce_doubles(mcb_handle *h, double *buf, long long n, int root, const char *what) {
  if (!h->comm) return 0;                              // a single process: nothing to sum (multipleProcesses_nompi)
  if (root >= h->commRanks) FAIL(h, "%s: root %d is not a rank of this communicator", what, root);
  NcclApi *a = nccl_api();
  CK(h, cudaSetDevice(h->device));
  if (root >= 0) NCK(h, a->Reduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, root, h->comm, h->stream));
  else NCK(h, a->AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, h->comm, h->stream));
  return 0;
}

int mcb_reduce_tallies(mcb_handle *h, int root) {
  if (!h) return 1;
  if (!h->haveGrid || !h->haveOptics) FAIL(h, "sumAcrossProcesses: problem not completely specified.");
  CK(h, cudaSetDevice(h->device));
  if (ensure_tallies(h)) return 1;
  return reduce_doubles(h, h->dTally, h->nTally, root, "sumAcrossProcesses");
}

int mcb_reduce_statistics(mcb_handle *h, int root) {
  if (!h) return 1;
  if (!h->haveGrid || !h->haveOptics) FAIL(h, "sumAcrossProcesses: problem not completely specified.");
  CK(h, cudaSetDevice(h->device));
  if (ensure_stats(h)) return 1;
  return reduce_doubles(h, (double *)h->dStats, 2 * h->nStats + 2, root, "sumAcrossProcesses");
}

// measured ceiling of divergent sector gathers (mcb_probe.cu): what the roofline of the L2-resident configurations
// is reported against
int mcb_debug_gather_probe(mcb_handle *h, int64_t bytes, int loadsInFlight, int blocksPerSM, int iterations,
                           double *gathersPerSecond) {
  if (!h || !gathersPerSecond) return 1;
  if (bytes < 4096 || blocksPerSM < 1 || blocksPerSM > 16 || iterations < 1) FAIL(h, "mcb_debug_gather_probe: bad arguments");
  CK(h, cudaSetDevice(h->device));
  const double r = mcb_run_gather_probe((size_t)bytes, loadsInFlight, blocksPerSM, iterations, h->numSMs, h->stream);
  if (r < 0.0) FAIL(h, "mcb_debug_gather_probe: failed (%g)", r);
  *gathersPerSecond = r;
  return 0;
}

// the vacuum-distance map of the staged domain (pack_field), for the tests of the leaping marcher
int mcb_debug_distance_map(mcb_handle *h, uint8_t *out, int64_t nBytes) {
  if (!h || !out) return 1;
  if (!h->haveGrid || !h->haveOptics) FAIL(h, "mcb_debug_distance_map: problem not completely specified.");
  if (!h->haveDist) FAIL(h, "mcb_debug_distance_map: this grid has no vacuum-distance map (narrow, irregular or bitmap-marched)");
  const int64_t cells = (int64_t)h->P.nx * h->P.ny * h->P.nz;
  if (nBytes < cells) FAIL(h, "mcb_debug_distance_map: buffer too small (%lld needed)", (long long)cells);
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaMemcpyAsync(out, h->dDist, (size_t)cells, cudaMemcpyDeviceToHost, h->stream));
  return settle(h);
}

}  // extern "C"
