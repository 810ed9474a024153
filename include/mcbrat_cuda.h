/*
 * mcbrat_cuda.h -- C ABI of the B200-native photon-tracing path of MCBRaT3D.
 *
 * Drop-in boundary: these entry points are what a thin ISO_C_BINDING layer inside the
 * reference's Fortran module procedures binds to (fortran/mcbrat_cuda_mod.f90 and
 * INTEGRATION.md show the binding).  Reference citations use
 *   INT = Integrators/monteCarloRadiativeTransfer.f95   OPT = src/opticalProperties.f95
 *   ILL = src/monteCarloIllumination.f95                EMI = src/emissionAndBroadBandWeights.f95
 *   DRV = Drivers/monteCarloDriver.f95
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; mcb_last_error() gives the
 *     text the Fortran shim hands to setStateToFailure(status, ...) (ErrorMessages.f95:225).
 *   - arrays use the reference's Fortran layout, x fastest: a(ix,iy,iz) is
 *     a[(ix-1) + nx*((iy-1) + ny*(iz-1))]; 4-D arrays add the component as slowest index;
 *     tables are values(step, entry), step fastest; indices are 1-based.
 *   - the library COPIES on every mcb_set_* (staged once into HBM) and never retains a
 *     host pointer; results are written into caller-provided buffers; any output pointer
 *     may be NULL (Fortran `optional`).
 *   - one handle per (process, GPU); calls on one handle are serialised by the caller.
 *   - plain pointers and sizes only: no torch / C++ types cross this boundary.
 */
#ifndef MCBRAT_CUDA_H
#define MCBRAT_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcb_handle mcb_handle;

enum { MCB_ARITH_FAST = 0, MCB_ARITH_REFERENCE = 1 };

/* Algorithmic choices: the optional arguments of specifyParameters (INT:1046-1073) that
 * the photon loop reads.  Defaults (mcb_default_options) are the integrator's (INT:53-96). */
typedef struct {
  int32_t useRayTracing;                    /* INT:53; 0 = maximum cross-section (INT:564-571) */
  int32_t useRussianRoulette;               /* INT:55                                          */
  float   russianRouletteW;                 /* INT:56                                          */
  int32_t useRussianRouletteForIntensity;   /* INT:90                                          */
  float   zetaMin;                          /* INT:91                                          */
  int32_t useHybridPhaseFunsForIntenCalcs;  /* INT:85                                          */
  int32_t numOrdersOrigPhaseFunIntenCalcs;  /* INT:87                                          */
  int32_t limitIntensityContributions;      /* INT:94                                          */
  float   maxIntensityContribution;         /* INT:95                                          */
  float   LW_flag;                          /* INT:68; > 0 switches on emission bookkeeping    */
  int32_t arithmetic;                       /* MCB_ARITH_FAST (default) | MCB_ARITH_REFERENCE  */
  /* Measurement knobs of the throughput kernels -- not part of the reference's interface.  0 = the
   * library's own choice (what every production call uses); tests and profiling scripts set them to
   * compare variants inside one process.  They never change results beyond f64 summation order.      */
  int32_t tuneKernel;                       /* MCB_KERNEL_PARK | MCB_KERNEL_POOL (flux-only, uniform grids)   */
  int32_t tuneLayout;                       /* MCB_LAYOUT_LINEAR | MCB_LAYOUT_BRICKS extinction field         */
  int32_t tuneBlocksPerSM;                  /* resident CTAs per SM of the flux kernels                       */
  int32_t tuneParkThreshold;                /* parked lanes that trigger an event phase (park kernel)         */
  int32_t tuneLeCarry;                      /* view rays parked between queue rounds: < 0 off, > 0 threshold  */
  int32_t tuneExtMask;                      /* treatment of fields too large for L2: < 0 off, > 0 on (default: on
                                             * above 48 MB); 1 = occupancy bitmap only, 2 = also the layer-cropped
                                             * field + compact event data of the pool flux kernel (what "on" means) */
  int32_t tuneBurst;                        /* cells per marching burst of the pool kernel (4 or 8)           */
  int32_t tuneLeap;                         /* vacuum leaps of the pool kernels: < 0 off, > 0 smallest distance */
  int32_t tuneLeapLanes;                    /* lanes of a warp that must want a leap for the warp to take one   */
  int32_t reserved[1];
} mcb_options;
enum { MCB_KERNEL_PARK = 1, MCB_KERNEL_POOL = 2 };
enum { MCB_LAYOUT_LINEAR = 1, MCB_LAYOUT_BRICKS = 2 };

/* Event counters of the last batch (algorithmic-bytes accounting). */
typedef struct {
  int64_t photons, crossings, scatters, surfaceHits, topExits, bad,
          leRays, leCrossings, rouletteKills, surfaceKills,
          leaps, leapCells,        /* vacuum / clear-sky leaps taken by the pool kernels; cells they crossed (also in crossings) */
          reserved[4];
} mcb_counters;

/* Trace record (fixed-random-number single-photon harness, north-star criterion (a)). */
enum {
  MCB_EV_BIRTH = 1, MCB_EV_SCATTER = 2, MCB_EV_SURFACE = 3, MCB_EV_EXIT_TOP = 4,
  MCB_EV_KILLED_SURFACE = 5, MCB_EV_KILLED_ROULETTE = 6, MCB_EV_BAD = 7,
  MCB_EV_LOCAL_ESTIMATE = 8, MCB_EV_RN_EXHAUSTED = 9, MCB_EV_NULL_COLLISION = 10
};
typedef struct {
  int32_t photon, kind, ix, iy, iz, component, phaseIndex, angleIndex, order, nrn;
  float   weight, tau;
  double  path, x, y, z;
  float   dir[3];
  int32_t pad;
} mcb_event;

/* ---- lifetime -------------------------------------------------------------------------- */
/* new_Integrator INT:129-132 / finalize_Integrator INT:1486 own the handle in the shim. */
int mcb_create(int device, mcb_handle **out);
int mcb_destroy(mcb_handle *h);
int mcb_last_error(const mcb_handle *h, char *buf, int len);
int mcb_version(void);
/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL = the handle's own. */
int mcb_set_stream(mcb_handle *h, void *cudaStream);
int mcb_synchronize(mcb_handle *h);

/* ---- staging: replaces the per-batch getInfo_Domain copies at INT:434-443, 1668-1673 --- */
/* grid edges: new_Integrator's getInfo_Domain(xPosition, yPosition, zPosition), INT:147-157 */
int mcb_set_grid(mcb_handle *h, int nx, int ny, int nz,
                 const double *xEdges, const double *yEdges, const double *zEdges);
/* getInfo_Domain(albedo, totalExt, cumExt, ssa, phaseFuncI), INT:441-443 */
int mcb_set_optics(mcb_handle *h, int nc, const double *totalExt, const double *cumExt,
                   const double *ssa, const int32_t *phaseIdx, double albedo);
/* ---- per-wavelength assembly on the device: read_SSPTable's inner loops (OPT:204-299) followed by
 * getOpticalPropertiesByComponent (OPT:1022-1061).  mcb_set_physical stages the wavelength-independent
 * state of type(commonDomain) once per run (OPT:63-75): massConc, Reff as (nPhys, nx, ny, nz) with the
 * component fastest, numConc = numConc(1,1,:) (nz values; NULL without gas components).  Per wavelength
 * mcb_assemble_optics takes what read_SSPTable reads for one lambdaIndex -- a few hundred bytes per
 * component -- and builds totalExt, cumExt, ssa, phaseFuncI and the packed copies in HBM; it replaces
 * mcb_set_optics for that wavelength.  setup != 0 leaves every phase-function index at 1 (OPT:278).     */
enum { MCB_COMP_VOLEXT = 0, MCB_COMP_ABSXSEC = 1, MCB_COMP_PROFILE = 2 };
typedef struct {
  int32_t kind;            /* extType "volExt" | "absXsec" (OPT:203) | explicit horizontally uniform profile  */
  int32_t physIndex;       /* volExt: 1-based slot of massConc / Reff ("comp-gasComp", OPT:264)               */
  int32_t nTable;          /* volExt: nReff; otherwise the number of levels                                   */
  int32_t zLevelBase;      /* 1-based (OPT:201)                                                               */
  const float *key;        /* volExt: phaseFunctionKeyT(nReff), default real                                   */
  const double *ext;       /* volExt: ExtinctionT(nReff); absXsec: xsec(nLevels); profile: extinction(nLevels) */
  const double *ssa;       /* volExt: SingleScatteringAlbedoT(nReff); profile: ssa(nLevels)                    */
  const int32_t *phaseIdx; /* profile: phaseFunctionIndex(nLevels) (e.g. calc_RayleighScattering, OPT:2052)    */
} mcb_component;
int mcb_set_physical(mcb_handle *h, int nPhys, const double *massConc, const double *Reff, const double *numConc);
int mcb_assemble_optics(mcb_handle *h, int nc, const mcb_component *comps, int setup, double albedo);
/* the dense arrays currently in HBM, in the domain's layout (any pointer may be NULL) */
int mcb_get_optics(mcb_handle *h, double *totalExt, double *cumExt, double *ssa, int32_t *phaseIdx);
/* getInfo_Domain(inversePhaseFuncs) INT:443 <- tabulateInversePhaseFunctions INT:280 */
int mcb_set_inverse_table(mcb_handle *h, int comp, int nS, int nE, const float *T);
/* computeInversePhaseFuncTable (INV:26-174) on the device instead of on the host: entry e of the component's
 * phase-function table is handed over at nAngles[e] points increasing in mu (the native angles reversed, or the
 * phase function at max(nMoments,2) Lobatto nodes, INV:87-112; mus / values concatenated over entries); the CDF,
 * the nS brackets and the analytic inversions run in HBM.  mcb_get_inverse_table reads a staged table back.   */
int mcb_build_inverse_table(mcb_handle *h, int comp, int nS, int nE, const int32_t *nAngles,
                            const float *mus, const float *values);
/* The same for a table whose entries are all stored as Legendre moments (entry e: nCoef[e] moments chi_1.., concatenated
 * in coefs), with nothing evaluated on the host: computeLobattoTerms NUM:27-114 (the max(nMoments,2) abscissas), the
 * phase function there (SPF:480-498, NUM:187-205) and the inversion all run in HBM.                              */
int mcb_build_inverse_table_legendre(mcb_handle *h, int comp, int nS, int nE, const int32_t *nCoef, const float *coefs);
int mcb_get_inverse_table(mcb_handle *h, int comp, float *T, int64_t nFloats);
/* tabulateForwardPhaseFunctions (OPT:1872-1934) on the device for tables stored as Legendre moments: entry e has
 * nCoef[e] moments chi_1.. (concatenated in coefs; 0 moments = isotropic); fills tabPhase and tabOrigPhase alike
 * (hybrid tables, OPT:1936-2050, and angle/value tables are staged with mcb_set_forward_table).                  */
int mcb_build_forward_table(mcb_handle *h, int comp, int nS, int nE, const int32_t *nCoef, const float *coefs);
/* tabulateForwardPhaseFunctions for any table: entry e is stored as Legendre moments (nAngles == NULL or nAngles[e] == 0)
 * or as nAngles[e] angle / value pairs, interpolated linearly in the cosine of the angle (SPF:499-527; concatenated in
 * angles / values).  hybridWidthDeg > 0: tabPhase gets the Gaussian forward peak of computeHybridPhaseFunctions
 * (OPT:1936-2050, hunt + bisection for the transition angle per entry), tabOrigPhase the original values.          */
int mcb_build_forward_table_general(mcb_handle *h, int comp, int nS, int nE, const int32_t *nCoef, const float *coefs,
                                    const int32_t *nAngles, const float *angles, const float *values, float hybridWidthDeg);
int mcb_get_forward_table(mcb_handle *h, int comp, float *T, int64_t nFloats);
/* getInfo_Domain(tabPhase, tabOrigPhase) INT:1672-1673 <- tabulateForwardPhaseFunctions INT:282 */
int mcb_set_forward_table(mcb_handle *h, int comp, int nS, int nE, const float *P, const float *Porig);
/* specifyParameters(intensityMus, intensityPhis): direction cosines as INT:1267-1269 builds them;
 * nDir = 0 switches intensity off (specifyParameters(computeIntensity=.false.), INT:1278-1284) */
int mcb_set_views(mcb_handle *h, int nDir, const float *dirCos);
void mcb_default_options(mcb_options *o);
int mcb_set_options(mcb_handle *h, const mcb_options *o);

/* ---- photon sources: replace new_PhotonStream + getNextPhoton (ILL:62-101, 431-522, 561-590)
 * The stream is never materialised; each photon samples its own start on the device.        */
int mcb_set_solar_source(mcb_handle *h, float solarMu, float solarAzimuthDeg);
/* Weights from emission_weighting (EMI:424-550): voxelCDF(nx,ny,nz), fracAtmsPower          */
int mcb_set_thermal_source(mcb_handle *h, double fracAtmsPower, const double *voxelCDF);
/* device build of the same CDF from the staged optics (EMI:498-522); temps(nx,ny,nz) in K;
 * temps == NULL reuses the temperatures of the previous call on this grid (many-wavelength runs) */
int mcb_build_thermal_source(mcb_handle *h, const double *temps, double lambda_um,
                             double surfaceTemp, double *fracAtmsPower, double *totalFlux);
/* read the staged Weights back (voxelWeights EMI:56-57, fracAtmsPower); either pointer may be NULL */
int mcb_get_thermal_source(mcb_handle *h, double *fracAtmsPower, double *voxelCDF, int64_t nDoubles);

/* getFrequencyDistr (EMI:552-573, called at DRV:439-445, 497-503): allocate totalPhotons photons to the
 * nLambda wavelength bins of a flux CDF -- one uniform draw + findCDFIndex per photon, on the device.
 * distribution(nLambda) receives the counts (they sum to totalPhotons).                                */
int mcb_frequency_distribution(mcb_handle *h, int nLambda, const double *cdf, int64_t totalPhotons,
                               uint64_t seed, int64_t *distribution);

/* ---- computeRadiativeTransfer (INT:209-218) ------------------------------------------- */
/* Zero the tallies (INT:247-272) and trace nPhotons photons with global ids
 * [firstPhotonId, firstPhotonId + nPhotons) of the stream `seed` (counter-based RNG: the
 * result does not depend on how a run is split into batches or GPUs).  Asynchronous on the
 * handle's stream; *nProcessed = photons started (INT:833, includes those later dropped).  */
int mcb_run_batch(mcb_handle *h, int64_t nPhotons, uint64_t seed, uint64_t firstPhotonId,
                  int64_t *nProcessed);
/* Same without zeroing: adds another range of photons to the current tallies. */
int mcb_accumulate_batch(mcb_handle *h, int64_t nPhotons, uint64_t seed, uint64_t firstPhotonId,
                         int64_t *nProcessed);
/* ---- the driver's batch loop and statistics (DRV:949-1052, 1188-1228) on the device --------
 * mcb_run_batches runs numBatches batches back to back (batch b traces photon ids firstPhotonId +
 * b*photonsPerBatch ...): after each batch the normalised results are folded into first and second
 * moments weighted by the photons of the batch (meanFlux*Stats, flux*Stats, absorbedProfileStats,
 * absorbedVolumeStats, RadianceStats of DRV:1023-1052).  Nothing returns to the host between batches.
 * Moments add across GPUs/ranks exactly as the driver's sumAcrossProcesses does (DRV:1151-1166):
 * mcb_stats_buffer exposes the f64 device buffer [moment1 : n][moment2 : n][totalNumPhotons]
 * [batchesCompleted] for one sum-reduce.  mcb_get_statistics finalises as DRV:1188-1228; every
 * output is stats(..., 1:2) in the Fortran layout (block of means, then block of standard errors);
 * meanFluxStats is (3,2): up, down, absorbed.  Any output pointer may be NULL.                     */
int mcb_stats_reset(mcb_handle *h);
int mcb_run_batches(mcb_handle *h, int64_t numBatches, int64_t photonsPerBatch, uint64_t seed,
                    uint64_t firstPhotonId, int64_t *nProcessed);
int mcb_stats_buffer(mcb_handle *h, void **devicePtr, int64_t *nDoubles);
int mcb_get_statistics(mcb_handle *h, double solarFlux, double *meanFluxStats, double *fluxUpStats,
                       double *fluxDownStats, double *fluxAbsorbedStats, double *absorbedProfileStats,
                       double *absorbedVolumeStats, double *radianceStats,
                       int64_t *totalNumPhotons, int64_t *batchesCompleted);

/* Device time of the last run/accumulate call (CUDA events on the launch stream), ms.    */
int mcb_last_batch_ms(mcb_handle *h, float *ms);
int mcb_get_counters(mcb_handle *h, mcb_counters *c);

/* ---- reportResults (INT:845-865) ---------------------------------------------------- */
/* Normalised exactly as computeRadiativeTransfer does (INT:294-388).  nPhotonsNormalise <= 0
 * uses the handle's own photon count; a multi-GPU run passes the global count after the
 * reduce.  Sizes: flux*(nx,ny); volumeAbsorption(nx,ny,nz); intensity(nx,ny,nDir);
 * intensityByComponent(nx,ny,nDir,0:nc).  Synchronises the stream.                       */
int mcb_get_results(mcb_handle *h, int64_t nPhotonsNormalise,
                    float *fluxUp, float *fluxDown, float *fluxAbsorbed,
                    float *volumeAbsorption, float *intensity, float *intensityByComponent);
/* The packed f64 tally buffer on the device: [fluxUp|fluxDown|fluxAbsorbed : 3*nx*ny]
 * [volumeAbsorption : nx*ny*nz][intensity : nx*ny*nDir][intensityByComponent : nx*ny*nDir*(nc+1)]
 * [intensityExcess : nDir*(nc+1)][photons started : 1].  This is what one NCCL/MPI sum-reduce
 * replaces sumAcrossProcesses with (DRV:1151-1166).                                      */
int mcb_tally_buffer(mcb_handle *h, void **devicePtr, int64_t *nDoubles);
int mcb_get_raw_tallies(mcb_handle *h, double *out, int64_t nDoubles);

/* ---- multi-GPU: multipleProcesses_mpi (MPIW:29-251) over NCCL ------------------------------
 * One process per GPU, the domain replicated in every GPU's HBM, photons split by global photon id
 * (firstPhotonId of mcb_run_batch / mcb_run_batches), and ONE sum-reduce of the packed buffer at the
 * end -- it replaces the nine sumAcrossProcesses calls at DRV:1151-1166 (MPI_REDUCE, MPIW:70-251).
 * libnccl is bound at run time (dlopen), so single-GPU hosts need no NCCL.
 *   mcb_comm_unique_id : rank 0 creates the 128-byte NCCL id; the HOST distributes it (the Fortran host
 *                        with one MPI_BCAST next to its MPI_INIT, MPIW:29-52; torchrun hosts with
 *                        torch.distributed; the C++ example through a file).
 *   mcb_comm_init      : initializeProcesses -- collective over all ranks.
 *   mcb_reduce_tallies / mcb_reduce_statistics : in place, asynchronous on the handle's stream;
 *                        root >= 0 reduces to that rank (MPI_REDUCE), root < 0 all-reduces.  No-ops on a
 *                        handle without a communicator (multipleProcesses_nompi.f95).
 *   mcb_comm_destroy   : finalizeProcesses (MPIW:62-68).                                               */
int mcb_comm_unique_id(void *id128);
int mcb_comm_init(mcb_handle *h, int nranks, int rank, const void *id128);
int mcb_comm_info(mcb_handle *h, int *nranks, int *rank, int *ncclVersion);
int mcb_reduce_tallies(mcb_handle *h, int root);
int mcb_reduce_statistics(mcb_handle *h, int root);
int mcb_comm_destroy(mcb_handle *h);

/* ---- trace harness ------------------------------------------------------------------ */
/* Photon p is born from and transported with rn[p*rnStride ...] (source draws first, then the
 * computeRT order).  Always reference arithmetic.  Events of photon 0 come first,
 * then photon 1, ...; at most maxEventsPerPhoton per photon, eventCap in total.  Tallies are
 * left as raw sums of the traced photons (read with mcb_get_raw_tallies).                */
int mcb_run_trace(mcb_handle *h, int64_t nPhotons, const float *rn, int64_t rnStride,
                  int32_t maxEventsPerPhoton, mcb_event *events, int64_t eventCap,
                  int64_t *nEvents);

/* ---- testing aid ------------------------------------------------------------------- */
/* The first n 32-bit outputs of the Philox4x32-10 stream of (seed, photon id).          */
int mcb_debug_philox(mcb_handle *h, uint64_t seed, uint64_t photon, int n, uint32_t *out);
/* Measured ceiling of the operation that bounds the photon kernels: fully divergent 4-byte gathers (one 32-byte
 * sector per lane per load, loadsInFlight independent loads per lane: 1, 2, 4, 8 or 16; negative: 16-byte loads,
 * -1, -4 or -8 of them) inside a buffer of `bytes`
 * bytes, launched like the flux kernels (persistent, 128 threads, blocksPerSM CTAs per SM).  Returns gathers per
 * second -- the denominator of bench.py's roofline.l2_gather.                                                    */
int mcb_debug_gather_probe(mcb_handle *h, int64_t bytes, int loadsInFlight, int blocksPerSM, int iterations,
                           double *gathersPerSecond);
/* The vacuum-distance map the packed extinction field of the staged domain is encoded with (nx*ny*nz bytes, x
 * fastest): per cell the Chebyshev distance, in cells, to the nearest cell with extinction, 0 for such a cell --
 * periodic in x and y, nothing above the top or below the surface, limited by 64.  The photon-pool kernels cross
 * that many cells of vacuum in one step (fewer when the boundary the ray is heading for is closer).  Fails if the staged grid is not marched that way.               */
int mcb_debug_distance_map(mcb_handle *h, uint8_t *out, int64_t nBytes);

#ifdef __cplusplus
}
#endif
#endif
