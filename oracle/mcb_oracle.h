/*
 * mcb_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the photon-tracing hot path of MCBRaT3D
 * (reference files are cited per function in mcb_oracle.c).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker / reported CPU baseline.  The
 * CUDA product (mcbrat3d_b200/csrc) never links, includes or calls it.
 *
 * PARITY STATUS: "parity unpinned".  The reference ships no tests, golden
 * vectors or fixtures for this path and cannot be
 * compiled in this image nor on the GPU box (no Fortran front end on either:
 * profiles/r02_compiler_probe_gpu_box.log).  oracle/ref_build/ holds the recipe
 * that compiles the UNMODIFIED reference sources into oracle/_ref and turns its
 * per-photon output into tests/golden/ref_trace_*.npz; tests/test_ref_fixtures.py
 * is the pin once those fixtures exist.  What IS pinned until then:
 *   - the RNG against the published MT19937 known-answer vectors
 *     (RandomNumbersForMC.f95:8-10 declares identity with mt19937ar-cok.c);
 *   - physical invariants (energy closure, Beer's law direct beam, isothermal
 *     thermal closure, plane-parallel answers from an independent solver).
 */
#ifndef MCB_ORACLE_H
#define MCB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- random numbers (RandomNumbersForMC.f95) ------------------------------------- */
typedef struct {
  int       mode;        /* 0 = MT19937, 1 = injected array                          */
  uint32_t  mt[624];
  int       mti;
  const float *inj;      /* injected random reals (trace harness)                    */
  int64_t   ninj, pos;
  int       exhausted;   /* set when more injected numbers were asked for than given */
  int64_t   ndrawn;
} orc_rng;

void     orc_rng_init_scalar(orc_rng *r, uint32_t seed);                 /* RNG:171-187 */
void     orc_rng_init_array(orc_rng *r, const uint32_t *key, int nkey);  /* RNG:189-241 */
void     orc_rng_init_injected(orc_rng *r, const float *vals, int64_t n);
uint32_t orc_rng_int(orc_rng *r);                                        /* RNG:245-260 */
double   orc_rng_double(orc_rng *r);                                     /* RNG:277-292 */
float    orc_rng_real(orc_rng *r);                                       /* RNG:294-301 */

/* ---- searches (numericUtilities.f95) ---------------------------------------------- */
int orc_findIndexDouble(double value, const double *table, int n, int firstGuess); /* NUM:206-260 */
int orc_findIndexMixed(float value, const double *table, int n, int firstGuess);   /* NUM:262-315 */
int orc_findCDFIndex(float value, const double *table, int n, int stride);          /* NUM:317-348 */

/* ---- event trace -------------------------------------------------------------------
 * One record per photon event.  Layout is mirrored by mcb_event in
 * include/mcbrat_cuda.h (the CUDA trace harness writes the same record).        */
enum {
  ORC_EV_BIRTH = 1, ORC_EV_SCATTER = 2, ORC_EV_SURFACE = 3, ORC_EV_EXIT_TOP = 4,
  ORC_EV_KILLED_SURFACE = 5, ORC_EV_KILLED_ROULETTE = 6, ORC_EV_BAD = 7,
  ORC_EV_LOCAL_ESTIMATE = 8, ORC_EV_RN_EXHAUSTED = 9, ORC_EV_NULL_COLLISION = 10
};

typedef struct {
  int32_t photon;      /* 0-based photon number inside the traced batch                 */
  int32_t kind;        /* ORC_EV_*                                                      */
  int32_t ix, iy, iz;  /* 1-based cell indices after the event                          */
  int32_t component;   /* scatter: component chosen; local estimate: direction (1-based) */
  int32_t phaseIndex;  /* scatter: phase-table entry (column) used                      */
  int32_t angleIndex;  /* scatter: inverse-table row used (1-based)                     */
  int32_t order;       /* scattering order after the event                              */
  int32_t nrn;         /* random numbers consumed by the run so far                      */
  float   weight;      /* photon weight after the event (local estimate: contribution)  */
  float   tau;         /* optical path sampled for the leg (local estimate: tau to edge) */
  double  path;        /* geometric length of the leg that ended in this event          */
  double  x, y, z;     /* position after the event                                      */
  float   dir[3];      /* direction cosines after the event                             */
  int32_t pad;
} orc_event;

/* ---- the domain's hot-path view (opticalProperties.f95 type(domain), OPT:77-111) --- */
typedef struct {
  int nx, ny, nz, nc;
  double *xE, *yE, *zE;            /* cell edges, n+1 each                          */
  double *totalExt;                /* (nx,ny,nz) x fastest                          */
  double *cumExt, *ssa;            /* (nx,ny,nz,nc)                                 */
  int32_t *phaseIdx;               /* (nx,ny,nz,nc), 1-based entry, 0 = none        */
  double albedo;                   /* Lambertian surface albedo                     */
  int   *invS, *invE;  float **inv;        /* inverse tables  (nS,nE) per component */
  int   *fwdS, *fwdE;  float **fwd, **fwdOrig; /* forward tables per component      */
} orc_domain;

orc_domain *orc_domain_new(int nx, int ny, int nz, int nc,
                           const double *xE, const double *yE, const double *zE,
                           const double *totalExt, const double *cumExt,
                           const double *ssa, const int32_t *phaseIdx, double albedo);
void orc_domain_set_inverse(orc_domain *d, int comp1, int nS, int nE, const float *T);
void orc_domain_set_forward(orc_domain *d, int comp1, int nS, int nE,
                            const float *P, const float *Porig);
void orc_domain_free(orc_domain *d);

/* accumulateExtinctionAlongPath, OPT:1656-1815.  Returns extAccumulated. */
float orc_march(const orc_domain *d, const float dir[3],
                double *x, double *y, double *z, int *ix, int *iy, int *iz,
                int hasTarget, float target, double *totalPath, int64_t *crossings);

/* ---- photon streams (monteCarloIllumination.f95 type(photonStream), ILL:35-42) ----- */
typedef struct {
  int64_t n, current;              /* current is 1-based like the reference          */
  double *x, *y, *z;
  float  *mu, *phi;
} orc_photons;

orc_photons *orc_photons_directional(float solarMu, float solarAzimuthDeg,
                                     int64_t n, orc_rng *r);              /* ILL:62-101 */
orc_photons *orc_photons_bbemission(double fracAtmsPower, const double *voxelCDF,
                                    int nx, int ny, int nz,
                                    int64_t n, orc_rng *r);               /* ILL:431-522 */
void orc_photons_free(orc_photons *p);

/* emission_weightingNEW, EMI:424-550 (no spectral response file).  Fills voxelCDF
 * (nx,ny,nz) and returns fracAtmsPower; totalFlux (W m^-2, monochromatic) optional. */
double orc_emission_weighting(const orc_domain *d, const double *temps, double lambda_um,
                              double sfcTemp, double *voxelCDF, double *totalFlux);

/* ---- integrator (monteCarloRadiativeTransfer.f95 type(integrator), INT:40-117) ------ */
typedef struct {
  int useRayTracing;                 /* .false. = maximum cross-section, INT:564-571   */
  int useRussianRoulette;
  float RussianRouletteW;
  int useRussianRouletteForIntensity;
  float zetaMin;
  int useHybridPhaseFunsForIntenCalcs;
  int numOrdersOrigPhaseFunIntenCalcs;
  int limitIntensityContributions;
  float maxIntensityContribution;
  float LW_flag;
} orc_options;

typedef struct {
  int64_t photons, crossings, scatters, surfaceHits, topExits, bad,
          leRays, leCrossings, rouletteKills, rnDrawn;
} orc_counters;

typedef struct {
  int nx, ny, nz, nc;
  int xyRegularlySpaced, zRegularlySpaced;
  double deltaX, deltaY, deltaZ, x0, y0, z0;
  double *xPosition, *yPosition, *zPosition;
  orc_options opt;
  int computeIntensity, nDir;
  float *intensityDirections;        /* (3,nDir)                                        */
  float *fluxUp, *fluxDown, *fluxAbsorbed;     /* (nx,ny)                               */
  float *volumeAbsorption;                     /* (nx,ny,nz)                            */
  float *intensity;                            /* (nx,ny,nDir)                          */
  float *intensityByComponent;                 /* (nx,ny,nDir,0:nc)                     */
  float *intensityExcess;                      /* (nDir,0:nc)                           */
  orc_counters cnt;
  orc_event *trace; int64_t traceCap, traceN; int32_t tracePhoton0;
} orc_integrator;

void orc_default_options(orc_options *o);
orc_integrator *orc_integrator_new(const orc_domain *d);                 /* INT:129-201 */
void orc_integrator_set_options(orc_integrator *g, const orc_options *o);/* INT:1046-1337 */
/* direction cosines as built by specifyParameters, INT:1267-1269 */
void orc_integrator_set_views(orc_integrator *g, int nDir, const float *mus, const float *phisDeg);
void orc_integrator_set_view_cosines(orc_integrator *g, int nDir, const float *dirCos);
void orc_integrator_set_trace(orc_integrator *g, orc_event *buf, int64_t cap);
void orc_integrator_free(orc_integrator *g);
void orc_make_direction_cosines(float mu, float phi, float out[3]);      /* INT:1876-1894 */

/* computeRadiativeTransfer, INT:209-391 (normalise=0 leaves raw sums, for the trace harness).
 * Returns 0 on success, 1 if no photon was processed (INT:831-839).                */
int orc_compute_radiative_transfer(orc_integrator *g, const orc_domain *d, orc_rng *r,
                                   orc_photons *p, int64_t numPhotonsPerBatch,
                                   int normalise, int64_t *numPhotonsProcessed);

/* Fixed-random-number single-photon trace harness: photon p uses rn[p*stride ...] for its
 * source draws and then its transport draws.  Tallies are left as raw sums.            */
int64_t orc_trace_photons(orc_integrator *g, const orc_domain *d,
                          int source, float solarMu, float solarAzimuthDeg,
                          double fracAtmsPower, const double *voxelCDF,
                          int64_t nPhotons, const float *rn, int64_t stride);

/* reportResults, INT:845-1042.  Any pointer may be NULL (Fortran optional).        */
void orc_report_results(const orc_integrator *g,
                        float *meanFluxUp, float *meanFluxDown, float *meanFluxAbsorbed,
                        float *fluxUp, float *fluxDown, float *fluxAbsorbed,
                        float *absorbedProfile, float *volumeAbsorption,
                        float *meanIntensity, float *intensity);

/* ---- the driver's batch loop + statistics (monteCarloDriver.f95:949-1052, 1188-1228) -
 * source: 0 = solar (solarMu, solarAzimuthDeg), 1 = thermal (fracAtmsPower, voxelCDF).
 * RNG seeded init_by_array{iseed, rank, thread} (DRV:901).  Outputs are (mean, stderr)
 * pairs: meanFlux*[2], flux*[nx*ny*2], absorbedProfile[nz*2], radiance[nx*ny*nDir*2];
 * moments are returned UN-finalised in *Stats when finalise == 0 so that several
 * workers can be summed first (DRV:1151-1166).                                       */
typedef struct {
  double *meanFluxUpStats, *meanFluxDownStats, *meanFluxAbsorbedStats;  /* [2]          */
  double *fluxUpStats, *fluxDownStats, *fluxAbsorbedStats;              /* [nx*ny*2]    */
  double *absorbedProfileStats;                                          /* [nz*2]       */
  double *absorbedVolumeStats;                                           /* [nx*ny*nz*2] or NULL */
  double *radianceStats;                                                 /* [nx*ny*nDir*2] or NULL */
} orc_stats;

int64_t orc_run_batches(orc_integrator *g, const orc_domain *d,
                        int source, float solarMu, float solarAzimuthDeg,
                        double fracAtmsPower, const double *voxelCDF,
                        int iseed, int rank, int thread,
                        int64_t numBatches, int64_t numPhotonsPerBatch,
                        orc_stats *st);
void orc_finalise_stats(double *stats, int64_t n, double solarFlux,
                        int64_t totalNumPhotons, int64_t batchesCompleted); /* DRV:1188-1228 */

/* ---- per-wavelength optical-property assembly: read_SSPTable's inner loops (OPT:204-299) followed by
 * getOpticalPropertiesByComponent (OPT:1022-1061).  The netCDF reads are not restated: the caller hands
 * over what they return for one lambdaIndex.                                                            */
enum { ORC_COMP_VOLEXT = 0, ORC_COMP_ABSXSEC = 1, ORC_COMP_PROFILE = 2 };
typedef struct {
  int32_t kind;            /* extType "volExt" | "absXsec" | an explicit horizontally uniform profile     */
  int32_t physIndex;       /* volExt: 1-based slot of massConc / Reff ("comp-gasComp", OPT:264)           */
  int32_t nTable;          /* volExt: nReff; otherwise number of levels                                   */
  int32_t zLevelBase;      /* 1-based (OPT:201)                                                           */
  const float *key;        /* volExt: phaseFunctionKeyT(nReff), default real (OPT:164, 243)               */
  const double *ext;       /* volExt: ExtinctionT(nReff); absXsec: xsec(nLevels); profile: extinction     */
  const double *ssa;       /* volExt: SingleScatteringAlbedoT(nReff); profile: ssa(nLevels)               */
  const int32_t *phaseIdx; /* profile: phaseFunctionIndex(nLevels)                                        */
} orc_component;
/* massConc, Reff: (nPhys, nx, ny, nz) component fastest (OPT:72-73); numConc: numConc(1,1,:) (OPT:223).
 * Outputs in the domain's layout.  Returns 0, or 1 = "Effective radius outside of table range" (OPT:289),
 * 2 = a component does not fit the vertical extent of the domain.                                        */
int orc_assemble_optics(int nx, int ny, int nz, int nPhys, const double *massConc, const double *Reff,
                        const double *numConc, int nc, const orc_component *comps, int setup,
                        double *totalExt, double *cumExt, double *ssa, int32_t *phaseIdx);

/* getFrequencyDistrNEW EMI:552-573: totalPhotons draws binned with findCDFIndex */
void orc_frequency_distribution(int numLambda, const double *CDF, int64_t totalPhotons, orc_rng *r, int64_t *distribution);

/* computeInversePhaseFunction INV:113-168 once the phase function is known at nAngles points increasing in mu
 * (native angles reversed, or max(nMoments,2) Lobatto nodes, INV:87-112): trapezoid CDF in mu, the bracket of every
 * probability (i-1)/(nSteps-1) by findIndex with the previous bracket as first guess, analytic inversion.         */
void orc_inverse_phase_function(int nAngles, const float *mus, const float *values, int nSteps, float *inverseTable);

/* tabulateForwardPhaseFunctions OPT:1872-1934 for one Legendre-stored entry: the phase function at nS equally spaced
 * angles 0..pi (OPT:1912-1913), value = sum (2l+1) chi_l P_l(cos angle) in single precision (SPF:480-498, NUM:187-205);
 * nCoef = 0 is the isotropic special case with value 1/2 (quirk q14).                                              */
void orc_forward_phase_function(int nCoef, const float *legendreCoefficients, int nS, float *values);

/* computeLobattoTerms NUM:27-114: abscissas (increasing) and weights of n-point Lobatto quadrature on [-1, 1] */
void orc_lobatto_terms(int n, float *mus, float *weights);

/* getPhaseFunctionValues_one SPF:448-531: Legendre-stored (nStored = 0, nCoef coefficients chi_1..chi_n) or stored as
 * nStored angle / value pairs (linear in the cosine of the angle)                                                 */
void orc_phase_function_values(int nCoef, const float *legendreCoefficients, int nStored, const float *storedAngle,
                               const float *storedValue, int nAngles, const float *scatteringAngle, float *value);

/* INV:97-112: mus (Lobatto abscissas) and the values of a Legendre-stored phase function there: max(nCoef, 2) each */
void orc_inversion_inputs_legendre(int nCoef, const float *legendreCoefficients, float *mus, float *values);

/* computeHybridPhaseFunctions OPT:1936-2050 for one entry tabulated at nAngles angles (radians); returns the
 * transition index (0 if the entry keeps its original values)                                                     */
int orc_hybrid_phase_function(int nAngles, const float *angles, const float *values, float gaussianWidth, float *newValues);

#ifdef __cplusplus
}
#endif
#endif
