/*
 * mcb_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See mcb_oracle.h.
 *
 * Plain-C restatement of the MCBRaT3D photon loop with the reference's mixed
 * f32/f64 arithmetic.  "INT" = Integrators/monteCarloRadiativeTransfer.f95,
 * "OPT" = src/opticalProperties.f95, "ILL" = src/monteCarloIllumination.f95,
 * "EMI" = src/emissionAndBroadBandWeights.f95, "RNG" = src/RandomNumbersForMC.f95,
 * "NUM" = src/numericUtilities.f95, "DRV" = Drivers/monteCarloDriver.f95.
 *
 * Arithmetic conventions
 *   - Fortran `real` = float, `real(8)` = double, default integer = int32_t.
 *   - Build with -ffp-contract=off: the reference has no fused multiply-adds.
 *   - Single-precision transcendental intrinsics (log, exp, cos, sin, acos) are
 *     evaluated as the correctly rounded value, (float)f((double)x).  The
 *     reference inherits whatever libm its compiler links; choosing the correctly
 *     rounded result makes the oracle reproducible across libms and lets the CUDA
 *     trace harness match it bit for bit.  sqrt and division are IEEE in both.
 *   - spacing()/tiny()/huge() follow gfortran semantics (spacing(0) = tiny).
 *
 * PARITY STATUS: parity unpinned (no reference fixtures exist; see header).
 */
#include "mcb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------- */
/* Fortran intrinsics                                                                    */
/* ------------------------------------------------------------------------------------- */
static inline double sp64(double x) {            /* spacing(real(8)) */
  x = fabs(x);
  if (x == 0.0) return DBL_MIN;
  double s = nextafter(x, INFINITY) - x;
  return s < DBL_MIN ? DBL_MIN : s;
}
static inline float sp32(float x) {              /* spacing(real) */
  x = fabsf(x);
  if (x == 0.0f) return FLT_MIN;
  float s = nextafterf(x, INFINITY) - x;
  return s < FLT_MIN ? FLT_MIN : s;
}
#define TINY32 FLT_MIN
#define HUGE64 DBL_MAX
static inline float f_log(float x)  { return (float)log((double)x); }
static inline float f_exp(float x)  { return (float)exp((double)x); }
static inline float f_cos(float x)  { return (float)cos((double)x); }
static inline float f_sin(float x)  { return (float)sin((double)x); }
static inline float f_acos(float x) { return (float)acos((double)x); }

/* real, parameter :: Pi = 3.14159265358979312 (INT:31, OPT:26) -> f32 */
static const float PI32 = 3.14159265358979312f;

/* ------------------------------------------------------------------------------------- */
/* RandomNumbersForMC.f95: MT19937 (mt19937ar-cok)                                       */
/* ------------------------------------------------------------------------------------- */
#define MT_N 624
#define MT_M 397

void orc_rng_init_scalar(orc_rng *r, uint32_t seed) {       /* RNG:171-187 */
  memset(r, 0, sizeof(*r));
  r->mt[0] = seed;
  for (int i = 1; i < MT_N; ++i)
    r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
  r->mti = MT_N;
}

void orc_rng_init_array(orc_rng *r, const uint32_t *key, int nkey) {  /* RNG:189-241 */
  orc_rng_init_scalar(r, 19650218u);
  int nFirstLoop = MT_N > nkey ? MT_N : nkey;
  int nWraps = 0;
  for (int k = 1; k <= nFirstLoop; ++k) {
    int i = (k + nWraps) % MT_N;
    int j = (k - 1) % nkey;
    if (i == 0) {
      r->mt[0] = r->mt[MT_N - 1];
      r->mt[1] = (r->mt[1] ^ ((r->mt[0] ^ (r->mt[0] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
      nWraps += 1;
    } else {
      r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    }
  }
  for (int i = nFirstLoop % MT_N + nWraps + 1; i <= MT_N - 1; ++i)
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
  r->mt[0] = r->mt[MT_N - 1];
  for (int i = 1; i <= nFirstLoop % MT_N + nWraps; ++i)
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
  r->mt[0] = 0x80000000u;
  r->mti = MT_N;
}

void orc_rng_init_injected(orc_rng *r, const float *vals, int64_t n) {
  memset(r, 0, sizeof(*r));
  r->mode = 1; r->inj = vals; r->ninj = n;
}

static inline uint32_t mt_twist(uint32_t u, uint32_t v) {   /* RNG:118-134 */
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}
static void mt_next_state(orc_rng *r) {                     /* RNG:136-154 */
  uint32_t *s = r->mt;
  int k;
  for (k = 0; k < MT_N - MT_M; ++k)      s[k] = s[k + MT_M] ^ mt_twist(s[k], s[k + 1]);
  for (k = MT_N - MT_M; k < MT_N - 1; ++k) s[k] = s[k + MT_M - MT_N] ^ mt_twist(s[k], s[k + 1]);
  s[MT_N - 1] = s[MT_M - 1] ^ mt_twist(s[MT_N - 1], s[0]);
  r->mti = 0;
}
uint32_t orc_rng_int(orc_rng *r) {                          /* RNG:245-260, temper RNG:156-166 */
  if (r->mti >= MT_N) mt_next_state(r);
  uint32_t x = r->mt[r->mti++];
  x ^= (x >> 11);
  x ^= (x << 7) & 0x9d2c5680u;
  x ^= (x << 15) & 0xefc60000u;
  x ^= (x >> 18);
  return x;
}
double orc_rng_double(orc_rng *r) {                         /* RNG:277-292 */
  return (double)orc_rng_int(r) / (4294967296.0 - 1.0);
}
float orc_rng_real(orc_rng *r) {                            /* RNG:294-301 */
  r->ndrawn++;
  if (r->mode == 1) {
    if (r->pos >= r->ninj) { r->exhausted = 1; return 0.25f; }
    return r->inj[r->pos++];
  }
  return (float)orc_rng_double(r);
}

/* ------------------------------------------------------------------------------------- */
/* numericUtilities.f95: table searches.  Tables are 1-based in the reference; here the   */
/* C array t[0..n-1] holds table(1..n) and results are the reference's 1-based values.    */
/* firstGuess <= 0 means "not present".                                                   */
/* ------------------------------------------------------------------------------------- */
#define T1(i) table[(i) - 1]
int orc_findIndexDouble(double value, const double *table, int n, int firstGuess) { /* NUM:206-260 */
  int lowerBound, upperBound, midPoint, increment;
  if (firstGuess > 0) {
    lowerBound = firstGuess; increment = 1;
    for (;;) {
      upperBound = lowerBound + increment < n ? lowerBound + increment : n;
      if (lowerBound == n || (T1(lowerBound) <= value && T1(upperBound) > value)) break;
      if (T1(lowerBound) > value) {
        upperBound = lowerBound;
        lowerBound = upperBound - increment > 1 ? upperBound - increment : 1;
      } else {
        lowerBound = upperBound;
      }
      increment *= 2;
    }
  } else {
    lowerBound = 0; upperBound = n;
  }
  for (;;) {
    if (lowerBound == n || upperBound <= lowerBound + 1) break;
    midPoint = (lowerBound + upperBound) / 2;
    if (value >= T1(midPoint)) lowerBound = midPoint; else upperBound = midPoint;
  }
  return lowerBound;
}
int orc_findIndexMixed(float value, const double *table, int n, int firstGuess) {  /* NUM:262-315 */
  return orc_findIndexDouble((double)value, table, n, firstGuess);
}
/* findCDFIndex_double NUM:317-348: smallest i with value <= table(i) (bisection form).
 * `stride` lets the caller walk the pointer slices colWeights/levelWeights (EMI:56-57). */
int orc_findCDFIndex(float value, const double *table, int n, int stride) {
  int lowerBound = 0, upperBound = n, midPoint;
  double v = (double)value;
  for (;;) {
    if (lowerBound == n || upperBound <= lowerBound + 1) break;
    midPoint = (lowerBound + upperBound) / 2;
    if (v > table[(int64_t)(midPoint - 1) * stride]) lowerBound = midPoint; else upperBound = midPoint;
  }
  return upperBound;
}
#undef T1

/* ------------------------------------------------------------------------------------- */
/* domain                                                                                */
/* ------------------------------------------------------------------------------------- */
static void *dup_mem(const void *src, size_t bytes) {
  void *p = malloc(bytes ? bytes : 1);
  if (src) memcpy(p, src, bytes);
  return p;
}

orc_domain *orc_domain_new(int nx, int ny, int nz, int nc,
                           const double *xE, const double *yE, const double *zE,
                           const double *totalExt, const double *cumExt,
                           const double *ssa, const int32_t *phaseIdx, double albedo) {
  orc_domain *d = (orc_domain *)calloc(1, sizeof(*d));
  size_t cells = (size_t)nx * ny * nz;
  d->nx = nx; d->ny = ny; d->nz = nz; d->nc = nc; d->albedo = albedo;
  d->xE = (double *)dup_mem(xE, sizeof(double) * (nx + 1));
  d->yE = (double *)dup_mem(yE, sizeof(double) * (ny + 1));
  d->zE = (double *)dup_mem(zE, sizeof(double) * (nz + 1));
  d->totalExt = (double *)dup_mem(totalExt, sizeof(double) * cells);
  d->cumExt = (double *)dup_mem(cumExt, sizeof(double) * cells * nc);
  d->ssa = (double *)dup_mem(ssa, sizeof(double) * cells * nc);
  d->phaseIdx = (int32_t *)dup_mem(phaseIdx, sizeof(int32_t) * cells * nc);
  d->invS = (int *)calloc(nc, sizeof(int)); d->invE = (int *)calloc(nc, sizeof(int));
  d->fwdS = (int *)calloc(nc, sizeof(int)); d->fwdE = (int *)calloc(nc, sizeof(int));
  d->inv = (float **)calloc(nc, sizeof(float *));
  d->fwd = (float **)calloc(nc, sizeof(float *));
  d->fwdOrig = (float **)calloc(nc, sizeof(float *));
  return d;
}
void orc_domain_set_inverse(orc_domain *d, int comp1, int nS, int nE, const float *T) {
  int c = comp1 - 1;
  free(d->inv[c]);
  d->invS[c] = nS; d->invE[c] = nE;
  d->inv[c] = (float *)dup_mem(T, sizeof(float) * (size_t)nS * nE);
}
void orc_domain_set_forward(orc_domain *d, int comp1, int nS, int nE,
                            const float *P, const float *Porig) {
  int c = comp1 - 1;
  free(d->fwd[c]); free(d->fwdOrig[c]);
  d->fwdS[c] = nS; d->fwdE[c] = nE;
  d->fwd[c] = (float *)dup_mem(P, sizeof(float) * (size_t)nS * nE);
  d->fwdOrig[c] = (float *)dup_mem(Porig ? Porig : P, sizeof(float) * (size_t)nS * nE);
}
void orc_domain_free(orc_domain *d) {
  if (!d) return;
  for (int c = 0; c < d->nc; ++c) { free(d->inv[c]); free(d->fwd[c]); free(d->fwdOrig[c]); }
  free(d->inv); free(d->fwd); free(d->fwdOrig);
  free(d->invS); free(d->invE); free(d->fwdS); free(d->fwdE);
  free(d->xE); free(d->yE); free(d->zE);
  free(d->totalExt); free(d->cumExt); free(d->ssa); free(d->phaseIdx);
  free(d);
}

#define CELL(d, ix, iy, iz) ((size_t)((ix) - 1) + (size_t)(d)->nx * ((size_t)((iy) - 1) + (size_t)(d)->ny * (size_t)((iz) - 1)))
#define CELLC(d, ix, iy, iz, c) (CELL(d, ix, iy, iz) + (size_t)(d)->nx * (d)->ny * (d)->nz * (size_t)((c) - 1))

/* ------------------------------------------------------------------------------------- */
/* accumulateExtinctionAlongPath, OPT:1656-1815                                           */
/* Edge arrays are 1-based in the reference: xPosition(i) == xE[i-1].                     */
/* ------------------------------------------------------------------------------------- */
float orc_march(const orc_domain *d, const float dir[3],
                double *xPos, double *yPos, double *zPos, int *xIndex, int *yIndex, int *zIndex,
                int hasTarget, float extToAccumulate, double *totalPathOut, int64_t *crossings) {
  const int nX = d->nx, nY = d->ny, nZ = d->nz;
  float extAccumulated = 0.0f;
  double totalPath = 0.0;
  int side[3], inc[3];
  for (int k = 0; k < 3; ++k) {                         /* OPT:1690-1692 */
    side[k] = dir[k] >= 0.0f ? 1 : 0;
    inc[k] = dir[k] >= 0.0f ? 1 : -1;
  }
  const double z0 = d->zE[0], zMax = d->zE[nZ];
  double x = *xPos, y = *yPos, z = *zPos;
  int ix = *xIndex, iy = *yIndex, iz = *zIndex;

  for (;;) {
    double step[3];
    /* OPT:1705-1712; xPosition(xIndex + side) -> xE[ix + side - 1] */
    step[0] = fabsf(dir[0]) >= 2.0f * TINY32 ? (d->xE[ix + side[0] - 1] - x) / (double)dir[0] : HUGE64;
    step[1] = fabsf(dir[1]) >= 2.0f * TINY32 ? (d->yE[iy + side[1] - 1] - y) / (double)dir[1] : HUGE64;
    step[2] = fabsf(dir[2]) >= 2.0f * TINY32 ? (d->zE[iz + side[2] - 1] - z) / (double)dir[2] : HUGE64;
    double thisStep = step[0];
    if (step[1] < thisStep) thisStep = step[1];
    if (step[2] < thisStep) thisStep = step[2];
    if (thisStep <= 0.0) { extAccumulated = -2.0f; break; }          /* OPT:1719-1722 */

    double thisCellExt = d->totalExt[CELL(d, ix, iy, iz)];           /* OPT:1727 */
    if (crossings) (*crossings)++;

    if (hasTarget) {                                                 /* OPT:1729-1739 */
      if ((double)extAccumulated + thisStep * thisCellExt > (double)extToAccumulate) {
        thisStep = (double)(extToAccumulate - extAccumulated) / thisCellExt;
        x = x + thisStep * (double)dir[0];
        y = y + thisStep * (double)dir[1];
        z = z + thisStep * (double)dir[2];
        totalPath = totalPath + thisStep;
        extAccumulated = extToAccumulate;
        break;
      }
    }
    extAccumulated = (float)((double)extAccumulated + thisStep * thisCellExt);   /* OPT:1743 */
    totalPath = totalPath + thisStep;

    if (step[0] <= thisStep) {                                       /* OPT:1752-1759 */
      x = d->xE[ix + side[0] - 1];
      ix = ix + inc[0];
    } else {
      x = x + thisStep * (double)dir[0];
      if (fabs(d->xE[ix + side[0] - 1] - x) <= 2.0 * sp64(x)) ix = ix + inc[0];
    }
    if (step[1] <= thisStep) {                                       /* OPT:1761-1768 */
      y = d->yE[iy + side[1] - 1];
      iy = iy + inc[1];
    } else {
      y = y + thisStep * (double)dir[1];
      if (fabs(d->yE[iy + side[1] - 1] - y) <= 2.0 * sp64(y)) iy = iy + inc[1];
    }
    if (step[2] <= thisStep) {                                       /* OPT:1770-1777 */
      z = d->zE[iz + side[2] - 1];
      iz = iz + inc[2];
    } else {
      z = z + thisStep * (double)dir[2];
      if (fabs(d->zE[iz + side[2] - 1] - z) <= 2.0 * sp64(z)) iz = iz + inc[2];
    }

    if (ix <= 0) {                                                   /* OPT:1782-1788 */
      ix = nX;
      x = d->xE[ix] + (double)(inc[0] * 2) * sp64(x);
    } else if (ix >= nX + 1) {
      ix = 1;
      x = d->xE[0] + (double)(inc[0] * 2) * sp64(x);
    }
    if (iy <= 0) {                                                   /* OPT:1790-1796: cellIncrement(1) (sic) */
      iy = nY;
      y = d->yE[iy] + (double)(inc[0] * 2) * sp64(y);
    } else if (iy >= nY + 1) {
      iy = 1;
      y = d->yE[0] + (double)(inc[0] * 2) * sp64(y);
    }
    if (iz > nZ) { z = zMax + 2.0 * sp64(zMax); break; }             /* OPT:1801-1804 */
    if (iz < 1)  { z = z0; break; }                                  /* OPT:1809-1812 */
  }
  *xPos = x; *yPos = y; *zPos = z; *xIndex = ix; *yIndex = iy; *zIndex = iz;
  if (totalPathOut) *totalPathOut = totalPath;
  return extAccumulated;
}

/* ------------------------------------------------------------------------------------- */
/* photon streams                                                                        */
/* ------------------------------------------------------------------------------------- */
static orc_photons *photons_alloc(int64_t n) {
  orc_photons *p = (orc_photons *)calloc(1, sizeof(*p));
  size_t m = (size_t)(n > 0 ? n : 1);
  p->n = n; p->current = 0;
  p->x = (double *)malloc(sizeof(double) * m); p->y = (double *)malloc(sizeof(double) * m);
  p->z = (double *)malloc(sizeof(double) * m);
  p->mu = (float *)malloc(sizeof(float) * m); p->phi = (float *)malloc(sizeof(float) * m);
  return p;
}
void orc_photons_free(orc_photons *p) {
  if (!p) return;
  free(p->x); free(p->y); free(p->z); free(p->mu); free(p->phi); free(p);
}

orc_photons *orc_photons_directional(float solarMu, float solarAzimuth, int64_t n, orc_rng *r) { /* ILL:62-101 */
  orc_photons *p = photons_alloc(n);
  for (int64_t i = 0; i < n; ++i) {                       /* ILL:88-92 */
    p->x[i] = (double)orc_rng_real(r);
    p->y[i] = (double)orc_rng_real(r);
  }
  const float zTop = 1.0f - sp32(1.0f);                   /* ILL:93 */
  const float mu = -fabsf(solarMu);                       /* ILL:95 */
  const float phi = solarAzimuth * f_acos(-1.0f) / 180.0f;/* ILL:96 */
  for (int64_t i = 0; i < n; ++i) { p->z[i] = (double)zTop; p->mu[i] = mu; p->phi[i] = phi; }
  p->current = 1;
  return p;
}

orc_photons *orc_photons_bbemission(double fracAtmsPower, const double *voxelWeights,
                                    int numX, int numY, int numZ, int64_t n, orc_rng *r) { /* ILL:431-522 */
  orc_photons *p = photons_alloc(n);
  const float pi32 = f_acos(-1.0f);
  /* pointer slices, EMI:56-57: colWeights(j,k) = voxelWeights(numX,j,k); levelWeights(k) = voxelWeights(numX,numY,k) */
  const double *levelBase = voxelWeights + (size_t)(numX - 1) + (size_t)numX * (size_t)(numY - 1);
  for (int64_t i = 0; i < n; ++i) {
    float RN = orc_rng_real(r);                            /* ILL:483 */
    if ((double)RN > fracAtmsPower) {                      /* surface, ILL:484-493 */
      p->x[i] = (double)orc_rng_real(r);
      p->y[i] = (double)orc_rng_real(r);
      for (;;) {
        p->mu[i] = sqrtf(orc_rng_real(r));
        if (fabsf(p->mu[i]) > 2.0f * TINY32) break;
        if (r->exhausted) break;
      }
      p->phi[i] = orc_rng_real(r) * 2.0f * pi32;
      p->z[i] = 0.0;
    } else {                                               /* atmosphere, ILL:494-510 */
      RN = orc_rng_real(r);
      int ik = orc_findCDFIndex(RN, levelBase, numZ, numX * numY);
      const double *colBase = voxelWeights + (size_t)(numX - 1) + (size_t)numX * numY * (size_t)(ik - 1);
      int ij = orc_findCDFIndex(RN, colBase, numY, numX);
      const double *voxBase = voxelWeights + (size_t)numX * ((size_t)(ij - 1) + (size_t)numY * (size_t)(ik - 1));
      int ii = orc_findCDFIndex(RN, voxBase, numX, 1);

      p->z[i] = ((double)(ik - 1) * 1.0 / (double)numZ) + (double)(orc_rng_real(r) / (float)numZ);
      if (ik == 1 && p->z[i] == 0.0) p->z[i] = 0.0 + sp64(1.0);
      if (ik == numZ && p->z[i] > 1.0 - 2.0 * sp64(1.0)) p->z[i] = p->z[i] - (2.0 * sp64(1.0));
      p->x[i] = ((double)(ii - 1) * 1.0 / (double)numX) + (double)(orc_rng_real(r) * (1.0f / (float)numX));
      p->y[i] = ((double)(ij - 1) * 1.0 / (double)numY) + (double)(orc_rng_real(r) * (1.0f / (float)numY));
      for (;;) {
        p->mu[i] = 1.0f - (2.0f * orc_rng_real(r));
        if (fabsf(p->mu[i]) > 2.0f * TINY32) break;
        if (r->exhausted) break;
      }
      p->phi[i] = orc_rng_real(r) * 2.0f * pi32;
    }
  }
  p->current = 1;
  return p;
}

/* ------------------------------------------------------------------------------------- */
/* emission_weightingNEW, EMI:424-550 (single wavelength, no instrument response file)    */
/* ------------------------------------------------------------------------------------- */
double orc_emission_weighting(const orc_domain *d, const double *atmsTemp, double lambda_um,
                              double sfcTemp, double *voxelWeights, double *totalFlux) {
  const double h = 6.62606957e-34, c = 2.99792458e+8, k = 1.3806488e-23;
  const double a = 2.0 * h * (c * c);
  const double Pi = 4.0 * atan(1.0);
  const int nx = d->nx, ny = d->ny, nz = d->nz, nc = d->nc;
  const size_t cells = (size_t)nx * ny * nz;
  const double emiss = 1.0 - d->albedo;
  const double lambda = lambda_um / 1.0e6;
  const double b = h * c / (k * lambda);
  const double areaX = d->xE[nx] - d->xE[0], areaY = d->yE[ny] - d->yE[0];
  double sfcPower;
  if (emiss == 0.0 || sfcTemp == 0.0) {
    sfcPower = 0.0;
  } else {
    double sfcPlanckRad = (a / (pow(lambda, 5.0) * (exp(b / sfcTemp) - 1.0))) / 1.0e6;
    sfcPower = Pi * emiss * sfcPlanckRad * areaX * areaY * (1000.0 * 1000.0);
  }
  double atmsPower = 0.0, previous = 0.0, corr_contrib, temp_sum, corr = 0.0;
  int anyCold = 0;
  for (size_t i = 0; i < cells; ++i) if (atmsTemp[i] <= 0.0) anyCold = 1;
  memset(voxelWeights, 0, sizeof(double) * cells);
  if (!anyCold) {
    for (int iz = 1; iz <= nz; ++iz)
      for (int iy = 1; iy <= ny; ++iy)
        for (int ix = 1; ix <= nx; ++ix) {
          size_t cell = CELL(d, ix, iy, iz);
          double atmsPlanckRad = (a / (pow(lambda, 5.0) * (exp(b / atmsTemp[cell]) - 1.0))) / 1.0e6;
          /* ext(:,:,:,j) rebuilt from totalExt and cumulativeExt, OPT:872-882 */
          double sumSsaExt = 0.0;
          for (int j = 1; j <= nc; ++j) {
            double extj = j == 1 ? d->totalExt[cell] * d->cumExt[CELLC(d, ix, iy, iz, 1)]
                                 : d->totalExt[cell] * (d->cumExt[CELLC(d, ix, iy, iz, j)] - d->cumExt[CELLC(d, ix, iy, iz, j - 1)]);
            sumSsaExt += d->ssa[CELLC(d, ix, iy, iz, j)] * extj;
          }
          double totalAbsCoef = d->totalExt[cell] - sumSsaExt;
          double dz = d->zE[iz] - d->zE[iz - 1];
          corr_contrib = (4.0 * Pi * atmsPlanckRad * totalAbsCoef * dz) - corr;   /* EMI:505-509 (Kahan) */
          temp_sum = previous + corr_contrib;
          corr = (temp_sum - previous) - corr_contrib;
          previous = temp_sum;
          voxelWeights[cell] = previous;
        }
  }
  double fracAtmsPower = 0.0;
  double last = voxelWeights[cells - 1];
  if (last > 0.0) {                                        /* EMI:512-521 */
    atmsPower = last * areaX * areaY * (1000.0 * 1000.0) / (double)(nx * ny);
    for (size_t i = 0; i < cells; ++i) voxelWeights[i] = voxelWeights[i] / last;
    voxelWeights[cells - 1] = 1.0;
    fracAtmsPower = atmsPower / (atmsPower + sfcPower);
  }
  if (totalFlux) *totalFlux = (atmsPower + sfcPower) / (areaX * areaY * (1000.0 * 1000.0));  /* EMI:536-538 */
  return fracAtmsPower;
}

/* ------------------------------------------------------------------------------------- */
/* integrator                                                                            */
/* ------------------------------------------------------------------------------------- */
void orc_default_options(orc_options *o) {                  /* INT:53-96 with DRV:74-82 driver defaults NOT applied */
  o->useRayTracing = 1; o->useRussianRoulette = 1; o->RussianRouletteW = 1.0f;
  o->useRussianRouletteForIntensity = 0; o->zetaMin = 0.3f;
  o->useHybridPhaseFunsForIntenCalcs = 0; o->numOrdersOrigPhaseFunIntenCalcs = 0;
  o->limitIntensityContributions = 0; o->maxIntensityContribution = FLT_MAX;
  o->LW_flag = -1.0f;
}

orc_integrator *orc_integrator_new(const orc_domain *d) {   /* INT:129-201 */
  orc_integrator *g = (orc_integrator *)calloc(1, sizeof(*g));
  const int numX = d->nx, numY = d->ny, numZ = d->nz;
  g->nx = numX; g->ny = numY; g->nz = numZ; g->nc = d->nc;
  g->xPosition = (double *)dup_mem(d->xE, sizeof(double) * (numX + 1));
  g->yPosition = (double *)dup_mem(d->yE, sizeof(double) * (numY + 1));
  g->zPosition = (double *)dup_mem(d->zE, sizeof(double) * (numZ + 1));
  g->x0 = g->xPosition[0]; g->y0 = g->yPosition[0]; g->z0 = g->zPosition[0];
  /* real :: deltaX, deltaY, deltaZ  -- f32 locals (INT:140, 166-168) */
  float deltaX = (float)(g->xPosition[1] - g->xPosition[0]);
  float deltaY = (float)(g->yPosition[1] - g->yPosition[0]);
  float deltaZ = (float)(g->zPosition[1] - g->zPosition[0]);
  int xyReg = 1, zReg = 1;
  for (int i = 0; i < numX; ++i)                           /* INT:169-172 */
    if (!(fabs((g->xPosition[i + 1] - g->xPosition[i]) - (double)deltaX) <= 2.0 * sp64(g->xPosition[i + 1]))) xyReg = 0;
  for (int i = 0; i < numY; ++i)
    if (!(fabs((g->yPosition[i + 1] - g->yPosition[i]) - (double)deltaY) <= 2.0 * sp64(g->yPosition[i + 1]))) xyReg = 0;
  for (int i = 0; i < numZ; ++i)                           /* INT:177-178 */
    if (!(fabs((g->zPosition[i + 1] - g->zPosition[i]) - (double)deltaZ) <= sp64(g->zPosition[i + 1]))) zReg = 0;
  if (xyReg) { g->xyRegularlySpaced = 1; g->deltaX = (double)deltaX; g->deltaY = (double)deltaY; }
  if (zReg)  { g->zRegularlySpaced = 1; g->deltaZ = (double)deltaZ; }
  size_t cols = (size_t)numX * numY;
  g->fluxUp = (float *)calloc(cols, sizeof(float));
  g->fluxDown = (float *)calloc(cols, sizeof(float));
  g->fluxAbsorbed = (float *)calloc(cols, sizeof(float));
  g->volumeAbsorption = (float *)calloc(cols * numZ, sizeof(float));
  orc_default_options(&g->opt);
  return g;
}
void orc_integrator_set_options(orc_integrator *g, const orc_options *o) { g->opt = *o; }

void orc_make_direction_cosines(float mu, float phi, float out[3]) {   /* INT:1876-1894 */
  float sinTheta = sqrtf(1.0f - mu * mu);
  float cosPhi = f_cos(phi), sinPhi = f_sin(phi);
  out[0] = sinTheta * cosPhi; out[1] = sinTheta * sinPhi; out[2] = mu;
}

static void alloc_intensity(orc_integrator *g, int nDir) {           /* INT:1245-1258, 1287-1291 */
  size_t cols = (size_t)g->nx * g->ny;
  free(g->intensityDirections); free(g->intensity); free(g->intensityByComponent); free(g->intensityExcess);
  g->nDir = nDir;
  g->intensityDirections = (float *)calloc((size_t)3 * nDir, sizeof(float));
  g->intensity = (float *)calloc(cols * nDir, sizeof(float));
  g->intensityByComponent = (float *)calloc(cols * nDir * (g->nc + 1), sizeof(float));
  g->intensityExcess = (float *)calloc((size_t)nDir * (g->nc + 1), sizeof(float));
  g->computeIntensity = 1;
}
void orc_integrator_set_views(orc_integrator *g, int nDir, const float *mus, const float *phisDeg) {
  alloc_intensity(g, nDir);
  for (int i = 0; i < nDir; ++i)                                      /* INT:1267-1269 */
    orc_make_direction_cosines(mus[i], phisDeg[i] * PI32 / 180.0f, g->intensityDirections + 3 * i);
}
void orc_integrator_set_view_cosines(orc_integrator *g, int nDir, const float *dirCos) {
  alloc_intensity(g, nDir);
  memcpy(g->intensityDirections, dirCos, sizeof(float) * 3 * nDir);
}
void orc_integrator_set_trace(orc_integrator *g, orc_event *buf, int64_t cap) {
  g->trace = buf; g->traceCap = cap; g->traceN = 0; g->tracePhoton0 = 0;
}
void orc_integrator_free(orc_integrator *g) {
  if (!g) return;
  free(g->xPosition); free(g->yPosition); free(g->zPosition);
  free(g->fluxUp); free(g->fluxDown); free(g->fluxAbsorbed); free(g->volumeAbsorption);
  free(g->intensityDirections); free(g->intensity); free(g->intensityByComponent); free(g->intensityExcess);
  free(g);
}

/* findXYIndicies INT:1551-1578 */
static void findXYIndicies(const orc_integrator *g, double xPos, double yPos, int *xIndex, int *yIndex) {
  const int nxp1 = g->nx + 1, nyp1 = g->ny + 1;         /* size(xPosition), size(yPosition) */
#define XP(i) g->xPosition[(i) - 1]
#define YP(i) g->yPosition[(i) - 1]
  if (g->xyRegularlySpaced) {
    int xi = (int)((xPos - g->x0) / g->deltaX) + 1; if (xi > nxp1 - 1) xi = nxp1 - 1;
    int yi = (int)((yPos - g->y0) / g->deltaY) + 1; if (yi > nyp1 - 1) yi = nyp1 - 1;
    if (fabs(XP(xi + 1) - xPos) < sp64(xPos)) xi = xi + 1;
    if (fabs(YP(yi + 1) - yPos) < sp64(yPos)) yi = yi + 1;
    if (xi == nxp1) xi = 1;
    if (yi == nyp1) yi = 1;
    *xIndex = xi; *yIndex = yi;
  } else {
    int xi = orc_findIndexDouble(xPos, g->xPosition, nxp1, *xIndex);
    int yi = orc_findIndexDouble(yPos, g->yPosition, nyp1, *yIndex);
    if (fabs(XP(xi) - xPos) < sp64(xPos)) xi = xi + 1;
    if (fabs(YP(yi) - yPos) < sp64(yPos)) yi = yi + 1;
    if (xi >= nxp1) xi = 1;
    if (yi >= nyp1) yi = 1;
    *xIndex = xi; *yIndex = yi;
  }
#undef XP
#undef YP
}
/* findZIndex INT:1580-1592 (only reached with zRegularlySpaced, INT:484-486) */
static void findZIndex(const orc_integrator *g, double zPos, int *zIndex) {
  if (g->zRegularlySpaced) {
    int zi = (int)((zPos - g->z0) / g->deltaZ) + 1; if (zi > g->nz) zi = g->nz;
    if (fabs(g->zPosition[zi] - zPos) < sp64(zPos)) zi = zi + 1;
    *zIndex = zi;
  } else {
    *zIndex = orc_findIndexDouble(zPos, g->zPosition, g->nz + 1, *zIndex);
  }
}

/* computeScatteringAngle INT:1594-1621; *kOut receives angleIndex */
static float computeScatteringAngle(float randomDeviate, const float *table, int numIntervals, int *kOut) {
  int angleIndex = (int)(randomDeviate * (float)numIntervals) + 1;
  *kOut = angleIndex;
  if (angleIndex < numIntervals) {
    float leftOver = randomDeviate - (float)(angleIndex - 1) / (float)numIntervals;
    return (1.0f - leftOver) * table[angleIndex - 1] + leftOver * table[angleIndex];
  }
  return table[numIntervals - 1];
}

/* NEXT_DIRECT INT:1921-1948 */
static void next_direct(orc_rng *r, float scatteringCosine, float S[3]) {
  float D = 2.0f, AX = 0.0f, AY = 0.0f, B;
  while (D > 1.0f) {
    AX = 1.0f - 2.0f * orc_rng_real(r);
    AY = 1.0f - 2.0f * orc_rng_real(r);
    D = AX * AX + AY * AY;
    if (r->exhausted) break;
  }
  B = sqrtf((1.0f - scatteringCosine * scatteringCosine) / D);
  AX = AX * B;
  AY = AY * B;
  B = S[0] * AX - S[1] * AY;
  D = scatteringCosine - B / (1.0f + fabsf(S[2]));
  S[0] = S[0] * D + AX;
  S[1] = S[1] * D - AY;
  S[2] = S[2] * scatteringCosine - copysignf(fabsf(B), S[2] * B);
}

/* lookUpPhaseFuncValsFromTable INT:1834-1873 (one angle) */
static float lookUpPhaseFuncVal(const float *table, int nAngleSteps, float scatteringAngle) {
  float deltaTheta = PI32 / (float)(nAngleSteps - 1);
  int angleIndex = (int)(scatteringAngle / deltaTheta) + 1;
  if (angleIndex < nAngleSteps) {
    float weight = 1.0f - (scatteringAngle - (float)(angleIndex - 1) * deltaTheta) / deltaTheta;
    return weight * table[angleIndex - 1] + (1.0f - weight) * table[angleIndex];
  }
  return table[nAngleSteps - 1];
}

static void trace_event(orc_integrator *g, const orc_rng *r, int32_t photon, int kind,
                        int ix, int iy, int iz, int component, int phaseIndex, int angleIndex,
                        int order, float weight, float tau, double path,
                        double x, double y, double z, const float *dir) {
  if (!g->trace || g->traceN >= g->traceCap) { if (g->trace) g->traceN++; return; }
  orc_event *e = &g->trace[g->traceN++];
  memset(e, 0, sizeof(*e));
  e->photon = photon; e->kind = kind; e->ix = ix; e->iy = iy; e->iz = iz;
  e->component = component; e->phaseIndex = phaseIndex; e->angleIndex = angleIndex;
  e->order = order; e->nrn = (int32_t)r->ndrawn; e->weight = weight; e->tau = tau; e->path = path;
  e->x = x; e->y = y; e->z = z;
  if (dir) { e->dir[0] = dir[0]; e->dir[1] = dir[1]; e->dir[2] = dir[2]; }
}

/* computeIntensityContribution INT:1623-1832 */
static void computeIntensityContribution(orc_integrator *g, const orc_domain *d, float photonWeight,
                                         double xPos, double yPos, double zPos,
                                         int xIndex, int yIndex, int zIndex,
                                         const float directionCosines[3], int component,
                                         orc_rng *r, int scatteringOrder, int32_t photonNo,
                                         float *contributions, int *xIndexF, int *yIndexF) {
  const int numIntensityDirections = g->nDir;
  const int zIndexMax = g->nz + 1;                         /* size(zPosition), INT:1677 */
  const float *ID = g->intensityDirections;
  for (int i = 0; i < numIntensityDirections; ++i) {
    float normalizedPhaseFunc;
    if (component == 0) {                                  /* INT:1688-1694 */
      normalizedPhaseFunc = 1.0f / PI32;
    } else if (component < 0) {                            /* INT:1695-1696 */
      normalizedPhaseFunc = 1.0f / (4.0f * PI32 * fabsf(ID[3 * i + 2]));
    } else {                                               /* INT:1697-1727 */
      float projection = 0.0f;
      for (int k = 0; k < 3; ++k) projection = projection + directionCosines[k] * ID[3 * i + k];
      if (fabsf(projection) > 1.0f) projection = copysignf(1.0f, projection);
      float scatteringAngle = f_acos(projection);
      int phaseFunctionIndex = d->phaseIdx[CELLC(d, xIndex, yIndex, zIndex, component)];
      int c = component - 1;
      const float *tab = (g->opt.useHybridPhaseFunsForIntenCalcs &&
                          scatteringOrder <= g->opt.numOrdersOrigPhaseFunIntenCalcs)
                             ? d->fwdOrig[c] : d->fwd[c];
      float phaseFunctionVal = lookUpPhaseFuncVal(tab + (size_t)(phaseFunctionIndex - 1) * d->fwdS[c],
                                                  d->fwdS[c], scatteringAngle);
      normalizedPhaseFunc = phaseFunctionVal / (4.0f * PI32 * fabsf(ID[3 * i + 2]));
    }

    double xTemp = xPos, yTemp = yPos, zTemp = zPos;
    int xF = xIndex, yF = yIndex, zF = zIndex;
    float tauToBoundary = 0.0f, contribution;
    g->cnt.leRays++;
    if (!g->opt.useRussianRouletteForIntensity) {          /* INT:1729-1752 */
      tauToBoundary = orc_march(d, ID + 3 * i, &xTemp, &yTemp, &zTemp, &xF, &yF, &zF, 0, 0.0f, NULL, &g->cnt.leCrossings);
      if (tauToBoundary >= 0.0f) contribution = photonWeight * normalizedPhaseFunc * f_exp(-tauToBoundary);
      else contribution = 0.0f;
    } else {                                               /* INT:1753-1813, Iwabuchi (2006) */
      float u = orc_rng_real(r);
      float tauFree = -f_log(u > TINY32 ? u : TINY32);
      if (PI32 * normalizedPhaseFunc <= g->opt.zetaMin) {  /* Eq 13 */
        tauToBoundary = orc_march(d, ID + 3 * i, &xTemp, &yTemp, &zTemp, &xF, &yF, &zF, 1, tauFree, NULL, &g->cnt.leCrossings);
        float test = orc_rng_real(r);
        if (test <= PI32 * normalizedPhaseFunc / g->opt.zetaMin && zF >= zIndexMax)
          contribution = photonWeight * g->opt.zetaMin / PI32;
        else contribution = 0.0f;
      } else {                                             /* Eq 14 */
        float pn = PI32 * normalizedPhaseFunc;
        float tauMax = -f_log(g->opt.zetaMin / (TINY32 > pn ? TINY32 : pn));
        tauToBoundary = orc_march(d, ID + 3 * i, &xTemp, &yTemp, &zTemp, &xF, &yF, &zF, 1, tauMax, NULL, &g->cnt.leCrossings);
        if (zF >= zIndexMax && tauToBoundary >= 0.0f) {
          contribution = photonWeight * normalizedPhaseFunc * f_exp(-tauToBoundary);
        } else if (tauToBoundary >= 0.0f && zF < 1) {
          /* the ray left through the surface: the reference would trace on with zIndex = 0
           * (out of bounds, INT:1793); it cannot reach the top from there, so 0.           */
          contribution = 0.0f;
        } else if (tauToBoundary >= 0.0f) {
          tauToBoundary = orc_march(d, ID + 3 * i, &xTemp, &yTemp, &zTemp, &xF, &yF, &zF, 1, tauFree, NULL, &g->cnt.leCrossings);
          if (zF >= zIndexMax) contribution = photonWeight * g->opt.zetaMin / PI32;
          else contribution = 0.0f;
        } else {
          contribution = 0.0f;
        }
      }
    }
    if (g->opt.limitIntensityContributions) {              /* INT:1815-1826; component -1 has no slot: use 0 */
      if (contribution > g->opt.maxIntensityContribution) {
        int cslot = component < 0 ? 0 : component;
        g->intensityExcess[i + (size_t)numIntensityDirections * cslot] =
            g->intensityExcess[i + (size_t)numIntensityDirections * cslot] + contribution - g->opt.maxIntensityContribution;
        contribution = g->opt.maxIntensityContribution;
      }
    }
    contributions[i] = contribution; xIndexF[i] = xF; yIndexF[i] = yF;
    trace_event(g, r, photonNo, ORC_EV_LOCAL_ESTIMATE, xF, yF, zF, i + 1, 0, 0, scatteringOrder,
                contribution, tauToBoundary, 0.0, xTemp, yTemp, zTemp, ID + 3 * i);
  }
}

static void add_intensity(orc_integrator *g, const float *contributions, const int *xF, const int *yF, int comp) {
  const size_t cols = (size_t)g->nx * g->ny;                /* INT:535-540, 696-701, 785-790 */
  for (int i = 0; i < g->nDir; ++i) {
    size_t col = (size_t)(xF[i] - 1) + (size_t)g->nx * (size_t)(yF[i] - 1);
    g->intensity[col + cols * i] = g->intensity[col + cols * i] + contributions[i];
    size_t k = col + cols * ((size_t)i + (size_t)g->nDir * comp);
    g->intensityByComponent[k] = g->intensityByComponent[k] + contributions[i];
  }
}

/* computeRT INT:393-841 (ray-tracing branch) */
/* makePeriodic INT:1898-1917: the function result is DEFAULT REAL (quirk q15), the bounds are real(8) */
static float makePeriodic(double a, double aMin, double aMax) {
  float m = (float)a;
  for (;;) {
    if ((double)m <= aMax && (double)m > aMin) break;
    if ((double)m > aMax) m = (float)((double)m - (aMax - aMin));
    else if ((double)m == aMin) m = (float)aMax;
    else m = (float)((double)m + (aMax - aMin));
  }
  return m;
}

static int computeRT(orc_integrator *g, const orc_domain *d, orc_rng *r, orc_photons *ph,
                     int64_t numPhotonsPerBatch, int64_t *numPhotonsProcessed) {
  const int numX = d->nx, numY = d->ny, numZ = d->nz, numComps = d->nc;
  const double albedo = d->albedo;
  const double x0 = g->x0, xMax = g->xPosition[numX];
  const double y0 = g->y0, yMax = g->yPosition[numY];
  const double z0 = g->z0, zMax = g->zPosition[numZ];
  float *contributions = NULL; int *xIndexF = NULL, *yIndexF = NULL;
  if (g->computeIntensity) {
    contributions = (float *)malloc(sizeof(float) * g->nDir);
    xIndexF = (int *)malloc(sizeof(int) * g->nDir); yIndexF = (int *)malloc(sizeof(int) * g->nDir);
  }
  double *cumTable = (double *)malloc(sizeof(double) * (numComps + 1));
  int64_t nPhotons = 0; int nBad = 0;
  (void)numY;
  /* INT:445-448: maximum cross-section (Marchuk 1980) instead of ray tracing; maxExtinction is DEFAULT REAL */
  const int useRayTracing = g->opt.useRayTracing, useMaxCrossSection = !useRayTracing;
  float maxExtinction = 0.0f;
  if (useMaxCrossSection) {
    double m = d->totalExt[0];
    const size_t cells = (size_t)numX * numY * numZ;
    for (size_t i = 1; i < cells; ++i) if (d->totalExt[i] > m) m = d->totalExt[i];
    maxExtinction = (float)m;
  }

  while (nPhotons < numPhotonsPerBatch) {                                   /* photonLoop INT:463 */
    if (!(ph->current > 0 && ph->current <= ph->n)) break;                  /* morePhotonsExist ILL:540-546 */
    double xPos = ph->x[ph->current - 1], yPos = ph->y[ph->current - 1], zPos = ph->z[ph->current - 1];
    float mu = ph->mu[ph->current - 1], phi = ph->phi[ph->current - 1];
    ph->current++;                                                          /* getNextPhoton ILL:561-590 */
    int scatteringOrder = 0;
    float directionCosines[3];
    orc_make_direction_cosines(mu, phi, directionCosines);
    float photonWeight = 1.0f;
    nPhotons = nPhotons + 1;
    const int32_t photonNo = g->tracePhoton0 + (int32_t)(nPhotons - 1);
    int xIndex = 1, yIndex = 1, zIndex = 1;
    xPos = x0 + xPos * (xMax - x0);                                         /* INT:480-483 */
    yPos = y0 + yPos * (yMax - y0);
    findXYIndicies(g, xPos, yPos, &xIndex, &yIndex);
    if (g->zRegularlySpaced) {
      zPos = z0 + zPos * (zMax - z0);
      findZIndex(g, zPos, &zIndex);
    } else {                                                                /* INT:491-493 (sic: z0 in domain units) */
      double remainder = (zPos - z0) * numZ - floor((zPos - z0) * numZ);
      int zi = (int)floor((zPos - z0) * numZ) + 1;
      zIndex = zi < numZ ? zi : numZ;
      zPos = g->zPosition[zIndex - 1] + remainder * (g->zPosition[zIndex] - g->zPosition[zIndex - 1]);
    }
    g->cnt.photons++;
    trace_event(g, r, photonNo, ORC_EV_BIRTH, xIndex, yIndex, zIndex, 0, 0, 0, 0, photonWeight, 0.0f, 0.0,
                xPos, yPos, zPos, directionCosines);

    if (g->opt.LW_flag > 0.0f) {                                            /* INT:504-542 */
      if (zPos > 0.0) {
        size_t col = (size_t)(xIndex - 1) + (size_t)numX * (size_t)(yIndex - 1);
        g->fluxAbsorbed[col] = g->fluxAbsorbed[col] - 1.0f;
        size_t cell = CELL(d, xIndex, yIndex, zIndex);
        g->volumeAbsorption[cell] = g->volumeAbsorption[cell] - 1.0f;
      }
      if (g->computeIntensity) {
        computeIntensityContribution(g, d, photonWeight, xPos, yPos, zPos, xIndex, yIndex, zIndex,
                                     directionCosines, zPos == 0.0 ? 0 : -1, r, scatteringOrder, photonNo,
                                     contributions, xIndexF, yIndexF);
        add_intensity(g, contributions, xIndexF, yIndexF, 0);
      }
    }

    for (;;) {                                                              /* scatteringLoop INT:548 */
      if (r->exhausted) {
        trace_event(g, r, photonNo, ORC_EV_RN_EXHAUSTED, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                    photonWeight, 0.0f, 0.0, xPos, yPos, zPos, directionCosines);
        break;
      }
      float u = orc_rng_real(r);
      float tauToTravel = -f_log(u > TINY32 ? u : TINY32);                  /* INT:554 */
      double path = 0.0;
      if (useRayTracing) {
        float tauAccumulated = orc_march(d, directionCosines, &xPos, &yPos, &zPos, &xIndex, &yIndex, &zIndex,
                                         1, tauToTravel, &path, &g->cnt.crossings);  /* INT:559-561 */
        if (tauAccumulated < 0.0f) {                                        /* INT:562-563 */
          nBad = nBad + 1; g->cnt.bad++;
          trace_event(g, r, photonNo, ORC_EV_BAD, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                      photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
          break;
        }
      } else {                                                              /* INT:564-571; the cell indices are NOT refreshed (sic) */
        xPos = (double)makePeriodic(xPos + (double)(directionCosines[0] * tauToTravel / maxExtinction), x0, xMax);
        yPos = (double)makePeriodic(yPos + (double)(directionCosines[1] * tauToTravel / maxExtinction), y0, yMax);
        zPos = zPos + (double)(directionCosines[2] * tauToTravel / maxExtinction);
      }
      if (zPos >= zMax) {                                                   /* INT:573-617 */
        if (useMaxCrossSection) {                                           /* INT:578-585: trace back to the domain top */
          xPos = (double)makePeriodic(xPos - (double)directionCosines[0] * fabs((zPos - zMax) / (double)directionCosines[2]), x0, xMax);
          yPos = (double)makePeriodic(yPos - (double)directionCosines[1] * fabs((zPos - zMax) / (double)directionCosines[2]), y0, yMax);
          findXYIndicies(g, xPos, yPos, &xIndex, &yIndex);
        }
        size_t col = (size_t)(xIndex - 1) + (size_t)numX * (size_t)(yIndex - 1);
        g->fluxUp[col] = g->fluxUp[col] + photonWeight;
        g->cnt.topExits++;
        trace_event(g, r, photonNo, ORC_EV_EXIT_TOP, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                    photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
        break;
      } else if (zPos <= z0 + sp64(z0)) {                                   /* INT:619-702 */
        if (useMaxCrossSection) {                                           /* INT:624-631: trace back to the domain base */
          xPos = (double)makePeriodic(xPos - (double)directionCosines[0] * fabs((zPos - z0) / (double)directionCosines[2]), x0, xMax);
          yPos = (double)makePeriodic(yPos - (double)directionCosines[1] * fabs((zPos - z0) / (double)directionCosines[2]), y0, yMax);
          findXYIndicies(g, xPos, yPos, &xIndex, &yIndex);
        }
        zIndex = 1;
        zPos = z0 + sp64(z0);
        size_t col = (size_t)(xIndex - 1) + (size_t)numX * (size_t)(yIndex - 1);
        g->fluxDown[col] = g->fluxDown[col] + photonWeight;
        g->cnt.surfaceHits++;
        scatteringOrder = scatteringOrder + 1;
        for (;;) {                                                          /* INT:655-662 */
          mu = sqrtf(orc_rng_real(r));
          if (fabsf(mu) > 2.0f * TINY32) break;
          if (r->exhausted) break;
        }
        phi = 2.0f * PI32 * orc_rng_real(r);                                /* INT:663 */
        photonWeight = (float)((double)photonWeight * albedo);              /* INT:673 */
        if (photonWeight <= TINY32) {                                       /* INT:675 */
          trace_event(g, r, photonNo, ORC_EV_KILLED_SURFACE, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                      photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
          break;
        }
        orc_make_direction_cosines(mu, phi, directionCosines);
        trace_event(g, r, photonNo, ORC_EV_SURFACE, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                    photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
        if (g->computeIntensity) {                                          /* INT:680-702 */
          computeIntensityContribution(g, d, photonWeight, xPos, yPos, zPos, xIndex, yIndex, zIndex,
                                       directionCosines, 0, r, scatteringOrder, photonNo,
                                       contributions, xIndexF, yIndexF);
          add_intensity(g, contributions, xIndexF, yIndexF, 0);
        }
      } else {                                                              /* scattering event INT:703-821 */
        if (useMaxCrossSection) {                                           /* INT:709-710: "physical" or "mathematical" event */
          const float rnPhys = orc_rng_real(r);
          if (!((double)rnPhys < d->totalExt[CELL(d, xIndex, yIndex, zIndex)] / (double)maxExtinction)) {
            trace_event(g, r, photonNo, ORC_EV_NULL_COLLISION, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                        photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
            continue;
          }
        }
        scatteringOrder = scatteringOrder + 1;
        g->cnt.scatters++;
        if (d->totalExt[CELL(d, xIndex, yIndex, zIndex)] <= 0.0) {          /* INT:728-754 */
          if (xPos - g->xPosition[xIndex - 1] <= 0.0 && directionCosines[0] > 0.0f) {
            xPos = xPos - sp64(xPos);
            xIndex = xIndex - 1;
            if (xIndex <= 0) {
              xIndex = numX;
              xPos = g->xPosition[xIndex - 1];
              xPos = xPos - 2.0 * sp64(xPos);
            }
          }
          if (yPos - g->yPosition[yIndex - 1] <= 0.0 && directionCosines[1] > 0.0f) {
            yPos = yPos - sp64(yPos);
            yIndex = yIndex - 1;
            if (yIndex <= 0) {
              yIndex = numY;
              /* INT:743 (sic: xPosition(yIndex)); out of bounds in the reference when ny > nx+1, clamped here */
              yPos = g->xPosition[(yIndex < numX + 1 ? yIndex : numX + 1) - 1];
              yPos = yPos - 2.0 * sp64(yPos);
            }
          }
          if (zPos - g->zPosition[zIndex - 1] <= 0.0 && directionCosines[2] > 0.0f) {
            zPos = zPos - sp64(zPos);
            zIndex = zIndex - 1;
          }
        }
        cumTable[0] = 0.0;                                                  /* INT:759-760 */
        for (int c = 1; c <= numComps; ++c) cumTable[c] = d->cumExt[CELLC(d, xIndex, yIndex, zIndex, c)];
        int component = orc_findIndexMixed(orc_rng_real(r), cumTable, numComps + 1, 0);
        float ssa = (float)d->ssa[CELLC(d, xIndex, yIndex, zIndex, component)];   /* INT:764 */
        if ((double)ssa < 1.0) {                                            /* INT:765-771 */
          double absorbed = (double)photonWeight * (1.0 - (double)ssa);
          size_t col = (size_t)(xIndex - 1) + (size_t)numX * (size_t)(yIndex - 1);
          g->fluxAbsorbed[col] = (float)((double)g->fluxAbsorbed[col] + absorbed);
          size_t cell = CELL(d, xIndex, yIndex, zIndex);
          g->volumeAbsorption[cell] = (float)((double)g->volumeAbsorption[cell] + absorbed);
          photonWeight = photonWeight * ssa;
        }
        if (g->computeIntensity) {                                          /* INT:776-800 */
          computeIntensityContribution(g, d, photonWeight, xPos, yPos, zPos, xIndex, yIndex, zIndex,
                                       directionCosines, component, r, scatteringOrder, photonNo,
                                       contributions, xIndexF, yIndexF);
          add_intensity(g, contributions, xIndexF, yIndexF, component);
        }
        if (g->opt.useRussianRoulette && photonWeight < g->opt.RussianRouletteW / 2.0f) {   /* INT:805-811 */
          if (orc_rng_real(r) >= photonWeight / g->opt.RussianRouletteW) photonWeight = 0.0f;
          else photonWeight = g->opt.RussianRouletteW;
        }
        int phaseFunctionIndex = d->phaseIdx[CELLC(d, xIndex, yIndex, zIndex, component)];  /* INT:816 */
        if (photonWeight <= TINY32) {                                       /* INT:812 */
          g->cnt.rouletteKills++;
          trace_event(g, r, photonNo, ORC_EV_KILLED_ROULETTE, xIndex, yIndex, zIndex, component,
                      phaseFunctionIndex, 0, scatteringOrder, photonWeight, tauToTravel, path,
                      xPos, yPos, zPos, directionCosines);
          break;
        }
        int c = component - 1, k = 0;
        float scatteringAngle = computeScatteringAngle(orc_rng_real(r),
                                    d->inv[c] + (size_t)(phaseFunctionIndex - 1) * d->invS[c], d->invS[c], &k);  /* INT:817-818 */
        next_direct(r, f_cos(scatteringAngle), directionCosines);           /* INT:819 */
        trace_event(g, r, photonNo, ORC_EV_SCATTER, xIndex, yIndex, zIndex, component, phaseFunctionIndex, k,
                    scatteringOrder, photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
      }
    }
  }
  free(contributions); free(xIndexF); free(yIndexF); free(cumTable);
  g->tracePhoton0 += (int32_t)nPhotons;
  g->cnt.rnDrawn = r->ndrawn;
  if (nPhotons > 0) { *numPhotonsProcessed = nPhotons; return 0; }          /* INT:831-839 */
  *numPhotonsProcessed = 0;
  return 1;
}

/* computeRadiativeTransfer INT:209-391 */
int orc_compute_radiative_transfer(orc_integrator *g, const orc_domain *d, orc_rng *r,
                                   orc_photons *p, int64_t numPhotonsPerBatch,
                                   int normalise, int64_t *numPhotonsProcessed) {
  const int numX = g->nx, numY = g->ny, numZ = g->nz, numComponents = g->nc;
  const size_t cols = (size_t)numX * numY;
  memset(g->fluxUp, 0, sizeof(float) * cols);                               /* INT:247-261 */
  memset(g->fluxDown, 0, sizeof(float) * cols);
  memset(g->fluxAbsorbed, 0, sizeof(float) * cols);
  memset(g->volumeAbsorption, 0, sizeof(float) * cols * numZ);
  if (g->intensity) memset(g->intensity, 0, sizeof(float) * cols * g->nDir);
  if (g->intensityByComponent) memset(g->intensityByComponent, 0, sizeof(float) * cols * g->nDir * (numComponents + 1));
  if (g->intensityExcess) memset(g->intensityExcess, 0, sizeof(float) * g->nDir * (numComponents + 1));

  int rc = computeRT(g, d, r, p, numPhotonsPerBatch, numPhotonsProcessed);  /* INT:291-292 */
  if (rc) return rc;

  if (g->computeIntensity && g->opt.limitIntensityContributions) {          /* INT:294-322 */
    for (int j = 0; j <= numComponents; ++j)
      for (int dd = 0; dd < g->nDir; ++dd) {
        float excess = g->intensityExcess[dd + (size_t)g->nDir * j];
        if (excess > 0.0f) {
          float *byc = g->intensityByComponent + cols * ((size_t)dd + (size_t)g->nDir * j);
          float s = 0.0f;
          for (size_t i = 0; i < cols; ++i) s = s + byc[i];
          for (size_t i = 0; i < cols; ++i) {
            g->intensity[i + cols * dd] = g->intensity[i + cols * dd] + (byc[i] / s) * excess;
          }
          for (size_t i = 0; i < cols; ++i) byc[i] = byc[i] + (byc[i] / s) * excess;
        }
      }
  }
  if (!normalise) return 0;

  float *numPhotonsPerColumn = (float *)malloc(sizeof(float) * cols);       /* INT:328-343 */
  if (g->xyRegularlySpaced) {
    float v = (float)(*numPhotonsProcessed) / (float)(numX * numY);
    for (size_t i = 0; i < cols; ++i) numPhotonsPerColumn[i] = v;
  } else {
    for (int j = 0; j < numY; ++j)
      for (int i = 0; i < numX; ++i) {
        float frac = (float)(((g->yPosition[j + 1] - g->yPosition[j]) * (g->xPosition[i + 1] - g->xPosition[i])) /
                             ((g->xPosition[numX] - g->xPosition[0]) * (g->yPosition[numY] - g->yPosition[0])));
        numPhotonsPerColumn[i + (size_t)numX * j] = frac * (float)(*numPhotonsProcessed);
      }
  }
  for (size_t i = 0; i < cols; ++i) {                                       /* INT:348-350 */
    g->fluxUp[i] = g->fluxUp[i] / numPhotonsPerColumn[i];
    g->fluxDown[i] = g->fluxDown[i] / numPhotonsPerColumn[i];
    g->fluxAbsorbed[i] = g->fluxAbsorbed[i] / numPhotonsPerColumn[i];
  }
  for (int k = 0; k < numZ; ++k)                                            /* INT:361-364 */
    for (size_t i = 0; i < cols; ++i)
      g->volumeAbsorption[i + cols * k] = (float)((double)g->volumeAbsorption[i + cols * k] /
          ((double)numPhotonsPerColumn[i] * (g->zPosition[k + 1] - g->zPosition[k]) * (double)1000.0f));
  if (g->computeIntensity) {                                                /* INT:369-379 (component 0 left as is) */
    for (int dd = 0; dd < g->nDir; ++dd)
      for (size_t i = 0; i < cols; ++i)
        g->intensity[i + cols * dd] = g->intensity[i + cols * dd] / numPhotonsPerColumn[i];
    for (int j = 1; j <= numComponents; ++j)
      for (int dd = 0; dd < g->nDir; ++dd)
        for (size_t i = 0; i < cols; ++i) {
          size_t k = i + cols * ((size_t)dd + (size_t)g->nDir * j);
          g->intensityByComponent[k] = g->intensityByComponent[k] / numPhotonsPerColumn[i];
        }
  }
  free(numPhotonsPerColumn);
  return 0;
}

/* Fixed-random-number single-photon trace harness (north-star criterion (a)): photon p is
 * born from and transported with the injected numbers rn[p*stride .. p*stride+stride-1]
 * (source draws first, then the computeRT order).  Tallies are raw sums over all photons. */
int64_t orc_trace_photons(orc_integrator *g, const orc_domain *d,
                          int source, float solarMu, float solarAzimuthDeg,
                          double fracAtmsPower, const double *voxelCDF,
                          int64_t nPhotons, const float *rn, int64_t stride) {
  const size_t cols = (size_t)g->nx * g->ny;
  memset(g->fluxUp, 0, sizeof(float) * cols);
  memset(g->fluxDown, 0, sizeof(float) * cols);
  memset(g->fluxAbsorbed, 0, sizeof(float) * cols);
  memset(g->volumeAbsorption, 0, sizeof(float) * cols * g->nz);
  if (g->intensity) memset(g->intensity, 0, sizeof(float) * cols * g->nDir);
  if (g->intensityByComponent) memset(g->intensityByComponent, 0, sizeof(float) * cols * g->nDir * (g->nc + 1));
  if (g->intensityExcess) memset(g->intensityExcess, 0, sizeof(float) * g->nDir * (g->nc + 1));
  int64_t total = 0;
  g->tracePhoton0 = 0;
  for (int64_t p = 0; p < nPhotons; ++p) {
    orc_rng rng;
    orc_rng_init_injected(&rng, rn + p * stride, stride);
    orc_photons *ph = source == 1
        ? orc_photons_bbemission(fracAtmsPower, voxelCDF, g->nx, g->ny, g->nz, 1, &rng)
        : orc_photons_directional(solarMu, solarAzimuthDeg, 1, &rng);
    int64_t done = 0;
    computeRT(g, d, &rng, ph, 1, &done);
    orc_photons_free(ph);
    total += done;
  }
  return total;
}

/* Fortran sum() of a real array: straightforward left-to-right accumulation in f32 */
static float sum_f32(const float *a, size_t n) { float s = 0.0f; for (size_t i = 0; i < n; ++i) s = s + a[i]; return s; }

/* reportResults INT:845-1042 */
void orc_report_results(const orc_integrator *g,
                        float *meanFluxUp, float *meanFluxDown, float *meanFluxAbsorbed,
                        float *fluxUp, float *fluxDown, float *fluxAbsorbed,
                        float *absorbedProfile, float *volumeAbsorption,
                        float *meanIntensity, float *intensity) {
  const size_t cols = (size_t)g->nx * g->ny;
  const int numColumns = (int)cols;
  if (meanFluxUp) *meanFluxUp = sum_f32(g->fluxUp, cols) / (float)numColumns;                 /* INT:881-884 */
  if (meanFluxDown) *meanFluxDown = sum_f32(g->fluxDown, cols) / (float)numColumns;
  if (meanFluxAbsorbed) *meanFluxAbsorbed = sum_f32(g->fluxAbsorbed, cols) / (float)numColumns;
  if (fluxUp) memcpy(fluxUp, g->fluxUp, sizeof(float) * cols);
  if (fluxDown) memcpy(fluxDown, g->fluxDown, sizeof(float) * cols);
  if (fluxAbsorbed) memcpy(fluxAbsorbed, g->fluxAbsorbed, sizeof(float) * cols);
  if (absorbedProfile)                                                                      /* INT:966 */
    for (int k = 0; k < g->nz; ++k)
      absorbedProfile[k] = sum_f32(g->volumeAbsorption + cols * k, cols) / (float)numColumns;
  if (volumeAbsorption) memcpy(volumeAbsorption, g->volumeAbsorption, sizeof(float) * cols * g->nz);
  if (meanIntensity && g->intensity)                                                        /* INT:989-991 */
    for (int dd = 0; dd < g->nDir; ++dd)
      meanIntensity[dd] = sum_f32(g->intensity + cols * dd, cols) / (float)numColumns;
  if (intensity && g->intensity) memcpy(intensity, g->intensity, sizeof(float) * cols * g->nDir);
}

/* ------------------------------------------------------------------------------------- */
/* driver batch loop and statistics, DRV:949-1052                                         */
/* ------------------------------------------------------------------------------------- */
static void add_moments(double *stats, const float *x, size_t n, double numPhotonsProcessed) {
  for (size_t i = 0; i < n; ++i) {
    double v = (double)x[i];
    stats[i] = stats[i] + v * numPhotonsProcessed;                          /* first moment  */
    stats[i + n] = stats[i + n] + numPhotonsProcessed * (v * v);            /* second moment */
  }
}

int64_t orc_run_batches(orc_integrator *g, const orc_domain *d,
                        int source, float solarMu, float solarAzimuthDeg,
                        double fracAtmsPower, const double *voxelCDF,
                        int iseed, int rank, int thread,
                        int64_t numBatches, int64_t numPhotonsPerBatch, orc_stats *st) {
  const size_t cols = (size_t)g->nx * g->ny;
  orc_rng rng;
  uint32_t key[3] = {(uint32_t)iseed, (uint32_t)rank, (uint32_t)thread};   /* DRV:901 */
  orc_rng_init_array(&rng, key, 3);
  float meanFluxUp, meanFluxDown, meanFluxAbsorbed;
  float *fluxUp = (float *)malloc(sizeof(float) * cols), *fluxDown = (float *)malloc(sizeof(float) * cols);
  float *fluxAbsorbed = (float *)malloc(sizeof(float) * cols);
  float *absorbedProfile = (float *)malloc(sizeof(float) * g->nz);
  float *absorbedVolume = st->absorbedVolumeStats ? (float *)malloc(sizeof(float) * cols * g->nz) : NULL;
  int64_t total = 0;
  for (int64_t b = 0; b < numBatches; ++b) {
    orc_photons *ph = source == 1
        ? orc_photons_bbemission(fracAtmsPower, voxelCDF, g->nx, g->ny, g->nz, numPhotonsPerBatch, &rng)
        : orc_photons_directional(solarMu, solarAzimuthDeg, numPhotonsPerBatch, &rng);
    int64_t done = 0;
    int rc = orc_compute_radiative_transfer(g, d, &rng, ph, numPhotonsPerBatch, 1, &done);
    orc_photons_free(ph);
    if (rc) break;
    total += done;
    orc_report_results(g, &meanFluxUp, &meanFluxDown, &meanFluxAbsorbed, fluxUp, fluxDown, fluxAbsorbed,
                       absorbedProfile, absorbedVolume, NULL, NULL);
    double np = (double)done;                                              /* DRV:1023-1052 */
    add_moments(st->meanFluxUpStats, &meanFluxUp, 1, np);
    add_moments(st->meanFluxDownStats, &meanFluxDown, 1, np);
    add_moments(st->meanFluxAbsorbedStats, &meanFluxAbsorbed, 1, np);
    add_moments(st->fluxUpStats, fluxUp, cols, np);
    add_moments(st->fluxDownStats, fluxDown, cols, np);
    add_moments(st->fluxAbsorbedStats, fluxAbsorbed, cols, np);
    add_moments(st->absorbedProfileStats, absorbedProfile, g->nz, np);
    if (absorbedVolume) add_moments(st->absorbedVolumeStats, absorbedVolume, cols * g->nz, np);
    if (g->computeIntensity && st->radianceStats) add_moments(st->radianceStats, g->intensity, cols * g->nDir, np);
  }
  free(fluxUp); free(fluxDown); free(fluxAbsorbed); free(absorbedProfile); free(absorbedVolume);
  return total;
}

/* DRV:1188-1228: stats holds n first moments followed by n second moments on entry;
 * on exit n means followed by n standard errors.                                       */
void orc_finalise_stats(double *stats, int64_t n, double solarFlux,
                        int64_t totalNumPhotons, int64_t batchesCompleted) {
  for (int64_t i = 0; i < n; ++i) {
    double m1 = solarFlux * stats[i] / (double)totalNumPhotons;
    double m2 = solarFlux * stats[i + n] / (double)totalNumPhotons;
    m2 = solarFlux * m2;
    double var = m2 - m1 * m1;
    if (var < 0.0) var = 0.0;
    stats[i] = m1;
    stats[i + n] = sqrt(var / (double)(batchesCompleted - 1));
  }
}

/* ------------------------------------------------------------------------------------- */
/* read_SSPTable inner loops (OPT:204-299) + getOpticalPropertiesByComponent (OPT:1022-1061) */
/* ------------------------------------------------------------------------------------- */
int orc_assemble_optics(int nx, int ny, int nz, int nPhys, const double *massConc, const double *Reff,
                        const double *numConc, int nc, const orc_component *comps, int setup,
                        double *totalExt, double *cumExt, double *ssa, int32_t *phaseIdx) {
  const size_t cols = (size_t)nx * ny, cells = cols * nz;
  int rc = 0;
  for (int c = 0; c < nc; ++c) {
    const orc_component *q = &comps[c];
    double *cumC = cumExt + cells * c, *ssaC = ssa + cells * c;
    int32_t *idxC = phaseIdx + cells * c;
    for (size_t i = 0; i < cells; ++i) { cumC[i] = 0.0; ssaC[i] = 0.0; idxC[i] = 0; }
    const int nLev = q->kind == ORC_COMP_VOLEXT ? nz : q->nTable;
    if (q->zLevelBase < 1 || q->zLevelBase + nLev - 1 > nz) return 2;              /* OPT:706-708 */
    if (q->kind == ORC_COMP_VOLEXT) {
      const int nReff = q->nTable;
      double *key8 = (double *)malloc(sizeof(double) * nReff);                     /* REAL(key(:),8), OPT:269 */
      float kmin = q->key[0], kmax = q->key[0];
      for (int i = 0; i < nReff; ++i) {
        key8[i] = (double)q->key[i];
        if (q->key[i] < kmin) kmin = q->key[i];
        if (q->key[i] > kmax) kmax = q->key[i];
      }
      for (size_t cell = 0; cell < cells; ++cell) {                                /* OPT:260-292 */
        const double m = massConc[(size_t)(q->physIndex - 1) + (size_t)nPhys * cell];
        const double re = Reff[(size_t)(q->physIndex - 1) + (size_t)nPhys * cell];
        double e = 0.0, w = 0.0; int32_t pi = 1;                                   /* OPT:253-255 defaults */
        if (m > 0.0 && re < (double)kmax && re >= (double)kmin) {
          const int il = orc_findIndexDouble(re, key8, nReff, 0);
          const double f = (re - (double)q->key[il - 1]) / (double)(q->key[il] - q->key[il - 1]);   /* OPT:272 */
          e = m * ((1 - f) * q->ext[il - 1] + f * q->ext[il]);
          w = (1 - f) * q->ssa[il - 1] + f * q->ssa[il];
          if (!setup) pi = f < 0.5 ? il : il + 1;                                  /* OPT:279-285 */
        } else if (m > 0.0) {
          rc = 1;                                                                  /* OPT:288-289 */
        }
        cumC[cell] = e; ssaC[cell] = w; idxC[cell] = pi;
      }
      free(key8);
    } else {
      for (int k = 0; k < nLev; ++k) {
        const int iz = q->zLevelBase - 1 + k;
        double e, w; int32_t pi;
        if (q->kind == ORC_COMP_ABSXSEC) { e = q->ext[k] * numConc[k] * 1000.0; w = 0.0; pi = 1; }   /* OPT:223-227 */
        else { e = q->ext[k]; w = q->ssa[k]; pi = q->phaseIdx[k]; }
        for (size_t j = 0; j < cols; ++j) {                                        /* spread(...), OPT:1038-1046 */
          cumC[j + cols * iz] = e; ssaC[j + cols * iz] = w; idxC[j + cols * iz] = pi;
        }
      }
    }
  }
  for (int c = 1; c < nc; ++c)                                                     /* OPT:1055-1057 */
    for (size_t i = 0; i < cells; ++i) cumExt[i + cells * c] = cumExt[i + cells * c] + cumExt[i + cells * (c - 1)];
  for (size_t i = 0; i < cells; ++i) totalExt[i] = cumExt[i + cells * (nc - 1)];
  for (int c = 0; c < nc; ++c)                                                     /* OPT:1059-1061 */
    for (size_t i = 0; i < cells; ++i)
      if (totalExt[i] > DBL_MIN) cumExt[i + cells * c] = cumExt[i + cells * c] / totalExt[i];
  return rc;
}

/* getFrequencyDistrNEW EMI:552-573 */
void orc_frequency_distribution(int numLambda, const double *CDF, int64_t totalPhotons, orc_rng *r, int64_t *distribution) {
  for (int i = 0; i < numLambda; ++i) distribution[i] = 0;
  for (int64_t n = 0; n < totalPhotons; ++n) {
    const float RN = orc_rng_real(r);
    const int i = orc_findCDFIndex(RN, CDF, numLambda, 1);
    distribution[i - 1] = distribution[i - 1] + 1;
  }
}

/* findIndexReal NUM:150-204: the single-precision twin of findIndexDouble */
static int findIndexReal(float value, const float *table, int n, int firstGuess) {
#define T1(i) table[(i) - 1]
  int lowerBound, upperBound, midPoint, increment;
  if (firstGuess > 0) {
    lowerBound = firstGuess; increment = 1;
    for (;;) {
      upperBound = lowerBound + increment < n ? lowerBound + increment : n;
      if (lowerBound == n || (T1(lowerBound) <= value && T1(upperBound) > value)) break;
      if (T1(lowerBound) > value) {
        upperBound = lowerBound;
        lowerBound = upperBound - increment > 1 ? upperBound - increment : 1;
      } else {
        lowerBound = upperBound;
      }
      increment *= 2;
    }
  } else {
    lowerBound = 0; upperBound = n;
  }
  for (;;) {
    if (lowerBound == n || upperBound <= lowerBound + 1) break;
    midPoint = (lowerBound + upperBound) / 2;
    if (value >= T1(midPoint)) lowerBound = midPoint; else upperBound = midPoint;
  }
  return lowerBound;
#undef T1
}

/* computeInversePhaseFunction INV:113-168 (arguments of acos outside [-1, 1] are clamped, as the host mirror does) */
void orc_inverse_phase_function(int nAngles, const float *mus, const float *values, int nSteps, float *inverseTable) {
  float *cdf = (float *)malloc(sizeof(float) * nAngles);
  int *indicies = (int *)malloc(sizeof(int) * nSteps);
  cdf[0] = 0.0f;
  for (int i = 1; i < nAngles; ++i)                                            /* INV:118-121 */
    cdf[i] = cdf[i - 1] + (mus[i] - mus[i - 1]) * 0.5f * (values[i] + values[i - 1]);
  const float last = cdf[nAngles - 1];
  for (int i = 0; i < nAngles; ++i) cdf[i] = cdf[i] / last;                    /* INV:125 */
  indicies[0] = findIndexReal(0.0f, cdf, nAngles, 0);                          /* INV:129-135 */
  for (int i = 2; i <= nSteps; ++i) {
    const float probabilityValue = (float)(i - 1) / (float)(nSteps - 1);
    indicies[i - 1] = findIndexReal(probabilityValue, cdf, nAngles, indicies[i - 2]);
  }
  for (int i = 1; i <= nSteps - 1; ++i) {                                      /* INV:137-167 */
    const float p = (float)(i - 1) / (float)(nSteps - 1);
    const int k = indicies[i - 1];                                             /* 1-based */
    const float c0 = cdf[k - 1], c1 = cdf[k], v0 = values[k - 1], v1 = values[k], m0 = mus[k - 1], m1 = mus[k];
    float arg;
    if (c1 - c0 <= sp32(c0)) {
      arg = m0;
    } else if (fabsf(v0 - v1) <= sp32(v0)) {
      arg = m0 + (m1 - m0) * (p - c0) / (c1 - c0);
    } else {
      arg = m0 + (m1 - m0) / (v0 - v1) * (v0 - sqrtf(((c1 - p) * (v0 * v0) + (p - c0) * (v1 * v1)) / (c1 - c0)));
    }
    arg = arg < -1.0f ? -1.0f : (arg > 1.0f ? 1.0f : arg);
    inverseTable[i - 1] = f_acos(arg);
  }
  inverseTable[nSteps - 1] = 0.0f;                                             /* INV:168 */
  free(cdf); free(indicies);
}

/* tabulateForwardPhaseFunctions OPT:1912-1913 + getPhaseFunctionValues SPF:480-498 + computeLegendrePolynomials NUM:187-205 */
void orc_forward_phase_function(int nCoef, const float *legendreCoefficients, int nS, float *values) {
  for (int i = 0; i < nS; ++i) {
    if (nCoef == 0) { values[i] = 0.5f; continue; }                               /* SPF:486-491 */
    const float angle = (float)i / (float)(nS - 1) * PI32;
    const float mu = f_cos(angle);
    float pm1 = 1.0f, p = mu;                                                      /* P_0, P_1 */
    float value = 1.0f * pm1;                                                      /* (2*0+1) * 1 * P_0 */
    value = value + (legendreCoefficients[0] * 3.0f) * p;
    for (int l = 1; l < nCoef; ++l) {
      const float pn = (((float)(2 * l + 1) * mu) * p - (float)l * pm1) / (float)(l + 1);
      pm1 = p; p = pn;
      value = value + (legendreCoefficients[l] * (float)(2 * (l + 1) + 1)) * p;
    }
    values[i] = value;
  }
}

/* ---- the rest of the table producers: Lobatto nodes, phase-function values for either storage kind,
 *      the inversion inputs of a Legendre-stored phase function, the hybrid (Gaussian forward peak) tables ---- */

/* computeLegendrePolynomials NUM:187-205 at one mu: P[0..maxL] by upward recursion in single precision */
static void legendre_polynomials(int maxL, float mu, float *P) {
  P[0] = 1.0f;
  P[1] = mu;
  for (int l = 1; l <= maxL - 1; ++l)
    P[l + 1] = (((float)(2 * l + 1) * mu) * P[l] - (float)l * P[l - 1]) / (float)(l + 1);
}

/* computeLobattoTerms NUM:27-114: abscissas and weights of n-point Lobatto quadrature on [-1, 1], Newton's method on
 * the zeros of P'_{n-1}.  The iteration is global as in the reference: every pass recomputes the polynomials at every
 * trial point, only the points still moving are updated, the loop ends when none moves (or after 26 passes) -- so the
 * weights are formed from the polynomials of the last pass.                                                       */
void orc_lobatto_terms(int n, float *mus, float *weights) {
  const float relativeAccuracy = 3.0f;
  const int maxIterations = 25;
  const float pi = f_acos(-1.0f);
  const int nTerms = n, midPoint = (nTerms + 1) / 2, m = midPoint - 1;
  float *trial = (float *)malloc(sizeof(float) * (m + 1)), *last = (float *)malloc(sizeof(float) * (m + 1));
  float *d1 = (float *)malloc(sizeof(float) * (m + 1)), *d2 = (float *)malloc(sizeof(float) * (m + 1));
  float *P = (float *)malloc(sizeof(float) * (size_t)(nTerms + 1) * (m + 1));        /* P[k * (nTerms + 1) + l] */
#define LP(l, k) P[(size_t)(k) * (nTerms + 1) + (l)]
  const float c1 = (nTerms % 2 == 1) ? 1.0f : 0.5f;
  for (int k = 0; k < m; ++k) trial[k] = f_sin(pi * ((float)(k + 1) - c1) / ((float)nTerms - 1.0f + 0.5f));
  for (int k = 0; k < m; ++k) {                                                        /* first Newton step */
    legendre_polynomials(nTerms - 1, trial[k], &LP(0, k));
    d1[k] = (float)(nTerms - 1) * (trial[k] * LP(nTerms - 1, k) - LP(nTerms - 2, k)) / (trial[k] * trial[k] - 1.0f);
    d2[k] = (2.0f * trial[k] * d1[k] - ((float)(nTerms * (nTerms - 1)) * LP(nTerms - 1, k))) / (1.0f - trial[k] * trial[k]);
    last[k] = trial[k];
    trial[k] = trial[k] - d1[k] / d2[k];
  }
  int i = 0;
  for (;;) {
    int moving = 0;
    for (int k = 0; k < m; ++k) if (fabsf(trial[k] - last[k]) > relativeAccuracy * sp32(trial[k])) moving = 1;
    if (!moving) break;
    for (int k = 0; k < m; ++k) legendre_polynomials(nTerms - 1, trial[k], &LP(0, k));
    for (int k = 0; k < m; ++k)
      if (fabsf(trial[k] - last[k]) > relativeAccuracy * sp32(trial[k])) {
        d1[k] = (float)(nTerms - 1) * (trial[k] * LP(nTerms - 1, k) - LP(nTerms - 2, k)) / (trial[k] * trial[k] - 1.0f);
        d2[k] = (2.0f * trial[k] * d1[k] - ((float)(nTerms * (nTerms - 1)) * LP(nTerms - 1, k))) / (1.0f - trial[k] * trial[k]);
        last[k] = trial[k];
        trial[k] = trial[k] - d1[k] / d2[k];
      }
    i = i + 1;
    if (i > maxIterations) break;
  }
  mus[0] = -1.0f;
  weights[0] = 2.0f / (float)(nTerms * (nTerms - 1));
  for (int k = 0; k < m; ++k) {                                                        /* mus(midPoint:2:-1) = -trialMus(:) */
    mus[midPoint - 1 - k] = -trial[k];
    weights[midPoint - 1 - k] = 2.0f / ((float)(nTerms * (nTerms - 1)) * (LP(nTerms - 1, k) * LP(nTerms - 1, k)));
  }
  if (nTerms % 2 == 0) {                                                               /* NUM:104-110, right-hand sides first */
    for (int k = 0; k < midPoint; ++k) { mus[midPoint + k] = -mus[midPoint - 1 - k]; weights[midPoint + k] = weights[midPoint - 1 - k]; }
  } else {
    float *tm = (float *)malloc(sizeof(float) * midPoint), *tw = (float *)malloc(sizeof(float) * midPoint);
    for (int k = 0; k < midPoint; ++k) { tm[k] = -mus[midPoint - 1 - k]; tw[k] = weights[midPoint - 1 - k]; }
    for (int k = 0; k < midPoint; ++k) { mus[midPoint - 1 + k] = tm[k]; weights[midPoint - 1 + k] = tw[k]; }
    free(tm); free(tw);
  }
#undef LP
  free(trial); free(last); free(d1); free(d2); free(P);
}

/* getPhaseFunctionValues_one SPF:448-531 for either storage kind: nCoef >= 0 Legendre coefficients (nStored = 0), or
 * nStored angle / value pairs (interpolated linearly in the cosine of the angle, findIndex NUM:206-260 without a
 * first guess, the last interval guarded as SPF:511-520)                                                          */
void orc_phase_function_values(int nCoef, const float *legendreCoefficients, int nStored, const float *storedAngle,
                               const float *storedValue, int nAngles, const float *scatteringAngle, float *value) {
  if (nStored <= 0) {
    float *P = (float *)malloc(sizeof(float) * (nCoef + 2));
    for (int i = 0; i < nAngles; ++i) {
      if (nCoef == 0) { value[i] = 0.5f; continue; }                                   /* SPF:486-491 */
      legendre_polynomials(nCoef, f_cos(scatteringAngle[i]), P);
      float v = 0.0f;                                                                  /* matmul, accumulated in order */
      for (int l = 0; l <= nCoef; ++l) v = v + ((l == 0 ? 1.0f : legendreCoefficients[l - 1]) * (float)(2 * l + 1)) * P[l];
      value[i] = v;
    }
    free(P);
    return;
  }
  for (int i = 0; i < nAngles; ++i) {
    int idx = findIndexReal(scatteringAngle[i], storedAngle, nStored, 0);
    idx = idx < 1 ? 1 : (idx > nStored ? nStored : idx);
    int ip1 = idx + 1;
    float dMu;
    if (idx < nStored) dMu = f_cos(storedAngle[ip1 - 1]) - f_cos(storedAngle[idx - 1]);
    else { dMu = FLT_MAX; ip1 = idx; }
    const float w = 1.0f - (f_cos(scatteringAngle[i]) - f_cos(storedAngle[idx - 1])) / dMu;
    value[i] = w * storedValue[idx - 1] + (1.0f - w) * storedValue[ip1 - 1];
  }
}

/* INV:97-112: what computeInversePhaseFunction inverts for a Legendre-stored phase function -- its values at the
 * max(nMoments, 2) Lobatto abscissas, increasing in mu.  mus, values: max(nCoef, 2) entries each.                 */
void orc_inversion_inputs_legendre(int nCoef, const float *legendreCoefficients, float *mus, float *values) {
  const int n = nCoef > 2 ? nCoef : 2;
  float *w = (float *)malloc(sizeof(float) * n), *ang = (float *)malloc(sizeof(float) * n), *v = (float *)malloc(sizeof(float) * n);
  orc_lobatto_terms(n, mus, w);
  for (int i = 0; i < n; ++i) ang[i] = f_acos(mus[n - 1 - i]);                         /* acos(mus(nAngles:1:-1)) */
  orc_phase_function_values(nCoef, legendreCoefficients, 0, NULL, NULL, n, ang, v);
  for (int i = 0; i < n; ++i) values[i] = v[n - 1 - i];                                /* values = values(nAngles:1:-1) */
  free(w); free(ang); free(v);
}

/* computeNormalization OPT:2027-2050 (transitionIndex 1-based; dot products accumulated in order, single precision) */
static float hybrid_normalization(int nAngles, const float *angleCosines, const float *values, const float *gaussianValues, int t) {
  float integralGaus = 0.0f, integralOrig = 0.0f;
  for (int j = 1; j <= t - 1; ++j)
    integralGaus = integralGaus + (0.5f * (gaussianValues[j - 1] + gaussianValues[j])) * (angleCosines[j - 1] - angleCosines[j]);
  for (int j = t; j <= nAngles - 1; ++j)
    integralOrig = integralOrig + (0.5f * (values[j - 1] + values[j])) * (angleCosines[j - 1] - angleCosines[j]);
  if (integralOrig >= 2.0f) return 1.0f / integralGaus;
  return (2.0f - integralOrig) / integralGaus;
}
static float hybrid_diff(int nAngles, const float *angleCosines, const float *values, const float *gaussianValues, int t) {   /* OPT:2011-2025 */
  const float P0 = hybrid_normalization(nAngles, angleCosines, values, gaussianValues, t);
  return P0 * gaussianValues[t - 1] - values[t - 1];
}

/* computeHybridPhaseFunctions OPT:1936-2009 for one table entry: a Gaussian of gaussianWidth degrees replaces the
 * forward peak, continuous with the original at the transition angle (hunt + bisection on the difference), scaled so
 * that the whole stays normalised.  Returns the transition index (0: the entry is left as it is).                  */
int orc_hybrid_phase_function(int nAngles, const float *angles, const float *values, float gaussianWidth, float *newValues) {
  float *angleCosines = (float *)malloc(sizeof(float) * nAngles), *gaussianValues = (float *)malloc(sizeof(float) * nAngles);
  const float width = gaussianWidth * PI32 / 180.0f;
  for (int i = 0; i < nAngles; ++i) {
    angleCosines[i] = f_cos(angles[i]);
    const float q = angles[i] / width;
    gaussianValues[i] = f_exp(-(q * q));
  }
  for (int i = 0; i < nAngles; ++i) newValues[i] = values[i];
  int transitionIndex = 0;
  int lowerBound = findIndexReal(width, angles, nAngles, 0) + 1;
  if (lowerBound < nAngles - 2) {
    float lowDiff = hybrid_diff(nAngles, angleCosines, values, gaussianValues, lowerBound), upDiff = 0.0f;
    int increment = 1, upperBound = lowerBound, root = 1;
    for (;;) {                                                                         /* hunt */
      upperBound = lowerBound + increment < nAngles - 1 ? lowerBound + increment : nAngles - 1;
      upDiff = hybrid_diff(nAngles, angleCosines, values, gaussianValues, upperBound);
      if (lowerBound == nAngles - 1) { root = 0; break; }
      if (lowDiff * upDiff < 0.0f) break;
      lowerBound = upperBound; lowDiff = upDiff; increment = increment * 2;
    }
    if (root) {
      while (upperBound > lowerBound + 1) {                                            /* bisection */
        const int midPoint = (lowerBound + upperBound) / 2;
        const float midDiff = hybrid_diff(nAngles, angleCosines, values, gaussianValues, midPoint);
        if (midDiff * upDiff < 0.0f) { lowerBound = midPoint; lowDiff = midDiff; }
        else { upperBound = midPoint; upDiff = midDiff; }
      }
      transitionIndex = lowerBound;
      const float P0 = hybrid_normalization(nAngles, angleCosines, values, gaussianValues, transitionIndex);
      for (int i = 0; i < transitionIndex; ++i) newValues[i] = P0 * gaussianValues[i];
    }
  }
  free(angleCosines); free(gaussianValues);
  return transitionIndex;
}
