// mcb_reference.cu -- photon kernels in REFERENCE arithmetic (sm_100a).
//
// The per-photon history below follows the reference's photon loop operation by operation
// (same f32/f64 mix, true divides, no fused multiply-adds: this file is compiled with
// -fmad=false), so that with injected random numbers it reproduces the reference's cell
// indices, event sequence and table look-ups exactly (north-star criterion (a)).
//   INT = Integrators/monteCarloRadiativeTransfer.f95, OPT = src/opticalProperties.f95,
//   ILL = src/monteCarloIllumination.f95, NUM = src/numericUtilities.f95.
// Single-precision log/exp/cos/sin/acos are evaluated as the correctly rounded value,
// (float)f((double)x); sqrt and division are IEEE (-prec-sqrt/-prec-div defaults).
//
// It serves two entry points: the fixed-random-number trace harness (mcb_run_trace) and
// full Philox-driven batches when mcb_options.arithmetic == MCB_ARITH_REFERENCE.  The
// throughput path is mcb_fast.cu.
#include "mcb_device.cuh"

namespace mcbref {

// ---- Fortran intrinsics ---------------------------------------------------------------
__device__ __forceinline__ double sp64(double x) {            // spacing(real(8)); spacing(0) = tiny
  x = fabs(x);
  if (x == 0.0) return DBL_MIN;
  double s = __longlong_as_double(__double_as_longlong(x) + 1LL) - x;
  return s < DBL_MIN ? DBL_MIN : s;
}
#define TINY32 FLT_MIN
__device__ __forceinline__ float f_log(float x)  { return (float)log((double)x); }
__device__ __forceinline__ float f_exp(float x)  { return (float)exp((double)x); }
__device__ __forceinline__ float f_cos(float x)  { return (float)cos((double)x); }
__device__ __forceinline__ float f_sin(float x)  { return (float)sin((double)x); }
__device__ __forceinline__ float f_acos(float x) { return (float)acos((double)x); }
#define PI32 3.14159265358979312f                              // INT:31

// ---- random-number policies -------------------------------------------------------------
struct PhiloxReal {                 // getRandomReal RNG:277-301: f32( u32 / (2^32 - 1) ) in [0,1]
  Philox g;
  __device__ __forceinline__ float real() { return (float)((double)g.next_u32() / 4294967295.0); }
  __device__ __forceinline__ bool exhausted() const { return false; }
  __device__ __forceinline__ int ndrawn() const { return (int)g.ndrawn; }
};
struct InjectedReal {               // trace harness: numbers supplied by the caller
  const float *p; long long n, pos; bool ex; int nd;
  __device__ __forceinline__ float real() {
    nd++;
    if (pos >= n) { ex = true; return 0.25f; }
    return p[pos++];
  }
  __device__ __forceinline__ bool exhausted() const { return ex; }
  __device__ __forceinline__ int ndrawn() const { return nd; }
};

// ---- trace sink ---------------------------------------------------------------------------
struct TraceSink {
  mcb_event *ev; int cap; int n; int photon;
};
template <bool TRACE, class RNG>
__device__ __forceinline__ void emit(TraceSink &t, const RNG &rng, int kind, int ix, int iy, int iz,
                                     int component, int phaseIndex, int angleIndex, int order,
                                     float weight, float tau, double path,
                                     double x, double y, double z, const float *dir) {
  if (!TRACE) return;
  if (t.n < t.cap) {
    mcb_event e;
    e.photon = t.photon; e.kind = kind; e.ix = ix; e.iy = iy; e.iz = iz;
    e.component = component; e.phaseIndex = phaseIndex; e.angleIndex = angleIndex;
    e.order = order; e.nrn = rng.ndrawn(); e.weight = weight; e.tau = tau; e.path = path;
    e.x = x; e.y = y; e.z = z; e.dir[0] = dir[0]; e.dir[1] = dir[1]; e.dir[2] = dir[2]; e.pad = 0;
    t.ev[t.n] = e;
  }
  t.n++;
}

// ---- searches (NUM:206-348); tables are 1-based in the reference -----------------------
__device__ int findIndexDouble(double value, const double *table, int n, int firstGuess) {
  int lowerBound, upperBound, midPoint, increment;
  if (firstGuess > 0) {
    lowerBound = firstGuess; increment = 1;
    for (;;) {
      upperBound = min(lowerBound + increment, n);
      if (lowerBound == n || (table[lowerBound - 1] <= value && table[upperBound - 1] > value)) break;
      if (table[lowerBound - 1] > value) {
        upperBound = lowerBound;
        lowerBound = max(upperBound - increment, 1);
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
    if (value >= table[midPoint - 1]) lowerBound = midPoint; else upperBound = midPoint;
  }
  return lowerBound;
}
__device__ int findCDFIndex(float value, const double *table, int n, long long stride) {  // NUM:317-348
  int lowerBound = 0, upperBound = n, midPoint;
  const double v = (double)value;
  for (;;) {
    if (lowerBound == n || upperBound <= lowerBound + 1) break;
    midPoint = (lowerBound + upperBound) / 2;
    if (v > table[(long long)(midPoint - 1) * stride]) lowerBound = midPoint; else upperBound = midPoint;
  }
  return upperBound;
}

#define CELL(P, ix, iy, iz) ((size_t)((ix) - 1) + (size_t)(P).nx * ((size_t)((iy) - 1) + (size_t)(P).ny * (size_t)((iz) - 1)))
#define CELLC(P, ix, iy, iz, c) (CELL(P, ix, iy, iz) + (size_t)(P).nx * (P).ny * (P).nz * (size_t)((c) - 1))

// ---- accumulateExtinctionAlongPath OPT:1656-1815 ---------------------------------------
__device__ float march(const DevDomain &P, const float *dir,
                       double &x, double &y, double &z, int &ix, int &iy, int &iz,
                       bool hasTarget, float extToAccumulate, double &totalPath,
                       unsigned long long &crossings) {
  const int nX = P.nx, nY = P.ny, nZ = P.nz;
  float extAccumulated = 0.0f;
  totalPath = 0.0;
  const int s0 = dir[0] >= 0.0f ? 1 : 0, s1 = dir[1] >= 0.0f ? 1 : 0, s2 = dir[2] >= 0.0f ? 1 : 0;  // OPT:1690
  const int i0 = dir[0] >= 0.0f ? 1 : -1, i1 = dir[1] >= 0.0f ? 1 : -1, i2 = dir[2] >= 0.0f ? 1 : -1; // OPT:1692
  const double z0 = P.zE[0], zMax = P.zE[nZ];
  const bool ok0 = fabsf(dir[0]) >= 2.0f * TINY32, ok1 = fabsf(dir[1]) >= 2.0f * TINY32,
             ok2 = fabsf(dir[2]) >= 2.0f * TINY32;
  const double d0 = (double)dir[0], d1 = (double)dir[1], d2 = (double)dir[2];
  for (;;) {
    const double st0 = ok0 ? (P.xE[ix + s0 - 1] - x) / d0 : DBL_MAX;          // OPT:1705-1712
    const double st1 = ok1 ? (P.yE[iy + s1 - 1] - y) / d1 : DBL_MAX;
    const double st2 = ok2 ? (P.zE[iz + s2 - 1] - z) / d2 : DBL_MAX;
    double thisStep = st0;
    if (st1 < thisStep) thisStep = st1;
    if (st2 < thisStep) thisStep = st2;
    if (thisStep <= 0.0) { extAccumulated = -2.0f; break; }                   // OPT:1719-1722
    const double thisCellExt = P.totalExt[CELL(P, ix, iy, iz)];               // OPT:1727
    crossings++;
    if (hasTarget) {                                                          // OPT:1729-1739
      if ((double)extAccumulated + thisStep * thisCellExt > (double)extToAccumulate) {
        thisStep = (double)(extToAccumulate - extAccumulated) / thisCellExt;
        x = x + thisStep * d0;
        y = y + thisStep * d1;
        z = z + thisStep * d2;
        totalPath = totalPath + thisStep;
        extAccumulated = extToAccumulate;
        break;
      }
    }
    extAccumulated = (float)((double)extAccumulated + thisStep * thisCellExt); // OPT:1743
    totalPath = totalPath + thisStep;
    if (st0 <= thisStep) {                                                    // OPT:1752-1759
      x = P.xE[ix + s0 - 1]; ix = ix + i0;
    } else {
      x = x + thisStep * d0;
      if (fabs(P.xE[ix + s0 - 1] - x) <= 2.0 * sp64(x)) ix = ix + i0;
    }
    if (st1 <= thisStep) {                                                    // OPT:1761-1768
      y = P.yE[iy + s1 - 1]; iy = iy + i1;
    } else {
      y = y + thisStep * d1;
      if (fabs(P.yE[iy + s1 - 1] - y) <= 2.0 * sp64(y)) iy = iy + i1;
    }
    if (st2 <= thisStep) {                                                    // OPT:1770-1777
      z = P.zE[iz + s2 - 1]; iz = iz + i2;
    } else {
      z = z + thisStep * d2;
      if (fabs(P.zE[iz + s2 - 1] - z) <= 2.0 * sp64(z)) iz = iz + i2;
    }
    if (ix <= 0) {                                                            // OPT:1782-1788
      ix = nX; x = P.xE[ix] + (double)(i0 * 2) * sp64(x);
    } else if (ix >= nX + 1) {
      ix = 1; x = P.xE[0] + (double)(i0 * 2) * sp64(x);
    }
    if (iy <= 0) {                                                            // OPT:1790-1796, cellIncrement(1) sic
      iy = nY; y = P.yE[iy] + (double)(i0 * 2) * sp64(y);
    } else if (iy >= nY + 1) {
      iy = 1; y = P.yE[0] + (double)(i0 * 2) * sp64(y);
    }
    if (iz > nZ) { z = zMax + 2.0 * sp64(zMax); break; }                      // OPT:1801-1804
    if (iz < 1) { z = z0; break; }                                            // OPT:1809-1812
  }
  return extAccumulated;
}

__device__ __forceinline__ void makeDirectionCosines(float mu, float phi, float *out) {  // INT:1876-1894
  const float sinTheta = sqrtf(1.0f - mu * mu);
  const float cosPhi = f_cos(phi), sinPhi = f_sin(phi);
  out[0] = sinTheta * cosPhi; out[1] = sinTheta * sinPhi; out[2] = mu;
}

__device__ void findXYIndicies(const DevDomain &P, double xPos, double yPos, int &xIndex, int &yIndex) { // INT:1551-1578
  const int nxp1 = P.nx + 1, nyp1 = P.ny + 1;
  if (P.xyRegular) {
    int xi = min((int)((xPos - P.x0) / P.deltaX) + 1, nxp1 - 1);
    int yi = min((int)((yPos - P.y0) / P.deltaY) + 1, nyp1 - 1);
    if (fabs(P.xE[xi] - xPos) < sp64(xPos)) xi = xi + 1;
    if (fabs(P.yE[yi] - yPos) < sp64(yPos)) yi = yi + 1;
    if (xi == nxp1) xi = 1;
    if (yi == nyp1) yi = 1;
    xIndex = xi; yIndex = yi;
  } else {
    int xi = findIndexDouble(xPos, P.xE, nxp1, xIndex);
    int yi = findIndexDouble(yPos, P.yE, nyp1, yIndex);
    if (fabs(P.xE[xi - 1] - xPos) < sp64(xPos)) xi = xi + 1;
    if (fabs(P.yE[yi - 1] - yPos) < sp64(yPos)) yi = yi + 1;
    if (xi >= nxp1) xi = 1;
    if (yi >= nyp1) yi = 1;
    xIndex = xi; yIndex = yi;
  }
}
__device__ void findZIndex(const DevDomain &P, double zPos, int &zIndex) {     // INT:1580-1592
  if (P.zRegular) {
    int zi = min((int)((zPos - P.z0) / P.deltaZ) + 1, P.nz);
    if (fabs(P.zE[zi] - zPos) < sp64(zPos)) zi = zi + 1;
    zIndex = zi;
  } else {
    zIndex = findIndexDouble(zPos, P.zE, P.nz + 1, zIndex);
  }
}

__device__ __forceinline__ float computeScatteringAngle(float randomDeviate, const float *table,
                                                        int numIntervals, int &k) {  // INT:1594-1621
  const int angleIndex = (int)(randomDeviate * (float)numIntervals) + 1;
  k = angleIndex;
  if (angleIndex < numIntervals) {
    const float leftOver = randomDeviate - (float)(angleIndex - 1) / (float)numIntervals;
    return (1.0f - leftOver) * table[angleIndex - 1] + leftOver * table[angleIndex];
  }
  return table[numIntervals - 1];
}

template <class RNG>
__device__ void next_direct(RNG &rng, float scatteringCosine, float *S) {      // INT:1921-1948
  float D = 2.0f, AX = 0.0f, AY = 0.0f, B;
  while (D > 1.0f) {
    AX = 1.0f - 2.0f * rng.real();
    AY = 1.0f - 2.0f * rng.real();
    D = AX * AX + AY * AY;
    if (rng.exhausted()) break;
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

__device__ __forceinline__ float lookUpPhaseFuncVal(const float *table, int nAngleSteps, float scatteringAngle) { // INT:1834-1873
  const float deltaTheta = PI32 / (float)(nAngleSteps - 1);
  const int angleIndex = (int)(scatteringAngle / deltaTheta) + 1;
  if (angleIndex < nAngleSteps) {
    const float weight = 1.0f - (scatteringAngle - (float)(angleIndex - 1) * deltaTheta) / deltaTheta;
    return weight * table[angleIndex - 1] + (1.0f - weight) * table[angleIndex];
  }
  return table[nAngleSteps - 1];
}

struct Counts { unsigned long long c[CNT_N]; };

// computeIntensityContribution INT:1623-1832 + the tally update at the call sites
template <class RNG, bool TRACE>
__device__ void intensityContribution(const DevDomain &P, float photonWeight,
                                      double xPos, double yPos, double zPos,
                                      int xIndex, int yIndex, int zIndex,
                                      const float *directionCosines, int component, int tallyComponent,
                                      RNG &rng, int scatteringOrder, TraceSink &ts, Counts &cnt) {
  const int zIndexMax = P.nz + 1;                                              // INT:1677
  const size_t cols = (size_t)P.nx * P.ny;
  for (int i = 0; i < P.nDir; ++i) {
    const float *vd = &P.viewDir[3 * i];
    float normalizedPhaseFunc;
    if (component == 0) {                                                      // INT:1688-1694
      normalizedPhaseFunc = 1.0f / PI32;
    } else if (component < 0) {                                                // INT:1695-1696
      normalizedPhaseFunc = 1.0f / (4.0f * PI32 * fabsf(vd[2]));
    } else {                                                                   // INT:1697-1727
      float projection = 0.0f;
      projection = projection + directionCosines[0] * vd[0];
      projection = projection + directionCosines[1] * vd[1];
      projection = projection + directionCosines[2] * vd[2];
      if (fabsf(projection) > 1.0f) projection = copysignf(1.0f, projection);
      const float scatteringAngle = f_acos(projection);
      const int phaseFunctionIndex = P.phaseIdx[CELLC(P, xIndex, yIndex, zIndex, component)];
      const int c = component - 1;
      const float *tab = (P.opt.useHybridPhaseFunsForIntenCalcs &&
                          scatteringOrder <= P.opt.numOrdersOrigPhaseFunIntenCalcs) ? P.fwdOrig[c] : P.fwd[c];
      const float phaseFunctionVal = lookUpPhaseFuncVal(tab + (size_t)(phaseFunctionIndex - 1) * P.fwdS[c],
                                                        P.fwdS[c], scatteringAngle);
      normalizedPhaseFunc = phaseFunctionVal / (4.0f * PI32 * fabsf(vd[2]));
    }
    double xTemp = xPos, yTemp = yPos, zTemp = zPos, pathDummy;
    int xF = xIndex, yF = yIndex, zF = zIndex;
    float tauToBoundary = 0.0f, contribution;
    cnt.c[CNT_LE_RAYS]++;
    if (!P.opt.useRussianRouletteForIntensity) {                               // INT:1729-1752
      tauToBoundary = march(P, vd, xTemp, yTemp, zTemp, xF, yF, zF, false, 0.0f, pathDummy, cnt.c[CNT_LE_CROSSINGS]);
      if (tauToBoundary >= 0.0f) contribution = photonWeight * normalizedPhaseFunc * f_exp(-tauToBoundary);
      else contribution = 0.0f;
    } else {                                                                   // INT:1753-1813
      const float u = rng.real();
      const float tauFree = -f_log(fmaxf(TINY32, u));
      if (PI32 * normalizedPhaseFunc <= P.opt.zetaMin) {                       // Iwabuchi (2006) Eq 13
        tauToBoundary = march(P, vd, xTemp, yTemp, zTemp, xF, yF, zF, true, tauFree, pathDummy, cnt.c[CNT_LE_CROSSINGS]);
        const float test = rng.real();
        if (test <= PI32 * normalizedPhaseFunc / P.opt.zetaMin && zF >= zIndexMax)
          contribution = photonWeight * P.opt.zetaMin / PI32;
        else contribution = 0.0f;
      } else {                                                                 // Eq 14
        const float pn = PI32 * normalizedPhaseFunc;
        const float tauMax = -f_log(P.opt.zetaMin / fmaxf(TINY32, pn));
        tauToBoundary = march(P, vd, xTemp, yTemp, zTemp, xF, yF, zF, true, tauMax, pathDummy, cnt.c[CNT_LE_CROSSINGS]);
        if (zF >= zIndexMax && tauToBoundary >= 0.0f) {
          contribution = photonWeight * normalizedPhaseFunc * f_exp(-tauToBoundary);
        } else if (tauToBoundary >= 0.0f && zF < 1) {
          contribution = 0.0f;   // left through the surface; the reference's onward trace is out of bounds (INT:1793)
        } else if (tauToBoundary >= 0.0f) {
          tauToBoundary = march(P, vd, xTemp, yTemp, zTemp, xF, yF, zF, true, tauFree, pathDummy, cnt.c[CNT_LE_CROSSINGS]);
          if (zF >= zIndexMax) contribution = photonWeight * P.opt.zetaMin / PI32;
          else contribution = 0.0f;
        } else {
          contribution = 0.0f;
        }
      }
    }
    if (P.opt.limitIntensityContributions) {                                   // INT:1815-1826
      if (contribution > P.opt.maxIntensityContribution) {
        const int cslot = component < 0 ? 0 : component;
        atomicAdd(&P.tally[P.offExcess + i + (long long)P.nDir * cslot],
                  (double)(contribution - P.opt.maxIntensityContribution));
        contribution = P.opt.maxIntensityContribution;
      }
    }
    // INT:535-540, 696-701, 785-790: added where the ray left the domain
    const size_t col = (size_t)(xF - 1) + (size_t)P.nx * (size_t)(yF - 1);
    if (contribution != 0.0f) {
      atomicAdd(&P.tally[P.offInt + col + cols * i], (double)contribution);
      atomicAdd(&P.tally[P.offIntByComp + col + cols * ((size_t)i + (size_t)P.nDir * tallyComponent)], (double)contribution);
    }
    emit<TRACE>(ts, rng, MCB_EV_LOCAL_ESTIMATE, xF, yF, zF, i + 1, 0, 0, scatteringOrder,
                contribution, tauToBoundary, 0.0, xTemp, yTemp, zTemp, vd);
  }
}

// One photon: source sampling (ILL:62-101 / ILL:431-522) + computeRT's loop body (INT:463-823)
// makePeriodic INT:1898-1917: the result is DEFAULT REAL (quirk q15), the bounds are real(8)
__device__ float makePeriodic(double a, double aMin, double aMax) {
  float m = (float)a;
  for (;;) {
    if ((double)m <= aMax && (double)m > aMin) break;
    if ((double)m > aMax) m = (float)((double)m - (aMax - aMin));
    else if ((double)m == aMin) m = (float)aMax;
    else m = (float)((double)m + (aMax - aMin));
  }
  return m;
}

template <class RNG, bool TRACE>
__device__ void photon_history(const DevDomain &P, RNG &rng, TraceSink &ts, Counts &cnt) {
  const int numX = P.nx, numY = P.ny, numZ = P.nz, numComps = P.nc;
  const size_t cols = (size_t)numX * numY;
  double xPos, yPos, zPos;
  float mu, phi;
  // ---- new_PhotonStream: this photon's entry ----
  if (P.source == 0) {                                                         // ILL:88-96
    xPos = (double)rng.real();
    yPos = (double)rng.real();
    zPos = (double)(1.0f - FLT_EPSILON);                                       // 1. - spacing(1.)
    mu = P.solarMu; phi = P.solarPhi;
  } else {                                                                     // ILL:481-515
    const float pi32 = f_acos(-1.0f);
    float RN = rng.real();
    if ((double)RN > P.fracAtmsPower) {
      xPos = (double)rng.real();
      yPos = (double)rng.real();
      for (;;) {
        mu = sqrtf(rng.real());
        if (fabsf(mu) > 2.0f * TINY32) break;
        if (rng.exhausted()) break;
      }
      phi = rng.real() * 2.0f * pi32;
      zPos = 0.0;
    } else {
      RN = rng.real();
      const double *levelBase = P.voxelCDF + (size_t)(numX - 1) + (size_t)numX * (size_t)(numY - 1);
      const int ik = findCDFIndex(RN, levelBase, numZ, (long long)numX * numY);
      const double *colBase = P.voxelCDF + (size_t)(numX - 1) + (size_t)numX * numY * (size_t)(ik - 1);
      const int ij = findCDFIndex(RN, colBase, numY, numX);
      const double *voxBase = P.voxelCDF + (size_t)numX * ((size_t)(ij - 1) + (size_t)numY * (size_t)(ik - 1));
      const int ii = findCDFIndex(RN, voxBase, numX, 1);
      zPos = ((double)(ik - 1) * 1.0 / (double)numZ) + (double)(rng.real() / (float)numZ);
      if (ik == 1 && zPos == 0.0) zPos = 0.0 + DBL_EPSILON;
      if (ik == numZ && zPos > 1.0 - 2.0 * DBL_EPSILON) zPos = zPos - (2.0 * DBL_EPSILON);
      xPos = ((double)(ii - 1) * 1.0 / (double)numX) + (double)(rng.real() * (1.0f / (float)numX));
      yPos = ((double)(ij - 1) * 1.0 / (double)numY) + (double)(rng.real() * (1.0f / (float)numY));
      for (;;) {
        mu = 1.0f - (2.0f * rng.real());
        if (fabsf(mu) > 2.0f * TINY32) break;
        if (rng.exhausted()) break;
      }
      phi = rng.real() * 2.0f * pi32;
    }
  }

  // ---- computeRT, INT:466-542 ----
  int scatteringOrder = 0;
  float directionCosines[3];
  makeDirectionCosines(mu, phi, directionCosines);
  float photonWeight = 1.0f;
  int xIndex = 1, yIndex = 1, zIndex = 1;
  xPos = P.x0 + xPos * (P.xMax - P.x0);
  yPos = P.y0 + yPos * (P.yMax - P.y0);
  findXYIndicies(P, xPos, yPos, xIndex, yIndex);
  if (P.zRegular) {
    zPos = P.z0 + zPos * (P.zMax - P.z0);
    findZIndex(P, zPos, zIndex);
  } else {                                                                     // INT:491-493
    const double remainder = (zPos - P.z0) * numZ - floor((zPos - P.z0) * numZ);
    zIndex = min((int)floor((zPos - P.z0) * numZ) + 1, numZ);
    zPos = P.zE[zIndex - 1] + remainder * (P.zE[zIndex] - P.zE[zIndex - 1]);
  }
  cnt.c[CNT_PHOTONS]++;
  emit<TRACE>(ts, rng, MCB_EV_BIRTH, xIndex, yIndex, zIndex, 0, 0, 0, 0, photonWeight, 0.0f, 0.0,
              xPos, yPos, zPos, directionCosines);

  if (P.opt.LW_flag > 0.0f) {                                                  // INT:504-542
    if (zPos > 0.0) {
      atomicAdd(&P.tally[P.offFluxAbs + (size_t)(xIndex - 1) + (size_t)numX * (size_t)(yIndex - 1)], -1.0);
      atomicAdd(&P.tally[P.offVolAbs + CELL(P, xIndex, yIndex, zIndex)], -1.0);
    }
    if (P.nDir > 0)
      intensityContribution<RNG, TRACE>(P, photonWeight, xPos, yPos, zPos, xIndex, yIndex, zIndex,
                                        directionCosines, zPos == 0.0 ? 0 : -1, 0, rng, scatteringOrder, ts, cnt);
  }

  for (;;) {                                                                   // scatteringLoop INT:548
    if (rng.exhausted()) {
      emit<TRACE>(ts, rng, MCB_EV_RN_EXHAUSTED, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                  photonWeight, 0.0f, 0.0, xPos, yPos, zPos, directionCosines);
      return;
    }
    const float u = rng.real();
    const float tauToTravel = -f_log(fmaxf(TINY32, u));                        // INT:554
    double path = 0.0;
    const bool useMaxCrossSection = !P.opt.useRayTracing;                      // INT:445
    if (!useMaxCrossSection) {
      const float tauAccumulated = march(P, directionCosines, xPos, yPos, zPos, xIndex, yIndex, zIndex,
                                         true, tauToTravel, path, cnt.c[CNT_CROSSINGS]);   // INT:559-561
      if (tauAccumulated < 0.0f) {                                             // INT:562-563
        cnt.c[CNT_BAD]++;
        emit<TRACE>(ts, rng, MCB_EV_BAD, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                    photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
        return;
      }
    } else {                                       // INT:564-571; the cell indices are NOT refreshed here (sic)
      xPos = (double)makePeriodic(xPos + (double)(directionCosines[0] * tauToTravel / P.maxExtinction), P.x0, P.xMax);
      yPos = (double)makePeriodic(yPos + (double)(directionCosines[1] * tauToTravel / P.maxExtinction), P.y0, P.yMax);
      zPos = zPos + (double)(directionCosines[2] * tauToTravel / P.maxExtinction);
    }
    if (useMaxCrossSection && (zPos >= P.zMax || zPos <= P.z0 + sp64(P.z0))) {  // INT:578-585, 624-631: trace back to the boundary
      const double zB = zPos >= P.zMax ? P.zMax : P.z0;
      xPos = (double)makePeriodic(xPos - (double)directionCosines[0] * fabs((zPos - zB) / (double)directionCosines[2]), P.x0, P.xMax);
      yPos = (double)makePeriodic(yPos - (double)directionCosines[1] * fabs((zPos - zB) / (double)directionCosines[2]), P.y0, P.yMax);
      findXYIndicies(P, xPos, yPos, xIndex, yIndex);
    }
    const size_t colIdx = (size_t)(xIndex - 1) + (size_t)numX * (size_t)(yIndex - 1);
    if (zPos >= P.zMax) {                                                      // INT:573-617
      atomicAdd(&P.tally[P.offFluxUp + colIdx], (double)photonWeight);
      cnt.c[CNT_TOP]++;
      emit<TRACE>(ts, rng, MCB_EV_EXIT_TOP, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                  photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
      return;
    } else if (zPos <= P.z0 + sp64(P.z0)) {                                    // INT:619-702
      zIndex = 1;
      zPos = P.z0 + sp64(P.z0);
      atomicAdd(&P.tally[P.offFluxDown + colIdx], (double)photonWeight);
      cnt.c[CNT_SURFACE]++;
      scatteringOrder = scatteringOrder + 1;
      for (;;) {                                                               // INT:655-662
        mu = sqrtf(rng.real());
        if (fabsf(mu) > 2.0f * TINY32) break;
        if (rng.exhausted()) break;
      }
      phi = 2.0f * PI32 * rng.real();                                          // INT:663
      photonWeight = (float)((double)photonWeight * P.albedo);                 // INT:673
      if (photonWeight <= TINY32) {                                            // INT:675
        emit<TRACE>(ts, rng, MCB_EV_KILLED_SURFACE, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                    photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
        return;
      }
      makeDirectionCosines(mu, phi, directionCosines);
      emit<TRACE>(ts, rng, MCB_EV_SURFACE, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                  photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
      if (P.nDir > 0)                                                          // INT:680-702
        intensityContribution<RNG, TRACE>(P, photonWeight, xPos, yPos, zPos, xIndex, yIndex, zIndex,
                                          directionCosines, 0, 0, rng, scatteringOrder, ts, cnt);
    } else {                                                                   // INT:703-821
      if (useMaxCrossSection) {                                                // INT:709-710: physical or mathematical event
        const float rnPhys = rng.real();
        if (!((double)rnPhys < P.totalExt[CELL(P, xIndex, yIndex, zIndex)] / (double)P.maxExtinction)) {
          emit<TRACE>(ts, rng, MCB_EV_NULL_COLLISION, xIndex, yIndex, zIndex, 0, 0, 0, scatteringOrder,
                      photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
          continue;
        }
      }
      scatteringOrder = scatteringOrder + 1;
      cnt.c[CNT_SCATTERS]++;
      if (P.totalExt[CELL(P, xIndex, yIndex, zIndex)] <= 0.0) {                // INT:728-754
        if (xPos - P.xE[xIndex - 1] <= 0.0 && directionCosines[0] > 0.0f) {
          xPos = xPos - sp64(xPos);
          xIndex = xIndex - 1;
          if (xIndex <= 0) {
            xIndex = numX;
            xPos = P.xE[xIndex - 1];
            xPos = xPos - 2.0 * sp64(xPos);
          }
        }
        if (yPos - P.yE[yIndex - 1] <= 0.0 && directionCosines[1] > 0.0f) {
          yPos = yPos - sp64(yPos);
          yIndex = yIndex - 1;
          if (yIndex <= 0) {
            yIndex = numY;
            yPos = P.xE[min(yIndex, numX + 1) - 1];                            // INT:743 sic: xPosition(yIndex)
            yPos = yPos - 2.0 * sp64(yPos);
          }
        }
        if (zPos - P.zE[zIndex - 1] <= 0.0 && directionCosines[2] > 0.0f) {
          zPos = zPos - sp64(zPos);
          zIndex = zIndex - 1;
        }
      }
      // findIndex(RN, (/ 0, cumExt(ix,iy,iz,:) /)), INT:759-760 / NUM:262-315 without a first guess
      const float rnComp = rng.real();
      int component;
      {
        int lowerBound = 0, upperBound = numComps + 1;
        for (;;) {
          if (lowerBound == numComps + 1 || upperBound <= lowerBound + 1) break;
          const int midPoint = (lowerBound + upperBound) / 2;
          const double t = midPoint == 1 ? 0.0 : P.cumExt[CELLC(P, xIndex, yIndex, zIndex, midPoint - 1)];
          if ((double)rnComp >= t) lowerBound = midPoint; else upperBound = midPoint;
        }
        component = lowerBound;
      }
      const float ssa = (float)P.ssa[CELLC(P, xIndex, yIndex, zIndex, component)];   // INT:764
      if ((double)ssa < 1.0) {                                                 // INT:765-771
        const double absorbed = (double)photonWeight * (1.0 - (double)ssa);
        atomicAdd(&P.tally[P.offFluxAbs + (size_t)(xIndex - 1) + (size_t)numX * (size_t)(yIndex - 1)], absorbed);
        atomicAdd(&P.tally[P.offVolAbs + CELL(P, xIndex, yIndex, zIndex)], absorbed);
        photonWeight = photonWeight * ssa;
      }
      if (P.nDir > 0)                                                          // INT:776-800
        intensityContribution<RNG, TRACE>(P, photonWeight, xPos, yPos, zPos, xIndex, yIndex, zIndex,
                                          directionCosines, component, component, rng, scatteringOrder, ts, cnt);
      if (P.opt.useRussianRoulette && photonWeight < P.opt.russianRouletteW / 2.0f) {   // INT:805-811
        if (rng.real() >= photonWeight / P.opt.russianRouletteW) photonWeight = 0.0f;
        else photonWeight = P.opt.russianRouletteW;
      }
      const int phaseFunctionIndex = P.phaseIdx[CELLC(P, xIndex, yIndex, zIndex, component)];  // INT:816
      if (photonWeight <= TINY32) {                                            // INT:812
        cnt.c[CNT_RR_KILLS]++;
        emit<TRACE>(ts, rng, MCB_EV_KILLED_ROULETTE, xIndex, yIndex, zIndex, component, phaseFunctionIndex, 0,
                    scatteringOrder, photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
        return;
      }
      const int c = component - 1;
      int k = 0;
      const float scatteringAngle = computeScatteringAngle(rng.real(),
          P.inv[c] + (size_t)(phaseFunctionIndex - 1) * P.invS[c], P.invS[c], k);          // INT:817-818
      next_direct(rng, f_cos(scatteringAngle), directionCosines);              // INT:819
      emit<TRACE>(ts, rng, MCB_EV_SCATTER, xIndex, yIndex, zIndex, component, phaseFunctionIndex, k,
                  scatteringOrder, photonWeight, tauToTravel, path, xPos, yPos, zPos, directionCosines);
    }
  }
  (void)cols;
}

__device__ __forceinline__ void flush_counts(const DevDomain &P, Counts &cnt) {
#pragma unroll
  for (int i = 0; i < CNT_N; ++i) {
    unsigned long long v = cnt.c[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&P.counters[i], v);
  }
}

// Full batches in reference arithmetic (Philox stream per photon id).
__global__ void __launch_bounds__(128)
batch_kernel(const __grid_constant__ DevDomain P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId) {
  Counts cnt;
#pragma unroll
  for (int i = 0; i < CNT_N; ++i) cnt.c[i] = 0;
  TraceSink ts{nullptr, 0, 0, 0};
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nPhotons;
       p += (long long)gridDim.x * blockDim.x) {
    PhiloxReal rng;
    rng.g.init(seed, firstPhotonId + (uint64_t)p);
    photon_history<PhiloxReal, false>(P, rng, ts, cnt);
  }
  flush_counts(P, cnt);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.tally[P.offPhotons], (double)nPhotons);
}

// Fixed-random-number trace harness: one thread per photon, events into its own segment.
__global__ void __launch_bounds__(64)
trace_kernel(const __grid_constant__ DevDomain P, long long nPhotons, const float *rn, long long rnStride,
             mcb_event *events, int maxEventsPerPhoton, int *eventCount) {
  Counts cnt;
#pragma unroll
  for (int i = 0; i < CNT_N; ++i) cnt.c[i] = 0;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nPhotons) {
    InjectedReal rng{rn + p * rnStride, rnStride, 0, false, 0};
    TraceSink ts{events + p * maxEventsPerPhoton, maxEventsPerPhoton, 0, (int)p};
    photon_history<InjectedReal, true>(P, rng, ts, cnt);
    eventCount[p] = ts.n;
  }
  flush_counts(P, cnt);
}

}  // namespace mcbref

// ---- launchers used by mcb_api.cu ---------------------------------------------------------
void mcb_launch_reference_batch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
                                int numSMs, cudaStream_t stream) {
  if (nPhotons <= 0) return;
  const int threads = 128;
  long long want = (nPhotons + threads - 1) / threads;
  long long cap = (long long)numSMs * 8;
  int blocks = (int)(want < cap ? want : cap);
  mcbref::batch_kernel<<<blocks, threads, 0, stream>>>(P, nPhotons, seed, firstPhotonId);
}

void mcb_launch_trace(const DevDomain &P, long long nPhotons, const float *rn, long long rnStride,
                      mcb_event *events, int maxEventsPerPhoton, int *eventCount, cudaStream_t stream) {
  if (nPhotons <= 0) return;
  const int threads = 64;
  int blocks = (int)((nPhotons + threads - 1) / threads);
  mcbref::trace_kernel<<<blocks, threads, 0, stream>>>(P, nPhotons, rn, rnStride, events, maxEventsPerPhoton, eventCount);
}
