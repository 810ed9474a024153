// mcb_fast.cu -- the throughput photon kernel (sm_100a).
//
// Same physics and event sequence as the reference's computeRT (INT:393-841) but laid out
// for the GPU instead of mirroring the reference's arithmetic:
//   * one persistent kernel, one photon per lane, lanes refilled from a global photon
//     counter (warp-aggregated atomic) so no lane idles while photons remain;
//   * ray marching (accumulateExtinctionAlongPath, OPT:1656-1815) as a parametric
//     Amanatides-Woo DDA in single precision: per cell one dependent gather of the packed
//     f32 extinction, one compare, one face update from shared-memory-staged edges; no
//     divides inside the cell loop; periodic x/y handled by shifting the leg origin;
//   * warp-level regrouping: lanes march together and park when they reach an event
//     (scatter / surface); the event code runs once enough lanes are parked, so the
//     expensive scatter path executes with many active lanes instead of one or two;
//   * Philox4x32-10 per-photon streams (mcb_device.cuh) instead of a sequential MT19937;
//   * tallies as f64 reductions into the packed tally buffer (RED.ADD.F64).
// Results agree with the reference arithmetic statistically (north-star criterion (b));
// bit-level trace parity is the job of mcb_reference.cu.
#include "mcb_device.cuh"

namespace mcbfast {

#define FULL 0xffffffffu
#define PI32 3.14159265358979312f
#define TINY32 FLT_MIN

enum { ST_DEAD = 0, ST_MARCH = 1, ST_SCATTER = 2, ST_SURFACE = 3, ST_DONE = 4 };

struct Rng {
  Philox g;
  __device__ __forceinline__ float real() {          // f32 in [0,1] built from 32 bits (RNG:286-300)
    return __uint2float_rn(g.next_u32()) * 2.3283064365386963e-10f;
  }
};

struct Ray {
  float ox, oy, oz;        // leg origin (x, y shifted by whole domain periods when wrapping)
  float dx, dy, dz;        // direction cosines
  float rx, ry, rz;        // reciprocal direction cosines (FLT_MAX-guarded)
  float t;                 // distance along the leg
  float tx, ty, tz;        // distance along the leg at which the next x/y/z face is met
  int ix, iy, iz;          // 0-based cell
};

struct Grid {              // shared-memory staged edges (f32)
  const float *sx, *sy, *sz;
  int nx, ny, nz;
  float Lx, Ly;
};

__device__ __forceinline__ float safe_rcp(float d) {
  return fabsf(d) >= 2.0f * TINY32 ? 1.0f / d : FLT_MAX;      // OPT:1705-1712 zero-direction guard
}

__device__ __forceinline__ void ray_start(Ray &r, const Grid &G) {
  r.rx = safe_rcp(r.dx); r.ry = safe_rcp(r.dy); r.rz = safe_rcp(r.dz);
  r.t = 0.0f;
  r.tx = r.rx == FLT_MAX ? FLT_MAX : (G.sx[r.ix + (r.dx >= 0.0f ? 1 : 0)] - r.ox) * r.rx;
  r.ty = r.ry == FLT_MAX ? FLT_MAX : (G.sy[r.iy + (r.dy >= 0.0f ? 1 : 0)] - r.oy) * r.ry;
  r.tz = r.rz == FLT_MAX ? FLT_MAX : (G.sz[r.iz + (r.dz >= 0.0f ? 1 : 0)] - r.oz) * r.rz;
}

// Advance across the face(s) reached at distance tmin.  Returns 0 inside, 1 out the top, 2 out the bottom.
__device__ __forceinline__ int ray_advance(Ray &r, const Grid &G, float tmin) {
  r.t = tmin;
  if (r.tx <= tmin) {
    if (r.dx >= 0.0f) { if (++r.ix >= G.nx) { r.ix = 0; r.ox -= G.Lx; } r.tx = (G.sx[r.ix + 1] - r.ox) * r.rx; }
    else              { if (--r.ix < 0) { r.ix = G.nx - 1; r.ox += G.Lx; } r.tx = (G.sx[r.ix] - r.ox) * r.rx; }
  }
  if (r.ty <= tmin) {
    if (r.dy >= 0.0f) { if (++r.iy >= G.ny) { r.iy = 0; r.oy -= G.Ly; } r.ty = (G.sy[r.iy + 1] - r.oy) * r.ry; }
    else              { if (--r.iy < 0) { r.iy = G.ny - 1; r.oy += G.Ly; } r.ty = (G.sy[r.iy] - r.oy) * r.ry; }
  }
  if (r.tz <= tmin) {
    if (r.dz >= 0.0f) { if (++r.iz >= G.nz) return 1; r.tz = (G.sz[r.iz + 1] - r.oz) * r.rz; }
    else              { if (--r.iz < 0) return 2; r.tz = (G.sz[r.iz] - r.oz) * r.rz; }
  }
  return 0;
}

// Trace to the boundary or to an optical-depth target (the local-estimate rays, INT:1734-1739,
// 1764-1795).  Returns the accumulated optical depth; *where = 0 stopped at target, 1 top, 2 bottom.
__device__ float ray_trace(Ray &r, const Grid &G, const float *__restrict__ ext32, bool hasTarget, float target,
                           int &where, unsigned &crossings) {
  float ext = 0.0f;
  for (;;) {
    const float tmin = fminf(r.tx, fminf(r.ty, r.tz));
    const float s = fmaxf(tmin - r.t, 0.0f);
    const float sig = __ldg(&ext32[(size_t)r.ix + (size_t)G.nx * ((size_t)r.iy + (size_t)G.ny * (size_t)r.iz)]);
    crossings++;
    const float e2 = fmaf(s, sig, ext);
    if (hasTarget && e2 > target) {
      r.t += (target - ext) / sig;
      where = 0;
      return target;
    }
    ext = e2;
    const int out = ray_advance(r, G, tmin);
    if (out) { where = out; return ext; }
  }
}

__device__ __forceinline__ int find_cell(const float *e, int n, float x, int guess) {
  int i = min(max(guess, 0), n - 1);
  while (i > 0 && x < e[i]) --i;
  while (i < n - 1 && x >= e[i + 1]) ++i;
  return i;
}
__device__ __forceinline__ int find_cell_bisect(const float *e, int n, float x) {
  int lo = 0, hi = n;                      // e[lo] <= x < e[hi]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (x >= e[mid]) lo = mid; else hi = mid; }
  return lo;
}

__device__ __forceinline__ void dir_from(float mu, float phi, float &dx, float &dy, float &dz) {  // INT:1876-1894
  const float st = sqrtf(fmaxf(1.0f - mu * mu, 0.0f));
  float s, c;
  __sincosf(phi, &s, &c);
  dx = st * c; dy = st * s; dz = mu;
}

__device__ __forceinline__ int cdf_search(const double *__restrict__ table, int n, long long stride, float value) { // NUM:317-348
  int lo = 0, hi = n;
  const double v = (double)value;
  while (hi > lo + 1) {
    const int mid = (lo + hi) >> 1;
    if (v > __ldg(&table[(long long)(mid - 1) * stride])) lo = mid; else hi = mid;
  }
  return hi;                                // 1-based
}

struct Counts { unsigned c[CNT_N]; };

// computeIntensityContribution (INT:1623-1832) for one event, all view directions.
__device__ void local_estimate(const DevDomain &P, const Grid &G, Rng &rng, const Ray &r0, float px, float py, float pz,
                               float w, int component, int tallyComponent, int order, Counts &cnt) {
  const size_t cols = (size_t)P.nx * P.ny;
  const size_t cell = (size_t)r0.ix + (size_t)P.nx * ((size_t)r0.iy + (size_t)P.ny * (size_t)r0.iz);
  for (int i = 0; i < P.nDir; ++i) {
    const float vx = P.viewDir[3 * i], vy = P.viewDir[3 * i + 1], vz = P.viewDir[3 * i + 2];
    float npf;
    if (component == 0) {
      npf = 1.0f / PI32;                                                         // INT:1694
    } else if (component < 0) {
      npf = 1.0f / (4.0f * PI32 * fabsf(vz));                                    // INT:1696
    } else {
      float proj = r0.dx * vx + r0.dy * vy + r0.dz * vz;                         // INT:1704-1706
      proj = fminf(fmaxf(proj, -1.0f), 1.0f);
      const float ang = acosf(proj);
      const int c = component - 1;
      const int pidx = (int)__ldg(&P.idx16[cell + (size_t)P.nx * P.ny * P.nz * c]);
      const float *tab = ((P.opt.useHybridPhaseFunsForIntenCalcs && order <= P.opt.numOrdersOrigPhaseFunIntenCalcs)
                              ? P.fwdOrig[c] : P.fwd[c]) + (size_t)(pidx - 1) * P.fwdS[c];
      const int nS = P.fwdS[c];                                                  // INT:1855-1870
      const float dTheta = PI32 / (float)(nS - 1);
      const int ai = (int)(ang / dTheta) + 1;
      float val;
      if (ai < nS) {
        const float wt = 1.0f - (ang - (float)(ai - 1) * dTheta) / dTheta;
        val = wt * __ldg(&tab[ai - 1]) + (1.0f - wt) * __ldg(&tab[ai]);
      } else {
        val = __ldg(&tab[nS - 1]);
      }
      npf = val / (4.0f * PI32 * fabsf(vz));                                     // INT:1726
    }
    Ray r;
    r.ox = px; r.oy = py; r.oz = pz; r.dx = vx; r.dy = vy; r.dz = vz;
    r.ix = r0.ix; r.iy = r0.iy; r.iz = r0.iz;
    ray_start(r, G);
    int where = 0;
    float contribution;
    cnt.c[CNT_LE_RAYS]++;
    if (!P.opt.useRussianRouletteForIntensity) {                                 // INT:1729-1752
      const float tau = ray_trace(r, G, P.ext32, false, 0.0f, where, cnt.c[CNT_LE_CROSSINGS]);
      contribution = w * npf * __expf(-tau);
    } else {                                                                     // INT:1753-1813
      const float tauFree = -__logf(fmaxf(TINY32, rng.real()));
      if (PI32 * npf <= P.opt.zetaMin) {                                         // Iwabuchi (2006) Eq 13
        ray_trace(r, G, P.ext32, true, tauFree, where, cnt.c[CNT_LE_CROSSINGS]);
        const float test = rng.real();
        contribution = (test <= PI32 * npf / P.opt.zetaMin && where == 1) ? w * P.opt.zetaMin / PI32 : 0.0f;
      } else {                                                                   // Eq 14
        const float tauMax = -__logf(P.opt.zetaMin / fmaxf(TINY32, PI32 * npf));
        const float tau = ray_trace(r, G, P.ext32, true, tauMax, where, cnt.c[CNT_LE_CROSSINGS]);
        if (where == 1) {
          contribution = w * npf * __expf(-tau);
        } else if (where == 0) {
          // continue from where the first trace stopped (INT:1793-1795)
          r.ox = fmaf(r.t, r.dx, r.ox); r.oy = fmaf(r.t, r.dy, r.oy); r.oz = fmaf(r.t, r.dz, r.oz);
          ray_start(r, G);
          ray_trace(r, G, P.ext32, true, tauFree, where, cnt.c[CNT_LE_CROSSINGS]);
          contribution = where == 1 ? w * P.opt.zetaMin / PI32 : 0.0f;
        } else {
          contribution = 0.0f;       // left through the bottom before tauMax: zIndexF < zIndexMax
          // (the reference then traces on with tauFree from z0 and still finds zIndexF < zIndexMax)
        }
      }
    }
    if (P.opt.limitIntensityContributions && contribution > P.opt.maxIntensityContribution) {   // INT:1815-1826
      const int cslot = component < 0 ? 0 : component;
      atomicAdd(&P.tally[P.offExcess + i + (long long)P.nDir * cslot],
                (double)(contribution - P.opt.maxIntensityContribution));
      contribution = P.opt.maxIntensityContribution;
    }
    if (contribution != 0.0f) {
      const size_t col = (size_t)r.ix + (size_t)P.nx * (size_t)r.iy;
      atomicAdd(&P.tally[P.offInt + col + cols * i], (double)contribution);
      atomicAdd(&P.tally[P.offIntByComp + col + cols * ((size_t)i + (size_t)P.nDir * tallyComponent)], (double)contribution);
    }
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
batch_kernel(const __grid_constant__ DevDomain P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
             unsigned long long *workCounter, int parkThreshold) {
  extern __shared__ float smem[];
  float *sx = smem, *sy = sx + (P.nx + 1), *sz = sy + (P.ny + 1);
  for (int i = threadIdx.x; i <= P.nx; i += THREADS) sx[i] = (float)P.xE[i];
  for (int i = threadIdx.x; i <= P.ny; i += THREADS) sy[i] = (float)P.yE[i];
  for (int i = threadIdx.x; i <= P.nz; i += THREADS) sz[i] = (float)P.zE[i];
  __syncthreads();
  Grid G{sx, sy, sz, P.nx, P.ny, P.nz, sx[P.nx] - sx[0], sy[P.ny] - sy[0]};
  const int lane = threadIdx.x & 31;
  const size_t cols = (size_t)P.nx * P.ny, cells = cols * P.nz;
  const float *__restrict__ ext32 = P.ext32;

  Counts cnt;
#pragma unroll
  for (int i = 0; i < CNT_N; ++i) cnt.c[i] = 0;

  Rng rng;
  Ray r;
  float w = 0.0f, tau = 0.0f, ext = 0.0f;
  int order = 0;
  int state = ST_DEAD;
  bool more = true;                      // photons may remain in the global counter

  for (;;) {
    // ---- refill dead lanes: one atomic per warp (getNextPhoton, ILL:561-590) ----
    {
      const unsigned dead = __ballot_sync(FULL, state == ST_DEAD);
      if (dead && more) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(workCounter, (unsigned long long)__popc(dead));
        base = __shfl_sync(FULL, base, 0);
        if (state == ST_DEAD) {
          const unsigned long long p = base + (unsigned long long)__popc(dead & ((1u << lane) - 1u));
          if (p < (unsigned long long)nPhotons) {
            rng.g.init(seed, firstPhotonId + p);
            float x01, y01, z01, mu, phi;
            if (P.source == 0) {                                               // ILL:88-96
              x01 = rng.real(); y01 = rng.real(); z01 = 1.0f - FLT_EPSILON;
              mu = P.solarMu; phi = P.solarPhi;
            } else {                                                           // ILL:481-515
              const float RN = rng.real();
              if ((double)RN > P.fracAtmsPower) {
                x01 = rng.real(); y01 = rng.real();
                do { mu = sqrtf(rng.real()); } while (!(fabsf(mu) > 2.0f * TINY32));
                phi = rng.real() * 2.0f * PI32;
                z01 = 0.0f;
              } else {
                const float q = rng.real();
                const double *levelBase = P.voxelCDF + (size_t)(P.nx - 1) + (size_t)P.nx * (size_t)(P.ny - 1);
                const int ik = cdf_search(levelBase, P.nz, (long long)cols, q);
                const double *colBase = P.voxelCDF + (size_t)(P.nx - 1) + cols * (size_t)(ik - 1);
                const int ij = cdf_search(colBase, P.ny, P.nx, q);
                const double *voxBase = P.voxelCDF + (size_t)P.nx * ((size_t)(ij - 1) + (size_t)P.ny * (size_t)(ik - 1));
                const int ii = cdf_search(voxBase, P.nx, 1, q);
                // uniform inside the chosen cell, nudged off its faces (ILL:500-505)
                z01 = ((float)(ik - 1) + fminf(fmaxf(rng.real(), 1e-6f), 1.0f - 1e-6f)) / (float)P.nz;
                x01 = ((float)(ii - 1) + fminf(rng.real(), 1.0f - 1e-6f)) / (float)P.nx;
                y01 = ((float)(ij - 1) + fminf(rng.real(), 1.0f - 1e-6f)) / (float)P.ny;
                do { mu = 1.0f - 2.0f * rng.real(); } while (!(fabsf(mu) > 2.0f * TINY32));
                phi = rng.real() * 2.0f * PI32;
              }
            }
            dir_from(mu, phi, r.dx, r.dy, r.dz);
            w = 1.0f; order = 0;
            // INT:478-494: unit square -> domain; thermal z01 indexes the (possibly irregular) level directly
            r.ox = sx[0] + x01 * G.Lx; r.oy = sy[0] + y01 * G.Ly;
            if (P.xyRegular) {
              r.ix = find_cell(sx, P.nx, r.ox, (int)(x01 * (float)P.nx));
              r.iy = find_cell(sy, P.ny, r.oy, (int)(y01 * (float)P.ny));
            } else {
              r.ix = find_cell_bisect(sx, P.nx, r.ox);
              r.iy = find_cell_bisect(sy, P.ny, r.oy);
            }
            if (P.zRegular) {
              r.oz = sz[0] + z01 * (sz[P.nz] - sz[0]);
              r.iz = find_cell(sz, P.nz, r.oz, (int)(z01 * (float)P.nz));
            } else {                                                           // INT:491-493
              const float zs = z01 * (float)P.nz;
              r.iz = min((int)zs, P.nz - 1);
              r.oz = sz[r.iz] + (zs - (float)r.iz) * (sz[r.iz + 1] - sz[r.iz]);
            }
            cnt.c[CNT_PHOTONS]++;
            if (P.opt.LW_flag > 0.0f) {                                        // INT:504-542
              if (r.oz > 0.0f) {
                atomicAdd(&P.tally[P.offFluxAbs + (size_t)r.ix + (size_t)P.nx * (size_t)r.iy], -1.0);
                atomicAdd(&P.tally[P.offVolAbs + (size_t)r.ix + (size_t)P.nx * ((size_t)r.iy + (size_t)P.ny * (size_t)r.iz)], -1.0);
              }
              if (P.nDir > 0)
                local_estimate(P, G, rng, r, r.ox, r.oy, r.oz, w, r.oz == 0.0f ? 0 : -1, 0, order, cnt);
            }
            tau = -__logf(fmaxf(TINY32, rng.real()));                          // INT:554
            ext = 0.0f;
            ray_start(r, G);
            state = ST_MARCH;
          } else {
            state = ST_DONE;
          }
        }
        if (base + (unsigned long long)__popc(dead) >= (unsigned long long)nPhotons) more = false;
      } else if (dead && !more) {
        if (state == ST_DEAD) state = ST_DONE;
      }
    }
    if (__all_sync(FULL, state == ST_DONE)) break;

    // ---- march phase: cross cells until enough lanes are parked at an event ----
    for (;;) {
      const bool marching = state == ST_MARCH;
      const unsigned mk = __ballot_sync(FULL, marching);
      const unsigned live = __ballot_sync(FULL, state != ST_DONE);
      // stop when nobody marches, or when the parked (event/dead) lanes reach the threshold
      if (mk == 0u || __popc(live & ~mk) >= parkThreshold) break;
      if (marching) {
        const float tmin = fminf(r.tx, fminf(r.ty, r.tz));
        const float s = fmaxf(tmin - r.t, 0.0f);
        const float sig = __ldg(&ext32[(size_t)r.ix + (size_t)P.nx * ((size_t)r.iy + (size_t)P.ny * (size_t)r.iz)]);
        cnt.c[CNT_CROSSINGS]++;
        const float e2 = fmaf(s, sig, ext);
        if (e2 > tau) {                                                        // OPT:1729-1738
          r.t += (tau - ext) / sig;
          state = ST_SCATTER;
        } else {
          ext = e2;
          const int out = ray_advance(r, G, tmin);
          if (out == 1) {                                                      // INT:573-617
            atomicAdd(&P.tally[P.offFluxUp + (size_t)r.ix + (size_t)P.nx * (size_t)r.iy], (double)w);
            cnt.c[CNT_TOP]++;
            state = ST_DEAD;
          } else if (out == 2) {
            state = ST_SURFACE;
          }
        }
      }
    }

    // ---- event phase ----
    if (state == ST_SURFACE) {                                                 // INT:619-702
      const size_t col = (size_t)r.ix + (size_t)P.nx * (size_t)r.iy;
      atomicAdd(&P.tally[P.offFluxDown + col], (double)w);
      cnt.c[CNT_SURFACE]++;
      order++;
      float mu;
      do { mu = sqrtf(rng.real()); } while (!(fabsf(mu) > 2.0f * TINY32));
      const float phi = 2.0f * PI32 * rng.real();
      w = (float)((double)w * P.albedo);
      if (w <= TINY32) {
        state = ST_DEAD;
      } else {
        const float px = fmaf(r.t, r.dx, r.ox), py = fmaf(r.t, r.dy, r.oy);
        r.ox = px; r.oy = py; r.oz = sz[0]; r.iz = 0;
        dir_from(mu, phi, r.dx, r.dy, r.dz);
        if (P.nDir > 0) local_estimate(P, G, rng, r, r.ox, r.oy, r.oz, w, 0, 0, order, cnt);
        tau = -__logf(fmaxf(TINY32, rng.real()));
        ext = 0.0f;
        ray_start(r, G);
        state = ST_MARCH;
      }
    } else if (state == ST_SCATTER) {                                          // INT:703-821
      order++;
      cnt.c[CNT_SCATTERS]++;
      const size_t cell = (size_t)r.ix + (size_t)P.nx * ((size_t)r.iy + (size_t)P.ny * (size_t)r.iz);
      const float rnComp = rng.real();                                         // drawn even when nc == 1 (INT:759)
      int comp = 1;
      for (int c = 1; c < P.nc; ++c)                                           // findIndex on (0, cumExt(:)), NUM:262-315
        if (rnComp >= __ldg(&P.cum32[cell + cells * (size_t)(c - 1)])) comp = c + 1;
      const float ssa = __ldg(&P.ssa32[cell + cells * (size_t)(comp - 1)]);
      if (ssa < 1.0f) {                                                        // INT:765-771
        const double absorbed = (double)w * (1.0 - (double)ssa);
        atomicAdd(&P.tally[P.offFluxAbs + (size_t)r.ix + (size_t)P.nx * (size_t)r.iy], absorbed);
        atomicAdd(&P.tally[P.offVolAbs + cell], absorbed);
        w *= ssa;
      }
      const float px = fmaf(r.t, r.dx, r.ox), py = fmaf(r.t, r.dy, r.oy), pz = fmaf(r.t, r.dz, r.oz);
      if (P.nDir > 0) local_estimate(P, G, rng, r, px, py, pz, w, comp, comp, order, cnt);   // INT:776-800
      if (P.opt.useRussianRoulette && w < P.opt.russianRouletteW * 0.5f) {     // INT:805-811
        if (rng.real() >= w / P.opt.russianRouletteW) w = 0.0f; else w = P.opt.russianRouletteW;
      }
      if (w <= TINY32) {
        cnt.c[CNT_RR_KILLS]++;
        state = ST_DEAD;
      } else {
        const int c = comp - 1;
        const int pidx = (int)__ldg(&P.idx16[cell + cells * (size_t)c]);
        const int nS = P.invS[c];
        const float *tab = P.inv[c] + (size_t)(pidx - 1) * nS;
        const float rn = rng.real();                                           // computeScatteringAngle INT:1594-1621
        const int k = (int)(rn * (float)nS) + 1;
        float theta;
        if (k < nS) {
          const float left = rn - (float)(k - 1) / (float)nS;
          theta = (1.0f - left) * __ldg(&tab[k - 1]) + left * __ldg(&tab[k]);
        } else {
          theta = __ldg(&tab[nS - 1]);
        }
        float sinT, cosT;
        __sincosf(theta, &sinT, &cosT);
        float AX, AY, D;                                                       // next_direct INT:1921-1948
        do {
          AX = 1.0f - 2.0f * rng.real();
          AY = 1.0f - 2.0f * rng.real();
          D = AX * AX + AY * AY;
        } while (D > 1.0f);
        float B = sinT * rsqrtf(D);
        AX *= B; AY *= B;
        B = r.dx * AX - r.dy * AY;
        D = cosT - B / (1.0f + fabsf(r.dz));
        const float ndx = r.dx * D + AX, ndy = r.dy * D - AY;
        const float ndz = r.dz * cosT - copysignf(fabsf(B), r.dz * B);
        r.ox = px; r.oy = py; r.oz = pz;
        r.dx = ndx; r.dy = ndy; r.dz = ndz;
        tau = -__logf(fmaxf(TINY32, rng.real()));
        ext = 0.0f;
        ray_start(r, G);
        state = ST_MARCH;
      }
    }
  }

  // ---- flush event counters: warp shuffle reduce, one atomic per warp ----
#pragma unroll
  for (int i = 0; i < CNT_N; ++i) {
    unsigned long long v = cnt.c[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    if (lane == 0 && v) atomicAdd(&P.counters[i], v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.tally[P.offPhotons], (double)nPhotons);
}

__global__ void philox_kat_kernel(uint64_t seed, uint64_t photon, int n, uint32_t *out) {
  Philox g;
  g.init(seed, photon);
  for (int i = 0; i < n; ++i) out[i] = g.next_u32();
}

}  // namespace mcbfast

void mcb_launch_fast_batch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
                           int numSMs, unsigned long long *workCounter, cudaStream_t stream) {
  if (nPhotons <= 0) return;
  constexpr int THREADS = 128;
  const size_t smem = sizeof(float) * (size_t)(P.nx + P.ny + P.nz + 3);
  static int blocksPerSM = 0;
  if (blocksPerSM == 0) {
    cudaFuncSetAttribute(mcbfast::batch_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, mcbfast::batch_kernel<THREADS>, THREADS, smem) != cudaSuccess ||
        blocksPerSM < 1)
      blocksPerSM = 1;
  }
  long long want = (nPhotons + THREADS - 1) / THREADS;
  long long cap = (long long)numSMs * blocksPerSM;          // persistent: every CTA resident, whole waves only
  const int blocks = (int)(want < cap ? want : cap);
  mcbfast::batch_kernel<THREADS><<<blocks, THREADS, smem, stream>>>(P, nPhotons, seed, firstPhotonId, workCounter, 12);
}

void mcb_launch_philox_kat(uint64_t seed, uint64_t photon, int n, uint32_t *out, cudaStream_t stream) {
  mcbfast::philox_kat_kernel<<<1, 1, 0, stream>>>(seed, photon, n, out);
}
