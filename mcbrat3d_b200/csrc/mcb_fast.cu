// mcb_fast.cu -- the throughput photon kernel (sm_100a).
//
// Same physics and event sequence as the reference's computeRT (INT:393-841) but laid out
// for the GPU instead of mirroring the reference's arithmetic:
//   * one persistent kernel, one photon per lane, lanes refilled from a global photon
//     counter (warp-aggregated atomic) so no lane idles while photons remain;
//   * ray marching (accumulateExtinctionAlongPath, OPT:1656-1815) as a parametric
//     Amanatides-Woo DDA in single precision that runs AHEAD of the extinction reads: the cells a
//     ray visits depend on geometry only, so a burst of B cells is stepped with pure ALU work, the
//     B gathers are issued together, and only then is optical depth accumulated and tested.  The
//     packed f32 extinction field carries a ghost shell (periodic replicas in x and y, empty cells
//     above and below), so the per-cell step has no wrap or exit tests at all -- those run once per
//     burst;
//   * the field is stored twice (mcb_device.cuh): as 2x2x2-cell bricks, one brick per 32-byte sector, for the
//     flux kernels (the measured limiter is the SM's L1TEX->XBAR request port, i.e. L1 misses: with bricks half
//     of a ray's steps stay inside the sector, whatever the axis), and x-fastest for local estimation, narrow
//     and irregular grids; fields too large for L2 add an occupancy bitmap so that clear-sky cells are never
//     gathered from HBM; everything a scattering event reads sits in one per-cell record (one gather per event);
//   * warp-level regrouping: lanes march in bursts and park when they reach an event (scatter /
//     surface / exit); once enough lanes are parked ONE event phase runs for all of them: tallies,
//     absorption, roulette, rebirth of finished lanes, one Philox block, the new direction, one
//     leg set-up -- so the expensive code executes once, with many active lanes;
//   * Philox4x32-10 per-photon streams (mcb_device.cuh) instead of a sequential MT19937;
//   * tallies: shared-memory-privatised f32 atomics flushed once per block when the column /
//     cell grid is small enough to be an atomic hot spot, f64 RED.ADD to the packed tally
//     buffer otherwise.  fluxAbsorbed is not tallied per event: the reference adds the same amount
//     to fluxAbsorbed(ix,iy) and volumeAbsorption(ix,iy,iz) (INT:765-771, 505-508), so the column
//     tally IS the column sum of the volume tally and a streaming kernel forms it after the launch
//     (one atomic per scattering event less on the request-bound path).
// Results agree with the reference arithmetic statistically (north-star criterion (b));
// bit-level trace parity is the job of mcb_reference.cu.
#include "mcb_march.cuh"
#include <cstdio>

namespace mcbfast {

// ---- local estimation: computeIntensityContribution (INT:1623-1832) as a warp-wide task queue ----
// The reference traces the nDir view rays of an event one after the other inside the photon loop.  Here an event
// POSTS a request (where, which way the photon was heading, weight, component, its Philox counter) into the
// warp's shared-memory slots; then ALL 32 lanes of the warp -- the marching ones included, they would idle during
// the event phase anyway -- pull (request, direction) tasks from a warp-local counter until none is left.  Every
// lane advances its current ray by one burst per iteration, so the marcher code stays convergent while rays of
// very different lengths are in flight; a lane that finishes a ray tallies it and takes the next task.  A lane that
// finds the queue empty goes back to its OWN photon leg (if it has one under way) and keeps marching it while the
// others finish their rays; a photon that reaches an event in here simply parks until the next event phase.
// Rays of very different lengths (a few cells inside the cloud, a hundred through clear sky to the top) leave a long
// tail: measured on C3 + 5 views, 74 % of the loop's iterations ran with the queue already empty and 7 lanes busy.
// So once the queue is empty and at most `carryThreshold` rays are still in flight, those rays are PARKED in shared
// memory (16 words) and the round ends; their lanes resume them at the start of the next round, next to the new
// requests, and a last round at the end of the kernel drains what is left.
#ifndef MCB_LE_BURST
#define MCB_LE_BURST 8          // cells per burst of a local-estimate ray (C3 + 5 views: 4 -> 5.9e7, 8 -> 6.4e7 photons/s)
#endif
#ifndef MCB_IRR_OCC
#define MCB_IRR_OCC 6            // CTAs per SM of the edge-table (stretched grid) flux kernel: 80 registers (4: 3.95e8, 6: 4.55e8, 8: 2.9e8 photons/s)
#endif
#ifndef MCB_LE_OCC
#define MCB_LE_OCC 5             // CTAs per SM of the local-estimation kernels: 96 registers, no spills (6: 80 registers, 314 B of spills, 8 % slower)
#endif
// cell indices travel as 16-bit fields of the request / parked-ray words (decoded unsigned): grids with views are
// limited to 65535 cells per axis, which mcb_api.cu enforces before the launch
#define LE_WORDS 13
#define LE_CARRY_WORDS 16      // a view ray parked between two rounds of the queue (see le_run)
enum { LE_PX = 0, LE_PY, LE_PZ, LE_DX, LE_DY, LE_DZ, LE_W, LE_IXY, LE_IZO, LE_COMP, LE_C0, LE_C1, LE_BLK };
enum { PH_IDLE = 0, PH_PLAIN, PH_E13, PH_E14A, PH_E14B, PH_PHOTON };

__device__ __forceinline__ void le_post(float *sle, int lane, const DevDomain &P, Rng &rng, const Ray &r0,
                                        float px, float py, float pz, float w, int component, int tallyComponent, int order,
                                        int pidx) {
  sle[LE_PX * 32 + lane] = px; sle[LE_PY * 32 + lane] = py; sle[LE_PZ * 32 + lane] = pz;
  sle[LE_DX * 32 + lane] = r0.dx; sle[LE_DY * 32 + lane] = r0.dy; sle[LE_DZ * 32 + lane] = r0.dz;
  sle[LE_W * 32 + lane] = w;
  sle[LE_IXY * 32 + lane] = __int_as_float(r0.ix | (r0.iy << 16));
  sle[LE_IZO * 32 + lane] = __int_as_float(r0.iz | (min(order, 65535) << 16));
  sle[LE_COMP * 32 + lane] = __int_as_float(((component + 1) & 0xff) | ((tallyComponent & 0xff) << 8) | (pidx << 16));
  sle[LE_C0 * 32 + lane] = __uint_as_float(rng.c0); sle[LE_C1 * 32 + lane] = __uint_as_float(rng.c1);
  sle[LE_BLK * 32 + lane] = __uint_as_float(rng.blk);
  // the request owns the next ceil(nDir/2) Philox blocks of this photon (one block serves two directions)
  if (P.opt.useRussianRouletteForIntensity) rng.blk += (uint32_t)((P.nDir + 1) >> 1);
}

template <bool REG, bool WIDE, bool MASK>
__device__ void le_run(const DevDomain &P, const Grid &G, const Tally &T, uint32_t k0, uint32_t k1, unsigned posted,
                       float *sle, unsigned *queue, int lane, Counts &cnt, Ray &photonRay, float &photonExt,
                       float photonTau, int &photonState, int carryThreshold) {
  const int nDir = P.nDir;
  const int nTasks = __popc(posted) * nDir;
  float *carry = carryThreshold >= 0 ? sle + LE_WORDS * 32 + 64 : nullptr;   // LE_CARRY_WORDS x 32, word-major like the slots
  const float invDir = 1.0f / (float)nDir;
  if (lane == 0) *queue = 0u;
  if ((posted >> lane) & 1u) queue[1 + __popc(posted & ((1u << lane) - 1u))] = (unsigned)lane;   // r-th request -> its lane
  __syncwarp();
  Ray r;
  r.ox = r.oy = r.oz = 0.0f; r.dx = r.dy = 0.0f; r.dz = 1.0f; r.rx = r.ry = r.rz = FLT_MAX;
  r.t = 0.0f; r.tx = r.ty = r.tz = FLT_MAX; r.ix = r.iy = r.iz = 0;
  float ext = 0.0f, tgt = FLT_MAX, w = 0.0f, npf = 0.0f, tauFree = 0.0f, uTest = 0.0f;
  int phase = PH_IDLE, dir = 0, comps = 0;
  bool done = false, mine = false;                       // mine: r holds this lane's own photon leg
  if (carry) {                                           // resume the ray this lane parked in the previous round
    const int packed = __float_as_int(carry[15 * 32 + lane]);
    if (packed != 0) {
      carry[15 * 32 + lane] = 0.0f;
      phase = (packed >> 16) & 0xf; dir = (packed >> 20) & 0xff; r.iz = packed & 0xffff;
      r.ox = carry[0 * 32 + lane]; r.oy = carry[1 * 32 + lane]; r.oz = carry[2 * 32 + lane];
      r.t = carry[3 * 32 + lane]; r.tx = carry[4 * 32 + lane]; r.ty = carry[5 * 32 + lane]; r.tz = carry[6 * 32 + lane];
      ext = carry[7 * 32 + lane]; tgt = carry[8 * 32 + lane]; w = carry[9 * 32 + lane]; npf = carry[10 * 32 + lane];
      tauFree = carry[11 * 32 + lane]; uTest = carry[12 * 32 + lane];
      const int ixy = __float_as_int(carry[13 * 32 + lane]);
      r.ix = ixy & 0xffff; r.iy = (int)((uint32_t)ixy >> 16);
      comps = __float_as_int(carry[14 * 32 + lane]);
      r.dx = P.viewDir[3 * dir]; r.dy = P.viewDir[3 * dir + 1]; r.dz = P.viewDir[3 * dir + 2];
      r.rx = safe_rcp(r.dx); r.ry = safe_rcp(r.dy); r.rz = safe_rcp(r.dz);
    }
  }
  for (;;) {
    if (phase == PH_IDLE && !done) {                     // take the next (request, direction) task
      const int t = (int)atomicAdd(queue, 1u);
      if (t >= nTasks) {
        done = true;
        if (photonState == ST_MARCH) { r = photonRay; ext = photonExt; tgt = photonTau; phase = PH_PHOTON; mine = true; }
      } else {
        const int req = (int)(((float)t + 0.5f) * invDir);        // t / nDir (t < 32 * MCB_MAX_DIR: exact)
        dir = t - req * nDir;
        const int s = (int)queue[1 + req];                         // lane that posted the req-th request
        const float vx = P.viewDir[3 * dir], vy = P.viewDir[3 * dir + 1], vz = P.viewDir[3 * dir + 2];
        const int ixy = __float_as_int(sle[LE_IXY * 32 + s]), izo = __float_as_int(sle[LE_IZO * 32 + s]);
        comps = __float_as_int(sle[LE_COMP * 32 + s]);
        const int component = (comps & 0xff) - 1, order = (int)((uint32_t)izo >> 16);
        r.ix = ixy & 0xffff; r.iy = (int)((uint32_t)ixy >> 16); r.iz = izo & 0xffff;
        w = sle[LE_W * 32 + s];
        if (component == 0) {
          // INT:1694.  A photon BORN at the surface (thermal source, scattering order 0) gives nothing to a downward view:
          // its view ray has no first step, the reference's marcher signals an error and INT:1745-1751 drop the
          // contribution (a photon REFLECTED there, order >= 1, does contribute weight / pi) -- tests/test_first_interaction.py
          npf = (order == 0 && vz < 0.0f) ? 0.0f : 1.0f / PI32;
        } else if (component < 0) {
          npf = P.viewNorm[dir];                                                   // INT:1696
        } else {
          float proj = sle[LE_DX * 32 + s] * vx + sle[LE_DY * 32 + s] * vy + sle[LE_DZ * 32 + s] * vz;   // INT:1704-1706
          proj = fminf(fmaxf(proj, -1.0f), 1.0f);
          const int c = component - 1;
          const int pidx = (int)((uint32_t)comps >> 16);                           // phase-function entry of the event's cell
          const float *tab = ((P.opt.useHybridPhaseFunsForIntenCalcs && order <= P.opt.numOrdersOrigPhaseFunIntenCalcs)
                                  ? P.fwdOrig[c] : P.fwd[c]) + (size_t)MCB_CHECK_INDEX(P, pidx - 1, P.fwdE[c]) * P.fwdS[c];
          const int nS = P.fwdS[c];                                                // INT:1855-1870: linear in the angle
          const float pos = acosf(proj) * P.fwdInvDTheta[c];                       // angle / dTheta
          const int ai = (int)pos + 1;
          float val;
          if (ai < nS) {
            const float wt = 1.0f - (pos - (float)(ai - 1));
            val = wt * __ldg(&tab[MCB_CHECK_INDEX(P, ai - 1, nS)]) + (1.0f - wt) * __ldg(&tab[MCB_CHECK_INDEX(P, ai, nS)]);
          } else {
            val = __ldg(&tab[nS - 1]);
          }
          npf = val * P.viewNorm[dir];                                             // INT:1726
        }
        r.ox = sle[LE_PX * 32 + s]; r.oy = sle[LE_PY * 32 + s]; r.oz = sle[LE_PZ * 32 + s];
        r.dx = vx; r.dy = vy; r.dz = vz;
        ray_start<REG>(r, P, G);
        ext = 0.0f;
        cnt.leRays++;
        if (!P.opt.useRussianRouletteForIntensity) {                               // INT:1729-1752
          tgt = FLT_MAX; phase = PH_PLAIN;
        } else {                                                                   // INT:1753-1813
          Rng q;
          q.c0 = __float_as_uint(sle[LE_C0 * 32 + s]); q.c1 = __float_as_uint(sle[LE_C1 * 32 + s]);
          q.blk = __float_as_uint(sle[LE_BLK * 32 + s]) + (uint32_t)(dir >> 1);
          const float4 u = q.block(k0, k1);
          const float uFree = (dir & 1) ? u.z : u.x;
          uTest = (dir & 1) ? u.w : u.y;
          tauFree = -__logf(fmaxf(TINY32, uFree));
          if (PI32 * npf <= P.opt.zetaMin) { tgt = tauFree; phase = PH_E13; }       // Iwabuchi (2006) Eq 13
          else { tgt = -__logf(P.opt.zetaMin / fmaxf(TINY32, PI32 * npf)); phase = PH_E14A; }   // Eq 14
        }
      }
    }
    {                                                    // queue empty and few rays left: park them, end the round
      const unsigned flying = __ballot_sync(FULL, phase != PH_IDLE && phase != PH_PHOTON);
      if ((nTasks == 0 || __any_sync(FULL, done)) && __popc(flying) <= max(carryThreshold, 0)) break;
    }
#ifdef MCB_LE_STATS
    {
      const unsigned act = __ballot_sync(FULL, phase != PH_IDLE && phase != PH_PHOTON), ph = __ballot_sync(FULL, phase == PH_PHOTON);
      if (lane == 0) {
        atomicAdd(&P.counters[16], 1ull); atomicAdd(&P.counters[17], (unsigned long long)__popc(act));
        atomicAdd(&P.counters[18], (unsigned long long)__popc(ph));
        if (*queue >= (unsigned)nTasks) { atomicAdd(&P.counters[19], 1ull); atomicAdd(&P.counters[20], (unsigned long long)__popc(act)); }
      }
    }
#endif
    if (phase != PH_IDLE) {
      unsigned crossed = 0u;
      const int ev = march_burst<REG, WIDE, MCB_LE_BURST, MASK, false>(r, P, G, ext, tgt, crossed);
      if (phase == PH_PHOTON) {
        cnt.crossings += crossed;
        if (ev != MARCH_ON) { photonState = ev; phase = PH_IDLE; }   // parked until the next event phase
      } else if ((cnt.leCrossings += crossed, ev != MARCH_ON)) {
        float contribution = 0.0f;
        bool finished = true;
        if (phase == PH_PLAIN) {                         // "extinction values < 0" signal a failed trace: no contribution (INT:1745-1751)
          contribution = ext < -1.0e-3f ? 0.0f : w * npf * __expf(-ext);
        } else if (phase == PH_E13) {
          contribution = (ev == MARCH_TOP && uTest <= PI32 * npf / P.opt.zetaMin) ? w * P.opt.zetaMin / PI32 : 0.0f;
        } else if (phase == PH_E14A) {
          if (ev == MARCH_TOP) {
            contribution = ext < -1.0e-3f ? 0.0f : w * npf * __expf(-ext);
          } else if (ev == MARCH_HIT) {                  // continue from where the first trace stopped (INT:1793-1795)
            float qx, qy, qz;
            ray_position(r, P, qx, qy, qz);
            if (!REG) seam_fix(P, G, r.ix, r.iy, qx, qy);
            r.ox = qx; r.oy = qy; r.oz = qz;
            ray_start<REG>(r, P, G);
            ext = 0.0f; tgt = tauFree; phase = PH_E14B;
            finished = false;
          }                                              // left through the surface before tauMax: nothing
        } else {                                         // PH_E14B
          contribution = ev == MARCH_TOP ? w * P.opt.zetaMin / PI32 : 0.0f;
        }
        if (finished) {
          const int component = (comps & 0xff) - 1, tallyComponent = (comps >> 8) & 0xff;
          if (P.opt.limitIntensityContributions && contribution > P.opt.maxIntensityContribution) {   // INT:1815-1826
            const int cslot = component < 0 ? 0 : component;
            atomicAdd(&P.tally[P.offExcess + dir + (long long)P.nDir * cslot],
                      (double)(contribution - P.opt.maxIntensityContribution));
            contribution = P.opt.maxIntensityContribution;
          }
          if (contribution != 0.0f) add_intensity(P, T, dir, r.ix + P.nx * r.iy, tallyComponent, contribution);
          phase = PH_IDLE;
        }
      }
    }
  }
  if (mine) { photonRay = r; photonExt = ext; }
  else if (carry && phase != PH_IDLE) {                  // park the unfinished view ray
    carry[0 * 32 + lane] = r.ox; carry[1 * 32 + lane] = r.oy; carry[2 * 32 + lane] = r.oz;
    carry[3 * 32 + lane] = r.t; carry[4 * 32 + lane] = r.tx; carry[5 * 32 + lane] = r.ty; carry[6 * 32 + lane] = r.tz;
    carry[7 * 32 + lane] = ext; carry[8 * 32 + lane] = tgt; carry[9 * 32 + lane] = w; carry[10 * 32 + lane] = npf;
    carry[11 * 32 + lane] = tauFree; carry[12 * 32 + lane] = uTest;
    carry[13 * 32 + lane] = __int_as_float(r.ix | (r.iy << 16));
    carry[14 * 32 + lane] = __int_as_float(comps);
    carry[15 * 32 + lane] = __int_as_float(r.iz | (phase << 16) | (dir << 20));
  }
  __syncwarp();
}

template <int THREADS, bool REG, bool WIDE, int MINBLOCKS, int BURST, bool LE, bool MASK, bool BRICK>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
batch_kernel(const __grid_constant__ DevDomain P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
             unsigned long long *workCounter, int parkThreshold, int leCarry, const SmemPlan plan) {
  extern __shared__ float smem[];
  __shared__ unsigned sCnt[4];            // rare events: surface hits, surface kills, roulette kills
  const int cols = P.nx * P.ny, cells = cols * P.nz;
  Grid G;
  G.sx = G.sy = G.sz = nullptr;
  if (!REG) {
    // ghost-extended edges: x, y continue periodically, z continues with the boundary spacing
    float *sx = smem + plan.edgesOff + GH, *sy = sx + (P.nx + 1 + 2 * GH), *sz = sy + (P.ny + 1 + 2 * GH);
    for (int i = (int)threadIdx.x - GH; i <= P.nx + GH; i += THREADS) {
      const int j = i < 0 ? i + P.nx * ((-i + P.nx - 1) / P.nx) : i;          // j in [0, ...)
      const int q = j / P.nx, m = j - q * P.nx;
      const int shift = (i < 0 ? -((-i + P.nx - 1) / P.nx) : 0) + q;
      sx[i] = (float)(P.xE[m] + (double)shift * (P.xMax - P.x0));
    }
    for (int i = (int)threadIdx.x - GH; i <= P.ny + GH; i += THREADS) {
      const int j = i < 0 ? i + P.ny * ((-i + P.ny - 1) / P.ny) : i;
      const int q = j / P.ny, m = j - q * P.ny;
      const int shift = (i < 0 ? -((-i + P.ny - 1) / P.ny) : 0) + q;
      sy[i] = (float)(P.yE[m] + (double)shift * (P.yMax - P.y0));
    }
    for (int i = (int)threadIdx.x - GH; i <= P.nz + GH; i += THREADS) {
      double e;
      if (i < 0) e = P.zE[0] + (double)i * (P.zE[1] - P.zE[0]);
      else if (i > P.nz) e = P.zE[P.nz] + (double)(i - P.nz) * (P.zE[P.nz] - P.zE[P.nz - 1]);
      else e = P.zE[i];
      sz[i] = (float)e;
    }
    G.sx = sx; G.sy = sy; G.sz = sz;
  }
  Tally T;
  T.cols = cols;
  T.sFlux = plan.fluxOff >= 0 ? smem + plan.fluxOff : nullptr;
  T.sVol = plan.volOff >= 0 ? smem + plan.volOff : nullptr;
  T.sInt = plan.intOff >= 0 ? smem + plan.intOff : nullptr;
  if (T.sFlux) for (int i = threadIdx.x; i < 2 * cols; i += THREADS) T.sFlux[i] = 0.0f;
  if (T.sVol) for (int i = threadIdx.x; i < cells; i += THREADS) T.sVol[i] = 0.0f;
  if (T.sInt) for (int i = threadIdx.x; i < cols * P.nDir; i += THREADS) T.sInt[i] = 0.0f;
  if (threadIdx.x < 4) sCnt[threadIdx.x] = 0u;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  float *sle = nullptr;
  unsigned *leQueue = nullptr;
  if (LE) {
    sle = smem + plan.leOff + (threadIdx.x >> 5) * plan.leStride;
    leQueue = (unsigned *)(sle + LE_WORDS * 32);
    if (leCarry >= 0) sle[LE_WORDS * 32 + 64 + 15 * 32 + lane] = 0.0f;      // no parked view ray
    __syncwarp();
  }
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);

  Counts cnt{0u, 0u, 0u, 0u};
  Rng rng;
  Ray r;
  float w = 0.0f, tau = 0.0f, ext = 0.0f, uNext = 0.0f;
  int order = 0;
  int state = ST_DEAD;
  bool more = true;                      // photons may remain in the global counter

  for (;;) {
    // =========================== event phase: every lane that is not marching ===========================
    float px = 0.0f, py = 0.0f, pz = 0.0f;
    int comp = 1, cell = 0, pidx = 0;
    bool posted = false;
    if (state == ST_TOP) {                                                     // INT:573-617
      add_flux(P, T, 0, r.ix + P.nx * r.iy, w);
      state = ST_DEAD;
    } else if (state == ST_SURFACE) {                                          // INT:619-702
      add_flux(P, T, 1, r.ix + P.nx * r.iy, w);
      atomicAdd(&sCnt[0], 1u);
      order++;
      w = (float)((double)w * P.albedo);
      if (w <= TINY32) {
        atomicAdd(&sCnt[1], 1u);                                               // absorbed by the surface
        state = ST_DEAD;
      } else {
        ray_position(r, P, px, py, pz);
        pz = P.fz0;
        if (LE && P.nDir > 0) { le_post(sle, lane, P, rng, r, px, py, pz, w, 0, 0, order, 0); posted = true; }
      }
    } else if (state == ST_SCATTER) {                                          // INT:703-811
      order++;
      cnt.scatters++;
      cell = r.ix + P.nx * (r.iy + P.ny * r.iz);
      // uNext (drawn with the previous event's block) picks the component and, rescaled to the chosen
      // component's interval (uniform again, conditional on the pick), decides the roulette
      float lo = 0.0f, hi = 1.0f, ssa;
      {                                                                        // the cell's event record: ONE gather
        const uint32_t *R = P.rec + ((size_t)MCB_CHECK_INDEX(P, cell, cells) << P.recShift);
        if (P.nc == 1) {
          const uint2 v = __ldg(reinterpret_cast<const uint2 *>(R));
          ssa = __uint_as_float(v.x); pidx = (int)(v.y & 0xffffu);
        } else if (P.nc == 2) {
          const uint4 v = __ldg(reinterpret_cast<const uint4 *>(R));
          const float cc = __uint_as_float(v.x);                               // findIndex on (0, cumExt(:)), NUM:262-315
          if (uNext >= cc) { comp = 2; lo = cc; } else { hi = cc; }
          ssa = __uint_as_float(comp == 1 ? v.y : v.z);
          pidx = (int)(comp == 1 ? (v.w & 0xffffu) : (v.w >> 16));
        } else {
          for (int c = 1; c < P.nc; ++c) {
            const float cc = __uint_as_float(__ldg(R + (c - 1)));
            if (uNext >= cc) { comp = c + 1; lo = cc; } else { hi = fminf(hi, cc); }
          }
          ssa = __uint_as_float(__ldg(R + (P.nc - 1) + (comp - 1)));
          const uint32_t pw = __ldg(R + (2 * P.nc - 1) + ((comp - 1) >> 1));
          pidx = (int)(((comp - 1) & 1) ? (pw >> 16) : (pw & 0xffffu));
        }
      }
      pidx = max(pidx, 1);                   // entry 0 marks a cell the component is absent from: never index before the table
      if (ssa < 1.0f) {                                                        // INT:765-771
        const float absorbed = w * (1.0f - ssa);
        add_vol(P, T, cell, absorbed);                     // fluxAbsorbed = column sum of this tally (column_absorption_kernel)
        w *= ssa;
      }
      ray_position(r, P, px, py, pz);
      if (!REG) seam_fix(P, G, r.ix, r.iy, px, py);
      if (LE && P.nDir > 0) { le_post(sle, lane, P, rng, r, px, py, pz, w, comp, comp, order, pidx); posted = true; }   // INT:776-800
      if (P.opt.useRussianRoulette && w < P.opt.russianRouletteW * 0.5f) {     // INT:805-811
        const float uRR = P.nc > 1 ? __fdividef(uNext - lo, fmaxf(hi - lo, TINY32)) : uNext;
        if (uRR >= w / P.opt.russianRouletteW) w = 0.0f; else w = P.opt.russianRouletteW;
      }
      if (w <= TINY32) {
        atomicAdd(&sCnt[2], 1u);                                               // killed by roulette
        state = ST_DEAD;
      }
    }
    if (LE && P.nDir > 0 && P.opt.LW_flag > 0.0f) {      // thermal runs: births post too (below), one slot per lane
      const unsigned pm = __ballot_sync(FULL, posted);
      // This round runs in the MIDDLE of the event phase: a marching lane must not take its own photon along (it could
      // come back at an event whose tallies, roulette and new leg this phase has already passed), so the queue sees
      // every lane as parked.
      int parked = ST_DEAD;
      if (pm) le_run<REG, WIDE, MASK>(P, G, T, k0, k1, pm, sle, leQueue, lane, cnt, r, ext, tau, parked, leCarry);
      posted = false;
    }
    // ---- finished lanes take the next photons: one atomic per warp (getNextPhoton, ILL:561-590) ----
    {
      const unsigned dead = __ballot_sync(FULL, state == ST_DEAD);
      if (dead) {
        if (more) {
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(workCounter, (unsigned long long)__popc(dead));
          base = __shfl_sync(FULL, base, 0);
          if (state == ST_DEAD) {
            const unsigned long long p = base + (unsigned long long)__popc(dead & ((1u << lane) - 1u));
            if (p < (unsigned long long)nPhotons) { rng.init(firstPhotonId + p); state = ST_BORN; }
            else state = ST_DONE;
          }
          if (base + (unsigned long long)__popc(dead) >= (unsigned long long)nPhotons) more = false;
        } else if (state == ST_DEAD) {
          state = ST_DONE;
        }
      }
    }
    if (__all_sync(FULL, state == ST_DONE)) {
      // last round of the view-ray queue: the requests of the final events and every parked ray, to the end
      if (LE && P.nDir > 0)
        le_run<REG, WIDE, MASK>(P, G, T, k0, k1, __ballot_sync(FULL, posted), sle, leQueue, lane, cnt, r, ext, tau, state, min(leCarry, 0));
      break;
    }
    // ---- one Philox block for every parked lane: (angle | position, azimuth | position, next optical depth, next pick) ----
    if (state != ST_MARCH && state != ST_DONE) {
      const float4 u = rng.block(k0, k1);
      float uTau = u.z;
      if (state == ST_BORN) {
        float x01, y01, z01;
        int bi = -1, bj = 0, bk = 0;                                           // birth cell of an atmospheric emission
        w = 1.0f; order = 0;
        if (P.source == 0) {                                                   // ILL:88-96
          x01 = u.x; y01 = u.y; z01 = 1.0f - FLT_EPSILON;
          r.dx = P.solarDir[0]; r.dy = P.solarDir[1]; r.dz = P.solarDir[2];
        } else {                                                               // ILL:481-515
          const float4 v = rng.block(k0, k1);
          float mu, phi;
          if ((double)u.x > P.fracAtmsPower) {                                 // surface emission
            x01 = u.y; y01 = v.w;
            mu = sqrtf(fmaxf(v.x, 1.0e-30f));                                  // ILL:489-491 retries on mu ~ 0
            phi = v.y * 2.0f * PI32;
            z01 = 0.0f;
          } else {                                                             // atmospheric emission
            const float q = u.y;
            // level and column from the compact column weights colCDF(ny,nz) = voxelCDF(nx,:,:) (EMI:56-57)
            const int ik = cdf_search(P.colCDF + (P.ny - 1), P.nz, (long long)P.ny, q);
            const int ij = cdf_search(P.colCDF + (size_t)P.ny * (size_t)(ik - 1), P.ny, 1, q);
            const double *voxBase = P.voxelCDF + (size_t)P.nx * ((size_t)(ij - 1) + (size_t)P.ny * (size_t)(ik - 1));
            const int ii = cdf_search(voxBase, P.nx, 1, q);
            // uniform inside the chosen cell, nudged off its faces (ILL:500-505)
            const float4 v2 = rng.block(k0, k1);
            z01 = ((float)(ik - 1) + fminf(fmaxf(v2.x, 1e-6f), 1.0f - 1e-6f)) / (float)P.nz;
            x01 = ((float)(ii - 1) + fminf(v2.y, 1.0f - 1e-6f)) / (float)P.nx;
            y01 = ((float)(ij - 1) + fminf(v2.z, 1.0f - 1e-6f)) / (float)P.ny;
            bi = ii - 1; bj = ij - 1; bk = ik - 1;           // the sampled cell itself: (cell + offset) / n may round across a face
            mu = 1.0f - 2.0f * v.x;                                            // ILL:507-509 retries on mu ~ 0
            if (!(fabsf(mu) > 2.0f * TINY32)) mu = 1.0e-30f;
            phi = v.y * 2.0f * PI32;
          }
          dir_from(mu, phi, r.dx, r.dy, r.dz);
        }
        // INT:478-494: unit square -> domain
        px = fmaf(x01, P.fLx, P.fx0); py = fmaf(y01, P.fLy, P.fy0);
        if (REG) {
          r.ix = min((int)(x01 * (float)P.nx), P.nx - 1);
          r.iy = min((int)(y01 * (float)P.ny), P.ny - 1);
          pz = fmaf(z01, P.fLz, P.fz0);
          r.iz = min((int)(z01 * (float)P.nz), P.nz - 1);
        } else {
          r.ix = find_cell(G.sx, P.nx, px);
          r.iy = find_cell(G.sy, P.ny, py);
          if (P.zRegular) {
            pz = fmaf(z01, P.fLz, P.fz0);
            r.iz = find_cell(G.sz, P.nz, pz);
          } else {                                                             // INT:491-493
            const float zs = z01 * (float)P.nz;
            r.iz = min((int)zs, P.nz - 1);
            pz = G.sz[r.iz] + (zs - (float)r.iz) * (G.sz[r.iz + 1] - G.sz[r.iz]);
          }
        }
        // uniform grids: unit-square fraction and cell coincide.  On stretched grids the reference maps the fraction
        // linearly into the domain and looks the cell up from the position (INT:478-494), which is kept above.
        if (REG && bi >= 0) { r.ix = bi; r.iy = bj; r.iz = bk; }
        if (P.opt.LW_flag > 0.0f) {                                            // INT:504-542
          if (pz > 0.0f) {
            add_vol(P, T, r.ix + P.nx * (r.iy + P.ny * r.iz), -1.0f);
          }
          if (LE && P.nDir > 0) { le_post(sle, lane, P, rng, r, px, py, pz, w, pz == 0.0f ? 0 : -1, 0, order, 0); posted = true; }
        }
      } else if (state == ST_SURFACE) {                                        // INT:655-676
        const float mu = sqrtf(fmaxf(u.x, 1.0e-30f));                          // retries on mu ~ 0
        dir_from(mu, 2.0f * PI32 * u.y, r.dx, r.dy, r.dz);
        r.iz = 0;
      } else {                                                                 // ST_SCATTER, INT:813-819
        const int c = comp - 1;
        const int nS = P.invS[c];
        const float *tab = P.inv[c] + (size_t)MCB_CHECK_INDEX(P, pidx - 1, P.invE[c]) * nS;
        const float rn = u.x;                                                  // computeScatteringAngle INT:1594-1621
        const int k = (int)(rn * (float)nS) + 1;
        float theta;
        if (k < nS) {
          const float left = rn - (float)(k - 1) / (float)nS;
          theta = (1.0f - left) * __ldg(&tab[MCB_CHECK_INDEX(P, k - 1, nS)]) + left * __ldg(&tab[MCB_CHECK_INDEX(P, k, nS)]);
        } else {
          theta = __ldg(&tab[nS - 1]);
        }
        float sinT, cosT;
        __sincosf(theta, &sinT, &cosT);
        // next_direct INT:1921-1948: the reference rejection-samples (AX, AY) in the unit disc and
        // normalises it, i.e. draws a uniform azimuth; the azimuth is drawn directly here.
        float AX, AY;
        __sincosf(2.0f * PI32 * u.y, &AY, &AX);
        AX *= sinT; AY *= sinT;
        const float B = r.dx * AX - r.dy * AY;
        const float D = cosT - __fdividef(B, 1.0f + fabsf(r.dz));
        const float ndx = r.dx * D + AX, ndy = r.dy * D - AY;
        const float ndz = r.dz * cosT - copysignf(fabsf(B), r.dz * B);
        r.dx = ndx; r.dy = ndy; r.dz = ndz;
      }
      // ---- common: start the next leg ----
      r.ox = px; r.oy = py; r.oz = pz;
      tau = -__logf(fmaxf(TINY32, uTau));                                      // INT:554
      uNext = u.w;
      ext = 0.0f;
      ray_start<REG>(r, P, G);
      state = ST_MARCH;
    }
    // the whole warp serves the posted requests (events above, emission at birth INT:513-542); every photon has its
    // next leg by now, so lanes that run out of requests march their own photon meanwhile
    if (LE && P.nDir > 0) {
      const unsigned pm = __ballot_sync(FULL, posted);
      if (pm) le_run<REG, WIDE, MASK>(P, G, T, k0, k1, pm, sle, leQueue, lane, cnt, r, ext, tau, state, leCarry);
    }

    // =========================== march phase: bursts until enough lanes are parked ===========================
    for (;;) {
      const unsigned mk = __ballot_sync(FULL, state == ST_MARCH);
      const unsigned live = __ballot_sync(FULL, state != ST_DONE);
      if (mk == 0u || __popc(live & ~mk) >= parkThreshold) break;
      if (state == ST_MARCH) {
        state = march_burst<REG, WIDE, BURST, MASK, BRICK>(r, P, G, ext, tau, cnt.crossings);
      }
    }
  }

  // ---- flush: event counters (warp shuffle reduce, one atomic per warp) ----
  {
    unsigned long long v[4] = {cnt.crossings, cnt.scatters, cnt.leRays, cnt.leCrossings};
    const int slot[4] = {CNT_CROSSINGS, CNT_SCATTERS, CNT_LE_RAYS, CNT_LE_CROSSINGS};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(FULL, v[i], o);
      if (lane == 0 && v[i]) atomicAdd(&P.counters[slot[i]], v[i]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sCnt[0]) atomicAdd(&P.counters[CNT_SURFACE], (unsigned long long)sCnt[0]);
    if (sCnt[1]) atomicAdd(&P.counters[CNT_SURFACE_KILLS], (unsigned long long)sCnt[1]);
    if (sCnt[2]) atomicAdd(&P.counters[CNT_RR_KILLS], (unsigned long long)sCnt[2]);
  }
  // ---- flush: privatised tallies, once per block, into the f64 tally buffer ----
  if (T.sFlux)
    for (int i = threadIdx.x; i < 2 * cols; i += THREADS) {
      const float v = T.sFlux[i];
      if (v != 0.0f) atomicAdd(&P.tally[P.offFluxUp + i], (double)v);       // fluxUp|fluxDown are contiguous
    }
  if (T.sVol)
    for (int i = threadIdx.x; i < cells; i += THREADS) {
      const float v = T.sVol[i];
      if (v != 0.0f) atomicAdd(&P.tally[P.offVolAbs + i], (double)v);
    }
  if (T.sInt)
    for (int i = threadIdx.x; i < cols * P.nDir; i += THREADS) {
      const float v = T.sInt[i];
      if (v != 0.0f) atomicAdd(&P.tally[P.offInt + i], (double)v);
    }
#ifdef MCB_LE_STATS
  if (blockIdx.x == 0 && threadIdx.x == 0)
    printf("LE_STATS (previous launches) iterations %llu le-lanes %llu photon-lanes %llu tail-iterations %llu tail-le-lanes %llu\n",
           P.counters[16], P.counters[17], P.counters[18], P.counters[19], P.counters[20]);
#endif
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&P.tally[P.offPhotons], (double)nPhotons);
    atomicAdd(&P.counters[CNT_PHOTONS], (unsigned long long)nPhotons);
  }
}

// fluxAbsorbed(ix,iy) = sum over iz of volumeAbsorption(ix,iy,iz) for everything accumulated so far (both kernels
// keep the two tallies equal by construction; the reference-arithmetic kernel still adds to both as the reference does)
__global__ void column_absorption_kernel(const DevDomain P) {
  const long long cols = (long long)P.nx * P.ny;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < cols; c += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < P.nz; ++k) s += P.tally[P.offVolAbs + c + cols * k];
    P.tally[P.offFluxAbs + c] = s;
  }
}

__global__ void philox_kat_kernel(uint64_t seed, uint64_t photon, int n, uint32_t *out) {
  Philox g;
  g.init(seed, photon);
  for (int i = 0; i < n; ++i) out[i] = g.next_u32();
}

}  // namespace mcbfast

template <bool REG, bool WIDE, int MINBLOCKS, int BURST, bool LE, bool MASK, bool BRICK>
static void launch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId, int numSMs,
                   unsigned long long *workCounter, cudaStream_t stream) {
  constexpr int THREADS = 128;
  auto kernel = mcbfast::batch_kernel<THREADS, REG, WIDE, MINBLOCKS, BURST, LE, MASK, BRICK>;
  // shared-memory plan: privatise the tallies when the column / cell grid is small enough to be an
  // atomic hot spot (homogeneous slabs, the 32-column step cloud); large grids spread their
  // atomics over many L2 lines and go straight to the f64 buffer.
  const int cols = P.nx * P.ny, cells = cols * P.nz;
  mcbfast::SmemPlan plan{-1, -1, -1, -1, -1, 0, 0};
  int off = 0;
  if (!REG) { plan.edgesOff = off; off += P.nx + P.ny + P.nz + 3 + 6 * MCB_GHOST; }
  const int budgetFloats = 9 * 1024;            // 36 KB per block keeps >= 6 blocks/SM resident
  if (cols <= 1024 && off + 2 * cols <= budgetFloats) { plan.fluxOff = off; off += 2 * cols; }
  if (cells <= 8192 && off + cells <= budgetFloats) { plan.volOff = off; off += cells; }
  if (P.nDir > 0 && cols * P.nDir <= 2048 && off + cols * P.nDir <= budgetFloats) { plan.intOff = off; off += cols * P.nDir; }
  // view rays still in flight when a round of the queue ends with no tasks left: at most `carry` of them are parked
  // (le_run).  That pays when a ray can be much longer than a burst -- grids deeper than 64 layers; on shallow grids
  // (the step cloud: 32 layers, 3.6 cells per ray) the 8 KB per block it takes away from L1 cost 10 %, so it is off (-1)
  const int carry = P.opt.tuneLeCarry < 0 ? -1 : P.opt.tuneLeCarry > 0 ? P.opt.tuneLeCarry : (P.nz > 64 ? 12 : -1);
  if (LE) {
    plan.leOff = off; plan.leStride = LE_WORDS * 32 + 64 + (carry >= 0 ? LE_CARRY_WORDS * 32 : 0);
    off += (THREADS / 32) * plan.leStride;
  }
  plan.totalFloats = off;
  const size_t smem = sizeof(float) * (size_t)off;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int blocksPerSM = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kernel, THREADS, smem) != cudaSuccess || blocksPerSM < 1)
    blocksPerSM = 1;
  const long long want = (nPhotons + THREADS - 1) / THREADS;
  const long long cap = (long long)numSMs * blocksPerSM;    // persistent: every CTA resident, whole waves only
  const int blocks = (int)(want < cap ? want : cap);
  // lanes parked before an event phase runs: 16 for flux runs; with local estimation the event phase is long and
  // shared by all 32 lanes, so waiting for 24 pays (C3 + 5 views: 16 -> 4.0e7, 24 -> 4.3e7 photons/s)
  int park = P.opt.tuneParkThreshold > 0 ? P.opt.tuneParkThreshold : (LE ? 24 : 16);
  if (park > 32) park = 32;
  kernel<<<blocks, THREADS, smem, stream>>>(P, nPhotons, seed, firstPhotonId, workCounter, park, carry, plan);
}

// mcb_pool.cu, mcb_pool_le.cu
bool mcb_pool_covers(const DevDomain &P);
bool mcb_pool_preferred(const DevDomain &P);
void mcb_launch_pool_batch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
                           int numSMs, unsigned long long *workCounter, cudaStream_t stream);
bool mcb_pool_le_covers(const DevDomain &P);
bool mcb_pool_le_preferred(const DevDomain &P);
bool mcb_pool_le_reads_bricks(const DevDomain &P);
void mcb_launch_pool_le_batch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
                              int numSMs, unsigned long long *workCounter, cudaStream_t stream);
#ifndef MCB_POOL_LE_DEFAULT
#define MCB_POOL_LE_DEFAULT 1                  // runs with view directions: 1 = the pool organisation is the default
#endif                                         // (one B200, r02, C3 + 5 views: 1.05e8 photons/s against the task queue's 8.5e7)

// which of the two organisations traces this run: mcb_options.tuneKernel, or where none is asked for the one that
// measured faster (flux-only: mcb_pool_preferred)
static bool runs_on_pool(const DevDomain &P) {
  if (P.nDir > 0) return mcb_pool_le_covers(P) && (P.opt.tuneKernel ? P.opt.tuneKernel == MCB_KERNEL_POOL
                                                                     : MCB_POOL_LE_DEFAULT != 0 && mcb_pool_le_preferred(P));
  return mcb_pool_covers(P) && (P.opt.tuneKernel ? P.opt.tuneKernel == MCB_KERNEL_POOL : mcb_pool_preferred(P));
}

// which layout of the extinction field the dispatcher below reads (the API packs that one)
bool mcb_fast_reads_bricks(const DevDomain &P) {
  const bool wide = P.nx >= MCB_GHOST && P.ny >= MCB_GHOST;
  if (P.nDir > 0) return runs_on_pool(P) && mcb_pool_le_reads_bricks(P);
  return P.uniform && wide && P.opt.tuneLayout != MCB_LAYOUT_LINEAR;
}


static void column_absorption(const DevDomain &P, int numSMs, cudaStream_t stream) {
  const long long cols = (long long)P.nx * P.ny;
  const int blocks = (int)((cols + 127) / 128 < (long long)numSMs * 8 ? (cols + 127) / 128 : (long long)numSMs * 8);
  mcbfast::column_absorption_kernel<<<blocks, 128, 0, stream>>>(P);
}

void mcb_launch_fast_batch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
                           int numSMs, unsigned long long *workCounter, cudaStream_t stream) {
  if (nPhotons <= 0) return;
  // uniform grids: the photon-pool kernels (mcb_pool.cu, mcb_pool_le.cu) or the park/regroup kernel below
  if (runs_on_pool(P)) {
    if (P.nDir > 0) mcb_launch_pool_le_batch(P, nPhotons, seed, firstPhotonId, numSMs, workCounter, stream);
    else mcb_launch_pool_batch(P, nPhotons, seed, firstPhotonId, numSMs, workCounter, stream);
    column_absorption(P, numSMs, stream);
    return;
  }
  // Kernel variants (defaults chosen from measurements on one B200):
  //   * flux-only, regular grid at least a ghost shell wide: bricked field, 64 registers = 8 CTAs/SM
  //     (C3: x-fastest 5.8e8 -> bricked 6.3e8 photons/s);
  //   * local estimation: x-fastest field, 80 registers = 6 CTAs/SM (a second ray per lane; bricks cost it 10 %);
  //   * fields too large for L2: the occupancy-bitmap variants (flux: 80 registers);
  //   * narrow grids (a period shorter than the ghost shell) and irregular grids: x-fastest field.
  // mcb_options.tuneBlocksPerSM = 6 | 8 and tuneLayout = MCB_LAYOUT_LINEAR are measurement knobs of the flux-only kernel.
  const int occEnv = P.opt.tuneBlocksPerSM > 0 ? P.opt.tuneBlocksPerSM : -1, linEnv = P.opt.tuneLayout == MCB_LAYOUT_LINEAR ? 1 : 0;
#define MCB_GO(REG, WIDE, OCC, LE, MASK, BRICK) \
    launch<REG, WIDE, OCC, 8, LE, MASK, BRICK>(P, nPhotons, seed, firstPhotonId, numSMs, workCounter, stream)
  const bool wide = P.nx >= MCB_GHOST && P.ny >= MCB_GHOST;      // no grid period shorter than the ghost shell
  const bool le = P.nDir > 0, mask = P.lin.mask != nullptr;
  if (P.uniform) {
    if (!wide) { if (le) MCB_GO(true, false, MCB_LE_OCC, true, false, false); else MCB_GO(true, false, 8, false, false, false); }
    else if (le) { if (mask) MCB_GO(true, true, MCB_LE_OCC, true, true, false); else MCB_GO(true, true, MCB_LE_OCC, true, false, false); }
    else if (mask) {
      if (linEnv) MCB_GO(true, true, 6, false, true, false);
      else if (occEnv >= 8) MCB_GO(true, true, 8, false, true, true);
      else MCB_GO(true, true, 6, false, true, true);
    } else {
      if (linEnv) MCB_GO(true, true, 8, false, false, false);
      else if (occEnv > 0 && occEnv < 8) MCB_GO(true, true, 6, false, false, true);
      else MCB_GO(true, true, 8, false, false, true);
    }
  } else {
    if (le) MCB_GO(false, false, 4, true, false, false); else MCB_GO(false, false, MCB_IRR_OCC, false, false, false);
  }
#undef MCB_GO
  column_absorption(P, numSMs, stream);
}

void mcb_launch_philox_kat(uint64_t seed, uint64_t photon, int n, uint32_t *out, cudaStream_t stream) {
  mcbfast::philox_kat_kernel<<<1, 1, 0, stream>>>(seed, photon, n, out);
}
