// mcb_pool_le.cu -- local estimation (computeIntensityContribution, INT:1623-1832) on the photon-pool organisation of
// mcb_pool.cu: photon legs AND view rays are jobs of the same kind -- "march this ray until its optical-depth target
// or the boundary" -- and every lane of a warp always has one.
//
// The round-1 kernel (mcb_fast.cu, le_run) alternated between marching photons (one per lane, the others idle once
// their photon had reached an event) and serving the view rays of the posted events from a warp-local task counter,
// with parked rays, tail handling and the set-up of every ray (Philox block, acos, table look-up, logarithms) done by
// whichever lanes happened to be free: 16.8 of 32 lanes active per instruction, issue slots 63 % busy, 8.3e7 photons/s
// on C3 + 5 views (ncu r01).  Here a warp owns
//     a photon pool      64 records in two stacks (READY legs / EVENTs), exactly as in mcb_pool.cu;
//     a request ring     32 requests, one per scattering / surface / emission event that wants radiances: position,
//                        weight, cell, component -- and, PER VIEW DIRECTION, the three numbers a view ray needs
//                        (normalised phase function value, free optical path, test variate), which the event phase
//                        computes with all 32 lanes busy (one Philox block per direction pair, one acos + table
//                        look-up per direction);
// and the main loop is: event phase when 32 events wait and the ring has room; lanes without a job take the next view
// ray (request, direction) -- or, when none is waiting, a READY photon leg; one burst for all 32 lanes; lanes whose
// ray ended tally it (view rays), continue it (Russian-roulette variant 14, second leg) or push the photon onto EVENT.
// View rays have priority, so the ring drains before new events are made: an event phase starts once every waiting
// ray has been taken, and 32 slots always suffice (thermal runs, where a dying photon and its successor's birth may
// both post, handle 16 events per phase).
//
// Same physics, same random numbers per photon and per (event, direction) as le_run: with the same seed the two
// kernels trace the same photon histories and the same view rays (tests/test_gpu_pool.py compares ray counts exactly
// and radiances to summation order).  Scope: uniform grids, up to MCB_POOL_LE_MAXDIR view directions -- grids at least a
// ghost shell wide with every field variant of mcb_pool.cu except CROP (vacuum / clear-layer leaps for photon legs AND
// view rays), narrower ones (the 32 x 1 x 32 step cloud) on the plain x-fastest field; everything else stays on mcb_fast.cu.
#include "mcb_march.cuh"

namespace mcbpoolle {

using namespace mcbfast;

#ifndef MCB_PLE_SPLIT
#define MCB_PLE_SPLIT false       // gathers of a burst in two halves (A/B: make variant FLAGS=-DMCB_PLE_SPLIT=true)
#endif
#ifndef MCB_PLE_BURST
#define MCB_PLE_BURST 8
#endif
#define PLE_WORDS 17
#define PLE_SLOTS 64
#define PLE_REQS 32
#define MCB_POOL_LE_MAXDIR 8
// photon records (as mcb_pool.cu); the scattering order rides in the cell words:
//   EVENT: PX.. = leg origin, TAU = distance along the leg, IXY = padded address of the hit cell, IZK = kind << 28 | order
//   READY: PX.. = leg origin, TAU = target optical depth, IXY = ix | iy << 16, IZK = iz | order << 16, TX.. = face distances
enum { PW_PX = 0, PW_PY, PW_PZ, PW_DX, PW_DY, PW_DZ, PW_W, PW_TAU, PW_UNEXT, PW_C0, PW_C1, PW_BLK, PW_IXY, PW_IZK,
       PW_TX, PW_TY, PW_TZ };
// request records: RQ_T + 3 * dir + {0: npf, 1: tauFree, 2: uTest}
enum { RQ_PX = 0, RQ_PY, RQ_PZ, RQ_W, RQ_IXY, RQ_IZ, RQ_COMP, RQ_T };
enum { JOB_NONE = 0, JOB_PHOTON = 1, JOB_PLAIN = 2, JOB_E13 = 3, JOB_E14A = 4, JOB_E14B = 5 };

// normalised phase function value of one (event, view direction) pair: INT:1694-1726
__device__ __forceinline__ float view_phase_value(const DevDomain &P, int component, int pidx, int order, int dir,
                                                  float dx, float dy, float dz) {
  // Lambertian surface, INT:1694; born AT the surface (order 0) and viewed from below: nothing, as in the reference, whose
  // marcher signals an error for the zero-length view ray (INT:1745-1751; tests/test_first_interaction.py)
  if (component == 0) return (order == 0 && P.viewDir[3 * dir + 2] < 0.0f) ? 0.0f : 1.0f / PI32;
  if (component < 0) return P.viewNorm[dir];                                           // isotropic emission, INT:1696
  float proj = dx * P.viewDir[3 * dir] + dy * P.viewDir[3 * dir + 1] + dz * P.viewDir[3 * dir + 2];   // INT:1704-1706
  proj = fminf(fmaxf(proj, -1.0f), 1.0f);
  const int c = component - 1;
  const float *tab = ((P.opt.useHybridPhaseFunsForIntenCalcs && order <= P.opt.numOrdersOrigPhaseFunIntenCalcs)
                          ? P.fwdOrig[c] : P.fwd[c]) + (size_t)MCB_CHECK_INDEX(P, pidx - 1, P.fwdE[c]) * P.fwdS[c];
  const int nS = P.fwdS[c];                                                            // INT:1855-1870: linear in the angle
  const float pos = acosf(proj) * P.fwdInvDTheta[c];
  const int ai = (int)pos + 1;
  float val;
  if (ai < nS) {
    const float wt = 1.0f - (pos - (float)(ai - 1));
    val = wt * __ldg(&tab[MCB_CHECK_INDEX(P, ai - 1, nS)]) + (1.0f - wt) * __ldg(&tab[MCB_CHECK_INDEX(P, ai, nS)]);
  } else {
    val = __ldg(&tab[nS - 1]);
  }
  return val * P.viewNorm[dir];                                                        // INT:1726
}

// WIDE = false: grids with a period shorter than the ghost shell (the 32 x 1 x 32 step cloud): indices that ran into the
// shell are folded with a remainder instead of one conditional add; x-fastest field, no bitmap, no leaps.
template <int THREADS, int MINBLOCKS, int BURST, bool MASK, bool BRICK, bool LEAP, bool WIDE = true>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
pool_le_kernel(const __grid_constant__ DevDomain P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
               unsigned long long *workCounter, const SmemPlan plan, const float leapBelow, const int leapLanes) {
  extern __shared__ float smem[];
  __shared__ unsigned sCnt[6];
  const int cols = P.nx * P.ny, cells = cols * P.nz;
  const int nDir = P.nDir;
  const bool useRR = P.opt.useRussianRouletteForIntensity != 0;
  const bool thermal = P.opt.LW_flag > 0.0f;
  Grid G;
  G.sx = G.sy = G.sz = nullptr;
  Tally T;
  T.cols = cols;
  T.sFlux = plan.fluxOff >= 0 ? smem + plan.fluxOff : nullptr;
  T.sVol = plan.volOff >= 0 ? smem + plan.volOff : nullptr;
  T.sInt = plan.intOff >= 0 ? smem + plan.intOff : nullptr;
  if (T.sFlux) for (int i = threadIdx.x; i < 2 * cols; i += THREADS) T.sFlux[i] = 0.0f;
  if (T.sVol) for (int i = threadIdx.x; i < cells; i += THREADS) T.sVol[i] = 0.0f;
  if (T.sInt) for (int i = threadIdx.x; i < cols * nDir; i += THREADS) T.sInt[i] = 0.0f;
  if (threadIdx.x < 6) sCnt[threadIdx.x] = 0u;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  const int rqWords = RQ_T + 3 * nDir;
  float *pool = smem + plan.poolOff + (threadIdx.x >> 5) * (PLE_WORDS * PLE_SLOTS + rqWords * PLE_REQS);
  float *ring = pool + PLE_WORDS * PLE_SLOTS;                 // word-major: ring[word * PLE_REQS + slot]
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t divDir = (65536u + (uint32_t)nDir - 1u) / (uint32_t)nDir;     // x / nDir = (x * divDir) >> 16 for x < 4096

  unsigned crossings = 0u, scatters = 0u, leRays = 0u, leCrossings = 0u;
  // the lane's job: a ray, its accumulated optical depth and target, and five words of payload
  //   photon leg: weight, uNext, Philox counter (c0, c1, blk), order
  //   view ray  : weight, npf, tauFree, uTest, (dir | comps << 8)
  Ray r;
  r.ox = r.oy = r.oz = 0.0f; r.dx = r.dy = 0.0f; r.dz = 1.0f; r.rx = r.ry = r.rz = FLT_MAX;
  r.t = 0.0f; r.tx = r.ty = r.tz = FLT_MAX; r.ix = r.iy = r.iz = 0;
  float ext = 0.0f, tgt = 0.0f, jw = 0.0f, ja = 0.0f;
  float vcur = 0.0f;                                          // <= -1: the ray's cell lies -vcur cells deep in vacuum (march_leap)
  uint32_t jb = 0u, jc = 0u, jd = 0u;
  int order = 0;
  int job = JOB_NONE;
  bool more = true;

  int nR = 0, nE = PLE_SLOTS;                                 // photon pool: starts as 64 dead photons (mcb_pool.cu)
  pool[PW_IZK * PLE_SLOTS + lane] = __int_as_float(ST_DEAD << 28);
  pool[PW_IZK * PLE_SLOTS + lane + 32] = __int_as_float(ST_DEAD << 28);
  int hReq = 0, hDir = 0, avail = 0, reqCount = 0;            // request ring: head request / direction, tasks waiting, requests held
  __syncwarp();
  const int maxEvents = thermal ? 16 : 32;                    // a thermal event may post twice: its scattering and the next birth

  for (;;) {
    // =========================== event phase ===========================
    for (;;) {
      const int jobs = __popc(__ballot_sync(FULL, job != JOB_NONE));
      const bool room = reqCount == 0;                       // every waiting view ray has been taken
      const bool go = (nE >= maxEvents && room) || (nE > 0 && nR == 0 && room && nE >= jobs);
      if (!go) break;
      const int cnt = min(maxEvents, nE);
      const int slot = PLE_SLOTS - nE + lane;
      nE -= cnt;
      int state = ST_DONE;
      float px = 0.0f, py = 0.0f, pz = 0.0f, dx = 0.0f, dy = 0.0f, dz = 1.0f, ew = 0.0f, eNext = 0.0f, eTau = 0.0f;
      int ix = 0, iy = 0, iz = 0, eOrder = 0;
      Rng rng;
      rng.c0 = rng.c1 = rng.blk = 0u;
      if (lane < cnt) {
        const int izk = __float_as_int(pool[PW_IZK * PLE_SLOTS + slot]);
        state = (int)((uint32_t)izk >> 28);
        eOrder = izk & 0x0fffffff;
        if (state != ST_DEAD) {
          const int raw = __float_as_int(pool[PW_IXY * PLE_SLOTS + slot]);
          const float t = pool[PW_TAU * PLE_SLOTS + slot];
          px = pool[PW_PX * PLE_SLOTS + slot]; py = pool[PW_PY * PLE_SLOTS + slot]; pz = pool[PW_PZ * PLE_SLOTS + slot];
          dx = pool[PW_DX * PLE_SLOTS + slot]; dy = pool[PW_DY * PLE_SLOTS + slot]; dz = pool[PW_DZ * PLE_SLOTS + slot];
          ew = pool[PW_W * PLE_SLOTS + slot]; eNext = pool[PW_UNEXT * PLE_SLOTS + slot];
          rng.c0 = __float_as_uint(pool[PW_C0 * PLE_SLOTS + slot]); rng.c1 = __float_as_uint(pool[PW_C1 * PLE_SLOTS + slot]);
          rng.blk = __float_as_uint(pool[PW_BLK * PLE_SLOTS + slot]);
          px = fmaf(t, dx, px); py = fmaf(t, dy, py); pz = fmaf(t, dz, pz);
          px -= P.fLx * floorf((px - P.fx0) * P.finvLx);
          py -= P.fLy * floorf((py - P.fy0) * P.finvLy);
          if (state == ST_SCATTER) {
            cell_decode<WIDE, BRICK>(P, raw, ix, iy, iz);
          } else {
            ix = min(max((int)((px - P.fx0) * P.finvhx), 0), P.nx - 1);
            iy = min(max((int)((py - P.fy0) * P.finvhy), 0), P.ny - 1);
            iz = state == ST_TOP ? P.nz - 1 : 0;
          }
        }
      }
      // ---- tallies, absorption, roulette; what the event asks of the view rays (INT:776-800, 676-702) ----
      int comp = 1, pidx = 1;
      int rqComp = -2, rqTally = 0, rqPidx = 0;              // -2: no request
      float rqW = 0.0f;
      if (state == ST_TOP) {
        add_flux(P, T, 0, ix + P.nx * iy, ew);
        state = ST_DEAD;
      } else if (state == ST_SURFACE) {
        add_flux(P, T, 1, ix + P.nx * iy, ew);
        atomicAdd(&sCnt[0], 1u);
        eOrder++;
        ew = (float)((double)ew * P.albedo);
        if (ew <= TINY32) {
          atomicAdd(&sCnt[1], 1u);
          state = ST_DEAD;
        } else {
          pz = P.fz0; iz = 0;
          rqComp = 0; rqTally = 0; rqPidx = 0; rqW = ew;
        }
      } else if (state == ST_SCATTER) {
        eOrder++;
        scatters++;
        const int cell = ix + P.nx * (iy + P.ny * iz);
        float lo = 0.0f, hi = 1.0f, ssa;
        {
          const uint32_t *R = P.rec + ((size_t)MCB_CHECK_INDEX(P, cell, cells) << P.recShift);
          if (P.nc == 1) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(R));
            ssa = __uint_as_float(v.x); pidx = (int)(v.y & 0xffffu);
          } else if (P.nc == 2) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(R));
            const float cc = __uint_as_float(v.x);
            if (eNext >= cc) { comp = 2; lo = cc; } else { hi = cc; }
            ssa = __uint_as_float(comp == 1 ? v.y : v.z);
            pidx = (int)(comp == 1 ? (v.w & 0xffffu) : (v.w >> 16));
          } else {
            for (int c = 1; c < P.nc; ++c) {
              const float cc = __uint_as_float(__ldg(R + (c - 1)));
              if (eNext >= cc) { comp = c + 1; lo = cc; } else { hi = fminf(hi, cc); }
            }
            ssa = __uint_as_float(__ldg(R + (P.nc - 1) + (comp - 1)));
            const uint32_t pw = __ldg(R + (2 * P.nc - 1) + ((comp - 1) >> 1));
            pidx = (int)(((comp - 1) & 1) ? (pw >> 16) : (pw & 0xffffu));
          }
        }
        pidx = max(pidx, 1);
        if (ssa < 1.0f) {
          add_vol(P, T, cell, ew * (1.0f - ssa));
          ew *= ssa;
        }
        rqComp = comp; rqTally = comp; rqPidx = pidx; rqW = ew;                          // before the roulette, INT:776
        if (P.opt.useRussianRoulette && ew < P.opt.russianRouletteW * 0.5f) {
          const float uRR = P.nc > 1 ? __fdividef(eNext - lo, fmaxf(hi - lo, TINY32)) : eNext;
          if (uRR >= ew / P.opt.russianRouletteW) ew = 0.0f; else ew = P.opt.russianRouletteW;
        }
        if (ew <= TINY32) {
          atomicAdd(&sCnt[2], 1u);
          state = ST_DEAD;
        }
      }
      // ---- post the requests of this phase: position, weight, cell, component; then, per direction, the ray's numbers ----
      for (int round = 0; round < 2; ++round) {
        if (round == 1) {
          // finished photons are replaced from the global counter (getNextPhoton, ILL:561-590); thermal births post too
          const unsigned dead = __ballot_sync(FULL, state == ST_DEAD);
          if (dead) {
            if (more) {
              unsigned long long base = 0;
              if (lane == 0) base = atomicAdd(workCounter, (unsigned long long)__popc(dead));
              base = __shfl_sync(FULL, base, 0);
              if (state == ST_DEAD) {
                const unsigned long long p = base + (unsigned long long)__popc(dead & below);
                if (p < (unsigned long long)nPhotons) { rng.init(firstPhotonId + p); state = ST_BORN; }
                else state = ST_DONE;
              }
              if (base + (unsigned long long)__popc(dead) >= (unsigned long long)nPhotons) more = false;
            } else if (state == ST_DEAD) {
              state = ST_DONE;
            }
          }
          rqComp = -2;
          if (state == ST_BORN) {
            const float4 u = rng.block(k0, k1);
            float x01, y01, z01;
            int bi = -1, bj = 0, bk = 0;
            ew = 1.0f; eOrder = 0;
            if (P.source == 0) {
              x01 = u.x; y01 = u.y; z01 = 1.0f - FLT_EPSILON;
              dx = P.solarDir[0]; dy = P.solarDir[1]; dz = P.solarDir[2];
            } else {
              const float4 v = rng.block(k0, k1);
              float mu, phi;
              if ((double)u.x > P.fracAtmsPower) {
                x01 = u.y; y01 = v.w;
                mu = sqrtf(fmaxf(v.x, 1.0e-30f));
                phi = v.y * 2.0f * PI32;
                z01 = 0.0f;
              } else {
                const float q = u.y;
                const int ik = cdf_search(P.colCDF + (P.ny - 1), P.nz, (long long)P.ny, q);
                const int ij = cdf_search(P.colCDF + (size_t)P.ny * (size_t)(ik - 1), P.ny, 1, q);
                const double *voxBase = P.voxelCDF + (size_t)P.nx * ((size_t)(ij - 1) + (size_t)P.ny * (size_t)(ik - 1));
                const int ii = cdf_search(voxBase, P.nx, 1, q);
                const float4 v2 = rng.block(k0, k1);
                z01 = ((float)(ik - 1) + fminf(fmaxf(v2.x, 1e-6f), 1.0f - 1e-6f)) / (float)P.nz;
                x01 = ((float)(ii - 1) + fminf(v2.y, 1.0f - 1e-6f)) / (float)P.nx;
                y01 = ((float)(ij - 1) + fminf(v2.z, 1.0f - 1e-6f)) / (float)P.ny;
                bi = ii - 1; bj = ij - 1; bk = ik - 1;
                mu = 1.0f - 2.0f * v.x;
                if (!(fabsf(mu) > 2.0f * TINY32)) mu = 1.0e-30f;
                phi = v.y * 2.0f * PI32;
              }
              dir_from(mu, phi, dx, dy, dz);
            }
            px = fmaf(x01, P.fLx, P.fx0); py = fmaf(y01, P.fLy, P.fy0); pz = fmaf(z01, P.fLz, P.fz0);
            ix = min((int)(x01 * (float)P.nx), P.nx - 1);
            iy = min((int)(y01 * (float)P.ny), P.ny - 1);
            iz = min((int)(z01 * (float)P.nz), P.nz - 1);
            if (bi >= 0) { ix = bi; iy = bj; iz = bk; }
            eTau = -__logf(fmaxf(TINY32, u.z));
            eNext = u.w;
            if (thermal) {                                                           // INT:504-542
              if (pz > 0.0f) add_vol(P, T, ix + P.nx * (iy + P.ny * iz), -1.0f);
              rqComp = pz == 0.0f ? 0 : -1; rqTally = 0; rqPidx = 0; rqW = ew;
            }
            state = ST_MARCH;                                                        // the leg is complete: nothing more to draw
          }
        }
        const unsigned posting = __ballot_sync(FULL, rqComp != -2);
        if (posting) {
          if (rqComp != -2) {
            const int s = (hReq + reqCount + __popc(posting & below)) & (PLE_REQS - 1);
            ring[RQ_PX * PLE_REQS + s] = px; ring[RQ_PY * PLE_REQS + s] = py; ring[RQ_PZ * PLE_REQS + s] = pz;
            ring[RQ_W * PLE_REQS + s] = rqW;
            ring[RQ_IXY * PLE_REQS + s] = __int_as_float(ix | (iy << 16));
            ring[RQ_IZ * PLE_REQS + s] = __int_as_float(iz);
            ring[RQ_COMP * PLE_REQS + s] = __int_as_float(((rqComp + 1) & 0xff) | ((rqTally & 0xff) << 8));
            // per direction: normalised phase function value; with Russian roulette the free path and the test variate,
            // one Philox block per direction pair -- the blocks le_post reserves: blk + (dir >> 1)
            for (int d = 0; d < nDir; d += 2) {
              float4 u = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
              if (useRR) { Rng q; q.c0 = rng.c0; q.c1 = rng.c1; q.blk = rng.blk + (uint32_t)(d >> 1); u = q.block(k0, k1); }
              ring[(RQ_T + 3 * d) * PLE_REQS + s] = view_phase_value(P, rqComp, rqPidx, eOrder, d, dx, dy, dz);
              ring[(RQ_T + 3 * d + 1) * PLE_REQS + s] = -__logf(fmaxf(TINY32, u.x));
              ring[(RQ_T + 3 * d + 2) * PLE_REQS + s] = u.y;
              if (d + 1 < nDir) {
                ring[(RQ_T + 3 * d + 3) * PLE_REQS + s] = view_phase_value(P, rqComp, rqPidx, eOrder, d + 1, dx, dy, dz);
                ring[(RQ_T + 3 * d + 4) * PLE_REQS + s] = -__logf(fmaxf(TINY32, u.z));
                ring[(RQ_T + 3 * d + 5) * PLE_REQS + s] = u.w;
              }
            }
            if (useRR) rng.blk += (uint32_t)((nDir + 1) >> 1);
          }
          const int n = __popc(posting);
          reqCount += n; avail += n * nDir;
        }
      }
      // ---- one Philox block per surviving event: new direction, next optical depth, next pick ----
      if (state == ST_SURFACE || state == ST_SCATTER) {
        const float4 u = rng.block(k0, k1);
        if (state == ST_SURFACE) {
          const float mu = sqrtf(fmaxf(u.x, 1.0e-30f));
          dir_from(mu, 2.0f * PI32 * u.y, dx, dy, dz);
        } else {
          const int c = comp - 1;
          const int nS = P.invS[c];
          const float *tab = P.inv[c] + (size_t)MCB_CHECK_INDEX(P, pidx - 1, P.invE[c]) * nS;
          const float rn = u.x;
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
          float AX, AY;
          __sincosf(2.0f * PI32 * u.y, &AY, &AX);
          AX *= sinT; AY *= sinT;
          const float Bq = dx * AX - dy * AY;
          const float D = cosT - __fdividef(Bq, 1.0f + fabsf(dz));
          const float ndx = dx * D + AX, ndy = dy * D - AY;
          const float ndz = dz * cosT - copysignf(fabsf(Bq), dz * Bq);
          dx = ndx; dy = ndy; dz = ndz;
        }
        eTau = -__logf(fmaxf(TINY32, u.z));
        eNext = u.w;
        state = ST_MARCH;
      }
      // ---- push the new legs, completely set up, onto READY ----
      const bool alive = state == ST_MARCH;
      const unsigned m = __ballot_sync(FULL, alive);
      if (alive) {
        const int s = nR + __popc(m & below);
        Ray q;
        q.ox = px; q.oy = py; q.oz = pz; q.dx = dx; q.dy = dy; q.dz = dz; q.ix = ix; q.iy = iy; q.iz = iz;
        ray_start<true>(q, P, G);
        pool[PW_TX * PLE_SLOTS + s] = q.tx; pool[PW_TY * PLE_SLOTS + s] = q.ty; pool[PW_TZ * PLE_SLOTS + s] = q.tz;
        pool[PW_PX * PLE_SLOTS + s] = px; pool[PW_PY * PLE_SLOTS + s] = py; pool[PW_PZ * PLE_SLOTS + s] = pz;
        pool[PW_DX * PLE_SLOTS + s] = dx; pool[PW_DY * PLE_SLOTS + s] = dy; pool[PW_DZ * PLE_SLOTS + s] = dz;
        pool[PW_W * PLE_SLOTS + s] = ew; pool[PW_TAU * PLE_SLOTS + s] = eTau; pool[PW_UNEXT * PLE_SLOTS + s] = eNext;
        pool[PW_C0 * PLE_SLOTS + s] = __uint_as_float(rng.c0); pool[PW_C1 * PLE_SLOTS + s] = __uint_as_float(rng.c1);
        pool[PW_BLK * PLE_SLOTS + s] = __uint_as_float(rng.blk);
        pool[PW_IXY * PLE_SLOTS + s] = __int_as_float(ix | (iy << 16));
        pool[PW_IZK * PLE_SLOTS + s] = __int_as_float(iz | (min(eOrder, 65535) << 16));
      }
      nR += __popc(m);
      __syncwarp();
    }

    // =========================== lanes without a job: a view ray first, else a photon leg ===========================
    {
      const unsigned idle = __ballot_sync(FULL, job == JOB_NONE);
      if (idle) {
        const int rank = __popc(idle & below);
        const int nIdle = __popc(idle);
        const int takeRays = min(nIdle, avail);
        if (job == JOB_NONE && rank < takeRays) {                     // the rank-th waiting (request, direction)
          const uint32_t x = (uint32_t)(hDir + rank);
          const uint32_t dq = (x * divDir) >> 16;
          const int dir = (int)(x - dq * (uint32_t)nDir);
          const int s = (hReq + (int)dq) & (PLE_REQS - 1);
          const int ixy = __float_as_int(ring[RQ_IXY * PLE_REQS + s]);
          r.ix = ixy & 0xffff; r.iy = (int)((uint32_t)ixy >> 16); r.iz = __float_as_int(ring[RQ_IZ * PLE_REQS + s]);
          r.ox = ring[RQ_PX * PLE_REQS + s]; r.oy = ring[RQ_PY * PLE_REQS + s]; r.oz = ring[RQ_PZ * PLE_REQS + s];
          r.dx = P.viewDir[3 * dir]; r.dy = P.viewDir[3 * dir + 1]; r.dz = P.viewDir[3 * dir + 2];
          ray_start<true>(r, P, G);
          ext = 0.0f; vcur = 0.0f;
          jw = ring[RQ_W * PLE_REQS + s];
          ja = ring[(RQ_T + 3 * dir) * PLE_REQS + s];                                   // npf
          jb = __float_as_uint(ring[(RQ_T + 3 * dir + 1) * PLE_REQS + s]);              // tauFree
          jc = __float_as_uint(ring[(RQ_T + 3 * dir + 2) * PLE_REQS + s]);              // uTest
          jd = (uint32_t)dir | ((uint32_t)__float_as_int(ring[RQ_COMP * PLE_REQS + s]) << 8);
          leRays++;
          if (!useRR) { tgt = FLT_MAX; job = JOB_PLAIN; }                                // INT:1729-1752
          else if (PI32 * ja <= P.opt.zetaMin) { tgt = __uint_as_float(jb); job = JOB_E13; }   // Iwabuchi (2006) Eq 13
          else { tgt = -__logf(P.opt.zetaMin / fmaxf(TINY32, PI32 * ja)); job = JOB_E14A; }    // Eq 14
        }
        if (takeRays > 0) {
          const uint32_t total = (uint32_t)(hDir + takeRays);
          const uint32_t dq = (total * divDir) >> 16;
          hReq = (hReq + (int)dq) & (PLE_REQS - 1); hDir = (int)(total - dq * (uint32_t)nDir);
          reqCount -= (int)dq; avail -= takeRays;
        }
        const int legs = min(nIdle - takeRays, nR);
        if (job == JOB_NONE && rank >= takeRays && rank - takeRays < legs) {
          const int s = nR - 1 - (rank - takeRays);
          r.ox = pool[PW_PX * PLE_SLOTS + s]; r.oy = pool[PW_PY * PLE_SLOTS + s]; r.oz = pool[PW_PZ * PLE_SLOTS + s];
          r.dx = pool[PW_DX * PLE_SLOTS + s]; r.dy = pool[PW_DY * PLE_SLOTS + s]; r.dz = pool[PW_DZ * PLE_SLOTS + s];
          jw = pool[PW_W * PLE_SLOTS + s]; tgt = pool[PW_TAU * PLE_SLOTS + s]; ja = pool[PW_UNEXT * PLE_SLOTS + s];
          jb = __float_as_uint(pool[PW_C0 * PLE_SLOTS + s]); jc = __float_as_uint(pool[PW_C1 * PLE_SLOTS + s]);
          jd = __float_as_uint(pool[PW_BLK * PLE_SLOTS + s]);
          const int ixy = __float_as_int(pool[PW_IXY * PLE_SLOTS + s]), izo = __float_as_int(pool[PW_IZK * PLE_SLOTS + s]);
          r.ix = ixy & 0xffff; r.iy = (int)((uint32_t)ixy >> 16); r.iz = izo & 0xffff; order = (int)((uint32_t)izo >> 16);
          r.tx = pool[PW_TX * PLE_SLOTS + s]; r.ty = pool[PW_TY * PLE_SLOTS + s]; r.tz = pool[PW_TZ * PLE_SLOTS + s];
          r.rx = safe_rcp(r.dx); r.ry = safe_rcp(r.dy); r.rz = safe_rcp(r.dz);
          r.t = 0.0f; ext = 0.0f; vcur = 0.0f;
          job = JOB_PHOTON;
        }
        nR -= legs;
        __syncwarp();
      }
    }
    if (!__any_sync(FULL, job != JOB_NONE)) break;           // no job, nothing ready, no ray waiting, no event: done

    // =========================== march: one burst for every lane ===========================
    int ev = MARCH_ON;
    unsigned crossed = 0u;
    if (LEAP) {                                                // first through vacuum in one leap where the cell is known to lie
      int D = job != JOB_NONE ? leap_distance(r, P, vcur, leapBelow) : 0;            // deep in it
      if (leapLanes > 1 && __popc(__ballot_sync(FULL, D > 0)) < leapLanes) D = 0;   // too few lanes to pay for the divergence
      if (D) ev = march_leap<MASK>(r, P, D, crossed, ext, tgt, &sCnt[4]);
    }
    if (job != JOB_NONE && ev == MARCH_ON)
      ev = march_burst<true, WIDE, BURST, MASK, BRICK, true, MCB_PLE_SPLIT, LEAP>(r, P, G, ext, tgt, crossed, LEAP ? &vcur : nullptr);
    if (job == JOB_PHOTON) crossings += crossed; else leCrossings += crossed;

    // =========================== rays that ended ===========================
    if (job > JOB_PHOTON && ev != MARCH_ON) {                  // a view ray: tally, or go on with its second leg
      float contribution = 0.0f;
      bool finished = true;
      const float npf = ja, tauFree = __uint_as_float(jb), uTest = __uint_as_float(jc);
      if (job == JOB_PLAIN) {
        contribution = jw * npf * __expf(-ext);
      } else if (job == JOB_E13) {
        contribution = (ev == MARCH_TOP && uTest <= PI32 * npf / P.opt.zetaMin) ? jw * P.opt.zetaMin / PI32 : 0.0f;
      } else if (job == JOB_E14A) {
        if (ev == MARCH_TOP) {
          contribution = jw * npf * __expf(-ext);
        } else if (ev == MARCH_HIT) {                          // continue from where the first trace stopped (INT:1793-1795)
          float qx, qy, qz;
          const int raw = r.ix;
          ray_position(r, P, qx, qy, qz);
          cell_decode<WIDE, BRICK>(P, raw, r.ix, r.iy, r.iz);
          r.ox = qx; r.oy = qy; r.oz = qz;
          ray_start<true>(r, P, G);
          ext = 0.0f; vcur = 0.0f; tgt = tauFree; job = JOB_E14B;
          finished = false;
        }
      } else {                                                 // JOB_E14B
        contribution = ev == MARCH_TOP ? jw * P.opt.zetaMin / PI32 : 0.0f;
      }
      if (finished) {
        const int dir = (int)(jd & 0xffu), comps = (int)(jd >> 8);
        const int component = (comps & 0xff) - 1, tallyComponent = (comps >> 8) & 0xff;
        if (P.opt.limitIntensityContributions && contribution > P.opt.maxIntensityContribution) {   // INT:1815-1826
          const int cslot = component < 0 ? 0 : component;
          atomicAdd(&P.tally[P.offExcess + dir + (long long)P.nDir * cslot], (double)(contribution - P.opt.maxIntensityContribution));
          contribution = P.opt.maxIntensityContribution;
        }
        if (contribution != 0.0f) {                            // the pixel the ray leaves through
          float qx, qy, qz;
          ray_position(r, P, qx, qy, qz);
          const int cx = min(max((int)((qx - P.fx0) * P.finvhx), 0), P.nx - 1);
          const int cy = min(max((int)((qy - P.fy0) * P.finvhy), 0), P.ny - 1);
          add_intensity(P, T, dir, cx + P.nx * cy, tallyComponent, contribution);
        }
        job = JOB_NONE;
      }
    }
    {                                                          // photons that reached an event go onto EVENT, raw
      const bool arrived = job == JOB_PHOTON && ev != MARCH_ON;
      const unsigned hit = __ballot_sync(FULL, arrived);
      if (hit) {
        if (arrived) {
          const int s = PLE_SLOTS - nE - 1 - __popc(hit & below);
          pool[PW_PX * PLE_SLOTS + s] = r.ox; pool[PW_PY * PLE_SLOTS + s] = r.oy; pool[PW_PZ * PLE_SLOTS + s] = r.oz;
          pool[PW_DX * PLE_SLOTS + s] = r.dx; pool[PW_DY * PLE_SLOTS + s] = r.dy; pool[PW_DZ * PLE_SLOTS + s] = r.dz;
          pool[PW_W * PLE_SLOTS + s] = jw; pool[PW_TAU * PLE_SLOTS + s] = r.t; pool[PW_UNEXT * PLE_SLOTS + s] = ja;
          pool[PW_C0 * PLE_SLOTS + s] = __uint_as_float(jb); pool[PW_C1 * PLE_SLOTS + s] = __uint_as_float(jc);
          pool[PW_BLK * PLE_SLOTS + s] = __uint_as_float(jd);
          pool[PW_IXY * PLE_SLOTS + s] = __int_as_float(r.ix);
          pool[PW_IZK * PLE_SLOTS + s] = __int_as_float((ev << 28) | order);
          job = JOB_NONE;
        }
        nE += __popc(hit);
        __syncwarp();
      }
    }
  }

  // ---- flush: counters, privatised tallies ----
  {
    unsigned long long v[4] = {crossings, scatters, leRays, leCrossings};
    const int slotOf[4] = {CNT_CROSSINGS, CNT_SCATTERS, CNT_LE_RAYS, CNT_LE_CROSSINGS};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(FULL, v[i], o);
      if (lane == 0 && v[i]) atomicAdd(&P.counters[slotOf[i]], v[i]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sCnt[0]) atomicAdd(&P.counters[CNT_SURFACE], (unsigned long long)sCnt[0]);
    if (sCnt[1]) atomicAdd(&P.counters[CNT_SURFACE_KILLS], (unsigned long long)sCnt[1]);
    if (sCnt[2]) atomicAdd(&P.counters[CNT_RR_KILLS], (unsigned long long)sCnt[2]);
    if (sCnt[4]) atomicAdd(&P.counters[CNT_LEAPS], (unsigned long long)sCnt[4]);
    if (sCnt[5]) atomicAdd(&P.counters[CNT_LEAP_CELLS], (unsigned long long)sCnt[5]);
  }
  if (T.sFlux)
    for (int i = threadIdx.x; i < 2 * cols; i += THREADS) {
      const float v = T.sFlux[i];
      if (v != 0.0f) atomicAdd(&P.tally[P.offFluxUp + i], (double)v);
    }
  if (T.sVol)
    for (int i = threadIdx.x; i < cells; i += THREADS) {
      const float v = T.sVol[i];
      if (v != 0.0f) atomicAdd(&P.tally[P.offVolAbs + i], (double)v);
    }
  if (T.sInt)
    for (int i = threadIdx.x; i < cols * nDir; i += THREADS) {
      const float v = T.sInt[i];
      if (v != 0.0f) atomicAdd(&P.tally[P.offInt + i], (double)v);
    }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&P.tally[P.offPhotons], (double)nPhotons);
    atomicAdd(&P.counters[CNT_PHOTONS], (unsigned long long)nPhotons);
  }
}

}  // namespace mcbpoolle

template <int MINBLOCKS, bool MASK, bool BRICK, bool LEAP, bool WIDE = true>
static void launch_pool_le(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId, int numSMs,
                           unsigned long long *workCounter, cudaStream_t stream) {
  constexpr int THREADS = 128;
  auto kernel = mcbpoolle::pool_le_kernel<THREADS, MINBLOCKS, MCB_PLE_BURST, MASK, BRICK, LEAP, WIDE>;
  const int cols = P.nx * P.ny, cells = cols * P.nz;
  mcbfast::SmemPlan plan{-1, -1, -1, -1, -1, 0, 0, 0};
  int off = 0;
  plan.poolOff = off; off += (THREADS / 32) * (PLE_WORDS * PLE_SLOTS + (mcbpoolle::RQ_T + 3 * P.nDir) * PLE_REQS);
  const int budgetFloats = 9 * 1024;             // small grids are atomic hot spots: privatise (as mcb_fast.cu does)
  int used = 0;
  if (cols <= 1024 && 2 * cols <= budgetFloats) { plan.fluxOff = off; off += 2 * cols; used += 2 * cols; }
  if (cells <= 8192 && used + cells <= budgetFloats) { plan.volOff = off; off += cells; used += cells; }
  if (cols * P.nDir <= 2048 && used + cols * P.nDir <= budgetFloats) { plan.intOff = off; off += cols * P.nDir; }
  plan.totalFloats = off;
  const size_t smem = sizeof(float) * (size_t)off;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int blocksPerSM = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kernel, THREADS, smem) != cudaSuccess || blocksPerSM < 1)
    blocksPerSM = 1;
  if (P.opt.tuneBlocksPerSM > 0 && P.opt.tuneBlocksPerSM < blocksPerSM) blocksPerSM = P.opt.tuneBlocksPerSM;
  const long long perBlock = (long long)(THREADS / 32) * PLE_SLOTS;
  const long long want = (nPhotons + perBlock - 1) / perBlock;
  const long long cap = (long long)numSMs * blocksPerSM;
  const int blocks = (int)(want < cap ? want : cap);
  const float leapBelow = P.opt.tuneLeap < 0 ? -FLT_MAX : -(float)(P.opt.tuneLeap >= 2 ? P.opt.tuneLeap : MCB_LEAP_MIN);
  kernel<<<blocks, THREADS, smem, stream>>>(P, nPhotons, seed, firstPhotonId, workCounter, plan, leapBelow, P.opt.tuneLeapLanes > 0 ? P.opt.tuneLeapLanes : MCB_LEAP_LANES);
}

// runs with view directions on uniform grids at least a ghost shell wide, up to MCB_POOL_LE_MAXDIR directions
bool mcb_pool_le_covers(const DevDomain &P) {
  const bool wide = P.nx >= MCB_GHOST && P.ny >= MCB_GHOST;
  return P.uniform && P.nDir > 0 && P.nDir <= MCB_POOL_LE_MAXDIR && P.nx <= 65535 && P.ny <= 65535 && P.nz <= 65535 &&
         (wide || (P.lin.mask == nullptr && P.opt.tuneLayout != MCB_LAYOUT_BRICKS));       // narrow grids: plain x-fastest field
}

// ... and is the default where it measured faster: grids at least a ghost shell wide (C3 + 5 views: 1.11e8 against the task
// queue's 8.5e7).  Narrow grids (the 32 x 1 x 32 step cloud): MCB_PLE_NARROW_DEFAULT.
#ifndef MCB_PLE_NARROW_DEFAULT
#define MCB_PLE_NARROW_DEFAULT 1            // C2 step cloud + 5 views, one B200, r02: 1.67e8 against the task queue's 1.59e8
#endif
bool mcb_pool_le_preferred(const DevDomain &P) {
  return (P.nx >= MCB_GHOST && P.ny >= MCB_GHOST) || MCB_PLE_NARROW_DEFAULT != 0;
}

// which layout of the extinction field this kernel reads (mcb_api.cu packs that one)
bool mcb_pool_le_reads_bricks(const DevDomain &P) { return P.opt.tuneLayout == MCB_LAYOUT_BRICKS; }

void mcb_launch_pool_le_batch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
                              int numSMs, unsigned long long *workCounter, cudaStream_t stream) {
  if (nPhotons <= 0) return;
  const bool mask = P.lin.mask != nullptr, brick = mcb_pool_le_reads_bricks(P);
  const int occ = P.opt.tuneBlocksPerSM ? P.opt.tuneBlocksPerSM : 5;
#define MCB_PLE_ARGS (P, nPhotons, seed, firstPhotonId, numSMs, workCounter, stream)
  if (!(P.nx >= MCB_GHOST && P.ny >= MCB_GHOST)) {          // a period shorter than the ghost shell
    if (occ >= 6) launch_pool_le<6, false, false, false, false> MCB_PLE_ARGS;
    else if (occ == 5) launch_pool_le<5, false, false, false, false> MCB_PLE_ARGS;
    else launch_pool_le<4, false, false, false, false> MCB_PLE_ARGS;
    return;
  }
#define MCB_PLE_GO2(OCC, LEAP) \
  do { if (mask) { if (brick) launch_pool_le<OCC, true, true, LEAP> MCB_PLE_ARGS; else launch_pool_le<OCC, true, false, LEAP> MCB_PLE_ARGS; } \
       else { if (brick) launch_pool_le<OCC, false, true, LEAP> MCB_PLE_ARGS; else launch_pool_le<OCC, false, false, LEAP> MCB_PLE_ARGS; } } while (0)
#define MCB_PLE_GO(OCC) do { if (P.leap) MCB_PLE_GO2(OCC, true); else MCB_PLE_GO2(OCC, false); } while (0)
  if (occ >= 6) MCB_PLE_GO(6); else if (occ == 5) MCB_PLE_GO(5); else MCB_PLE_GO(4);
#undef MCB_PLE_GO
#undef MCB_PLE_GO2
#undef MCB_PLE_ARGS
}
