// mcb_pool.cu -- the photon-pool flux kernel (sm_100a): the same physics and event sequence as mcb_fast.cu's
// park/regroup megakernel (computeRT, INT:393-841), organised so that no lane waits for another.
//
// The park/regroup kernel keeps one photon per lane: a lane that reaches an event idles until half the warp has
// arrived, and the event phase then runs with the other half idle (ncu, round 1: 20 of 32 lanes active per
// instruction -- march code 26, event code 16).  Here every warp owns a POOL of 64 photons: 32 are in the lanes'
// registers, marching; the others wait in shared memory as 17-word records in one of two stacks that grow towards
// each other inside one array of 64 slots,
//     READY  [0, nR)        legs that have been set up (origin, direction, cell, target optical depth) and wait for a lane
//     EVENT  [64 - nE, 64)  photons that reached an event (scattering, surface, top) and wait for the event phase
// and the warp alternates between
//     march  -- one burst for all 32 lanes; a lane whose photon reaches an event pushes it onto EVENT and pops a
//               READY leg in its place, so all 32 lanes march in (almost) every burst;
//     events -- as soon as 32 events wait: one event phase with all 32 lanes (tallies, absorption, roulette, rebirth
//               from the global photon counter, one Philox block, new direction), each result pushed onto READY.
// Both stacks are addressed by rank among the pushing / popping lanes (ballot + popc), so the 32 lanes always touch
// 32 consecutive slots: conflict-free shared-memory access without any per-slot bookkeeping, and nR, nE are plain
// warp-uniform registers.  Photons in the pool + in registers never exceed 64 (a birth only ever replaces a death),
// so the stacks cannot collide.
//
// A record is the universal hand-over format: position, direction, weight, target optical depth, the uniform that
// picks the component at the next scattering, the photon's Philox counter, its cell and the kind of event.
// Everything that is work per EVENT rather than per cell runs in the event phase, where all 32 lanes have an event:
// a marching lane pushes its ray RAW (leg origin, distance along the leg, padded address of the hit cell) and the
// event phase decodes cell and position; the event phase also sets the next leg up completely (first face
// distances), so a lane that pops a READY leg only loads it.  (First version: decode and leg set-up ran in the march
// loop with the ~12 lanes that had just arrived -- 200 warp-instructions per burst at 12 of 32 lanes, ncu r02.)
//
// Variants (template parameters, chosen by mcb_launch_pool_batch from what the staging found):
//   LEAP  -- the domain has vacuum (or, on bitmap-marched fields, layers that are clear throughout) worth leaping: a lane
//            whose last gather says that its cell lies D >= 2 cells deep in it starts its iteration with march_leap
//            (mcb_march.cuh) and then runs its burst from the landing cell;
//   MASK  -- fields too large for L2, marched through the occupancy bitmap;
//   CROP  -- fields too large for L2 whose cloud occupies a band of layers: the band as its own bricked, L2-resident field;
//            event records and the absorption tally of the cells inside the per-column ranges in compact arrays
//            (DevDomain::colTab / recC / tallyC), the tally added into the dense one after the launch.
// Scope: flux / absorption runs (no view directions) on uniform grids at least a ghost shell wide -- C1, C3, C4, C5.
// Everything else stays on mcb_fast.cu (view directions: mcb_pool_le.cu).  Statistical parity with the reference arithmetic (criterion (b)) is tested
// like the park kernel's; with the same seed the two kernels trace the SAME photon histories (same Philox blocks in
// the same order per photon), so their tallies agree to summation order -- tests/test_gpu_pool.py.
#include "mcb_march.cuh"

#ifndef MCB_CROP_OCC
#define MCB_CROP_OCC 6            // layer-cropped variant: 80 registers (C5, r02: 6 CTAs/SM 4.86e8, 7 CTAs/SM with spills 4.30e8),
                                  // gathers in two halves as on the other L2-resident fields (4.86e8 vs 4.78e8)
#endif
#ifndef MCB_CROP_SPLIT
#define MCB_CROP_SPLIT true
#endif

namespace mcbpool {

using namespace mcbfast;

#define POOL_WORDS 17
#define POOL_SLOTS 64
// EVENT records: PX.. = leg origin, TAU = distance along the leg, IXY = padded address of the hit cell, IZK = kind << 28
// READY records: PX.. = leg origin, TAU = target optical depth, IXY / IZK = cell, TX.. = first face distances
enum { PW_PX = 0, PW_PY, PW_PZ, PW_DX, PW_DY, PW_DZ, PW_W, PW_TAU, PW_UNEXT, PW_C0, PW_C1, PW_BLK, PW_IXY, PW_IZK,
       PW_TX, PW_TY, PW_TZ };

template <int THREADS, int MINBLOCKS, int BURST, bool SPLIT, bool MASK, bool BRICK, bool LEAP, bool CROP>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
pool_kernel(const __grid_constant__ DevDomain P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
            unsigned long long *workCounter, const SmemPlan plan, const float leapBelow, const int leapLanes) {
  extern __shared__ float smem[];
  __shared__ unsigned sCnt[6];            // rare events: surface hits, surface kills, roulette kills
  const int cols = P.nx * P.ny, cells = cols * P.nz;
  Grid G;
  G.sx = G.sy = G.sz = nullptr;
  Tally T;
  T.cols = cols;
  T.sFlux = plan.fluxOff >= 0 ? smem + plan.fluxOff : nullptr;
  T.sVol = plan.volOff >= 0 ? smem + plan.volOff : nullptr;
  T.sInt = nullptr;
  if (T.sFlux) for (int i = threadIdx.x; i < 2 * cols; i += THREADS) T.sFlux[i] = 0.0f;
  if (T.sVol) for (int i = threadIdx.x; i < cells; i += THREADS) T.sVol[i] = 0.0f;
  if (threadIdx.x < 6) sCnt[threadIdx.x] = 0u;
  __syncthreads();

#ifdef MCB_POOL_SMEM_INV         // A/B build: the inverse table of a one-component, one-entry domain staged in shared memory
  const float *sInv = nullptr;
  if (plan.intOff >= 0) {
    float *t = smem + plan.intOff;
    for (int i = threadIdx.x; i < P.invS[0]; i += THREADS) t[i] = __ldg(P.inv[0] + i);
    __syncthreads();
    sInv = t;
  }
#endif
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  float *pool = smem + plan.poolOff + (threadIdx.x >> 5) * (POOL_WORDS * POOL_SLOTS);   // word-major: pool[word * 64 + slot]
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);

  unsigned crossings = 0u, scatters = 0u;
  // the lane's marching photon
  Ray r;
  r.ox = r.oy = r.oz = 0.0f; r.dx = r.dy = 0.0f; r.dz = 1.0f; r.rx = r.ry = r.rz = FLT_MAX;
  r.t = 0.0f; r.tx = r.ty = r.tz = FLT_MAX; r.ix = r.iy = r.iz = 0;
  float ext = 0.0f, tau = 0.0f, w = 0.0f, uNext = 0.0f;
  float vcur = 0.0f;                     // <= -1: the ray's cell is vacuum, and so is everything within -vcur - 1 cells of it
  uint32_t c0 = 0u, c1 = 0u, blk = 0u;
  bool have = false;
  bool more = true;                      // photons may remain in the global counter

  // the pool starts as 64 dead photons waiting for the event phase, which is where photons are born
  int nR = 0, nE = POOL_SLOTS;
  pool[PW_IZK * POOL_SLOTS + lane] = __int_as_float(ST_DEAD << 28);
  pool[PW_IZK * POOL_SLOTS + lane + 32] = __int_as_float(ST_DEAD << 28);
  __syncwarp();

  for (;;) {
    // =========================== event phase: 32 waiting events, 32 lanes ===========================
    while (nE >= 32 || (nE > 0 && nR == 0 && nE >= __popc(__ballot_sync(FULL, have)))) {
      const int cnt = min(32, nE);
      const int slot = POOL_SLOTS - nE + lane;               // the cnt most recent events
      nE -= cnt;
      int state = ST_DONE;
      float px = 0.0f, py = 0.0f, pz = 0.0f, dx = 0.0f, dy = 0.0f, dz = 1.0f, ew = 0.0f, eNext = 0.0f, eTau = 0.0f;
      int ix = 0, iy = 0, iz = 0;
      int raw = -1;
      Rng rng;
      rng.c0 = rng.c1 = rng.blk = 0u;
      if (lane < cnt) {
        const int izk = __float_as_int(pool[PW_IZK * POOL_SLOTS + slot]);
        state = (int)((uint32_t)izk >> 28);
        if (state != ST_DEAD) {
          raw = __float_as_int(pool[PW_IXY * POOL_SLOTS + slot]);
          const float t = pool[PW_TAU * POOL_SLOTS + slot];
          px = pool[PW_PX * POOL_SLOTS + slot]; py = pool[PW_PY * POOL_SLOTS + slot]; pz = pool[PW_PZ * POOL_SLOTS + slot];
          dx = pool[PW_DX * POOL_SLOTS + slot]; dy = pool[PW_DY * POOL_SLOTS + slot]; dz = pool[PW_DZ * POOL_SLOTS + slot];
          ew = pool[PW_W * POOL_SLOTS + slot]; eNext = pool[PW_UNEXT * POOL_SLOTS + slot];
          rng.c0 = __float_as_uint(pool[PW_C0 * POOL_SLOTS + slot]); rng.c1 = __float_as_uint(pool[PW_C1 * POOL_SLOTS + slot]);
          rng.blk = __float_as_uint(pool[PW_BLK * POOL_SLOTS + slot]);
          // where the leg ended, folded back into the periodic domain (ray_position), and in which cell
          px = fmaf(t, dx, px); py = fmaf(t, dy, py); pz = fmaf(t, dz, pz);
          px -= P.fLx * floorf((px - P.fx0) * P.finvLx);
          py -= P.fLy * floorf((py - P.fy0) * P.finvLy);
          if (state == ST_SCATTER && CROP) {
            // layer-cropped field: the address knows its cell; an event outside the cropped layers (raw = INT_MIN:
            // molecular scattering in a layer that is clear throughout) is located by its position
            if (raw != INT_MIN) {
              cell_decode<true, true>(P, P.crp, raw, ix, iy, iz);
              iz += GH + P.cropLo;                             // (the cropped field has no ghost layers)
            } else {
              ix = min(max((int)((px - P.fx0) * P.finvhx), 0), P.nx - 1);
              iy = min(max((int)((py - P.fy0) * P.finvhy), 0), P.ny - 1);
              iz = min(max((int)((pz - P.fz0) * P.finvhz), 0), P.nz - 1);
            }
          } else if (state == ST_SCATTER) {
            cell_decode<true, BRICK>(P, raw, ix, iy, iz);
          } else {                                         // left through the top / reached the surface: column of the exit point
            ix = min(max((int)((px - P.fx0) * P.finvhx), 0), P.nx - 1);
            iy = min(max((int)((py - P.fy0) * P.finvhy), 0), P.ny - 1);
            iz = state == ST_TOP ? P.nz - 1 : 0;
          }
        }
      }
      int comp = 1, pidx = 1;
      if (state == ST_TOP) {                                                     // INT:573-617
        add_flux(P, T, 0, ix + P.nx * iy, ew);
        state = ST_DEAD;
      } else if (state == ST_SURFACE) {                                          // INT:619-702
        add_flux(P, T, 1, ix + P.nx * iy, ew);
        atomicAdd(&sCnt[0], 1u);
        ew = (float)((double)ew * P.albedo);
        if (ew <= TINY32) {
          atomicAdd(&sCnt[1], 1u);                                               // absorbed by the surface
          state = ST_DEAD;
        } else {
          pz = P.fz0; iz = 0;
        }
      } else if (state == ST_SCATTER) {                                          // INT:703-811
        scatters++;
        const int cell = ix + P.nx * (iy + P.ny * iz);
        int ci = -1;                                     // CROP: the cell's index in the column-compressed arrays
        float lo = 0.0f, hi = 1.0f, ssa;
        {                                                                        // the cell's event record: ONE gather
          const uint32_t *R = P.rec + ((size_t)MCB_CHECK_INDEX(P, cell, cells) << P.recShift);
          if (CROP) {                                // the column-compressed copy of the records (stays in L2)
            const uint2 ct = __ldg(P.colTab + MCB_CHECK_INDEX(P, (ix + GH) + P.lin.nxp * (iy + GH), P.lin.nxp * P.lin.nyp));
            const unsigned lo = ct.y & 0xffffu, rel = (unsigned)iz - lo;
            if (rel < (ct.y >> 16) - lo) {
              ci = (int)MCB_CHECK_INDEX(P, ct.x + rel, P.nCompact);
              R = P.recC + ((size_t)ci << P.recShift);
            }
          }
          if (P.nc == 1) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(R));
            ssa = __uint_as_float(v.x); pidx = (int)(v.y & 0xffffu);
          } else if (P.nc == 2) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(R));
            const float cc = __uint_as_float(v.x);                               // findIndex on (0, cumExt(:)), NUM:262-315
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
        pidx = max(pidx, 1);                   // entry 0 marks a cell the component is absent from: never index before the table
        if (ssa < 1.0f) {                                                        // INT:765-771
          // fluxAbsorbed = column sum of this tally (column_absorption_kernel).  CROP: the absorbing cells are the cells of
          // the column ranges, and their f64 tally lives in the compact array too (18 MB on C5, L2-resident; the dense
          // 127 MB array drew 10 GB of DRAM traffic per 1e7 photons and pushed the field out of L2): expanded after the launch
          if (CROP && ci >= 0) atomicAdd(&P.tallyC[ci], (double)(ew * (1.0f - ssa)));
          else add_vol(P, T, cell, ew * (1.0f - ssa));
          ew *= ssa;
        }
        if (P.opt.useRussianRoulette && ew < P.opt.russianRouletteW * 0.5f) {    // INT:805-811
          const float uRR = P.nc > 1 ? __fdividef(eNext - lo, fmaxf(hi - lo, TINY32)) : eNext;
          if (uRR >= ew / P.opt.russianRouletteW) ew = 0.0f; else ew = P.opt.russianRouletteW;
        }
        if (ew <= TINY32) {
          atomicAdd(&sCnt[2], 1u);                                               // killed by roulette
          state = ST_DEAD;
        }
      }
      // ---- finished photons are replaced from the global counter: one atomic per warp (getNextPhoton, ILL:561-590) ----
      {
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
      }
      // ---- one Philox block per event: (angle | position, azimuth | position, next optical depth, next pick) ----
      const bool alive = state != ST_DONE;
      if (alive) {
        const float4 u = rng.block(k0, k1);
        if (state == ST_BORN) {
          float x01, y01, z01;
          int bi = -1, bj = 0, bk = 0;
          ew = 1.0f;
          if (P.source == 0) {                                                   // ILL:88-96
            x01 = u.x; y01 = u.y; z01 = 1.0f - FLT_EPSILON;
            dx = P.solarDir[0]; dy = P.solarDir[1]; dz = P.solarDir[2];
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
              const int ik = cdf_search(P.colCDF + (P.ny - 1), P.nz, (long long)P.ny, q);
              const int ij = cdf_search(P.colCDF + (size_t)P.ny * (size_t)(ik - 1), P.ny, 1, q);
              const double *voxBase = P.voxelCDF + (size_t)P.nx * ((size_t)(ij - 1) + (size_t)P.ny * (size_t)(ik - 1));
              const int ii = cdf_search(voxBase, P.nx, 1, q);
              const float4 v2 = rng.block(k0, k1);                               // uniform inside the cell (ILL:500-505)
              z01 = ((float)(ik - 1) + fminf(fmaxf(v2.x, 1e-6f), 1.0f - 1e-6f)) / (float)P.nz;
              x01 = ((float)(ii - 1) + fminf(v2.y, 1.0f - 1e-6f)) / (float)P.nx;
              y01 = ((float)(ij - 1) + fminf(v2.z, 1.0f - 1e-6f)) / (float)P.ny;
              bi = ii - 1; bj = ij - 1; bk = ik - 1;
              mu = 1.0f - 2.0f * v.x;                                            // ILL:507-509 retries on mu ~ 0
              if (!(fabsf(mu) > 2.0f * TINY32)) mu = 1.0e-30f;
              phi = v.y * 2.0f * PI32;
            }
            dir_from(mu, phi, dx, dy, dz);
          }
          px = fmaf(x01, P.fLx, P.fx0); py = fmaf(y01, P.fLy, P.fy0); pz = fmaf(z01, P.fLz, P.fz0);   // INT:478-494
          ix = min((int)(x01 * (float)P.nx), P.nx - 1);
          iy = min((int)(y01 * (float)P.ny), P.ny - 1);
          iz = min((int)(z01 * (float)P.nz), P.nz - 1);
          if (bi >= 0) { ix = bi; iy = bj; iz = bk; }
          if (P.opt.LW_flag > 0.0f && pz > 0.0f) add_vol(P, T, ix + P.nx * (iy + P.ny * iz), -1.0f);   // INT:504-508
        } else if (state == ST_SURFACE) {                                        // INT:655-676
          const float mu = sqrtf(fmaxf(u.x, 1.0e-30f));                          // retries on mu ~ 0
          dir_from(mu, 2.0f * PI32 * u.y, dx, dy, dz);
        } else {                                                                 // ST_SCATTER, INT:813-819
          const int c = comp - 1;
          const int nS = P.invS[c];
          const float *tab = P.inv[c] + (size_t)MCB_CHECK_INDEX(P, pidx - 1, P.invE[c]) * nS;
          const float rn = u.x;                                                  // computeScatteringAngle INT:1594-1621
          const int k = (int)(rn * (float)nS) + 1;
          float theta;
#ifdef MCB_POOL_SMEM_INV
          if (sInv) {
            theta = k < nS ? (1.0f - (rn - (float)(k - 1) / (float)nS)) * sInv[k - 1] + (rn - (float)(k - 1) / (float)nS) * sInv[k] : sInv[nS - 1];
          } else
#endif
          if (k < nS) {
            const float left = rn - (float)(k - 1) / (float)nS;
            theta = (1.0f - left) * __ldg(&tab[MCB_CHECK_INDEX(P, k - 1, nS)]) + left * __ldg(&tab[MCB_CHECK_INDEX(P, k, nS)]);
          } else {
            theta = __ldg(&tab[nS - 1]);
          }
          float sinT, cosT;
          __sincosf(theta, &sinT, &cosT);
          float AX, AY;                                                          // next_direct INT:1921-1948 (uniform azimuth)
          __sincosf(2.0f * PI32 * u.y, &AY, &AX);
          AX *= sinT; AY *= sinT;
          const float Bq = dx * AX - dy * AY;
          const float D = cosT - __fdividef(Bq, 1.0f + fabsf(dz));
          const float ndx = dx * D + AX, ndy = dy * D - AY;
          const float ndz = dz * cosT - copysignf(fabsf(Bq), dz * Bq);
          dx = ndx; dy = ndy; dz = ndz;
        }
        eTau = -__logf(fmaxf(TINY32, u.z));                                      // INT:554
        eNext = u.w;
      }
      // ---- push the new legs, completely set up, onto READY ----
      const unsigned m = __ballot_sync(FULL, alive);
      if (alive) {
        const int s = nR + __popc(m & below);
        {
          Ray q;
          q.ox = px; q.oy = py; q.oz = pz; q.dx = dx; q.dy = dy; q.dz = dz; q.ix = ix; q.iy = iy; q.iz = iz;
          ray_start<true>(q, P, G);
          pool[PW_TX * POOL_SLOTS + s] = q.tx; pool[PW_TY * POOL_SLOTS + s] = q.ty; pool[PW_TZ * POOL_SLOTS + s] = q.tz;
        }
        pool[PW_PX * POOL_SLOTS + s] = px; pool[PW_PY * POOL_SLOTS + s] = py; pool[PW_PZ * POOL_SLOTS + s] = pz;
        pool[PW_DX * POOL_SLOTS + s] = dx; pool[PW_DY * POOL_SLOTS + s] = dy; pool[PW_DZ * POOL_SLOTS + s] = dz;
        pool[PW_W * POOL_SLOTS + s] = ew; pool[PW_TAU * POOL_SLOTS + s] = eTau; pool[PW_UNEXT * POOL_SLOTS + s] = eNext;
        pool[PW_C0 * POOL_SLOTS + s] = __uint_as_float(rng.c0); pool[PW_C1 * POOL_SLOTS + s] = __uint_as_float(rng.c1);
        pool[PW_BLK * POOL_SLOTS + s] = __uint_as_float(rng.blk);
        pool[PW_IXY * POOL_SLOTS + s] = __int_as_float(ix | (iy << 16));
        pool[PW_IZK * POOL_SLOTS + s] = __int_as_float(iz);
      }
      nR += __popc(m);
      __syncwarp();
    }

    // =========================== lanes without a photon take a READY leg ===========================
    {
      const unsigned idle = __ballot_sync(FULL, !have);
      if (idle && nR > 0) {
        const int rank = __popc(idle & below);
        if (!have && rank < nR) {
          const int s = nR - 1 - rank;
          r.ox = pool[PW_PX * POOL_SLOTS + s]; r.oy = pool[PW_PY * POOL_SLOTS + s]; r.oz = pool[PW_PZ * POOL_SLOTS + s];
          r.dx = pool[PW_DX * POOL_SLOTS + s]; r.dy = pool[PW_DY * POOL_SLOTS + s]; r.dz = pool[PW_DZ * POOL_SLOTS + s];
          w = pool[PW_W * POOL_SLOTS + s]; tau = pool[PW_TAU * POOL_SLOTS + s]; uNext = pool[PW_UNEXT * POOL_SLOTS + s];
          c0 = __float_as_uint(pool[PW_C0 * POOL_SLOTS + s]); c1 = __float_as_uint(pool[PW_C1 * POOL_SLOTS + s]);
          blk = __float_as_uint(pool[PW_BLK * POOL_SLOTS + s]);
          const int ixy = __float_as_int(pool[PW_IXY * POOL_SLOTS + s]);
          r.ix = ixy & 0xffff; r.iy = (int)((uint32_t)ixy >> 16);
          r.iz = __float_as_int(pool[PW_IZK * POOL_SLOTS + s]);
          r.tx = pool[PW_TX * POOL_SLOTS + s]; r.ty = pool[PW_TY * POOL_SLOTS + s]; r.tz = pool[PW_TZ * POOL_SLOTS + s];
          r.rx = safe_rcp(r.dx); r.ry = safe_rcp(r.dy); r.rz = safe_rcp(r.dz);
          r.t = 0.0f; ext = 0.0f; vcur = 0.0f;
          have = true;
        }
        nR -= min(__popc(idle), nR);
        __syncwarp();
      }
    }
    if (!__any_sync(FULL, have)) break;                    // nothing marching, nothing ready, no event waiting: done

    // =========================== march: one burst for every lane ===========================
    // (a lane whose cell is known to lie deep enough in vacuum crosses that in one leap first: march_leap)
    int ev = MARCH_ON;
    if (LEAP) {
      int D = have ? leap_distance(r, P, vcur, leapBelow) : 0;
      if (leapLanes > 1 && __popc(__ballot_sync(FULL, D > 0)) < leapLanes) D = 0;   // too few lanes to pay for the divergence
      if (D) ev = march_leap<MASK || CROP>(r, P, D, crossings, ext, tau, &sCnt[4]); // (implies have)
    }
    if (have && ev == MARCH_ON)
      ev = march_burst<true, true, BURST, MASK, BRICK, true, SPLIT, LEAP, CROP>(r, P, G, ext, tau, crossings, LEAP ? &vcur : nullptr);

    // =========================== photons that reached an event go onto EVENT ===========================
    {
      const bool arrived = have && ev != MARCH_ON;
      const unsigned hit = __ballot_sync(FULL, arrived);
      if (hit) {
        if (arrived) {
          const int s = POOL_SLOTS - nE - 1 - __popc(hit & below);       // the ray as the burst left it (RAW)
          pool[PW_PX * POOL_SLOTS + s] = r.ox; pool[PW_PY * POOL_SLOTS + s] = r.oy; pool[PW_PZ * POOL_SLOTS + s] = r.oz;
          pool[PW_DX * POOL_SLOTS + s] = r.dx; pool[PW_DY * POOL_SLOTS + s] = r.dy; pool[PW_DZ * POOL_SLOTS + s] = r.dz;
          pool[PW_W * POOL_SLOTS + s] = w; pool[PW_TAU * POOL_SLOTS + s] = r.t; pool[PW_UNEXT * POOL_SLOTS + s] = uNext;
          pool[PW_C0 * POOL_SLOTS + s] = __uint_as_float(c0); pool[PW_C1 * POOL_SLOTS + s] = __uint_as_float(c1);
          pool[PW_BLK * POOL_SLOTS + s] = __uint_as_float(blk);
          pool[PW_IXY * POOL_SLOTS + s] = __int_as_float(r.ix);
          pool[PW_IZK * POOL_SLOTS + s] = __int_as_float(ev << 28);
          have = false;
        }
        nE += __popc(hit);
        __syncwarp();
      }
    }
  }

  // ---- flush: event counters (warp shuffle reduce, one atomic per warp) ----
  {
    unsigned long long v[2] = {crossings, scatters};
    const int slotOf[2] = {CNT_CROSSINGS, CNT_SCATTERS};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
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
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&P.tally[P.offPhotons], (double)nPhotons);
    atomicAdd(&P.counters[CNT_PHOTONS], (unsigned long long)nPhotons);
  }
}

}  // namespace mcbpool

template <int MINBLOCKS, int BURST, bool SPLIT, bool MASK, bool BRICK, bool LEAP, bool CROP = false>
static void launch_pool(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId, int numSMs,
                        unsigned long long *workCounter, cudaStream_t stream) {
  constexpr int THREADS = 128;
  auto kernel = mcbpool::pool_kernel<THREADS, MINBLOCKS, BURST, SPLIT, MASK, BRICK, LEAP, CROP>;
  const int cols = P.nx * P.ny, cells = cols * P.nz;
  mcbfast::SmemPlan plan{-1, -1, -1, -1, -1, 0, 0, 0};
  int off = 0;
  plan.poolOff = off; off += (THREADS / 32) * POOL_WORDS * POOL_SLOTS;
  const int budgetFloats = 9 * 1024;             // small grids are atomic hot spots: privatise (as mcb_fast.cu does)
  if (cols <= 1024 && 2 * cols <= budgetFloats) { plan.fluxOff = off; off += 2 * cols; }
  if (cells <= 8192 && (plan.fluxOff >= 0 ? 2 * cols : 0) + cells <= budgetFloats) { plan.volOff = off; off += cells; }
#ifdef MCB_POOL_SMEM_INV
  if (P.nc == 1 && P.invE[0] == 1 && P.invS[0] <= 10240) { plan.intOff = off; off += P.invS[0]; }
#endif
  plan.totalFloats = off;
  const size_t smem = sizeof(float) * (size_t)off;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int blocksPerSM = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kernel, THREADS, smem) != cudaSuccess || blocksPerSM < 1)
    blocksPerSM = 1;
  if (P.opt.tuneBlocksPerSM > 0 && P.opt.tuneBlocksPerSM < blocksPerSM) blocksPerSM = P.opt.tuneBlocksPerSM;
  const long long perBlock = (long long)(THREADS / 32) * POOL_SLOTS;       // photons a CTA holds at once
  const long long want = (nPhotons + perBlock - 1) / perBlock;
  const long long cap = (long long)numSMs * blocksPerSM;    // persistent: every CTA resident, whole waves only
  const int blocks = (int)(want < cap ? want : cap);
  // vacuum leaps: from a distance of tuneLeap cells (default MCB_LEAP_MIN; < 0: never)
  const float leapBelow = P.opt.tuneLeap < 0 ? -FLT_MAX : -(float)(P.opt.tuneLeap >= 2 ? P.opt.tuneLeap : MCB_LEAP_MIN);
  kernel<<<blocks, THREADS, smem, stream>>>(P, nPhotons, seed, firstPhotonId, workCounter, plan, leapBelow, P.opt.tuneLeapLanes > 0 ? P.opt.tuneLeapLanes : MCB_LEAP_LANES);
}

void mcb_launch_expand_compact_tally(const DevDomain &P, int numSMs, cudaStream_t stream);     // mcb_stage.cu

// the pool kernel covers flux-only runs on uniform grids at least a ghost shell wide
bool mcb_pool_covers(const DevDomain &P) {
  return P.uniform && P.nx >= MCB_GHOST && P.ny >= MCB_GHOST && P.nDir == 0 && P.nx <= 65535 && P.ny <= 65535 && P.nz <= 65535;
}

// ... and is the default where it measured faster (one B200, r02: C3 6.56e8 -> 7.48e8 photons/s, C3 Mie 6.28e8 -> 7.0e8,
// C5 3.76e8 -> 4.41e8): grids whose tallies go straight to the f64 buffer.  On grids small enough for shared-memory
// tallies the park/regroup kernel stays ahead (C1 9.9e8 vs 9.3e8, C4 5.0e9 vs 4.2e9): its photons are short-lived and
// the pool's 17 KB of shared memory per CTA come on top of the privatised tallies.
bool mcb_pool_preferred(const DevDomain &P) {
  const long long cols = (long long)P.nx * P.ny;
  return mcb_pool_covers(P) && cols > 1024;
}

void mcb_launch_pool_batch(const DevDomain &P, long long nPhotons, uint64_t seed, uint64_t firstPhotonId,
                           int numSMs, unsigned long long *workCounter, cudaStream_t stream) {
  if (nPhotons <= 0) return;
  const bool mask = P.lin.mask != nullptr, brick = P.opt.tuneLayout != MCB_LAYOUT_LINEAR;
  // tuneBurst: 8 = eight cells per burst, all gathers up front; 4 = four; 44 = eight cells, gathers in two halves (SPLIT)
  // default: split gathers on L2-resident fields (C3: 7.09e8 -> 7.27e8); with the occupancy bitmap (C5) the second
  // round trip costs more than the saved gathers (4.07e8 vs 3.92e8)
  const int burst = P.opt.tuneBurst ? P.opt.tuneBurst : (mask ? 8 : 44);
  // register budget: 8 CTAs/SM = 64 registers (spills), 7 = 72, 6 = 80.  Measured (one B200, r02, photons/s): C3 with split
  // gathers 6: 7.27e8, 7: 7.48e8, 8: 6.0e8; C5 (bitmap, eight gathers up front) 6: 4.41e8, 7: 4.03e8
  const int occ = P.opt.tuneBlocksPerSM ? P.opt.tuneBlocksPerSM : (mask ? 6 : 7);
  // LEAP variants (vacuum / clear-layer leaps, clamped gathers) only where the staging found space worth leaping (P.leap):
  // on a scene without any (C3 with a Rayleigh background) the machinery alone costs 7 % (r02: 7.40e8 -> 6.88e8)
#define MCB_POOL_GO(OCC, B, SPLIT, MASK, BRICK) \
  do { if (P.leap) launch_pool<OCC, B, SPLIT, MASK, BRICK, true>(P, nPhotons, seed, firstPhotonId, numSMs, workCounter, stream); \
       else launch_pool<OCC, B, SPLIT, MASK, BRICK, false>(P, nPhotons, seed, firstPhotonId, numSMs, workCounter, stream); } while (0)
#define MCB_POOL_LAYOUT(OCC, B, SPLIT) \
  do { if (mask) { if (brick) MCB_POOL_GO(OCC, B, SPLIT, true, true); else MCB_POOL_GO(OCC, B, SPLIT, true, false); } \
       else { if (brick) MCB_POOL_GO(OCC, B, SPLIT, false, true); else MCB_POOL_GO(OCC, B, SPLIT, false, false); } } while (0)
  // the layer-cropped field instead of the bitmap (fields too large for L2 whose cloud band fits; tuneExtMask = 1 keeps
  // the bitmap)
  if (mask && P.crp.ext && P.colTab && P.opt.tuneExtMask != 1) {
#define MCB_POOL_CROP(OCC, SPLIT) \
  do { if (P.leap) launch_pool<OCC, 8, SPLIT, false, true, true, true>(P, nPhotons, seed, firstPhotonId, numSMs, workCounter, stream); \
       else launch_pool<OCC, 8, SPLIT, false, true, false, true>(P, nPhotons, seed, firstPhotonId, numSMs, workCounter, stream); } while (0)
    const int occC = P.opt.tuneBlocksPerSM ? P.opt.tuneBlocksPerSM : MCB_CROP_OCC;
    const bool splitC = P.opt.tuneBurst ? P.opt.tuneBurst == 44 : MCB_CROP_SPLIT;
    if (occC >= 7) { if (splitC) MCB_POOL_CROP(7, true); else MCB_POOL_CROP(7, false); }
    else { if (splitC) MCB_POOL_CROP(6, true); else MCB_POOL_CROP(6, false); }
#undef MCB_POOL_CROP
    mcb_launch_expand_compact_tally(P, numSMs, stream);
    return;
  }
  if (occ >= 8) { if (burst == 44) MCB_POOL_LAYOUT(8, 8, true); else MCB_POOL_LAYOUT(8, 8, false); }
  else if (occ == 7) { if (burst == 44) MCB_POOL_LAYOUT(7, 8, true); else MCB_POOL_LAYOUT(7, 8, false); }
  else { if (burst == 4) MCB_POOL_LAYOUT(6, 4, false); else if (burst == 44) MCB_POOL_LAYOUT(6, 8, true); else MCB_POOL_LAYOUT(6, 8, false); }
#undef MCB_POOL_LAYOUT
#undef MCB_POOL_GO
}
