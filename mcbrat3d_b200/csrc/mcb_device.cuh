// mcb_device.cuh -- device-side data model shared by the photon kernels (sm_100a).
//
// Everything the photon loop reads is staged once into HBM by the mcb_set_* calls
// (mcb_api.cu) and described to the kernels by one DevDomain parameter block.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/mcbrat_cuda.h"

#define MCB_MAX_COMP 8          // optical components per domain (the reference's decks use <= 4)
#define MCB_MAX_DIR 32          // view directions kept in the parameter block
#define MCB_GHOST 8             // ghost cells on every side of the packed extinction field = longest marching burst
#define MCB_LEAP_MIN 2          // smallest vacuum distance the pool kernels leap from (mcb_options.tuneLeap overrides); one
                                // B200, r02, C3: 2 -> 7.84e8, 3 -> 7.77e8, 4 -> 7.69e8 photons/s (C3 + views 1.101e8 / 1.093e8)
#define MCB_LEAP_LANES 1        // lanes of a warp that must want a leap for the warp to run the leap code (tuneLeapLanes).
                                // 8 measured 1 % faster on C3, but then a photon's leaps -- and with them the last bits of its
                                // positions -- depend on which other photons share its warp: results would no longer be
                                // independent of the batch split and of the number of GPUs (tests: ..._independent_of_batch_split)
#define MCB_LEAP_CAP 64         // largest vacuum distance (cells) encoded in the packed field = longest leap of the pool kernels

struct DevDomain {
  // ---- grid (OPT:77-83, INT:60-66) ----
  int nx, ny, nz, nc;
  const double *xE, *yE, *zE;                 // edges, n+1 each
  int xyRegular, zRegular;                    // INT:163-181 (f32-rounded spacing test, quirk q1)
  int uniform;                                // all three spacings uniform to 1e-6 relative: the throughput kernel steps
                                              // such grids incrementally even when q1 sends the reference down its
                                              // irregular-grid path (0.05 km is not exactly representable in f32)
  double deltaX, deltaY, deltaZ;
  double x0, y0, z0, xMax, yMax, zMax;
  // ---- optical properties, reference layout and precision ----
  const double *totalExt;                     // (nx,ny,nz)
  const double *cumExt, *ssa;                 // (nx,ny,nz,nc)
  const int32_t *phaseIdx;                    // (nx,ny,nz,nc)
  double albedo;
  float maxExtinction;                        // real(maxval(totalExt)), INT:448 (maximum cross-section only)
  // ---- packed single-precision copies for the fast kernel ----
  // The padded extinction field (ghost shell of MCB_GHOST cells: periodic replicas in x, y; zeros above and below)
  // is kept in two layouts, each described by an ExtField:
  //   lin -- x fastest: cell (ix,iy,iz) at ix + nxp*(iy + nyp*iz).  Read by the local-estimation kernels and on narrow
  //          or irregular grids.
  //   brk -- 2x2x2-cell bricks, one brick per 32-byte sector, bricks x-fastest (padded dimensions rounded up to even):
  //          address(i,j,k) = 8*(((k>>1)*by + (j>>1))*bx + (i>>1)) + 4*(k&1) + 2*(j&1) + (i&1)
  //                         = 4i - 3(i&1) + cY j - (cY-2)(j&1) + cZ k - (cZ-4)(k&1),   cY = 4 bx, cZ = 4 bx by.
  //          Whatever axis a ray steps along, the next cell is in the same sector half of the time (x-fastest rows:
  //          7/8 of the x steps, none of the others), which is what bounds the flux kernels: L1TEX->XBAR requests.
  // Fields too large for L2 (C5: 78 MB) also carry an occupancy bitmap per layout: bit p (p = absolute padded address)
  // is set where the cell's extinction differs from its layer's clear-sky value layerExt[iz + G] (the layer minimum; 0
  // in the ghost layers).  The marcher reads the bitmap (1 bit per cell, L1/L2-resident) and gathers the field only
  // where the bit is set, so clear-sky cells -- most of a cloud scene -- never go to HBM.  Exact: both branches give
  // (float)totalExt.
  struct ExtField {
    const float *ext;                         // points AT the first real cell
    const uint32_t *mask;                     // ceil(padded / 32) words, or nullptr
    int nxp, nyp, origin;                     // padded row / slice lengths; address of the first real cell
    int cY, cZ;                               // bricked layout only
    uint32_t divSliceM, divRowM;              // address -> (ix,iy,iz): q = (M * n) >> S, exact for n < 2^31
    int divSliceS, divRowS;
    long long padded;                         // number of padded cells
  } lin, brk, crp;
  // crp: fields too large for L2 whose cloud occupies a band of layers -- the bricked field of the layers
  // [cropLo, cropLo + cropN) only (both even; periodic ghost shell in x and y, none in z), small enough to stay in L2.
  // Outside the band every layer is clear throughout and the marcher takes layerExt.  ext == nullptr: not built.
  int cropLo, cropN;
  const float *layerExt;                      // nz + 2G (+2) clear-sky values; sign bit set: the WHOLE layer has this value
                                              // (no bitmap look-up needed there)
  int leap;                                   // the packed fields carry the vacuum distances / the layer tables below are
                                              // worth leaping with: the pool kernels run their LEAP variants
  const float *layerLeap;                     // nz + 2G (+2): minus the distance, in layers, to the nearest layer that is not
                                              // clear throughout (0 for such a layer): march_leap crosses that many at once
  const float *layerCum;                      // nz + 1: clear-sky optical depth per unit |1/mu| from the surface to each edge
  // Column-compressed event records (photon-pool flux kernel on fields too large for L2): per padded column the range
  // of layers [lo, hi) outside which every cell has its layer's clear-sky value, and for the cells inside the ranges --
  // a few per cent of a cloud scene -- the event record, the cell index and an f64 absorption tally, stored densely column by
  // column: 35 MB instead of 253 MB on C5, so the records of the scattering cells stay in L2 next to the cropped field.
  // (Measured, r02: marching on this storage as well -- one table look-up per crossing, then the gather -- was 5 % SLOWER
  // than the bitmap: what bounds C5 is the dependent look-up -> gather chain of a crossing, not where the gather is served.)
  const uint2 *colTab;                        // lin.nxp x lin.nyp: .x = compact index of layer lo, .y = lo | hi << 16
  const uint32_t *recC;                       // nCompact << recShift: the cells' event records
  const uint32_t *cellC;                      // nCompact: ix + nx * (iy + ny * iz)
  double *tallyC;                             // nCompact: volume-absorption tally of those cells (added into tally after a launch)
  long long nCompact;
  uint32_t divColsM, divNxM; int divColsS, divNxS;   // unpadded cell -> (ix, iy, iz): division by nx * ny and by nx
  // one record per cell with everything a scattering event reads, so an event costs ONE gather (a 32 B sector for
  // nc <= 3) instead of a dependent chain through three arrays: 2^recShift u32 words =
  // [f32 cumExt(c), c = 1..nc-1][f32 ssa(c), c = 1..nc][u16 phase index pairs], zero-padded
  const uint32_t *rec;
  int recShift;
  float fx0, fy0, fz0, fLx, fLy, fLz;         // single-precision grid scalars (fast kernel, constant bank)
  float fhx, fhy, fhz, finvLx, finvLy, finvhx, finvhy, finvhz, fzMax;
  // ---- tables ----
  const float *inv[MCB_MAX_COMP];  int invS[MCB_MAX_COMP];  int invE[MCB_MAX_COMP], fwdE[MCB_MAX_COMP];   // steps, entries
  const float *fwd[MCB_MAX_COMP];  const float *fwdOrig[MCB_MAX_COMP];  int fwdS[MCB_MAX_COMP];
  float fwdInvDTheta[MCB_MAX_COMP];           // (nS - 1) / pi: the forward tables are on equal angle steps (OPT:1912-1913)
  // ---- views ----
  int nDir;
  float viewDir[3 * MCB_MAX_DIR];
  float viewNorm[MCB_MAX_DIR];                // 1 / (4 pi |mu_view|), INT:1696, 1726
  // ---- options ----
  mcb_options opt;
  // ---- source ----
  int source;                                 // 0 solar, 1 thermal
  float solarMu, solarPhi;                    // ILL:95-96 values
  float solarDir[3];                          // their direction cosines (INT:1876-1894)
  double fracAtmsPower;
  const double *voxelCDF;                     // (nx,ny,nz)
  const double *colCDF;                       // (ny,nz): voxelCDF(nx,:,:), the column weights of EMI:56 gathered into one
                                              // small array (the level weights are its rows' last entries)
  // ---- tallies: packed f64 buffer ----
  double *tally;
  long long offFluxUp, offFluxDown, offFluxAbs, offVolAbs, offInt, offIntByComp, offExcess, offPhotons;
  unsigned long long *counters;               // mcb_counters as 16 x u64
};

// Bounds-checked build (libmcbrat_cuda_dbg.so, -DMCB_BOUNDS_CHECK): every gather index of the fast kernel is tested
// against its array and violations are counted in counters[CNT_BAD] (the access is clamped, not made).  The pool
// has no compute-sanitizer, so tests/test_gpu_bounds.py runs the cases through this build and asserts zero.
#ifdef MCB_BOUNDS_CHECK
#define MCB_LEAP_STATS
#define MCB_CHECK_INDEX(P, i, n) mcb_checked_index((P), (long long)(i), (long long)(n))
__device__ __forceinline__ long long mcb_checked_index(const struct DevDomain &P, long long i, long long n);
#else
#define MCB_CHECK_INDEX(P, i, n) (i)
#endif

__host__ __device__ inline long long mcb_brick_address(int i, int j, int k, int bx, int by) {   // padded coordinates
  return ((((long long)(k >> 1) * by + (j >> 1)) * bx + (i >> 1)) << 3) | ((k & 1) << 2) | ((j & 1) << 1) | (i & 1);
}

enum { CNT_PHOTONS = 0, CNT_CROSSINGS, CNT_SCATTERS, CNT_SURFACE, CNT_TOP, CNT_BAD,
       CNT_LE_RAYS, CNT_LE_CROSSINGS, CNT_RR_KILLS, CNT_SURFACE_KILLS, CNT_N,     // [CNT_N] is the photon work counter
       CNT_LEAPS, CNT_LEAP_CELLS, CNT_END };

#ifdef MCB_BOUNDS_CHECK
__device__ __forceinline__ long long mcb_checked_index(const DevDomain &P, long long i, long long n) {
  if (i < 0 || i >= n) { atomicAdd(&P.counters[CNT_BAD], 1ull); return 0; }
  return i;
}
#endif

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011).  One stream per photon:
// key = (seed lo, seed hi), counter = (photon id lo, photon id hi, draw block, 0).
// Replaces RandomNumbersForMC's sequential MT19937 stream (RNG:118-301): the numbers a
// photon sees depend only on (seed, photon id), never on batch or GPU decomposition.
// ---------------------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1, c0, c1, blk;
  uint32_t b0, b1, b2, b3;
  int have;
  unsigned ndrawn;

  __device__ __forceinline__ void init(uint64_t seed, uint64_t photon) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
    c0 = (uint32_t)photon; c1 = (uint32_t)(photon >> 32);
    blk = 0; have = 0; ndrawn = 0;
  }
  __device__ __forceinline__ void refill() {
    uint32_t x0 = c0, x1 = c1, x2 = blk, x3 = 0u, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
      uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    b0 = x0; b1 = x1; b2 = x2; b3 = x3;
    blk++; have = 4;
  }
  __device__ __forceinline__ uint32_t next_u32() {
    if (have == 0) refill();
    ndrawn++; have--;
    uint32_t r = b0; b0 = b1; b1 = b2; b2 = b3;
    return r;
  }
};
