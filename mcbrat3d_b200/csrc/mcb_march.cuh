// mcb_march.cuh -- what the throughput kernels share: the per-photon Philox stream, the ray state, the burst marcher
// (accumulateExtinctionAlongPath, OPT:1656-1815, as a geometry-ahead DDA over the ghost-shelled f32 extinction field),
// source / rotation helpers and the tally updates.  Included by mcb_fast.cu (park/regroup megakernel, local
// estimation) and mcb_pool.cu (photon-pool flux kernel).
#pragma once
#include "mcb_device.cuh"

namespace mcbfast {

#define FULL 0xffffffffu
#define PI32 3.14159265358979312f
#define TINY32 FLT_MIN
#define GH MCB_GHOST
// gather of the padded extinction field (rel = index relative to the first real cell; may be negative in the shell)
#ifdef MCB_BOUNDS_CHECK
#define EXT_AT(P, F, rel) __ldg((F).ext + (mcb_checked_index((P), (long long)(rel) + (F).origin, (F).padded) - (F).origin))
#else
#define EXT_AT(P, F, rel) __ldg((F).ext + (rel))
#endif

enum { ST_DEAD = 0, ST_MARCH = 1, ST_SCATTER = 2, ST_SURFACE = 3, ST_TOP = 4, ST_BORN = 5, ST_DONE = 6 };

// launch-time layout of the dynamic shared memory
struct SmemPlan {
  int edgesOff;            // float[(nx+1+2G) + (ny+1+2G) + (nz+1+2G)] ghost-extended edges (irregular grids; -1 otherwise)
  int fluxOff;             // float[2*cols]  privatised fluxUp|fluxDown           (-1: global atomics)
  int volOff;              // float[cells]   privatised volumeAbsorption           (-1: global atomics)
  int intOff;              // float[cols*nDir] privatised intensity                (-1: global atomics)
  int leOff;               // per warp: LE_WORDS x 32 request slots + 64 words of queue state: task counter, rank -> lane map
  int leStride;            // words per warp: the above (+ LE_CARRY_WORDS x 32 for parked view rays when leCarry >= 0)
  int totalFloats;
  int poolOff;             // photon-pool kernel (mcb_pool.cu): per warp POOL_WORDS x POOL_SLOTS record words
};

struct Rng {
  uint32_t c0, c1, blk;
  __device__ __forceinline__ void init(uint64_t photon) { c0 = (uint32_t)photon; c1 = (uint32_t)(photon >> 32); blk = 0; }
  // One Philox4x32-10 block = four uniform reals.  Every call site is reached by all the lanes that
  // take part in the event phase together, so the ten rounds run convergently.
  __device__ __forceinline__ float4 block(uint32_t k0, uint32_t k1) {
    uint32_t x0 = c0, x1 = c1, x2 = blk, x3 = 0u, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
      const uint32_t y0 = hi1 ^ x1 ^ a, y2 = hi0 ^ x3 ^ b;
      x0 = y0; x1 = lo1; x2 = y2; x3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    blk++;
    // f32 built from 32 bits like RNG:286-300, but on [0,1): the scale is the float just below 2^-32, so the 128
    // largest integers (which round to 2^32) give 1 - 2^-24 instead of 1.  The closed upper end of the reference's
    // generator only matters to its bit-exact traces (mcb_reference.cu); here a draw of exactly 1 could pick a trailing
    // component of zero extinction (cumExt = 1) whose phase-function entry is 0.
    const float S = __uint_as_float(0x2f7fffffu);
    return make_float4(__uint2float_rn(x0) * S, __uint2float_rn(x1) * S, __uint2float_rn(x2) * S, __uint2float_rn(x3) * S);
  }
};

struct Ray {
  float ox, oy, oz;        // leg origin (irregular grids: x, y shifted by whole domain periods when wrapping)
  float dx, dy, dz;        // direction cosines
  float rx, ry, rz;        // reciprocal direction cosines (FLT_MAX-guarded)
  float t;                 // distance along the leg
  float tx, ty, tz;        // distance along the leg at which the next x/y/z face is met
  int ix, iy, iz;          // 0-based cell; inside a burst x, y may run into the periodic ghost shell
};

struct Grid {
  const float *sx, *sy, *sz;     // shared-memory edges (irregular grids), pointing at edge 0 of ghost-extended
};                               // arrays; every scalar of the grid is a field of the DevDomain parameter block

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float safe_rcp(float d) {
  return fabsf(d) >= 2.0f * TINY32 ? rcp_approx(d) : FLT_MAX;            // OPT:1705-1712 zero-direction guard
}

template <bool REG>
__device__ __forceinline__ float edge_x(const DevDomain &P, const Grid &G, int i) { return REG ? fmaf((float)i, P.fhx, P.fx0) : G.sx[i]; }
template <bool REG>
__device__ __forceinline__ float edge_y(const DevDomain &P, const Grid &G, int i) { return REG ? fmaf((float)i, P.fhy, P.fy0) : G.sy[i]; }
template <bool REG>
__device__ __forceinline__ float edge_z(const DevDomain &P, const Grid &G, int i) { return REG ? fmaf((float)i, P.fhz, P.fz0) : G.sz[i]; }

template <bool REG>
__device__ __forceinline__ void ray_start(Ray &r, const DevDomain &P, const Grid &G) {
  r.rx = safe_rcp(r.dx); r.ry = safe_rcp(r.dy); r.rz = safe_rcp(r.dz);
  r.t = 0.0f;
  r.tx = r.rx == FLT_MAX ? FLT_MAX : fmaxf((edge_x<REG>(P, G, r.ix + (r.dx >= 0.0f ? 1 : 0)) - r.ox) * r.rx, 0.0f);
  r.ty = r.ry == FLT_MAX ? FLT_MAX : fmaxf((edge_y<REG>(P, G, r.iy + (r.dy >= 0.0f ? 1 : 0)) - r.oy) * r.ry, 0.0f);
  r.tz = r.rz == FLT_MAX ? FLT_MAX : fmaxf((edge_z<REG>(P, G, r.iz + (r.dz >= 0.0f ? 1 : 0)) - r.oz) * r.rz, 0.0f);
}

// Position on the leg, folded back into the periodic domain (the leg origin is only ever shifted by
// whole periods, so the fold is valid for both grid kinds and after a burst has been rolled back).
__device__ __forceinline__ void ray_position(const Ray &r, const DevDomain &P, float &px, float &py, float &pz) {
  px = fmaf(r.t, r.dx, r.ox); py = fmaf(r.t, r.dy, r.oy); pz = fmaf(r.t, r.dz, r.oz);
  px -= P.fLx * floorf((px - P.fx0) * P.finvLx);
  py -= P.fLy * floorf((py - P.fy0) * P.finvLy);
}

// Edge-table grids: make an event's folded position agree with its cell.  ray_position() folds by the rounded quotient
// (p - x0) / L, the marcher wraps the integer cell index; for an event within rounding of the periodic seam (~1e-7 of
// the events) the two can disagree by one whole period -- position at one end of the domain, cell at the other.  On
// uniform grids the next leg then merely flies one period too far inside that cell (face distances advance by whole
// cell widths from the first, clamped one); with edge tables every face distance is recomputed from the absolute edge,
// so the first steps of the leg get NEGATIVE lengths, its optical depth goes negative and a local-estimate
// contribution w * P * exp(-tau) blows up (seen once in 3.7e8 scatterings: tests/test_gpu_first_interaction.py).
// The position is moved by the one period that puts it next to its cell.
__device__ __forceinline__ void seam_fix(const DevDomain &P, const Grid &G, int ix, int iy, float &px, float &py) {
  const float ex = px - 0.5f * (G.sx[ix] + G.sx[ix + 1]), ey = py - 0.5f * (G.sy[iy] + G.sy[iy + 1]);
  px -= ex > 0.5f * P.fLx ? P.fLx : (ex < -0.5f * P.fLx ? -P.fLx : 0.0f);
  py -= ey > 0.5f * P.fLy ? P.fLy : (ey < -0.5f * P.fLy ? -P.fLy : 0.0f);
}

__device__ __forceinline__ int find_cell(const float *e, int n, float x) {
  int lo = 0, hi = n;                      // e[lo] <= x < e[hi]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (x >= e[mid]) lo = mid; else hi = mid; }
  return lo;
}

// fold a cell index that ran into the periodic ghost shell (|excursion| <= MCB_GHOST) back into [0, n).
// WIDE grids (nx, ny >= MCB_GHOST, decided at launch) need one conditional add each way and no branch; only
// narrower ones (the 1-column step cloud) can be several periods off.
template <bool WIDE>
__device__ __forceinline__ int wrap_index(int i, int n) {
  if (WIDE) {
    i += i < 0 ? n : 0;
    i -= i >= n ? n : 0;
  } else {                               // one remainder instead of a loop per period: the lanes of a warp stay together
    i %= n;
    i += i < 0 ? n : 0;
  }
  return i;
}

// padded linear cell (relative to the first real cell) -> (ix, iy, iz): two divisions by
// launch-invariant divisors with the precomputed multipliers of the parameter block (exact for every
// padded cell index < 2^31); x, y come back folded into the domain.
template <bool WIDE, bool BRICK>
__device__ __forceinline__ void cell_decode(const DevDomain &P, const DevDomain::ExtField &F, int rel, int &ix, int &iy, int &iz);
template <bool WIDE, bool BRICK>
__device__ __forceinline__ void cell_decode(const DevDomain &P, int rel, int &ix, int &iy, int &iz) {
  cell_decode<WIDE, BRICK>(P, BRICK ? P.brk : P.lin, rel, ix, iy, iz);
}
template <bool WIDE, bool BRICK>
__device__ __forceinline__ void cell_decode(const DevDomain &P, const DevDomain::ExtField &F, int rel, int &ix, int &iy, int &iz) {
  const uint32_t c = (uint32_t)(rel + F.origin);
  if (BRICK) {
    const uint32_t b = c >> 3;                                                  // brick; the low three bits are (z, y, x) inside it
    const uint32_t bz = (uint32_t)(((uint64_t)F.divSliceM * b) >> F.divSliceS);
    const uint32_t rem = b - bz * (uint32_t)(F.cZ >> 2);
    const uint32_t by = (uint32_t)(((uint64_t)F.divRowM * rem) >> F.divRowS);
    const uint32_t bx = rem - by * (uint32_t)(F.cY >> 2);
    ix = wrap_index<WIDE>((int)(2u * bx + (c & 1u)) - GH, P.nx);
    iy = wrap_index<WIDE>((int)(2u * by + ((c >> 1) & 1u)) - GH, P.ny);
    iz = (int)(2u * bz + ((c >> 2) & 1u)) - GH;
  } else {
    const uint32_t z = (uint32_t)(((uint64_t)F.divSliceM * c) >> F.divSliceS);
    const uint32_t rem = c - z * (uint32_t)(F.nxp * F.nyp);
    const uint32_t y = (uint32_t)(((uint64_t)F.divRowM * rem) >> F.divRowS);
    ix = wrap_index<WIDE>((int)(rem - y * (uint32_t)F.nxp) - GH, P.nx);
    iy = wrap_index<WIDE>((int)y - GH, P.ny);
    iz = (int)z - GH;
  }
}

// bricked layout: address of cell (ix, iy, iz) relative to the first real cell (the ghost depth is even, so parities carry over)
__device__ __forceinline__ int brick_rel(const DevDomain::ExtField &F, int ix, int iy, int iz) {
  return 4 * ix - 3 * (ix & 1) + F.cY * iy - (F.cY - 2) * (iy & 1) + F.cZ * iz - (F.cZ - 4) * (iz & 1);
}
// address step of the next move along one axis: inside the brick (u) or on to the neighbouring brick (S - u);
// after every move the two alternate: next = s*S - current
__device__ __forceinline__ int brick_step(int parity, int s, int u, int S) {
  const int m = (parity != 0) == (s > 0) ? S - u : u;
  return s > 0 ? m : -m;
}

// what a burst ends with IS the lane's next state (no translation in the hot loop)
enum { MARCH_ON = ST_MARCH, MARCH_HIT = ST_SCATTER, MARCH_BOTTOM = ST_SURFACE, MARCH_TOP = ST_TOP };

// One burst of the marcher (accumulateExtinctionAlongPath, OPT:1697-1814): B cells.  The cells a
// ray visits and the lengths of its segments depend on geometry only, never on the extinction read,
// so the DDA runs B cells AHEAD (pure ALU; per axis one compare, one predicated index step and one
// predicated face-distance step), the B extinction gathers are issued together (B loads in flight
// per lane instead of one dependent load per cell), and only then is the optical depth accumulated
// and tested against the target (OPT:1729-1738).  Inside the burst the indices may run up to B cells
// into the ghost shell of the padded extinction field: periodic replicas in x and y (so no wrap test
// per cell, OPT:1782-1796), empty cells above the top and below the surface (so no exit test per
// cell, OPT:1801-1812, and no false hits there).  After the burst:
//   * the target fell inside cell k: the ray is put AT the event (r.t = distance where the target
//     optical depth is met, (ix,iy,iz) decoded from the saved padded cell) -> MARCH_HIT.  Face
//     distances tx/ty/tz are stale then; every caller starts a new leg there;
//   * the ray left through the top / the surface: r.t = distance to that boundary, (ix,iy) = column of
//     the exit point -> MARCH_TOP / MARCH_BOTTOM;
//   * otherwise x, y are folded back into the domain -> MARCH_ON.
//
// MASK (fields too large for L2): the geometry loop also fetches, per cell, the occupancy-bitmap word and the layer's
// clear-sky extinction (both L1/L2-resident); the gather of the big field is then issued only where the bit is set.
//
// BRICK: the field is read in its 2x2x2-brick layout (mcb_device.cuh).  The address is carried along incrementally: a
// move along an axis adds that axis' current step, and the step then alternates between "inside the brick" and "on to
// the next brick" -- two predicated integer instructions per axis instead of one, no multiplications.
//
// RAW (photon-pool kernel): a burst that ends at an event leaves the ray in its raw state -- r.t at the event, r.ix =
// the hit cell's padded address (hits), indices untouched (exits) -- and the caller decodes cell and position later,
// where all 32 lanes have an event to decode (mcb_pool.cu's event phase), instead of here with a third of the lanes.
// SPLIT: the gathers of the second half of the burst are issued only by lanes whose target was not met in the first
// half (one more dependent round trip per burst, fewer requests on the L1TEX -> L2 path that bounds the kernel).
//
// Vacuum-distance encoding (mcb_stage.cu, fields read without MASK): a cell without extinction holds -D, D = its
// Chebyshev distance in cells to the nearest cell that has some, so every gathered value is clamped at 0 before it is
// used.  vlast (photon-pool kernels): where the burst ends with MARCH_ON it receives a lower bound of -D for the cell
// the ray is in now (the last gathered value + 1: D changes by at most 1 between neighbours) -- march_leap's input.
// ENC = false (pool kernels on a domain that was staged without the encoding, DevDomain::leap == 0): no clamp.
//
// CROP (pool flux kernel, fields too large for L2; BRICK, no MASK): the layer-cropped field P.crp (mcb_device.cuh).  A
// cell inside the cropped layers is gathered from it like any cell of an L2-resident field -- address by arithmetic, ONE
// round trip per burst; a cell outside takes its layer's clear-sky value from the (L1-resident) layer table, and the two
// loads do not depend on each other.  A hit outside the crop hands back INT_MIN instead of an address.
template <bool REG, bool WIDE, int B, bool MASK, bool BRICK, bool RAW = false, bool SPLIT = false, bool ENC = true, bool CROP = false>
__device__ __forceinline__ int march_burst(Ray &r, const DevDomain &P, const Grid &G,
                                           float &ext, float target, unsigned &crossings, float *vlast = nullptr) {
  static_assert(!CROP || (RAW && REG && !MASK && BRICK), "layer-cropped field: pool flux kernel on uniform grids, bricked");
  const DevDomain::ExtField &F = CROP ? P.crp : BRICK ? P.brk : P.lin;
  float tE[B], sg[B];
  int ck[B];
  uint32_t mw[MASK ? B : 1];
  const float t0 = r.t;
  const int sx = r.dx >= 0.0f ? 1 : -1, sy = r.dy >= 0.0f ? 1 : -1, sz = r.dz >= 0.0f ? 1 : -1;
  int a = 0, dax = 0, day = 0, daz = 0, SX = 0, SY = 0, SZ = 0;
  if (BRICK) {
    a = brick_rel(F, r.ix, r.iy, CROP ? r.iz - P.cropLo : r.iz);          // (cropLo is even: parities carry over)
    SX = 8 * sx; SY = 2 * F.cY * sy; SZ = 2 * F.cZ * sz;
    dax = brick_step(r.ix & 1, sx, 1, 8); day = brick_step(r.iy & 1, sy, 2, 2 * F.cY); daz = brick_step(r.iz & 1, sz, 4, 2 * F.cZ);
  }
#pragma unroll
  for (int k = 0; k < B; ++k) {
    const float tmin = fminf(fminf(r.tx, r.ty), r.tz);
    ck[k] = BRICK ? a : r.ix + F.nxp * (r.iy + F.nyp * r.iz);
    if (CROP && (unsigned)(r.iz - P.cropLo) >= (unsigned)P.cropN) {           // outside the cropped layers: clear sky
      ck[k] = INT_MIN;
      sg[k] = fabsf(__ldg(P.layerExt + MCB_CHECK_INDEX(P, r.iz + GH, P.nz + 2 * GH)));
    }
    tE[k] = tmin;
    if (MASK) {
      // the layer's clear-sky extinction; its sign bit says that the whole layer has that value, so the bitmap word is
      // fetched only in layers that hold cloud somewhere (C5: 64 of 150 layers)
      const float lv = __ldg(P.layerExt + MCB_CHECK_INDEX(P, r.iz + GH, P.nz + 2 * GH));
      sg[k] = fabsf(lv);
      mw[k] = 0u;
      if (!signbit(lv)) mw[k] = __ldg(F.mask + MCB_CHECK_INDEX(P, (uint32_t)(ck[k] + F.origin) >> 5, (F.padded + 31) >> 5));
    }
    if (REG) {
      // one compare and predicated updates per axis, spelled out so that the index step is not widened into a
      // select followed by an add
      if (BRICK) {
#define MCB_STEP(T, I, S, H, RR, DA, SS) asm("{\n\t.reg .pred p;\n\t.reg .f32 q;\n\tsetp.le.f32 p, %0, %4;\n\tabs.f32 q, %7;\n\t" \
                                     "@p add.s32 %1, %1, %5;\n\t@p fma.rn.f32 %0, %6, q, %0;\n\t@p add.s32 %2, %2, %3;\n\t@p sub.s32 %3, %8, %3;\n\t}" \
                                     : "+f"(T), "+r"(I), "+r"(a), "+r"(DA) : "f"(tmin), "r"(S), "f"(H), "f"(RR), "r"(SS))
        MCB_STEP(r.tx, r.ix, sx, P.fhx, r.rx, dax, SX);
        MCB_STEP(r.ty, r.iy, sy, P.fhy, r.ry, day, SY);
        MCB_STEP(r.tz, r.iz, sz, P.fhz, r.rz, daz, SZ);
#undef MCB_STEP
      } else {
#define MCB_STEP(T, I, S, H, RR) asm("{\n\t.reg .pred p;\n\t.reg .f32 q;\n\tsetp.le.f32 p, %0, %2;\n\tabs.f32 q, %5;\n\t" \
                                     "@p add.s32 %1, %1, %3;\n\t@p fma.rn.f32 %0, %4, q, %0;\n\t}" \
                                     : "+f"(T), "+r"(I) : "f"(tmin), "r"(S), "f"(H), "f"(RR))
        MCB_STEP(r.tx, r.ix, sx, P.fhx, r.rx);
        MCB_STEP(r.ty, r.iy, sy, P.fhy, r.ry);
        MCB_STEP(r.tz, r.iz, sz, P.fhz, r.rz);
#undef MCB_STEP
      }
    } else {
      { const bool c = r.tx <= tmin; r.ix += c ? sx : 0; const float nt = (G.sx[r.ix + (sx > 0 ? 1 : 0)] - r.ox) * r.rx; r.tx = c ? nt : r.tx; }
      { const bool c = r.ty <= tmin; r.iy += c ? sy : 0; const float nt = (G.sy[r.iy + (sy > 0 ? 1 : 0)] - r.oy) * r.ry; r.ty = c ? nt : r.ty; }
      { const bool c = r.tz <= tmin; r.iz += c ? sz : 0; const float nt = (G.sz[r.iz + (sz > 0 ? 1 : 0)] - r.oz) * r.rz; r.tz = c ? nt : r.tz; }
    }
  }
  // accumulate until the target is passed (OPT:1729-1738); from there on acc / tS stay frozen at the
  // ENTRY of the hit cell, so only its extinction and index have to be carried along
  float acc = ext, tS = t0, hS = 1.0f;
  int hC = 0;
  int hK = B - 1;                                      // burst position of the hit cell
  bool found = false;
  constexpr int H = SPLIT ? B / 2 : B;                 // gathers issued up front
#pragma unroll
  for (int k = 0; k < H; ++k) {
    if (CROP) {
      if (ck[k] != INT_MIN) sg[k] = EXT_AT(P, F, ck[k]);
    } else if (MASK) {                                 // bit p of the bitmap: the shift count wraps modulo 32
      if (__funnelshift_r(mw[k], 0u, (uint32_t)(ck[k] + F.origin)) & 1u) sg[k] = EXT_AT(P, F, ck[k]);
    } else {
      sg[k] = EXT_AT(P, F, ck[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < B; ++k) {
    if (SPLIT && k == H) {                             // second half: only where the first half did not reach the target
#pragma unroll
      for (int j = H; j < B; ++j) {
        if (CROP) {
          if (ck[j] != INT_MIN) { sg[j] = 0.0f; if (!found) sg[j] = EXT_AT(P, F, ck[j]); }
        } else if (MASK) {
          if (!found && (__funnelshift_r(mw[j], 0u, (uint32_t)(ck[j] + F.origin)) & 1u)) sg[j] = EXT_AT(P, F, ck[j]);
        } else {
          sg[j] = 0.0f;
          if (!found) sg[j] = EXT_AT(P, F, ck[j]);
        }
      }
    }
#ifdef MCB_NO_LEAP                                            // A/B build without vacuum leaps (make variant FLAGS=-DMCB_NO_LEAP)
    const float sk = sg[k];
#else
    const float sk = (MASK || CROP || !ENC) ? sg[k] : fmaxf(sg[k], 0.0f);       // vacuum cells hold -D
#endif
    const float en = fmaf(tE[k] - tS, sk, acc);
    const bool h = !found && en > target;
    hS = h ? sk : hS; hC = h ? ck[k] : hC; hK = h ? k : hK;
    found = found || h;
    acc = found ? acc : en; tS = found ? tS : tE[k];
  }
  if (found) {
    crossings += (unsigned)(hK + 1);               // cells entered up to and including the hit cell
    r.t = tS + __fdividef(target - acc, hS);           // where the target optical depth is met (OPT:1731)
    if (RAW) r.ix = hC; else cell_decode<WIDE, BRICK>(P, hC, r.ix, r.iy, r.iz);
    return MARCH_HIT;
  }
  ext = acc;
  if ((unsigned)r.iz >= (unsigned)P.nz) {              // left the domain during this burst
    const bool top = r.iz > 0;
    const float tX = ((top ? P.fzMax : P.fz0) - r.oz) * r.rz;
    unsigned nv = 1u;                                  // cells entered before the boundary was reached
#pragma unroll
    for (int k = 1; k < B; ++k) nv += tE[k - 1] < tX ? 1u : 0u;
    crossings += nv;
    r.t = tX;
    if (RAW) return top ? MARCH_TOP : MARCH_BOTTOM;
    float px, py, pz;
    ray_position(r, P, px, py, pz);
    if (REG) {
      r.ix = min(max((int)((px - P.fx0) * P.finvhx), 0), P.nx - 1);
      r.iy = min(max((int)((py - P.fy0) * P.finvhy), 0), P.ny - 1);
    } else {
      r.ix = find_cell(G.sx, P.nx, px);
      r.iy = find_cell(G.sy, P.ny, py);
    }
    r.iz = top ? P.nz - 1 : 0;
    return top ? MARCH_TOP : MARCH_BOTTOM;
  }
  crossings += (unsigned)B;
  r.t = tS;
  if (vlast) *vlast = (MASK || CROP) ? __ldg(P.layerLeap + r.iz + GH) : sg[B - 1] + 1.0f;
  if (REG) {
    r.ix = wrap_index<WIDE>(r.ix, P.nx);
    r.iy = wrap_index<WIDE>(r.iy, P.ny);
  } else {                                             // keep (edge - origin) invariant under the fold
    while (r.ix < 0) { r.ix += P.nx; r.ox += P.fLx; }
    while (r.ix >= P.nx) { r.ix -= P.nx; r.ox -= P.fLx; }
    while (r.iy < 0) { r.iy += P.ny; r.oy += P.fLy; }
    while (r.iy >= P.ny) { r.iy -= P.ny; r.oy -= P.fLy; }
  }
  return MARCH_ON;
}

// One leap through vacuum (photon-pool kernels, uniform grids): the ray is in a cell whose vacuum distance is at least
// D (mcb_stage.cu: every cell within D - 1 cells of it, Chebyshev, has no extinction; the caller has limited D by the
// number of layers up to the boundary the ray is heading for -- leap_distance -- so the cube ends on that boundary at
// the latest), so it can go straight to where it leaves that cube -- D faces away along one axis -- without looking at
// a single cell: no optical depth accumulates on the way (OPT:1729-1738 adds 0 per cell).  The DDA state is advanced by
// whole faces (face distances stay on the same lattice tx0 + n * step as the cell-by-cell walk) and the cells passed are
// counted (one per face crossed: the cell the ray is in now up to the last one before the landing cell, exactly what
// bursts over the same path would have counted).  Landing on the top / the surface ends the leg like a burst does
// (RAW); otherwise the caller marches on from the landing cell with a burst IN THE SAME ITERATION, whose last gather
// says whether the next iteration starts with a leap again.  (Measured, r02: a leap that took the place of the burst --
// with one gather at the landing cell to chain leaps -- covered 42 % of the C3 crossings and gained nothing: its
// divergent code and the extra dependent round trip per iteration cost what the saved bursts had bought.)
// how far the ray may leap from its cell (0: march): v = what is known about the cell (march_burst's vlast)
__device__ __forceinline__ int leap_distance(const Ray &r, const DevDomain &P, float v, float leapBelow) {
#ifdef MCB_NO_LEAP
  return 0;
#endif
  if (!(v <= leapBelow)) return 0;
  const int D = min((int)(-v), r.dz >= 0.0f ? P.nz - r.iz : r.iz + 1);
  return D >= 2 ? D : 0;
}
//
// MASK (fields read through the occupancy bitmap, whose clear sky is not vacuum): D counts layers that are clear
// throughout (layerLeap), the optical depth of the leap is that of the horizontally uniform clear sky (layerExt,
// layerCum) -- and if the target falls inside it the leap is not taken (the burst that follows finds the event).
// stat: two shared-memory counters (leaps taken, cells they crossed), added to once per warp (MCB_LEAP_STATS builds).
template <bool MASK>
__device__ __forceinline__ int march_leap(Ray &r, const DevDomain &P, const int D, unsigned &crossings,
                                          float &ext, const float target, unsigned *stat) {
  const float k = (float)(D - 1);
  const float ax = fabsf(r.rx) * P.fhx, ay = fabsf(r.ry) * P.fhy, az = fabsf(r.rz) * P.fhz;   // path between two faces (inf: axis not moving)
  const float ex = fmaf(k, ax, r.tx), ey = fmaf(k, ay, r.ty), ez = fmaf(k, az, r.tz);          // where the cube ends
  const float te = fminf(fminf(ex, ey), ez);
  // faces crossed on the way: all D along the axis the ray leaves through, fewer along the others
  // (selects, not branches: the lanes in here are few enough as it is; an axis that does not move has t = FLT_MAX,
  // its quotient converts to a large negative integer and is discarded)
  int nx = ex == te ? D : min(__float2int_rz((te - r.tx) * (fabsf(r.dx) * P.finvhx)) + 1, D);
  int ny = ey == te ? D : min(__float2int_rz((te - r.ty) * (fabsf(r.dy) * P.finvhy)) + 1, D);
  int nz = ez == te ? D : min(__float2int_rz((te - r.tz) * (fabsf(r.dz) * P.finvhz)) + 1, D);
  nx = r.tx <= te ? nx : 0;
  ny = r.ty <= te ? ny : 0;
  nz = r.tz <= te ? nz : 0;
  if (MASK) {                                                // clear-sky optical depth from r.t to te
    const float *L = P.layerExt + GH;
    const bool up = r.dz >= 0.0f;
    float tauL = fabsf(__ldg(L + r.iz));
    if (nz == 0) {
      tauL *= te - r.t;
    } else {                                                 // rest of this layer + whole layers + the part of the landing layer
      const int izN = r.iz + (up ? nz : -nz);                // (the ghost layer outside the domain has no extinction)
      const float c0 = __ldg(P.layerCum + (up ? r.iz + 1 : r.iz)), c1 = __ldg(P.layerCum + (up ? izN : izN + 1));
      tauL = tauL * (r.tz - r.t) + fabsf(c1 - c0) * fabsf(r.rz) + fabsf(__ldg(L + izN)) * (te - fmaf((float)(nz - 1), az, r.tz));
    }
    if (ext + tauL > target) return MARCH_ON;
    ext += tauL;
  }
  if (nx) { r.tx = fmaf((float)nx, ax, r.tx); r.ix += r.dx >= 0.0f ? nx : -nx; }
  if (ny) { r.ty = fmaf((float)ny, ay, r.ty); r.iy += r.dy >= 0.0f ? ny : -ny; }
  if (nz) { r.tz = fmaf((float)nz, az, r.tz); r.iz += r.dz >= 0.0f ? nz : -nz; }
  crossings += (unsigned)(nx + ny + nz);
#ifdef MCB_LEAP_STATS                                          // counters leaps / leapCells: the bounds-checked build only (the
  {                                                           // reduction costs the 72-register flux kernel 9 % on C3, r02)
    const unsigned lanes = __activemask();
    const unsigned cells = __reduce_add_sync(lanes, (unsigned)(nx + ny + nz));
    if ((threadIdx.x & 31) == __ffs(lanes) - 1) { atomicAdd(&stat[0], (unsigned)__popc(lanes)); atomicAdd(&stat[1], cells); }
  }
#endif
  if ((unsigned)r.iz >= (unsigned)P.nz) {                    // landed on the boundary: the leg ends there
    const bool top = r.iz > 0;
    r.t = ((top ? P.fzMax : P.fz0) - r.oz) * r.rz;
    return top ? MARCH_TOP : MARCH_BOTTOM;
  }
  r.t = te;
  r.ix = wrap_index<true>(r.ix, P.nx);
  r.iy = wrap_index<true>(r.iy, P.ny);
  return MARCH_ON;
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

// ---- tallies ------------------------------------------------------------------------------
struct Tally {
  float *sFlux, *sVol, *sInt;       // shared-memory privatised copies (or nullptr)
  int cols;
};
__device__ __forceinline__ void add_flux(const DevDomain &P, const Tally &T, int which, int col, float v) {
  col = (int)MCB_CHECK_INDEX(P, col, T.cols);
  if (T.sFlux) atomicAdd(&T.sFlux[which * T.cols + col], v);
  else atomicAdd(&P.tally[(which == 0 ? P.offFluxUp : which == 1 ? P.offFluxDown : P.offFluxAbs) + col], (double)v);
}
__device__ __forceinline__ void add_vol(const DevDomain &P, const Tally &T, int cell, float v) {
  cell = (int)MCB_CHECK_INDEX(P, cell, (long long)T.cols * P.nz);
  if (T.sVol) atomicAdd(&T.sVol[cell], v);
  else atomicAdd(&P.tally[P.offVolAbs + cell], (double)v);
}
__device__ __forceinline__ void add_intensity(const DevDomain &P, const Tally &T, int dir, int col, int comp, float v) {
  col = (int)MCB_CHECK_INDEX(P, col, T.cols); dir = (int)MCB_CHECK_INDEX(P, dir, P.nDir); comp = (int)MCB_CHECK_INDEX(P, comp, P.nc + 1);
  if (T.sInt) atomicAdd(&T.sInt[dir * T.cols + col], v);
  else atomicAdd(&P.tally[P.offInt + col + (long long)T.cols * dir], (double)v);
  atomicAdd(&P.tally[P.offIntByComp + col + (long long)T.cols * (dir + (long long)P.nDir * comp)], (double)v);
}

struct Counts { unsigned crossings, scatters, leRays, leCrossings; };

}  // namespace mcbfast
