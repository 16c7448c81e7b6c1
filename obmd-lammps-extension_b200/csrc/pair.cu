// PairDPD::compute on the cell grid (replaces NStencilBin::create + NPairBin::build +
// PairDPD::compute, pair_dpd.cpp:66-163, and Pair::virial_fdotr_compute, pair.cpp:1735).
//
// Full-neighbour evaluation: every atom walks its 27 cells and accumulates only its own
// force, so there are no atomics, no reverse halo, and the result is independent of launch
// geometry.  The pair expression follows pair_dpd.cpp:117-146:
//   wd = 1 - r/rc;  fpair = ((a0*wd - gamma*wd^2*(dr.dv)/r) + sigma*wd*xi/sqrt(dt)) / r
// The sequential RanMars stream (pair_dpd.cpp:122) cannot be reproduced in parallel; xi is a
// symmetric counter-based N(0,1) variate keyed on (min tag, max tag, seed, timestep), identical
// for the (i,j) and (j,i) evaluations so Newton's third law holds exactly (DESIGN.md "RNG").
//
// Kernels in this translation unit:
//   * k_pair_brick (pair_brick.cuh): the production path for the cell-sorted atoms -- bricks of
//     C x R x zc cells whose window the copy engine stages (cp.async.bulk + mbarrier) from the 48 + 16-byte
//     atom records the sort wrote, fp32 screen, fp64 evaluation;
//   * k_pair_row (pair_row.cuh): the thread-staged row tile of round 1; it folds the per-cell lists of
//     appended atoms (slab arrivals, ghosts, insertions) into its window, so it takes the columns whose
//     stencil holds such atoms, and everything when the records are stale;
//   * k_pair_dpd: per-atom kernel that walks global memory (L1/L2) with a minimum-image test: a
//     distance screen over the 9 stencil rows (the 3 z-cells of a row are one contiguous slot range
//     because z is the fastest cell index) that pushes the slots inside the cutoff to a per-thread
//     list, then a dense evaluation of the list.  Used for energy/virial tallies, for the few atoms
//     appended after the sort, as the overflow path of the tiled kernels, and under OBMD_B200_TILE=0;
//   * k_pair_new_atoms / k_pair_add_new: the incremental update after USHER insertions.
//
// Roofline: algorithmic traffic is 80 B/atom (x,v,tag,type in, f out); neighbour reads are served by
// shared memory / L2.  The kernels are issue/latency bound in fp64, not HBM bound.
#include "ctx.h"

#define EPSILON_DPD 1.0e-10  // pair_dpd.cpp:36
#define PAIR_T 128           // threads per block
#define PAIR_HCAP 32         // listed hits per thread (mean 12.6 at rho = 3); overflow is evaluated inline

__device__ __forceinline__ double dpd_gaussian(int tagi, int tagj, unsigned step, unsigned seed)
{
  unsigned lo = (unsigned) min(tagi, tagj), hi = (unsigned) max(tagi, tagj);
  unsigned long long key = ((unsigned long long) lo << 32) | hi;
  unsigned long long ctr = ((unsigned long long) seed << 32) | step;
  unsigned long long z = key * 0x9E3779B97F4A7C15ULL + ctr * 0xD1B54A32D192ED03ULL;
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  unsigned h = (unsigned) (z >> 32), l = (unsigned) z;
  float u1 = ((float) h + 1.0f) * 2.3283064365386963e-10f;
  float u2 = (float) l * 2.3283064365386963e-10f;
  float r;  // MUFU.SQRT: the default IEEE sqrtf costs a Newton fix-up the Box-Muller radius does not need
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
  return (double) (r * __cosf(6.283185307179586f * u2));
}

struct PairAcc {
  double fx, fy, fz, e, v0, v1, v2, v3, v4, v5;
};

struct AtomI {
  double x, y, z, vx, vy, vz;
  int tag, type;
  bool bnd;  // sits in a boundary cell in y or z: needs the minimum-image shift
};

// Domain::pbc (domain.cpp:529-640) applied on the fly, with exactly the arithmetic of k_pbc: the
// forces are then the same whether or not the stored coordinates have been wrapped yet, which is
// what lets the pair kernel run on the side stream while pre_exchange still needs the
// unwrapped positions (USHER sums must see the reference's pre-pbc coordinates).
__device__ __forceinline__ double wrap1(double x, double lo, double hi, double prd)
{
  if (x < lo) x += prd;
  if (x >= hi) {
    x -= prd;
    x = fmax(x, lo);
  }
  return x;
}

__device__ __forceinline__ void min_image(const Grid &g, double &dely, double &delz)
{
  if (dely > g.half[1]) dely -= g.prd[1];
  else if (dely < -g.half[1]) dely += g.prd[1];
  if (delz > g.half[2]) delz -= g.prd[2];
  else if (delz < -g.half[2]) delz += g.prd[2];
}

// pair_dpd.cpp:117-146 for a pair already known to be inside the cutoff
template <bool SINGLE_TYPE, bool TALLY, bool PREWRAP>
__device__ __forceinline__ void pair_eval_xyz(const Planes &p, const PairTab &pt, const Grid &g, int j,
                                              double xj, double yj, double zj, const AtomI &a, unsigned step,
                                              bool random_on, PairAcc &acc);

template <bool SINGLE_TYPE, bool TALLY, bool PREWRAP>
__device__ __forceinline__ void pair_eval(const Planes &p, const PairTab &pt, const Grid &g, int j,
                                          const AtomI &a, unsigned step, bool random_on, PairAcc &acc)
{
  pair_eval_xyz<SINGLE_TYPE, TALLY, PREWRAP>(p, pt, g, j, p.x[j], p.y[j], p.z[j], a, step, random_on, acc);
}

// the neighbour's position is passed in (global memory or the tile kernel's shared-memory copy)
template <bool SINGLE_TYPE, bool TALLY, bool PREWRAP>
__device__ __forceinline__ void pair_eval_xyz(const Planes &p, const PairTab &pt, const Grid &g, int j,
                                              double xj, double yj, double zj, const AtomI &a, unsigned step,
                                              bool random_on, PairAcc &acc)
{
  double delx = a.x - xj, dely, delz;
  if (a.bnd) {
    dely = a.y - (PREWRAP ? wrap1(yj, g.lo[1], g.hi[1], g.prd[1]) : yj);
    delz = a.z - (PREWRAP ? wrap1(zj, g.lo[2], g.hi[2], g.prd[2]) : zj);
    min_image(g, dely, delz);
  } else {
    dely = a.y - yj;
    delz = a.z - zj;
  }
  double rsq = delx * delx + dely * dely + delz * delz;
  int tj = SINGLE_TYPE ? 1 : p.type[j];
  // one rsqrt instead of sqrt + two divisions (pair_dpd.cpp:118-121 computes r = sqrt(rsq),
  // rinv = 1/r, wd = 1 - r/cut): same values to ~2 ulp, well inside the 1e-9 force tolerance
  if (rsq < EPSILON_DPD * EPSILON_DPD) return;
  double rinv = rsqrt(rsq);
  double r = rsq * rinv;
  double delvx = a.vx - p.vx[j], delvy = a.vy - p.vy[j], delvz = a.vz - p.vz[j];
  double dot = delx * delvx + dely * delvy + delz * delvz;
  double wd = 1.0 - r * pt.cutinv[a.type][tj];
  double fpair = pt.a0[a.type][tj] * wd;
  fpair -= pt.gamma[a.type][tj] * wd * wd * dot * rinv;
  if (random_on) {
    double randnum = dpd_gaussian(a.tag, p.tag[j], step, pt.seed);
    fpair += pt.sigma[a.type][tj] * wd * randnum * pt.dtinvsqrt;
  }
  fpair *= rinv;
  acc.fx += delx * fpair;
  acc.fy += dely * fpair;
  acc.fz += delz * fpair;
  if (TALLY) {
    acc.e += 0.5 * (0.5 * pt.a0[a.type][tj] * pt.cut[a.type][tj] * wd * wd);
    double h = 0.5 * fpair;
    acc.v0 += h * delx * delx;
    acc.v1 += h * dely * dely;
    acc.v2 += h * delz * delz;
    acc.v3 += h * delx * dely;
    acc.v4 += h * delx * delz;
    acc.v5 += h * dely * delz;
  }
}

template <bool SINGLE_TYPE, bool PREWRAP>
__device__ __forceinline__ bool pair_screen(const Planes &p, const PairTab &pt, const Grid &g, int j,
                                            const AtomI &a)
{
  double delx = a.x - p.x[j], dely, delz;
  if (a.bnd) {
    dely = a.y - (PREWRAP ? wrap1(p.y[j], g.lo[1], g.hi[1], g.prd[1]) : p.y[j]);
    delz = a.z - (PREWRAP ? wrap1(p.z[j], g.lo[2], g.hi[2], g.prd[2]) : p.z[j]);
    min_image(g, dely, delz);
  } else {
    dely = a.y - p.y[j];
    delz = a.z - p.z[j];
  }
  double rsq = delx * delx + dely * dely + delz * delz;
  int tj = SINGLE_TYPE ? 1 : p.type[j];
  return rsq < pt.cutsq[a.type][tj];
}

// slot of the g-th atom of a column set (intervals in order, then the inserted atoms); -1 if none
__device__ __forceinline__ int colrange_slot(const ColRange &cr, int gidx, const Grid &g,
                                             const int *__restrict__ cell_start, const Extras &ex)
{
  const int nyz = g.ny * g.nz;
  for (int k = 0; k < cr.n; k++) {
    int b = cell_start[cr.a[k] * nyz], e = cell_start[(cr.b[k] + 1) * nyz];
    if (gidx < e - b) return b + gidx;
    gidx -= e - b;
  }
  return cr.extras ? extras_slot(ex, cr.extras, gidx) : -1;
}

// ---- one atom against its stencil, straight from global memory (L1/L2) ------------------------
// Used for explicit slot lists (recompute near insertions), the atoms appended after the sort, the
// tally variant, and as the fallback of the tiled kernel.  The order in which contributions are
// accumulated is part of the contract (the tiled kernel reproduces it bit for bit): rows (ox, oy)
// ascending; inside a row the cells cz-1, cz, cz+1 in that (wrapped) order; hits beyond the list
// capacity and atoms from the per-cell extra lists are added on the spot, listed hits at the end.
template <bool SINGLE_TYPE, bool TALLY, bool PREWRAP, typename HitT>
__device__ __forceinline__ void pair_atom_global(const Planes &p, const Grid &g, const PairTab &pt,
                                                 const Extras &ex, const int *__restrict__ cell_start,
                                                 const int *__restrict__ extra_head,
                                                 const int *__restrict__ extra_next, int i, unsigned live,
                                                 unsigned step, bool random_on, HitT *myhits, int hstride,
                                                 PairAcc &acc)
{
  const int nextra = ex.narr + ex.nghost + ex.nins;
  AtomI a;
  a.x = p.x[i];
  a.y = PREWRAP ? wrap1(p.y[i], g.lo[1], g.hi[1], g.prd[1]) : p.y[i];
  a.z = PREWRAP ? wrap1(p.z[i], g.lo[2], g.hi[2], g.prd[2]) : p.z[i];
  a.vx = p.vx[i]; a.vy = p.vy[i]; a.vz = p.vz[i];
  a.tag = p.tag[i];
  a.type = SINGLE_TYPE ? 1 : p.type[i];
  const int cx = cell_cx(g, a.x), cy = cell_cy(g, a.y), cz = cell_cz(g, a.z);
  // minimum image for atoms in a face layer; before Domain::pbc has run (PREWRAP) also for the
  // second layer, because a face-layer neighbour may still carry unwrapped coordinates.  The
  // choice is made per warp so the screen loop does not diverge (both paths give the same
  // numbers for interior atoms: wrap and shift are no-ops there).
  const int lay = PREWRAP ? 1 : 0;
  a.bnd = (cy <= lay || cy >= g.ny - 1 - lay || cz <= lay || cz >= g.nz - 1 - lay);
  a.bnd = __any_sync(live, a.bnd);
  // the three stencil cells of a row in window order; away from the faces they are one segment
  int segb[3], sege[3], nseg;
  int nh = 0;
  for (int ox = -1; ox <= 1; ox++) {
    const int bx = cx + ox;
    if (bx < 0 || bx >= g.nx) continue;
    for (int oy = -1; oy <= 1; oy++) {
      int by = cy + oy;
      by = by < 0 ? by + g.ny : (by >= g.ny ? by - g.ny : by);
      const int row = (bx * g.ny + by) * g.nz;
      if (cz == 0) {
        nseg = 2;
        segb[0] = cell_start[row + g.nz - 1]; sege[0] = cell_start[row + g.nz];
        segb[1] = cell_start[row]; sege[1] = cell_start[row + 2];
      } else if (cz == g.nz - 1) {
        nseg = 2;
        segb[0] = cell_start[row + cz - 1]; sege[0] = cell_start[row + cz + 1];
        segb[1] = cell_start[row]; sege[1] = cell_start[row + 1];
      } else {
        nseg = 1;
        segb[0] = cell_start[row + cz - 1]; sege[0] = cell_start[row + cz + 2];
      }
      for (int sgi = 0; sgi < nseg; sgi++) {
        const int jb = segb[sgi], je = sege[sgi];
        // four candidates per trip: the twelve coordinate loads are issued back to back
        for (int j0 = jb; j0 < je; j0 += 4) {
          bool h[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int j = min(j0 + u, je - 1);  // clamped lanes repeat the last atom, masked below
            h[u] = pair_screen<SINGLE_TYPE, PREWRAP>(p, pt, g, j, a) && (j0 + u < je) && (j != i);
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (h[u]) {
              if (nh < PAIR_HCAP) {
                myhits[nh * hstride] = (HitT) (j0 + u - i + 32768);  // slot offset from i, biased
                nh++;
              } else {
                pair_eval<SINGLE_TYPE, TALLY, PREWRAP>(p, pt, g, j0 + u, a, step, random_on, acc);
              }
            }
          }
        }
      }
      if (nextra > 0) {  // atoms appended after the sort hang off per-cell lists
        for (int oz = -1; oz <= 1; oz++) {
          int bz = cz + oz;
          bz = bz < 0 ? bz + g.nz : (bz >= g.nz ? bz - g.nz : bz);
          for (int e = extra_head[row + bz]; e >= 0; e = extra_next[e]) {
            int j = ex.base + e;
            if (j == i) continue;
            if (pair_screen<SINGLE_TYPE, PREWRAP>(p, pt, g, j, a))
              pair_eval<SINGLE_TYPE, TALLY, PREWRAP>(p, pt, g, j, a, step, random_on, acc);
          }
        }
      }
    }
  }
  for (int k = 0; k < nh; k++)
    pair_eval<SINGLE_TYPE, TALLY, PREWRAP>(p, pt, g, (int) myhits[k * hstride] - 32768 + i, a, step,
                                           random_on, acc);
}

template <bool SINGLE_TYPE, bool TALLY, bool PREWRAP>
__global__ void __launch_bounds__(PAIR_T)
    k_pair_dpd(Planes p, ColRange cr, Extras ex, Grid g, PairTab pt,
               const int *__restrict__ cell_start, const int *__restrict__ extra_head,
               const int *__restrict__ extra_next, double *__restrict__ fx, double *__restrict__ fy,
               double *__restrict__ fz, unsigned step, int random_on, double *__restrict__ partials,
               double *__restrict__ vatom)
{
  __shared__ int hits[PAIR_HCAP * PAIR_T];
  const int i = colrange_slot(cr, blockIdx.x * PAIR_T + threadIdx.x, g, cell_start, ex);
  PairAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const unsigned live = __ballot_sync(0xffffffffu, i >= 0);
  if (i >= 0) {
    pair_atom_global<SINGLE_TYPE, TALLY, PREWRAP, int>(p, g, pt, ex, cell_start, extra_head, extra_next, i,
                                                       live, step, random_on != 0, hits + threadIdx.x,
                                                       PAIR_T, acc);
    fx[i] = acc.fx;
    fy[i] = acc.fy;
    fz[i] = acc.fz;
    // per-atom virial xx (what `compute stress/atom` gets from ev_tally: half of every pair term)
    if (TALLY && vatom) vatom[i] = acc.v0;
  }
  if (TALLY) {
    // deterministic block reduction of (e, v[6]) -> partials[block][7]
    __shared__ double sh[7][PAIR_T / 32];
    double v[7] = {acc.e, acc.v0, acc.v1, acc.v2, acc.v3, acc.v4, acc.v5};
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 7; k++) {
      double s = v[k];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
      if (lane == 0) sh[k][w] = s;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
      double s = 0.0;
      for (int ww = 0; ww < PAIR_T / 32; ww++) s += sh[threadIdx.x][ww];
      partials[blockIdx.x * 7 + threadIdx.x] = s;
    }
  }
}

#include "pair_row.cuh"
#include "pair_brick.cuh"

// ---- forces after USHER insertions, without recomputing anybody from scratch -----------------
// obmd_run computes PairDPD on the side stream while USHER is still searching, i.e. without the
// atoms it is about to create.  Afterwards (a) every new atom gets its full force, one warp per
// atom: lanes 0..26 own one stencil cell each and the partial sums are combined by a fixed
// shuffle tree; (b) every old atom in the 27 cells around an insertion (the list k_mark_redo
// builds) adds the terms of the new atoms in its stencil, in per-cell list order.  Both orders
// are functions of the data layout only, so the result is deterministic.
template <bool SINGLE_TYPE>
__global__ void __launch_bounds__(128)
    k_pair_new_atoms(Planes p, Extras ex, Grid g, PairTab pt, const int *__restrict__ cell_start,
                     const int *__restrict__ extra_head, const int *__restrict__ extra_next,
                     double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz, unsigned step,
                     int random_on)
{
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= ex.nins) return;
  const int i = ex.base + ex.narr + ex.nghost + k;
  AtomI a;
  a.x = p.x[i]; a.y = p.y[i]; a.z = p.z[i];
  a.vx = p.vx[i]; a.vy = p.vy[i]; a.vz = p.vz[i];
  a.tag = p.tag[i];
  a.type = SINGLE_TYPE ? 1 : p.type[i];
  a.bnd = true;
  PairAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < 27) {
    const int bx = cell_cx(g, a.x) + lane / 9 - 1;
    if (bx >= 0 && bx < g.nx) {
      int by = cell_cy(g, a.y) + (lane / 3) % 3 - 1, bz = cell_cz(g, a.z) + lane % 3 - 1;
      by = by < 0 ? by + g.ny : (by >= g.ny ? by - g.ny : by);
      bz = bz < 0 ? bz + g.nz : (bz >= g.nz ? bz - g.nz : bz);
      const int cc = (bx * g.ny + by) * g.nz + bz;
      for (int j = cell_start[cc]; j < cell_start[cc + 1]; j++)
        if (pair_screen<SINGLE_TYPE, false>(p, pt, g, j, a))
          pair_eval<SINGLE_TYPE, false, false>(p, pt, g, j, a, step, random_on != 0, acc);
      for (int e = extra_head[cc]; e >= 0; e = extra_next[e]) {
        const int j = ex.base + e;
        if (j != i && pair_screen<SINGLE_TYPE, false>(p, pt, g, j, a))
          pair_eval<SINGLE_TYPE, false, false>(p, pt, g, j, a, step, random_on != 0, acc);
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    acc.fx += __shfl_down_sync(0xffffffffu, acc.fx, d);
    acc.fy += __shfl_down_sync(0xffffffffu, acc.fy, d);
    acc.fz += __shfl_down_sync(0xffffffffu, acc.fz, d);
  }
  if (lane == 0) {
    fx[i] = acc.fx;
    fy[i] = acc.fy;
    fz[i] = acc.fz;
  }
}

template <bool SINGLE_TYPE>
__global__ void __launch_bounds__(128)
    k_pair_add_new(Planes p, Extras ex, Grid g, PairTab pt, const int *__restrict__ extra_head,
                   const int *__restrict__ extra_next, double *__restrict__ fx, double *__restrict__ fy,
                   double *__restrict__ fz, unsigned step, int random_on, const int *__restrict__ list,
                   const int *__restrict__ count, int cap)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= min(*count, cap)) return;
  const int i = list[q], first_new = ex.narr + ex.nghost;
  if (i >= ex.base + first_new) return;  // the new atoms were done by k_pair_new_atoms
  AtomI a;
  a.x = p.x[i]; a.y = p.y[i]; a.z = p.z[i];
  a.vx = p.vx[i]; a.vy = p.vy[i]; a.vz = p.vz[i];
  a.tag = p.tag[i];
  a.type = SINGLE_TYPE ? 1 : p.type[i];
  a.bnd = true;
  const int cx = cell_cx(g, a.x), cy = cell_cy(g, a.y), cz = cell_cz(g, a.z);
  PairAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // all 27 list heads first (independent loads; almost all of them are empty), then the few lists
  int head[27];
#pragma unroll
  for (int nb = 0; nb < 27; nb++) {
    const int bx = cx + nb / 9 - 1;
    int by = cy + (nb / 3) % 3 - 1, bz = cz + nb % 3 - 1;
    by = by < 0 ? by + g.ny : (by >= g.ny ? by - g.ny : by);
    bz = bz < 0 ? bz + g.nz : (bz >= g.nz ? bz - g.nz : bz);
    head[nb] = (bx < 0 || bx >= g.nx) ? -1 : extra_head[(bx * g.ny + by) * g.nz + bz];
  }
#pragma unroll
  for (int nb = 0; nb < 27; nb++)
    for (int e = head[nb]; e >= 0; e = extra_next[e])
      if (e >= first_new && pair_screen<SINGLE_TYPE, false>(p, pt, g, ex.base + e, a))
        pair_eval<SINGLE_TYPE, false, false>(p, pt, g, ex.base + e, a, step, random_on != 0, acc);
  fx[i] += acc.fx;
  fy[i] += acc.fy;
  fz[i] += acc.fz;
}

int pair_add_insertions(obmd_ctx *c, long long step, const int *list, const int *count, int cap)
{
  ScopedTimer t(c, "pair");
  if (c->nins == 0) return OBMD_OK;
  const bool single = (c->ntypes == 1);
  const int random_on = c->temperature > 0.0 ? 1 : 0;
  c->pt.dtinvsqrt = 1.0 / sqrt(c->dt);
  const Extras ex = extras_of(c);
  if (single) {
    LAUNCH(k_pair_new_atoms<true>, cdiv(c->nins, 4), 128, 0, c->cur, ex, c->grid, c->pt, c->cell_start,
           c->extra_head, c->extra_next, c->fx, c->fy, c->fz, (unsigned) step, random_on);
    LAUNCH(k_pair_add_new<true>, cdiv(cap, 128), 128, 0, c->cur, ex, c->grid, c->pt, c->extra_head,
           c->extra_next, c->fx, c->fy, c->fz, (unsigned) step, random_on, list, count, cap);
  } else {
    LAUNCH(k_pair_new_atoms<false>, cdiv(c->nins, 4), 128, 0, c->cur, ex, c->grid, c->pt, c->cell_start,
           c->extra_head, c->extra_next, c->fx, c->fy, c->fz, (unsigned) step, random_on);
    LAUNCH(k_pair_add_new<false>, cdiv(cap, 128), 128, 0, c->cur, ex, c->grid, c->pt, c->extra_head,
           c->extra_next, c->fx, c->fy, c->fz, (unsigned) step, random_on, list, count, cap);
  }
  CK(cudaGetLastError());
  return OBMD_OK;
}

// Fixed-order final reduction of block partials (one block).
__global__ void k_pair_tally_final(const double *__restrict__ partials, int nblocks, StepScalars *sc)
{
  __shared__ double sh[256];
  for (int k = 0; k < 7; k++) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partials[b * 7 + k];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = blockDim.x / 2; d > 0; d >>= 1) {
      if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (k == 0) sc->eng = sh[0];
      else sc->virial[k - 1] = sh[0];
    }
    __syncthreads();
  }
}

// diagnostic: the counter-based N(0,1) variate of pairs (tag0 + k, tag0 + k + 1) at one timestep
__global__ void k_gaussian_sample(int n, int tag0, unsigned step, unsigned seed, double *__restrict__ out)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = dpd_gaussian(tag0 + k, tag0 + k + 1, step, seed);
}

int pair_gaussian_sample(obmd_ctx *c, int n, int tag0, long long step, double *out)
{
  double *d = nullptr;
  CK(cudaMalloc(&d, sizeof(double) * n));
  LAUNCH(k_gaussian_sample, cdiv(n, 256), 256, 0, n, tag0, (unsigned) step, (unsigned) c->pair_seed, d);
  CK(cudaMemcpyAsync(out, d, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return OBMD_OK;
}

ColRange all_columns(const obmd_ctx *c)
{
  ColRange cr;
  cr.n = 1;
  cr.a[0] = 0;
  cr.b[0] = c->grid.nx - 1;  // ghost columns hold no sorted atoms, so this is every sorted atom
  cr.a[1] = cr.b[1] = 0;
  cr.extras = 1;
  return cr;
}

// upper bound of the atoms in a column set, for sizing the grid without a device read-back
static int colrange_bound(const obmd_ctx *c, const ColRange &cr)
{
  long long cols = 0;
  for (int k = 0; k < cr.n; k++) cols += cr.b[k] - cr.a[k] + 1;
  (void) cols;
  // no per-column counts on the host: launch for every sorted atom, surplus threads exit at once
  return c->nsorted + (cr.extras ? c->nextra : 0);
}

// Tile shape of the row kernel for this box and density: rows per tile R (1, 2, 4), z-cells per tile
// zc, threads T, staged-atom capacity, ranked by a small cost model fitted to measurements;
// OBMD_B200_ROWS / _ZC / _PT override.
struct RowPlan {
  int R, nchunk, zc, T, cap;
  size_t smem;
};

static int pair_row_plan(obmd_ctx *c, bool single, RowPlan &best)
{
  const Grid &g = c->grid;
  const double rho = (double) c->nsorted / ((double) (g.own_b - g.own_a + 1) * g.ny * g.nz);
  double best_score = -1.0;
  for (int R = 1; R <= 4; R *= 2) {
    if (c->row_rows > 0 && R != c->row_rows) continue;
    if (R > g.ny) continue;
    const int NR = 3 * (R + 2);
    for (int nchunk = 1; nchunk <= g.nz; nchunk++) {
      const int zc = cdiv(g.nz, nchunk);
      if (zc > PR_MAXZC) continue;
      if (c->row_zc > 0 && zc != cdiv(g.nz, cdiv(g.nz, c->row_zc))) continue;
      const double atoms = R * zc * rho;
      int T = 32 * cdiv((long long) (atoms * 1.03 + 2.0), 32);
      if (c->row_threads > 0) T = c->row_threads / 32 * 32;
      if (T < 32) T = 32;
      if (T > PR_TMAX) continue;
      const int nw = zc + 2;
      // staged-atom capacity: the mean plus the larger of 4 % + 16 and 3.5 sigma of a Poisson count (a
      // dense DPD fluid fluctuates far less); a tile that still overflows takes the global-memory path
      const double nmean = NR * nw * rho;
      const double slack = fmax(0.04 * nmean + 16.0, 3.5 * sqrt(nmean));
      const int cap = ((int) ((nmean + slack) * c->row_capscale) + 7) / 8 * 8;
      const size_t smem = pair_row_smem_bytes(R, cap, nw, T, single);
      if (smem > 200 * 1024) continue;
      int ctas = (int) ((227 * 1024) / (smem + 1024));
      if (ctas > 2048 / T) ctas = 2048 / T;
      if (ctas > 32) ctas = 32;
      const double warps = ctas * (T / 32.0);
      const double passes = ceil(atoms / T);
      const double lane_eff = atoms / (passes * T);
      const double redundancy = (double) NR * nw / (R * zc);
      // measured on B200 (tools/pair_bench.py, 1 M atoms): lanes doing work x amortisation of the
      // per-CTA set-up (~40 atoms' worth) x occupancy (saturating at 14 warps/SM), a penalty for the
      // halo redundancy, and one for CTAs of more than 5 warps (their barriers couple too many warps:
      // 288-thread tiles ran 40 % slower than 160-thread tiles at the same occupancy)
      const double score = lane_eff * (atoms / (atoms + 40.0)) * (warps < 14.0 ? warps / 14.0 : 1.0) *
          (T > 256 ? 0.7 : (T > 192 ? 0.8 : (T > 160 ? 0.9 : 1.0))) / (1.0 + 0.03 * redundancy);
      if (score > best_score) {
        best_score = score;
        best = {R, nchunk, zc, T, cap, smem};
      }
    }
  }
  NEED(best_score > 0.0, OBMD_ERR_CAPACITY, "pair_row: no tile shape fits (nz %d, density %.2f)", g.nz, rho);
  return OBMD_OK;
}

// Brick shape of the bulk-staged kernel (pair_brick.cuh) for this box and density: columns C and rows R
// per brick (1, 2, 4), z-cells per brick zc, threads T, landing capacity; ranked by a small cost model
// fitted to tools/pair_bench.py sweeps.  OBMD_B200_COLS / _ROWS / _ZC / _PT override.
struct BrickPlan {
  int C, R, nchunk, zc, T, capL;
  size_t smem;
};

static int pair_brick_plan(obmd_ctx *c, bool single, BrickPlan &best)
{
  const Grid &g = c->grid;
  const double rho = (double) c->nsorted / ((double) (g.own_b - g.own_a + 1) * g.ny * g.nz);
  const int ncols = g.own_b - g.own_a + 1;
  // the density profile along the open axis is not flat (buffers vs region of interest): the landing
  // capacity follows the fullest cell column
  const double rho_col = (double) c->colmax / ((double) g.ny * g.nz);
  const double rho_max = rho_col > rho ? rho_col : rho;
  double best_score = -1.0;
  for (int C = 1; C <= 4; C *= 2) {
    if (c->row_cols > 0 && C != c->row_cols) continue;
    if (C > 1 && C > ncols) continue;
    for (int R = 1; R <= 4; R *= 2) {
      if (c->row_rows > 0 && R != c->row_rows) continue;
      if (R > 1 && R > g.ny) continue;
      const int NROW = (C + 2) * (R + 2);
      for (int nchunk = 1; nchunk <= g.nz; nchunk++) {
        const int zc = cdiv(g.nz, nchunk);
        if (zc > PB_MAXZC) continue;
        if (c->row_zc > 0 && zc != cdiv(g.nz, cdiv(g.nz, c->row_zc))) continue;
        const double atoms = C * R * zc * rho;
        int T = 32 * cdiv((long long) (atoms * 1.03 + 2.0), 32);
        if (c->row_threads > 0) T = c->row_threads / 32 * 32;
        if (T < 64) T = 64;
        if (T > PB_TMAX) T = PB_TMAX;  // several passes over the brick's atoms
        const int nw = zc + 2;
        // landing capacity: the mean plus the larger of 5 % + 16 and 2.5 sigma of a Poisson count (a dense
        // DPD fluid fluctuates far less); a brick that still overflows takes the global-memory path
        const double nmean = NROW * nw * rho_max;
        const double slack = fmax(0.05 * nmean + 16.0, 2.5 * sqrt(nmean));
        const int capL = ((int) ((nmean + slack) * c->row_capscale) + 7) / 8 * 8;
        if (capL > 60000) continue;  // landing indices are 16-bit
        const size_t smem = pair_brick_smem_bytes(C, R, capL, nw, T, single);
        if (smem > 200 * 1024) continue;
        int ctas = (int) ((227 * 1024) / (smem + 1024));
        if (ctas > 2048 / T) ctas = 2048 / T;
        if (ctas > 32) ctas = 32;
        const double warps = ctas * (T / 32.0);
        const double passes = ceil(atoms / T);
        const double lane_eff = atoms / (passes * T);
        const double redundancy = (double) NROW * nw / (C * R * zc);
        const double score = lane_eff * (atoms / (atoms + 40.0)) * (warps < 24.0 ? warps / 24.0 : 1.0) /
            (1.0 + 0.10 * redundancy);
        if (score > best_score) {
          best_score = score;
          best = {C, R, nchunk, zc, T, capL, smem};
        }
      }
    }
  }
  NEED(best_score > 0.0, OBMD_ERR_CAPACITY, "pair_brick: no brick shape fits (nz %d, density %.2f)", g.nz, rho);
  return OBMD_OK;
}

static int colrange_columns(const ColRange &cr)
{
  int n = 0;
  for (int k = 0; k < cr.n; k++) n += cr.b[k] - cr.a[k] + 1;
  return n;
}

template <int C, int R, bool S>
static int pair_launch_brick_t(obmd_ctx *c, long long step, const ColRange &cr, cudaStream_t st, const BrickPlan &pl)
{
  static size_t smem_set = 0;  // per instantiation
  if (pl.smem > smem_set) {
    CK(cudaFuncSetAttribute(k_pair_brick<C, R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pl.smem));
    smem_set = pl.smem;
  }
  int ncoltiles = 0;
  for (int k = 0; k < cr.n; k++) ncoltiles += (cr.b[k] - cr.a[k] + C) / C;
  const int ntiles = ncoltiles * cdiv(c->grid.ny, R) * pl.nchunk;
  const int random_on = c->temperature > 0.0 ? 1 : 0;
  LAUNCH_ON(st, (k_pair_brick<C, R, S>), ntiles, pl.T, pl.smem, c->rec, c->aux, c->cur, cr, c->grid, c->pt,
            c->cell_start, c->fx, c->fy, c->fz, (unsigned) step, random_on, pl.nchunk, pl.zc, pl.capL);
  CK(cudaGetLastError());
  return OBMD_OK;
}

static int pair_launch_brick(obmd_ctx *c, long long step, const ColRange &cr, cudaStream_t st)
{
  const bool single = (c->ntypes == 1);
  BrickPlan pl;
  int rc = pair_brick_plan(c, single, pl);
  if (rc) return rc;
#define BRICK_CASE(NC_, NR_)                                                       \
  if (pl.C == NC_ && pl.R == NR_)                                                  \
    return single ? pair_launch_brick_t<NC_, NR_, true>(c, step, cr, st, pl)       \
                  : pair_launch_brick_t<NC_, NR_, false>(c, step, cr, st, pl)
  BRICK_CASE(1, 1); BRICK_CASE(1, 2); BRICK_CASE(1, 4);
  BRICK_CASE(2, 1); BRICK_CASE(2, 2); BRICK_CASE(2, 4);
  BRICK_CASE(4, 1); BRICK_CASE(4, 2); BRICK_CASE(4, 4);
#undef BRICK_CASE
  return obmd_fail(c, OBMD_ERR_STATE, "pair_brick: no kernel for the planned shape");
}

static int pair_launch_row(obmd_ctx *c, long long step, const ColRange &cr, cudaStream_t st, bool prewrap)
{
  const bool single = (c->ntypes == 1);
  const int random_on = c->temperature > 0.0 ? 1 : 0;
  const Grid &g = c->grid;
  RowPlan pl;
  int rc = pair_row_plan(c, single, pl);
  if (rc) return rc;
  if (pl.smem > c->row_smem_set) {
#define ROW_ATTR(R, S, P) \
  CK(cudaFuncSetAttribute(k_pair_row<R, S, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pl.smem))
    ROW_ATTR(1, true, true); ROW_ATTR(1, true, false); ROW_ATTR(1, false, true); ROW_ATTR(1, false, false);
    ROW_ATTR(2, true, true); ROW_ATTR(2, true, false); ROW_ATTR(2, false, true); ROW_ATTR(2, false, false);
    ROW_ATTR(4, true, true); ROW_ATTR(4, true, false); ROW_ATTR(4, false, true); ROW_ATTR(4, false, false);
    c->row_smem_set = pl.smem;
  }
  const int ntiles = colrange_columns(cr) * cdiv(g.ny, pl.R) * pl.nchunk;
#define ROW_ARGS                                                                                       \
  c->cur, cr, extras_of(c), c->grid, c->pt, c->cell_start, c->extra_head, c->extra_next, c->fx, c->fy, \
      c->fz, (unsigned) step, random_on, pl.nchunk, pl.zc, pl.cap
#define ROW_LAUNCH(R)                                                                                    \
  do {                                                                                                   \
    if (single && prewrap) LAUNCH_ON(st, (k_pair_row<R, true, true>), ntiles, pl.T, pl.smem, ROW_ARGS);  \
    else if (single) LAUNCH_ON(st, (k_pair_row<R, true, false>), ntiles, pl.T, pl.smem, ROW_ARGS);       \
    else if (prewrap) LAUNCH_ON(st, (k_pair_row<R, false, true>), ntiles, pl.T, pl.smem, ROW_ARGS);      \
    else LAUNCH_ON(st, (k_pair_row<R, false, false>), ntiles, pl.T, pl.smem, ROW_ARGS);                  \
  } while (0)
  if (pl.R == 1) ROW_LAUNCH(1);
  else if (pl.R == 2) ROW_LAUNCH(2);
  else ROW_LAUNCH(4);
  CK(cudaGetLastError());
  return OBMD_OK;
}

static int pair_launch(obmd_ctx *c, long long step, bool tally, const ColRange &cr, cudaStream_t st,
                       bool prewrap = false)
{
  const bool single = (c->ntypes == 1);
  const int random_on = c->temperature > 0.0 ? 1 : 0;
  c->pt.dtinvsqrt = 1.0 / sqrt(c->dt);
  const Grid &g = c->grid;
  if (!tally && c->tile_kernel && cr.n > 0 && c->nsorted > 0) {
    // cell-sorted atoms of the column set: the brick kernel (pair_brick.cuh, staged by the copy engine from
    // the records the sort wrote) for the columns whose stencil holds no appended atoms, the thread-staged
    // row kernel (pair_row.cuh) for the others (the two columns next to each slab face; every column when
    // insertions are pending or the records are stale).  The atoms appended after the sort go through
    // the list kernel below.
    int ca = 0, cb = -1;  // columns the brick kernel may take
    if (c->tile_kernel >= 2 && c->nins == 0 && c->rec_valid) {
      if (c->narr + c->nghost == 0) { ca = 0; cb = g.nx - 1; }
      else { ca = g.own_a + 2; cb = g.own_b - 2; }
    }
    ColRange in, out[2];
    memset(&in, 0, sizeof(in));
    memset(out, 0, sizeof(out));
    int nout = 0;
    for (int k = 0; k < cr.n; k++) {
      const int a = cr.a[k], b = cr.b[k];
      const int ia = a > ca ? a : ca, ib = b < cb ? b : cb;
      if (ia <= ib) {
        in.a[in.n] = ia; in.b[in.n] = ib; in.n++;
        if (a < ia) { ColRange &o = out[nout / 2]; o.a[o.n] = a; o.b[o.n] = ia - 1; o.n++; nout++; }
        if (ib < b) { ColRange &o = out[nout / 2]; o.a[o.n] = ib + 1; o.b[o.n] = b; o.n++; nout++; }
      } else {
        ColRange &o = out[nout / 2]; o.a[o.n] = a; o.b[o.n] = b; o.n++; nout++;
      }
    }
    int rc = OBMD_OK;
    if (in.n > 0 && (rc = pair_launch_brick(c, step, in, st))) return rc;
    for (int k = 0; k < 2; k++)
      if (out[k].n > 0 && (rc = pair_launch_row(c, step, out[k], st, prewrap))) return rc;
    if (!cr.extras || c->narr + c->nins == 0) return OBMD_OK;
    ColRange xr;
    memset(&xr, 0, sizeof(xr));
    xr.extras = cr.extras;
    return pair_launch(c, step, false, xr, st, prewrap);
  }
  const int nmax = cr.n == 0 ? c->nextra : colrange_bound(c, cr);
  if (nmax <= 0) return OBMD_OK;
  const int nb = cdiv(nmax, PAIR_T);
  if (tally && c->vatom_cap < c->cap) {  // per-atom virial xx for the stress profile (obmd_profile_x2)
    cudaFree(c->vatom);
    c->vatom = nullptr;
    CK(cudaMalloc(&c->vatom, sizeof(double) * c->cap));
    c->vatom_cap = c->cap;
  }
  if (tally && nb * 7 > c->partial_cap) {
    if (c->partials) cudaFree(c->partials);
    c->partial_cap = nb * 7 + 1024;
    CK(cudaMalloc(&c->partials, sizeof(double) * c->partial_cap));
  }
#define PAIR_ARGS                                                                             \
  c->cur, cr, extras_of(c), c->grid, c->pt, c->cell_start, c->extra_head, c->extra_next,          \
      c->fx, c->fy, c->fz, (unsigned) step, random_on, c->partials, (tally ? c->vatom : nullptr)
  if (prewrap) {  // side-stream launch of obmd_run: coordinates may not be wrapped yet
    if (single) LAUNCH_ON(st, (k_pair_dpd<true, false, true>), nb, PAIR_T, 0, PAIR_ARGS);
    else LAUNCH_ON(st, (k_pair_dpd<false, false, true>), nb, PAIR_T, 0, PAIR_ARGS);
  } else if (single && !tally) LAUNCH_ON(st, (k_pair_dpd<true, false, false>), nb, PAIR_T, 0, PAIR_ARGS);
  else if (single && tally) LAUNCH_ON(st, (k_pair_dpd<true, true, false>), nb, PAIR_T, 0, PAIR_ARGS);
  else if (!single && !tally) LAUNCH_ON(st, (k_pair_dpd<false, false, false>), nb, PAIR_T, 0, PAIR_ARGS);
  else LAUNCH_ON(st, (k_pair_dpd<false, true, false>), nb, PAIR_T, 0, PAIR_ARGS);
  if (tally) LAUNCH_ON(st, k_pair_tally_final, 1, 256, 0, c->partials, nb, c->sc_d);
  CK(cudaGetLastError());
  return OBMD_OK;
}

// forces of the atoms in a column set only (used by obmd_run to overlap the interior with USHER)
int pair_compute_cols(obmd_ctx *c, long long step, const ColRange &cr, cudaStream_t st)
{
  NEED(c->cells_valid, OBMD_ERR_STATE, "pair_compute_cols: cell list is stale");
  ScopedTimer t(c, "pair", st);
  return pair_launch(c, step, false, cr, st, true);
}

int pair_compute(obmd_ctx *c, long long step, int eflag, int vflag)
{
  NEED(c->cells_valid, OBMD_ERR_STATE, "obmd_pair_compute: cell list is stale (call obmd_build_cells)");
  ScopedTimer t(c, "pair");
  int rc = pair_launch(c, step, eflag || vflag, all_columns(c), c->stream);
  if (rc) return rc;
  c->forces_valid = true;
  c->vatom_valid = (eflag || vflag);
  return OBMD_OK;
}
