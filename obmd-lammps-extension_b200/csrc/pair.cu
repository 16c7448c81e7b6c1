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
// Roofline: algorithmic traffic is 80 B/atom (x,v,tag,type in, f out); neighbour reads are
// served by L1/L2.  The kernel is fp64-pipe bound, not HBM bound.
#include "ctx.h"

#define EPSILON_DPD 1.0e-10  // pair_dpd.cpp:36

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
  float r = sqrtf(-2.0f * __logf(u1));
  return (double) (r * __cosf(6.283185307179586f * u2));
}

struct PairAcc {
  double fx, fy, fz, e, v0, v1, v2, v3, v4, v5;
};

template <bool SINGLE_TYPE, bool TALLY>
__device__ __forceinline__ void pair_ij(const Planes &p, const PairTab &pt, int j, double xi, double yi,
                                        double zi, double vxi, double vyi, double vzi, int tagi, int ti,
                                        const Grid &g, unsigned step, bool random_on, PairAcc &a)
{
  double delx = xi - p.x[j];
  double dely = yi - p.y[j];
  double delz = zi - p.z[j];
  // minimum image in y,z (boundary f p p); positions are wrapped, one shift suffices
  if (dely > g.half[1]) dely -= g.prd[1];
  else if (dely < -g.half[1]) dely += g.prd[1];
  if (delz > g.half[2]) delz -= g.prd[2];
  else if (delz < -g.half[2]) delz += g.prd[2];
  double rsq = delx * delx + dely * dely + delz * delz;
  int tj = SINGLE_TYPE ? 1 : p.type[j];
  if (rsq < pt.cutsq[ti][tj]) {
    double r = sqrt(rsq);
    if (r < EPSILON_DPD) return;
    double rinv = 1.0 / r;
    double delvx = vxi - p.vx[j], delvy = vyi - p.vy[j], delvz = vzi - p.vz[j];
    double dot = delx * delvx + dely * delvy + delz * delvz;
    double wd = 1.0 - r / pt.cut[ti][tj];
    double fpair = pt.a0[ti][tj] * wd;
    fpair -= pt.gamma[ti][tj] * wd * wd * dot * rinv;
    if (random_on) {
      double randnum = dpd_gaussian(tagi, p.tag[j], step, pt.seed);
      fpair += pt.sigma[ti][tj] * wd * randnum * pt.dtinvsqrt;
    }
    fpair *= rinv;
    a.fx += delx * fpair;
    a.fy += dely * fpair;
    a.fz += delz * fpair;
    if (TALLY) {
      a.e += 0.5 * (0.5 * pt.a0[ti][tj] * pt.cut[ti][tj] * wd * wd);
      double h = 0.5 * fpair;
      a.v0 += h * delx * delx;
      a.v1 += h * dely * dely;
      a.v2 += h * delz * delz;
      a.v3 += h * delx * dely;
      a.v4 += h * delx * delz;
      a.v5 += h * dely * delz;
    }
  }
}

template <bool SINGLE_TYPE, bool TALLY>
__global__ void __launch_bounds__(128)
    k_pair_dpd(Planes p, int n, int nsorted, int nextra, Grid g, PairTab pt,
               const int *__restrict__ cell_start, const int *__restrict__ extra_head,
               const int *__restrict__ extra_next, double *__restrict__ fx, double *__restrict__ fy,
               double *__restrict__ fz, unsigned step, int random_on, double *__restrict__ partials)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  PairAcc a = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (i < n) {
    double xi = p.x[i], yi = p.y[i], zi = p.z[i];
    double vxi = p.vx[i], vyi = p.vy[i], vzi = p.vz[i];
    int tagi = p.tag[i];
    int ti = SINGLE_TYPE ? 1 : p.type[i];
    int cx = cell_cx(g, xi), cy = cell_cy(g, yi), cz = cell_cz(g, zi);
    for (int ox = -1; ox <= 1; ox++) {
      int bx = cx + ox;
      if (bx < 0 || bx >= g.nx) continue;
      for (int oy = -1; oy <= 1; oy++) {
        int by = cy + oy;
        by = by < 0 ? by + g.ny : (by >= g.ny ? by - g.ny : by);
        int row = (bx * g.ny + by) * g.nz;
        for (int oz = -1; oz <= 1; oz++) {
          int bz = cz + oz;
          bz = bz < 0 ? bz + g.nz : (bz >= g.nz ? bz - g.nz : bz);
          int cc = row + bz;
          int jb = cell_start[cc], je = cell_start[cc + 1];
          for (int j = jb; j < je; j++) {
            if (j == i) continue;
            pair_ij<SINGLE_TYPE, TALLY>(p, pt, j, xi, yi, zi, vxi, vyi, vzi, tagi, ti, g, step,
                                        random_on != 0, a);
          }
          if (nextra > 0) {
            for (int e = extra_head[cc]; e >= 0; e = extra_next[e]) {
              int j = nsorted + e;
              if (j == i) continue;
              pair_ij<SINGLE_TYPE, TALLY>(p, pt, j, xi, yi, zi, vxi, vyi, vzi, tagi, ti, g, step,
                                          random_on != 0, a);
            }
          }
        }
      }
    }
    fx[i] = a.fx;
    fy[i] = a.fy;
    fz[i] = a.fz;
  }
  if (TALLY) {
    // deterministic block reduction of (e, v[6]) -> partials[block][7]
    __shared__ double sh[7][4];
    double v[7] = {a.e, a.v0, a.v1, a.v2, a.v3, a.v4, a.v5};
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
      for (int ww = 0; ww < (int) (blockDim.x >> 5); ww++) s += sh[threadIdx.x][ww];
      partials[blockIdx.x * 7 + threadIdx.x] = s;
    }
  }
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

int pair_compute(obmd_ctx *c, long long step, int eflag, int vflag)
{
  NEED(c->cells_valid, OBMD_ERR_STATE, "obmd_pair_compute: cell list is stale (call obmd_build_cells)");
  ScopedTimer t(c, "pair");
  const int n = c->nsorted + c->nextra;
  const int nb = cdiv(n, 128);
  const bool tally = eflag || vflag;
  const bool single = (c->ntypes == 1);
  const int random_on = c->temperature > 0.0 ? 1 : 0;
  c->pt.dtinvsqrt = 1.0 / sqrt(c->dt);
  if (tally && nb * 7 > c->partial_cap) {
    if (c->partials) cudaFree(c->partials);
    c->partial_cap = nb * 7 + 1024;
    CK(cudaMalloc(&c->partials, sizeof(double) * c->partial_cap));
  }
#define PAIR_ARGS                                                                              \
  c->cur, n, c->nsorted, c->nextra, c->grid, c->pt, c->cell_start, c->extra_head, c->extra_next, \
      c->fx, c->fy, c->fz, (unsigned) step, random_on, c->partials
  if (single && !tally) LAUNCH((k_pair_dpd<true, false>), nb, 128, 0, PAIR_ARGS);
  else if (single && tally) LAUNCH((k_pair_dpd<true, true>), nb, 128, 0, PAIR_ARGS);
  else if (!single && !tally) LAUNCH((k_pair_dpd<false, false>), nb, 128, 0, PAIR_ARGS);
  else LAUNCH((k_pair_dpd<false, true>), nb, 128, 0, PAIR_ARGS);
  if (tally) LAUNCH(k_pair_tally_final, 1, 256, 0, c->partials, nb, c->sc_d);
  CK(cudaGetLastError());
  c->forces_valid = true;
  return OBMD_OK;
}
