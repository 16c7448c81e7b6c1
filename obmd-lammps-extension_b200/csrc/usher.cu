// Batched USHER on the device: try_inserting (fix_obmd_merged.cpp:869-1225), usher (:1518-1616),
// energy (:1774-1810), PairDPD::single (pair_dpd.cpp:401-418).  Compiled with -fmad=false: every
// sum, product and quotient that reaches a trial coordinate is evaluated exactly as the x86-64
// reference build evaluates it (no FMA contraction, IEEE sqrt and division, contributions summed
// in ascending LAMMPS local index).
//
// The reference inserts serially: trial t sees the atoms accepted by trials < t.  Here
//   round 0   k_usher_eval evaluates every trial of the call speculatively against the
//             pre-insertion configuration, one lane group (8, 16 or 32 lanes) per trial;
//   k_usher_diff / k_usher_dirty find the trials whose evaluation points lie within the cutoff of an
//             atom accepted by an earlier trial;
//   k_usher_fix (one thread-block cluster, no host round trips) re-runs those trials from the first
//             affected point, with the accepted atoms visible through per-cell lists, and repeats
//             diff -> dirty -> re-run until nothing changes.  Trial t depends only on trials < t, so
//             the fixed point equals the serial result bit for bit.
//
// One energy evaluation: the group keeps an fp32 copy (relative to the corner of the point's cell,
// periodic image already chosen) of the atoms of the 27 surrounding cells in shared memory and
// refreshes it only when the point changes cell.  Candidates are screened in fp32 against a widened
// cutoff (a superset of the exact test), the few survivors are evaluated in fp64 with the
// reference's expressions (minimum-image loop, exact cutoff test, sqrt, division), ranked by local
// index and summed left to right by one lane per quantity.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "ctx.h"
#include "glibc_pow_tables.h"

namespace cg = cooperative_groups;

#define EPSILON_FIX 1.0e-6   // fix_obmd_merged.cpp:62
#define EPSILON_DPD 1.0e-10  // pair_dpd.cpp:36

#define USH_CAND 144  // atoms in the 27 stencil cells of a point (mean 81 at rho = 3)
#define USH_CAP 48    // fp32-screened neighbours of a point (mean 13 at rho = 3)
#define USH_T 128     // threads per CTA of the round-0 kernel
#define FIX_T 512     // threads per CTA there (one warp per re-run trial: 128 trials at a time)
#ifndef USHER_DEBUG
#define USHER_DEBUG 0  // 1: OBMD_B200_DEBUG_FIX=1 prints phase times of the fix-up kernel (costs registers)
#endif
#define KEY_DROP 0x7fffff00  // keys of screened candidates that fail the exact cutoff test (+ slot)

// per lane group: the stencil of the cell the point is in, and the scratch of one energy evaluation.
// FAT (re-run rounds, where latency is everything and shared memory is plentiful): the exact
// coordinates and summation keys of the candidates are kept too, so an evaluation never leaves the SM.
template <bool FAT>
struct UGroupFat {};
template <>
struct UGroupFat<true> {
  double cx[USH_CAND], cy[USH_CAND], cz[USH_CAND];
  int ckey[USH_CAND];
};
template <bool FAT>
struct __align__(16) UGroupT : UGroupFat<FAT> {
  uint4 scr[USH_CAND];  // x, y, z (fp32 bits, relative to the cell corner), w = slot (>= 0) or -(entry + 1)
  double e[USH_CAP], fx[USH_CAP], fy[USH_CAP], fz[USH_CAP];
  int key[USH_CAP];              // LAMMPS local index of the contributing atom
  unsigned short hq[USH_CAP];    // candidate index of each screened neighbour
};
typedef UGroupT<false> UGroup;

struct UsherArgs {
  Planes p;
  Grid g;
  FixDev fd;
  UsherCfg u;
  Extras ex;
  const int *cell_start, *extra_head, *extra_next;
  const int *keys;    // summation-order key of every stored atom: LAMMPS local index, or the tag (see usher_pass)
  int spec_key0;      // key of the atom trial 0 of the call would create (later trials: + global trial index)
  // atoms accepted so far in this call ("spec" entries, append-only; a dead entry has trial INT_MAX).
  // There are few of them (one per accepted trial), so a gather filters the whole array against
  // its 27-cell window instead of walking per-cell lists.
  int *spec_trial, *spec_of_trial;
  double *spec_pos;
  int spec_cap;
  obmd_trial *cur;
  double *path;
  int *npts;
  float *bbox;
  int *kfirst, *list;
  int2 *pairs;        // (trial, change record) pairs that passed the bounding-box test
  int pair_cap;
  ChgRec *chg;
  int *acc_list;
  StepScalars *sc;
  int ntr, limit;
  double rc, rsqmax;  // dirty test: cutoff with a small margin
  float wf[3], pf[3], ipf[3];  // fp32 screen: cell widths, periods, inverse periods
  float cut32;        // widened squared cutoff of the fp32 screen
  int debug;          // OBMD_B200_DEBUG_FIX: print the phase times of the fix-up kernel
};

// the accepted-atom entries of this call, cached in shared memory by the fix-up kernel at the start of
// every re-run round (entries beyond the cache are read from global memory)
// The cached entries hang off per-cell chains (head[cell & hmask] -> next[e]; an entry remembers its cell),
// so a gather finds the accepted atoms of its 27 cells with one shared-memory load per cell.
struct SpecSm {
  const double *x, *y, *z;
  const int *tr, *cell, *next, *head;
  int hmask;
  int n_sm, n_all;
};

// Pull a line into L1 with a real load whose result is never used (no scoreboard wait): unlike
// prefetch.global.L1 it cannot be dropped.
__device__ __forceinline__ void touch_l1(const void *p)
{
  unsigned d;
  asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(d) : "l"(p) : "memory");
}

__device__ __forceinline__ int ranpark_pow(long long n)
{
  unsigned long long m = 2147483647ULL, a = 16807ULL, r = 1ULL;
  while (n > 0) {
    if (n & 1) r = (r * a) % m;
    a = (a * a) % m;
    n >>= 1;
  }
  return (int) r;
}

__device__ __forceinline__ double ranpark_uniform(int &seed)  // random_park.cpp:43-50
{
  int k = seed / 127773;
  seed = 16807 * (seed - k * 127773) - 2836 * k;
  if (seed < 0) seed += 2147483647;
  return (1.0 / 2147483647) * seed;
}

// pow() as the reference evaluates it.  usher() takes ds = dsovlp - pow(4 eps / E, 1/12) when E > uovlp
// (fix_obmd_merged.cpp:1571); the reference calls GNU libm, whose pow (glibc 2.39,
// sysdeps/ieee754/dbl-64/e_pow.c, the FMA variant x86-64 hosts dispatch to) is accurate to ~0.52 ULP but
// not correctly rounded, and CUDA's pow differs from it in the last bit for ~0.1 % of the arguments --
// enough to send a trial down another path.  This is that algorithm restated operation by operation
// (log in double-double with the 128-entry table, y * log, exp with the 128-entry table; explicit
// fma / mul / add exactly where the reference build has them), for positive normal x and the
// exponent range the main path of e_pow.c covers; anything else falls back to CUDA's pow.  The
// tables come from tools/gen_glibc_pow_tables.py.  Checked against libm on 4e7 arguments (CPU, same
// sequence) and through tests/golden/b9_usher_overlap.npz.
__device__ __noinline__ double glibc_pow(double x, double y)
{
  const unsigned long long ix = (unsigned long long) __double_as_longlong(x);
  const unsigned long long iy = (unsigned long long) __double_as_longlong(y);
  const unsigned topx = (unsigned) (ix >> 52), topy = (unsigned) (iy >> 52) & 0x7ff;
  if (topx - 1u >= 0x7feu || topy - 0x3beu > 0x7fu) return pow(x, y);
  const unsigned long long tmp = ix - 0x3fe6955500000000ULL;
  const int i = (int) ((tmp >> 45) & 0x7f);
  const int k = (int) ((long long) tmp >> 52);
  const double z = __longlong_as_double((long long) (ix - (tmp & 0xfff0000000000000ULL)));
  const double kd = (double) k;
#define GPD(q) __longlong_as_double((long long) (q))
  const double invc = GPD(GP_LOG_TAB[i][0]), logc = GPD(GP_LOG_TAB[i][1]), logctail = GPD(GP_LOG_TAB[i][2]);
  const double A0 = GPD(GP_LOG_HDR[2]), A1 = GPD(GP_LOG_HDR[3]), A2 = GPD(GP_LOG_HDR[4]), A3 = GPD(GP_LOG_HDR[5]);
  const double A4 = GPD(GP_LOG_HDR[6]), A5 = GPD(GP_LOG_HDR[7]), A6 = GPD(GP_LOG_HDR[8]);
  const double t1 = __fma_rn(kd, GPD(GP_LOG_HDR[0]), logc);
  const double lo1 = __fma_rn(kd, GPD(GP_LOG_HDR[1]), logctail);
  const double r = __fma_rn(z, invc, -1.0);
  const double ar = __dmul_rn(A0, r);
  const double p12 = __fma_rn(r, A2, A1), p34 = __fma_rn(r, A4, A3);
  const double t2 = __dadd_rn(r, t1);
  const double lo2 = __dadd_rn(__dsub_rn(t1, t2), r);
  const double ar2 = __dmul_rn(r, ar), ar3 = __dmul_rn(r, ar2);
  const double lo3 = __fma_rn(ar, r, -ar2);
  const double hi = __dadd_rn(t2, ar2);
  const double p56 = __fma_rn(r, A6, A5);
  const double lo4 = __dadd_rn(__dsub_rn(t2, hi), ar2);
  double q = __fma_rn(p56, ar2, p34);
  q = __fma_rn(ar2, q, p12);
  double lo = __dadd_rn(__dadd_rn(__dadd_rn(lo1, lo2), lo3), lo4);
  lo = __fma_rn(ar3, q, lo);
  const double lh = __dadd_rn(hi, lo);
  const double ll = __dadd_rn(__dsub_rn(hi, lh), lo);
  const double ehi = __dmul_rn(y, lh);
  const double elo = __fma_rn(y, ll, __fma_rn(lh, y, -ehi));
  const unsigned tope = (unsigned) ((unsigned long long) __double_as_longlong(ehi) >> 52) & 0x7ff;
  if (tope - 0x3c9u > 0x3eu) return pow(x, y);
  const double Shift = GPD(GP_EXP_HDR[1]);
  const double kds = __fma_rn(ehi, GPD(GP_EXP_HDR[0]), Shift);
  const unsigned long long ki = (unsigned long long) __double_as_longlong(kds);
  const double kd2 = __dsub_rn(kds, Shift);
  double rr = __fma_rn(kd2, GPD(GP_EXP_HDR[2]), ehi);
  rr = __fma_rn(kd2, GPD(GP_EXP_HDR[3]), rr);
  rr = __dadd_rn(elo, rr);
  const unsigned idx = 2u * (unsigned) (ki & 0x7f);
  const unsigned long long sbits = GP_EXP_TAB[idx + 1] + (ki << 45);
  const double c23 = __fma_rn(rr, GPD(GP_EXP_HDR[5]), GPD(GP_EXP_HDR[4]));
  const double tr = __dadd_rn(rr, GPD(GP_EXP_TAB[idx]));
  const double r2 = __dmul_rn(rr, rr);
  const double c45 = __fma_rn(rr, GPD(GP_EXP_HDR[7]), GPD(GP_EXP_HDR[6]));
  double t = __fma_rn(c23, r2, tr);
  t = __fma_rn(c45, __dmul_rn(r2, r2), t);
  const double scale = GPD(sbits);
#undef GPD
  return __fma_rn(t, scale, scale);
}

// PairDPD::single (pair_dpd.cpp:401-418, factor_dpd = 1) and fpair*del (:1801-1805)
__device__ __forceinline__ void usher_single(double a0, double cut, double rsq, double &e, double &fx,
                                             double &fy, double &fz)
{
  double r = sqrt(rsq);
  double fforce, phi;
  if (r < EPSILON_DPD) {
    fforce = 0.0;
    phi = 0.0;
  } else {
    double rinv = 1.0 / r;
    double wd = 1.0 - (cut == 1.0 ? r : r / cut);  // r / 1.0 == r exactly: skip the division
    fforce = a0 * wd * 1.0 * rinv;
    phi = 0.5 * a0 * cut * wd * wd;
    phi = 1.0 * phi;
  }
  e = phi;
  fx = fforce * fx;
  fy = fforce * fy;
  fz = fforce * fz;
}

// ---- stencil gather -------------------------------------------------------------------------------
// The lanes of a group own the 27 cells around (cx, cy, cz).  Each lane reads its cells' ranges (sorted
// atoms plus the atoms appended after the sort: slab runs, earlier passes), a group scan gives it
// its place in the candidate list, and it stores the fp32 screening copy of its own atoms directly:
// coordinates relative to the corner of the centre cell, shifted to the periodic image next to it
// (the lane knows whether its cell was reached across a periodic face).  Atoms in a face cell may
// still carry coordinates Domain::pbc has not wrapped yet; only those cells pay for a generic fix.
// Atoms accepted earlier in this call by trials < t are filtered against the 27-cell window.
struct LaneCell {  // per lane, fixed for the kernel: offsets of the cells it owns
  int ox[2], oy[2], oz[2];
};

template <int G, bool FAT>
__device__ __forceinline__ int usher_gather(const UsherArgs &A, UGroupT<FAT> &S, const LaneCell &lc, int cx, int cy,
                                            int cz, int t, double ox, double oy, double oz, int gl, int gbase,
                                            unsigned gmask, bool have_extras, const SpecSm *sp,
                                            long long *pc = nullptr)
{
  constexpr int NCL = (27 + G - 1) / G;
  long long gc0 = pc ? clock64() : 0;
  static_assert(NCL <= 2, "lane groups of at least 16 lanes");
  constexpr unsigned GM = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  const Grid &g = A.g;
  int pf_a = 0, pf_b = 0;  // re-run rounds: slots of the five z-cells around cz in this lane's column
  if (G == 32 && gl < 25) {
    // the next gather of this trial is at most one cell away: pull the cell_start rows of the 5 x 5
    // columns around this cell into L1 now, so that its first dependent round trip is short
    const int a = (gl * 13) >> 6;  // gl / 5 for gl < 25
    const int bx = cx + a - 2;
    int by = cy + (gl - 5 * a) - 2;
    by = by < 0 ? by + g.ny : (by >= g.ny ? by - g.ny : by);
    if (bx >= 0 && bx < g.nx && by >= 0 && by < g.ny) {
      const int *row = A.cell_start + (bx * g.ny + by) * g.nz;
      const int z0 = max(cz - 2, 0), z1 = min(cz + 3, g.nz);
      if (sp) {
        // A re-run round is a handful of dependent chains and every gather is two or three dependent
        // trips to L2 / HBM.  Read the slot range of the column's five cells now (used at the end of
        // this gather) and prefetch those atoms: the next gather then finds everything in L1.
        pf_a = __ldg(row + z0);
        pf_b = __ldg(row + z1);
      } else {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(row + z0));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(row + z1));
      }
    }
  }
  int jb[NCL], cnt[NCL], hd[NCL], c2s[NCL];
  double oyl[NCL], ozl[NCL];
  bool face[NCL];
  int tot = 0;
#pragma unroll
  for (int s = 0; s < NCL; s++) {
    jb[s] = 0;
    cnt[s] = 0;
    hd[s] = -1;
    c2s[s] = -1;
    oyl[s] = oy;
    ozl[s] = oz;
    face[s] = false;
    const int bx = cx + lc.ox[s];
    if (gl + s * G < 27 && bx >= 0 && bx < g.nx) {
      int by = cy + lc.oy[s], bz = cz + lc.oz[s];  // offsets are +-1: wrap without a division
      // a cell reached across the low face holds atoms one period above their image next to the centre
      if (by < 0) { by += g.ny; oyl[s] = oy + g.prd[1]; }
      else if (by >= g.ny) { by -= g.ny; oyl[s] = oy - g.prd[1]; }
      if (bz < 0) { bz += g.nz; ozl[s] = oz + g.prd[2]; }
      else if (bz >= g.nz) { bz -= g.nz; ozl[s] = oz - g.prd[2]; }
      face[s] = by == 0 || by == g.ny - 1 || bz == 0 || bz == g.nz - 1;
      const int c2 = (bx * g.ny + by) * g.nz + bz;
      jb[s] = __ldg(A.cell_start + c2);
      cnt[s] = __ldg(A.cell_start + c2 + 1) - jb[s];
      tot += cnt[s];
      if (have_extras) {
        hd[s] = __ldg(A.extra_head + c2);
        for (int e = hd[s]; e >= 0; e = __ldg(A.extra_next + e)) tot++;
      }
      if (sp) {  // atoms accepted earlier in this call by trials < t (a retired entry has trial INT_MAX)
        c2s[s] = c2;
        for (int e = sp->head[c2 & sp->hmask]; e >= 0; e = sp->next[e])
          if (sp->cell[e] == c2 && sp->tr[e] < t) tot++;
      }
    }
  }
  int inc = tot;
#pragma unroll
  for (int d = 1; d < G; d <<= 1) {
    const int v = __shfl_up_sync(gmask, inc, d, G);
    if (gl >= d) inc += v;
  }
  int ncand = __shfl_sync(gmask, inc, G - 1, G);
  if (ncand > USH_CAND) {
    if (gl == 0) A.sc->usher_error = OBMD_ERR_CAPACITY;
    return 0;
  }
  int off = inc - tot;  // exclusive
  if (pc) { long long c1 = clock64(); pc[8] += c1 - gc0; gc0 = c1; }
#pragma unroll
  for (int s = 0; s < NCL; s++) {
    const float ty = ((float) lc.oy[s] + 0.5f) * A.wf[1], tz = ((float) lc.oz[s] + 0.5f) * A.wf[2];
    // the cell's sorted atoms, four at a time with all twelve coordinate loads in flight at once (a
    // cell holds ~3 atoms: one round trip in almost every case), then its list of appended atoms
    for (int k0 = 0; k0 < cnt[s]; k0 += 4) {
      double xj[4], yj[4], zj[4];
      int kj[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int j = jb[s] + min(k0 + q, cnt[s] - 1);
        xj[q] = __ldg(A.p.x + j);
        yj[q] = __ldg(A.p.y + j);
        zj[q] = __ldg(A.p.z + j);
        if (FAT) kj[q] = __ldg(A.keys + j);
      }
      // the exact evaluation will want the local index of the few candidates inside the cutoff
      if (!FAT) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.keys + jb[s] + k0));
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (k0 + q >= cnt[s]) break;
        const float rx = (float) (xj[q] - ox);
        float ry = (float) (yj[q] - oyl[s]), rz = (float) (zj[q] - ozl[s]);
        if (face[s]) {  // coordinates Domain::pbc has not wrapped yet
          ry -= A.pf[1] * rintf((ry - ty) * A.ipf[1]);
          rz -= A.pf[2] * rintf((rz - tz) * A.ipf[2]);
        }
        S.scr[off] = make_uint4(__float_as_uint(rx), __float_as_uint(ry), __float_as_uint(rz),
                                (unsigned) (jb[s] + k0 + q));
        if constexpr (FAT) {
          S.cx[off] = xj[q];
          S.cy[off] = yj[q];
          S.cz[off] = zj[q];
          S.ckey[off] = kj[q];
        }
        off++;
      }
    }
    for (int e = hd[s]; e >= 0; e = __ldg(A.extra_next + e)) {
      const int j = A.ex.base + e;
      const double xj = __ldg(A.p.x + j), yj = __ldg(A.p.y + j), zj = __ldg(A.p.z + j);
      const float rx = (float) (xj - ox);
      float ry = (float) (yj - oyl[s]), rz = (float) (zj - ozl[s]);
      if (face[s]) {
        ry -= A.pf[1] * rintf((ry - ty) * A.ipf[1]);
        rz -= A.pf[2] * rintf((rz - tz) * A.ipf[2]);
      }
      S.scr[off] = make_uint4(__float_as_uint(rx), __float_as_uint(ry), __float_as_uint(rz), (unsigned) j);
      if constexpr (FAT) {
        S.cx[off] = xj;
        S.cy[off] = yj;
        S.cz[off] = zj;
        S.ckey[off] = __ldg(A.keys + j);
      }
      off++;
    }
    if (sp && c2s[s] >= 0) {
      for (int e = sp->head[c2s[s] & sp->hmask]; e >= 0; e = sp->next[e]) {
        const int tr = sp->tr[e];
        if (sp->cell[e] != c2s[s] || tr >= t) continue;
        const double xj = sp->x[e], yj = sp->y[e], zj = sp->z[e];
        const float rx = (float) (xj - ox);
        float ry = (float) (yj - oyl[s]), rz = (float) (zj - ozl[s]);
        if (face[s]) {  // an end point exactly on the upper periodic face is binned into layer 0
          ry -= A.pf[1] * rintf((ry - ty) * A.ipf[1]);
          rz -= A.pf[2] * rintf((rz - tz) * A.ipf[2]);
        }
        S.scr[off] = make_uint4(__float_as_uint(rx), __float_as_uint(ry), __float_as_uint(rz), (unsigned) (-(e + 1)));
        if constexpr (FAT) {
          S.cx[off] = xj;
          S.cy[off] = yj;
          S.cz[off] = zj;
          S.ckey[off] = A.spec_key0 + tr;
        }
        off++;
      }
    }
  }
  if (pc) { long long c1 = clock64(); pc[9] += c1 - gc0; gc0 = c1; }
  if (sp && pf_b > pf_a) {
    // one 128-byte line holds 16 coordinates / 32 keys: a column's five cells (~15 atoms) are one or
    // two lines per plane
    const char *px = reinterpret_cast<const char *>(A.p.x), *py = reinterpret_cast<const char *>(A.p.y);
    const char *pz = reinterpret_cast<const char *>(A.p.z), *pk = reinterpret_cast<const char *>(A.keys);
    const size_t d0 = ((size_t) pf_a * 8) & ~(size_t) 127, d1 = (size_t) pf_b * 8;
    for (size_t o = d0; o < d1; o += 128) {
      touch_l1(px + o);
      touch_l1(py + o);
      touch_l1(pz + o);
    }
    for (size_t o = ((size_t) pf_a * 4) & ~(size_t) 127; o < (size_t) pf_b * 4; o += 128)
      touch_l1(pk + o);
  }
  if (sp && sp->n_all > sp->n_sm) {
    // accepted atoms beyond the shared-memory cache (thousands of insertions in one call): the entries
    // whose position falls into the 27-cell window (a superset of those within the cutoff of any point
    // of the centre cell) are appended, already in screening form
    const float lox = -1.001f * A.wf[0], hix = 2.001f * A.wf[0], loy = -1.001f * A.wf[1], hiy = 2.001f * A.wf[1];
    const float loz = -1.001f * A.wf[2], hiz = 2.001f * A.wf[2];
    for (int e0 = sp->n_sm; e0 < sp->n_all; e0 += G) {
      const int e = e0 + gl;
      bool in = false;
      float rx = 0.f, ry = 0.f, rz = 0.f;
      double xs = 0.0, ys = 0.0, zs = 0.0;
      int tr = 0x7fffffff;
      if (e < sp->n_all) {
        tr = __ldcg(A.spec_trial + e);
        if (tr < t) {  // a retired entry has trial INT_MAX
          xs = __ldcg(A.spec_pos + 3 * (size_t) e);
          ys = __ldcg(A.spec_pos + 3 * (size_t) e + 1);
          zs = __ldcg(A.spec_pos + 3 * (size_t) e + 2);
          rx = (float) (xs - ox);
          ry = (float) (ys - oy);
          rz = (float) (zs - oz);
          ry -= A.pf[1] * rintf((ry - 0.5f * A.wf[1]) * A.ipf[1]);
          rz -= A.pf[2] * rintf((rz - 0.5f * A.wf[2]) * A.ipf[2]);
          in = rx > lox && rx < hix && ry > loy && ry < hiy && rz > loz && rz < hiz;
        }
      }
      const unsigned m = (G == 32) ? __ballot_sync(gmask, in) : ((__ballot_sync(gmask, in) >> gbase) & GM);
      if (in) {
        const int pos = ncand + __popc(m & ((1u << gl) - 1u));
        if (pos < USH_CAND) {
          S.scr[pos] = make_uint4(__float_as_uint(rx), __float_as_uint(ry), __float_as_uint(rz), (unsigned) (-(e + 1)));
          if constexpr (FAT) {
            S.cx[pos] = xs;
            S.cy[pos] = ys;
            S.cz[pos] = zs;
            S.ckey[pos] = A.spec_key0 + tr;
          }
        }
      }
      ncand += __popc(m);
    }
    if (ncand > USH_CAND) {
      if (gl == 0) A.sc->usher_error = OBMD_ERR_CAPACITY;
      ncand = 0;
    }
  }
  __syncwarp(gmask);
  if (pc) { long long c1 = clock64(); pc[10] += c1 - gc0; }
  return ncand;
}

// ---- one energy evaluation (:1774-1810) -------------------------------------------------------------
// Returns E in every lane of the group and f[d] in lane d + 1 (lane 0: E) as `mine`.
template <bool ST, int G, bool FAT>
__device__ __forceinline__ void usher_energy(const UsherArgs &A, const PairTab &pt, UGroupT<FAT> &S, const double c[3],
                                             double ox, double oy, double oz, int nc, bool wrap,
                                             const double k1[3], int gl, int gbase, unsigned gmask,
                                             double &E, double f[3], long long *pc)
{
  constexpr unsigned GM = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  constexpr int NS = (USH_CAP + G - 1) / G;
  long long c0 = pc ? clock64() : 0;
  const Grid &g = A.g;
  const int it = A.fd.ntype;
  // fp32 screen: a superset of the exact test (coordinates relative to the cell corner are below three
  // cell widths, so the absolute error of rsq is ~1e-6; the margin is 1e-3)
  const float px = (float) (c[0] - ox), py = (float) (c[1] - oy), pz = (float) (c[2] - oz);
  const unsigned lt = (1u << gl) - 1u;
  int nh = 0;
  for (int q0 = 0; q0 < nc; q0 += G) {
    const int q = q0 + gl;
    bool hit = false;
    if (q < nc) {
      const uint4 s = S.scr[q];
      const float dx = px - __uint_as_float(s.x), dy = py - __uint_as_float(s.y), dz = pz - __uint_as_float(s.z);
      hit = dx * dx + dy * dy + dz * dz < A.cut32;
    }
    const unsigned m = (G == 32) ? __ballot_sync(gmask, hit) : ((__ballot_sync(gmask, hit) >> gbase) & GM);
    if (hit) {
      const int s2 = nh + __popc(m & lt);
      if (s2 < USH_CAP) S.hq[s2] = (unsigned short) q;
    }
    nh += __popc(m);
  }
  if (nh > USH_CAP) {
    if (gl == 0) A.sc->usher_error = OBMD_ERR_CAPACITY;
    nh = USH_CAP;
  }
  __syncwarp(gmask);
  if (pc) { long long c1 = clock64(); pc[2] += c1 - c0; c0 = c1; }
  // exact evaluation of the screened candidates: separation with the reference's minimum_image loop
  // (domain.cpp:980-1008, y and z periodic; skipped when the stencil touches no periodic face, where
  // it cannot change anything), exact cutoff test, PairDPD::single.  Results stay in registers until
  // every key is known.
  double ev[NS], xv[NS], yv[NS], zv[NS];
  int kv[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    const int k = gl + s * G;
    kv[s] = 0;
    ev[s] = xv[s] = yv[s] = zv[s] = 0.0;
    if (k < nh) {
      const int hq = S.hq[k];
      const int code = (int) S.scr[hq].w;
      double xj, yj, zj;
      int key, tj = it;
      if constexpr (FAT) {
        xj = S.cx[hq];
        yj = S.cy[hq];
        zj = S.cz[hq];
        key = S.ckey[hq];
        if (!ST && code >= 0) tj = __ldg(A.p.type + code);
      } else if (code >= 0) {
        xj = __ldg(A.p.x + code);
        yj = __ldg(A.p.y + code);
        zj = __ldg(A.p.z + code);
        key = __ldg(A.keys + code);
        if (!ST) tj = __ldg(A.p.type + code);
      } else {  // created earlier in this call: local index nbase + (rank among the accepted trials);
                // the trial index orders the created atoms the same way
        const int e = -code - 1;
        const double *sp = A.spec_pos + 3 * (size_t) e;
        xj = __ldcg(sp);
        yj = __ldcg(sp + 1);
        zj = __ldcg(sp + 2);
        key = A.spec_key0 + __ldcg(A.spec_trial + e);
      }
      double delx = c[0] - xj, dely = c[1] - yj, delz = c[2] - zj;
      if (wrap) {
        while (fabs(dely) > g.half[1]) {
          if (dely < 0.0) dely += g.prd[1];
          else dely -= g.prd[1];
        }
        while (fabs(delz) > g.half[2]) {
          if (delz < 0.0) delz += g.prd[2];
          else delz -= g.prd[2];
        }
      }
      const double rsq = delx * delx + dely * dely + delz * delz;
      if (rsq < (ST ? k1[2] : pt.cutsq[it][tj])) {
        xv[s] = delx;
        yv[s] = dely;
        zv[s] = delz;
        if (ST) usher_single(k1[0], k1[1], rsq, ev[s], xv[s], yv[s], zv[s]);
        else usher_single(pt.a0[it][tj], pt.cut[it][tj], rsq, ev[s], xv[s], yv[s], zv[s]);
      } else
        key = KEY_DROP + k;  // outside the exact cutoff: contributes +0.0 behind every real term
      kv[s] = key;
      S.key[k] = key;
    }
  }
  const int nh4 = (nh + 3) & ~3;
  if (gl < nh4 - nh) S.key[nh + gl] = 0x7fffffff;  // pad: the rank loop reads four keys per trip
  __syncwarp(gmask);
  if (pc) { long long c1 = clock64(); pc[3] += c1 - c0; c0 = c1; }
  // every contribution goes to its rank by local index (keys are distinct), then left-to-right sums
  // in ascending local index, one lane per quantity -- the reference's loop order (:1790-1807)
#pragma unroll
  for (int s = 0; s < NS; s++) {
    const int k = gl + s * G;
    if (k < nh) {
      const int key = kv[s];
      int r = 0;
      for (int k2 = 0; k2 < nh4; k2 += 4) {
        const int4 kk = *reinterpret_cast<const int4 *>(&S.key[k2]);
        r += (kk.x < key) + (kk.y < key) + (kk.z < key) + (kk.w < key);
      }
      S.e[r] = ev[s];
      S.fx[r] = xv[s];
      S.fy[r] = yv[s];
      S.fz[r] = zv[s];
    }
  }
  if (gl < nh4 - nh) {  // adding the +0.0 padding is exact
    S.e[nh + gl] = 0.0;
    S.fx[nh + gl] = 0.0;
    S.fy[nh + gl] = 0.0;
    S.fz[nh + gl] = 0.0;
  }
  __syncwarp(gmask);
  if (pc) { long long c1 = clock64(); pc[4] += c1 - c0; c0 = c1; }
  double acc = 0.0;
  if (gl < 4) {
    const double *arr = gl == 0 ? S.e : (gl == 1 ? S.fx : (gl == 2 ? S.fy : S.fz));
    for (int r = 0; r < nh4; r += 4) {
      const double2 a01 = *reinterpret_cast<const double2 *>(arr + r), a23 = *reinterpret_cast<const double2 *>(arr + r + 2);
      acc += a01.x;
      acc += a01.y;
      acc += a23.x;
      acc += a23.y;
    }
  }
  E = __shfl_sync(gmask, acc, 0, G);
  f[0] = __shfl_sync(gmask, acc, 1, G);
  f[1] = __shfl_sync(gmask, acc, 2, G);
  f[2] = __shfl_sync(gmask, acc, 3, G);
  if (pc) { long long c1 = clock64(); pc[5] += c1 - c0; pc[7] += nh; }
}

// ---- trials ------------------------------------------------------------------------------------
// Every group of G lanes runs whole trials.  `list` == nullptr: trial li is li (round 0, handed out
// from a device counter because trial lengths differ widely); otherwise the trials of `list` are
// re-run from their first affected evaluation point (static assignment: group first, first + stride ...).
template <bool ST, int G, bool FAT>
__device__ __forceinline__ void usher_trials(const UsherArgs &A, const PairTab &pt, UGroupT<FAT> *groups,
                                             const int *list, int nl, int *cursor, int first, int stride,
                                             const SpecSm *sp)
{
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1), gbase = lane & ~(G - 1);
  constexpr unsigned GM = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  const unsigned gmask = GM << gbase;
  UGroupT<FAT> &S = groups[(threadIdx.x >> 5) * (32 / G) + (lane / G)];
  const Grid &g = A.g;
  const FixDev &fd = A.fd;
  const UsherCfg &u = A.u;
  const double k1[3] = {pt.a0[fd.ntype][fd.ntype], pt.cut[fd.ntype][fd.ntype], pt.cutsq[fd.ntype][fd.ntype]};
  const bool have_extras = (A.ex.narr + A.ex.nghost + A.ex.nins) > 0;
  LaneCell lc;
  for (int s = 0; s < 2; s++) {
    const int n = min(gl + s * G, 26);
    const int a = (n * 11) >> 5, b = (((n * 11) >> 5) * 11) >> 5;  // n / 3, n / 9 for n < 32
    lc.oz[s] = n - 3 * a - 1;
    lc.oy[s] = a - 3 * b - 1;
    lc.ox[s] = b - 1;
  }

  bool active = false, more = true;
  int t = 0, side = 0, i = 0, np = 0, nc = 0, gx = 0, gy = 0, gz = 0, nextli = first;
  bool wrap = true;
  double c[3] = {0, 0, 0}, start[3] = {0, 0, 0}, E = 0.0, ox = 0.0, oy = 0.0, oz = 0.0;
  float blo[3] = {0, 0, 0}, bhi[3] = {0, 0, 0};
  double *mypath = A.path;
  unsigned long long tdbg0 = 0;
  int idbg0 = 0, ngath = 0;
  const bool dbgp = USHER_DEBUG && A.debug && list;
  long long pcl[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (;;) {
    if (!active && more) {
      int li;
      if (cursor) {
        li = 0;
        if (gl == 0) li = atomicAdd(cursor, 1);
        li = __shfl_sync(gmask, li, 0, G);
      } else {
        li = nextli;
        nextli += stride;
      }
      if (li < nl) {
        t = list ? __ldcg(list + li) : li;
        const int tg = t + u.tg0;  // position of this trial in the (replicated) RanPark stream
        side = (tg < u.ntrL) ? 0 : 1;
        const obmd_block &reg = fd.region[4 + side];
        // trial start (:923-927): three consecutive RanPark draws, x then y then z
        int seed = (int) (((unsigned long long) u.seed0 * (unsigned long long) ranpark_pow(3LL * tg)) %
                          2147483647ULL);
        for (int d = 0; d < 3; d++) {
          c[d] = reg.lo[d] + ranpark_uniform(seed) * (reg.hi[d] - reg.lo[d]);
          start[d] = c[d];
        }
        mypath = A.path + (size_t) t * (u.natt + 1) * 3;
        i = np = 0;
        for (int d = 0; d < 3; d++) {
          blo[d] = 3.0e38f;
          bhi[d] = -3.0e38f;
        }
        if (list) {
          // Re-run of a trial whose neighbourhood changed: everything before the first affected
          // evaluation point is unchanged, so resume there (the state at the top of iteration k is
          // exactly (path[k], i = k)).  The stored bounding box already covers the kept prefix.
          i = np = __ldcg(A.kfirst + t);
          for (int d = 0; d < 3; d++) {
            c[d] = __ldcg(mypath + 3 * i + d);
            blo[d] = __ldcg(A.bbox + (size_t) t * 6 + d);
            bhi[d] = __ldcg(A.bbox + (size_t) t * 6 + 3 + d);
          }
        }
        gx = -(1 << 30);
        E = 0.0;
        active = true;
        if (USHER_DEBUG && A.debug && list) {
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tdbg0));
          idbg0 = i;
          ngath = 0;
          for (int q = 0; q < 12; q++) pcl[q] = 0;
        }
      } else
        more = false;
    }
    if (!__any_sync(0xffffffffu, active)) break;
    if (!active) continue;

    const obmd_block &reg = fd.region[4 + side];
    long long cA = dbgp ? clock64() : 0;
    // the point is inside the insertion region, hence inside the box: the cell index is at most one
    // period off (y == yhi exactly), so the wrap needs no division
    const int cx = cell_cx(g, c[0]);
    int cy = (int) floor((c[1] - g.lo[1]) * g.inv[1]), cz = (int) floor((c[2] - g.lo[2]) * g.inv[2]);
    if (cy < 0 || cy >= 2 * g.ny) cy = cell_wrap(cy, g.ny);
    else if (cy >= g.ny) cy -= g.ny;
    if (cz < 0 || cz >= 2 * g.nz) cz = cell_wrap(cz, g.nz);
    else if (cz >= g.nz) cz -= g.nz;
    if (dbgp) { long long c1 = clock64(); pcl[0] += c1 - cA; cA = c1; }
    if (cx != gx || cy != gy || cz != gz) {
      ox = g.lo[0] + (cx + g.cxoff) * g.w[0];
      oy = g.lo[1] + cy * g.w[1];
      oz = g.lo[2] + cz * g.w[2];
      nc = usher_gather<G, FAT>(A, S, lc, cx, cy, cz, t, ox, oy, oz, gl, gbase, gmask, have_extras, sp,
                           dbgp ? pcl : nullptr);
      gx = cx;
      gy = cy;
      gz = cz;
      ngath++;
      // pre_exchange runs before Domain::pbc: an atom that has just left through a periodic face still
      // carries its unwrapped coordinate and sits in a face cell, so stencils reaching the face layer
      // (points in the two outermost layers) need the reference's minimum-image loop
      wrap = g.ny < 6 || g.nz < 6 || cy <= 1 || cy >= g.ny - 2 || cz <= 1 || cz >= g.nz - 2;
    }
    if (dbgp) { long long c1 = clock64(); pcl[1] += c1 - cA; cA = c1; }
    double f[3];
    usher_energy<ST, G, FAT>(A, pt, S, c, ox, oy, oz, nc, wrap, k1, gl, gbase, gmask, E, f, dbgp ? pcl : nullptr);
    if (dbgp) cA = clock64();
    if (gl == 0) {
      mypath[np * 3 + 0] = c[0];
      mypath[np * 3 + 1] = c[1];
      mypath[np * 3 + 2] = c[2];
    }
    np++;
    for (int d = 0; d < 3; d++) {  // conservative fp32 bounding box of the evaluated points
      blo[d] = fminf(blo[d], __double2float_rd(c[d]));
      bhi[d] = fmaxf(bhi[d], __double2float_ru(c[d]));
    }
    bool done = false;
    if (E < fd.etarget + EPSILON_FIX) done = true;
    else {
      const double fabs_ = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
      if (fabs_ < EPSILON_FIX) {  // the reference `continue`s without progress: report it
        if (gl == 0) A.sc->usher_error = OBMD_ERR_USHER;
        done = true;
      } else {
        // the four quotients by |f| (:1586-1600), one lane each: the division sequence is issued once
        const double num = gl == 0 ? E - fd.etarget : (gl == 1 ? f[0] : (gl == 2 ? f[1] : f[2]));
        const double quo = num / fabs_;
        double ds;
        if (E > fd.uovlp) ds = fd.dsovlp - glibc_pow(4 * fd.eps / E, 1.0 / 12.0);
        else ds = fmin(__shfl_sync(gmask, quo, 0, G), fd.ds0);
        c[0] += __shfl_sync(gmask, quo, 1, G) * ds;
        c[1] += __shfl_sync(gmask, quo, 2, G) * ds;
        c[2] += __shfl_sync(gmask, quo, 3, G) * ds;
        if (!block_inside(reg, c[0], c[1], c[2])) done = true;  // check_mol_region (:1700-1712)
        else {
          i += 1;
          if (i >= u.natt) done = true;
        }
      }
    }
    if (dbgp) { long long c1 = clock64(); pcl[6] += c1 - cA; }
    if (!done) continue;

    bool ok = (E < fd.etarget + EPSILON_FIX);
    // check_mol_proc / check_proc (:1645-1696) on the owning rank: half-open box in x,y; the top z
    // face belongs to the last processor
    if (ok) {
      const bool in = c[0] >= fd.boxlo[0] && c[0] < fd.boxhi[0] && c[1] >= fd.boxlo[1] && c[1] < fd.boxhi[1] &&
          c[2] >= fd.boxlo[2];
      if (!in) ok = false;
    }
    if (ok && !block_inside(reg, c[0], c[1], c[2])) ok = false;  // :1107-1108
    if (gl == 0) {
      obmd_trial r;
      for (int d = 0; d < 3; d++) {
        r.start[d] = start[d];
        r.end[d] = c[d];
      }
      r.energy = E;
      r.iter = i;
      r.accepted = ok ? 1 : 0;
      r.side = side;
      r.tag = 0;
      A.cur[t] = r;
      A.npts[t] = np;
      for (int d = 0; d < 3; d++) {
        A.bbox[(size_t) t * 6 + d] = blo[d];
        A.bbox[(size_t) t * 6 + 3 + d] = bhi[d];
      }
    }
    active = false;
    if (USHER_DEBUG && A.debug && list && gl == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      printf("  trial %d: evals %d gathers %d nc %d ns %llu | cyc cell %lld gather %lld (cells %lld atoms %lld spec %lld) screen %lld exact %lld rank %lld sum %lld step %lld nh %lld\n",
             t, np - idbg0, ngath, nc, t1 - tdbg0, pcl[0], pcl[1], pcl[8], pcl[9], pcl[10], pcl[2], pcl[3], pcl[4], pcl[5], pcl[6], pcl[7]);
    }
  }
}

template <bool ST, int G, int MINB>
__global__ void __launch_bounds__(USH_T, MINB) k_usher_eval(const __grid_constant__ UsherArgs A, const __grid_constant__ PairTab pt, int nl)
{
  extern __shared__ __align__(16) unsigned char usm[];
  if (blockIdx.x == 0 && threadIdx.x == 0 && nl > 0) A.sc->nrounds_work++;
  usher_trials<ST, G, false>(A, pt, reinterpret_cast<UGroup *>(usm), nullptr, nl, &A.sc->trial_cursor, 0, 0, nullptr);
}

// ---- maxattempt > 1 ------------------------------------------------------------------------------
// (:912-1172): insertion slot k draws attempts from the RNG stream until one is accepted or
// maxattempt have failed, then slot k+1 continues where it stopped.  Given the accept flag of every
// possible attempt, the attempts actually made are a prefix of the stream: a slot boundary sits
// behind every accepted attempt and behind every maxattempt-th consecutive failure.
// nused = index of the slots-th boundary + 1.  One CTA of nt <= 1024 threads, chunked scans.
__device__ void usher_resolve(const obmd_trial *cur, int ntr, int slots, int M, StepScalars *sc, int *s_last,
                              int *s_cnt)
{
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (ntr + nt - 1) / nt, b = tid * per, e = min(b + per, ntr);
  int last = -1;
  for (int a = b; a < e; a++)
    if (__ldcg(&cur[a].accepted)) last = a;
  s_last[tid] = last;
  __syncthreads();
  if (tid == 0) {  // exclusive running maximum over the chunks
    int run = -1;
    for (int k = 0; k < nt; k++) {
      const int v = s_last[k];
      s_last[k] = run;
      run = max(run, v);
    }
  }
  __syncthreads();
  int la = s_last[tid], nb = 0;
  for (int a = b; a < e; a++) {
    const bool acc = __ldcg(&cur[a].accepted) != 0;
    nb += (acc || ((a - la) % M == 0)) ? 1 : 0;
    if (acc) la = a;
  }
  s_cnt[tid] = nb;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int k = 0; k < nt; k++) {
      const int v = s_cnt[k];
      s_cnt[k] = run;
      run += v;
    }
    if (run < slots) sc->nused = ntr;  // cannot happen for ntr == slots * M; keep the state sane
  }
  __syncthreads();
  int c0 = s_cnt[tid];
  if (c0 < slots && c0 + nb >= slots) {
    la = s_last[tid];
    for (int a = b; a < e; a++) {
      const bool acc = __ldcg(&cur[a].accepted) != 0;
      if (acc || ((a - la) % M == 0)) {
        if (++c0 == slots) {
          sc->nused = a + 1;
          break;
        }
      }
      if (acc) la = a;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) k_usher_resolve(const obmd_trial *cur, int ntr, int slots, int M,
                                                        StepScalars *sc)
{
  __shared__ int s_last[1024], s_cnt[1024];
  usher_resolve(cur, ntr, slots, M, sc, s_last, s_cnt);
}

// ---- diff: which trials changed the set of accepted atoms -----------------------------------------
// The accepted atom of trial t, if any, is spec entry spec_of_trial[t].  A trial whose latest
// evaluation disagrees with it (accepted flag or position; maxattempt > 1: only attempts the serial
// loop really makes, t < nused, can insert) retires the old entry, publishes a new one in the
// entry array and leaves a change record for the dirty test.  Entry order is irrelevant:
// contributions are ordered by local index when they are summed.
__device__ __forceinline__ void usher_diff_one(const UsherArgs &A, int t, int par)
{
  const obmd_trial *b = A.cur + t;
  const int newacc = __ldcg(&b->accepted) && (!A.limit || t < *(volatile int *) &A.sc->nused);
  const int eold = __ldcg(A.spec_of_trial + t);
  double o[3] = {0, 0, 0}, n[3];
  for (int d = 0; d < 3; d++) n[d] = __ldcg(&b->end[d]);
  bool ch = (eold >= 0) != (newacc != 0);
  if (eold >= 0) {
    for (int d = 0; d < 3; d++) o[d] = __ldcg(A.spec_pos + 3 * (size_t) eold + d);
    if (newacc && (o[0] != n[0] || o[1] != n[1] || o[2] != n[2])) ch = true;
  }
  if (!ch) return;
  ChgRec r;
  for (int d = 0; d < 3; d++) {
    r.o[d] = o[d];
    r.n[d] = n[d];
  }
  r.q = t;
  r.flags = (eold >= 0 ? 1 : 0) | (newacc ? 2 : 0);
  A.chg[atomicAdd(&A.sc->nchg2[par], 1)] = r;
  if (eold >= 0) A.spec_trial[eold] = 0x7fffffff;
  int enew = -1;
  if (newacc) {
    enew = atomicAdd(&A.sc->nspec, 1);
    if (enew >= A.spec_cap) {
      A.sc->usher_error = OBMD_ERR_CAPACITY + 1000;
      enew = -1;
    } else {
      for (int d = 0; d < 3; d++) A.spec_pos[3 * (size_t) enew + d] = n[d];
      A.spec_trial[enew] = t;
    }
  }
  A.spec_of_trial[t] = enew;
}

__global__ void k_usher_diff(const __grid_constant__ UsherArgs A)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < A.ntr) usher_diff_one(A, t, 0);
}

// ---- dirty test -----------------------------------------------------------------------------------
// all points involved lie inside the box, so one periodic shift suffices
__device__ __forceinline__ bool near_pt(const double *a, double bx, double by, double bz, const Grid &g,
                                        double rsqmax)
{
  double dx = a[0] - bx, dy = a[1] - by, dz = a[2] - bz;
  if (dy > g.half[1]) dy -= g.prd[1];
  else if (dy < -g.half[1]) dy += g.prd[1];
  if (dz > g.half[2]) dz -= g.prd[2];
  else if (dz < -g.half[2]) dz += g.prd[2];
  return dx * dx + dy * dy + dz * dz <= rsqmax;
}

// is coordinate v (or a periodic image) within rc of the interval [lo, hi]?
__device__ __forceinline__ bool axis_near(double v, double lo, double hi, double rc, double L, bool periodic)
{
  lo -= rc;
  hi += rc;
  if (v >= lo && v <= hi) return true;
  if (!periodic) return false;
  return (v + L >= lo && v + L <= hi) || (v - L >= lo && v - L <= hi);
}

__device__ __forceinline__ bool near_bbox(const double *a, const float *bb, const Grid &g, double rc)
{
  return axis_near(a[0], bb[0], bb[3], rc, 0.0, false) && axis_near(a[1], bb[1], bb[4], rc, g.prd[1], true) &&
      axis_near(a[2], bb[2], bb[5], rc, g.prd[2], true);
}

// A trial must be re-run when an earlier trial whose state changed has its old or new atom within
// the cutoff of any point the trial evaluated (a superset test with a small margin).
// kfirst[t] = index of the first such point: the re-run resumes there.  One warp per trial.
template <bool CACHED>  // CACHED: the inputs were written by an earlier kernel (L1 may hold them)
__device__ __forceinline__ void usher_dirty_one(const UsherArgs &A, int t, int par, int lane)
{
  const Grid &g = A.g;
  const double *mp = A.path + (size_t) t * (A.u.natt + 1) * 3;
  float bb[6];
  for (int d = 0; d < 6; d++) bb[d] = __ldcg(A.bbox + (size_t) t * 6 + d);
  const int np = __ldcg(A.npts + t), nchg = A.sc->nchg2[par];
  int kf = np;  // first affected point found so far
  // lanes prefilter 32 change records at a time against the trial's bounding box; every record that
  // passes is then tested against all evaluated points by the whole warp (points are read coalesced)
  for (int q0 = 0; q0 < nchg; q0 += 32) {
    const int qi = q0 + lane;
    ChgRec r;
    r.flags = 0;
    r.q = 1 << 30;
    for (int d = 0; d < 3; d++) r.o[d] = r.n[d] = 0.0;
    if (qi < nchg) {
      const ChgRec *src = A.chg + qi;
      if (CACHED) r = *src;
      else {
        for (int d = 0; d < 3; d++) {
          r.o[d] = __ldcg(&src->o[d]);
          r.n[d] = __ldcg(&src->n[d]);
        }
        r.q = __ldcg(&src->q);
        r.flags = __ldcg(&src->flags);
      }
    }
    const bool oldacc = r.q < t && (r.flags & 1) && near_bbox(r.o, bb, g, A.rc);
    const bool newacc = r.q < t && (r.flags & 2) && near_bbox(r.n, bb, g, A.rc);
    unsigned m = __ballot_sync(0xffffffffu, oldacc || newacc);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1u;
      double o[3], n[3];
      for (int d = 0; d < 3; d++) {
        o[d] = __shfl_sync(0xffffffffu, r.o[d], src);
        n[d] = __shfl_sync(0xffffffffu, r.n[d], src);
      }
      const bool oa = __shfl_sync(0xffffffffu, (int) oldacc, src) != 0;
      const bool na = __shfl_sync(0xffffffffu, (int) newacc, src) != 0;
      for (int k0 = 0; k0 < kf; k0 += 32) {  // kf stays warp-uniform: it is only set from ballots
        const int k = k0 + lane;
        bool hit = false;
        if (k < kf) {
          const double bx = __ldcg(mp + 3 * k), by = __ldcg(mp + 3 * k + 1), bz = __ldcg(mp + 3 * k + 2);
          hit = (oa && near_pt(o, bx, by, bz, g, A.rsqmax)) || (na && near_pt(n, bx, by, bz, g, A.rsqmax));
        }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm) {
          kf = k0 + __ffs(hm) - 1;
          break;
        }
      }
    }
  }
  if (kf < np && lane == 0) {
    A.kfirst[t] = kf;
    A.list[atomicAdd(&A.sc->ndirty2[par], 1)] = t;
  }
}

__global__ void k_usher_dirty(const __grid_constant__ UsherArgs A)
{
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= A.ntr || A.sc->nchg2[0] == 0) return;
  usher_dirty_one<true>(A, t, 0, threadIdx.x & 31);
}

// ---- re-run rounds, all in one launch ---------------------------------------------------------------
#define SPEC_SM 1024  // accepted-atom entries cached in shared memory (more are read from global memory)
#define SPEC_H 2048   // heads of their per-cell chains (cell index modulo SPEC_H)
typedef UGroupT<true> UGroupFix;

// Dirty test of a re-run round, in three sweeps separated by cluster barriers.  The change records
// are few (one per re-run trial at most), so they are staged in shared memory once per CTA.
//  1. every THREAD takes a trial: one round trip for its bounding box, then the records at shared-memory
//     speed; a (trial, record) pair that passes the box test goes to a queue;
//  2. one warp per queued pair tests all evaluated points of the trial at once and lowers kfirst[t];
//  3. the trials with kfirst[t] < npts[t] form the list of the next round.
__device__ void usher_dirty_sweep1(const UsherArgs &A, int nchg, ChgRec *srec, int gtid, int gthreads)
{
  const Grid &g = A.g;
  for (int k = threadIdx.x; k < nchg * (int) (sizeof(ChgRec) / 4); k += blockDim.x)
    reinterpret_cast<int *>(srec)[k] = __ldcg(reinterpret_cast<const int *>(A.chg) + k);
  __syncthreads();
  for (int t = gtid; t < A.ntr; t += gthreads) {
    float bb[6];
    for (int d = 0; d < 6; d++) bb[d] = __ldcg(A.bbox + (size_t) t * 6 + d);
    A.kfirst[t] = 0x7fffffff;
    for (int q = 0; q < nchg; q++) {
      const ChgRec &r = srec[q];
      if (r.q >= t) continue;
      const bool oa = (r.flags & 1) && near_bbox(r.o, bb, g, A.rc);
      const bool na = (r.flags & 2) && near_bbox(r.n, bb, g, A.rc);
      if (!(oa || na)) continue;
      const int k = atomicAdd(&A.sc->npairs, 1);
      if (k < A.pair_cap) A.pairs[k] = make_int2(t, q);
      else A.sc->usher_error = OBMD_ERR_CAPACITY + 1000;
    }
  }
}

__device__ void usher_dirty_sweep2(const UsherArgs &A, int gw, int nwarp, int lane)
{
  const Grid &g = A.g;
  const int npairs = min(*(volatile int *) &A.sc->npairs, A.pair_cap);
  for (int k = gw; k < npairs; k += nwarp) {
    const int2 pr = __ldcg(A.pairs + k);
    const int t = pr.x;
    const ChgRec *src = A.chg + pr.y;
    double o[3], n[3];
    for (int d = 0; d < 3; d++) {
      o[d] = __ldcg(&src->o[d]);
      n[d] = __ldcg(&src->n[d]);
    }
    const int flags = __ldcg(&src->flags);
    const int np = __ldcg(A.npts + t);
    const double *mp = A.path + (size_t) t * (A.u.natt + 1) * 3;
    for (int k0 = 0; k0 < np; k0 += 32) {
      const int kk = k0 + lane;
      bool hit = false;
      if (kk < np) {
        const double bx = __ldcg(mp + 3 * kk), by = __ldcg(mp + 3 * kk + 1), bz = __ldcg(mp + 3 * kk + 2);
        hit = ((flags & 1) && near_pt(o, bx, by, bz, g, A.rsqmax)) || ((flags & 2) && near_pt(n, bx, by, bz, g, A.rsqmax));
      }
      const unsigned hm = __ballot_sync(0xffffffffu, hit);
      if (hm) {
        if (lane == 0) atomicMin(&A.kfirst[t], k0 + __ffs(hm) - 1);
        break;
      }
    }
  }
}

__device__ void usher_dirty_sweep3(const UsherArgs &A, int par, int gtid, int gthreads)
{
  for (int t = gtid; t < A.ntr; t += gthreads)
    if (__ldcg(A.kfirst + t) < __ldcg(A.npts + t)) A.list[atomicAdd(&A.sc->ndirty2[par], 1)] = t;
}

// One thread-block cluster.  Round r >= 1: re-run the trials the dirty test of round r-1 listed (one
// warp per trial: the rounds are a handful of dependent 40-iteration chains, latency is what
// counts), then resolve / diff / dirty as after round 0, until a round changes nothing.  Finally the
// accepted trials are compacted in order into acc_list (they become local indices nbase, nbase+1, ...
// and tags maxtag+1, ...).
template <bool ST, int FIX_CL>
__global__ void __launch_bounds__(FIX_T, 1)
    k_usher_fix(const __grid_constant__ UsherArgs A, const __grid_constant__ PairTab pt)
{
  extern __shared__ __align__(16) unsigned char usm[];
  __shared__ int s_a[FIX_T], s_b[FIX_T];
  __shared__ int s_base;
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int) cl.block_rank();
  const int tid = threadIdx.x;
  const int nwarp = FIX_CL * (FIX_T / 32), gw = rank * (FIX_T / 32) + (tid >> 5);
  const int gtid = rank * FIX_T + tid, gthreads = FIX_CL * FIX_T;
  StepScalars *sc = A.sc;
  volatile int *ndirty2 = sc->ndirty2, *nchg2 = sc->nchg2;
  UGroupFix *groups = reinterpret_cast<UGroupFix *>(usm);
  constexpr size_t GROUP_BYTES = sizeof(UGroupFix) * (FIX_T / 32);
  double *ssx = reinterpret_cast<double *>(usm + GROUP_BYTES), *ssy = ssx + SPEC_SM, *ssz = ssy + SPEC_SM;
  int *sstr = reinterpret_cast<int *>(ssz + SPEC_SM), *scell = sstr + SPEC_SM, *snext = scell + SPEC_SM;
  int *shead = snext + SPEC_SM;
  unsigned long long tdbg[5] = {0, 0, 0, 0, 0};
  auto now = []() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
  };
  for (int round = 1;; round++) {
    const int par = round & 1, pprev = par ^ 1;
    const int nd = ndirty2[pprev];
    if (nd == 0) break;
    if (USHER_DEBUG && A.debug) tdbg[0] = now();
    if (round > A.ntr + 2) {  // cannot happen: trial t settles after at most t + 1 rounds
      if (gtid == 0) sc->usher_error = OBMD_ERR_STATE;
      break;
    }
    // the atoms accepted so far, into shared memory
    SpecSm sp;
    sp.n_all = min(*(volatile int *) &sc->nspec, A.spec_cap);
    sp.n_sm = min(sp.n_all, SPEC_SM);
    sp.x = ssx;
    sp.y = ssy;
    sp.z = ssz;
    sp.tr = sstr;
    sp.cell = scell;
    sp.next = snext;
    sp.head = shead;
    sp.hmask = SPEC_H - 1;
    for (int h = tid; h < SPEC_H; h += FIX_T) shead[h] = -1;
    __syncthreads();
    for (int e = tid; e < sp.n_sm; e += FIX_T) {
      const double xs = __ldcg(A.spec_pos + 3 * (size_t) e), ys = __ldcg(A.spec_pos + 3 * (size_t) e + 1);
      const double zs = __ldcg(A.spec_pos + 3 * (size_t) e + 2);
      const int tr = __ldcg(A.spec_trial + e);
      ssx[e] = xs;
      ssy[e] = ys;
      ssz[e] = zs;
      sstr[e] = tr;
      if (tr != 0x7fffffff) {  // live: thread it into the chain of its cell (the cell a trial's gather computes)
        const int c2 = (cell_cx(A.g, xs) * A.g.ny + cell_cy(A.g, ys)) * A.g.nz + cell_cz(A.g, zs);
        scell[e] = c2;
        snext[e] = atomicExch(&shead[c2 & (SPEC_H - 1)], e);
      }
    }
    __syncthreads();
    if (USHER_DEBUG && A.debug) tdbg[1] = now();
    // listed trial li runs on CTA li % FIX_CL: a round of a few dozen trials is spread over all SMs of
    // the cluster (each chain then has a scheduler and most of an L1 to itself)
    usher_trials<ST, 32, true>(A, pt, groups, A.list, nd, nullptr, (tid >> 5) * FIX_CL + rank, nwarp, &sp);
    if (gtid == 0) sc->nrounds_work++;
    __threadfence();
    cl.sync();
    if (USHER_DEBUG && A.debug) tdbg[2] = now();
    if (A.limit) {
      if (rank == 0) usher_resolve(A.cur, A.ntr, A.u.slots, A.u.maxatt, sc, s_a, s_b);
      __threadfence();
      cl.sync();
      for (int t = gtid; t < A.ntr; t += gthreads) usher_diff_one(A, t, par);
    } else {
      for (int k = gtid; k < nd; k += gthreads) usher_diff_one(A, __ldcg(A.list + k), par);
    }
    if (gtid == 0) {  // the counters of the previous round have been consumed by everybody
      ndirty2[pprev] = 0;
      nchg2[pprev] = 0;
    }
    __threadfence();
    cl.sync();
    if (USHER_DEBUG && A.debug) tdbg[3] = now();
    const int nchg = nchg2[par];
    if (nchg > 0) {
      if ((size_t) nchg * sizeof(ChgRec) <= GROUP_BYTES) {
        usher_dirty_sweep1(A, nchg, reinterpret_cast<ChgRec *>(usm), gtid, gthreads);
        __threadfence();
        cl.sync();
        usher_dirty_sweep2(A, gw, nwarp, tid & 31);
        __threadfence();
        cl.sync();
        usher_dirty_sweep3(A, par, gtid, gthreads);
        if (gtid == 0) sc->npairs = 0;
      } else
        for (int t = gw; t < A.ntr; t += nwarp) usher_dirty_one<false>(A, t, par, tid & 31);
    }
    __threadfence();
    cl.sync();
    if (USHER_DEBUG && A.debug && gtid == 0) {
      tdbg[4] = now();
      printf("fix round %d: nd %d nchg %d nspec %d | load %llu eval %llu diff %llu dirty %llu ns\n", round, nd, nchg,
             sp.n_all, tdbg[1] - tdbg[0], tdbg[2] - tdbg[1], tdbg[3] - tdbg[2], tdbg[4] - tdbg[3]);
    }
  }
  // ordered compaction of the accepted trials (CTA 0): every thread owns a contiguous run of trials,
  // a block scan of the counts gives its offset
  if (rank == 0) {
    const int nlim = A.limit ? *(volatile int *) &sc->nused : A.ntr;
    const int per = (nlim + FIX_T - 1) / FIX_T, tb = tid * per, te = min(tb + per, nlim);
    int cnt = 0;
    for (int t = tb; t < te; t++) cnt += (__ldcg(A.spec_of_trial + t) >= 0) ? 1 : 0;
    s_a[tid] = cnt;
    __syncthreads();
    if (tid < 32) {  // exclusive scan of FIX_T counts by one warp, a contiguous run per lane
      constexpr int RUN = FIX_T / 32;
      int s = 0;
      for (int k = 0; k < RUN; k++) s += s_a[tid * RUN + k];
      int inc = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (tid >= d) inc += v;
      }
      int run = inc - s;
      for (int k = 0; k < RUN; k++) {
        const int v = s_a[tid * RUN + k];
        s_b[tid * RUN + k] = run;
        run += v;
      }
      if (tid == 31) s_base = inc;
    }
    __syncthreads();
    int off = s_b[tid];
    for (int t = tb; t < te; t++)
      if (__ldcg(A.spec_of_trial + t) >= 0) A.acc_list[off++] = t;
    if (tid == 0) sc->naccept = s_base;
  }
}

// create_atom + tag/mask/image/v assignment (:1121-1150, atom_vec.cpp:1592-1612), appended after
// every other extra atom; the new atoms are then threaded into the per-cell lists the pair kernel
// walks (fix_link_extras).
__global__ void k_usher_commit(Planes p, Extras ex, int tag0, FixDev fd, obmd_trial *__restrict__ cur,
                               const int *__restrict__ acc_list, int nacc, double *fx, double *fy, double *fz)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nacc) {
    obmd_trial &r = cur[acc_list[k]];
    int s = ex.base + ex.narr + ex.nghost + ex.nins + k;  // ex.nins: atoms of an earlier pass
    p.x[s] = r.end[0];
    p.y[s] = r.end[1];
    p.z[s] = r.end[2];
    p.vx[s] = p.vy[s] = p.vz[s] = 0.0;
    p.type[s] = fd.ntype;
    p.tag[s] = tag0 + k + 1;
    p.image[s] = IMG0;
    p.lidx[s] = s;  // appended: local index == slot (serial runs: nlocal-after-deletion + k)
    fx[s] = fy[s] = fz[s] = 0.0;
    r.tag = tag0 + k + 1;
  }
}

// ---- host side ---------------------------------------------------------------------------------

// Device buffers of the USHER trials.  Regrowing them costs a dozen cudaFree/cudaMalloc pairs (each
// an implicit device synchronisation, ~150 ms in all) in the middle of a step, so they start at 64 Ki
// trials (80 MB; the 4 M-atom workload hovers around 16 k trials per step) and double; the pinned host copy of the trial records is only made when a caller asks for them.
static int ensure_trials(obmd_ctx *c, int ntr, int natt)
{
  if (ntr <= c->tr_cap && natt <= c->tr_natt) return OBMD_OK;
  int cap = c->tr_cap;
  if (ntr > cap) {
    cap = cap < 65536 ? 65536 : cap;
    while (cap < ntr) cap *= 2;
  }
  int na = natt > c->tr_natt ? natt : c->tr_natt;
  cudaFree(c->tr_d); cudaFree(c->tr_path); cudaFree(c->tr_npts); cudaFree(c->tr_list);
  cudaFree(c->tr_kfirst); cudaFree(c->acc_list); cudaFree(c->tr_bbox); cudaFree(c->tr_chg);
  cudaFree(c->spec_trial); cudaFree(c->spec_of_trial); cudaFree(c->spec_pos); cudaFree(c->tr_pairs);
  CK(cudaMalloc(&c->tr_pairs, sizeof(int) * 2 * 8 * (size_t) cap));
  CK(cudaMalloc(&c->tr_bbox, sizeof(float) * 6 * cap));
  CK(cudaMalloc(&c->tr_d, sizeof(obmd_trial) * cap));
  CK(cudaMalloc(&c->tr_path, sizeof(double) * 3 * (size_t) cap * (na + 1)));
  CK(cudaMalloc(&c->tr_npts, sizeof(int) * cap));
  CK(cudaMalloc(&c->tr_list, sizeof(int) * cap));
  CK(cudaMalloc(&c->tr_kfirst, sizeof(int) * cap));
  CK(cudaMalloc(&c->tr_chg, sizeof(ChgRec) * cap));
  CK(cudaMalloc(&c->acc_list, sizeof(int) * cap));
  // accepted-atom entries are append-only within a call: one per change of a trial's accepted state
  CK(cudaMalloc(&c->spec_trial, sizeof(int) * 2 * cap));
  CK(cudaMalloc(&c->spec_pos, sizeof(double) * 3 * 2 * (size_t) cap));
  CK(cudaMalloc(&c->spec_of_trial, sizeof(int) * cap));
  c->tr_cap = cap;
  c->tr_natt = na;
  return OBMD_OK;
}

int usher_trials_to_host(obmd_ctx *c, obmd_trial *dst, int ncopy)
{
  if (ncopy <= 0) return OBMD_OK;
  if (ncopy > c->tr_h_cap) {
    if (c->tr_h) cudaFreeHost(c->tr_h);
    c->tr_h = nullptr;
    c->tr_h_cap = 0;
    int cap = ncopy < 4096 ? 4096 : 2 * ncopy;
    CK(cudaMallocHost(&c->tr_h, sizeof(obmd_trial) * cap));
    c->tr_h_cap = cap;
  }
  CK(cudaMemcpyAsync(c->tr_h, c->tr_d, sizeof(obmd_trial) * ncopy, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  memcpy(dst, c->tr_h, sizeof(obmd_trial) * ncopy);
  return OBMD_OK;
}

template <bool ST, int G, int MINB>
static int launch_eval(obmd_ctx *c, const UsherArgs &A, int ntr)
{
  const size_t smem = sizeof(UGroup) * (USH_T / 32) * (32 / G);
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    CK(cudaFuncSetAttribute((k_usher_eval<ST, G, MINB>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    attr_done = true;
  }
  const int groups_per_cta = (USH_T / 32) * (32 / G);
  int grid = cdiv(ntr, groups_per_cta);
  const int full = MINB * NSM;
  if (grid > full) grid = full;
  LAUNCH((k_usher_eval<ST, G, MINB>), grid, USH_T, smem, A, c->pt, ntr);
  return OBMD_OK;
}

template <bool ST, int CL>
static int launch_fix(obmd_ctx *c, const UsherArgs &A)
{
  const size_t smem = sizeof(UGroupFix) * (FIX_T / 32) + SPEC_SM * (3 * sizeof(double) + 3 * sizeof(int)) +
      SPEC_H * sizeof(int);
  static bool attr_done = false;
  if (!attr_done) {
    CK(cudaFuncSetAttribute((k_usher_fix<ST, CL>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CL, 1, 1);
  cfg.blockDim = dim3(FIX_T, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  if (CL > 8) {  // more than 8 CTAs per cluster is a non-portable size: opt in once
    static bool np_done = false;
    if (!np_done) {
      CK(cudaFuncSetAttribute((k_usher_fix<ST, CL>), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      np_done = true;
    }
  }
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, (k_usher_fix<ST, CL>), A, c->pt));
  c->launches++;
  return OBMD_OK;
}

// One pass of try_inserting for ntr_loc trials on this rank (maxattempt == 1: both sides together;
// maxattempt > 1: one side).  Enqueues round 0, the first diff / dirty test and the fix-up kernel,
// then reads the outcome back (the only host round trip of the pass).
int usher_pass(obmd_ctx *c, const UsherCfg &u, int ntr_loc, int limit, int maxtag, int *nacc_out, int *nused_out,
               int *rounds_out)
{
  int rc;
  if ((rc = ensure_trials(c, ntr_loc, c->fa.nattempt))) return rc;
  UsherArgs A;
  memset(&A, 0, sizeof(A));
  A.p = c->cur;
  A.g = c->grid;
  A.fd = c->fd;
  A.u = u;
  A.ex = extras_of(c);
  // Order of the USHER energy sums.  Serial runs: ascending LAMMPS local index, the reference's loop
  // order (fix_obmd_merged.cpp:1790-1807).  Slab runs (and serial runs with obmd_set_usher_order(1)):
  // ascending tag, which does not depend on the decomposition -- any number of ranks then produces
  // bit-identical trial outcomes, equal to the serial tag-ordered run.  Atoms created earlier in the
  // call come after every existing atom in both orders, among themselves in trial order.
  const bool by_tag = c->nranks > 1 || c->usher_by_tag;
  A.keys = by_tag ? c->cur.tag : c->cur.lidx;
  A.spec_key0 = by_tag ? maxtag + 1 + u.tg0 : u.nbase;
  A.cell_start = c->cell_start;
  A.extra_head = c->extra_head;
  A.extra_next = c->extra_next;
  A.spec_trial = c->spec_trial;
  A.spec_of_trial = c->spec_of_trial;
  A.spec_pos = c->spec_pos;
  A.spec_cap = 2 * c->tr_cap;
  A.cur = c->tr_d;
  A.path = c->tr_path;
  A.npts = c->tr_npts;
  A.bbox = c->tr_bbox;
  A.kfirst = c->tr_kfirst;
  A.list = c->tr_list;
  A.pairs = reinterpret_cast<int2 *>(c->tr_pairs);
  A.pair_cap = 8 * c->tr_cap;
  A.chg = c->tr_chg;
  A.acc_list = c->acc_list;
  A.sc = c->sc_d;
  A.ntr = ntr_loc;
  A.limit = limit;
  // the bounding boxes are fp32 (rounded outwards); 1e-3 covers the prefilter's own rounding
  A.rc = c->pt.cutmax * (1.0 + 1e-9) + 1e-3;
  A.rsqmax = c->pt.cutmax * c->pt.cutmax * (1.0 + 1e-9);
  for (int d = 0; d < 3; d++) {
    A.wf[d] = (float) c->grid.w[d];
    A.pf[d] = (float) c->grid.prd[d];
    A.ipf[d] = 1.0f / A.pf[d];
  }
  A.cut32 = (float) c->pt.cutmax * (float) c->pt.cutmax * 1.001f;
  {  // OBMD_B200_DEBUG_FIX=N (builds with -DUSHER_DEBUG=1): phase times of the fix-up kernel from call N on
    static long long ncall = 0;
    const char *e = getenv("OBMD_B200_DEBUG_FIX");
    A.debug = (e && ++ncall >= atoll(e)) ? 1 : 0;
  }
  // nchg2, ndirty2, naccept, nrounds_work, trial_cursor, nused, nused_prev, nspec, npairs
  CK(cudaMemsetAsync(&c->sc_d->nchg2[0], 0, 11 * sizeof(int), c->stream));
  CK(cudaMemsetAsync(c->spec_of_trial, 0xff, sizeof(int) * ntr_loc, c->stream));
  const bool st = (c->ntypes == 1);
  {
    ScopedTimer t5(c, "usher_eval0");
    // one warp per trial; 8 resident CTAs per SM (64 registers) when there are more trials than the
    // 6-CTA build (80 registers) can hold at once; OBMD_B200_USH_G=16 packs two trials per warp
    const int G = c->ush_group;
    const bool dense = c->ush_ctas_per_sm >= 8;
    if (st) {
      if (G == 16) rc = launch_eval<true, 16, 6>(c, A, ntr_loc);
      else if (dense) rc = launch_eval<true, 32, 8>(c, A, ntr_loc);
      else if (c->ush_ctas_per_sm == 5) rc = launch_eval<true, 32, 5>(c, A, ntr_loc);
      else if (c->ush_ctas_per_sm == 7) rc = launch_eval<true, 32, 7>(c, A, ntr_loc);
      else rc = launch_eval<true, 32, 6>(c, A, ntr_loc);
    } else {
      if (G == 16) rc = launch_eval<false, 16, 6>(c, A, ntr_loc);
      else if (dense) rc = launch_eval<false, 32, 8>(c, A, ntr_loc);
      else if (c->ush_ctas_per_sm == 5) rc = launch_eval<false, 32, 5>(c, A, ntr_loc);
      else if (c->ush_ctas_per_sm == 7) rc = launch_eval<false, 32, 7>(c, A, ntr_loc);
      else rc = launch_eval<false, 32, 6>(c, A, ntr_loc);
    }
    if (rc) return rc;
  }
  // Round 0 fills the machine; obmd_run hooks in here to start the pair kernel on the side stream
  // right behind it (OBMD_B200_PAIR_FIRST=0).
  if (c->round0_hook) {
    int (*hook)(obmd_ctx *) = c->round0_hook;
    c->round0_hook = nullptr;
    if ((rc = hook(c))) return rc;
  }
  {
    ScopedTimer t6(c, "usher_fix");
    if (limit) LAUNCH(k_usher_resolve, 1, 1024, 0, c->tr_d, ntr_loc, u.slots, u.maxatt, c->sc_d);
    LAUNCH(k_usher_diff, cdiv(ntr_loc, 256), 256, 0, A);
    LAUNCH(k_usher_dirty, cdiv(ntr_loc, 4), 128, 0, A);
    // one cluster of 8 CTAs (OBMD_B200_FIX_CL=1: a single CTA)
    if (c->fix_cluster == 1) rc = st ? launch_fix<true, 1>(c, A) : launch_fix<false, 1>(c, A);
    else if (c->fix_cluster == 16) rc = st ? launch_fix<true, 16>(c, A) : launch_fix<false, 16>(c, A);
    else rc = st ? launch_fix<true, 8>(c, A) : launch_fix<false, 8>(c, A);
    if (rc) return rc;
  }
  if ((rc = read_scalars(c))) return rc;
  const int err = c->sc_h->usher_error;
  NEED(err != OBMD_ERR_USHER, OBMD_ERR_USHER,
       "fix obmd: USHER force vanished with E above target (the reference would not terminate)");
  NEED(err != OBMD_ERR_CAPACITY, OBMD_ERR_CAPACITY,
       "fix obmd: more than %d neighbours near the cutoff of a USHER point (or %d atoms in its stencil)",
       USH_CAP, USH_CAND);
  NEED(err != OBMD_ERR_CAPACITY + 1000, OBMD_ERR_CAPACITY, "fix obmd: too many accepted-state changes in one call");
  NEED(err != OBMD_ERR_STATE, OBMD_ERR_STATE, "fix obmd: USHER commit did not converge");
  *nacc_out = c->sc_h->naccept;
  *nused_out = limit ? c->sc_h->nused : ntr_loc;
  *rounds_out = c->sc_h->nrounds_work;
  return OBMD_OK;
}

int usher_commit(obmd_ctx *c, int tag0, int nacc)
{
  LAUNCH(k_usher_commit, cdiv(nacc, 128), 128, 0, c->cur, extras_of(c), tag0, c->fd, c->tr_d, c->acc_list, nacc,
         c->fx, c->fy, c->fz);
  return OBMD_OK;
}
