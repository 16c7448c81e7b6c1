// FixObmdMerged on the device (fix_obmd_merged.cpp) plus the FixNVE / Domain::pbc kernels that
// keep the state device-resident.  Compiled with -fmad=false: the bit-exact parts (USHER sums,
// trial coordinates, weights, integrator) must not be contracted into FMAs, because the
// reference is an x86-64 build without FMA.
//
//   pre_exchange (:550-633)      -> k_scan_atoms, k_delete_fixup, k_usher_eval/_diff/_dirty/_acc,
//                                   k_usher_commit
//   post_force   (:1229-1240)    -> k_weight_sums, k_apply_forces
//   FixNVE       (fix_nve.cpp)   -> k_nve_initial, k_nve_final
//   Domain::pbc  (domain.cpp)    -> k_pbc
#include "ctx.h"

#define EPSILON_FIX 1.0e-6   // fix_obmd_merged.cpp:62
#define EPSILON_DPD 1.0e-10  // pair_dpd.cpp:36
#define MY_PI 3.14159265358979323846

// =====================================================================================
// integrator and periodic wrap
// =====================================================================================

// fix_nve.cpp:134-145: dtfm = dtf/mass[type]; v += dtfm*f; x += dtv*v
__global__ void k_nve_initial(Planes p, const double *__restrict__ fx, const double *__restrict__ fy,
                              const double *__restrict__ fz, int n, double dtv, double dtf, PairTab pt)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double dtfm = dtf / pt.mass[p.type[i]];
  double vx = p.vx[i] + dtfm * fx[i];
  double vy = p.vy[i] + dtfm * fy[i];
  double vz = p.vz[i] + dtfm * fz[i];
  p.vx[i] = vx;
  p.vy[i] = vy;
  p.vz[i] = vz;
  p.x[i] += dtv * vx;
  p.y[i] += dtv * vy;
  p.z[i] += dtv * vz;
}

// fix_nve.cpp:236-244
__global__ void k_nve_final(Planes p, const double *__restrict__ fx, const double *__restrict__ fy,
                            const double *__restrict__ fz, int n, double dtf, PairTab pt)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double dtfm = dtf / pt.mass[p.type[i]];
  p.vx[i] += dtfm * fx[i];
  p.vy[i] += dtfm * fy[i];
  p.vz[i] += dtfm * fz[i];
}

// domain.cpp:529-640 for boundary f p p: wrap y and z, adjust image flags
__device__ __forceinline__ void wrap_dim(double &x, int &image, double lo, double hi, double period,
                                         int shift)
{
  if (x < lo) {
    x += period;
    int idim = (image >> shift) & IMGMASK;
    int other = image ^ (idim << shift);
    idim = (idim - 1) & IMGMASK;
    image = other | (idim << shift);
  }
  if (x >= hi) {
    x -= period;
    x = fmax(x, lo);
    int idim = (image >> shift) & IMGMASK;
    int other = image ^ (idim << shift);
    idim = (idim + 1) & IMGMASK;
    image = other | (idim << shift);
  }
}

__global__ void k_pbc(Planes p, int n, Grid g)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double y = p.y[i], z = p.z[i];
  if (y < g.lo[1] || y >= g.hi[1] || z < g.lo[2] || z >= g.hi[2]) {
    int im = p.image[i];
    wrap_dim(y, im, g.lo[1], g.hi[1], g.prd[1], IMGBITS);
    wrap_dim(z, im, g.lo[2], g.hi[2], g.prd[2], IMG2BITS);
    p.y[i] = y;
    p.z[i] = z;
    p.image[i] = im;
  }
}

int fix_nve_initial(obmd_ctx *c)
{
  NEED(c->n > 0 && c->dt > 0.0, OBMD_ERR_STATE, "obmd_nve_initial: atoms and timestep must be set");
  ScopedTimer t(c, "nve");
  LAUNCH(k_nve_initial, cdiv(c->n, 256), 256, 0, c->cur, c->fx, c->fy, c->fz, c->n, c->dt,
         0.5 * c->dt, c->pt);
  c->cells_valid = false;
  return OBMD_OK;
}

int fix_nve_final(obmd_ctx *c)
{
  NEED(c->n > 0 && c->dt > 0.0, OBMD_ERR_STATE, "obmd_nve_final: atoms and timestep must be set");
  ScopedTimer t(c, "nve");
  LAUNCH(k_nve_final, cdiv(c->n, 256), 256, 0, c->cur, c->fx, c->fy, c->fz, c->n, 0.5 * c->dt, c->pt);
  return OBMD_OK;
}

int fix_pbc(obmd_ctx *c)
{
  ScopedTimer t(c, "pbc");
  LAUNCH(k_pbc, cdiv(c->n, 256), 256, 0, c->cur, c->n, c->grid);
  return OBMD_OK;
}

// =====================================================================================
// pre_exchange, part 1: scan + deletion bookkeeping
// =====================================================================================

// One pass over the live atoms: Group::count for both buffers (group.cpp:793-811, closed block on
// the not-yet-wrapped coordinates, Q9), find_maxid (:1860-1875), and the list of survivors whose
// local index is >= nprime (they will be moved into the holes the deletions leave).
__global__ void k_scan_atoms(Planes p, int nlive, FixDev fd, StepScalars *sc, int *__restrict__ tail_key,
                             int *__restrict__ tail_val, int tail_cap)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int cl = 0, cr = 0, mt = 0;
  if (i < nlive) {
    double x = p.x[i], y = p.y[i], z = p.z[i];
    cl = block_inside(fd.region[0], x, y, z);
    cr = block_inside(fd.region[1], x, y, z);
    mt = p.tag[i];
    int li = p.lidx[i];
    if (li >= nlive) {
      int k = atomicAdd(&sc->ntail, 1);
      if (k < tail_cap) {
        tail_key[k] = li;
        tail_val[k] = i;
      }
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    cl += __shfl_down_sync(0xffffffffu, cl, d);
    cr += __shfl_down_sync(0xffffffffu, cr, d);
    mt = max(mt, __shfl_down_sync(0xffffffffu, mt, d));
  }
  if ((threadIdx.x & 31) == 0) {
    if (cl) atomicAdd(&sc->cnt_left, cl);
    if (cr) atomicAdd(&sc->cnt_right, cr);
    if (mt) atomicMax(&sc->maxtag, mt);
  }
}

// In-block bitonic sort (descending by key) of npad = power of two entries in global memory.
__device__ void bitonic_desc(int *key, int *val, int npad)
{
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          bool desc = ((i & k) == 0);
          int a = key[i], b = key[ixj];
          if (desc ? (a < b) : (a > b)) {
            key[i] = b;
            key[ixj] = a;
            int t = val[i];
            val[i] = val[ixj];
            val[ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  }
}

// try_deleting (:639-865), ATOM mode.  The parked atoms [nlive, nlive+ndel) are the marked set.
//  * momentum of the deleted atoms, accumulated in the reference's order (descending local index,
//    :827-837) so vnewl/vnewr are bit-identical;
//  * the reverse loop of AtomVec::copy(nlocal-1, i) (:827-841) expressed on local indices.  Walking
//    the positions downwards, every survivor is enqueued and every hole dequeues the oldest entry
//    and re-enqueues it (the "current last atom" is always the oldest survivor not yet settled), so
//    the entry created at position p settles at p iff p < nprime.  Only the ndel positions
//    >= nprime need the walk: src[k] = storage slot of the survivor that ends up in the k-th
//    effective hole (holes counted downwards, skipping the run of holes at the very top).
__global__ void k_delete_fixup(Planes p, int nlive, int ndel, int npad_del, int npad_tail, FixDev fd,
                               PairTab pt, StepScalars *sc, int *del_key, int *del_val, int *tail_key,
                               int *tail_val, int *del_src)
{
  for (int k = threadIdx.x; k < npad_del; k += blockDim.x) {
    if (k < ndel) {
      del_key[k] = p.lidx[nlive + k];
      del_val[k] = nlive + k;
    } else {
      del_key[k] = -1;
      del_val[k] = -1;
    }
  }
  int ntail = sc->ntail;
  for (int k = ntail + threadIdx.x; k < npad_tail; k += blockDim.x) {
    tail_key[k] = -1;
    tail_val[k] = -1;
  }
  __syncthreads();
  bitonic_desc(del_key, del_val, npad_del);
  bitonic_desc(tail_key, tail_val, npad_tail);
  if (threadIdx.x == 0) {
    const int n = nlive + ndel;  // nlocal before the deletion; nprime = nlive
    double mid = 0.5 * (fd.boxhi[0] + fd.boxlo[0]);
    double vl0 = 0, vl1 = 0, vl2 = 0, vr0 = 0, vr1 = 0, vr2 = 0;
    for (int k = 0; k < ndel; k++) {
      int s = del_val[k];
      double m = pt.mass[p.type[s]];
      if (p.x[s] < mid) {
        vl0 += m * p.vx[s];
        vl1 += m * p.vy[s];
        vl2 += m * p.vz[s];
      } else {
        vr0 += m * p.vx[s];
        vr1 += m * p.vy[s];
        vr2 += m * p.vz[s];
      }
    }
    int z = 0;
    while (z < ndel && del_key[z] == n - 1 - z) z++;
    const int heff = ndel - z;
    int is = 0, id = z, bad = 0;
    for (int k = 0; k < heff; k++) {
      int pos = n - 1 - z - k;
      if (is < ntail && tail_key[is] == pos) del_src[k] = tail_val[is++];
      else if (id < ndel && del_key[id] == pos) {
        del_src[k] = del_src[id - z];
        id++;
      } else
        bad = 1;
    }
    if (is != ntail) bad = 1;
    for (int k = 0; k < heff && !bad; k++) {
      int h = del_key[z + k];
      if (h < nlive) p.lidx[del_src[k]] = h;
    }
    if (bad) sc->usher_error = OBMD_ERR_STATE;  // bookkeeping invariant violated
    sc->vnewl[0] = vl0;
    sc->vnewl[1] = vl1;
    sc->vnewl[2] = vl2;
    sc->vnewr[0] = vr0;
    sc->vnewr[1] = vr1;
    sc->vnewr[2] = vr2;
  }
}

static int next_pow2(int v)
{
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static int ranpark_skip(int seed, long long n)  // random_park.cpp:43-50 applied n times
{
  unsigned long long m = 2147483647ULL, a = 16807ULL, r = 1ULL;
  while (n > 0) {
    if (n & 1) r = (r * a) % m;
    a = (a * a) % m;
    n >>= 1;
  }
  return (int) (((unsigned long long) seed * r) % m);
}

// =====================================================================================
// pre_exchange, part 2: batched USHER (try_inserting :869-1225, usher :1518-1616,
// energy :1774-1810, PairDPD::single pair_dpd.cpp:401-418)
// =====================================================================================

#define USH_WARPS 4
#define USH_CAP 192  // contributions per energy evaluation held per warp

struct UsherCfg {
  int ntrL, ntr;      // trials for the left region come first (maxattempt == 1)
  int seed0;          // RanPark state before the first trial draw
  int nbase;          // local index of the first inserted atom (= nlocal after deletion)
  int nacc_prev;      // accepted trials of the previous round (acc_list entries)
  int natt;
};

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

__device__ __forceinline__ double ranpark_uniform(int &seed)
{
  int k = seed / 127773;
  seed = 16807 * (seed - k * 127773) - 2836 * k;
  if (seed < 0) seed += 2147483647;
  return (1.0 / 2147483647) * seed;
}

struct WarpHits {
  int key[USH_CAP];
  double e[USH_CAP], fx[USH_CAP], fy[USH_CAP], fz[USH_CAP];
  int skey[USH_CAP];  // rank-sorted order: slot -> source
};

// contribution of one neighbour at (xj,yj,zj) of type tj to the trial point: literal energy()
// body (:1791-1806) + PairDPD::single
__device__ __forceinline__ bool usher_contrib(const double c[3], double xj, double yj, double zj, int it,
                                              int tj, const Grid &g, const PairTab &pt, double &e,
                                              double &fx, double &fy, double &fz)
{
  double delx = c[0] - xj, dely = c[1] - yj, delz = c[2] - zj;
  while (fabs(dely) > g.half[1]) {  // domain.cpp:980-1008, y and z periodic
    if (dely < 0.0) dely += g.prd[1];
    else dely -= g.prd[1];
  }
  while (fabs(delz) > g.half[2]) {
    if (delz < 0.0) delz += g.prd[2];
    else delz -= g.prd[2];
  }
  double rsq = delx * delx + dely * dely + delz * delz;
  if (!(rsq < pt.cutsq[it][tj])) return false;
  double r = sqrt(rsq);
  double fforce, phi;
  if (r < EPSILON_DPD) {
    fforce = 0.0;
    phi = 0.0;
  } else {
    double rinv = 1.0 / r;
    double wd = 1.0 - r / pt.cut[it][tj];
    fforce = pt.a0[it][tj] * wd * 1.0 * rinv;
    phi = 0.5 * pt.a0[it][tj] * pt.cut[it][tj] * wd * wd;
    phi = 1.0 * phi;
  }
  e = phi;
  fx = fforce * delx;
  fy = fforce * dely;
  fz = fforce * delz;
  return true;
}

// One warp evaluates E and f at point c: gathers contributions from the 27 cells and from the
// atoms accepted earlier in this pre_exchange (previous round's state), orders them by LAMMPS
// local index and sums them left to right -- the reference's summation order (:1790-1807).
__device__ void usher_energy(const double c[3], int t, const Planes &p, const Grid &g, const PairTab &pt,
                             const FixDev &fd, const int *__restrict__ cell_start, const UsherCfg &u,
                             const int *__restrict__ acc_list, const obmd_trial *__restrict__ prev,
                             WarpHits &H, StepScalars *sc, double &E, double f[3])
{
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int it = fd.ntype;
  int nh = 0;
  int cx = cell_cx(g, c[0]), cy = cell_cy(g, c[1]), cz = cell_cz(g, c[2]);
  for (int ox = -1; ox <= 1; ox++) {
    int bx = cx + ox;
    if (bx < 0 || bx >= g.nx) continue;
    for (int oy = -1; oy <= 1; oy++) {
      int by = cell_wrap(cy + oy, g.ny);
      for (int oz = -1; oz <= 1; oz++) {
        int bz = cell_wrap(cz + oz, g.nz);
        int cc = (bx * g.ny + by) * g.nz + bz;
        int jb = cell_start[cc], je = cell_start[cc + 1];
        for (int j0 = jb; j0 < je; j0 += 32) {
          int j = j0 + lane;
          bool hit = false;
          double e = 0, fx = 0, fy = 0, fz = 0;
          int key = 0;
          if (j < je) {
            hit = usher_contrib(c, p.x[j], p.y[j], p.z[j], it, p.type[j], g, pt, e, fx, fy, fz);
            key = p.lidx[j];
          }
          unsigned m = __ballot_sync(0xffffffffu, hit);
          if (hit) {
            int s = nh + __popc(m & lt);
            if (s < USH_CAP) {
              H.key[s] = key;
              H.e[s] = e;
              H.fx[s] = fx;
              H.fy[s] = fy;
              H.fz[s] = fz;
            }
          }
          nh += __popc(m);
        }
      }
    }
  }
  // atoms created by earlier trials of this call (they sit at local indices nbase + rank)
  for (int k0 = 0; k0 < u.nacc_prev; k0 += 32) {
    int k = k0 + lane;
    bool hit = false;
    double e = 0, fx = 0, fy = 0, fz = 0;
    if (k < u.nacc_prev) {
      int tp = acc_list[k];
      if (tp < t) {
        const obmd_trial &q = prev[tp];
        hit = usher_contrib(c, q.end[0], q.end[1], q.end[2], it, fd.ntype, g, pt, e, fx, fy, fz);
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      int s = nh + __popc(m & lt);
      if (s < USH_CAP) {
        H.key[s] = u.nbase + k;
        H.e[s] = e;
        H.fx[s] = fx;
        H.fy[s] = fy;
        H.fz[s] = fz;
      }
    }
    nh += __popc(m);
  }
  if (nh > USH_CAP) {
    if (lane == 0) sc->usher_error = OBMD_ERR_CAPACITY;
    nh = USH_CAP;
  }
  __syncwarp();
  // rank sort by key (keys are distinct local indices)
  for (int s = lane; s < nh; s += 32) {
    int k = H.key[s], r = 0;
    for (int q = 0; q < nh; q++) r += (H.key[q] < k);
    H.skey[r] = s;
  }
  __syncwarp();
  double Es = 0.0, f0 = 0.0, f1 = 0.0, f2 = 0.0;
  if (lane == 0) {
    for (int r = 0; r < nh; r++) {
      int s = H.skey[r];
      Es += H.e[s];
      f0 += H.fx[s];
      f1 += H.fy[s];
      f2 += H.fz[s];
    }
  }
  E = __shfl_sync(0xffffffffu, Es, 0);
  f[0] = __shfl_sync(0xffffffffu, f0, 0);
  f[1] = __shfl_sync(0xffffffffu, f1, 0);
  f[2] = __shfl_sync(0xffffffffu, f2, 0);
  __syncwarp();
}

// One warp per trial in `list`.  Writes the trial record and the points at which the energy was
// evaluated (needed by the dirty test of the next round).
__global__ void __launch_bounds__(USH_WARPS * 32)
    k_usher_eval(const int *__restrict__ list, int nlist, Planes p, Grid g, PairTab pt, FixDev fd,
                 const int *__restrict__ cell_start, UsherCfg u, const int *__restrict__ acc_list,
                 const obmd_trial *__restrict__ prev, obmd_trial *__restrict__ cur,
                 double *__restrict__ path, int *__restrict__ npts, StepScalars *sc)
{
  __shared__ WarpHits hits[USH_WARPS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int li = blockIdx.x * USH_WARPS + w;
  if (li >= nlist) return;
  const int t = list ? list[li] : li;
  const int side = (t < u.ntrL) ? 0 : 1;
  const obmd_block &reg = fd.region[4 + side];

  // trial start (:923-927): three consecutive RanPark draws, x then y then z
  int seed = (int) (((unsigned long long) u.seed0 * (unsigned long long) ranpark_pow(3LL * t)) %
                    2147483647ULL);
  double c[3], start[3];
  for (int d = 0; d < 3; d++) {
    c[d] = reg.lo[d] + ranpark_uniform(seed) * (reg.hi[d] - reg.lo[d]);
    start[d] = c[d];
  }

  double *mypath = path + (size_t) t * (u.natt + 1) * 3;
  int i = 0, np = 0;
  double E = 0.0, f[3];
  while (i < u.natt) {
    usher_energy(c, t, p, g, pt, fd, cell_start, u, acc_list, prev, hits[w], sc, E, f);
    if (lane == 0) {
      mypath[np * 3 + 0] = c[0];
      mypath[np * 3 + 1] = c[1];
      mypath[np * 3 + 2] = c[2];
    }
    np++;
    if (E < fd.etarget + EPSILON_FIX) break;
    double fabs_ = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    if (fabs_ < EPSILON_FIX) {  // the reference `continue`s without progress: report it
      if (lane == 0) sc->usher_error = OBMD_ERR_USHER;
      break;
    }
    double ds;
    if (E > fd.uovlp) ds = fd.dsovlp - pow(4 * fd.eps / E, 1.0 / 12.0);
    else ds = fmin((E - fd.etarget) / fabs_, fd.ds0);
    c[0] += f[0] / fabs_ * ds;
    c[1] += f[1] / fabs_ * ds;
    c[2] += f[2] / fabs_ * ds;
    if (!block_inside(reg, c[0], c[1], c[2])) break;  // check_mol_region (:1700-1712)
    i += 1;
  }
  bool ok = (E < fd.etarget + EPSILON_FIX);
  // check_mol_proc / check_proc (:1645-1696) on the owning rank: half-open box in x,y; the top z
  // face belongs to the last processor
  if (ok) {
    bool in = c[0] >= fd.boxlo[0] && c[0] < fd.boxhi[0] && c[1] >= fd.boxlo[1] && c[1] < fd.boxhi[1] &&
        c[2] >= fd.boxlo[2];
    if (!in) ok = false;
  }
  if (ok && !block_inside(reg, c[0], c[1], c[2])) ok = false;  // :1107-1108
  if (lane == 0) {
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
    cur[t] = r;
    npts[t] = np;
  }
}

// changed[t] = the trial's state (accepted, end) differs between the two latest rounds
__global__ void k_usher_diff(const obmd_trial *__restrict__ prev, const obmd_trial *__restrict__ cur,
                             int ntr, int first_round, int *__restrict__ changed, StepScalars *sc)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntr) return;
  int ch;
  if (first_round) ch = cur[t].accepted;
  else {
    const obmd_trial &a = prev[t], &b = cur[t];
    ch = (a.accepted != b.accepted) ||
        (b.accepted && (a.end[0] != b.end[0] || a.end[1] != b.end[1] || a.end[2] != b.end[2]));
  }
  changed[t] = ch;
  if (ch) atomicAdd(&sc->nchanged, 1);
}

__device__ __forceinline__ bool near_pt(const double *a, const double *b, const Grid &g, double rsqmax)
{
  double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  dy -= g.prd[1] * rint(dy / g.prd[1]);
  dz -= g.prd[2] * rint(dz / g.prd[2]);
  return dx * dx + dy * dy + dz * dz <= rsqmax;
}

// A trial must be re-run when an earlier trial whose state changed has its old or new atom
// within the cutoff of any point the trial evaluated (a superset test with a small margin).
__global__ void k_usher_dirty(const obmd_trial *__restrict__ prev, const obmd_trial *__restrict__ cur,
                              int ntr, int first_round, const int *__restrict__ changed,
                              const double *__restrict__ path, const int *__restrict__ npts, int natt,
                              Grid g, double rsqmax, int *__restrict__ list, StepScalars *sc)
{
  int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (t >= ntr) return;
  const double *mp = path + (size_t) t * (natt + 1) * 3;
  int np = npts[t];
  bool dirty = false;
  for (int q = lane; q < t && !dirty; q += 32) {
    if (!changed[q]) continue;
    bool oldacc = !first_round && prev[q].accepted, newacc = cur[q].accepted;
    for (int k = 0; k < np && !dirty; k++) {
      if (oldacc && near_pt(prev[q].end, mp + 3 * k, g, rsqmax)) dirty = true;
      if (newacc && near_pt(cur[q].end, mp + 3 * k, g, rsqmax)) dirty = true;
    }
  }
  dirty = __any_sync(0xffffffffu, dirty);
  if (dirty && lane == 0) list[atomicAdd(&sc->ndirty, 1)] = t;
}

// prev <- cur and the ordered list of accepted trial ids (one block, ordered compaction)
__global__ void k_usher_acc(obmd_trial *__restrict__ prev, const obmd_trial *__restrict__ cur, int ntr,
                            int *__restrict__ acc_list, StepScalars *sc)
{
  __shared__ int base;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < ntr; t0 += blockDim.x) {
    int t = t0 + threadIdx.x;
    int a = 0;
    if (t < ntr) {
      prev[t] = cur[t];
      a = cur[t].accepted;
    }
    unsigned m = __ballot_sync(0xffffffffu, a);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int off = base;
    for (int k = 0; k < w; k++) off += wsum[k];
    if (a) acc_list[off + __popc(m & ((1u << lane) - 1u))] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int k = 0; k < (int) (blockDim.x >> 5); k++) s += wsum[k];
      base += s;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) sc->naccept = base;
}

// create_atom + tag/mask/image/v assignment (:1121-1150, atom_vec.cpp:1592-1612), appended after
// the sorted range; also threads the new atoms into per-cell lists the pair kernel walks.
__global__ void k_usher_commit(Planes p, int nsorted, int maxtag, int nbase, Grid g, FixDev fd,
                               obmd_trial *__restrict__ cur, const int *__restrict__ acc_list, int nacc,
                               int *__restrict__ extra_head, int *__restrict__ extra_next,
                               int *__restrict__ extra_cell, double *fx, double *fy, double *fz)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nacc) {
    obmd_trial &r = cur[acc_list[k]];
    int s = nsorted + k;
    p.x[s] = r.end[0];
    p.y[s] = r.end[1];
    p.z[s] = r.end[2];
    p.vx[s] = p.vy[s] = p.vz[s] = 0.0;
    p.type[s] = fd.ntype;
    p.tag[s] = maxtag + k + 1;
    p.image[s] = IMG0;
    p.lidx[s] = nbase + k;
    fx[s] = fy[s] = fz[s] = 0.0;
    r.tag = maxtag + k + 1;
    extra_cell[k] = (cell_cx(g, r.end[0]) * g.ny + cell_cy(g, r.end[1])) * g.nz + cell_cz(g, r.end[2]);
  }
}

// single thread: push in descending k so each cell's list runs in ascending k (deterministic)
__global__ void k_extra_link(const int *__restrict__ extra_cell, int nacc, int *__restrict__ extra_head,
                             int *__restrict__ extra_next)
{
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int k = nacc - 1; k >= 0; k--) {
      int cc = extra_cell[k];
      extra_next[k] = extra_head[cc];
      extra_head[cc] = k;
    }
}

__global__ void k_extra_clear(const int *__restrict__ extra_cell, int nacc, int *__restrict__ extra_head)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nacc) extra_head[extra_cell[k]] = -1;
}

int fix_clear_extras(obmd_ctx *c)
{
  if (c->nextra > 0)
    LAUNCH(k_extra_clear, cdiv(c->nextra, 128), 128, 0, c->extra_cell, c->nextra, c->extra_head);
  return OBMD_OK;
}

static int ensure_trials(obmd_ctx *c, int ntr, int natt)
{
  if (ntr <= c->tr_cap && natt <= c->tr_natt) return OBMD_OK;
  int cap = ntr > c->tr_cap ? ntr + ntr / 2 + 256 : c->tr_cap;
  int na = natt > c->tr_natt ? natt : c->tr_natt;
  cudaFree(c->tr_d); cudaFree(c->tr_prev_d); cudaFree(c->tr_path); cudaFree(c->tr_npts);
  cudaFree(c->tr_list); cudaFree(c->tr_changed); cudaFree(c->acc_list);
  if (c->tr_h) cudaFreeHost(c->tr_h);
  CK(cudaMalloc(&c->tr_d, sizeof(obmd_trial) * cap));
  CK(cudaMalloc(&c->tr_prev_d, sizeof(obmd_trial) * cap));
  CK(cudaMalloc(&c->tr_path, sizeof(double) * 3 * (size_t) cap * (na + 1)));
  CK(cudaMalloc(&c->tr_npts, sizeof(int) * cap));
  CK(cudaMalloc(&c->tr_list, sizeof(int) * cap));
  CK(cudaMalloc(&c->tr_changed, sizeof(int) * cap));
  CK(cudaMalloc(&c->acc_list, sizeof(int) * cap));
  CK(cudaMallocHost(&c->tr_h, sizeof(obmd_trial) * cap));
  c->tr_cap = cap;
  c->tr_natt = na;
  return OBMD_OK;
}

int read_scalars(obmd_ctx *c)
{
  CK(cudaMemcpyAsync(c->sc_h, c->sc_d, sizeof(StepScalars), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return OBMD_OK;
}

int fix_pre_exchange(obmd_ctx *c, obmd_trial *trials, int maxtrials)
{
  NEED(c->have_fix, OBMD_ERR_STATE, "obmd_pre_exchange: call obmd_fix_create first");
  NEED(c->cells_valid, OBMD_ERR_STATE, "obmd_pre_exchange: cell list is stale (call obmd_build_cells)");
  const obmd_fix_args &fa = c->fa;
  obmd_pre_exchange_report &rep = c->rep;
  memset(&rep, 0, sizeof(rep));
  const int nlive = c->nsorted, ndel = c->nparked;
  int rc;
  {
    ScopedTimer t(c, "pre_exchange");
    // zero cnt_left, cnt_right, maxtag, ntail, nonfinite, usher_error, nchanged, ndirty, naccept
    CK(cudaMemsetAsync(c->sc_d, 0, offsetof(StepScalars, rng_seed), c->stream));
    CK(cudaMemsetAsync(c->sc_d->vnewl, 0, 6 * sizeof(double), c->stream));
    if (ndel > c->del_cap) {
      cudaFree(c->del_key); cudaFree(c->del_val); cudaFree(c->tail_key); cudaFree(c->tail_val);
      cudaFree(c->del_src);
      c->del_cap = next_pow2(ndel + 1024);
      CK(cudaMalloc(&c->del_src, sizeof(int) * c->del_cap));
      CK(cudaMalloc(&c->del_key, sizeof(int) * c->del_cap));
      CK(cudaMalloc(&c->del_val, sizeof(int) * c->del_cap));
      CK(cudaMalloc(&c->tail_key, sizeof(int) * c->del_cap));
      CK(cudaMalloc(&c->tail_val, sizeof(int) * c->del_cap));
    }
    LAUNCH(k_scan_atoms, cdiv(nlive > 0 ? nlive : 1, 256), 256, 0, c->cur, nlive, c->fd, c->sc_d,
           c->tail_key, c->tail_val, c->del_cap);
    if (ndel > 0) {
      int npad = next_pow2(ndel);
      LAUNCH(k_delete_fixup, 1, 1024, 0, c->cur, nlive, ndel, npad, npad, c->fd, c->pt, c->sc_d,
             c->del_key, c->del_val, c->tail_key, c->tail_val, c->del_src);
    }
    if ((rc = read_scalars(c))) return rc;
  }
  const StepScalars &s = *c->sc_h;
  NEED(s.usher_error == 0, OBMD_ERR_STATE, "fix obmd: deletion bookkeeping invariant violated");
  // try_deleting x2 (:578-579): one uniform() per deleted atom (Q7)
  c->fs.rng_seed = ranpark_skip(c->fs.rng_seed, ndel);
  c->fs.ndeleted += ndel;
  c->n = nlive;  // the parked atoms are gone
  c->nparked = 0;
  rep.ndel = ndel;
  rep.cnt_left = s.cnt_left;
  rep.cnt_right = s.cnt_right;
  rep.maxtag = s.maxtag;
  for (int d = 0; d < 3; d++) {
    rep.vnewl[d] = s.vnewl[d];
    rep.vnewr[d] = s.vnewr[d];
  }
  // :586-589 (mol_len = 1 in ATOM mode); truncation toward zero
  const double mol_len = 1.0;
  int stevL = -static_cast<int>((static_cast<double>(s.cnt_left) / mol_len - fa.alpha * fa.nbuf) *
                                c->dt / fa.tau);
  int stevR = -static_cast<int>((static_cast<double>(s.cnt_right) / mol_len - fa.alpha * fa.nbuf) *
                                c->dt / fa.tau);
  rep.ninsert_left = stevL;
  rep.ninsert_right = stevR;
  const int ntrL = stevL > 0 ? stevL : 0, ntrR = stevR > 0 ? stevR : 0;
  const int ntr = ntrL + ntrR;

  int nacc = 0;
  if (ntr > 0) {
    ScopedTimer t(c, "usher");
    NEED(fa.maxattempt == 1, OBMD_ERR_ARG,
         "fix obmd: maxattempt > 1 is not implemented on the device path (DESIGN.md 5.4)");
    if ((rc = ensure_trials(c, ntr, fa.nattempt))) return rc;
    if ((rc = obmd_ensure_capacity(c, nlive + ntr))) return rc;
    UsherCfg u;
    u.ntrL = ntrL;
    u.ntr = ntr;
    u.seed0 = c->fs.rng_seed;
    u.nbase = nlive;
    u.nacc_prev = 0;
    u.natt = fa.nattempt;
    const double rsqmax = c->pt.cutmax * c->pt.cutmax * (1.0 + 1e-9);
    int nlist = ntr;
    const int *list = nullptr;
    int round = 0;
    for (;; round++) {
      NEED(round <= ntr + 1, OBMD_ERR_STATE, "fix obmd: USHER commit did not converge");
      LAUNCH(k_usher_eval, cdiv(nlist, USH_WARPS), USH_WARPS * 32, 0, list, nlist, c->cur, c->grid, c->pt,
             c->fd, c->cell_start, u, c->acc_list, c->tr_prev_d, c->tr_d, c->tr_path, c->tr_npts, c->sc_d);
      CK(cudaMemsetAsync(&c->sc_d->nchanged, 0, 2 * sizeof(int), c->stream));  // nchanged, ndirty
      LAUNCH(k_usher_diff, cdiv(ntr, 256), 256, 0, c->tr_prev_d, c->tr_d, ntr, round == 0 ? 1 : 0,
             c->tr_changed, c->sc_d);
      LAUNCH(k_usher_dirty, cdiv(ntr, 4), 128, 0, c->tr_prev_d, c->tr_d, ntr, round == 0 ? 1 : 0,
             c->tr_changed, c->tr_path, c->tr_npts, u.natt, c->grid, rsqmax, c->tr_list, c->sc_d);
      LAUNCH(k_usher_acc, 1, 1024, 0, c->tr_prev_d, c->tr_d, ntr, c->acc_list, c->sc_d);
      if ((rc = read_scalars(c))) return rc;
      NEED(c->sc_h->usher_error != OBMD_ERR_USHER, OBMD_ERR_USHER,
           "fix obmd: USHER force vanished with E above target (the reference would not terminate)");
      NEED(c->sc_h->usher_error != OBMD_ERR_CAPACITY, OBMD_ERR_CAPACITY,
           "fix obmd: more than %d neighbours within the cutoff of a USHER point", USH_CAP);
      nacc = c->sc_h->naccept;
      if (c->sc_h->nchanged == 0 || c->sc_h->ndirty == 0) break;
      nlist = c->sc_h->ndirty;
      list = c->tr_list;
      u.nacc_prev = nacc;
    }
    rep.usher_rounds = round + 1;
    // every trial consumes exactly three draws (ATOM mode, uniform distribution, block region)
    c->fs.rng_seed = ranpark_skip(c->fs.rng_seed, 3LL * ntr);
    if (nacc > 0) {
      LAUNCH(k_usher_commit, cdiv(nacc, 128), 128, 0, c->cur, nlive, s.maxtag, nlive, c->grid, c->fd,
             c->tr_d, c->acc_list, nacc, c->extra_head, c->extra_next, c->extra_cell, c->fx, c->fy, c->fz);
      LAUNCH(k_extra_link, 1, 32, 0, c->extra_cell, nacc, c->extra_head, c->extra_next);
    }
    if (trials && maxtrials > 0) {
      int ncopy = ntr < maxtrials ? ntr : maxtrials;
      CK(cudaMemcpyAsync(c->tr_h, c->tr_d, sizeof(obmd_trial) * ncopy, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      memcpy(trials, c->tr_h, sizeof(obmd_trial) * ncopy);
    }
    CK(cudaGetLastError());
  }
  c->nextra = nacc;
  c->n = nlive + nacc;
  c->fs.ninserted += nacc;
  rep.ntrials = ntr;
  rep.naccept = nacc;
  rep.nlocal = c->n;

  // :608-630 flux forces
  const double ly = c->boxhi[1] - c->boxlo[1], lz = c->boxhi[2] - c->boxlo[2];
  const double area = ly * lz;
  const double dt = c->dt;
  obmd_fix_state &fs = c->fs;
  fs.simulation_time += dt;
  double factor = fa.pxx + fa.dpxx * sin(2.0 * MY_PI * fa.freq * fs.simulation_time);
  fs.momentumForce_left[0] = s.vnewl[0] / (dt) + factor * area;
  fs.momentumForce_left[1] = s.vnewl[1] / (dt);
  fs.momentumForce_left[2] = s.vnewl[2] / (dt);
  fs.shearForce_left[0] = 0.0;
  fs.shearForce_left[1] = fa.pxy * area;
  fs.shearForce_left[2] = fa.pxz * area;
  fs.momentumForce_right[0] = s.vnewr[0] / (dt) -fa.pxx * area;
  fs.momentumForce_right[1] = s.vnewr[1] / (dt);
  fs.momentumForce_right[2] = s.vnewr[2] / (dt);
  fs.shearForce_right[0] = 0.0;
  fs.shearForce_right[1] = -fa.pxy * area;
  fs.shearForce_right[2] = -fa.pxz * area;
  return OBMD_OK;
}

// =====================================================================================
// post_force: buffer-region force distribution
// =====================================================================================

struct ForceCfg {
  int nint;          // disjoint cell-column intervals to visit (x slowest => contiguous slots)
  int cxa[4], cxb[4];
  int active[4];     // region r has a non-zero total force
  double F[4][3];    // momentumForce_left/right, shearForce_left/right
};

// g_par_local_charged (:1312-1340); the fall-through (Q13) contributes 0
__device__ __forceinline__ double g_par_local(const FixDev &fd, double mass, double xv)
{
  double upper_x = fd.boxhi[0], lower_x = fd.boxlo[0];
  double buffer_size = fd.buffer_size, g_fac = fd.g_fac, carg;
  if (xv < lower_x + buffer_size) {
    if (xv < (lower_x + (1.0 - g_fac) * buffer_size)) return mass;
    carg = 1.0 / g_fac * MY_PI * (xv - buffer_size - lower_x) / (-buffer_size) - MY_PI;
    return 0.5 * (1.0 + cos(carg)) * mass;
  }
  if (xv > upper_x - buffer_size) {
    if (xv > (upper_x - (1.0 - g_fac) * buffer_size)) return mass;
    carg = 1.0 / g_fac * MY_PI * (xv - upper_x + buffer_size) / (buffer_size) -MY_PI;
    return 0.5 * (1.0 + cos(carg)) * mass;
  }
  return 0.0;
}

// g_par_global_charged (:1271-1299): both buffer tests are evaluated (no else)
__device__ __forceinline__ double g_par_global_term(const FixDev &fd, double mass, double xv)
{
  double upper_x = fd.boxhi[0], lower_x = fd.boxlo[0];
  double buffer_size = fd.buffer_size, g_fac = fd.g_fac, carg, g = 0.0;
  if (xv < lower_x + buffer_size) {
    if (xv < (lower_x + (1.0 - g_fac) * buffer_size)) g += mass;
    else {
      carg = 1.0 / g_fac * MY_PI * (xv - buffer_size - lower_x) / (-buffer_size) - MY_PI;
      g += 0.5 * (1.0 + cos(carg)) * mass;
    }
  }
  if (xv > upper_x - buffer_size) {
    if (xv > (upper_x - (1.0 - g_fac) * buffer_size)) g += mass;
    else {
      carg = 1.0 / g_fac * MY_PI * (xv - upper_x + buffer_size) / (buffer_size) -MY_PI;
      g += 0.5 * (1.0 + cos(carg)) * mass;
    }
  }
  return g;
}

#define FB_BLOCKS (NSM * 2)
#define FB_THREADS 256

// Pass 1 (g_par_global_charged :1244, g_perp_global_charged :1344): weight sums of the four
// regions.  Fixed grid, grid-stride over the contiguous slot ranges, warp-shuffle + block tree,
// one partial per block: the final value does not depend on scheduling.
__global__ void __launch_bounds__(FB_THREADS)
    k_weight_sums(Planes p, int nsorted, int n, Grid g, FixDev fd, PairTab pt, ForceCfg fc,
                  const int *__restrict__ cell_start, double *__restrict__ partials)
{
  double s[4] = {0, 0, 0, 0};
  const int nyz = g.ny * g.nz;
  for (int iv = 0; iv <= fc.nint; iv++) {
    int b, e;
    if (iv < fc.nint) {
      b = cell_start[fc.cxa[iv] * nyz];
      e = cell_start[(fc.cxb[iv] + 1) * nyz];
    } else {  // atoms inserted after the sort
      b = nsorted;
      e = n;
    }
    for (int i = b + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += gridDim.x * blockDim.x) {
      double x = p.x[i], y = p.y[i], z = p.z[i];
      double m = pt.mass[p.type[i]];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (!fc.active[r]) continue;
        if (!block_inside(fd.region[r], x, y, z)) continue;
        s[r] += (r < 2) ? g_par_global_term(fd, m, x) : m;
      }
    }
  }
  __shared__ double sh[4][FB_THREADS / 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    double v = s[r];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) sh[r][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int ww = 0; ww < FB_THREADS / 32; ww++) v += sh[threadIdx.x][ww];
    partials[blockIdx.x * 4 + threadIdx.x] = v;
  }
}

// Pass 2 (reg_force :1414-1428, reg_force_perp :1480-1494): f += F * w / G.  Each block first
// re-reduces the FB_BLOCKS partials in a fixed order (no extra launch, no atomics).
__global__ void __launch_bounds__(FB_THREADS)
    k_apply_forces(Planes p, int nsorted, int n, Grid g, FixDev fd, PairTab pt, ForceCfg fc,
                   const int *__restrict__ cell_start, const double *__restrict__ partials,
                   double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
                   StepScalars *sc)
{
  __shared__ double G[4];
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int b = 0; b < FB_BLOCKS; b++) v += partials[b * 4 + threadIdx.x];
    G[threadIdx.x] = v;
    if (blockIdx.x == 0) sc->gsum[threadIdx.x] = v;
  }
  __syncthreads();
  const int nyz = g.ny * g.nz;
  for (int iv = 0; iv <= fc.nint; iv++) {
    int b, e;
    if (iv < fc.nint) {
      b = cell_start[fc.cxa[iv] * nyz];
      e = cell_start[(fc.cxb[iv] + 1) * nyz];
    } else {
      b = nsorted;
      e = n;
    }
    for (int i = b + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += gridDim.x * blockDim.x) {
      double x = p.x[i], y = p.y[i], z = p.z[i];
      double m = pt.mass[p.type[i]];
      double ax = fx[i], ay = fy[i], az = fz[i];
      bool touched = false;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (!fc.active[r]) continue;
        if (!block_inside(fd.region[r], x, y, z)) continue;
        double gloc = (r < 2) ? g_par_local(fd, m, x) : m;
        ax += fc.F[r][0] * gloc / G[r];
        ay += fc.F[r][1] * gloc / G[r];
        az += fc.F[r][2] * gloc / G[r];
        touched = true;
      }
      if (touched) {
        fx[i] = ax;
        fy[i] = ay;
        fz[i] = az;
      }
    }
  }
}

int fix_post_force(obmd_ctx *c)
{
  NEED(c->have_fix, OBMD_ERR_STATE, "obmd_post_force: call obmd_fix_create first");
  NEED(c->cells_valid, OBMD_ERR_STATE, "obmd_post_force: cell list is stale");
  ScopedTimer t(c, "post_force");
  const obmd_fix_state &fs = c->fs;
  ForceCfg fc;
  memset(&fc, 0, sizeof(fc));
  const double *F[4] = {fs.momentumForce_left, fs.momentumForce_right, fs.shearForce_left,
                        fs.shearForce_right};
  const Grid &g = c->grid;
  int lo[4], hi[4], nr = 0;
  for (int r = 0; r < 4; r++) {
    for (int d = 0; d < 3; d++) fc.F[r][d] = F[r][d];
    fc.active[r] = (F[r][0] != 0.0 || F[r][1] != 0.0 || F[r][2] != 0.0);
    if (!fc.active[r]) continue;
    const obmd_block &b = c->fd.region[r];
    int a = (int) floor((b.lo[0] - g.lo[0]) * g.inv[0]) - 1;  // one column of slack for rounding
    int e = (int) floor((b.hi[0] - g.lo[0]) * g.inv[0]) + 1;
    a = a < 0 ? 0 : a;
    e = e > g.nx - 1 ? g.nx - 1 : e;
    if (e < a) {
      fc.active[r] = 0;
      continue;
    }
    lo[nr] = a;
    hi[nr] = e;
    nr++;
  }
  if (nr == 0) return OBMD_OK;
  // merge the column intervals into disjoint ones (insertion sort by lo, then sweep)
  for (int i = 1; i < nr; i++)
    for (int k = i; k > 0 && lo[k] < lo[k - 1]; k--) {
      int t1 = lo[k]; lo[k] = lo[k - 1]; lo[k - 1] = t1;
      t1 = hi[k]; hi[k] = hi[k - 1]; hi[k - 1] = t1;
    }
  fc.nint = 0;
  for (int i = 0; i < nr; i++) {
    if (fc.nint > 0 && lo[i] <= fc.cxb[fc.nint - 1] + 1) {
      if (hi[i] > fc.cxb[fc.nint - 1]) fc.cxb[fc.nint - 1] = hi[i];
    } else {
      fc.cxa[fc.nint] = lo[i];
      fc.cxb[fc.nint] = hi[i];
      fc.nint++;
    }
  }
  if (c->partial_cap < FB_BLOCKS * 4) {
    if (c->partials) cudaFree(c->partials);
    c->partial_cap = FB_BLOCKS * 4 + 1024;
    CK(cudaMalloc(&c->partials, sizeof(double) * c->partial_cap));
  }
  const int n = c->nsorted + c->nextra;
  LAUNCH(k_weight_sums, FB_BLOCKS, FB_THREADS, 0, c->cur, c->nsorted, n, c->grid, c->fd, c->pt, fc,
         c->cell_start, c->partials);
  LAUNCH(k_apply_forces, FB_BLOCKS, FB_THREADS, 0, c->cur, c->nsorted, n, c->grid, c->fd, c->pt, fc,
         c->cell_start, c->partials, c->fx, c->fy, c->fz, c->sc_d);
  CK(cudaGetLastError());
  return OBMD_OK;
}

// =====================================================================================
// observables
// =====================================================================================

__global__ void k_kinetic(Planes p, int n, PairTab pt, double *__restrict__ partials)
{
  double s[4] = {0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double m = pt.mass[p.type[i]];
    double vx = p.vx[i], vy = p.vy[i], vz = p.vz[i];
    s[0] += m * (vx * vx + vy * vy + vz * vz);
    s[1] += m * vx;
    s[2] += m * vy;
    s[3] += m * vz;
  }
  __shared__ double sh[4][FB_THREADS / 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = 0; r < 4; r++) {
    double v = s[r];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) sh[r][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int ww = 0; ww < FB_THREADS / 32; ww++) v += sh[threadIdx.x][ww];
    partials[blockIdx.x * 4 + threadIdx.x] = v;
  }
}

__global__ void k_kinetic_final(const double *__restrict__ partials, StepScalars *sc, int n)
{
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int b = 0; b < FB_BLOCKS; b++) v += partials[b * 4 + threadIdx.x];
    sc->kin[threadIdx.x] = v;
  }
  if (threadIdx.x == 4) sc->kin[4] = (double) n;
}

int fix_reduce_kinetic(obmd_ctx *c, double out[5])
{
  if (c->partial_cap < FB_BLOCKS * 4) {
    if (c->partials) cudaFree(c->partials);
    c->partial_cap = FB_BLOCKS * 4 + 1024;
    CK(cudaMalloc(&c->partials, sizeof(double) * c->partial_cap));
  }
  LAUNCH(k_kinetic, FB_BLOCKS, FB_THREADS, 0, c->cur, c->n, c->pt, c->partials);
  LAUNCH(k_kinetic_final, 1, 32, 0, c->partials, c->sc_d, c->n);
  int rc = read_scalars(c);
  if (rc) return rc;
  for (int k = 0; k < 5; k++) out[k] = c->sc_h->kin[k];
  return OBMD_OK;
}

__global__ void k_profile_x(Planes p, int n, PairTab pt, double lo, double inv, int nbins,
                            double *__restrict__ count, double *__restrict__ mvx)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b = (int) floor((p.x[i] - lo) * inv);
  b = min(max(b, 0), nbins - 1);
  atomicAdd(&count[b], 1.0);
  atomicAdd(&mvx[b], pt.mass[p.type[i]] * p.vx[i]);
}

int fix_profile_x(obmd_ctx *c, int nbins, double *count, double *mvx)
{
  NEED(nbins > 0 && nbins <= 65536, OBMD_ERR_ARG, "obmd_profile_x: bad nbins");
  double *d = nullptr;
  CK(cudaMalloc(&d, sizeof(double) * 2 * nbins));
  CK(cudaMemsetAsync(d, 0, sizeof(double) * 2 * nbins, c->stream));
  LAUNCH(k_profile_x, cdiv(c->n, 256), 256, 0, c->cur, c->n, c->pt, c->boxlo[0],
         nbins / (c->boxhi[0] - c->boxlo[0]), nbins, d, d + nbins);
  CK(cudaMemcpyAsync(count, d, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(mvx, d + nbins, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return OBMD_OK;
}
