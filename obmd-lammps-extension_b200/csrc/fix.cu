// FixObmdMerged on the device (fix_obmd_merged.cpp) plus the FixNVE / Domain::pbc kernels that
// keep the state device-resident.  Compiled with -fmad=false: the bit-exact parts (USHER sums,
// trial coordinates, weights, integrator) must not be contracted into FMAs, because the
// reference is an x86-64 build without FMA.
//
//   pre_exchange (:550-633)      -> k_scan_ids, k_count_regions, k_delete_fixup; batched USHER in usher.cu
//   post_force   (:1229-1240)    -> k_weight_sums (last block adds the partials), k_apply_forces
//   FixNVE       (fix_nve.cpp)   -> k_nve_initial, k_nve_final
//   Domain::pbc  (domain.cpp)    -> k_pbc
#include <chrono>

#include "ctx.h"

#define EPSILON_FIX 1.0e-6   // fix_obmd_merged.cpp:62
#define EPSILON_DPD 1.0e-10  // pair_dpd.cpp:36
#define MY_PI 3.14159265358979323846
#define FB_BLOCKS (NSM * 6)  // fixed grid of the grid-stride region kernels (6 CTAs per SM: ~1 buffer atom per thread at 1 M atoms)
#define FB_THREADS 256

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

__global__ void k_pbc(Planes p, ColRange cr, Extras ex, Grid g, const int *__restrict__ cell_start)
{
  int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nyz = g.ny * g.nz;
  int i = -1;
  for (int k = 0; k < cr.n; k++) {
    int b = cell_start[cr.a[k] * nyz], e = cell_start[(cr.b[k] + 1) * nyz];
    if (gidx < e - b) {
      i = b + gidx;
      break;
    }
    gidx -= e - b;
  }
  if (i < 0 && cr.extras && gidx >= 0) i = extras_slot(ex, cr.extras, gidx);
  if (i < 0) return;
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
  c->cells_valid = c->rec_valid = false;
  return OBMD_OK;
}

int fix_nve_final(obmd_ctx *c)
{
  NEED(c->n > 0 && c->dt > 0.0, OBMD_ERR_STATE, "obmd_nve_final: atoms and timestep must be set");
  ScopedTimer t(c, "nve");
  LAUNCH(k_nve_final, cdiv(c->n, 256), 256, 0, c->cur, c->fx, c->fy, c->fz, c->n, 0.5 * c->dt, c->pt);
  c->rec_valid = false;  // velocities changed
  return OBMD_OK;
}

int fix_pbc_cols(obmd_ctx *c, const ColRange &cr, cudaStream_t st)
{
  NEED(c->cells_valid, OBMD_ERR_STATE, "obmd_pbc: cell list is stale (call obmd_build_cells)");
  ScopedTimer t(c, "pbc", st);
  const int n = c->nsorted + c->nextra;
  if (n > 0) LAUNCH_ON(st, k_pbc, cdiv(n, 256), 256, 0, c->cur, cr, extras_of(c), c->grid, c->cell_start);
  return OBMD_OK;
}

int fix_pbc(obmd_ctx *c)
{
  ColRange cr = all_columns(c);
  cr.extras = 2;  // ghosts are wrapped too: they are this rank's private copies
  return fix_pbc_cols(c, cr, c->stream);
}

// =====================================================================================
// pre_exchange, part 1: scan + deletion bookkeeping
// =====================================================================================

// One pass over the live atoms: Group::count for both buffers (group.cpp:793-811, closed block on
// the not-yet-wrapped coordinates, Q9), find_maxid (:1860-1875), and the list of survivors whose
// local index is >= nprime (they will be moved into the holes the deletions leave).
// The id pass touches only the two int planes (8 B/atom); positions are read only in the cell
// columns that overlap the two buffer regions, so the interior can be wrapped and its forces
// computed on the side stream while this runs (obmd_run).
__global__ void k_scan_ids(Planes p, int nlive, int want_tail, StepScalars *sc, int *__restrict__ tail_key,
                           int *__restrict__ tail_val, int tail_cap)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int mt = 0;
  if (i < nlive) {
    mt = p.tag[i];
    int li = want_tail ? p.lidx[i] : -1;
    if (li >= nlive) {
      int k = atomicAdd(&sc->ntail, 1);
      if (k < tail_cap) {
        tail_key[k] = li;
        tail_val[k] = i;
      }
    }
  }
  for (int d = 16; d > 0; d >>= 1) mt = max(mt, __shfl_down_sync(0xffffffffu, mt, d));
  __shared__ int wmax[8];  // one atomic per CTA instead of one per warp (all to the same address)
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mt;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int) (blockDim.x >> 5); k++) mt = max(mt, wmax[k]);
    if (mt) atomicMax(&sc->maxtag, mt);
  }
}

__global__ void k_count_regions(Planes p, ColRange cr, Extras ex, Grid g, FixDev fd,
                                const int *__restrict__ cell_start, StepScalars *sc)
{
  const int nyz = g.ny * g.nz;
  int cl = 0, cright = 0;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ex.narr + ex.nins; q += gridDim.x * blockDim.x) {
    int i = extras_slot(ex, 1, q);  // arrivals from neighbour slabs count as owned
    cl += block_inside(fd.region[0], p.x[i], p.y[i], p.z[i]);
    cright += block_inside(fd.region[1], p.x[i], p.y[i], p.z[i]);
  }
  for (int k = 0; k < cr.n; k++) {
    int b = cell_start[cr.a[k] * nyz], e = cell_start[(cr.b[k] + 1) * nyz];
    for (int i = b + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += gridDim.x * blockDim.x) {
      double x = p.x[i], y = p.y[i], z = p.z[i];
      cl += block_inside(fd.region[0], x, y, z);
      cright += block_inside(fd.region[1], x, y, z);
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    cl += __shfl_down_sync(0xffffffffu, cl, d);
    cright += __shfl_down_sync(0xffffffffu, cright, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (cl) atomicAdd(&sc->cnt_left, cl);
    if (cright) atomicAdd(&sc->cnt_right, cright);
  }
}

// In-block bitonic sort (descending by key) of two independent (key, value) arrays of npad = power
// of two entries each, stage by stage together (one barrier per stage for both).
__device__ void bitonic_desc2(int *key0, int *val0, int *key1, int *val1, int npad)
{
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i2 = threadIdx.x; i2 < 2 * npad; i2 += blockDim.x) {
        int *key = i2 < npad ? key0 : key1, *val = i2 < npad ? val0 : val1;
        const int i = i2 < npad ? i2 : i2 - npad;
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
#define DEL_SMEM 1024  // deletions handled entirely in shared memory (more fall back to global)

__global__ void __launch_bounds__(1024)
    k_delete_fixup(Planes p, int nlive, int ndel, int npad_del, int npad_tail, FixDev fd, PairTab pt,
                   StepScalars *sc, int *del_key, int *del_val, int *tail_key, int *tail_val,
                   int *del_src, double *del_mv, int do_walk)
{
  __shared__ int s_dkey[DEL_SMEM], s_tkey[DEL_SMEM], s_tval[DEL_SMEM], s_src[DEL_SMEM];
  __shared__ double s_mv[4 * DEL_SMEM];
  const bool small = (npad_del <= DEL_SMEM && npad_tail <= DEL_SMEM);
  if (small) del_mv = s_mv;  // the six serial sums below then run out of shared memory
  const int ntail = do_walk ? sc->ntail : 0;
  // working arrays: shared memory when they fit (the usual few hundred deletions per step)
  int *dkey = small ? s_dkey : del_key, *dval = small ? s_src : del_val;
  int *tkey = small ? s_tkey : tail_key, *tval = small ? s_tval : tail_val;
  for (int k = threadIdx.x; k < npad_del; k += blockDim.x) {
    dkey[k] = (k < ndel) ? p.lidx[nlive + k] : -1;
    dval[k] = (k < ndel) ? nlive + k : -1;
  }
  for (int k = threadIdx.x; k < npad_tail; k += blockDim.x) {
    tkey[k] = (k < ntail) ? tail_key[k] : -1;
    tval[k] = (k < ntail) ? tail_val[k] : -1;
  }
  __syncthreads();
  if (small && npad_del <= 256 && npad_tail <= 256) {
    // few entries (the usual case): rank sort, two barriers instead of one per bitonic stage.  Rank =
    // number of larger keys (+ earlier equal ones: only the -1 padding repeats), descending order.
    const int t = threadIdx.x;
    int dk = -1, dv = -1, tk = -1, tv = -1, dr = 0, tr = 0;
    if (t < npad_del) {
      dk = dkey[t]; dv = dval[t];
      for (int j = 0; j < npad_del; j++) dr += (dkey[j] > dk) || (dkey[j] == dk && j < t);
    }
    if (t < npad_tail) {
      tk = tkey[t]; tv = tval[t];
      for (int j = 0; j < npad_tail; j++) tr += (tkey[j] > tk) || (tkey[j] == tk && j < t);
    }
    __syncthreads();
    if (t < npad_del) { dkey[dr] = dk; dval[dr] = dv; }
    if (t < npad_tail) { tkey[tr] = tk; tval[tr] = tv; }
    __syncthreads();
  } else
    bitonic_desc2(dkey, dval, tkey, tval, npad_del);  // npad_del == npad_tail (fix_pre_begin)
  // momentum contributions m*v of the deleted atoms, in descending local index order
  const double mid = 0.5 * (fd.boxhi[0] + fd.boxlo[0]);
  for (int k = threadIdx.x; k < ndel; k += blockDim.x) {
    int s = dval[k];
    double m = pt.mass[p.type[s]];
    del_mv[4 * k + 0] = m * p.vx[s];
    del_mv[4 * k + 1] = m * p.vy[s];
    del_mv[4 * k + 2] = m * p.vz[s];
    del_mv[4 * k + 3] = (p.x[s] < mid) ? 0.0 : 1.0;
  }
  __syncthreads();
  if (threadIdx.x < 6) {  // six serial sums (left/right x three components), reference order (:827-837)
    const int comp = threadIdx.x % 3;
    const double side = (threadIdx.x < 3) ? 0.0 : 1.0;
    double acc = 0.0;
    for (int k = 0; k < ndel; k++)
      if (del_mv[4 * k + 3] == side) acc += del_mv[4 * k + comp];
    if (threadIdx.x < 3) sc->vnewl[comp] = acc;
    else sc->vnewr[comp] = acc;
  }
  if (threadIdx.x == 32 && do_walk) {  // hole walk, on a different warp than the sums
    const int n = nlive + ndel;  // nlocal before the deletion; nprime = nlive
    int *src = small ? s_src : del_src;  // dval is dead once del_mv is written (small: same array)
    int z = 0;
    while (z < ndel && dkey[z] == n - 1 - z) z++;
    const int heff = ndel - z;
    int is = 0, id = z, bad = 0;
    for (int k = 0; k < heff; k++) {
      int pos = n - 1 - z - k;
      if (is < ntail && tkey[is] == pos) src[k] = tval[is++];
      else if (id < ndel && dkey[id] == pos) {
        src[k] = src[id - z];
        id++;
      } else
        bad = 1;
    }
    if (is != ntail) bad = 1;
    for (int k = 0; k < heff && !bad; k++) {
      int h = dkey[z + k];
      if (h < nlive) p.lidx[src[k]] = h;
    }
    if (bad) sc->usher_error = OBMD_ERR_STATE;  // bookkeeping invariant violated
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

// pre_exchange, part 2 (batched USHER: try_inserting :869-1225, usher :1518-1616) lives in usher.cu.

#define LINK_MAX 6000  // atoms per unsorted extra group (two int arrays in 48 KB of shared memory)

// cells of the extra atoms [first, first+count) (indices relative to the extras base)
__global__ void k_extra_cells(Planes p, int base, int first, int count, Grid g, int *__restrict__ extra_cell)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  int s = base + first + k;
  int lc = cell_cx(g, p.x[s]);
  lc = min(max(lc, 0), g.nx - 1);
  extra_cell[first + k] = (lc * g.ny + cell_cy(g, p.y[s])) * g.nz + cell_cz(g, p.z[s]);
}

// A group that arrives sorted by cell (a neighbour's boundary column): runs of equal cells become
// the lists.  Precondition: no earlier group has atoms in these cells.
__global__ void k_link_sorted(const int *__restrict__ extra_cell, int first, int count,
                              int *__restrict__ extra_head, int *__restrict__ extra_next)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  int e = first + k, cc = extra_cell[e];
  if (k == 0 || extra_cell[e - 1] != cc) extra_head[cc] = e;
  extra_next[e] = (k + 1 < count && extra_cell[e + 1] == cc) ? e + 1 : -1;
}

// An unsorted group (arrivals, a neighbour's own arrivals, USHER insertions): atom k links to the
// next k' > k of the group in the same cell, the last one to whatever the cell's list held before,
// and the first becomes the head.  One CTA, cells staged in shared memory, O(count) per thread,
// deterministic.
__global__ void k_link_group(const int *__restrict__ extra_cell, int first, int count,
                             int *__restrict__ extra_head, int *__restrict__ extra_next)
{
  extern __shared__ int cellsh[];  // [count] cells, [count] old heads
  int *oldh = cellsh + count;
  for (int k = threadIdx.x; k < count; k += blockDim.x) {
    cellsh[k] = extra_cell[first + k];
    oldh[k] = extra_head[cellsh[k]];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < count; k += blockDim.x) {
    int cc = cellsh[k], nxt = -1;
    for (int q = k + 1; q < count; q++)
      if (cellsh[q] == cc) {
        nxt = first + q;
        break;
      }
    extra_next[first + k] = (nxt >= 0) ? nxt : oldh[k];
    bool is_first = true;
    for (int q = 0; q < k; q++)
      if (cellsh[q] == cc) {
        is_first = false;
        break;
      }
    if (is_first) extra_head[cc] = first + k;
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

int read_scalars(obmd_ctx *c)
{
  const auto t0 = std::chrono::steady_clock::now();
  CK(cudaMemcpyAsync(c->sc_h, c->sc_d, sizeof(StepScalars), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (c->timing) {  // host wall time spent waiting here (includes the GPU work queued before it)
    TimerSlot &sl = c->timers["host_wait_scalars"];
    sl.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    sl.launches++;
  }
  return OBMD_OK;
}

// local cell columns overlapped by region block r (one column of slack each side for rounding),
// clipped to the columns this rank owns; e < a when the region misses the slab
static void region_cols(const obmd_ctx *c, int r, int margin, int &a, int &e)
{
  const Grid &g = c->grid;
  const obmd_block &b = c->fd.region[r];
  a = (int) floor((b.lo[0] - g.lo[0]) * g.inv[0]) - 1 - margin - g.cxoff;
  e = (int) floor((b.hi[0] - g.lo[0]) * g.inv[0]) + 1 + margin - g.cxoff;
  a = a < g.own_a ? g.own_a : a;
  e = e > g.own_b ? g.own_b : e;
}

static int merge_intervals(int *lo, int *hi, int n)
{
  for (int i = 1; i < n; i++)
    for (int k = i; k > 0 && lo[k] < lo[k - 1]; k--) {
      int t = lo[k]; lo[k] = lo[k - 1]; lo[k - 1] = t;
      t = hi[k]; hi[k] = hi[k - 1]; hi[k - 1] = t;
    }
  int m = 0;
  for (int i = 0; i < n; i++) {
    if (hi[i] < lo[i]) continue;
    if (m > 0 && lo[i] <= hi[m - 1] + 1) {
      if (hi[i] > hi[m - 1]) hi[m - 1] = hi[i];
    } else {
      lo[m] = lo[i];
      hi[m] = hi[i];
      m++;
    }
  }
  return m;
}

// columns of regions r0 and r1 as a ColRange (at most two disjoint intervals)
ColRange fix_region_columns(const obmd_ctx *c, int r0, int r1, int margin)
{
  int lo[2], hi[2];
  region_cols(c, r0, margin, lo[0], hi[0]);
  region_cols(c, r1, margin, lo[1], hi[1]);
  int m = merge_intervals(lo, hi, 2);
  ColRange cr;
  memset(&cr, 0, sizeof(cr));
  cr.n = m;
  for (int k = 0; k < m; k++) {
    cr.a[k] = lo[k];
    cr.b[k] = hi[k];
  }
  return cr;
}

// cells + per-cell lists for the extra atoms [first, first+count) (relative to the extras base)
int fix_link_extras(obmd_ctx *c, int first, int count, bool sorted_by_cell)
{
  if (count <= 0) return OBMD_OK;
  LAUNCH(k_extra_cells, cdiv(count, 128), 128, 0, c->cur, c->nsorted, first, count, c->grid, c->extra_cell);
  if (sorted_by_cell)
    LAUNCH(k_link_sorted, cdiv(count, 128), 128, 0, c->extra_cell, first, count, c->extra_head, c->extra_next);
  else {
    NEED(count <= LINK_MAX, OBMD_ERR_CAPACITY, "more than %d atoms in one unsorted extra group", LINK_MAX);
    LAUNCH(k_link_group, 1, 512, 2 * sizeof(int) * count, c->extra_cell, first, count, c->extra_head,
           c->extra_next);
  }
  return OBMD_OK;
}

// ---- pair forces near freshly inserted atoms -------------------------------------------------
// obmd_run computes the pair forces on the side stream while USHER is still deciding the
// insertions.  Afterwards only the atoms within one cell of a new atom (and the new atoms
// themselves) are recomputed, now with the new atoms in the lists: every atom's force is always
// the complete, fixed-order sum over its stencil, so the result equals a single late evaluation.
__global__ void k_mark_redo(Extras ex, Grid g, const int *__restrict__ cell_start,
                            const int *__restrict__ extra_cell, const int *__restrict__ extra_head,
                            const int *__restrict__ extra_next, int *__restrict__ flag,
                            int *__restrict__ list, int *__restrict__ count, int cap)
{
  int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  int k = gidx / 27, nb = gidx % 27;
  if (k >= ex.nins) return;
  int cc = extra_cell[ex.narr + ex.nghost + k];
  int cz = cc % g.nz, cy = (cc / g.nz) % g.ny, cx = cc / (g.nz * g.ny);
  int bx = cx + nb / 9 - 1, by = cy + (nb / 3) % 3 - 1, bz = cz + nb % 3 - 1;
  if (bx < g.own_a || bx > g.own_b) return;  // ghost columns hold no atoms of this rank
  by = by < 0 ? by + g.ny : (by >= g.ny ? by - g.ny : by);
  bz = bz < 0 ? bz + g.nz : (bz >= g.nz ? bz - g.nz : bz);
  int c2 = (bx * g.ny + by) * g.nz + bz;
  for (int j = cell_start[c2]; j < cell_start[c2 + 1]; j++)
    if (atomicExch(&flag[j], 1) == 0) {
      int q = atomicAdd(count, 1);
      if (q < cap) list[q] = j;
    }
  for (int e = extra_head[c2]; e >= 0; e = extra_next[e]) {
    if (e >= ex.narr && e < ex.narr + ex.nghost) continue;  // ghosts belong to the neighbour rank
    int j = ex.base + e;
    if (atomicExch(&flag[j], 1) == 0) {
      int q = atomicAdd(count, 1);
      if (q < cap) list[q] = j;
    }
  }
}

__global__ void k_clear_redo(int *__restrict__ flag, const int *__restrict__ list, const int *__restrict__ count,
                             int cap, StepScalars *sc)
{
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  int n = min(*count, cap);
  if (q < n) flag[list[q]] = 0;
  // more atoms around the insertions than the list holds (24 per cell): the surplus atoms would keep their
  // flag and miss the new atoms' force terms; reported at the next read of the step scalars
  if (q == 0 && *count > cap) sc->redo_overflow = 1;
}

int fix_redo_pair_near_insertions(obmd_ctx *c, long long step)
{
  if (c->nins == 0) return OBMD_OK;
  const int cap = c->nins * 27 * 24;  // 24 atoms per cell on average would be 8x the fluid density
  if (c->redo_cap < c->cap) {
    cudaFree(c->redo_flag);
    CK(cudaMalloc(&c->redo_flag, sizeof(int) * c->cap));
    CK(cudaMemsetAsync(c->redo_flag, 0, sizeof(int) * c->cap, c->stream));
    c->redo_cap = c->cap;
  }
  if (!c->redo_count) CK(cudaMalloc(&c->redo_count, sizeof(int)));
  if (cap > c->redo_list_cap) {
    cudaFree(c->redo_list);
    c->redo_list_cap = 2 * cap < (1 << 20) ? (1 << 20) : 2 * cap;  // regrowing mid-run stalls the step
    CK(cudaMalloc(&c->redo_list, sizeof(int) * c->redo_list_cap));
  }
  CK(cudaMemsetAsync(c->redo_count, 0, sizeof(int), c->stream));
  LAUNCH(k_mark_redo, cdiv(c->nins * 27, 128), 128, 0, extras_of(c), c->grid, c->cell_start, c->extra_cell,
         c->extra_head, c->extra_next, c->redo_flag, c->redo_list, c->redo_count, cap);
  int rc = pair_add_insertions(c, step, c->redo_list, c->redo_count, cap);
  if (rc) return rc;
  LAUNCH(k_clear_redo, cdiv(cap, 256), 256, 0, c->redo_flag, c->redo_list, c->redo_count, cap, c->sc_d);
  return OBMD_OK;
}

// copy the 13 per-atom planes of `count` atoms from slot src to slot dst (dst < src) via scratch
__global__ void k_move_atoms(Planes p, Planes tmp, double *fx, double *fy, double *fz, double *ftmp, int src,
                             int dst, int count, int phase)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  if (phase == 0) {
    int s = src + k;
    tmp.x[k] = p.x[s]; tmp.y[k] = p.y[s]; tmp.z[k] = p.z[s];
    tmp.vx[k] = p.vx[s]; tmp.vy[k] = p.vy[s]; tmp.vz[k] = p.vz[s];
    tmp.tag[k] = p.tag[s]; tmp.type[k] = p.type[s]; tmp.image[k] = p.image[s];
    ftmp[3 * k] = fx[s]; ftmp[3 * k + 1] = fy[s]; ftmp[3 * k + 2] = fz[s];
  } else {
    int d = dst + k;
    p.x[d] = tmp.x[k]; p.y[d] = tmp.y[k]; p.z[d] = tmp.z[k];
    p.vx[d] = tmp.vx[k]; p.vy[d] = tmp.vy[k]; p.vz[d] = tmp.vz[k];
    p.tag[d] = tmp.tag[k]; p.type[d] = tmp.type[k]; p.image[d] = tmp.image[k];
    p.lidx[d] = d;
    fx[d] = ftmp[3 * k]; fy[d] = ftmp[3 * k + 1]; fz[d] = ftmp[3 * k + 2];
  }
}

// Slab runs, after the forces are complete: the ghosts are dead; close the gap they leave between
// the arrivals and the inserted atoms so that [0, n) is again every atom this rank owns.
int fix_compact_extras(obmd_ctx *c)
{
  if (c->nghost == 0) return OBMD_OK;
  fix_clear_extras(c);
  const int src = c->nsorted + c->narr + c->nghost, dst = c->nsorted + c->narr;
  if (c->nins > 0) {
    LAUNCH(k_move_atoms, cdiv(c->nins, 128), 128, 0, c->cur, c->alt, c->fx, c->fy, c->fz, c->stage_d, src, dst,
           c->nins, 0);
    LAUNCH(k_move_atoms, cdiv(c->nins, 128), 128, 0, c->cur, c->alt, c->fx, c->fy, c->fz, c->stage_d, src, dst,
           c->nins, 1);
  }
  c->n = c->nsorted + c->narr + c->nins;
  c->nsorted = c->n;  // no longer cell-sorted as a whole, but contiguous; the next sort rebuilds
  c->narr = c->nghost = c->nins = c->nextra = 0;
  c->cells_valid = c->rec_valid = false;
  return OBMD_OK;
}

__global__ void k_pack_red(const StepScalars *sc, double *red, int ndel)
{
  if (threadIdx.x == 0) {
    red[0] = sc->cnt_left;
    red[1] = sc->cnt_right;
    red[2] = ndel;
    for (int d = 0; d < 3; d++) {
      red[3 + d] = sc->vnewl[d];
      red[6 + d] = sc->vnewr[d];
    }
    red[9] = sc->maxtag;  // reduced with max
  }
}

// First half of pre_exchange: everything that only needs the atoms parked outside the box
// (try_deleting :639-865).  Slab runs call it before the halo exchange because the arrivals are
// stored where the parked atoms sit.
int fix_pre_begin(obmd_ctx *c)
{
  NEED(c->have_fix, OBMD_ERR_STATE, "obmd_pre_exchange: call obmd_fix_create first");
  NEED(c->cells_valid, OBMD_ERR_STATE, "obmd_pre_exchange: cell list is stale (call obmd_build_cells)");
  ScopedTimer t(c, "pre_exchange");
  const int nlive = c->nsorted, ndel = c->nparked;
  const bool serial = (c->nranks == 1);
  // zero cnt_left, cnt_right, maxtag, ntail, nonfinite, usher_error, nchanged, ndirty, naccept
  CK(cudaMemsetAsync(c->sc_d, 0, offsetof(StepScalars, rng_seed), c->stream));
  CK(cudaMemsetAsync(c->sc_d->vnewl, 0, 6 * sizeof(double), c->stream));
  if (ndel > c->del_cap) {
    cudaFree(c->del_key); cudaFree(c->del_val); cudaFree(c->tail_key); cudaFree(c->tail_val);
    cudaFree(c->del_src); cudaFree(c->del_mv);
    c->del_cap = next_pow2(ndel + 1024);
    CK(cudaMalloc(&c->del_src, sizeof(int) * c->del_cap));
    CK(cudaMalloc(&c->del_mv, sizeof(double) * 4 * c->del_cap));
    CK(cudaMalloc(&c->del_key, sizeof(int) * c->del_cap));
    CK(cudaMalloc(&c->del_val, sizeof(int) * c->del_cap));
    CK(cudaMalloc(&c->tail_key, sizeof(int) * c->del_cap));
    CK(cudaMalloc(&c->tail_val, sizeof(int) * c->del_cap));
  }
  if (serial)  // tags + survivors that will move into holes; slab runs scan after the arrivals are in
    LAUNCH(k_scan_ids, cdiv(nlive > 0 ? nlive : 1, 256), 256, 0, c->cur, nlive, 1, c->sc_d, c->tail_key,
           c->tail_val, c->del_cap);
  if (ndel > 0) {
    int npad = next_pow2(ndel);
    LAUNCH(k_delete_fixup, 1, 1024, 0, c->cur, nlive, ndel, npad, npad, c->fd, c->pt, c->sc_d, c->del_key,
           c->del_val, c->tail_key, c->tail_val, c->del_src, c->del_mv, serial ? 1 : 0);
  }
  c->ndel_local = ndel;
  c->n = nlive;  // the parked atoms are gone
  c->nparked = 0;
  c->pre_begun = true;
  return OBMD_OK;
}

int fix_pre_exchange(obmd_ctx *c, obmd_trial *trials, int maxtrials)
{
  NEED(c->have_fix, OBMD_ERR_STATE, "obmd_pre_exchange: call obmd_fix_create first");
  NEED(c->cells_valid, OBMD_ERR_STATE, "obmd_pre_exchange: cell list is stale (call obmd_build_cells)");
  const obmd_fix_args &fa = c->fa;
  obmd_pre_exchange_report &rep = c->rep;
  memset(&rep, 0, sizeof(rep));
  const bool serial = (c->nranks == 1);
  int rc;
  if (!c->pre_begun && (rc = fix_pre_begin(c))) return rc;
  c->pre_begun = false;
  const int nlive = c->nsorted;
  int ndel = c->ndel_local, cntL, cntR, maxtag;
  double vnewl[3], vnewr[3];
  {
    ScopedTimer t(c, "pre_exchange");
    if (!serial)
      LAUNCH(k_scan_ids, cdiv(nlive + c->narr > 0 ? nlive + c->narr : 1, 256), 256, 0, c->cur, nlive + c->narr, 0,
             c->sc_d, c->tail_key, c->tail_val, c->del_cap);
    ColRange bufcols = fix_region_columns(c, 0, 1, 0);
    LAUNCH(k_count_regions, FB_BLOCKS, FB_THREADS, 0, c->cur, bufcols, extras_of(c), c->grid, c->fd,
           c->cell_start, c->sc_d);
    if (serial) {
      if ((rc = read_scalars(c))) return rc;
      const StepScalars &s = *c->sc_h;
      NEED(s.usher_error == 0, OBMD_ERR_STATE, "fix obmd: deletion bookkeeping invariant violated");
      NEED(s.redo_overflow == 0, OBMD_ERR_CAPACITY,
           "fix obmd: more than 24 atoms per cell around the insertions of the previous step (pair update list)");
      cntL = s.cnt_left; cntR = s.cnt_right; maxtag = s.maxtag;
      for (int d = 0; d < 3; d++) { vnewl[d] = s.vnewl[d]; vnewr[d] = s.vnewr[d]; }
    } else {
      // the ~14 MPI_Allreduce/Scan calls of the reference (:604-605,:677-678, group.cpp:809,
      // :1868) become one packed sum + one max over NVLink
      LAUNCH(k_pack_red, 1, 32, 0, c->sc_d, c->red_d, ndel);
      if ((rc = comm_allreduce(c, c->red_d, 9, 1))) return rc;
      CK(cudaMemcpyAsync(c->red_h, c->red_d, 10 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      cntL = (int) c->red_h[0]; cntR = (int) c->red_h[1]; ndel = (int) c->red_h[2];
      for (int d = 0; d < 3; d++) { vnewl[d] = c->red_h[3 + d]; vnewr[d] = c->red_h[6 + d]; }
      maxtag = (int) c->red_h[9];
    }
  }
  // try_deleting x2 (:578-579): one uniform() per deleted atom (Q7)
  c->fs.rng_seed = ranpark_skip(c->fs.rng_seed, ndel);
  c->fs.ndeleted += ndel;
  rep.ndel = ndel;
  rep.cnt_left = cntL;
  rep.cnt_right = cntR;
  rep.maxtag = maxtag;
  for (int d = 0; d < 3; d++) {
    rep.vnewl[d] = vnewl[d];
    rep.vnewr[d] = vnewr[d];
  }
  // :586-589 (mol_len = 1 in ATOM mode); truncation toward zero
  const double mol_len = 1.0;
  int stevL = -static_cast<int>((static_cast<double>(cntL) / mol_len - fa.alpha * fa.nbuf) * c->dt / fa.tau);
  int stevR = -static_cast<int>((static_cast<double>(cntR) / mol_len - fa.alpha * fa.nbuf) * c->dt / fa.tau);
  rep.ninsert_left = stevL;
  rep.ninsert_right = stevR;
  const int ntrL = stevL > 0 ? stevL : 0, ntrR = stevR > 0 ? stevR : 0;
  const int ntr = ntrL + ntrR;
  // slab runs: rank 0 owns the left insertion region, the last rank the right one; both draw from
  // the same replicated RanPark stream (skip-ahead), so no trial data crosses ranks
  const bool doL = (c->myrank == 0), doR = (c->myrank == c->nranks - 1);
  if (!serial) {
    const Grid &g = c->grid;
    const double slab_lo = g.lo[0] + c->gx0 / g.inv[0], slab_hi = g.lo[0] + c->gx1 / g.inv[0];
    // an atom inserted within the cutoff of the inner slab face would act on the neighbour's atoms only
    // through a halo that was exchanged before it existed: keep the insertion regions one cutoff away
    NEED(!doL || c->fd.region[4].hi[0] <= slab_hi - c->pt.cutmax, OBMD_ERR_ARG,
         "slab run: region5 (x <= %g) must end at least one cutoff before the first rank's slab face (x = %g)",
         c->fd.region[4].hi[0], slab_hi);
    NEED(!doR || c->fd.region[5].lo[0] >= slab_lo + c->pt.cutmax, OBMD_ERR_ARG,
         "slab run: region6 (x >= %g) must start at least one cutoff behind the last rank's slab face (x = %g)",
         c->fd.region[5].lo[0], slab_lo);
  }
  const int ntr_loc = (doL ? ntrL : 0) + (doR ? ntrR : 0);
  const int tg0 = doL ? 0 : ntrL;

  int nacc = 0, naccL = 0, naccR = 0, rounds = 0, nused_total = 0, ncopied = 0;
  if ((ntr == 0 || ntr_loc == 0) && c->round0_hook) {  // no USHER work on this rank this step
    int (*hook)(obmd_ctx *) = c->round0_hook;
    c->round0_hook = nullptr;
    if ((rc = hook(c))) return rc;
  }
  if (ntr > 0) {
    ScopedTimer t(c, "usher");
    HostTimer hu(c, "host_usher");
    const int M = fa.maxattempt;
    NEED(M >= 1, OBMD_ERR_ARG, "fix obmd: maxattempt must be >= 1");
    // maxattempt == 1: trial t of either side sits at a fixed place of the RanPark stream, both sides
    // go through the commit rounds together.  maxattempt > 1: the attempts a side consumes depend on
    // its outcomes (k_usher_resolve), so the left side is finished and committed first and the right
    // side starts where it stopped.
    struct Pass { int nL, nR, slots; };
    Pass passes[2];
    int npass = 1;
    if (M == 1) passes[0] = {ntrL, ntrR, 0};
    else {
      passes[0] = {ntrL * M, 0, ntrL};
      passes[1] = {0, ntrR * M, ntrR};
      npass = 2;
    }
    for (int ip = 0; ip < npass; ip++) {
      const Pass &ps = passes[ip];
      const int ntr_p = ps.nL + ps.nR;
      if (ntr_p == 0) continue;
      const int ntr_loc = (doL ? ps.nL : 0) + (doR ? ps.nR : 0);
      const int limit = M > 1 ? 1 : 0;
      int nacc_p = 0, nused_p = ntr_p;
      if (ntr_loc > 0) {
        if ((rc = obmd_ensure_capacity(c, nlive + c->nextra + ntr_loc))) return rc;
        UsherCfg u;
        u.ntrL = ps.nL;
        u.ntr = ntr_p;
        u.seed0 = c->fs.rng_seed;
        u.nbase = nlive + c->nextra;
        u.natt = fa.nattempt;
        u.tg0 = doL ? 0 : ps.nL;
        u.slots = ps.slots;
        u.maxatt = M;
        // round 0, the first diff / dirty test and the fix-up kernel that runs every re-run round are
        // enqueued back to back; the outcome is read back once (usher.cu)
        int rounds_p = 0;
        int nused_loc = 0;
        if ((rc = usher_pass(c, u, ntr_loc, limit, maxtag, &nacc_p, &nused_loc, &rounds_p))) return rc;
        if (limit) nused_p = nused_loc;  // maxattempt == 1: every trial of the pass (on any rank) is made
        rounds += rounds_p;
      }

      if (npass == 1) {
        naccL = doL ? nacc_p : 0;  // serial: both sides live in one list, the split is not needed
        naccR = doR && !doL ? nacc_p : 0;
        if (!serial) {  // creation counts of both edge ranks: natoms and the right rank's first new tag
          ScopedTimer t3(c, "usher_allreduce");
          HostTimer h3(c, "host_allreduce");
          double h[2] = {(double) naccL, (double) naccR};
          CK(cudaMemcpyAsync(c->red_d, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
          if ((rc = comm_allreduce(c, c->red_d, 2, 0))) return rc;
          CK(cudaMemcpyAsync(c->red_h, c->red_d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
          CK(cudaStreamSynchronize(c->stream));
          naccL = (int) c->red_h[0];
          naccR = (int) c->red_h[1];
        }
      }
      int nacc_before = nacc;  // atoms created by earlier passes, on any rank (tags continue behind them)
      if (npass == 2 && !serial) {
        // maxattempt > 1 in a slab run: only the rank owning the side ran the pass; how many attempts it
        // consumed (where the other side starts in the replicated RanPark stream) and how many atoms it
        // created (natoms, the next tag) go to everybody
        ScopedTimer t3(c, "usher_allreduce");
        double h[2] = {ntr_loc > 0 ? (double) nused_p : 0.0, ntr_loc > 0 ? (double) nacc_p : 0.0};
        CK(cudaMemcpyAsync(c->red_d, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
        if ((rc = comm_allreduce(c, c->red_d, 2, 0))) return rc;
        CK(cudaMemcpyAsync(c->red_h, c->red_d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        nused_p = (int) c->red_h[0];
        if (ip == 0) naccL = (int) c->red_h[1];
        else naccR = (int) c->red_h[1];
        nacc_before = (ip == 0) ? 0 : naccL;
      }
      // every attempt made consumes exactly three draws (ATOM mode, uniform distribution, block region)
      c->fs.rng_seed = ranpark_skip(c->fs.rng_seed, 3LL * nused_p);
      nused_total += nused_p;
      if (nacc_p > 0) {
        // the side stream may still be walking the per-cell extra lists: link the new atoms only
        // after its pair kernel has finished (obmd_run sets commit_wait)
        ScopedTimer t4(c, "usher_commit");
        if (c->commit_wait) CK(cudaStreamWaitEvent(c->stream, c->commit_wait, 0));
        const int tag0 = maxtag + nacc_before + ((npass == 1 && doR && !doL) ? naccL : 0);
        if ((rc = usher_commit(c, tag0, nacc_p))) return rc;
        if ((rc = fix_link_extras(c, c->narr + c->nghost + c->nins, nacc_p, false))) return rc;
      }
      if (trials && maxtrials > ncopied && ntr_loc > 0) {
        const int nloc_used = limit ? (nused_p < ntr_loc ? nused_p : ntr_loc) : ntr_loc;
        const int ncopy = nloc_used < maxtrials - ncopied ? nloc_used : maxtrials - ncopied;
        if ((rc = usher_trials_to_host(c, trials + ncopied, ncopy))) return rc;
        ncopied += ncopy;
      }
      // the next pass sees this pass's atoms as appended atoms (per-cell lists, local indices)
      nacc += nacc_p;
      c->nins = nacc;
      c->nextra = c->narr + c->nghost + c->nins;
      c->n = nlive + c->nextra;
    }
    rep.usher_rounds = rounds;
    c->fs.ntrials += nused_total;
    c->fs.nrounds += rounds;
    CK(cudaGetLastError());
  }
  c->nins = nacc;
  c->nextra = c->narr + c->nghost + c->nins;
  c->n = nlive + c->nextra;
  const int nacc_all = serial ? nacc : naccL + naccR;
  c->fs.ninserted += nacc_all;
  rep.ntrials = nused_total;
  rep.naccept = nacc_all;
  rep.nlocal = nlive + c->narr + c->nins;

  struct { double vnewl[3], vnewr[3]; } s;
  for (int d = 0; d < 3; d++) { s.vnewl[d] = vnewl[d]; s.vnewr[d] = vnewr[d]; }
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


// Pass 1 (g_par_global_charged :1244, g_perp_global_charged :1344): weight sums of the four
// regions.  Fixed grid, grid-stride over the contiguous slot ranges, warp-shuffle + block tree,
// one partial per block: the final value does not depend on scheduling.
__global__ void __launch_bounds__(FB_THREADS)
    k_weight_sums(Planes p, Extras ex, Grid g, FixDev fd, PairTab pt, ForceCfg fc,
                  const int *__restrict__ cell_start, double *__restrict__ partials, double *__restrict__ gsum,
                  int *__restrict__ done)
{
  double s[4] = {0, 0, 0, 0};
  const int nyz = g.ny * g.nz;
  for (int iv = 0; iv <= fc.nint; iv++) {
    int b, e;
    if (iv < fc.nint) {
      b = cell_start[fc.cxa[iv] * nyz];
      e = cell_start[(fc.cxb[iv] + 1) * nyz];
    } else {  // atoms appended after the sort that this rank owns (arrivals, insertions)
      b = 0;
      e = ex.narr + ex.nins;
    }
    for (int ii = b + blockIdx.x * blockDim.x + threadIdx.x; ii < e; ii += gridDim.x * blockDim.x) {
      const int i = (iv < fc.nint) ? ii : extras_slot(ex, 1, ii);
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
  __shared__ int s_last;
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
  // The block that finishes last adds the per-block partials in a fixed order (chunks of 16 blocks, then
  // the chunks): G does not depend on which block that is, and no second launch is needed.
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done, 1) == (int) gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double ch[4][64];
  const int nb = (int) gridDim.x, nch = (nb + 15) / 16;  // FB_BLOCKS <= 1024
  if (threadIdx.x < 4 * 64) {
    const int r = threadIdx.x & 3, q = threadIdx.x >> 2;
    double v = 0.0;
    if (q < nch)
      for (int b = q * 16; b < min(q * 16 + 16, nb); b++) v += __ldcg(partials + b * 4 + r);
    ch[r][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int q = 0; q < nch; q++) v += ch[threadIdx.x][q];
    gsum[threadIdx.x] = v;
    if (threadIdx.x == 0) *done = 0;
  }
}

// Pass 2 (reg_force :1414-1428, reg_force_perp :1480-1494): f += F * w / G.  Each block first
// re-reduces the FB_BLOCKS partials in a fixed order (no extra launch, no atomics).
__global__ void __launch_bounds__(FB_THREADS)
    k_apply_forces(Planes p, Extras ex, Grid g, FixDev fd, PairTab pt, ForceCfg fc,
                   const int *__restrict__ cell_start, const double *__restrict__ gsum,
                   double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
                   StepScalars *sc)
{
  __shared__ double G[4];
  if (threadIdx.x < 4) {
    G[threadIdx.x] = gsum[threadIdx.x];
    if (blockIdx.x == 0) sc->gsum[threadIdx.x] = gsum[threadIdx.x];
  }
  __syncthreads();
  const int nyz = g.ny * g.nz;
  for (int iv = 0; iv <= fc.nint; iv++) {
    int b, e;
    if (iv < fc.nint) {
      b = cell_start[fc.cxa[iv] * nyz];
      e = cell_start[(fc.cxb[iv] + 1) * nyz];
    } else {  // atoms appended after the sort that this rank owns (arrivals, insertions)
      b = 0;
      e = ex.narr + ex.nins;
    }
    for (int ii = b + blockIdx.x * blockDim.x + threadIdx.x; ii < e; ii += gridDim.x * blockDim.x) {
      const int i = (iv < fc.nint) ? ii : extras_slot(ex, 1, ii);
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

static void region_cols(const obmd_ctx *c, int r, int margin, int &a, int &e);
static int merge_intervals(int *lo, int *hi, int n);

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
  int lo[4], hi[4], nr = 0, nactive = 0;
  for (int r = 0; r < 4; r++) {
    for (int d = 0; d < 3; d++) fc.F[r][d] = F[r][d];
    fc.active[r] = (F[r][0] != 0.0 || F[r][1] != 0.0 || F[r][2] != 0.0);
    if (!fc.active[r]) continue;
    nactive++;
    int a, e;
    region_cols(c, r, 0, a, e);  // local columns; empty when the region misses this slab
    if (e < a) continue;
    lo[nr] = a;
    hi[nr] = e;
    nr++;
  }
  if (nactive == 0) return OBMD_OK;  // identical on every rank: the forces are replicated
  nr = merge_intervals(lo, hi, nr);
  fc.nint = nr;
  for (int k = 0; k < nr; k++) {
    fc.cxa[k] = lo[k];
    fc.cxb[k] = hi[k];
  }
  if (c->partial_cap < FB_BLOCKS * 4) {
    if (c->partials) cudaFree(c->partials);
    c->partial_cap = FB_BLOCKS * 4 + 1024;
    CK(cudaMalloc(&c->partials, sizeof(double) * c->partial_cap));
  }
  const Extras ex = extras_of(c);
  LAUNCH(k_weight_sums, FB_BLOCKS, FB_THREADS, 0, c->cur, ex, c->grid, c->fd, c->pt, fc, c->cell_start,
         c->partials, c->red_d, &c->sc_d->wsum_done);
  if (c->nranks > 1) {  // a buffer may span slabs: G is a global sum (fix_obmd_merged.cpp:1305,1378)
    int rc = comm_allreduce(c, c->red_d, 4, 0);
    if (rc) return rc;
  }
  LAUNCH(k_apply_forces, FB_BLOCKS, FB_THREADS, 0, c->cur, ex, c->grid, c->fd, c->pt, fc, c->cell_start,
         c->red_d, c->fx, c->fy, c->fz, c->sc_d);
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

// 1-D profiles along x (what `compute chunk/atom bin/1d x` + `fix ave/chunk` accumulate): per bin the
// atom count, sum m vx, sum m v^2 and sum of the per-atom stress xx = -(m vx^2 + W_xx)
// (compute_stress_atom.cpp: kinetic + virial contribution, pressure*volume units).
__global__ void k_profile_x(Planes p, int n, PairTab pt, double lo, double inv, int nbins,
                            const double *__restrict__ vatom, double *__restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b = (int) floor((p.x[i] - lo) * inv);
  b = min(max(b, 0), nbins - 1);
  const double m = pt.mass[p.type[i]], vx = p.vx[i], vy = p.vy[i], vz = p.vz[i];
  atomicAdd(&out[b], 1.0);
  atomicAdd(&out[nbins + b], m * vx);
  atomicAdd(&out[2 * nbins + b], m * (vx * vx + vy * vy + vz * vz));
  if (vatom) atomicAdd(&out[3 * nbins + b], -(m * vx * vx + vatom[i]));
}

int fix_profile_x(obmd_ctx *c, int nbins, double *count, double *mvx, double *mv2, double *sxx)
{
  NEED(nbins > 0 && nbins <= 65536, OBMD_ERR_ARG, "obmd_profile_x: bad nbins");
  NEED(!sxx || (c->vatom && c->vatom_valid), OBMD_ERR_STATE,
       "obmd_profile_x2: the stress profile needs obmd_pair_compute(..., vflag = 1) on the current positions");
  double *d = nullptr;
  CK(cudaMalloc(&d, sizeof(double) * 4 * nbins));
  CK(cudaMemsetAsync(d, 0, sizeof(double) * 4 * nbins, c->stream));
  LAUNCH(k_profile_x, cdiv(c->n, 256), 256, 0, c->cur, c->n, c->pt, c->boxlo[0],
         nbins / (c->boxhi[0] - c->boxlo[0]), nbins, sxx ? c->vatom : nullptr, d);
  CK(cudaMemcpyAsync(count, d, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(mvx, d + nbins, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
  if (mv2) CK(cudaMemcpyAsync(mv2, d + 2 * nbins, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
  if (sxx) CK(cudaMemcpyAsync(sxx, d + 3 * nbins, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return OBMD_OK;
}
