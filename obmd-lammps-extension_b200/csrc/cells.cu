// Binned cell list build: replaces NBinStandard::bin_atoms (nbin_standard.cpp:221-261),
// NBin::coord2bin (nbin.cpp:139-170) and -- because the pair kernel walks cells directly --
// the role of NPairBin::build (npair_bin.cpp:52-253) and Atom::sort (atom.cpp:2288-2378).
//
// Counting sort of the SoA planes by cell: k_bin (cell id + arrival rank + histogram) ->
// exclusive scan -> k_perm (slot -> old slot) -> k_cell_order (sort each cell's few entries by
// old slot so the layout, and with it every fp64 summation order downstream, is independent of
// atomic arrival order) -> k_gather (move the 10 planes).  Atoms outside the box in x get the
// key ncell and land behind the live range, where pre_exchange deletes them.
//
// All kernels are HBM-bound streaming kernels: coalesced SoA accesses, grids sized to cover n.
#include "ctx.h"

__global__ void k_bin(Planes p, int n, Grid g, int *__restrict__ cid, int *__restrict__ rank,
                      int *__restrict__ cell_count, StepScalars *sc)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = p.x[i], y = p.y[i], z = p.z[i];
  if (!(isfinite(x) && isfinite(y) && isfinite(z))) {
    sc->nonfinite = 1;  // domain.cpp:540-545
    x = g.lo[0] - 1.0;  // park it
    y = g.lo[1];
    z = g.lo[2];
  }
  int c;
  if (x < g.lo[0] || x > g.hi[0])  // fix_obmd_merged.cpp:669: strictly outside [boxlo, boxhi]
    c = g.ncell;
  else {
    int lc = cell_cx(g, x);
    if (lc < g.own_a) c = g.ncell + 1;       // left the slab: migrates to the left neighbour
    else if (lc > g.own_b) c = g.ncell + 2;  // ... to the right neighbour
    else c = (lc * g.ny + cell_cy(g, y)) * g.nz + cell_cz(g, z);
  }
  cid[i] = c;
  rank[i] = atomicAdd(&cell_count[c], 1);
}

// ---- exclusive scan of cell_count[0..m) into cell_start[0..m], three small kernels -----
#define SCAN_T 512
#define SCAN_E 8  // elements per thread -> 4096 per block

__global__ void k_scan_block(const int *__restrict__ in, int *__restrict__ out, int *__restrict__ sums,
                             int m)
{
  __shared__ int warp_tot[SCAN_T / 32];
  int base = blockIdx.x * SCAN_T * SCAN_E + threadIdx.x * SCAN_E;
  int v[SCAN_E], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    v[k] = (base + k < m) ? in[base + k] : 0;
    s += v[k];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    int t = (lane < SCAN_T / 32) ? warp_tot[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, t, d);
      if (lane >= d) t += u;
    }
    if (lane < SCAN_T / 32) warp_tot[lane] = t;  // inclusive over warps
  }
  __syncthreads();
  int excl = inc - s + (w ? warp_tot[w - 1] : 0);
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    if (base + k < m) out[base + k] = excl;
    excl += v[k];
  }
  if (threadIdx.x == SCAN_T - 1) sums[blockIdx.x] = warp_tot[SCAN_T / 32 - 1];
}

__global__ void k_scan_sums(int *sums, int nb)
{
  // one block; sequential over chunks of blockDim with a running carry
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += blockDim.x) {
    int i = b0 + threadIdx.x;
    int v = (i < nb) ? sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < blockDim.x; d <<= 1) {
      int t = (threadIdx.x >= d) ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) sums[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += sh[threadIdx.x];
    __syncthreads();
  }
}

__global__ void k_scan_add(int *__restrict__ out, const int *__restrict__ sums, int m, int total_slot,
                           int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] += sums[i / (SCAN_T * SCAN_E)];
  if (i == 0) out[total_slot] = n;
}

__global__ void k_perm(const int *__restrict__ cid, const int *__restrict__ rank,
                       const int *__restrict__ cell_start, int *__restrict__ perm, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) perm[cell_start[cid[i]] + rank[i]] = i;
}

// One thread per cell: its entries are few (rho * rc^3 ~ 3), insertion sort in place.
// Thread c < nx also reports the population of cell column c (the density profile along the open axis is
// not flat: the pair kernel sizes its shared-memory window for the fullest column).
__global__ void k_cell_order(const int *__restrict__ cell_start, int *__restrict__ perm, int ncell, int nx, int nyz,
                             StepScalars *sc)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nx) atomicMax(&sc->colmax, cell_start[(c + 1) * nyz] - cell_start[c * nyz]);
  if (c >= ncell) return;
  int b = cell_start[c], e = cell_start[c + 1];
  for (int a = b + 1; a < e; a++) {
    int v = perm[a], k = a - 1;
    while (k >= b && perm[k] > v) {
      perm[k + 1] = perm[k];
      k--;
    }
    perm[k + 1] = v;
  }
}

// Permutes the planes into cell order and writes the records the pair kernel stages by bulk copy
// (pair_brick.cuh): four 16-byte stores per thread, contiguous per warp.
__global__ void k_gather(Planes src, Planes dst, AtomRec *__restrict__ rec, int4 *__restrict__ aux,
                         const int *__restrict__ perm, int n, int renumber)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int j = perm[i];
  const double x = src.x[j], y = src.y[j], z = src.z[j], vx = src.vx[j], vy = src.vy[j], vz = src.vz[j];
  const int tag = src.tag[j], type = src.type[j];
  dst.x[i] = x;
  dst.y[i] = y;
  dst.z[i] = z;
  dst.vx[i] = vx;
  dst.vy[i] = vy;
  dst.vz[i] = vz;
  dst.tag[i] = tag;
  dst.lidx[i] = renumber ? i : src.lidx[j];  // slab mode: local order = storage order
  dst.type[i] = type;
  dst.image[i] = src.image[j];
  double2 *r = reinterpret_cast<double2 *>(rec + i);
  r[0] = make_double2(x, y);
  r[1] = make_double2(z, vx);
  r[2] = make_double2(vy, vz);
  aux[i] = make_int4(tag, type, 0, 0);
}

int cells_build(obmd_ctx *c)
{
  NEED(c->pair_ready, OBMD_ERR_STATE, "obmd_build_cells: call obmd_pair_init first");
  NEED(c->n > 0, OBMD_ERR_STATE, "obmd_build_cells: no atoms uploaded");
  const Grid &g = c->grid;
  const int n = c->n, m = g.ncell + 3;  // + parked, migrate-left, migrate-right
  {
    ScopedTimer t(c, "bin");
    CK(cudaMemsetAsync(c->cell_count, 0, sizeof(int) * (m + 1), c->stream));
    CK(cudaMemsetAsync(&c->sc_d->nonfinite, 0, 2 * sizeof(int), c->stream));  // nonfinite, colmax
    LAUNCH(k_bin, cdiv(n, 256), 256, 0, c->cur, n, g, c->cid, c->rank, c->cell_count, c->sc_d);
    int nb = cdiv(m, SCAN_T * SCAN_E);
    LAUNCH(k_scan_block, nb, SCAN_T, 0, c->cell_count, c->cell_start, c->scan_tmp, m);
    LAUNCH(k_scan_sums, 1, 1024, 0, c->scan_tmp, nb);
    LAUNCH(k_scan_add, cdiv(m, 256), 256, 0, c->cell_start, c->scan_tmp, m, m, n);
  }
  {
    ScopedTimer t(c, "sort");
    LAUNCH(k_perm, cdiv(n, 256), 256, 0, c->cid, c->rank, c->cell_start, c->perm, n);
    LAUNCH(k_cell_order, cdiv(g.ncell, 128), 128, 0, c->cell_start, c->perm, g.ncell, g.nx, g.ny * g.nz, c->sc_d);
    LAUNCH(k_gather, cdiv(n, 256), 256, 0, c->cur, c->alt, c->rec, c->aux, c->perm, n, c->nranks > 1 ? 1 : 0);
    Planes t2 = c->cur;
    c->cur = c->alt;
    c->alt = t2;
  }
  // segment starts (sorted | parked | migrate-left | migrate-right) are needed on the host
  int *seg = reinterpret_cast<int *>(c->red_h);
  CK(cudaMemcpyAsync(seg, &c->cell_start[g.ncell], sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (c->nranks > 1) {  // slot ranges of the two boundary columns: what the halo exchange sends
    const int nyz = g.ny * g.nz;
    CK(cudaMemcpyAsync(seg + 4, &c->cell_start[g.own_a * nyz], sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(seg + 5, &c->cell_start[(g.own_a + 1) * nyz], sizeof(int), cudaMemcpyDeviceToHost,
                       c->stream));
    CK(cudaMemcpyAsync(seg + 6, &c->cell_start[g.own_b * nyz], sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(seg + 7, &c->cell_start[(g.own_b + 1) * nyz], sizeof(int), cudaMemcpyDeviceToHost,
                       c->stream));
  }
  CK(cudaMemcpyAsync(&c->sc_h->nonfinite, &c->sc_d->nonfinite, 2 * sizeof(int), cudaMemcpyDeviceToHost,
                     c->stream));  // nonfinite, colmax
  CK(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 4; k++) c->col_ab[k] = (c->nranks > 1) ? seg[4 + k] : 0;
  NEED(!c->sc_h->nonfinite, OBMD_ERR_NONFINITE, "Non-numeric atom coords - simulation unstable");
  c->nsorted = seg[0];
  c->colmax = c->sc_h->colmax;
  c->nparked = seg[1] - seg[0];
  c->seg_migL = seg[1];
  c->seg_migR = seg[2];
  c->nmigL = seg[2] - seg[1];
  c->nmigR = seg[3] - seg[2];
  NEED(c->nranks > 1 || (c->nmigL == 0 && c->nmigR == 0), OBMD_ERR_STATE, "cell build: stray migrants");
  c->nextra = c->narr = c->nghost = c->nins = 0;
  c->cells_valid = true;
  c->rec_valid = true;
  c->forces_valid = false;
  c->vatom_valid = false;
  return OBMD_OK;
}
