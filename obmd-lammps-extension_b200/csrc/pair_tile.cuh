// Tiled PairDPD kernel (included by pair.cu): one CTA per cell row (chunk), neighbour positions
// staged into shared memory by TMA bulk copies.
//
// A tile is the atoms of one (column, row) pair over up to P3_ZC z-cells: one contiguous slot range,
// because z is the fastest cell index.  Its stencil is nine rows over the same z window plus one
// cell on each side; per row that is one slot-contiguous piece, plus a single wrapped cell at a
// periodic face.  One thread issues `cp.async.bulk` copies of the x, y, z planes of those pieces
// (completion counted on an mbarrier) while the other threads build the index tables.  In shared
// memory the cells of a row sit in window order, so the stencil cells cz-1, cz, cz+1 of an atom are
// at most two contiguous index ranges in every row, also across the wrap.
//   phase 1  screen the ~81 candidates from shared memory, list the hits (16-bit staged indices)
//   phase 2  evaluate the ~12.6 hits densely; positions from shared memory, v and tag from L1/L2
// The accumulation order is exactly that of pair_atom_global (pair.cu), so both kernels give the
// same bits for the same atom.
#pragma once

#define P3_T 160     // threads per CTA (a 44-cell row holds 132 +- 12 atoms at rho = 3)
#define P3_ZC 48     // z-cells per tile at most
#define P3_W (P3_ZC + 2)
#define P3_CAP 1344  // staged atoms (9 rows x 46 cells x 3 = 1242 on average for a 44-cell row)
#define P3_PARTS 27

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
  return (unsigned) __cvta_generic_to_shared(p);
}

struct TileHdr {
  unsigned long long bar;       // mbarrier the bulk copies complete on
  int part_gb[P3_PARTS];        // first global slot staged for the piece (even: 16-byte aligned source)
  int part_off[P3_PARTS];       // its index in the staged arrays (even)
  int part_n[P3_PARTS];         // staged elements (even)
  int row_base[9];              // first cell index of stencil row r, -1 if that column does not exist
  int nparts, nstaged, tb, te;  // tile atoms are slots [tb, te)
  int wbeg[9][P3_W], wend[9][P3_W];  // staged index range of window cell w (w = 0 is cell z0-1) of row r
};

#define P3_SMEM (3 * 8 * P3_CAP + 4 * P3_CAP + 2 * PAIR_HCAP * P3_T)

template <bool SINGLE_TYPE, bool PREWRAP>
__global__ void __launch_bounds__(P3_T)
    k_pair_tile(Planes p, Extras ex, Grid g, PairTab pt, const int *__restrict__ cell_start,
                const int *__restrict__ extra_head, const int *__restrict__ extra_next,
                double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz, unsigned step,
                int random_on, int nchunk, int zc)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sx = reinterpret_cast<double *>(smem_raw);
  double *sy = sx + P3_CAP, *sz = sy + P3_CAP;
  int *sg = reinterpret_cast<int *>(sz + P3_CAP);                          // staged index -> global slot
  unsigned short *hits = reinterpret_cast<unsigned short *>(sg + P3_CAP);  // [PAIR_HCAP][P3_T]
  __shared__ TileHdr H;

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % nchunk, rowid = blockIdx.x / nchunk;
  const int cy = rowid % g.ny, lc = g.own_a + rowid / g.ny;
  const int z0 = chunk * zc, z1 = min(z0 + zc, g.nz);  // tile cells [z0, z1)
  const int nw = z1 - z0 + 2;                          // window cells z0-1 .. z1
  const bool lowrap = (z0 == 0), hiwrap = (z1 == g.nz);

  if (tid < 9) {
    int bx = lc + tid / 3 - 1, by = cy + tid % 3 - 1;
    by = by < 0 ? by + g.ny : (by >= g.ny ? by - g.ny : by);
    H.row_base[tid] = (bx < 0 || bx >= g.nx) ? -1 : (bx * g.ny + by) * g.nz;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&H.bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // ---- pieces: [wrapped low cell] [main run of cells] [wrapped high cell] per existing row --------
  // Lane l < 27 of warp 0 owns piece l = 3 * row + kind: one global round trip for all slot ranges,
  // a warp scan for the staged offsets, then every lane issues the TMA copies of its own piece.
  if (tid < 32) {
    const int r = tid / 3, kind = tid % 3;
    int gal = 0, cnt = 0;
    if (tid < 27 && H.row_base[r] >= 0 && (kind == 1 || (kind == 0 && lowrap) || (kind == 2 && hiwrap))) {
      const int ca = kind == 0 ? g.nz - 1 : (kind == 1 ? max(z0 - 1, 0) : 0);
      const int cb = kind == 0 ? g.nz - 1 : (kind == 1 ? min(z1, g.nz - 1) : 0);
      const int gb = cell_start[H.row_base[r] + ca], ge = cell_start[H.row_base[r] + cb + 1];
      gal = gb & ~1;
      cnt = ((ge - gal) + 1) & ~1;
    }
    if (tid == 27) H.tb = cell_start[(lc * g.ny + cy) * g.nz + z0];
    if (tid == 28) H.te = cell_start[(lc * g.ny + cy) * g.nz + z1];
    int off = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, off, d);
      if (tid >= d) off += v;
    }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    off -= cnt;
    if (tid < 27) {
      H.part_gb[tid] = gal;
      H.part_off[tid] = off;
      H.part_n[tid] = cnt;
    }
    if (tid == 0) {
      H.nparts = 27;
      H.nstaged = total;
    }
    if (total <= P3_CAP) {
      if (tid == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&H.bar)),
                     "r"(3u * 8u * (unsigned) total)
                     : "memory");
      __syncwarp();
      if (cnt > 0) {  // TMA: x, y, z planes of this piece, all counted on the one mbarrier
        const unsigned nb = 8u * (unsigned) cnt;
        const double *src[3] = {p.x + gal, p.y + gal, p.z + gal};
        double *dst[3] = {sx + off, sy + off, sz + off};
#pragma unroll
        for (int d = 0; d < 3; d++)
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(dst[d])),
              "l"(src[d]), "r"(nb), "r"(smem_u32(&H.bar))
              : "memory");
      }
    }
  }
  __syncthreads();
  const bool staged = H.nstaged <= P3_CAP;
  if (staged) {
    // meanwhile the other threads build the window tables and the staged-index -> slot map
    for (int q = tid; q < 9 * nw; q += P3_T) {
      const int r = q / nw, w = q % nw;
      int b = 0, e = 0;
      if (H.row_base[r] >= 0) {
        const int zu = z0 - 1 + w;  // unwrapped cell of window position w
        int pi = 3 * r, cell;
        if (zu < 0) {
          cell = g.nz - 1;          // the wrapped low cell: piece kind 0
        } else if (zu >= g.nz) {
          cell = zu - g.nz;         // the wrapped high cell: piece kind 2
          pi += 2;
        } else {
          cell = zu;                // main run: piece kind 1
          pi += 1;
        }
        b = H.part_off[pi] + (cell_start[H.row_base[r] + cell] - H.part_gb[pi]);
        e = H.part_off[pi] + (cell_start[H.row_base[r] + cell + 1] - H.part_gb[pi]);
      }
      H.wbeg[r][w] = b;
      H.wend[r][w] = e;
    }
    for (int k = 0; k < H.nparts; k++)
      for (int q = tid; q < H.part_n[k]; q += P3_T) sg[H.part_off[k] + q] = H.part_gb[k] + q;
    unsigned ok = 0;
    while (!ok)
      asm volatile(
          "{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }"
          : "=r"(ok)
          : "r"(smem_u32(&H.bar))
          : "memory");
  }
  __syncthreads();

  const int nextra = ex.narr + ex.nghost + ex.nins;
  const int ntile = H.te - H.tb;
  for (int k0 = 0; k0 < ntile; k0 += P3_T) {
    const int i = H.tb + k0 + tid;
    const bool mine = (k0 + tid) < ntile;
    const unsigned live = __ballot_sync(0xffffffffu, mine);
    if (!mine) continue;
    PairAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (!staged) {  // unusually dense neighbourhood: plain global path, same accumulation order
      pair_atom_global<SINGLE_TYPE, false, PREWRAP, int>(p, g, pt, ex, cell_start, extra_head, extra_next, i,
                                                         live, step, random_on != 0,
                                                         reinterpret_cast<int *>(smem_raw) + tid, P3_T, acc);
      fx[i] = acc.fx;
      fy[i] = acc.fy;
      fz[i] = acc.fz;
      continue;
    }
    AtomI a;
    a.x = p.x[i];
    a.y = PREWRAP ? wrap1(p.y[i], g.lo[1], g.hi[1], g.prd[1]) : p.y[i];
    a.z = PREWRAP ? wrap1(p.z[i], g.lo[2], g.hi[2], g.prd[2]) : p.z[i];
    a.vx = p.vx[i]; a.vy = p.vy[i]; a.vz = p.vz[i];
    a.tag = p.tag[i];
    a.type = SINGLE_TYPE ? 1 : p.type[i];
    const int cz = min(max(cell_cz(g, a.z), z0), z1 - 1);  // the sort key of a tile atom lies in [z0, z1)
    const int lay = PREWRAP ? 1 : 0;
    a.bnd = (cy <= lay || cy >= g.ny - 1 - lay || cz <= lay || cz >= g.nz - 1 - lay);
    a.bnd = __any_sync(live, a.bnd);
    const int w = cz - z0 + 1;  // window position of the atom's own cell
    unsigned short *myhits = hits + tid;
    int nh = 0;
    for (int r = 0; r < 9; r++) {
      if (H.row_base[r] < 0) continue;
      // staged ranges of window cells w-1, w, w+1, merged where they touch
      int rb[3], re[3], nr = 0;
#pragma unroll
      for (int ww = -1; ww <= 1; ww++) {
        const int b = H.wbeg[r][w + ww], e = H.wend[r][w + ww];
        if (nr > 0 && re[nr - 1] == b) re[nr - 1] = e;
        else {
          rb[nr] = b;
          re[nr] = e;
          nr++;
        }
      }
      for (int sgi = 0; sgi < nr; sgi++) {
        const int qe = re[sgi];
        for (int q0 = rb[sgi]; q0 < qe; q0 += 4) {
          bool h[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int q = min(q0 + u, qe - 1);
            double delx = a.x - sx[q], dely, delz;
            if (a.bnd) {
              dely = a.y - (PREWRAP ? wrap1(sy[q], g.lo[1], g.hi[1], g.prd[1]) : sy[q]);
              delz = a.z - (PREWRAP ? wrap1(sz[q], g.lo[2], g.hi[2], g.prd[2]) : sz[q]);
              min_image(g, dely, delz);
            } else {
              dely = a.y - sy[q];
              delz = a.z - sz[q];
            }
            const double rsq = delx * delx + dely * dely + delz * delz;
            const int tj = SINGLE_TYPE ? 1 : p.type[sg[q]];
            h[u] = (rsq < pt.cutsq[a.type][tj]) && (q0 + u < qe) && (sg[q] != i);
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (h[u]) {
              if (nh < PAIR_HCAP) {
                myhits[nh * P3_T] = (unsigned short) (q0 + u);
                nh++;
              } else {
                pair_eval<SINGLE_TYPE, false, PREWRAP>(p, pt, g, sg[q0 + u], a, step, random_on != 0, acc);
              }
            }
          }
        }
      }
      if (nextra > 0) {  // atoms appended after the sort hang off per-cell lists (global memory)
        for (int oz = -1; oz <= 1; oz++) {
          int bz = cz + oz;
          bz = bz < 0 ? bz + g.nz : (bz >= g.nz ? bz - g.nz : bz);
          for (int e = extra_head[H.row_base[r] + bz]; e >= 0; e = extra_next[e]) {
            const int j = ex.base + e;
            if (j == i) continue;
            if (pair_screen<SINGLE_TYPE, PREWRAP>(p, pt, g, j, a))
              pair_eval<SINGLE_TYPE, false, PREWRAP>(p, pt, g, j, a, step, random_on != 0, acc);
          }
        }
      }
    }
    for (int k = 0; k < nh; k++) {
      const int q = myhits[k * P3_T];
      pair_eval_xyz<SINGLE_TYPE, false, PREWRAP>(p, pt, g, sg[q], sx[q], sy[q], sz[q], a, step, random_on != 0,
                                                 acc);
    }
    fx[i] = acc.fx;
    fy[i] = acc.fy;
    fz[i] = acc.fz;
  }
}
