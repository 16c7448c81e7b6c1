// Brick-tiled PairDPD kernel (included by pair.cu): the production path for cell-sorted atoms whose
// stencil holds no appended atoms (every tile of a single-GPU step; the interior columns of a slab).
// pair_dpd.cpp:66-163 with NStencilBin/NPairBin folded in (nstencil_bin.cpp, npair_bin.cpp:52-253).
//
// A tile is a brick of C columns x R rows x zc z-cells (C, R in {1, 2, 4}).  Its window -- the brick plus
// one cell all round -- is (C+2)(R+2) cell ROWS, and because z is the fastest cell index the part of a
// row the window needs is ONE contiguous run of slots.  The sort's gather leaves, next to the planes, a
// 48-byte record (x, v) and a 16-byte record (tag, type) per atom, so a row is staged by TWO bulk copies
// of the copy engine (`cp.async.bulk`, UBLKCP) whatever its length: ~50-70 copies per CTA of ~250 atoms,
// issued by one warp against one mbarrier; no thread touches a staged atom before it sits in shared
// memory.  Rows are kept in memory order ("landing" order), so there is no reordering pass and any
// brick shape works: a brick of 4 x 4 x 5 cells stages 3.2 atoms per tile atom where the single-column
// row tile of pair_row.cuh staged 5.3, and its CTA has 8 warps where that one had 5.
//   * the cell behind a periodic z face (tiles at either end of a row) is a separate small copy landing
//     directly in front of / behind the row's run, so the row stays contiguous;
//   * one elementwise pass (a warp per row) applies the periodic image shift -- exactly the reference's
//     ghost atoms: comm_brick.cpp:735-760 adds the period once, pair_dpd.cpp:101-104 subtracts plain
//     coordinates -- and, for atoms Domain::pbc has not wrapped yet, the wrap itself (k_pbc's
//     arithmetic), in place, and overwrites the 16-byte record with {fp32 position relative to the
//     tile corner, tag}: the screening copy;
//   * phase 1: each thread screens the 9 row segments of its atom (cells w-1..w+1 of the 3 x 3 rows
//     around it) in fp32 with a conservative margin, one LDS.128 per candidate, hit bits compacted to
//     a per-thread list of landing indices;
//   * phase 2: the ~12.6 listed hits are evaluated in fp64 two at a time (exact cutoff test first).
// Accumulation order (a function of the layout only): rows (dx, dy) ascending, slots ascending.
#pragma once

#define PB_TMAX 256
#define PB_HCAP 32        // hit slots per thread (mean 12.6 at rho = 3); more -> test-and-evaluate path
#define PB_MAXROWS 36     // (C+2)(R+2), C, R <= 4
#define PB_MAXZC 60

struct BrickHdr {
  int row_base[PB_MAXROWS];   // first cell of window row r = wx*(R+2) + wy; -1: not needed / column outside the grid
  double yshift[PB_MAXROWS];
  int main_g[PB_MAXROWS], main_n[PB_MAXROWS];  // run of the row over the window's z-cells inside the box: first slot, atoms
  int wlo_g[PB_MAXROWS], wlo_n[PB_MAXROWS];    // the cell behind the low periodic z face (tile starts at z = 0)
  int whi_g[PB_MAXROWS], whi_n[PB_MAXROWS];    // the cell behind the high face (tile ends at z = nz)
  int row_l[PB_MAXROWS + 1];                   // landing offset of the row; [NROW] = staged atoms
  int tb[16], tcum[17];                        // tile rows tr = tc*R + ty: first slot, running atom count
  int yface[PB_MAXROWS];                       // row lies in a periodic y face layer
  int issued;
  double org[3];                               // low corner of the tile (origin of the fp32 screening copy)
};

struct BrickSmem {
  AtomRec *rec;          // [capL] landing array: positions and velocities
  float4 *scr;           // [capL + 4] lands as {tag, type, 0, 0}; then {fp32 position relative to the corner, tag}
  int *cb;               // [NROW][nw + 1] landing index where window cell w of row r starts
  unsigned short *hits;  // [PB_HCAP][T] landing indices of each thread's hits
  unsigned char *type;   // [capL] (multi-type systems only)
};

static inline size_t pair_brick_smem_bytes(int C, int R, int capL, int nw, int T, bool single_type)
{
  const int NROW = (C + 2) * (R + 2);
  size_t b = (size_t) capL * 48 + (size_t) (capL + 4) * 16 + (size_t) 4 * (NROW * (nw + 1) + 4) + (size_t) 2 * PB_HCAP * T;
  if (!single_type) b += (size_t) capL + 16;
  const size_t fallback = (size_t) T * PAIR_HCAP * 4;  // int hit lists of pair_atom_global
  if (b < fallback) b = fallback;
  return (b + 15) / 16 * 16;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// global -> shared bulk copy by the copy engine; 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_global, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_global), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 1/sqrt(x) for 1e-20 <= x <= cutoff^2: fp32 MUFU.RSQ seed (relative error 2^-22), two Newton steps in
// fp64 (1.5 e^2 each: 5e-14, then rounding level).  CUDA's rsqrt() spends ~30 instructions on the same
// value because it also handles denormals, infinities and negative arguments.
__device__ __forceinline__ double brick_rsqrt(double x)
{
  float s;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"((float) x));
  double y = (double) s;
  const double h = 0.5 * x;
  y = fma(y, fma(-h, y * y, 0.5), y);
  y = fma(y, fma(-h, y * y, 0.5), y);
  return y;
}

struct BrickAtom {
  double x, y, z, vx, vy, vz;
  int tag, type;
};

__device__ __forceinline__ void brick_load(const BrickSmem &S, int l, double &x, double &y, double &z, double &vx,
                                           double &vy, double &vz)
{
  const double2 *r = reinterpret_cast<const double2 *>(S.rec + l);
  const double2 a = r[0], b = r[1], c = r[2];
  x = a.x; y = a.y; z = b.x; vx = b.y; vy = c.x; vz = c.y;
}

// pair_dpd.cpp:117-146 for the staged atom with landing index l (test-and-evaluate path)
template <bool SINGLE_TYPE>
__device__ __forceinline__ void pair_brick_eval(const BrickSmem &S, const PairTab &pt, int l, const BrickAtom &a,
                                                unsigned step, bool random_on, double &fxa, double &fya,
                                                double &fza)
{
  double xj, yj, zj, vxj, vyj, vzj;
  brick_load(S, l, xj, yj, zj, vxj, vyj, vzj);
  const double delx = a.x - xj, dely = a.y - yj, delz = a.z - zj;
  const double rsq = delx * delx + dely * dely + delz * delz;
  const int tj = SINGLE_TYPE ? 1 : (int) S.type[l];
  if (rsq >= pt.cutsq[a.type][tj] || rsq < EPSILON_DPD * EPSILON_DPD) return;
  const double rinv = brick_rsqrt(rsq);
  const double r = rsq * rinv;
  const double dot = delx * (a.vx - vxj) + dely * (a.vy - vyj) + delz * (a.vz - vzj);
  const double wd = 1.0 - r * pt.cutinv[a.type][tj];
  double t = pt.a0[a.type][tj] - pt.gamma[a.type][tj] * wd * dot * rinv;
  if (random_on)
    t += pt.sigma[a.type][tj] * pt.dtinvsqrt * dpd_gaussian(a.tag, __float_as_int(S.scr[l].w), step, pt.seed);
  const double fpair = wd * rinv * t;
  fxa += delx * fpair;
  fya += dely * fpair;
  fza += delz * fpair;
}

// Two listed hits at once (independent dependency chains); `two` == false evaluates only l0.
// Contributions are added l0 first, then l1.
template <bool SINGLE_TYPE>
__device__ __forceinline__ void pair_brick_eval2(const BrickSmem &S, const PairTab &pt, int l0, int l1, bool two,
                                                 const BrickAtom &a, unsigned step, bool random_on, double &fxa,
                                                 double &fya, double &fza)
{
  double x0, y0, z0, u0, v0, w0, x1, y1, z1, u1, v1, w1;
  brick_load(S, l0, x0, y0, z0, u0, v0, w0);
  brick_load(S, l1, x1, y1, z1, u1, v1, w1);
  const int tg0 = __float_as_int(S.scr[l0].w), tg1 = __float_as_int(S.scr[l1].w);
  const int t0 = SINGLE_TYPE ? 1 : (int) S.type[l0], t1 = SINGLE_TYPE ? 1 : (int) S.type[l1];
  const double dx0 = a.x - x0, dy0 = a.y - y0, dz0 = a.z - z0;
  const double dx1 = a.x - x1, dy1 = a.y - y1, dz1 = a.z - z1;
  const double rsq0 = dx0 * dx0 + dy0 * dy0 + dz0 * dz0, rsq1 = dx1 * dx1 + dy1 * dy1 + dz1 * dz1;
  // the exact cutoff test (pair_dpd.cpp:116; phase 1 screens in fp32 with a margin) and :119
  const bool ok0 = rsq0 < pt.cutsq[a.type][t0] && rsq0 >= EPSILON_DPD * EPSILON_DPD;
  const bool ok1 = two && rsq1 < pt.cutsq[a.type][t1] && rsq1 >= EPSILON_DPD * EPSILON_DPD;
  const double rinv0 = brick_rsqrt(ok0 ? rsq0 : 1.0), rinv1 = brick_rsqrt(ok1 ? rsq1 : 1.0);
  const double r0 = rsq0 * rinv0, r1 = rsq1 * rinv1;
  const double dot0 = dx0 * (a.vx - u0) + dy0 * (a.vy - v0) + dz0 * (a.vz - w0);
  const double dot1 = dx1 * (a.vx - u1) + dy1 * (a.vy - v1) + dz1 * (a.vz - w1);
  const double wd0 = 1.0 - r0 * pt.cutinv[a.type][t0], wd1 = 1.0 - r1 * pt.cutinv[a.type][t1];
  double s0 = pt.a0[a.type][t0] - pt.gamma[a.type][t0] * wd0 * dot0 * rinv0;
  double s1 = pt.a0[a.type][t1] - pt.gamma[a.type][t1] * wd1 * dot1 * rinv1;
  if (random_on) {
    s0 += pt.sigma[a.type][t0] * pt.dtinvsqrt * dpd_gaussian(a.tag, tg0, step, pt.seed);
    s1 += pt.sigma[a.type][t1] * pt.dtinvsqrt * dpd_gaussian(a.tag, tg1, step, pt.seed);
  }
  const double fp0 = ok0 ? wd0 * rinv0 * s0 : 0.0, fp1 = ok1 ? wd1 * rinv1 * s1 : 0.0;
  if (ok0) {
    fxa += dx0 * fp0;
    fya += dy0 * fp0;
    fza += dz0 * fp0;
  }
  if (ok1) {
    fxa += dx1 * fp1;
    fya += dy1 * fp1;
    fza += dz1 * fp1;
  }
}

template <int C, int R, bool SINGLE_TYPE>
__global__ void __launch_bounds__(PB_TMAX, 3)
    k_pair_brick(const AtomRec *__restrict__ grec, const int4 *__restrict__ gaux, Planes p, ColRange cr, Grid g,
                 PairTab pt, const int *__restrict__ cell_start, double *__restrict__ fx, double *__restrict__ fy,
                 double *__restrict__ fz, unsigned step, int random_on, int nchunk, int zc, int capL)
{
  constexpr int RP = R + 2, NROW = (C + 2) * RP, NTR = C * R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ BrickHdr H;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;

  // ---- which brick ------------------------------------------------------------------------------
  const int nstrip = (g.ny + R - 1) / R;
  int id = blockIdx.x;
  const int chunk = id % nchunk;
  id /= nchunk;
  const int cy = (id % nstrip) * R;
  int ct = id / nstrip, lc = -1, cend = -1;
  for (int k = 0; k < cr.n; k++) {
    const int nt = (cr.b[k] - cr.a[k] + C) / C;
    if (lc < 0 && ct >= 0 && ct < nt) {
      lc = cr.a[k] + ct * C;
      cend = cr.b[k];
    }
    ct -= nt;
  }
  const int z0 = chunk * zc, z1 = min(z0 + zc, g.nz);  // tile cells [z0, z1)
  if (lc < 0 || z0 >= z1) return;
  const int ncolsv = min(C, cend - lc + 1), nrowsv = min(R, g.ny - cy);  // columns / rows of the brick that exist
  const int nw = z1 - z0 + 2, cbs = nw + 1;                              // window cells z0-1 .. z1

  BrickSmem S;
  S.rec = reinterpret_cast<AtomRec *>(smem_raw);
  S.scr = reinterpret_cast<float4 *>(S.rec + capL);
  S.cb = reinterpret_cast<int *>(S.scr + capL + 4);
  S.hits = reinterpret_cast<unsigned short *>(S.cb + NROW * cbs + (NROW * cbs & 1) + 2);
  S.type = reinterpret_cast<unsigned char *>(S.hits + PB_HCAP * T);

  // ---- round trip 1: geometry and slot runs of the window rows, slot ranges of the tile rows ---------
  if (tid == 0) {
    mbar_init(&mbar, 1);
    H.issued = 0;
  }
  if (tid < NROW) {
    const int wx = tid / RP, wy = tid - wx * RP;
    const int bx = lc - 1 + wx;
    int by = cy - 1 + wy;
    double sh = 0.0;
    if (by < 0) { by += g.ny; sh = -g.prd[1]; }
    else if (by >= g.ny) { by -= g.ny; sh = g.prd[1]; }
    const bool needed = wx <= ncolsv + 1 && wy <= nrowsv + 1 && bx >= 0 && bx < g.nx;
    const int rb = needed ? (bx * g.ny + by) * g.nz : -1;
    int mg = 0, mn = 0, lg = 0, ln = 0, hg = 0, hn = 0;
    if (rb >= 0) {
      const int zlo = max(z0 - 1, 0), zhi = min(z1, g.nz - 1);
      mg = cell_start[rb + zlo];
      mn = cell_start[rb + zhi + 1] - mg;
      if (z0 == 0) {
        lg = cell_start[rb + g.nz - 1];
        ln = cell_start[rb + g.nz] - lg;
      }
      if (z1 == g.nz) {
        hg = cell_start[rb];
        hn = cell_start[rb + 1] - hg;
      }
    }
    H.row_base[tid] = rb;
    H.yshift[tid] = sh;
    H.yface[tid] = (by == 0 || by == g.ny - 1) ? 1 : 0;
    H.main_g[tid] = mg; H.main_n[tid] = mn;
    H.wlo_g[tid] = lg; H.wlo_n[tid] = ln;
    H.whi_g[tid] = hg; H.whi_n[tid] = hn;
  } else if (tid < NROW + NTR) {
    const int tr = tid - NROW, tc = tr / R, ty = tr - tc * R;
    int b = 0, n = 0;
    if (tc < ncolsv && ty < nrowsv) {
      const int rb = ((lc + tc) * g.ny + cy + ty) * g.nz;
      b = cell_start[rb + z0];
      n = cell_start[rb + z1] - b;
    }
    H.tb[tr] = b;
    H.tcum[tr + 1] = n;  // turned into a running count below
  } else if (tid == NROW + NTR) {
    H.org[0] = g.lo[0] + (lc + g.cxoff) * (g.prd[0] / g.gnx);
    H.org[1] = g.lo[1] + cy * (g.prd[1] / g.ny);
    H.org[2] = g.lo[2] + z0 * (g.prd[2] / g.nz);
  }
  __syncthreads();

  // ---- warp 0: landing offsets of the rows, then the bulk copies ---------------------------------
  if (warp == 0) {
    const int r0 = 2 * lane, r1 = 2 * lane + 1;  // two rows per lane (NROW <= 36 < 64)
    const int n0 = r0 < NROW ? H.main_n[r0] + H.wlo_n[r0] + H.whi_n[r0] : 0;
    const int n1 = r1 < NROW ? H.main_n[r1] + H.wlo_n[r1] + H.whi_n[r1] : 0;
    int inc = n0 + n1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += v;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    const int off0 = inc - n0 - n1, off1 = off0 + n0;
    if (r0 < NROW) H.row_l[r0] = off0;
    if (r1 < NROW) H.row_l[r1] = off1;
    if (lane == 0) {
      H.row_l[NROW] = total;
      int run = 0;
      H.tcum[0] = 0;
      for (int k = 1; k <= NTR; k++) {
        run += H.tcum[k];
        H.tcum[k] = run;
      }
    }
    const bool ok = total <= capL;
    if (ok) {
      if (lane == 0) {
        H.issued = 1;
        mbar_arrive_expect_tx(&mbar, (unsigned) total * 64u);
      }
    }
  }
  __syncthreads();
  const bool staged = H.issued != 0;
  const int nland = H.row_l[NROW];

  if (staged) {
    // one copy per thread: (row, piece, record kind); every warp feeds the copy engine
    for (int q = tid; q < NROW * 6; q += T) {
      const int r = q / 6, k = q - r * 6, pc = k >> 1;
      const int ln = H.wlo_n[r], mn = H.main_n[r];
      const int n = pc == 0 ? ln : (pc == 1 ? mn : H.whi_n[r]);
      if (n > 0) {
        const int gs = pc == 0 ? H.wlo_g[r] : (pc == 1 ? H.main_g[r] : H.whi_g[r]);
        const int l = H.row_l[r] + (pc >= 1 ? ln : 0) + (pc == 2 ? mn : 0);
        if (k & 1) bulk_g2s(S.scr + l, gaux + gs, (unsigned) n * 16u, &mbar);
        else bulk_g2s(S.rec + l, grec + gs, (unsigned) n * 48u, &mbar);
      }
    }
    // ---- while the copies fly: landing index where window cell w of row r starts (second round trip
    // to the cell table; its lines were touched by the header loads) -----------------------------------
    for (int q = tid; q < NROW * cbs; q += T) {
      const int r = q / cbs, w = q - r * cbs;
      const int rb = H.row_base[r];
      int v = 0;
      if (rb >= 0) {
        const int z = z0 - 1 + w;
        if (z < 0) v = 0;
        else if (z <= g.nz) v = H.wlo_n[r] + (cell_start[rb + z] - H.main_g[r]);
        else v = H.wlo_n[r] + H.main_n[r] + H.whi_n[r];
      }
      S.cb[q] = H.row_l[r] + v;
    }
    mbar_wait(&mbar, 0);
  }
  __syncthreads();

  if (staged) {
    // ---- elementwise pass, a warp per row: wrap + image shift in place, screening copy ---------------
    const double ox = H.org[0], oy = H.org[1], oz = H.org[2];
    const int nwarp = T >> 5;
    for (int r = warp; r < NROW; r += nwarp) {
      const int rl = H.row_l[r], re = H.row_l[r + 1];
      const int lo_end = rl + H.wlo_n[r], hi_beg = re - H.whi_n[r];
      const double ys = H.yshift[r];
      // before Domain::pbc only atoms of the face layers can still carry unwrapped coordinates: the rows
      // of a y face layer, and in every row the cells z = 0 and z = nz-1 (the ends of the row's run when
      // the window reaches them, and the two wrapped cells)
      const bool yface = H.yface[r] != 0;
      const int zf_lo = (z0 <= 1) ? S.cb[r * cbs + (z0 == 0 ? 2 : 1)] : rl;        // atoms below: cells <= 0 (and the wrapped low cell)
      const int zf_hi = (z1 >= g.nz - 1) ? S.cb[r * cbs + (z1 == g.nz ? nw - 2 : nw - 1)] : re;  // atoms from here: cell nz-1 (and the wrapped high cell)
      for (int l = rl + lane; l < re; l += 32) {
        double2 *rr = reinterpret_cast<double2 *>(S.rec + l);
        const double2 xy = rr[0];
        const double z = rr[1].x;
        // Domain::pbc (k_pbc's arithmetic; a no-op for wrapped coordinates), then the image next to the tile
        const double y2 = (yface ? wrap1(xy.y, g.lo[1], g.hi[1], g.prd[1]) : xy.y) + ys;
        const bool zface = l < zf_lo || l >= zf_hi;
        const double z2 = (zface ? wrap1(z, g.lo[2], g.hi[2], g.prd[2]) : z) + (l < lo_end ? -g.prd[2] : (l >= hi_beg ? g.prd[2] : 0.0));
        if (y2 != xy.y) rr[0].y = y2;
        if (z2 != z) rr[1].x = z2;
        const float4 aux = S.scr[l];  // {tag, type, 0, 0} as landed
        if (!SINGLE_TYPE) S.type[l] = (unsigned char) __float_as_int(aux.y);
        S.scr[l] = make_float4((float) (xy.x - ox), (float) (y2 - oy), (float) (z2 - oz), aux.x);
      }
    }
    if (tid < 4) S.scr[nland + tid] = make_float4(1.0e30f, 1.0e30f, 1.0e30f, 0.0f);  // read by the last screen group, masked
  }
  __syncthreads();

  const int ntile = H.tcum[NTR];
  for (int k0 = 0; k0 < ntile; k0 += T) {
    const int k = k0 + tid;
    const bool mine = k < ntile;
    const unsigned live = __ballot_sync(0xffffffffu, mine);
    if (!mine) continue;
    int tr = 0;
#pragma unroll
    for (int q = 1; q < NTR; q++) tr += (k >= H.tcum[q]) ? 1 : 0;
    const int kin = k - H.tcum[tr];
    const int i = H.tb[tr] + kin;
    if (!staged) {  // unusually dense neighbourhood: plain global path (planes; wraps on the fly)
      PairAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      Extras ex0 = {0, 0, 0, 0};
      pair_atom_global<SINGLE_TYPE, false, true, int>(p, g, pt, ex0, cell_start, nullptr, nullptr, i, live, step,
                                                      random_on != 0, reinterpret_cast<int *>(smem_raw) + tid, T, acc);
      fx[i] = acc.fx;
      fy[i] = acc.fy;
      fz[i] = acc.fz;
      continue;
    }
    const int tc = tr / R, ty = tr - tc * R;
    const int *cbo = S.cb + ((tc + 1) * RP + ty + 1) * cbs;  // own row
    const int ls = cbo[1] + kin;                            // my landing index
    int w;                                                  // my window cell: cbo[w] <= ls < cbo[w + 1]
    {
      int lo = 1, hi = nw - 2;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (cbo[mid] <= ls) lo = mid;
        else hi = mid - 1;
      }
      w = lo;
    }
    BrickAtom a;
    brick_load(S, ls, a.x, a.y, a.z, a.vx, a.vy, a.vz);
    const float4 me = S.scr[ls];
    a.tag = __float_as_int(me.w);
    a.type = SINGLE_TYPE ? 1 : (int) S.type[ls];
    double fxa = 0.0, fya = 0.0, fza = 0.0;
    const unsigned hbase = smem_u32(S.hits + tid);
    unsigned hp = hbase;  // next free slot of my hit list: hits[k*T + tid]
    const int *cb0 = S.cb + (tc * RP + ty) * cbs + (w - 1);
    // ---- phase 1: fp32 screen of the nine row segments.  The sign of rsq - cut is funnel-shifted into a
    // word (first candidate highest), the word is bit-reversed so that candidate j sits at bit j, and its
    // set bits are appended to the hit list.  The margin makes the screen a superset of the exact test.
    const float cut32 = (float) pt.cutmax * (float) pt.cutmax * 1.0001f;
    int nh = 0;
    bool fast = true;
#pragma unroll
    for (int dx = 0; dx < 3; dx++) {
#pragma unroll
      for (int dy = 0; dy < 3; dy++) {
        const int *cbr = cb0 + (dx * RP + dy) * cbs;
        const int qb = cbr[0], n = cbr[3] - qb;  // candidates of this segment
        if (n > 32) fast = false;
        const int qe = qb + min(n, 32);
        unsigned word = 0;
        int cnt = 0;
        for (int q0 = qb; q0 < qe; q0 += 4, cnt += 4) {
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const float4 c = S.scr[q0 + u];
            const float ddx = me.x - c.x, ddy = me.y - c.y, ddz = me.z - c.z;
            const float rmc = fmaf(ddz, ddz, fmaf(ddy, ddy, fmaf(ddx, ddx, -cut32)));  // rsq - cut: sign bit = hit
            word = __funnelshift_l(__float_as_uint(rmc), word, 1);
          }
        }
        word = cnt > 0 ? (__brev(word) >> (32 - cnt)) : 0u;
        if (n < 32) word &= (1u << n) - 1u;  // a group may run up to 3 candidates past the segment
        if (dx == 1 && dy == 1) {            // not myself
          const int so = ls - qb;
          if (so < 32) word &= ~(1u << so);
        }
        nh += __popc(word);
        if (nh <= PB_HCAP) {
          while (word) {
            const int kbit = __ffs(word) - 1;
            word &= word - 1u;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(hp), "r"(qb + kbit) : "memory");
            hp += 2 * T;
          }
        }
      }
    }
    fast = fast && nh <= PB_HCAP;
    if (fast) {
      // ---- phase 2: the listed hits in fp64, two at a time --------------------------------------------
      for (unsigned hq = hbase; hq < hp; hq += 4 * T) {
        unsigned l0, l1;
        const bool two = hq + 2 * T < hp;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(l0) : "r"(hq) : "memory");
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(l1) : "r"(two ? hq + 2 * T : hq) : "memory");
        pair_brick_eval2<SINGLE_TYPE>(S, pt, (int) l0, (int) l1, two, a, step, random_on != 0, fxa, fya, fza);
      }
    } else {
      // an unusually crowded neighbourhood: test and evaluate in place, same order
#pragma unroll
      for (int dx = 0; dx < 3; dx++)
#pragma unroll
        for (int dy = 0; dy < 3; dy++) {
          const int *cbr = cb0 + (dx * RP + dy) * cbs;
          for (int q = cbr[0]; q < cbr[3]; q++)
            if (q != ls) pair_brick_eval<SINGLE_TYPE>(S, pt, q, a, step, random_on != 0, fxa, fya, fza);
        }
    }
    fx[i] = fxa;
    fy[i] = fya;
    fz[i] = fza;
  }
}
