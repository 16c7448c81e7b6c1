// Row-tiled PairDPD kernel (included by pair.cu): the thread-staged tile of round 1.  Since round 2 the
// brick kernel (pair_brick.cuh) is the production path; this one takes the columns whose stencil holds
// atoms appended after the sort (it folds their per-cell lists into the window) and everything when the
// sort's atom records are stale.
//
// A tile is the atoms of R (1, 2 or 4) y-adjacent cell rows of one column over a run of z-cells: R
// contiguous slot ranges, because z is the fastest cell index.  Its stencil is the 3 x (R+2) rows
// around it over the same z window plus one cell on either side.  The CTA copies x, v and tag of
// that window into shared memory ONCE with 8-byte `cp.async` copies (no register staging, every
// copy in flight at once), in window-cell-major order (w, ry, rx, atom).  In that order the nine
// stencil cells an atom needs from one window cell are contiguous, and for R = 1 all 27 are, so
//   phase 1  each thread screens its ~81 candidates in flat loops over an fp32 copy of the positions
//            (one LDS.128 per candidate, conservative margin) and keeps the hits as register bitmasks,
//   phase 2  it evaluates the ~12.6 hits densely in fp64 (exact cutoff test first, then rsqrt, drag,
//            counter RNG), positions, velocities
//            and tags all from shared memory -- the kernel's only global traffic is the staging
//            copy (served by L2 after the first touch) and the force store.
// Periodic images are materialised while staging: a row reached across the y face, or a
// window cell reached across the z face, is stored shifted by the box length, exactly as
// the reference's ghost atoms are (comm_brick.cpp:735-760 adds the period once, pair_dpd.cpp:101-104
// then subtracts plain coordinates).  The inner loops therefore carry no minimum-image logic.
// Atoms appended after the sort (insertions, slab arrivals and ghosts) hang off per-cell lists; the
// staging copy folds them into their cells, so they are ordinary candidates too.
//
// Accumulation order (fixed by the layout, independent of launch geometry): candidate order
// (w, ry, rx, slot).  An atom with more than 64 candidates in one window cell takes a
// test-and-evaluate loop in the same order.
#pragma once

#define PR_HCAP 32        // hit slots per thread (mean 12.6 hits at rho = 3); more -> test-and-evaluate path
#define PR_MAXZC 62       // z-cells per tile at most
#define PR_TMAX 288

struct RowHdr {
  int row_base[18];  // first cell index of stencil row r = ry*3 + rx (R <= 4), -1 if that column does not exist
  double yshift[18];
  int tb[4], te[4];  // tile atoms of tile row ty are slots [tb, te)
  int nstaged;
  double org[3];     // low corner of the tile (origin of the fp32 screening copy)
};

struct RowSmem {
  double *pos;          // [cap][3]
  double *vel;          // [cap][3]
  int *tag;             // [cap]
  int *cbeg;            // [NR*nw + 1] staged offset of cell (w, r), index w*NR + r
  int *gbeg;            // [NR*nw]     first global slot of that cell
  float4 *scr;          // [cap] fp32 copy of pos relative to the tile origin (screening only)
  unsigned char *hits;  // [PR_HCAP][T] hit list of each thread: byte offsets from its first candidate
  unsigned char *type;  // [cap] (multi-type systems only)
};

// dynamic shared memory of one CTA
static inline size_t pair_row_smem_bytes(int R, int cap, int nw, int T, bool single_type)
{
  const int NR = 3 * (R + 2);
  size_t b = (size_t) cap * (16 + 6 * 8 + 4) + (size_t) 4 * (2 * NR * nw + 2) + (size_t) PR_HCAP * T;
  if (!single_type) b += (size_t) cap;
  const size_t fallback = (size_t) T * PAIR_HCAP * 4;  // int hit lists of pair_atom_global
  if (b < fallback) b = fallback;
  return (b + 15) / 16 * 16;
}

__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src_global)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)),
               "l"(src_global)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src_global)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)),
               "l"(src_global)
               : "memory");
}

struct RowAtom {
  double x, y, z, vx, vy, vz;
  int tag, type;
};

// pair_dpd.cpp:117-146 for candidate q of the staged window
template <bool SINGLE_TYPE>
__device__ __forceinline__ void pair_row_eval(const RowSmem &S, const PairTab &pt, int q, const RowAtom &a,
                                              unsigned step, bool random_on, double &fxa, double &fya,
                                              double &fza)
{
  const double *pj = S.pos + 3 * q, *vj = S.vel + 3 * q;
  const double delx = a.x - pj[0], dely = a.y - pj[1], delz = a.z - pj[2];
  const double rsq = delx * delx + dely * dely + delz * delz;
  const int tj = SINGLE_TYPE ? 1 : (int) S.type[q];
  // the exact cutoff test (pair_dpd.cpp:116; phase 1 screens in fp32 with a margin) and :119
  if (rsq >= pt.cutsq[a.type][tj] || rsq < EPSILON_DPD * EPSILON_DPD) return;
  const double rinv = rsqrt(rsq);
  const double r = rsq * rinv;
  const double delvx = a.vx - vj[0], delvy = a.vy - vj[1], delvz = a.vz - vj[2];
  const double dot = delx * delvx + dely * delvy + delz * delvz;
  const double wd = 1.0 - r * pt.cutinv[a.type][tj];
  // fpair = (a0*wd - gamma*wd^2*dot/r + sigma*wd*xi/sqrt(dt)) / r, factored as wd/r * (...)
  double t = pt.a0[a.type][tj] - pt.gamma[a.type][tj] * wd * dot * rinv;
  if (random_on) t += pt.sigma[a.type][tj] * pt.dtinvsqrt * dpd_gaussian(a.tag, S.tag[q], step, pt.seed);
  const double fpair = wd * rinv * t;
  fxa += delx * fpair;
  fya += dely * fpair;
  fza += delz * fpair;
}

// Two candidates at once (independent dependency chains: twice the loads in flight and two fp64
// chains interleaved); `two` == false evaluates only q0.  Contributions are added q0 first, then q1.
template <bool SINGLE_TYPE>
__device__ __forceinline__ void pair_row_eval2(const RowSmem &S, const PairTab &pt, int q0, int q1, bool two,
                                               const RowAtom &a, unsigned step, bool random_on, double &fxa,
                                               double &fya, double &fza)
{
  const double *p0 = S.pos + 3 * q0, *v0 = S.vel + 3 * q0, *p1 = S.pos + 3 * q1, *v1 = S.vel + 3 * q1;
  const double dx0 = a.x - p0[0], dy0 = a.y - p0[1], dz0 = a.z - p0[2];
  const double dx1 = a.x - p1[0], dy1 = a.y - p1[1], dz1 = a.z - p1[2];
  const double du0 = a.vx - v0[0], dv0 = a.vy - v0[1], dw0 = a.vz - v0[2];
  const double du1 = a.vx - v1[0], dv1 = a.vy - v1[1], dw1 = a.vz - v1[2];
  const int tg0 = S.tag[q0], tg1 = S.tag[q1];
  const int t0 = SINGLE_TYPE ? 1 : (int) S.type[q0], t1 = SINGLE_TYPE ? 1 : (int) S.type[q1];
  const double rsq0 = dx0 * dx0 + dy0 * dy0 + dz0 * dz0, rsq1 = dx1 * dx1 + dy1 * dy1 + dz1 * dz1;
  // the exact cutoff test (pair_dpd.cpp:116; phase 1 screens in fp32 with a margin) and :119
  const bool ok0 = rsq0 < pt.cutsq[a.type][t0] && rsq0 >= EPSILON_DPD * EPSILON_DPD;
  const bool ok1 = two && rsq1 < pt.cutsq[a.type][t1] && rsq1 >= EPSILON_DPD * EPSILON_DPD;
  const double rinv0 = rsqrt(ok0 ? rsq0 : 1.0), rinv1 = rsqrt(ok1 ? rsq1 : 1.0);
  const double r0 = rsq0 * rinv0, r1 = rsq1 * rinv1;
  const double dot0 = dx0 * du0 + dy0 * dv0 + dz0 * dw0, dot1 = dx1 * du1 + dy1 * dv1 + dz1 * dw1;
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

template <bool SINGLE_TYPE>
__device__ __forceinline__ bool pair_row_hit(const RowSmem &S, const PairTab &pt, int q, const RowAtom &a)
{
  const double *pj = S.pos + 3 * q;
  const double dx = a.x - pj[0], dy = a.y - pj[1], dz = a.z - pj[2];
  const double rsq = dx * dx + dy * dy + dz * dz;
  const int tj = SINGLE_TYPE ? 1 : (int) S.type[q];
  return rsq < pt.cutsq[a.type][tj];
}

template <int R, bool SINGLE_TYPE, bool PREWRAP>
__global__ void __launch_bounds__(PR_TMAX, 1)
    k_pair_row(Planes p, ColRange cr, Extras ex, Grid g, PairTab pt, const int *__restrict__ cell_start,
               const int *__restrict__ extra_head, const int *__restrict__ extra_next,
               double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz, unsigned step,
               int random_on, int nchunk, int zc, int cap)
{
  constexpr int NR = 3 * (R + 2);  // stencil rows per window cell
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ RowHdr H;
  const int tid = threadIdx.x, T = blockDim.x;

  // ---- which tile ---------------------------------------------------------------------------
  const int nstrip = (g.ny + R - 1) / R;
  const int chunk = blockIdx.x % nchunk;
  const int rowid = blockIdx.x / nchunk;
  // odd ny, R = 2: the last strip is shifted back onto rows ny-2, ny-1 and owns only the last one, so
  // every tile sees all 3 x (R+2) stencil rows (the tail over-read of phase 1 relies on that)
  int cy = (rowid % nstrip) * R, ty_first = 0;
  if (cy + R > g.ny) {
    ty_first = cy + R - g.ny;
    cy = g.ny - R;
  }
  int col = rowid / nstrip, lc = -1;  // col-th column of the set
  for (int k = 0; k < cr.n; k++) {
    const int len = cr.b[k] - cr.a[k] + 1;
    if (lc < 0 && col >= 0 && col < len) lc = cr.a[k] + col;
    col -= len;
  }
  const int z0 = chunk * zc, z1 = min(z0 + zc, g.nz);  // tile cells [z0, z1)
  if (lc < 0 || z0 >= z1) return;
  const int nw = z1 - z0 + 2;                          // window cells z0-1 .. z1
  const int ncw = NR * nw;

  RowSmem S;
  S.scr = reinterpret_cast<float4 *>(smem_raw);
  S.pos = reinterpret_cast<double *>(S.scr + cap);
  S.vel = S.pos + 3 * cap;
  S.tag = reinterpret_cast<int *>(S.vel + 3 * cap);
  S.cbeg = S.tag + cap;
  S.gbeg = S.cbeg + ncw + 1;
  S.hits = reinterpret_cast<unsigned char *>(S.gbeg + ncw);
  S.type = S.hits + PR_HCAP * T;

  if (tid < NR) {
    const int ry = tid / 3, rx = tid - 3 * ry;
    const int bx = lc + rx - 1;
    int by = cy + ry - 1;
    double sh = 0.0;
    if (by < 0) { by += g.ny; sh = -g.prd[1]; }
    else if (by >= g.ny) { by -= g.ny; sh = g.prd[1]; }
    H.row_base[tid] = (bx < 0 || bx >= g.nx) ? -1 : (bx * g.ny + by) * g.nz;
    H.yshift[tid] = sh;
  }
  if (tid == 31) {
    H.org[0] = g.lo[0] + (lc + g.cxoff) * (g.prd[0] / g.gnx);
    H.org[1] = g.lo[1] + cy * (g.prd[1] / g.ny);
    H.org[2] = g.lo[2] + z0 * (g.prd[2] / g.nz);
  }
  __syncthreads();

  // ---- cell populations of the window (sorted range + appended atoms), then their scan ---------
  const bool have_extras = (ex.narr + ex.nghost + ex.nins) > 0;
  for (int q = tid; q < ncw; q += T) {
    const int r = q / nw, w = q - r * nw;  // threads walk z: coalesced reads (ncw <= 1152: a few trips)
    int cnt = 0, gb = 0;
    if (H.row_base[r] >= 0) {
      int zu = z0 - 1 + w;
      zu = zu < 0 ? zu + g.nz : (zu >= g.nz ? zu - g.nz : zu);
      const int cc = H.row_base[r] + zu;
      gb = cell_start[cc];
      cnt = cell_start[cc + 1] - gb;
      if (have_extras)
        for (int e = extra_head[cc]; e >= 0; e = extra_next[e]) cnt++;
    }
    S.cbeg[w * NR + r] = cnt;
    S.gbeg[w * NR + r] = gb;
  }
  if (tid >= T - 2 * R) {  // slot ranges of the tile rows (same round trip as the counts)
    const int k = T - 1 - tid, ty = k >> 1;
    int v = 0;
    if (ty >= ty_first) v = cell_start[(lc * g.ny + cy + ty) * g.nz + ((k & 1) ? z1 : z0)];
    if (k & 1) H.te[ty] = v;
    else H.tb[ty] = v;
  }
  __syncthreads();
  if (tid < 32) {  // exclusive scan of the counts by one warp, a contiguous run per lane
    const int per = (ncw + 31) / 32, b = tid * per, e = min(b + per, ncw);
    int s = 0;
    for (int k = b; k < e; k++) s += S.cbeg[k];
    int inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, d);
      if (tid >= d) inc += v;
    }
    int run = inc - s;
    for (int k = b; k < e; k++) {
      const int v = S.cbeg[k];
      S.cbeg[k] = run;
      run += v;
    }
    if (tid == 31) {
      S.cbeg[ncw] = inc;
      H.nstaged = inc;
    }
  }
  __syncthreads();
  const bool staged = H.nstaged + 4 <= cap;  // 4 slack entries: phase 1 reads up to 3 past a segment

  // ---- staging copy: one thread per staged atom (balanced, consecutive threads copy consecutive
  // atoms), every copy asynchronous.  The cell of each staged index is first written to a table
  // (it aliases the tag array: entry d is read by the thread that then issues the copy into tag[d]).
  if (staged) {
    for (int ci = tid; ci < ncw; ci += T)
      for (int d = S.cbeg[ci]; d < S.cbeg[ci + 1]; d++) S.tag[d] = ci;
    __syncthreads();
    // Velocities and tags are copied asynchronously as they are.  Positions go through registers: the
    // thread applies the periodic image shift of the atom's window cell (and, before Domain::pbc has
    // run, the wrap itself), stores the fp64 position and the fp32 screening copy relative to the
    // tile's low corner (|coordinate| < ~70 cell widths, so the absolute error is < 1e-5 and the
    // relative margin on cutsq in phase 1 covers it many times over).  Two atoms per trip keep six
    // position loads in flight per thread.
    const double ox = H.org[0], oy = H.org[1], oz = H.org[2];
    for (int d0 = tid; d0 < H.nstaged; d0 += 2 * T) {
      int dd[2], jj[2];
      double ys[2], zs[2];
      bool on[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int d = d0 + u * T;
        on[u] = d < H.nstaged;
        dd[u] = on[u] ? d : d0;
        const int ci = S.tag[dd[u]];
        const int k = dd[u] - S.cbeg[ci];
        int j = S.gbeg[ci] + k;
        const int w = ci / NR, r = ci - w * NR;
        const int zraw = z0 - 1 + w;
        if (have_extras) {  // atoms appended after the sort sit behind the cell's sorted atoms
          const int zu = zraw < 0 ? zraw + g.nz : (zraw >= g.nz ? zraw - g.nz : zraw);
          const int cc = H.row_base[r] + zu;
          const int ns = cell_start[cc + 1] - S.gbeg[ci];
          if (k >= ns) {
            int e = extra_head[cc];
            for (int m = ns; m < k; m++) e = extra_next[e];
            j = ex.base + e;
          }
        }
        jj[u] = j;
        ys[u] = H.yshift[r];
        zs[u] = zraw < 0 ? -g.prd[2] : (zraw >= g.nz ? g.prd[2] : 0.0);
      }
      double xx[2], yy[2], zz[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        xx[u] = p.x[jj[u]];
        yy[u] = p.y[jj[u]];
        zz[u] = p.z[jj[u]];
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (!on[u]) continue;
        const int d = dd[u], j = jj[u];
        cp_async8(S.vel + 3 * d, p.vx + j);
        cp_async8(S.vel + 3 * d + 1, p.vy + j);
        cp_async8(S.vel + 3 * d + 2, p.vz + j);
        cp_async4(S.tag + d, p.tag + j);  // entry d of the cell table was consumed above, by this thread
        if (!SINGLE_TYPE) S.type[d] = (unsigned char) p.type[j];
        const double y = (PREWRAP ? wrap1(yy[u], g.lo[1], g.hi[1], g.prd[1]) : yy[u]) + ys[u];
        const double z = (PREWRAP ? wrap1(zz[u], g.lo[2], g.hi[2], g.prd[2]) : zz[u]) + zs[u];
        S.pos[3 * d] = xx[u];
        S.pos[3 * d + 1] = y;
        S.pos[3 * d + 2] = z;
        S.scr[d] = make_float4((float) (xx[u] - ox), (float) (y - oy), (float) (z - oz), 0.0f);
      }
    }
    if (tid < 4) {  // slack entries read by the last screen group (masked afterwards): far away
      const int d = H.nstaged + tid;
      S.scr[d] = make_float4(1.0e30f, 1.0e30f, 1.0e30f, 0.0f);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  int ncum[R + 1];  // tile atoms of rows 0..ty-1
  ncum[0] = 0;
#pragma unroll
  for (int ty = 0; ty < R; ty++) ncum[ty + 1] = ncum[ty] + (H.te[ty] - H.tb[ty]);
  const int ntile = ncum[R];
  for (int k0 = 0; k0 < ntile; k0 += T) {
    const int k = k0 + tid;
    const bool mine = k < ntile;
    const unsigned live = __ballot_sync(0xffffffffu, mine);
    if (!mine) continue;
    int ty = 0;
#pragma unroll
    for (int q = 1; q < R; q++) ty += (k >= ncum[q]) ? 1 : 0;
    int i = H.tb[0] + k;
#pragma unroll
    for (int q = 1; q < R; q++)
      if (ty == q) i = H.tb[q] + (k - ncum[q]);
    if (!staged) {  // unusually dense neighbourhood: plain global path
      PairAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      pair_atom_global<SINGLE_TYPE, false, PREWRAP, int>(p, g, pt, ex, cell_start, extra_head, extra_next, i,
                                                         live, step, random_on != 0,
                                                         reinterpret_cast<int *>(smem_raw) + tid, T, acc);
      fx[i] = acc.fx;
      fy[i] = acc.fy;
      fz[i] = acc.fz;
      continue;
    }
    // window cell of atom i: the own-row cell whose slot range holds i (exact: no coordinate
    // arithmetic that could disagree with the sort key)
    const int rc = (ty + 1) * 3 + 1;  // own row, own column
    int w;
    {
      int lo = 1, hi = nw - 2;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (S.gbeg[mid * NR + rc] <= i) lo = mid;
        else hi = mid - 1;
      }
      w = lo;
    }
    const int qs = S.cbeg[w * NR + rc] + (i - S.gbeg[w * NR + rc]);
    RowAtom a;
    a.x = S.pos[3 * qs]; a.y = S.pos[3 * qs + 1]; a.z = S.pos[3 * qs + 2];
    a.vx = S.vel[3 * qs]; a.vy = S.vel[3 * qs + 1]; a.vz = S.vel[3 * qs + 2];
    a.tag = S.tag[qs];
    a.type = SINGLE_TYPE ? 1 : (int) S.type[qs];
    // candidate segments: window cells w-1, w, w+1; in each the rows ty..ty+2 (nine cells) are contiguous
    int sb[3], se[3];
#pragma unroll
    for (int s = 0; s < 3; s++) {
      sb[s] = S.cbeg[(w - 1 + s) * NR + ty * 3];
      se[s] = S.cbeg[(w - 1 + s) * NR + ty * 3 + 9];
    }
    double fxa = 0.0, fya = 0.0, fza = 0.0;
    const unsigned hbase = (unsigned) __cvta_generic_to_shared(S.hits + tid);
    unsigned hp = hbase;  // next free slot of my hit list: hits[k*T + tid]
    bool fast = (se[0] - sb[0] <= 64) && (se[1] - sb[1] <= 64) && (se[2] - sb[2] <= 64) && (se[2] - sb[0] <= 256);
    if (fast) {
      // ---- phase 1: fp32 screen, one LDS.128 per candidate.  The sign of rsq - cut is funnel-shifted
      // into a word (first candidate highest), the word is bit-reversed so that candidate j sits at
      // bit j, and its set bits are appended to the hit list as byte offsets from sb[0].  The margin
      // on the cutoff makes the screen a superset of the exact fp64 test of phase 2.
      const float4 me = S.scr[qs];
      const float cut32 = (float) pt.cutmax * (float) pt.cutmax * 1.0001f;
      int nh = 0;
#pragma unroll
      for (int s = 0; s < 3; s++) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
          unsigned word = 0;
          const int qb = sb[s] + 32 * half, qe = min(se[s], qb + 32);
          int cnt = 0;
          for (int q0 = qb; q0 < qe; q0 += 4, cnt += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const float4 c = S.scr[q0 + u];
              const float dx = me.x - c.x, dy = me.y - c.y, dz = me.z - c.z;
              const float rsq = dx * dx + dy * dy + dz * dz;
              word = __funnelshift_l(__float_as_uint(rsq - cut32), word, 1);
            }
          }
          word = cnt > 0 ? (__brev(word) >> (32 - cnt)) : 0u;
          // a group may run up to 3 candidates past the segment: drop those bits
          const int n = se[s] - qb;  // candidates of this word (<= 0: none)
          if (n < 32) word &= (n > 0) ? ((1u << n) - 1u) : 0u;
          if (s == 1) {  // not myself (my own cell is in the middle segment)
            const int so = qs - qb;
            if (so >= 0 && so < 32) word &= ~(1u << so);
          }
          nh += __popc(word);
          if (nh <= PR_HCAP) {
            const int off0 = qb - sb[0];
            while (word) {
              const int kbit = __ffs(word) - 1;
              word &= word - 1u;
              asm volatile("st.shared.u8 [%0], %1;" ::"r"(hp), "r"(off0 + kbit) : "memory");
              hp += T;
            }
          }
        }
      }
      fast = nh <= PR_HCAP;
      if (fast) {
        // ---- phase 2: evaluate the listed hits in fp64 ---------------------------------------------
        for (unsigned hq = hbase; hq < hp; hq += 2 * T) {
          unsigned off0, off1;
          const bool two = hq + T < hp;
          asm volatile("ld.shared.u8 %0, [%1];" : "=r"(off0) : "r"(hq) : "memory");
          asm volatile("ld.shared.u8 %0, [%1];" : "=r"(off1) : "r"(two ? hq + T : hq) : "memory");
          pair_row_eval2<SINGLE_TYPE>(S, pt, sb[0] + (int) off0, sb[0] + (int) off1, two, a, step, random_on != 0,
                                      fxa, fya, fza);
        }
      }
    }
    if (!fast) {
      // an unusually crowded neighbourhood: test and evaluate in place, same order
#pragma unroll
      for (int s = 0; s < 3; s++)
        for (int q = sb[s]; q < se[s]; q++)
          if (q != qs) pair_row_eval<SINGLE_TYPE>(S, pt, q, a, step, random_on != 0, fxa, fya, fza);
    }
    fx[i] = fxa;
    fy[i] = fya;
    fz[i] = fza;
  }
}
