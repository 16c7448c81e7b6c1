// C-ABI entry points of libobmd_b200.so (include/obmd_b200.h): context, parameters,
// host<->device marshalling in LAMMPS layout, the per-step hook sequence.
#include <cmath>
#include <cstddef>
#include <cstdlib>

#include <chrono>

#include "ctx.h"

char g_create_err[1024] = "no error";

int obmd_fail(obmd_ctx *c, int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(c ? c->err : g_create_err, 1024, fmt, ap);
  va_end(ap);
  return code;
}

extern "C" int obmd_abi_version(void) { return OBMD_ABI_VERSION; }

extern "C" const char *obmd_last_error(const obmd_ctx *c) { return c ? c->err : g_create_err; }

extern "C" int obmd_create(obmd_ctx **out, int device)
{
  obmd_ctx *c = nullptr;
  if (!out) return obmd_fail(c, OBMD_ERR_ARG, "obmd_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return obmd_fail(c, OBMD_ERR_CUDA, "obmd_create: no CUDA device (%s); there is no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev)
    return obmd_fail(c, OBMD_ERR_ARG, "obmd_create: device %d out of range [0,%d)", device, ndev);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return obmd_fail(c, OBMD_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return obmd_fail(c, OBMD_ERR_CUDA,
                     "obmd_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                     device, prop.major, prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess)
    return obmd_fail(c, OBMD_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  c = new obmd_ctx();
  c->device = device;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if ((e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
      (e = cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_lo)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&c->ev_sorted, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&c->ev_int, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
      (e = cudaMalloc(&c->sc_d, sizeof(StepScalars))) != cudaSuccess ||
      (e = cudaMallocHost(&c->sc_h, sizeof(StepScalars))) != cudaSuccess ||
      (e = cudaMalloc(&c->red_d, 64 * sizeof(double))) != cudaSuccess ||
      (e = cudaMallocHost(&c->red_h, 64 * sizeof(double))) != cudaSuccess) {
    obmd_fail(nullptr, OBMD_ERR_CUDA, "obmd_create: %s", cudaGetErrorString(e));
    delete c;
    return OBMD_ERR_CUDA;
  }
  cudaMemset(c->sc_d, 0, sizeof(StepScalars));
  memset(c->sc_h, 0, sizeof(StepScalars));
  memset(&c->pt, 0, sizeof(c->pt));
  memset(&c->grid, 0, sizeof(c->grid));
  memset(&c->fa, 0, sizeof(c->fa));
  memset(&c->fs, 0, sizeof(c->fs));
  memset(&c->fd, 0, sizeof(c->fd));
  memset(&c->rep, 0, sizeof(c->rep));
  memset(c->setflag, 0, sizeof(c->setflag));
  memset(&c->cur, 0, sizeof(Planes));
  memset(&c->alt, 0, sizeof(Planes));
  for (int t = 0; t <= MAXT; t++) c->mass[t] = 0.0;
  if (const char *e2 = getenv("OBMD_B200_OVERLAP")) c->overlap = atoi(e2) != 0;
  if (const char *e3 = getenv("OBMD_B200_TILE")) c->tile_kernel = atoi(e3);
  if (const char *e4 = getenv("OBMD_B200_ZC")) c->row_zc = atoi(e4);
  if (const char *e6 = getenv("OBMD_B200_ROWS")) c->row_rows = atoi(e6);
  if (const char *e7 = getenv("OBMD_B200_COLS")) c->row_cols = atoi(e7);
  if (const char *e9 = getenv("OBMD_B200_PAIR_FIRST")) c->pair_first = atoi(e9);
  if (const char *e8 = getenv("OBMD_B200_CAPSCALE")) c->row_capscale = atof(e8);
  if (const char *e5 = getenv("OBMD_B200_PT")) c->row_threads = atoi(e5);
  if (const char *e10 = getenv("OBMD_B200_USH_G")) {
    const int gsz = atoi(e10);
    if (gsz == 16 || gsz == 32) c->ush_group = gsz;
  }
  if (const char *e12 = getenv("OBMD_B200_FIX_CL")) {
    const int v = atoi(e12);
    c->fix_cluster = (v == 1 || v == 16) ? v : 8;
  }
  if (const char *e11 = getenv("OBMD_B200_USH_CTAS")) {
    const int v = atoi(e11);
    if (v >= 1 && v <= 16) c->ush_ctas_per_sm = v;
  }
  snprintf(c->err, sizeof(c->err), "no error");
  *out = c;
  return OBMD_OK;
}

static void free_planes(Planes &p)
{
  cudaFree(p.x); cudaFree(p.y); cudaFree(p.z); cudaFree(p.vx); cudaFree(p.vy); cudaFree(p.vz);
  cudaFree(p.tag); cudaFree(p.lidx); cudaFree(p.type); cudaFree(p.image);
  memset(&p, 0, sizeof(p));
}

extern "C" void obmd_destroy(obmd_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->stream2);
  free_planes(c->cur);
  free_planes(c->alt);
  cudaFree(c->fx); cudaFree(c->fy); cudaFree(c->fz); cudaFree(c->rec); cudaFree(c->aux);
  bond_free(c);
  cudaFree(c->cid); cudaFree(c->rank); cudaFree(c->perm);
  cudaFree(c->cell_count); cudaFree(c->cell_start); cudaFree(c->scan_tmp);
  cudaFree(c->extra_head); cudaFree(c->extra_next); cudaFree(c->extra_cell);
  cudaFree(c->stage_d); cudaFree(c->stage_i);
  cudaFree(c->sc_d); cudaFreeHost(c->sc_h);
  cudaFree(c->red_d); cudaFreeHost(c->red_h);
  cudaFree(c->sendL); cudaFree(c->sendR); cudaFree(c->recvL); cudaFree(c->recvR);
  cudaFree(c->del_key); cudaFree(c->del_val); cudaFree(c->tail_key); cudaFree(c->tail_val); cudaFree(c->del_src); cudaFree(c->del_mv);
  cudaFree(c->tr_d); cudaFree(c->tr_path); cudaFree(c->tr_npts);
  cudaFree(c->tr_list); cudaFree(c->tr_kfirst); cudaFree(c->tr_chg); cudaFree(c->acc_list); cudaFree(c->tr_bbox);
  cudaFree(c->spec_trial); cudaFree(c->spec_of_trial); cudaFree(c->spec_pos); cudaFree(c->tr_pairs);
  if (c->tr_h) cudaFreeHost(c->tr_h);
  cudaFree(c->partials);
  cudaFree(c->vatom);
  cudaFree(c->redo_flag); cudaFree(c->redo_list); cudaFree(c->redo_count);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
  for (auto &p : c->pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  cudaEventDestroy(c->ev_sorted); cudaEventDestroy(c->ev_int);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->stream2);
  delete c;
}

// ---- domain ------------------------------------------------------------------------

extern "C" int obmd_set_box(obmd_ctx *c, const double boxlo[3], const double boxhi[3])
{
  NEED(c && boxlo && boxhi, OBMD_ERR_ARG, "obmd_set_box: NULL argument");
  for (int d = 0; d < 3; d++)
    NEED(boxhi[d] > boxlo[d], OBMD_ERR_ARG, "obmd_set_box: box has no extent in dimension %d", d);
  for (int d = 0; d < 3; d++) {
    c->boxlo[d] = boxlo[d];
    c->boxhi[d] = boxhi[d];
  }
  c->sublo = boxlo[0];
  c->subhi = boxhi[0];
  c->have_box = true;
  c->pair_ready = false;
  return OBMD_OK;
}

extern "C" int obmd_set_masses(obmd_ctx *c, int ntypes, const double *mass)
{
  NEED(c && mass, OBMD_ERR_ARG, "obmd_set_masses: NULL argument");
  NEED(ntypes >= 1 && ntypes <= MAXT, OBMD_ERR_ARG, "obmd_set_masses: ntypes must be in [1,%d]", MAXT);
  for (int t = 1; t <= ntypes; t++)
    NEED(mass[t] > 0.0, OBMD_ERR_ARG, "Invalid mass value for type %d", t);
  c->ntypes = ntypes;
  for (int t = 1; t <= ntypes; t++) c->mass[t] = mass[t];
  c->pair_ready = false;
  return OBMD_OK;
}

extern "C" int obmd_set_timestep(obmd_ctx *c, double dt)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  NEED(dt > 0.0, OBMD_ERR_ARG, "obmd_set_timestep: dt must be > 0");
  c->dt = dt;
  return OBMD_OK;
}

// ---- pair_style dpd -------------------------------------------------------------------

extern "C" int obmd_pair_settings(obmd_ctx *c, double temperature, double cut_global, int seed)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  NEED(seed > 0, OBMD_ERR_ARG, "Illegal pair_style command");  // pair_dpd.cpp:205
  NEED(cut_global > 0.0, OBMD_ERR_ARG, "Illegal pair_style command");
  c->temperature = temperature;
  c->cut_global = cut_global;
  c->pair_seed = seed;
  for (int i = 1; i <= MAXT; i++)  // :211-216 reset cutoffs that have been explicitly set
    for (int j = i; j <= MAXT; j++)
      if (c->setflag[i][j]) c->pt.cut[i][j] = cut_global;
  c->have_pair = true;
  c->pair_ready = false;
  return OBMD_OK;
}

extern "C" int obmd_pair_coeff(obmd_ctx *c, int itype, int jtype, double a0, double gamma, double cut)
{
  NEED(c && c->have_pair, OBMD_ERR_STATE, "obmd_pair_coeff: call obmd_pair_settings first");
  NEED(c->ntypes > 0, OBMD_ERR_STATE, "obmd_pair_coeff: call obmd_set_masses first");
  NEED(itype >= 1 && jtype >= 1 && itype <= c->ntypes && jtype <= c->ntypes, OBMD_ERR_ARG,
       "Incorrect args for pair coefficients");
  int i = itype < jtype ? itype : jtype, j = itype < jtype ? jtype : itype;
  c->pt.a0[i][j] = a0;
  c->pt.gamma[i][j] = gamma;
  c->pt.cut[i][j] = cut > 0.0 ? cut : c->cut_global;
  c->setflag[i][j] = true;
  c->pair_ready = false;
  return OBMD_OK;
}

int setup_grid(obmd_ctx *c);

__global__ void k_fill_int(int *p, int n, int v)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

extern "C" int obmd_pair_init(obmd_ctx *c, double boltz)
{
  NEED(c && c->have_pair && c->have_box && c->ntypes > 0, OBMD_ERR_STATE,
       "obmd_pair_init: box, masses and pair settings must be set first");
  PairTab &pt = c->pt;
  pt.ntypes = c->ntypes;
  pt.cutmax = 0.0;
  c->boltz = boltz;
  for (int i = 1; i <= c->ntypes; i++)
    for (int j = i; j <= c->ntypes; j++) {
      NEED(c->setflag[i][j], OBMD_ERR_STATE, "All pair coeffs are not set");  // pair_dpd.cpp:281
      pt.sigma[i][j] = sqrt(2.0 * boltz * c->temperature * pt.gamma[i][j]);
      pt.cut[j][i] = pt.cut[i][j];
      pt.a0[j][i] = pt.a0[i][j];
      pt.gamma[j][i] = pt.gamma[i][j];
      pt.sigma[j][i] = pt.sigma[i][j];
      pt.cutsq[i][j] = pt.cutsq[j][i] = pt.cut[i][j] * pt.cut[i][j];
      pt.cutinv[i][j] = pt.cutinv[j][i] = 1.0 / pt.cut[i][j];
      if (pt.cut[i][j] > pt.cutmax) pt.cutmax = pt.cut[i][j];
    }
  for (int t = 0; t <= MAXT; t++) pt.mass[t] = c->mass[t];
  pt.seed = (unsigned) c->pair_seed;
  pt.dtinvsqrt = c->dt > 0.0 ? 1.0 / sqrt(c->dt) : 0.0;
  c->pair_ready = true;
  return setup_grid(c);
}

// Owned cell columns [c0, c1) of `rank` in an x-slab decomposition of `ncols` columns over `nranks` ranks.
// Even split; when `edge_cols` (the columns an edge rank needs to hold its insertion region plus one
// cutoff) exceeds an even share and there are interior ranks, the two edge slabs get `edge_cols` columns
// and the interior ranks share the rest evenly -- BASELINE config 5: 16 M atoms on 8 GPUs, buffer
// 0.15 Lx > Lx / 8 (SURVEY 8(e) "non-uniform slabs").  Pure host arithmetic: no device, no context.
extern "C" int obmd_slab_partition(int ncols, int nranks, int rank, int edge_cols, int *c0, int *c1)
{
  if (!c0 || !c1 || nranks < 1 || rank < 0 || rank >= nranks || ncols < 2 * nranks) return OBMD_ERR_ARG;
  const int P = nranks, r = rank;
  const int per = ncols / P, rem = ncols % P;
  *c0 = r * per + (r < rem ? r : rem);
  *c1 = *c0 + per + (r < rem ? 1 : 0);
  if (P > 2 && edge_cols > per) {
    const int inner = ncols - 2 * edge_cols;
    if (inner < 2 * (P - 2)) return OBMD_ERR_ARG;
    const int iper = inner / (P - 2), irem = inner % (P - 2);
    if (r == 0) { *c0 = 0; *c1 = edge_cols; }
    else if (r == P - 1) { *c0 = ncols - edge_cols; *c1 = ncols; }
    else {
      const int k = r - 1;
      *c0 = edge_cols + k * iper + (k < irem ? k : irem);
      *c1 = *c0 + iper + (k < irem ? 1 : 0);
    }
  }
  return OBMD_OK;
}

// Cell grid (cell size >= cutmax) and this rank's window of it: owned columns [gx0, gx1) plus one
// ghost column towards each x-neighbour.
int setup_grid(obmd_ctx *c)
{
  const PairTab &pt = c->pt;
  Grid &g = c->grid;
  int nn[3];
  for (int d = 0; d < 3; d++) {
    double L = c->boxhi[d] - c->boxlo[d];
    nn[d] = (int) floor(L / pt.cutmax);
    if (nn[d] < 1) nn[d] = 1;
    g.lo[d] = c->boxlo[d];
    g.hi[d] = c->boxhi[d];
    g.prd[d] = L;
    g.half[d] = 0.5 * L;
    g.inv[d] = nn[d] / L;
  }
  NEED(nn[1] >= 3 && nn[2] >= 3, OBMD_ERR_ARG,
       "obmd_pair_init: periodic box needs >= 3 cells of size cut (%g) in y and z", pt.cutmax);
  NEED(nn[0] >= 2 * c->nranks, OBMD_ERR_ARG, "slab run: fewer than 2 cell columns per rank (%d columns, %d ranks)",
       nn[0], c->nranks);
  const int P = c->nranks, r = c->myrank;
  // The edge ranks run the USHER trials of their side, so the insertion regions (region5 / region6)
  // plus one cutoff must fit inside the first / last slab (obmd_slab_partition widens them if needed).
  int edge = 0;
  if (P > 2 && c->have_fix) {
    const double w = (c->boxhi[0] - c->boxlo[0]) / nn[0];
    const int needL = (int) ceil((c->fd.region[4].hi[0] + pt.cutmax - c->boxlo[0]) / w) + 1;
    const int needR = (int) ceil((c->boxhi[0] - c->fd.region[5].lo[0] + pt.cutmax) / w) + 1;
    edge = needL > needR ? needL : needR;
  }
  NEED(obmd_slab_partition(nn[0], P, r, edge, &c->gx0, &c->gx1) == OBMD_OK, OBMD_ERR_ARG,
       "slab run: the insertion regions leave fewer than 2 cell columns per interior rank (%d columns, %d ranks)",
       nn[0], P);
  g.gnx = nn[0];
  g.cxoff = c->gx0 > 0 ? c->gx0 - 1 : 0;
  const int hi = c->gx1 < nn[0] ? c->gx1 + 1 : nn[0];
  g.nx = hi - g.cxoff;
  g.own_a = c->gx0 - g.cxoff;
  g.own_b = c->gx1 - 1 - g.cxoff;
  g.ny = nn[1];
  g.nz = nn[2];
  for (int d = 0; d < 3; d++) g.w[d] = g.prd[d] / nn[d];
  NEED((double) g.nx * nn[1] * nn[2] < 2.0e9, OBMD_ERR_CAPACITY, "obmd_pair_init: too many cells");
  g.ncell = g.nx * nn[1] * nn[2];
  if (g.ncell + 6 > c->cell_cap) {
    cudaFree(c->cell_count); cudaFree(c->cell_start); cudaFree(c->scan_tmp); cudaFree(c->extra_head);
    c->cell_cap = g.ncell + 6;
    CK(cudaMalloc(&c->cell_count, sizeof(int) * c->cell_cap));
    CK(cudaMalloc(&c->cell_start, sizeof(int) * c->cell_cap));
    CK(cudaMalloc(&c->scan_tmp, sizeof(int) * (c->cell_cap / 4096 + 2)));
    CK(cudaMalloc(&c->extra_head, sizeof(int) * c->cell_cap));
  }
  LAUNCH(k_fill_int, cdiv(c->cell_cap, 256), 256, 0, c->extra_head, c->cell_cap, -1);
  CK(cudaStreamSynchronize(c->stream));
  c->nextra = c->narr = c->nghost = c->nins = 0;
  c->cells_valid = c->rec_valid = false;
  return OBMD_OK;
}

extern "C" double obmd_pair_single(const obmd_ctx *c, int itype, int jtype, double rsq, double *fforce)
{
  // pair_dpd.cpp:401-418 with factor_dpd = 1
  double r = sqrt(rsq);
  if (!c || r < 1.0e-10) {
    if (fforce) *fforce = 0.0;
    return 0.0;
  }
  double rinv = 1.0 / r;
  double wd = 1.0 - r / c->pt.cut[itype][jtype];
  if (fforce) *fforce = c->pt.a0[itype][jtype] * wd * 1.0 * rinv;
  return 0.5 * c->pt.a0[itype][jtype] * c->pt.cut[itype][jtype] * wd * wd;
}

// ---- fix obmd ----------------------------------------------------------------------------

static double ranpark_uniform_host(int &seed)  // random_park.cpp:43-50
{
  int k = seed / 127773;
  seed = 16807 * (seed - k * 127773) - 2836 * k;
  if (seed < 0) seed += 2147483647;
  return (1.0 / 2147483647) * seed;
}

extern "C" int obmd_fix_create(obmd_ctx *c, const obmd_fix_args *a)
{
  NEED(c && a, OBMD_ERR_ARG, "obmd_fix_create: NULL argument");
  NEED(c->have_box && c->ntypes > 0, OBMD_ERR_STATE, "obmd_fix_create: set box and masses first");
  NEED(a->seed > 0, OBMD_ERR_ARG, "Illegal fix obmd command");                         // :91
  NEED(a->ntype > 0 && a->ntype <= c->ntypes, OBMD_ERR_ARG,
       "Invalid atom type in fix obmd command");                                        // :171-172
  NEED(a->buffer_size > 0.0, OBMD_ERR_ARG,
       "Illegal fix obmd command. Parameter buffersize should be larger than 0.0");     // :1988-1990
  NEED(a->g_fac >= 0.0 && a->g_fac <= 1.0, OBMD_ERR_ARG,
       "Illegal fix obmd command. Parameter gfac should be between 0.0 and 1.0");       // :1995-1996
  NEED(a->step_parallel == 0, OBMD_ERR_ARG,
       "For now, only the smooth transition distribution function is implemented.");    // :2013-2015
  NEED(a->step_perp == 1, OBMD_ERR_ARG,
       "For now, only the heaviside step distribution function is implemented.");       // :2019-2020
  NEED(a->usherflag == 1, OBMD_ERR_ARG, "You must pick either usher or near");          // :2163
  NEED(a->nattempt >= 1 && a->nattempt <= 4096, OBMD_ERR_ARG, "Illegal fix obmd command");
  NEED(a->maxattempt >= 1, OBMD_ERR_ARG, "Illegal fix obmd command");
  NEED(a->tau != 0.0, OBMD_ERR_ARG, "Illegal fix obmd command");
  const obmd_block &r1 = a->region[0];  // :181-191 insertion region inside the box
  NEED(!(r1.lo[0] < c->boxlo[0] || r1.hi[0] > c->boxhi[0] || r1.lo[1] < c->boxlo[1] ||
         r1.hi[1] > c->boxhi[1] || r1.lo[2] < c->boxlo[2] || r1.hi[2] > c->boxhi[2]),
       OBMD_ERR_ARG, "Insertion region extends outside simulation box");
  c->fa = *a;
  memset(&c->fs, 0, sizeof(c->fs));
  c->fs.rng_seed = a->seed;  // :275-276 RanPark(seed), 30 warm-up draws
  for (int ii = 0; ii < 30; ii++) ranpark_uniform_host(c->fs.rng_seed);
  FixDev &fd = c->fd;
  for (int r = 0; r < 6; r++) fd.region[r] = a->region[r];
  for (int d = 0; d < 3; d++) {
    fd.boxlo[d] = c->boxlo[d];
    fd.boxhi[d] = c->boxhi[d];
  }
  fd.buffer_size = a->buffer_size;
  fd.g_fac = a->g_fac;
  fd.ntype = a->ntype;
  fd.etarget = a->etarget;
  fd.ds0 = a->ds0;
  fd.uovlp = a->uovlp;
  fd.dsovlp = a->dsovlp;
  fd.eps = a->eps;
  fd.nattempt = a->nattempt;
  c->have_fix = true;
  // the slab partition depends on the insertion regions (setup_grid): redo it if atoms are not resident yet
  if (c->nranks > 2 && c->pair_ready && c->n == 0) return setup_grid(c);
  return OBMD_OK;
}

extern "C" int obmd_fix_set_flux(obmd_ctx *c, double pxx, double pxy, double pxz, double dpxx, double freq,
                                 double alpha, double tau, double nbuf)
{
  NEED(c && c->have_fix, OBMD_ERR_STATE, "obmd_fix_set_flux: call obmd_fix_create first");
  NEED(tau != 0.0, OBMD_ERR_ARG, "Illegal fix obmd command");
  c->fa.pxx = pxx; c->fa.pxy = pxy; c->fa.pxz = pxz; c->fa.dpxx = dpxx; c->fa.freq = freq;
  c->fa.alpha = alpha; c->fa.tau = tau; c->fa.nbuf = nbuf;
  return OBMD_OK;
}

extern "C" int obmd_set_usher_order(obmd_ctx *c, int by_tag)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  c->usher_by_tag = by_tag != 0;
  return OBMD_OK;
}

// Comm::exchange + Comm::borders for callers that drive the hooks themselves (obmd_run does this itself)
extern "C" int obmd_exchange(obmd_ctx *c)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  if (c->nranks == 1) return OBMD_OK;
  NEED(c->have_fix && c->cells_valid, OBMD_ERR_STATE, "obmd_exchange: call obmd_build_cells first");
  int rc = fix_pre_begin(c);
  if (rc) return rc;
  return comm_exchange(c);
}

extern "C" int obmd_fix_get_state(const obmd_ctx *c, obmd_fix_state *out)
{
  if (!c || !out || !c->have_fix) return OBMD_ERR_STATE;
  *out = c->fs;
  return OBMD_OK;
}

extern "C" int obmd_fix_set_state(obmd_ctx *c, const obmd_fix_state *in)
{
  NEED(c && in && c->have_fix, OBMD_ERR_STATE, "obmd_fix_set_state: call obmd_fix_create first");
  NEED(in->rng_seed > 0, OBMD_ERR_ARG, "Invalid seed for Park random # generator");
  c->fs = *in;
  return OBMD_OK;
}

// ---- atoms: host <-> device -----------------------------------------------------------

template <typename T>
static cudaError_t regrow(T *&p, size_t oldn, size_t newn, cudaStream_t s)
{
  T *q = nullptr;
  cudaError_t e = cudaMalloc(&q, sizeof(T) * newn);
  if (e != cudaSuccess) return e;
  if (p && oldn) e = cudaMemcpyAsync(q, p, sizeof(T) * oldn, cudaMemcpyDeviceToDevice, s);
  cudaStreamSynchronize(s);
  cudaFree(p);
  p = q;
  return e;
}

int obmd_ensure_capacity(obmd_ctx *c, int cap)
{
  if (cap <= c->cap) return OBMD_OK;
  cudaStreamSynchronize(c->stream2);  // side-stream kernels may still hold the old pointers
  size_t o = (size_t) c->cap, nn = (size_t) cap + (size_t) cap / 8 + 4096;
  NEED(nn < 2000000000ull, OBMD_ERR_CAPACITY, "too many atoms for 32-bit indexing");
  Planes *pl[2] = {&c->cur, &c->alt};
  for (int k = 0; k < 2; k++) {
    Planes &p = *pl[k];
    CK(regrow(p.x, o, nn, c->stream)); CK(regrow(p.y, o, nn, c->stream)); CK(regrow(p.z, o, nn, c->stream));
    CK(regrow(p.vx, o, nn, c->stream)); CK(regrow(p.vy, o, nn, c->stream)); CK(regrow(p.vz, o, nn, c->stream));
    CK(regrow(p.tag, o, nn, c->stream)); CK(regrow(p.lidx, o, nn, c->stream));
    CK(regrow(p.type, o, nn, c->stream)); CK(regrow(p.image, o, nn, c->stream));
  }
  CK(regrow(c->fx, o, nn, c->stream)); CK(regrow(c->fy, o, nn, c->stream)); CK(regrow(c->fz, o, nn, c->stream));
  CK(regrow(c->rec, o, nn, c->stream)); CK(regrow(c->aux, o, nn, c->stream));
  CK(regrow(c->cid, o, nn, c->stream)); CK(regrow(c->rank, o, nn, c->stream));
  CK(regrow(c->perm, o, nn, c->stream));
  CK(regrow(c->extra_next, o, nn, c->stream)); CK(regrow(c->extra_cell, o, nn, c->stream));
  CK(regrow(c->stage_d, 9 * o, 9 * nn, c->stream));
  CK(regrow(c->stage_i, 3 * o, 3 * nn, c->stream));
  c->cap = (int) nn;
  return OBMD_OK;
}

__global__ void k_unpack_upload(Planes p, double *fx, double *fy, double *fz, const double *sx,
                                const double *sv, const double *sf, const int *st, const int *sg,
                                const int *si, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  p.x[i] = sx[3 * i]; p.y[i] = sx[3 * i + 1]; p.z[i] = sx[3 * i + 2];
  p.vx[i] = sv[3 * i]; p.vy[i] = sv[3 * i + 1]; p.vz[i] = sv[3 * i + 2];
  if (sf) { fx[i] = sf[3 * i]; fy[i] = sf[3 * i + 1]; fz[i] = sf[3 * i + 2]; }
  else { fx[i] = 0.0; fy[i] = 0.0; fz[i] = 0.0; }
  p.type[i] = st[i];
  p.tag[i] = sg[i];
  p.image[i] = si ? si[i] : IMG0;
  p.lidx[i] = i;
}

extern "C" int obmd_upload(obmd_ctx *c, int n, const double *x, const double *v, const double *f,
                           const int *type, const int *tag, const int *image, int capacity)
{
  NEED(c && x && v && type && tag, OBMD_ERR_ARG, "obmd_upload: NULL argument");
  NEED(n > 0, OBMD_ERR_ARG, "obmd_upload: n must be > 0");
  int rc = obmd_ensure_capacity(c, capacity > n ? capacity : n + n / 8 + 1024);
  if (rc) return rc;
  if (c->pair_ready) fix_clear_extras(c);
  double *sx = c->stage_d, *sv = c->stage_d + 3 * (size_t) c->cap, *sf = c->stage_d + 6 * (size_t) c->cap;
  int *st = c->stage_i, *sg = c->stage_i + c->cap, *si = c->stage_i + 2 * (size_t) c->cap;
  CK(cudaMemcpyAsync(sx, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(sv, v, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  if (f) CK(cudaMemcpyAsync(sf, f, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(st, type, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(sg, tag, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
  if (image) CK(cudaMemcpyAsync(si, image, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
  LAUNCH(k_unpack_upload, cdiv(n, 256), 256, 0, c->cur, c->fx, c->fy, c->fz, sx, sv, f ? sf : nullptr, st,
         sg, image ? si : nullptr, n);
  CK(cudaStreamSynchronize(c->stream));
  c->n = n;
  c->nsorted = c->nparked = c->nextra = c->narr = c->nghost = c->nins = 0;
  c->pre_begun = false;
  c->cells_valid = c->rec_valid = false;
  c->forces_valid = (f != nullptr);
  return OBMD_OK;
}

__global__ void k_unpack_xv(Planes p, const double *sx, const double *sv, int n)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int i = p.lidx[s];
  p.x[s] = sx[3 * i]; p.y[s] = sx[3 * i + 1]; p.z[s] = sx[3 * i + 2];
  p.vx[s] = sv[3 * i]; p.vy[s] = sv[3 * i + 1]; p.vz[s] = sv[3 * i + 2];
}

extern "C" int obmd_upload_xv(obmd_ctx *c, int n, const double *x, const double *v)
{
  NEED(c && x && v, OBMD_ERR_ARG, "obmd_upload_xv: NULL argument");
  NEED(n == c->n, OBMD_ERR_STATE, "obmd_upload_xv: n (%d) differs from the resident atom count (%d)", n,
       c->n);
  double *sx = c->stage_d, *sv = c->stage_d + 3 * (size_t) c->cap;
  CK(cudaMemcpyAsync(sx, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(sv, v, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  LAUNCH(k_unpack_xv, cdiv(n, 256), 256, 0, c->cur, sx, sv, n);
  CK(cudaStreamSynchronize(c->stream));
  c->cells_valid = c->rec_valid = false;
  return OBMD_OK;
}

__global__ void k_pack_download(Planes p, const double *fx, const double *fy, const double *fz, double *sx,
                                double *sv, double *sf, int *st, int *sg, int *si, int n)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int i = p.lidx[s];
  sx[3 * i] = p.x[s]; sx[3 * i + 1] = p.y[s]; sx[3 * i + 2] = p.z[s];
  sv[3 * i] = p.vx[s]; sv[3 * i + 1] = p.vy[s]; sv[3 * i + 2] = p.vz[s];
  sf[3 * i] = fx[s]; sf[3 * i + 1] = fy[s]; sf[3 * i + 2] = fz[s];
  st[i] = p.type[s];
  sg[i] = p.tag[s];
  si[i] = p.image[s];
}

extern "C" int obmd_download(obmd_ctx *c, int *n_out, double *x, double *v, double *f, int *type, int *tag,
                             int *image)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  NEED(c->nparked == 0, OBMD_ERR_STATE,
       "obmd_download: atoms outside the box are parked; call obmd_pre_exchange first");
  if (c->nghost > 0) {
    int rc = fix_compact_extras(c);
    if (rc) return rc;
  }
  const int n = c->n;
  if (n_out) *n_out = n;
  if (n == 0) return OBMD_OK;
  double *sx = c->stage_d, *sv = c->stage_d + 3 * (size_t) c->cap, *sf = c->stage_d + 6 * (size_t) c->cap;
  int *st = c->stage_i, *sg = c->stage_i + c->cap, *si = c->stage_i + 2 * (size_t) c->cap;
  LAUNCH(k_pack_download, cdiv(n, 256), 256, 0, c->cur, c->fx, c->fy, c->fz, sx, sv, sf, st, sg, si, n);
  if (x) CK(cudaMemcpyAsync(x, sx, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  if (v) CK(cudaMemcpyAsync(v, sv, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  if (f) CK(cudaMemcpyAsync(f, sf, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  if (type) CK(cudaMemcpyAsync(type, st, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
  if (tag) CK(cudaMemcpyAsync(tag, sg, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
  if (image) CK(cudaMemcpyAsync(image, si, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return OBMD_OK;
}

extern "C" int obmd_nlocal(const obmd_ctx *c) { return c ? c->n - c->nghost : 0; }  // owned atoms

// ---- hooks -------------------------------------------------------------------------------

extern "C" int obmd_nve_initial(obmd_ctx *c)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  return fix_nve_initial(c);
}

extern "C" int obmd_nve_final(obmd_ctx *c)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  return fix_nve_final(c);
}

extern "C" int obmd_build_cells(obmd_ctx *c)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  if (c->nghost > 0) {  // a hook-driven slab step leaves the ghosts of the previous exchange behind
    int rc = fix_compact_extras(c);
    if (rc) return rc;
  }
  fix_clear_extras(c);
  return cells_build(c);
}

extern "C" int obmd_pre_exchange(obmd_ctx *c, obmd_pre_exchange_report *report, obmd_trial *trials,
                                 int maxtrials)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  int rc = fix_pre_exchange(c, trials, maxtrials);
  if (report) *report = c->rep;
  return rc;
}

extern "C" int obmd_pbc(obmd_ctx *c)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  return fix_pbc(c);
}

extern "C" int obmd_pair_compute(obmd_ctx *c, long long ntimestep, int eflag, int vflag)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  NEED(c->nparked == 0, OBMD_ERR_STATE,
       "obmd_pair_compute: %d atoms are outside the box in x; obmd_pre_exchange must delete them first",
       c->nparked);
  return pair_compute(c, ntimestep, eflag, vflag);
}

extern "C" int obmd_pair_tallies(obmd_ctx *c, double *eng_vdwl, double virial[6])
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  int rc = read_scalars(c);
  if (rc) return rc;
  if (eng_vdwl) *eng_vdwl = c->sc_h->eng;
  if (virial)
    for (int k = 0; k < 6; k++) virial[k] = c->sc_h->virial[k];
  return OBMD_OK;
}

extern "C" int obmd_post_force(obmd_ctx *c)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  return fix_post_force(c);
}

extern "C" int obmd_zero_forces(obmd_ctx *c)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  if (c->n > 0) {
    CK(cudaMemsetAsync(c->fx, 0, sizeof(double) * c->n, c->stream));
    CK(cudaMemsetAsync(c->fy, 0, sizeof(double) * c->n, c->stream));
    CK(cudaMemsetAsync(c->fz, 0, sizeof(double) * c->n, c->stream));
  }
  return OBMD_OK;
}

extern "C" int obmd_setup(obmd_ctx *c, long long ntimestep)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  NEED(c->pair_ready, OBMD_ERR_STATE, "obmd_setup: pair style must be initialised");
  int rc;
  fix_clear_extras(c);
  if ((rc = obmd_ensure_capacity(c, c->n + c->n / 32 + 4096))) return rc;
  if ((rc = cells_build(c))) return rc;
  NEED(c->nparked == 0, OBMD_ERR_STATE, "obmd_setup: %d atoms lie outside the box in x", c->nparked);
  if (c->nranks > 1 && (rc = comm_exchange(c))) return rc;
  if ((rc = fix_pbc(c))) return rc;
  if ((rc = pair_compute(c, ntimestep, 0, 0))) return rc;
  if (c->have_fix && (rc = fix_post_force(c))) return rc;
  if ((rc = fix_compact_extras(c))) return rc;
  CK(cudaStreamSynchronize(c->stream));
  return OBMD_OK;
}

extern "C" int obmd_comm_slab(const obmd_ctx *c, double *lo, double *hi)
{
  if (!c || !lo || !hi || !c->pair_ready) return OBMD_ERR_STATE;
  const Grid &g = c->grid;
  *lo = g.lo[0] + c->gx0 / g.inv[0];
  *hi = (c->gx1 >= g.gnx) ? g.hi[0] : g.lo[0] + c->gx1 / g.inv[0];
  return OBMD_OK;
}

// launched from inside fix_pre_exchange, right after USHER round 0 has been enqueued
static int pair_on_side_stream(obmd_ctx *c)
{
  CK(cudaEventRecord(c->ev_sorted, c->stream));
  CK(cudaStreamWaitEvent(c->stream2, c->ev_sorted, 0));
  int rc = pair_compute_cols(c, c->hook_step, all_columns(c), c->stream2);
  if (rc) return rc;
  CK(cudaEventRecord(c->ev_int, c->stream2));
  return OBMD_OK;
}

extern "C" int obmd_run(obmd_ctx *c, long long first_step, int nsteps)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  NEED(c->have_fix && c->pair_ready, OBMD_ERR_STATE, "obmd_run: pair style and fix must be initialised");
  int rc;
  for (int s = 0; s < nsteps; s++) {  // verlet.cpp:229-363
    const long long step = first_step + s;
    const auto host_t0 = std::chrono::steady_clock::now();
    struct HostStep {  // host wall time of the step's enqueue work (timing runs only)
      obmd_ctx *c;
      std::chrono::steady_clock::time_point t0;
      ~HostStep()
      {
        if (!c->timing) return;
        TimerSlot &sl = c->timers["host_step"];
        sl.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        sl.launches++;
      }
    } host_step{c, host_t0};
    if ((rc = fix_nve_initial(c))) return rc;
    fix_clear_extras(c);
    if ((rc = obmd_ensure_capacity(c, c->n + c->n / 32 + 4096))) return rc;
    {
      HostTimer hc(c, "host_cells");
      if ((rc = cells_build(c))) return rc;
    }
    if (c->nranks > 1) {
      HostTimer hx(c, "host_comm");
      // slab run: deletion sums first (the arrivals reuse the parked slots), then migration + halo
      // (Comm::exchange / borders); the rest of pre_exchange sees the ghosts
      if ((rc = fix_pre_begin(c))) return rc;
      if ((rc = comm_exchange(c))) return rc;
    }
    if (c->overlap) {
      // PairDPD for every atom this rank owns runs on the side stream.  It starts right behind
      // USHER round 0 (which fills the machine by itself) and overlaps the re-run rounds, which are
      // a handful of latency-bound CTAs.  The pair kernel wraps coordinates on the fly, so it does
      // not need Domain::pbc first; atoms next to an insertion are recomputed afterwards
      // (fix_redo_pair_near_insertions).
      c->hook_step = step;
      // The pair kernel starts before pre_exchange's first host round trip: for large systems it is the
      // longer branch anyway (4 M atoms: 2.42 -> 2.28 ms/step), and for small ones it keeps the GPU busy
      // while the host reads the counts back and prepares the USHER launches (no difference on a quiet
      // host, the difference between 1.3 and 1.7 ms/step on a noisy one).  OBMD_B200_PAIR_FIRST=0
      // launches it right behind USHER round 0 instead.
      // Small systems in steady exchange: USHER round 0 is the longer branch and runs faster with the
      // machine to itself, the pair kernel then overlaps the re-run rounds (a handful of CTAs).  Large
      // systems: the pair kernel is the longer branch and starts first.  OBMD_B200_PAIR_FIRST overrides.
      const bool pair_first = c->pair_first < 0 ? (c->n > 3000000) : (c->pair_first != 0);
      if (pair_first) {
        if ((rc = pair_on_side_stream(c))) return rc;
        c->round0_hook = nullptr;
      } else
        c->round0_hook = pair_on_side_stream;
      c->commit_wait = c->ev_int;
      {
        HostTimer hp(c, "host_pre_exchange");
        rc = fix_pre_exchange(c, nullptr, 0);
      }
      c->commit_wait = nullptr;
      if (rc) return rc;
      NEED(c->round0_hook == nullptr, OBMD_ERR_STATE, "obmd_run: side-stream pair launch was skipped");
      CK(cudaStreamWaitEvent(c->stream, c->ev_int, 0));
      if ((rc = fix_pbc(c))) return rc;
      if ((rc = fix_redo_pair_near_insertions(c, step))) return rc;
      c->forces_valid = true;
    } else {
      if ((rc = fix_pre_exchange(c, nullptr, 0))) return rc;
      if ((rc = fix_pbc(c))) return rc;
      if ((rc = pair_compute(c, step, 0, 0))) return rc;
    }
    {
      HostTimer hf(c, "host_post");
      if ((rc = fix_post_force(c))) return rc;
      if ((rc = fix_compact_extras(c))) return rc;
      if ((rc = fix_nve_final(c))) return rc;
    }
  }
  {
    HostTimer hs(c, "host_final_sync");
    CK(cudaStreamSynchronize(c->stream));
  }
  return OBMD_OK;
}

extern "C" int obmd_reduce_kinetic(obmd_ctx *c, double out[5])
{
  NEED(c && out, OBMD_ERR_ARG, "NULL argument");
  return fix_reduce_kinetic(c, out);
}

extern "C" int obmd_profile_x(obmd_ctx *c, int nbins, double *count, double *mvx)
{
  NEED(c && count && mvx, OBMD_ERR_ARG, "NULL argument");
  return fix_profile_x(c, nbins, count, mvx, nullptr, nullptr);
}

extern "C" int obmd_profile_x2(obmd_ctx *c, int nbins, double *count, double *mvx, double *mv2, double *sxx)
{
  NEED(c && count && mvx, OBMD_ERR_ARG, "NULL argument");
  return fix_profile_x(c, nbins, count, mvx, mv2, sxx);
}

extern "C" int obmd_pair_gaussian_sample(obmd_ctx *c, int n, int tag0, long long ntimestep, double *out)
{
  NEED(c && out && n > 0 && tag0 > 0, OBMD_ERR_ARG, "obmd_pair_gaussian_sample: bad argument");
  NEED(c->have_pair, OBMD_ERR_STATE, "obmd_pair_gaussian_sample: call obmd_pair_settings first");
  return pair_gaussian_sample(c, n, tag0, ntimestep, out);
}

// ---- timing ------------------------------------------------------------------------------

static cudaEvent_t pool_get(obmd_ctx *c)
{
  if (!c->ev_pool.empty()) {
    cudaEvent_t e = c->ev_pool.back();
    c->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void timer_begin(obmd_ctx *c, cudaEvent_t *e0, cudaStream_t st)
{
  *e0 = pool_get(c);
  cudaEventRecord(*e0, st);
}

void timer_end(obmd_ctx *c, const char *name, cudaEvent_t e0, long long launches, cudaStream_t st)
{
  cudaEvent_t e1 = pool_get(c);
  cudaEventRecord(e1, st);
  c->pending.push_back(PendingTimer{name, e0, e1, launches});
}

static void timers_resolve(obmd_ctx *c)
{
  if (c->pending.empty()) return;
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->stream2);
  for (const PendingTimer &p : c->pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.e0, p.e1);
    TimerSlot &s = c->timers[p.name];
    s.ms += ms;
    s.launches += p.launches;
    c->ev_pool.push_back(p.e0);
    c->ev_pool.push_back(p.e1);
  }
  c->pending.clear();
}

extern "C" int obmd_timing_enable(obmd_ctx *c, int on)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  timers_resolve(c);
  c->timing = on != 0;
  return OBMD_OK;
}

extern "C" int obmd_timing_get(obmd_ctx *c, const char *name, double *ms, long long *launches)
{
  NEED(c && name, OBMD_ERR_ARG, "NULL argument");
  timers_resolve(c);
  auto it = c->timers.find(name);
  if (ms) *ms = it == c->timers.end() ? 0.0 : it->second.ms;
  if (launches) *launches = it == c->timers.end() ? 0 : it->second.launches;
  return OBMD_OK;
}

extern "C" int obmd_timing_reset(obmd_ctx *c)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  timers_resolve(c);
  c->timers.clear();
  return OBMD_OK;
}

extern "C" long long obmd_launch_count(const obmd_ctx *c) { return c ? c->launches : 0; }
extern "C" void *obmd_stream(const obmd_ctx *c) { return c ? (void *) c->stream : nullptr; }
