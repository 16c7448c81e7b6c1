// Internal context of libobmd_b200.so (not part of the ABI; see include/obmd_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "obmd_b200.h"

#define MAXT OBMD_MAXTYPES
#define IMG0 ((512 << 20) | (512 << 10) | 512)  // lmptype.h IMGMAX in all three fields
#define IMGMASK 1023
#define IMGBITS 10
#define IMG2BITS 20

#define NSM 148  // B200 SM count: grids are sized in multiples of this where they are persistent

// ---- device-side parameter blocks (passed by value to kernels) -----------------

// SoA planes of one atom buffer.  Index = storage slot (cell-sorted after build_cells).
struct Planes {
  double *x, *y, *z, *vx, *vy, *vz;
  int *tag, *lidx, *type, *image;
};

// One cell-sorted atom as the pair kernel stages it (pair_brick.cuh): a 48-byte record of positions
// and velocities plus a 16-byte record {tag, type, 0, 0}.  The copy engine's granule is 16 bytes, so any
// run of either is a legal bulk copy.  Written by the sort's gather next to the planes.
struct __align__(16) AtomRec {
  double x, y, z, vx, vy, vz;
};

// Cell grid: cell size >= max cutoff; linear index (cx*ny + cy)*nz + cz, x slowest, z fastest,
// so each open-boundary buffer is one contiguous slot range and a z-row of a stencil is contiguous.
struct Grid {
  int nx, ny, nz, ncell;  // nx = columns stored on this rank (owned + ghost columns)
  int gnx, cxoff;         // global column count; global index of local column 0
  int own_a, own_b;       // owned local columns (inclusive); the others hold ghosts only
  double lo[3], hi[3], prd[3], inv[3], half[3];
  double w[3];            // cell widths prd / n
};

// Atoms appended after the cell-sorted range [0, base): arrivals from neighbour slabs (mine),
// ghosts of the neighbour slabs (not mine), atoms inserted by USHER (mine).  They hang off the
// per-cell extra lists until the next sort.
struct Extras {
  int base, narr, nghost, nins;
};

struct PairTab {
  int ntypes;
  double a0[MAXT + 1][MAXT + 1], gamma[MAXT + 1][MAXT + 1], sigma[MAXT + 1][MAXT + 1];
  double cut[MAXT + 1][MAXT + 1], cutsq[MAXT + 1][MAXT + 1], cutinv[MAXT + 1][MAXT + 1];
  double cutmax;
  double dtinvsqrt;
  unsigned seed;
  double mass[MAXT + 1];
};

struct FixDev {  // what the fix kernels need
  obmd_block region[6];
  double boxlo[3], boxhi[3];
  double buffer_size, g_fac;
  int ntype;
  double etarget, ds0, uovlp, dsovlp, eps;
  int nattempt;
};

// Device-resident per-step scalars (one struct, read back with a single small copy).
struct StepScalars {
  int cnt_left, cnt_right;  // Group::count
  int maxtag;
  int ntail;                // survivors with lidx >= nprime found by the scan
  int nonfinite;            // non-numeric coordinate seen while binning
  int colmax;               // atoms in the fullest cell column (cells.cu; sizes the pair kernel's shared-memory window)
  int usher_error;          // see OBMD_ERR_USHER / capacity
  // USHER commit rounds run without host round trips: counters are double-buffered by round parity
  int nchg2[2];             // trials whose state changed in round r (index r & 1)
  int ndirty2[2];           // trials to re-run in round r + 1
  int naccept;              // accepted trials after the latest round
  int nrounds_work;         // rounds whose evaluation had work to do
  int trial_cursor;         // next entry of the round's trial list to hand to a warp
  int nused, nused_prev;    // maxattempt > 1: attempts of the RNG stream the serial loop consumes (a prefix)
  int nspec;                // accepted-atom entries published so far in this call (usher.cu)
  int npairs;               // (trial, change record) pairs queued by the dirty test of a re-run round
  int rng_seed;             // RanPark state after the deletion draws
  int wsum_done;            // blocks of k_weight_sums that have published their partials (reset by the last one)
  int redo_overflow;        // k_mark_redo found more atoms near the insertions than its list holds (sticky,
                            // outside the range pre_exchange zeroes)
  double vnewl[3], vnewr[3];
  double gsum[4];           // G_left, G_right, Gperp_left, Gperp_right
  double kin[5];
  double eng, virial[6];
};

struct PendingTimer;
// A trial whose state changed in a commit round: old / new atom position (if it was / is accepted)
struct ChgRec {
  double o[3], n[3];
  int q, flags;  // trial index; bit 0: old state was accepted, bit 1: new state is accepted
};
struct TimerSlot {
  double ms = 0.0;
  long long launches = 0;
};

struct obmd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;   // main stream (high priority): everything the fix orders
  cudaStream_t stream2 = nullptr;  // side stream: interior pbc + PairDPD overlapped with USHER
  cudaEvent_t ev_sorted = nullptr, ev_int = nullptr;
  cudaEvent_t commit_wait = nullptr;  // if set, pre_exchange waits for it before it links new atoms
  int (*round0_hook)(obmd_ctx *) = nullptr;  // called once, right after USHER round 0 is enqueued
  long long hook_step = 0;
  int *redo_flag = nullptr, *redo_list = nullptr, *redo_count = nullptr;  // pair recompute near insertions
  int redo_cap = 0, redo_list_cap = 0;
  bool overlap = true;
  int pair_first = -1;  // side-stream pair kernel before pre_exchange (1), right behind USHER round 0 (0), by size (-1)
  int tile_kernel = 2;       // OBMD_B200_TILE: 2 bulk-staged bricks (pair_brick.cuh), 1 thread-staged row tiles (pair_row.cuh), 0 per-atom list kernel
  int row_rows = 0, row_zc = 0, row_threads = 0, row_cols = 0;  // tuning overrides (OBMD_B200_ROWS, _ZC, _PT, _COLS); 0 = automatic
  size_t row_smem_set = 0;
  double row_capscale = 1.0;  // OBMD_B200_CAPSCALE (tests): < 1 forces tiles onto the overflow path
  char err[1024] = {0};
  long long launches = 0;
  bool timing = false;
  std::map<std::string, TimerSlot> timers;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<struct PendingTimer> pending;

  // domain
  bool have_box = false;
  double boxlo[3], boxhi[3];
  double sublo, subhi;
  int ntypes = 0;
  double mass[MAXT + 1];
  double dt = 0.0;
  // pair
  bool have_pair = false, pair_ready = false;
  double temperature = 0.0, cut_global = 0.0, boltz = 1.0;
  int pair_seed = 0;
  bool setflag[MAXT + 1][MAXT + 1];
  PairTab pt;
  Grid grid;
  // fix
  bool have_fix = false;
  obmd_fix_args fa;
  obmd_fix_state fs;
  FixDev fd;

  // atoms
  int cap = 0;      // capacity of every per-atom plane
  int n = 0;        // atoms stored (live + parked + extras)
  int nsorted = 0;  // [0,nsorted) cell-sorted live atoms
  int nparked = 0;  // [nsorted, nsorted+nparked) atoms outside the box in x (before pre_exchange)
  int nextra = 0;   // atoms appended after the sort: [nsorted, nsorted+nextra) = arrivals|ghosts|inserted
  int narr = 0, nghost = 0, nins = 0;
  int colmax = 0;   // atoms in the fullest cell column at the latest cell build
  int nmigL = 0, nmigR = 0, seg_migL = 0, seg_migR = 0;  // migrant segments after the sort
  int ndel_local = 0;
  int col_ab[4] = {0, 0, 0, 0};  // slot ranges of the first / last owned column (halo source)
  bool pre_begun = false;
  // slab decomposition (comm.cu); nranks == 1: whole box, no communication
  int myrank = 0, nranks = 1;
  void *nccl = nullptr;        // ncclComm_t
  int gx0 = 0, gx1 = 0;        // owned global columns [gx0, gx1)
  double *sendL = nullptr, *sendR = nullptr, *recvL = nullptr, *recvR = nullptr;
  int comm_cap = 0;            // capacity of each buffer in atoms (8 doubles per atom)
  double *red_d = nullptr;     // 32 doubles for the packed allreduces
  double *red_h = nullptr;     // pinned
  bool cells_valid = false;
  bool forces_valid = false;
  Planes cur, alt;
  AtomRec *rec = nullptr;      // records of the sorted atoms [0, nsorted), as of the latest cell build
  int4 *aux = nullptr;         // {tag, type, 0, 0} of the same atoms
  bool rec_valid = false;      // nothing has changed x, v, tag or type of a sorted atom since (Domain::pbc may have run)
  double *fx = nullptr, *fy = nullptr, *fz = nullptr;
  int *cid = nullptr, *rank = nullptr, *perm = nullptr;
  int *cell_count = nullptr, *cell_start = nullptr, *scan_tmp = nullptr;
  int cell_cap = 0;
  int *extra_head = nullptr, *extra_next = nullptr, *extra_cell = nullptr;
  // staging for host <-> device in LAMMPS layout
  double *stage_d = nullptr;  // 9*cap doubles
  int *stage_i = nullptr;     // 3*cap ints
  // scalars
  StepScalars *sc_d = nullptr;
  StepScalars *sc_h = nullptr;  // pinned
  // deletion scratch
  int del_cap = 0;
  int *del_key = nullptr, *del_val = nullptr, *tail_key = nullptr, *tail_val = nullptr;
  int *del_src = nullptr;
  double *del_mv = nullptr;
  // USHER scratch
  int tr_cap = 0, tr_natt = 0;
  obmd_trial *tr_d = nullptr;       // trial records (state of the latest round)
  obmd_trial *tr_h = nullptr;       // pinned copy (allocated when a caller asks for the trial records)
  int tr_h_cap = 0;
  double *tr_path = nullptr;        // [tr_cap][(natt+1)*3]
  int *tr_npts = nullptr;
  float *tr_bbox = nullptr;         // [tr_cap][6] bounding box of the evaluated points (rounded outwards)
  int *tr_list = nullptr;           // trials to re-run this round
  int *tr_pairs = nullptr;          // [8 * tr_cap] int2 queue of the dirty test
  int *tr_kfirst = nullptr;         // first evaluation point of a listed trial that must be redone
  struct ChgRec *tr_chg = nullptr;  // compact records of the trials whose state changed this round
  int *acc_list = nullptr;          // accepted trial ids in order (final)
  // atoms accepted so far in this call: append-only entries (usher.cu)
  int *spec_trial = nullptr;
  int *spec_of_trial = nullptr;     // [tr_cap] entry of trial t's accepted atom, -1 if none
  double *spec_pos = nullptr;
  int ush_group = 32;               // lanes per trial in round 0 (OBMD_B200_USH_G: 8, 16, 32)
  bool usher_by_tag = false;        // serial runs: sum USHER contributions in tag order (obmd_set_usher_order)
  int fix_cluster = 16;             // CTAs in the cluster of the USHER fix-up kernel (OBMD_B200_FIX_CL: 1, 8 or 16)
  int ush_ctas_per_sm = 6;          // persistent grid of round 0 (OBMD_B200_USH_CTAS)
  double *partials = nullptr;       // block partial sums for deterministic reductions
  double *vatom = nullptr;          // per-atom virial xx of the latest tally call (stress profile)
  int vatom_cap = 0;
  bool vatom_valid = false;
  int partial_cap = 0;
  // bonded path (bond.cu): harmonic coefficients, topology in CSR form indexed by atom tag, tag -> slot map
  int nbondtypes = 0, bond_style = 0;  // 0 harmonic, 1 fene
  double bond_k[MAXT + 1], bond_r0[MAXT + 1], bond_eps[MAXT + 1], bond_sig[MAXT + 1];
  int nbonds = 0, bond_maxtag = 0;
  int *bond_off = nullptr, *bond_partner = nullptr, *bond_ptype = nullptr, *tag_map = nullptr;
  double *bond_e = nullptr;         // [7] bond energy and virial of the latest obmd_bond_compute(eflag = 1)
  // last pre_exchange report
  obmd_pre_exchange_report rep;
};

// ---- error helpers -----------------------------------------------------------------
int obmd_fail(obmd_ctx *c, int code, const char *fmt, ...);
extern char g_create_err[1024];

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return obmd_fail(c, OBMD_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
                       cudaGetErrorString(e_));                                               \
  } while (0)

#define NEED(cond, code, ...) \
  do {                        \
    if (!(cond)) return obmd_fail(c, code, __VA_ARGS__); \
  } while (0)

// Kernel launch with bookkeeping; `c` must be in scope.
#define LAUNCH(kern, grid, block, smem, ...)                    \
  do {                                                          \
    kern<<<(grid), (block), (smem), c->stream>>>(__VA_ARGS__);  \
    c->launches++;                                              \
  } while (0)
#define LAUNCH_ON(st, kern, grid, block, smem, ...)             \
  do {                                                          \
    kern<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);       \
    c->launches++;                                              \
  } while (0)

// A set of cell columns (cx intervals; x is the slowest cell index, so each interval is one
// contiguous slot range) plus, optionally, the atoms inserted after the sort.
struct ColRange {
  int n;            // number of intervals (0..2)
  int a[2], b[2];   // inclusive column intervals
  int extras;       // 0: none, 1: the extras this rank owns (arrivals + insertions), 2: all extras
};

// Sync-free kernel-family timer: records a CUDA event pair on the launching stream and defers
// cudaEventElapsedTime to obmd_timing_get(), so timing can stay on inside a timed region.
struct PendingTimer {
  const char *name;
  cudaEvent_t e0, e1;
  long long launches;
};
void timer_begin(obmd_ctx *c, cudaEvent_t *e0, cudaStream_t st);
void timer_end(obmd_ctx *c, const char *name, cudaEvent_t e0, long long launches, cudaStream_t st);

struct ScopedTimer {
  obmd_ctx *c;
  const char *name;
  long long l0;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr;
  ScopedTimer(obmd_ctx *c_, const char *n, cudaStream_t s = nullptr)
      : c(c_), name(n), l0(c_->launches), st(s ? s : c_->stream)
  {
    if (c->timing) timer_begin(c, &e0, st);
  }
  ~ScopedTimer()
  {
    if (c->timing && e0) timer_end(c, name, e0, c->launches - l0, st);
  }
};

// Host wall-clock time of a scope, accumulated under `name` when timing is on (diagnostics: where the
// enqueueing thread spends its time, as opposed to the CUDA-event timers above).
struct HostTimer {
  obmd_ctx *c;
  const char *name;
  std::chrono::steady_clock::time_point t0;
  HostTimer(obmd_ctx *c_, const char *n) : c(c_), name(n), t0(std::chrono::steady_clock::now()) {}
  ~HostTimer()
  {
    if (!c->timing) return;
    TimerSlot &sl = c->timers[name];
    sl.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    sl.launches++;
  }
};

static inline int cdiv(long long a, int b) { return (int) ((a + b - 1) / b); }

// ---- internal entry points of the translation units ---------------------------------
int obmd_ensure_capacity(obmd_ctx *c, int cap);
int cells_build(obmd_ctx *c);                                   // cells.cu
int pair_compute(obmd_ctx *c, long long step, int eflag, int vflag);  // pair.cu
int pair_compute_cols(obmd_ctx *c, long long step, const ColRange &cr, cudaStream_t st);
int pair_add_insertions(obmd_ctx *c, long long step, const int *list, const int *count, int cap);
int fix_redo_pair_near_insertions(obmd_ctx *c, long long step);
int fix_pbc_cols(obmd_ctx *c, const ColRange &cr, cudaStream_t st);
ColRange all_columns(const obmd_ctx *c);
ColRange fix_region_columns(const obmd_ctx *c, int r0, int r1, int margin);
int fix_nve_initial(obmd_ctx *c);                               // fix.cu
int fix_nve_final(obmd_ctx *c);
int fix_pbc(obmd_ctx *c);
int fix_pre_begin(obmd_ctx *c);
int fix_pre_exchange(obmd_ctx *c, obmd_trial *trials, int maxtrials);
int fix_post_force(obmd_ctx *c);
int fix_reduce_kinetic(obmd_ctx *c, double out[5]);
int fix_profile_x(obmd_ctx *c, int nbins, double *count, double *mvx, double *mv2, double *sxx);
int pair_gaussian_sample(obmd_ctx *c, int n, int tag0, long long step, double *out);
int fix_clear_extras(obmd_ctx *c);
void bond_free(obmd_ctx *c);  // bond.cu
int read_scalars(obmd_ctx *c);  // copies *sc_d to *sc_h and synchronises the stream
int comm_exchange(obmd_ctx *c);  // comm.cu: migrate + halo + ghost lists
int comm_allreduce(obmd_ctx *c, double *dev, int nsum, int nmax);  // first nsum: sum, next nmax: max
int fix_link_extras(obmd_ctx *c, int first, int count, bool sorted_by_cell);
int fix_compact_extras(obmd_ctx *c);
struct UsherCfg {
  int ntrL, ntr;      // trials for the left region come first (maxattempt == 1)
  int seed0;          // RanPark state before the first trial draw
  int nbase;          // local index of the first inserted atom (= nlocal after deletion)
  int natt;
  int tg0;            // global index of local trial 0 (slab runs: the right rank starts at ntrL)
  int slots, maxatt;  // maxattempt > 1 (one side per pass): insertions wanted, attempts allowed for each
};

int usher_pass(obmd_ctx *c, const UsherCfg &u, int ntr_loc, int limit, int maxtag, int *nacc_out, int *nused_out,
               int *rounds_out);
int usher_commit(obmd_ctx *c, int tag0, int nacc);
int usher_trials_to_host(obmd_ctx *c, obmd_trial *dst, int ncopy);

// ---- small device helpers shared by the TUs --------------------------------------
static inline Extras extras_of(const obmd_ctx *c)
{
  Extras e;
  e.base = c->nsorted;
  e.narr = c->narr;
  e.nghost = c->nghost;
  e.nins = c->nins;
  return e;
}

#ifdef __CUDACC__
// local column of coordinate x (global column clamped to the box, minus this rank's offset)
__device__ __forceinline__ int cell_cx(const Grid &g, double x)
{
  int k = (int) floor((x - g.lo[0]) * g.inv[0]);
  return min(max(k, 0), g.gnx - 1) - g.cxoff;
}
// slot of the q-th extra atom selected by mode (1: mine, 2: all); -1 if q is past the end
__device__ __forceinline__ int extras_slot(const Extras &e, int mode, int q)
{
  if (mode == 2) return q < e.narr + e.nghost + e.nins ? e.base + q : -1;
  if (q < e.narr) return e.base + q;
  q -= e.narr;
  return q < e.nins ? e.base + e.narr + e.nghost + q : -1;
}
__device__ __forceinline__ int cell_wrap(int k, int n)
{
  k %= n;
  return k < 0 ? k + n : k;
}
__device__ __forceinline__ int cell_cy(const Grid &g, double y)
{
  return cell_wrap((int) floor((y - g.lo[1]) * g.inv[1]), g.ny);
}
__device__ __forceinline__ int cell_cz(const Grid &g, double z)
{
  return cell_wrap((int) floor((z - g.lo[2]) * g.inv[2]), g.nz);
}
__device__ __forceinline__ bool block_inside(const obmd_block &b, double x, double y, double z)
{
  return x >= b.lo[0] && x <= b.hi[0] && y >= b.lo[1] && y <= b.hi[1] && z >= b.lo[2] && z <= b.hi[2];
}
#endif
