/* obmd_b200.h -- C ABI of libobmd_b200.so: the B200 (sm_100a) implementation of the
 * OBMD-DPD hot path of chocolatesoup/OBMD-LAMMPS-extension.
 *
 * This is the drop-in boundary.  A LAMMPS style shim (host/lammps_plugin.cpp) or any
 * other host binds exactly these entry points; every one cites the reference
 * interface it replaces (paths relative to /root/reference/code/src).  Plain pointers
 * and sizes only: no C++ or torch types cross this line.  All functions return 0 on
 * success and a negative obmd_status on failure; obmd_last_error() gives the message
 * the shim hands to LAMMPS' error->all() (reference error behaviour:
 * fix_obmd_merged.cpp:72,91,171-191,2160-2163; pair_dpd.cpp:197-205).
 * There is no CPU fallback: without a CUDA device obmd_create() fails.
 *
 * Host arrays use LAMMPS' layout (atom.h): x,v,f = double[n][3] (AoS), type/tag/image =
 * int[n] (LAMMPS_SMALLBIG: tagint = imageint = int32), index = LAMMPS local index.
 * Device arrays are cell-sorted SoA planes; every atom carries its LAMMPS local
 * index so that deletion/insertion bookkeeping and USHER summation order are
 * reproduced exactly.
 */
#ifndef OBMD_B200_H
#define OBMD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define OBMD_MAXTYPES 8
#define OBMD_ABI_VERSION 1

typedef struct obmd_ctx obmd_ctx;

enum obmd_status {
  OBMD_OK = 0,
  OBMD_ERR_CUDA = -1,      /* CUDA runtime error (message has the call site) */
  OBMD_ERR_ARG = -2,       /* illegal argument, mirrors "Illegal fix obmd command" */
  OBMD_ERR_STATE = -3,     /* call order violated (e.g. compute before upload) */
  OBMD_ERR_CAPACITY = -4,  /* a fixed-capacity device buffer overflowed */
  OBMD_ERR_NONFINITE = -5, /* non-numeric coordinates (domain.cpp:540-545) */
  OBMD_ERR_USHER = -6      /* |f| < 1e-6 with E > etarget: the reference spins forever (Q10) */
};

/* region block (region_block.cpp:289-293: closed interval on all three axes) */
typedef struct { double lo[3], hi[3]; } obmd_block;

/* `fix ID group obmd ntype nfreq seed pxx pxy pxz dpxx freq alpha tau nbuf keyword ...`
 * after parsing (fix_obmd_merged.cpp:88-164 positional, :1881-2164 keywords; defaults
 * as in :1883-1939).  ATOM mode, usher, charged 0. */
typedef struct {
  int ntype, nfreq, seed;
  double pxx, pxy, pxz, dpxx, freq, alpha, tau, nbuf;
  obmd_block region[6];   /* region1..6: left/right buffer, left/right shear, left/right insertion */
  double buffer_size, g_fac;
  int step_parallel, step_perp; /* must be 0 / 1 (Q4) */
  int maxattempt;
  int usherflag;
  double etarget, ds0, dtheta0, uovlp, dsovlp, eps;
  int nattempt;
} obmd_fix_args;

/* Restartable fix state (the reference sets restart_global but saves nothing). */
typedef struct {
  int rng_seed;            /* RanPark state (random_park.cpp:43-50) */
  double simulation_time;  /* fix_obmd_merged.cpp:610 */
  double momentumForce_left[3], momentumForce_right[3];
  double shearForce_left[3], shearForce_right[3];
  long long ninserted, ndeleted;
  long long ntrials, nrounds; /* USHER trials run / speculative-commit rounds, cumulative */
} obmd_fix_state;

/* Per-call report of pre_exchange (what the reference prints to stdout, Q8). */
typedef struct {
  int ndel;                 /* atoms deleted (try_deleting, :639-865) */
  int cnt_left, cnt_right;  /* Group::count(igroup, region1/2) (:582-583) */
  int ninsert_left, ninsert_right; /* :586-589 */
  int ntrials, naccept;     /* USHER trials run / atoms created */
  int usher_rounds;         /* speculative-commit rounds used */
  int maxtag;               /* find_maxid before insertion (:1860-1875) */
  int nlocal;               /* atoms after the call */
  double vnewl[3], vnewr[3];
} obmd_pre_exchange_report;

/* One USHER trial (diagnostics / parity tests). */
typedef struct {
  double start[3], end[3];
  double energy;
  int iter, accepted, side, tag;
} obmd_trial;

/* ---- lifetime --------------------------------------------------------------- */
int obmd_abi_version(void);
/* device: CUDA ordinal.  Fails with OBMD_ERR_CUDA when no sm_100 device is usable. */
int obmd_create(obmd_ctx **out, int device);
void obmd_destroy(obmd_ctx *ctx);
const char *obmd_last_error(const obmd_ctx *ctx); /* ctx may be NULL: last create error */

/* ---- domain / atom types (domain.h boxlo/boxhi, atom->mass) ------------------- */
/* boundary f p p, orthogonal.  Needs >= 3 cells of size max(cut) in y and z. */
int obmd_set_box(obmd_ctx *ctx, const double boxlo[3], const double boxhi[3]);
int obmd_set_masses(obmd_ctx *ctx, int ntypes, const double *mass /* [ntypes+1], 1-based */);
int obmd_set_timestep(obmd_ctx *ctx, double dt); /* update->dt */

/* ---- pair_style dpd (DPD-BASIC/pair_dpd.cpp) --------------------------------- */
/* PairDPD::settings (:195-216): pair_style dpd T cut_global seed */
int obmd_pair_settings(obmd_ctx *ctx, double temperature, double cut_global, int seed);
/* PairDPD::coeff (:222-251): pair_coeff itype jtype a0 gamma [cut]; cut <= 0 -> cut_global */
int obmd_pair_coeff(obmd_ctx *ctx, int itype, int jtype, double a0, double gamma, double cut);
/* PairDPD::init_style + init_one (:257-291): sigma = sqrt(2 kB T gamma), symmetrise, build the
 * cell grid.  boltz = force->boltz (1 in lj units). */
int obmd_pair_init(obmd_ctx *ctx, double boltz);
/* PairDPD::single (:401-418), host-side scalar for callers that probe the potential. */
double obmd_pair_single(const obmd_ctx *ctx, int itype, int jtype, double rsq, double *fforce);

/* ---- fix obmd (OBMD/fix_obmd_merged.cpp) ------------------------------------- */
/* constructor + init (:66-287, :321-536): validates args, seeds RanPark(seed) and warms it
 * 30 draws, zeroes simulation_time and the flux forces (Q3). */
int obmd_fix_create(obmd_ctx *ctx, const obmd_fix_args *args);
int obmd_fix_set_flux(obmd_ctx *ctx, double pxx, double pxy, double pxz, double dpxx, double freq,
                      double alpha, double tau, double nbuf); /* v_ variables, :563-572 */
int obmd_fix_get_state(const obmd_ctx *ctx, obmd_fix_state *out);
int obmd_fix_set_state(obmd_ctx *ctx, const obmd_fix_state *in);

/* ---- atoms: host <-> device --------------------------------------------------- */
/* Upload n owned atoms in LAMMPS local order (atom->x, v, type, tag, image; f may be NULL).
 * natoms_capacity: device capacity to reserve (>= n; insertions grow nlocal). */
int obmd_upload(obmd_ctx *ctx, int n, const double *x, const double *v, const double *f,
                const int *type, const int *tag, const int *image, int capacity);
/* Refresh only positions/velocities of the n atoms already resident (same local order). */
int obmd_upload_xv(obmd_ctx *ctx, int n, const double *x, const double *v);
/* Download in LAMMPS local order; any pointer may be NULL.  Returns nlocal in *n. */
int obmd_download(obmd_ctx *ctx, int *n, double *x, double *v, double *f, int *type, int *tag,
                  int *image);
int obmd_nlocal(const obmd_ctx *ctx);

/* ---- per-step hooks, in the order Verlet::run calls them (verlet.cpp:229-363) -- */
/* FixNVE::initial_integrate (fix_nve.cpp:80-145) */
int obmd_nve_initial(obmd_ctx *ctx);
/* NBinStandard::bin_atoms + NPairBin::build role (nbin_standard.cpp:221-261,
 * npair_bin.cpp:52-253): cell index, counting sort of the SoA planes.  Atoms outside the
 * box in x are parked behind the live range for pre_exchange to delete. */
int obmd_build_cells(obmd_ctx *ctx);
/* FixObmdMerged::pre_exchange (:550-633): delete, count, USHER insert, flux forces.
 * Requires obmd_build_cells() on the current positions.  report/trials may be NULL. */
int obmd_pre_exchange(obmd_ctx *ctx, obmd_pre_exchange_report *report, obmd_trial *trials,
                      int maxtrials);
/* Domain::pbc (domain.cpp:529-640) for boundary f p p: wrap y,z and update image flags. */
int obmd_pbc(obmd_ctx *ctx);
/* Verlet::force_clear + PairDPD::compute (:66-163) (+ virial_fdotr_compute, pair.cpp:1735).
 * ntimestep keys the counter RNG.  eflag/vflag != 0 also reduce eng_vdwl / virial[6]. */
int obmd_pair_compute(obmd_ctx *ctx, long long ntimestep, int eflag, int vflag);
int obmd_pair_tallies(obmd_ctx *ctx, double *eng_vdwl, double virial[6]);
/* FixObmdMerged::post_force (:1229-1240): reg_force x2 + reg_force_perp x2 */
int obmd_post_force(obmd_ctx *ctx);
/* Verlet::force_clear on the device copy (lets a host shim read back only what post_force added). */
int obmd_zero_forces(obmd_ctx *ctx);
/* FixNVE::final_integrate (fix_nve.cpp:185-244) */
int obmd_nve_final(obmd_ctx *ctx);
/* Verlet::setup (verlet.cpp:92-163): [exchange, borders,] neighbour build, pair forces,
 * fix setup -> post_force.  ntimestep keys the counter RNG. */
int obmd_setup(obmd_ctx *ctx, long long ntimestep);
/* nsteps full timesteps, device resident, in exactly the order above.  first_step = the
 * timestep number of the first step (update->ntimestep + 1). */
int obmd_run(obmd_ctx *ctx, long long first_step, int nsteps);

/* ---- observables (compute temp / ke, used by the statistical parity checks) ---- */
/* out[0] = sum m v^2 (2*KE), out[1..3] = total momentum, out[4] = nlocal */
int obmd_reduce_kinetic(obmd_ctx *ctx, double out[5]);
/* 1-D profiles along x in nbins bins over the box: count[nbins], vx sums [nbins] */
int obmd_profile_x(obmd_ctx *ctx, int nbins, double *count, double *mvx);
/* The profiles `compute chunk/atom bin/1d x` + `fix ave/chunk ... density/number vx temp c_stress[1]`
 * accumulate (compute_chunk_atom.cpp, fix_ave_chunk.cpp, compute_stress_atom.cpp): per bin the atom
 * count, sum m vx, sum m v^2 and the sum of the per-atom stress xx = -(m vx^2 + W_xx,i), where W_xx,i
 * is the atom's share of the pair virial as Pair::ev_tally assigns it (pair.cpp:1068-1230).  mv2 and
 * sxx may be NULL; sxx needs obmd_pair_compute(..., vflag != 0) on the current positions first. */
int obmd_profile_x2(obmd_ctx *ctx, int nbins, double *count, double *mvx, double *mv2, double *sxx);
/* Diagnostic for the statistical parity tests: out[k] = the N(0,1) variate PairDPD uses for the pair
 * (tag0 + k, tag0 + k + 1) at timestep ntimestep (stands in for RanMars::gaussian, random_mars.cpp:103). */
int obmd_pair_gaussian_sample(obmd_ctx *ctx, int n, int tag0, long long ntimestep, double *out);

/* ---- multi-GPU: x-slab decomposition over NCCL (replaces CommBrick::exchange/borders, ------
 *      comm_brick.cpp:652-978, and the fix's MPI_Allreduce/Scan calls) ------------ */
/* One process per GPU.  Rank 0 creates the id (ncclGetUniqueId), the host broadcasts the 128 bytes
 * (MPI in LAMMPS, torch.distributed in bench.py), every rank calls obmd_comm_init before
 * obmd_upload.  Rank r then owns a contiguous range of cell columns along x and uploads only the
 * atoms inside it; obmd_run / obmd_setup do the migration, the halo and the packed allreduces.
 * The insertion regions (region5/6) must lie inside the first / last rank's slab. */
int obmd_comm_unique_id(char *out128);
int obmd_comm_init(obmd_ctx *ctx, int rank, int nranks, const char *id128);
int obmd_comm_finalize(obmd_ctx *ctx);
/* The slab partition itself (pure host arithmetic; no device or context needed): owned cell columns
 * [c0, c1) of `rank`.  edge_cols = columns an edge rank must hold (its insertion region + one cutoff), 0
 * for an even split.  Replaces the processor grid of comm_brick.cpp / procmap.cpp for one decomposed axis. */
int obmd_slab_partition(int ncols, int nranks, int rank, int edge_cols, int *c0, int *c1);
/* x-range [lo, hi) of the columns this rank owns (what the host uses to split the atoms). */
int obmd_comm_slab(const obmd_ctx *ctx, double *lo, double *hi);
/* Comm::exchange + Comm::borders (comm_brick.cpp:652-978) for callers that drive the per-step hooks
 * themselves in a slab run: after obmd_build_cells, before obmd_pre_exchange (obmd_run does it itself).
 * No-op on one rank. */
int obmd_exchange(obmd_ctx *ctx);
/* Order in which USHER sums the contributions to a trial point's energy and force (energy(),
 * fix_obmd_merged.cpp:1790-1807).  0 (default, one rank): ascending LAMMPS local index -- the
 * reference's loop order, bit-identical to the serial reference.  1: ascending atom tag -- independent
 * of the decomposition; slab runs always use it, so a slab run on any number of ranks reproduces a
 * one-rank run with by_tag = 1 bit for bit as long as both see the same atoms (the reference's own
 * multi-rank sums differ from its serial ones in the same way: per-rank partial sums + MPI_Allreduce). */
int obmd_set_usher_order(obmd_ctx *ctx, int by_tag);

/* ---- bonded forces (BASELINE config 4, the bonded path only) ----------------------- */
/* bond_style harmonic + bond_coeff t K r0 (MOLECULE/bond_harmonic.cpp:126-149): k and r0 are 1-based,
 * nbondtypes + 1 entries each (entry 0 unused), nbondtypes <= OBMD_MAXTYPES. */
int obmd_bond_style_harmonic(obmd_ctx *ctx, int nbondtypes, const double *k, const double *r0);
/* bond_style fene + bond_coeff t K R0 epsilon sigma (MOLECULE/bond_fene.cpp:53-135, :150-175); a bond
 * longer than 2 R0 makes obmd_bond_compute fail with "Bad FENE bond" (:94). */
int obmd_bond_style_fene(obmd_ctx *ctx, int nbondtypes, const double *k, const double *r0,
                         const double *epsilon, const double *sigma);
/* The bond topology by atom tags, one entry per bond (Atom::bond_type / bond_atom as read_data leaves
 * them with newton_bond on; NTopoBondAll::build, ntopo_bond_all.cpp:36-77, turns them into the bond list
 * every step through Atom::map and Domain::closest_image).  The topology is static: molecule-aware
 * deletion / insertion (fix_obmd_merged.cpp:709-821, :993-1027, :1132-1168) is not implemented, and
 * obmd_bond_compute reports an error when a bonded atom has been deleted. */
int obmd_upload_bonds(obmd_ctx *ctx, int nbonds, const int *type, const int *tag1, const int *tag2);
/* BondHarmonic::compute (MOLECULE/bond_harmonic.cpp:48-103): adds the bond forces to the device forces
 * (call after obmd_pair_compute, which overwrites them).  eflag != 0 also reduces the bond energy and
 * the six virial components (Bond::ev_tally, bond.cpp:174-240; order xx yy zz xy xz yz). */
int obmd_bond_compute(obmd_ctx *ctx, int eflag);
int obmd_bond_tallies(obmd_ctx *ctx, double *ebond, double virial[6] /* may be NULL */);

/* ---- timing / introspection ----------------------------------------------------- */
/* Accumulated CUDA-event milliseconds and launch counts per kernel family since the last
 * reset.  names: "bin", "sort", "pair", "post_force", "nve", "pre_exchange", "usher", "pbc". */
int obmd_timing_enable(obmd_ctx *ctx, int on);
int obmd_timing_get(obmd_ctx *ctx, const char *name, double *ms, long long *launches);
int obmd_timing_reset(obmd_ctx *ctx);
long long obmd_launch_count(const obmd_ctx *ctx); /* kernels launched since create */
void *obmd_stream(const obmd_ctx *ctx);            /* cudaStream_t the kernels run on */

#ifdef __cplusplus
}
#endif
#endif
