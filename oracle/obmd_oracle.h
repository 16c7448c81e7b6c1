/* TEST INFRASTRUCTURE -- CPU oracle for the OBMD-DPD hot path.  NOT the product.
 *
 * A plain-C, single-thread restatement of the reference algorithm
 * (chocolatesoup/OBMD-LAMMPS-extension: LAMMPS 7Feb2024 + src/OBMD).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (obmd-lammps-extension_b200/)
 * never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * against outputs of the unmodified reference binary (oracle/_ref/lmp_serial,
 * built by oracle/build_ref.sh) stored under tests/golden/ by
 * tests/golden/make_golden.py: RanPark trial stream + USHER verdicts and accepted
 * positions (bit-exact), deletion/insertion bookkeeping and local order
 * (bit-exact), buffer force distribution (bit-exact), deterministic DPD forces.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/code/src unless noted).
 */
#ifndef OBMD_ORACLE_H
#define OBMD_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXTYPES 8

/* LAMMPS_SMALLBIG image flags: 10 bits per dimension (lmptype.h:90-116). */
#define ORC_IMGMASK 1023
#define ORC_IMGMAX 512
#define ORC_IMGBITS 10
#define ORC_IMG2BITS 20

typedef struct { double lo[3], hi[3]; } orc_block; /* region block, closed (region_block.cpp:289) */

/* Host atom arrays in LAMMPS layout: AoS double[N][3], int per-atom vectors. */
typedef struct {
  int nlocal;   /* owned atoms */
  int nmax;     /* capacity of the arrays below */
  long natoms;  /* atom->natoms */
  double *x, *v, *f; /* [nmax*3] */
  int *type, *tag, *image;
} orc_atoms;

typedef struct {
  /* domain: boundary f p p, orthogonal */
  double boxlo[3], boxhi[3];
  int ntypes;
  double mass[ORC_MAXTYPES + 1];
  /* pair_style dpd T rc seed ; pair_coeff i j a0 gamma rc (pair_dpd.cpp:195-291) */
  double temperature, cut_global, boltz;
  int pair_seed;
  double a0[ORC_MAXTYPES + 1][ORC_MAXTYPES + 1];
  double gamma[ORC_MAXTYPES + 1][ORC_MAXTYPES + 1];
  double sigma[ORC_MAXTYPES + 1][ORC_MAXTYPES + 1];
  double cut[ORC_MAXTYPES + 1][ORC_MAXTYPES + 1];
  double cutsq[ORC_MAXTYPES + 1][ORC_MAXTYPES + 1];
  double dt;
  /* fix obmd arguments (fix_obmd_merged.cpp:88-164, 1881-2164) */
  int ntype, nfreq, seed;
  double pxx, pxy, pxz, dpxx, freq, alpha, tau, nbuf;
  orc_block region[6]; /* region1..6: leftB rightB leftshear rightshear insertL insertR */
  double buffer_size, g_fac;
  int maxattempt;
  int usherflag;
  double etarget, ds0, dtheta0, uovlp, dsovlp, eps;
  int nattempt;
  /* fix state */
  int rng_seed;           /* RanPark state (random_park.cpp:24-50) */
  double simulation_time; /* Q3: explicitly 0 at construction */
  double momentumForce_left[3], momentumForce_right[3];
  double shearForce_left[3], shearForce_right[3];
  double vnewl[3], vnewr[3];
  long ninserted, ndeleted;
  int maxtag_all;
  /* last-call diagnostics */
  int last_cnt_left, last_cnt_right, last_ninsert_left, last_ninsert_right;
  int last_ndel, last_ntrials, last_naccept;
  int use_cells; /* 1: USHER energy through a cell list (same sums, same order) */
} orc_fix;

/* One USHER trial record (for bit-exact comparison of the trial stream). */
typedef struct {
  double start[3];  /* trial start drawn from RanPark (fix_obmd_merged.cpp:923-927) */
  double end[3];    /* coords when usher() returned */
  double energy;    /* value returned by usher() */
  int iter;         /* iterations */
  int accepted;     /* 1 if the atom was created */
  int side;         /* 0 = region5 (left), 1 = region6 (right) */
  int tag;          /* tag given to the new atom (0 if none) */
} orc_trial;

/* ---- RNGs ---- */
double orc_ranpark_uniform(int *seed);                 /* random_park.cpp:43-50 */
int orc_ranpark_skip(int seed, long n);                /* n calls of uniform() at once */
typedef struct { double u[98]; int i97, j97; double c, cd, cm; int save; double second; } orc_ranmars;
void orc_ranmars_init(orc_ranmars *r, int seed);       /* random_mars.cpp:30-69 */
double orc_ranmars_uniform(orc_ranmars *r);            /* random_mars.cpp:82-98 */
double orc_ranmars_gaussian(orc_ranmars *r);           /* random_mars.cpp:103-122 */

/* ---- pair dpd ---- */
void orc_pair_init(orc_fix *p);                        /* init_one: sigma, cutsq (pair_dpd.cpp:279-291) */
double orc_pair_single(const orc_fix *p, int itype, int jtype, double rsq, double *fforce); /* :401-418 */
/* Symmetric counter-based N(0,1) variate the CUDA kernel uses in place of the
   sequential RanMars stream (pair_dpd.cpp:122).  NOT in the reference: a
   design substitution, statistical parity only (DESIGN.md "RNG"). */
double orc_pair_gaussian(int tagi, int tagj, unsigned step, unsigned seed);
/* Full-neighbour DPD forces over owned atoms, y/z periodic by minimum image, x open.
   rng_mode 0: no random term (T=0 parity); 1: orc_pair_gaussian.  Writes f (overwrites),
   returns eng_vdwl in *eng and the 6 virial components in virial (pair.cpp:1735 convention). */
void orc_pair_compute(const orc_fix *p, orc_atoms *a, int rng_mode, unsigned step,
                      double *eng, double *virial);

/* ---- fix obmd ---- */
void orc_fix_init(orc_fix *p);                         /* ctor: RanPark(seed)+30 draws (:275-276) */
void orc_fix_pre_exchange(orc_fix *p, orc_atoms *a, orc_trial *trials, int maxtrials); /* :550-633 */
void orc_fix_post_force(const orc_fix *p, orc_atoms *a);  /* :1229-1240 */
double orc_g_par_local(const orc_fix *p, double mass, double xv); /* :1312-1340 */
/* USHER on one point (ATOM mode): returns energy, updates coord, *iter. (:1518-1616) */
double orc_usher(const orc_fix *p, const orc_atoms *a, const orc_block *reg, double *coord, int *iter);
double orc_energy(const orc_fix *p, const orc_atoms *a, int itype, const double *coord, double *fusher); /* :1774-1810 */

/* ---- bonded forces (BASELINE config 4, SURVEY 8f N3: the bonded path only) ---- */
/* BondHarmonic::compute (MOLECULE/bond_harmonic.cpp:48-103) over a bond list given by atom tags
   (NTopoBondAll::build, ntopo_bond_all.cpp:36-77: atom->map + Domain::closest_image, here the minimum
   image in y and z).  k, r0: 1-based per bond type.  Adds to a->f, returns the bond energy. */
double orc_bond_harmonic(const orc_fix *p, orc_atoms *a, int nbonds, const int *btype, const int *tag1,
                         const int *tag2, const double *k, const double *r0);

/* BondFENE::compute (MOLECULE/bond_fene.cpp:53-135), same bond list; k, r0, epsilon, sigma 1-based.
   Returns the bond energy; -1e300 if a bond is longer than 2 r0 ("Bad FENE bond"). */
double orc_bond_fene(const orc_fix *p, orc_atoms *a, int nbonds, const int *btype, const int *tag1,
                     const int *tag2, const double *k, const double *r0, const double *epsilon,
                     const double *sigma);

/* ---- integrator / geometry ---- */
void orc_nve_initial(const orc_fix *p, orc_atoms *a);  /* fix_nve.cpp:80-145 */
void orc_nve_final(const orc_fix *p, orc_atoms *a);    /* fix_nve.cpp:185-244 */
void orc_pbc(const orc_fix *p, orc_atoms *a);          /* domain.cpp:529-640 (y,z only) */
/* One Verlet step in the reference's order (verlet.cpp:229-363). */
void orc_step(orc_fix *p, orc_atoms *a, int rng_mode, unsigned step, orc_trial *trials, int maxtrials);

#ifdef __cplusplus
}
#endif
#endif
