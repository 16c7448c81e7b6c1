// LAMMPS style shim: registers `fix obmd` and `pair_style dpd` from a plugin and forwards them to
// libobmd_b200.so through the C ABI (include/obmd_b200.h).  Input-script syntax, hook points and
// coefficients are the reference's (OBMD/fix_obmd_merged.{h,cpp}, DPD-BASIC/pair_dpd.{h,cpp}); this
// file only marshals arguments and LAMMPS-owned host pointers.  Two routes (DESIGN.md / INTEGRATION.md):
//   stage A  `run N`: LAMMPS' own Verlet drives the hooks; FixNVE, Domain::pbc and Comm stay on the host,
//            so every step uploads x,v (48 B/atom) and downloads f (24 B/atom);
//   stage B  `run_obmd N`: a command style registered by the same plugin runs the whole timestep
//            device-resident (obmd_run, the hook sequence of Verlet::run, verlet.cpp:229-363) and
//            brings the atoms back to the host only on thermo / dump steps and at the end.
//
// Build (needs the LAMMPS headers, i.e. only where the reference tree is mounted):
//   make -C obmd-lammps-extension_b200/host plugin LAMMPS_SRC=/root/reference/code/src
// Use:  first line of the deck:  plugin load /path/to/libobmd_b200_lammps.so
// The classes are deliberately NOT named FixObmdMerged / PairDPD: the host binary exports its own
// symbols (ENABLE_EXPORTS) and would interpose them (SURVEY 8b "Gotcha").
#include "atom.h"
#include "atom_vec.h"
#include "bond.h"
#include "comm.h"
#include "command.h"
#include "domain.h"
#include "error.h"
#include "finish.h"
#include "fix.h"
#include "force.h"
#include "group.h"
#include "integrate.h"
#include "lammpsplugin.h"
#include "memory.h"
#include "modify.h"
#include "output.h"
#include "pair.h"
#include "region.h"
#include "timer.h"
#include "update.h"
#include "version.h"

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "obmd_b200.h"

using namespace LAMMPS_NS;
using namespace FixConst;

namespace {

// One device context per LAMMPS instance, shared by the pair style and the fix.
struct Shared {
  obmd_ctx *ctx = nullptr;
  int users = 0;
  bool box_set = false, pair_ready = false;
  int resident = 0;       // atoms resident on the device (LAMMPS local order)
  bigint xv_step = -1;    // timestep whose x,v are on the device
};
std::map<LAMMPS *, Shared> g_shared;

Shared &acquire(LAMMPS *lmp)
{
  Shared &s = g_shared[lmp];
  if (!s.ctx) {
    int dev = 0;
    if (const char *e = getenv("OBMD_B200_DEVICE")) dev = atoi(e);
    if (obmd_create(&s.ctx, dev) != OBMD_OK)
      lmp->error->all(FLERR, "obmd_b200: {}", obmd_last_error(nullptr));
  }
  s.users++;
  return s;
}

void release(LAMMPS *lmp)
{
  auto it = g_shared.find(lmp);
  if (it == g_shared.end()) return;
  if (--it->second.users <= 0) {
    obmd_destroy(it->second.ctx);
    g_shared.erase(it);
  }
}

#define OBMD_CK(lmp, s, call)                                                        \
  do {                                                                               \
    if ((call) != OBMD_OK) (lmp)->error->all(FLERR, "{}", obmd_last_error((s).ctx)); \
  } while (0)

void push_domain(LAMMPS *lmp, Shared &s)
{
  Domain *d = lmp->domain;
  if (d->triclinic) lmp->error->all(FLERR, "obmd_b200 requires an orthogonal box");
  if (d->xperiodic || !d->yperiodic || !d->zperiodic)
    lmp->error->all(FLERR, "obmd_b200 requires boundary f p p");
  if (lmp->comm->nprocs > 1)
    lmp->error->all(FLERR, "obmd_b200 plugin (stage A) runs on one MPI rank; use obmd_run for slabs");
  OBMD_CK(lmp, s, obmd_set_box(s.ctx, d->boxlo, d->boxhi));
  OBMD_CK(lmp, s, obmd_set_masses(s.ctx, lmp->atom->ntypes, lmp->atom->mass));
  OBMD_CK(lmp, s, obmd_set_timestep(s.ctx, lmp->update->dt));
  s.box_set = true;
}

void upload_all(LAMMPS *lmp, Shared &s)
{
  Atom *a = lmp->atom;
  OBMD_CK(lmp, s, obmd_upload(s.ctx, a->nlocal, &a->x[0][0], &a->v[0][0], nullptr, a->type, a->tag,
                              a->image, a->nmax));
  s.resident = a->nlocal;
  s.xv_step = lmp->update->ntimestep;
}

// ------------------------------------------------------------------------------------------
// pair_style dpd T rc seed ; pair_coeff i j a0 gamma [rc]      (DPD-BASIC/pair_dpd.cpp)
// ------------------------------------------------------------------------------------------
class PairDPDB200 : public Pair {
 public:
  PairDPDB200(LAMMPS *lmp) : Pair(lmp), sh(acquire(lmp))
  {
    writedata = 1;
    no_virial_fdotr_compute = 1;  // the device returns the virial itself
  }
  ~PairDPDB200() override
  {
    if (allocated) {
      memory->destroy(setflag);
      memory->destroy(cutsq);
      memory->destroy(cut);
      memory->destroy(a0);
      memory->destroy(gamma);
    }
    release(lmp);
  }

  void settings(int narg, char **arg) override  // pair_dpd.cpp:195-216
  {
    if (narg != 3) error->all(FLERR, "Illegal pair_style command");
    temperature = utils::numeric(FLERR, arg[0], false, lmp);
    cut_global = utils::numeric(FLERR, arg[1], false, lmp);
    seed = utils::inumeric(FLERR, arg[2], false, lmp);
    if (seed <= 0) error->all(FLERR, "Illegal pair_style command");
    if (allocated)
      for (int i = 1; i <= atom->ntypes; i++)
        for (int j = i; j <= atom->ntypes; j++)
          if (setflag[i][j]) cut[i][j] = cut_global;
  }

  void coeff(int narg, char **arg) override  // pair_dpd.cpp:222-251
  {
    if (narg < 4 || narg > 5) error->all(FLERR, "Incorrect args for pair coefficients");
    if (!allocated) allocate();
    int ilo, ihi, jlo, jhi;
    utils::bounds(FLERR, arg[0], 1, atom->ntypes, ilo, ihi, error);
    utils::bounds(FLERR, arg[1], 1, atom->ntypes, jlo, jhi, error);
    double a0_one = utils::numeric(FLERR, arg[2], false, lmp);
    double gamma_one = utils::numeric(FLERR, arg[3], false, lmp);
    double cut_one = cut_global;
    if (narg == 5) cut_one = utils::numeric(FLERR, arg[4], false, lmp);
    int count = 0;
    for (int i = ilo; i <= ihi; i++)
      for (int j = MAX(jlo, i); j <= jhi; j++) {
        a0[i][j] = a0_one;
        gamma[i][j] = gamma_one;
        cut[i][j] = cut_one;
        setflag[i][j] = 1;
        count++;
      }
    if (count == 0) error->all(FLERR, "Incorrect args for pair coefficients");
  }

  void init_style() override  // pair_dpd.cpp:257-273, but no neighbour request: cells live on the GPU
  {
    if (force->newton_pair == 0 && comm->me == 0)
      error->warning(FLERR, "Pair dpd needs newton pair on for momentum conservation");
    // The device mirrors atoms by LAMMPS local index and refreshes only x and v while nlocal is
    // unchanged: a spatial re-sort (Atom::sort, atom.cpp:2288-2378) would silently permute type and
    // tag under it.  The reference decks all run with sorting off; refuse anything else.
    if (atom->sortfreq > 0)
      error->all(FLERR, "pair_style dpd (obmd_b200 plugin) requires 'atom_modify sort 0 0'");
    push_domain(lmp, sh);
    OBMD_CK(lmp, sh, obmd_pair_settings(sh.ctx, temperature, cut_global, seed));
    for (int i = 1; i <= atom->ntypes; i++)
      for (int j = i; j <= atom->ntypes; j++) {
        if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
        OBMD_CK(lmp, sh, obmd_pair_coeff(sh.ctx, i, j, a0[i][j], gamma[i][j], cut[i][j]));
      }
    OBMD_CK(lmp, sh, obmd_pair_init(sh.ctx, force->boltz));
    sh.pair_ready = true;
  }

  double init_one(int i, int j) override  // pair_dpd.cpp:279-291
  {
    if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
    cut[j][i] = cut[i][j];
    a0[j][i] = a0[i][j];
    gamma[j][i] = gamma[i][j];
    return cut[i][j];
  }

  void compute(int eflag, int vflag) override  // pair_dpd.cpp:66-163
  {
    ev_init(eflag, vflag);
    Atom *a = atom;
    OBMD_CK(lmp, sh, obmd_set_timestep(sh.ctx, update->dt));
    if (sh.resident != a->nlocal) upload_all(lmp, sh);
    else OBMD_CK(lmp, sh, obmd_upload_xv(sh.ctx, a->nlocal, &a->x[0][0], &a->v[0][0]));
    OBMD_CK(lmp, sh, obmd_build_cells(sh.ctx));
    OBMD_CK(lmp, sh, obmd_pair_compute(sh.ctx, update->ntimestep, eflag_either, vflag_either));
    int n = 0;
    if ((int) fbuf.size() < 3 * a->nmax) fbuf.resize(3 * (size_t) a->nmax);
    OBMD_CK(lmp, sh, obmd_download(sh.ctx, &n, nullptr, nullptr, fbuf.data(), nullptr, nullptr, nullptr));
    double **f = a->f;
    for (int i = 0; i < n; i++) {
      f[i][0] += fbuf[3 * i];
      f[i][1] += fbuf[3 * i + 1];
      f[i][2] += fbuf[3 * i + 2];
    }
    if (eflag_either || vflag_either) {
      double e, v[6];
      OBMD_CK(lmp, sh, obmd_pair_tallies(sh.ctx, &e, v));
      if (eflag_global) eng_vdwl += e;
      if (vflag_global)
        for (int k = 0; k < 6; k++) virial[k] += v[k];
    }
  }

  double single(int, int, int itype, int jtype, double rsq, double, double factor_dpd,
                double &fforce) override  // pair_dpd.cpp:401-418
  {
    double ff = 0.0;
    double phi = obmd_pair_single(sh.ctx, itype, jtype, rsq, &ff);
    fforce = ff * factor_dpd;
    return factor_dpd * phi;
  }

 private:
  void allocate()
  {
    allocated = 1;
    int n = atom->ntypes;
    memory->create(setflag, n + 1, n + 1, "pair:setflag");
    for (int i = 1; i <= n; i++)
      for (int j = i; j <= n; j++) setflag[i][j] = 0;
    memory->create(cutsq, n + 1, n + 1, "pair:cutsq");
    memory->create(cut, n + 1, n + 1, "pair:cut");
    memory->create(a0, n + 1, n + 1, "pair:a0");
    memory->create(gamma, n + 1, n + 1, "pair:gamma");
  }
  Shared &sh;
  double cut_global = 0.0, temperature = 0.0;
  int seed = 0;
  double **cut = nullptr, **a0 = nullptr, **gamma = nullptr;
  std::vector<double> fbuf;
};

// ------------------------------------------------------------------------------------------
// bond_style harmonic ; bond_coeff t K r0                         (MOLECULE/bond_harmonic.cpp)
// Stage A: the pair style has put this step's atoms on the device; the bond list NTopoBondAll::build
// (ntopo_bond_all.cpp:36-77) would walk is handed over as tags once (static topology).
// ------------------------------------------------------------------------------------------
class BondHarmonicB200 : public Bond {
 public:
  BondHarmonicB200(LAMMPS *lmp) : Bond(lmp), sh(acquire(lmp)) {}
  ~BondHarmonicB200() override
  {
    if (allocated && !copymode) {
      memory->destroy(setflag);
      memory->destroy(k);
      memory->destroy(r0);
    }
    release(lmp);
  }

  void compute(int eflag, int vflag) override  // bond_harmonic.cpp:48-103
  {
    ev_init(eflag, vflag);
    Atom *a = atom;
    if (eflag_atom || vflag_atom) error->all(FLERR, "bond_style harmonic (obmd_b200 plugin): per-atom tallies are not available");
    if (sh.resident != a->nlocal)
      error->all(FLERR, "bond_style harmonic (obmd_b200 plugin) needs pair_style dpd from the same plugin (it uploads the atoms)");
    if (topo_bonds != a->nbonds || topo_step < 0) upload_topology();
    OBMD_CK(lmp, sh, obmd_zero_forces(sh.ctx));
    OBMD_CK(lmp, sh, obmd_bond_compute(sh.ctx, (eflag_either || vflag_either) ? 1 : 0));
    int n = 0;
    if ((int) fbuf.size() < 3 * a->nmax) fbuf.resize(3 * (size_t) a->nmax);
    OBMD_CK(lmp, sh, obmd_download(sh.ctx, &n, nullptr, nullptr, fbuf.data(), nullptr, nullptr, nullptr));
    double **f = a->f;
    for (int i = 0; i < n; i++) {
      f[i][0] += fbuf[3 * i];
      f[i][1] += fbuf[3 * i + 1];
      f[i][2] += fbuf[3 * i + 2];
    }
    if (eflag_either || vflag_either) {
      double e, v[6];
      OBMD_CK(lmp, sh, obmd_bond_tallies(sh.ctx, &e, v));
      if (eflag_global) energy += e;
      if (vflag_global)
        for (int q = 0; q < 6; q++) virial[q] += v[q];
    }
  }

  void coeff(int narg, char **arg) override  // bond_harmonic.cpp:123-143
  {
    if (narg != 3) error->all(FLERR, "Incorrect args for bond coefficients");
    if (!allocated) allocate();
    int ilo, ihi;
    utils::bounds(FLERR, arg[0], 1, atom->nbondtypes, ilo, ihi, error);
    const double k_one = utils::numeric(FLERR, arg[1], false, lmp);
    const double r0_one = utils::numeric(FLERR, arg[2], false, lmp);
    int count = 0;
    for (int i = ilo; i <= ihi; i++) {
      k[i] = k_one;
      r0[i] = r0_one;
      setflag[i] = 1;
      count++;
    }
    if (count == 0) error->all(FLERR, "Incorrect args for bond coefficients");
    topo_step = -1;  // coefficients go to the device with the next topology upload
  }

  double equilibrium_distance(int i) override { return r0[i]; }  // :149-152

  void write_restart(FILE *fp) override  // :158-162
  {
    fwrite(&k[1], sizeof(double), atom->nbondtypes, fp);
    fwrite(&r0[1], sizeof(double), atom->nbondtypes, fp);
  }

  void read_restart(FILE *fp) override  // :168-180 (one rank)
  {
    allocate();
    utils::sfread(FLERR, &k[1], sizeof(double), atom->nbondtypes, fp, nullptr, error);
    utils::sfread(FLERR, &r0[1], sizeof(double), atom->nbondtypes, fp, nullptr, error);
    for (int i = 1; i <= atom->nbondtypes; i++) setflag[i] = 1;
  }

  void write_data(FILE *fp) override  // :186-189
  {
    for (int i = 1; i <= atom->nbondtypes; i++) fprintf(fp, "%d %g %g\n", i, k[i], r0[i]);
  }

  double single(int type, double rsq, int, int, double &fforce) override  // :193-201
  {
    const double r = sqrt(rsq);
    const double dr = r - r0[type];
    const double rk = k[type] * dr;
    fforce = 0;
    if (r > 0.0) fforce = -2.0 * rk / r;
    return rk * dr;
  }

 private:
  void allocate()
  {
    allocated = 1;
    const int np1 = atom->nbondtypes + 1;
    memory->create(k, np1, "bond:k");
    memory->create(r0, np1, "bond:r0");
    memory->create(setflag, np1, "bond:setflag");
    for (int i = 1; i < np1; i++) setflag[i] = 0;
  }

  // Atom::bond_type / bond_atom (newton_bond on: every bond stored once) -> (type, tag, tag)
  void upload_topology()
  {
    Atom *a = atom;
    if (a->nbondtypes > OBMD_MAXTYPES) error->all(FLERR, "obmd_b200: too many bond types");
    for (int i = 1; i <= a->nbondtypes; i++)
      if (!setflag[i]) error->all(FLERR, "All bond coeffs are not set");
    OBMD_CK(lmp, sh, obmd_bond_style_harmonic(sh.ctx, a->nbondtypes, k, r0));
    std::vector<int> bt, t1, t2;
    for (int i = 0; i < a->nlocal; i++)
      for (int m = 0; m < a->num_bond[i]; m++) {
        if (a->bond_type[i][m] <= 0) continue;
        bt.push_back(a->bond_type[i][m]);
        t1.push_back(a->tag[i]);
        t2.push_back(a->bond_atom[i][m]);
      }
    OBMD_CK(lmp, sh, obmd_upload_bonds(sh.ctx, (int) bt.size(), bt.data(), t1.data(), t2.data()));
    topo_bonds = a->nbonds;
    topo_step = update->ntimestep;
  }

  Shared &sh;
  double *k = nullptr, *r0 = nullptr;
  bigint topo_bonds = -1, topo_step = -1;
  std::vector<double> fbuf;
};

// ------------------------------------------------------------------------------------------
// fix ID group obmd ntype nfreq seed pxx pxy pxz dpxx freq alpha tau nbuf keyword value ...
// (OBMD/fix_obmd_merged.cpp:66-287 constructor, :1881-2164 options)
// ------------------------------------------------------------------------------------------
class FixObmdB200 : public Fix {
 public:
  FixObmdB200(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg), sh(acquire(lmp))
  {
    if (narg < 14) error->all(FLERR, "Illegal fix obmd command (check # of arguments)");
    if (igroup != 0) error->all(FLERR, "obmd_b200: fix obmd supports group all only");
    restart_global = 1;
    time_depend = 1;
    memset(&fa, 0, sizeof(fa));
    fa.ntype = utils::inumeric(FLERR, arg[3], false, lmp);
    fa.nfreq = utils::inumeric(FLERR, arg[4], false, lmp);
    fa.seed = utils::inumeric(FLERR, arg[5], false, lmp);
    if (fa.seed <= 0) error->all(FLERR, "Illegal fix obmd command");
    double *dst[8] = {&fa.pxx, &fa.pxy, &fa.pxz, &fa.dpxx, &fa.freq, &fa.alpha, &fa.tau, &fa.nbuf};
    for (int k = 0; k < 8; k++) {
      if (strstr(arg[6 + k], "v_") == arg[6 + k]) varname[k] = &arg[6 + k][2];
      else *dst[k] = utils::numeric(FLERR, arg[6 + k], false, lmp);
    }
    options(narg - 14, &arg[14]);
    if (fa.ntype <= 0 || fa.ntype > atom->ntypes)
      error->all(FLERR, "Invalid atom type in fix obmd command");
    if (atom->tag_enable == 0) error->all(FLERR, "Cannot use fix obmd unless atoms have IDs");
    force_reneighbor = 1;  // :279-281
    next_reneighbor = update->ntimestep + 1;
  }
  ~FixObmdB200() override { release(lmp); }

  int setmask() override  // :311-317
  {
    int mask = 0;
    mask |= PRE_EXCHANGE;
    mask |= POST_FORCE;
    return mask;
  }

  void init() override  // :321-536 (region lookup; variables are resolved each step)
  {
    for (int r = 0; r < 6; r++) {
      Region *reg = domain->get_region_by_id(idregion[r]);
      if (!reg) error->all(FLERR, "Region ID for fix obmd does not exist");
      if (reg->bboxflag == 0) error->all(FLERR, "Fix obmd region does not support a bounding box");
      fa.region[r].lo[0] = reg->extent_xlo; fa.region[r].hi[0] = reg->extent_xhi;
      fa.region[r].lo[1] = reg->extent_ylo; fa.region[r].hi[1] = reg->extent_yhi;
      fa.region[r].lo[2] = reg->extent_zlo; fa.region[r].hi[2] = reg->extent_zhi;
    }
    if (!sh.box_set) push_domain(lmp, sh);
    if (!created) {
      OBMD_CK(lmp, sh, obmd_fix_create(sh.ctx, &fa));
      created = true;
    }
    if (have_pending_state) {
      OBMD_CK(lmp, sh, obmd_fix_set_state(sh.ctx, &pending_state));
      have_pending_state = false;
    }
  }

  void setup(int vflag) override  // :542-545
  {
    if (utils::strmatch(update->integrate_style, "^verlet")) post_force(vflag);
  }

  void pre_exchange() override  // :550-633
  {
    refresh_variables();
    Atom *a = atom;
    OBMD_CK(lmp, sh, obmd_set_timestep(sh.ctx, update->dt));
    upload_all(lmp, sh);
    OBMD_CK(lmp, sh, obmd_build_cells(sh.ctx));
    obmd_pre_exchange_report rep;
    OBMD_CK(lmp, sh, obmd_pre_exchange(sh.ctx, &rep, nullptr, 0));
    if (rep.ndel || rep.naccept) {
      // the fix changed the atom arrays: refresh LAMMPS' copy in local order (same bookkeeping as
      // AtomVec::copy / create_atom, :827-841, :1121-1150)
      while (rep.nlocal > a->nmax) a->avec->grow(0);
      // group masks are not mirrored on the device: survivors moved into holes come from the tail,
      // so remember (tag -> mask) of the tail and of every old slot's tag
      const int nold = a->nlocal;
      std::vector<tagint> oldtag(a->tag, a->tag + nold);
      std::unordered_map<tagint, int> tailmask;
      for (int i = nold - rep.ndel < 0 ? 0 : nold - rep.ndel; i < nold; i++) tailmask[a->tag[i]] = a->mask[i];
      int n = 0;
      OBMD_CK(lmp, sh, obmd_download(sh.ctx, &n, &a->x[0][0], &a->v[0][0], nullptr, a->type, a->tag,
                                     a->image));
      const int nkeep = nold - rep.ndel;
      for (int i = 0; i < nkeep && i < n; i++)
        if (a->tag[i] != oldtag[i]) {
          auto it = tailmask.find(a->tag[i]);
          if (it != tailmask.end()) a->mask[i] = it->second;
        }
      for (int i = nkeep; i < n; i++) a->mask[i] = 1 | groupbit;
      a->natoms += (bigint) rep.naccept - (bigint) rep.ndel;
      a->nlocal = n;
      a->nghost = 0;
      if (a->map_style != Atom::MAP_NONE) {
        a->map_init();
        a->map_set();
      }
      sh.resident = n;
    }
    next_reneighbor += fa.nfreq;  // :632
  }

  // restart_global = 1 (:74): the reference sets the flag but saves nothing, so RanPark state and
  // simulation_time are lost across a restart file.  Here the fix state goes in and out.
  void write_restart(FILE *fp) override
  {
    obmd_fix_state st;
    if (obmd_fix_get_state(sh.ctx, &st) != OBMD_OK) return;
    if (comm->me == 0) {
      int size = (int) sizeof(st);
      fwrite(&size, sizeof(int), 1, fp);
      fwrite(&st, 1, sizeof(st), fp);
    }
  }
  void restart(char *buf) override
  {
    memcpy(&pending_state, buf, sizeof(pending_state));
    have_pending_state = true;  // applied in init(), once the device fix exists
  }

  bool has_variables() const
  {
    for (int k = 0; k < 8; k++)
      if (!varname[k].empty()) return true;
    return false;
  }
  void refresh_for_step() { refresh_variables(); }
  void after_device_run(bigint ntimestep) { next_reneighbor = ntimestep + fa.nfreq; }
  int group_bit() const { return groupbit; }

  void post_force(int) override  // :1229-1240
  {
    if (update->ntimestep % nevery) return;
    Atom *a = atom;
    if (sh.resident != a->nlocal || !sh.pair_ready) return;  // setup() before the first pair compute
    OBMD_CK(lmp, sh, obmd_zero_forces(sh.ctx));
    OBMD_CK(lmp, sh, obmd_post_force(sh.ctx));
    int n = 0;
    if ((int) fbuf.size() < 3 * a->nmax) fbuf.resize(3 * (size_t) a->nmax);
    OBMD_CK(lmp, sh, obmd_download(sh.ctx, &n, nullptr, nullptr, fbuf.data(), nullptr, nullptr, nullptr));
    double **f = a->f;
    for (int i = 0; i < n; i++) {
      f[i][0] += fbuf[3 * i];
      f[i][1] += fbuf[3 * i + 1];
      f[i][2] += fbuf[3 * i + 2];
    }
  }

 private:
  void refresh_variables()  // :563-572; v_pxx/v_pxy/v_pxz work here (the reference segfaults, Q2)
  {
    bool any = false;
    double *dst[8] = {&fa.pxx, &fa.pxy, &fa.pxz, &fa.dpxx, &fa.freq, &fa.alpha, &fa.tau, &fa.nbuf};
    for (int k = 0; k < 8; k++) {
      if (varname[k].empty()) continue;
      int iv = input_variable_find(varname[k].c_str());
      if (iv < 0) error->all(FLERR, "Variable name for fix obmd does not exist");
      *dst[k] = input_variable_equal(iv);
      any = true;
    }
    if (any)
      OBMD_CK(lmp, sh, obmd_fix_set_flux(sh.ctx, fa.pxx, fa.pxy, fa.pxz, fa.dpxx, fa.freq, fa.alpha,
                                         fa.tau, fa.nbuf));
  }
  int input_variable_find(const char *name);
  double input_variable_equal(int iv);

  void options(int narg, char **arg)  // :1881-2164, ATOM mode keywords
  {
    fa.g_fac = 0.25;
    fa.buffer_size = 0.30 * (domain->boxhi[0] - domain->boxlo[0]);
    fa.etarget = 3.6; fa.ds0 = 0.1; fa.dtheta0 = 0.35; fa.uovlp = 10000.0; fa.dsovlp = 3.0;
    fa.eps = 0.15; fa.nattempt = 40; fa.maxattempt = 1;
    fa.step_parallel = 0; fa.step_perp = 1; fa.usherflag = 0;
    int nearflag = 0, iarg = 0;
    while (iarg < narg) {
      if (strncmp(arg[iarg], "region", 6) == 0 && strlen(arg[iarg]) == 7 && arg[iarg][6] >= '1' &&
          arg[iarg][6] <= '6') {
        if (!domain->get_region_by_id(arg[iarg + 1]))
          error->all(FLERR, "Region ID {} for fix obmd does not exist", arg[iarg + 1]);
        idregion[arg[iarg][6] - '1'] = arg[iarg + 1];
        iarg += 2;
      } else if (strcmp(arg[iarg], "buffersize") == 0) {
        fa.buffer_size = utils::numeric(FLERR, arg[iarg + 1], false, lmp);
        if (fa.buffer_size <= 0.0)
          error->all(FLERR, "Illegal fix obmd command. Parameter buffersize should be larger than 0.0");
        iarg += 2;
      } else if (strcmp(arg[iarg], "gfac") == 0) {
        fa.g_fac = utils::numeric(FLERR, arg[iarg + 1], false, lmp);
        if (fa.g_fac < 0.0 || fa.g_fac > 1.0)
          error->all(FLERR, "Illegal fix obmd command. Parameter gfac should be between 0.0 and 1.0");
        iarg += 2;
      } else if (strcmp(arg[iarg], "stepparallel") == 0) {
        fa.step_parallel = utils::inumeric(FLERR, arg[iarg + 1], false, lmp);
        if (fa.step_parallel != 0)
          error->all(FLERR, "For now, only the smooth transition distribution function is implemented.");
        iarg += 2;
      } else if (strcmp(arg[iarg], "stepperp") == 0) {
        fa.step_perp = utils::inumeric(FLERR, arg[iarg + 1], false, lmp);
        if (fa.step_perp != 1)
          error->all(FLERR, "For now, only the heaviside step distribution function is implemented.");
        iarg += 2;
      } else if (strcmp(arg[iarg], "maxattempt") == 0) {
        fa.maxattempt = utils::inumeric(FLERR, arg[iarg + 1], false, lmp);
        iarg += 2;
      } else if (strcmp(arg[iarg], "usher") == 0) {
        if (iarg + 9 > narg) error->all(FLERR, "Illegal fix obmd command");
        fa.usherflag = utils::inumeric(FLERR, arg[iarg + 1], false, lmp);
        if (fa.usherflag == 1 && nearflag == 1) error->all(FLERR, "You can not have both usher and near");
        fa.etarget = utils::numeric(FLERR, arg[iarg + 2], false, lmp);
        fa.ds0 = utils::numeric(FLERR, arg[iarg + 3], false, lmp);
        fa.dtheta0 = utils::numeric(FLERR, arg[iarg + 4], false, lmp);
        fa.uovlp = utils::numeric(FLERR, arg[iarg + 5], false, lmp);
        fa.dsovlp = utils::numeric(FLERR, arg[iarg + 6], false, lmp);
        fa.eps = utils::numeric(FLERR, arg[iarg + 7], false, lmp);
        fa.nattempt = utils::inumeric(FLERR, arg[iarg + 8], false, lmp);
        iarg += 9;
      } else if (strcmp(arg[iarg], "charged") == 0) {
        if (utils::inumeric(FLERR, arg[iarg + 1], false, lmp) == 1)
          error->all(FLERR, "You can not use charged without MOLECULE mode");
        iarg += 2;
      } else if (strcmp(arg[iarg], "near") == 0) {
        nearflag = 1;
        error->all(FLERR, "obmd_b200: only usher insertion is implemented on the device path");
      } else
        error->all(FLERR, "Illegal fix obmd command");
    }
    if (fa.usherflag == 0 && nearflag == 0) error->all(FLERR, "You must pick either usher or near");
    for (int r = 0; r < 6; r++)
      if (idregion[r].empty())
        error->all(FLERR, r < 2 ? "Must specify a region in fix obmd"
                                : "Region ID for fix obmd does not exist");
  }

  Shared &sh;
  obmd_fix_args fa;
  bool created = false;
  obmd_fix_state pending_state;
  bool have_pending_state = false;
  std::string idregion[6];
  std::string varname[8];
  std::vector<double> fbuf;
};

// ------------------------------------------------------------------------------------------
// run_obmd N : N timesteps device-resident (stage B).  Same set-up as `run` (Run::command, run.cpp:37-260:
// lmp->init(), Integrate::setup -> step-0 forces, fix setup, thermo header), then obmd_run in chunks that
// end on the output steps (Output::next); the atoms come back to the host only there and at the end.
// Needs: pair_style dpd and fix obmd from this plugin, fix nve on group all (its two half-kicks run on the
// device: fix_nve.cpp:134-145,236-244), atom_modify sort 0 0, one MPI rank.
// ------------------------------------------------------------------------------------------
class RunObmd : public Command {
 public:
  RunObmd(LAMMPS *lmp) : Command(lmp) {}
  void command(int narg, char **arg) override
  {
    if (narg != 1) error->all(FLERR, "Illegal run_obmd command (usage: run_obmd N)");
    if (domain->box_exist == 0) error->all(FLERR, "run_obmd command before simulation box is defined");
    const bigint nsteps = utils::bnumeric(FLERR, arg[0], false, lmp);
    if (nsteps < 0 || nsteps > MAXSMALLINT) error->all(FLERR, "Invalid run_obmd command N value: {}", nsteps);
    FixObmdB200 *fobmd = nullptr;
    for (auto &f : modify->get_fix_list())
      if (!fobmd) fobmd = dynamic_cast<FixObmdB200 *>(f);
    if (!fobmd) error->all(FLERR, "run_obmd requires fix obmd from the obmd_b200 plugin");
    if (!dynamic_cast<PairDPDB200 *>(force->pair))
      error->all(FLERR, "run_obmd requires pair_style dpd from the obmd_b200 plugin");
    if (modify->get_fix_by_style("^nve$").empty())
      error->all(FLERR, "run_obmd integrates with FixNVE on the device: define `fix ID all nve`");
    if (modify->get_fix_list().size() != 2)
      error->all(FLERR, "run_obmd supports exactly two fixes: nve and obmd (others run on the host: use run)");
    Shared &sh = acquire(lmp);
    sh.users--;  // borrowed: the pair style and the fix own the context

    update->whichflag = 1;
    timer->init_timeout();
    update->nsteps = (int) nsteps;
    update->firststep = update->ntimestep;
    update->laststep = update->ntimestep + nsteps;
    update->beginstep = update->firststep;
    update->endstep = update->laststep;
    lmp->init();
    update->integrate->setup(1);  // host path once: exchange, step-0 forces, fix setup, thermo header

    // everything to the device, forces included (Integrate::setup left the complete forces on the host)
    Atom *a = atom;
    OBMD_CK(lmp, sh, obmd_set_timestep(sh.ctx, update->dt));
    OBMD_CK(lmp, sh, obmd_upload(sh.ctx, a->nlocal, &a->x[0][0], &a->v[0][0], &a->f[0][0], a->type, a->tag,
                                 a->image, a->nmax));
    sh.resident = a->nlocal;
    {  // atoms->natoms follows the fix's counters from here on
      obmd_fix_state st0;
      obmd_fix_get_state(sh.ctx, &st0);
      seen_ins = st0.ninserted;
      seen_del = st0.ndeleted;
    }

    timer->init();
    timer->barrier_start();
    const bool per_step = fobmd->has_variables();  // v_ arguments are evaluated on the host every step
    while (update->ntimestep < update->laststep) {
      bigint next = update->laststep;
      if (output->next > update->ntimestep && output->next < next) next = output->next;
      if (per_step) {
        next = update->ntimestep + 1;
        update->ntimestep = next;  // variables see the step being taken, as in Verlet::run
        fobmd->refresh_for_step();
        update->ntimestep = next - 1;
      }
      const int n = (int) (next - update->ntimestep);
      OBMD_CK(lmp, sh, obmd_run(sh.ctx, (long long) update->ntimestep + 1, n));
      update->ntimestep = next;
      if (next == output->next || next == update->laststep) {
        const bool out = (next == output->next);
        to_host(sh, fobmd, out);
        if (out) {
          timer->stamp();
          output->write(next);
          timer->stamp(Timer::OUTPUT);
        }
      }
    }
    timer->barrier_stop();
    fobmd->after_device_run(update->ntimestep);
    update->integrate->cleanup();
    Finish finish(lmp);
    finish.end(1);
    update->whichflag = 0;
    update->firststep = update->laststep = 0;
    update->beginstep = update->endstep = 0;
  }

 private:
  // the device state in LAMMPS' arrays, local order (what AtomVec::copy / create_atom would have left,
  // :827-841, :1121-1150); on output steps also the pair tallies thermo may ask for
  void to_host(Shared &sh, FixObmdB200 *fobmd, bool tallies)
  {
    Atom *a = atom;
    if (tallies) {
      // energy / virial of this step's configuration (same step key: the same random forces), then the
      // buffer forces post_force had added
      OBMD_CK(lmp, sh, obmd_pair_compute(sh.ctx, (long long) update->ntimestep, 1, 1));
      OBMD_CK(lmp, sh, obmd_post_force(sh.ctx));
      double e, v[6];
      OBMD_CK(lmp, sh, obmd_pair_tallies(sh.ctx, &e, v));
      force->pair->eng_vdwl = e;
      force->pair->eng_coul = 0.0;
      for (int k = 0; k < 6; k++) force->pair->virial[k] = v[k];
      update->eflag_global = update->vflag_global = update->ntimestep;
    }
    const int n = obmd_nlocal(sh.ctx);
    while (n > a->nmax) a->avec->grow(0);
    int nn = 0;
    OBMD_CK(lmp, sh, obmd_download(sh.ctx, &nn, &a->x[0][0], &a->v[0][0], &a->f[0][0], a->type, a->tag, a->image));
    for (int i = 0; i < nn; i++) a->mask[i] = 1 | fobmd->group_bit();
    obmd_fix_state st;
    obmd_fix_get_state(sh.ctx, &st);
    a->natoms += (bigint) (st.ninserted - seen_ins) - (bigint) (st.ndeleted - seen_del);
    seen_ins = st.ninserted;
    seen_del = st.ndeleted;
    a->nlocal = nn;
    a->nghost = 0;
    if (a->map_style != Atom::MAP_NONE) {
      a->map_init();
      a->map_set();
    }
    sh.resident = nn;
  }
  long long seen_ins = 0, seen_del = 0;
};

}    // namespace

#include "input.h"
#include "variable.h"
int FixObmdB200::input_variable_find(const char *name) { return input->variable->find(name); }
double FixObmdB200::input_variable_equal(int iv)
{
  if (!input->variable->equalstyle(iv)) error->all(FLERR, "Variable for fix obmd is invalid style");
  return input->variable->compute_equal(iv);
}

static Pair *pair_creator(LAMMPS *lmp) { return new PairDPDB200(lmp); }
static Bond *bond_creator(LAMMPS *lmp) { return new BondHarmonicB200(lmp); }
static Command *run_creator(LAMMPS *lmp) { return new RunObmd(lmp); }
static Fix *fix_creator(LAMMPS *lmp, int argc, char **argv) { return new FixObmdB200(lmp, argc, argv); }

extern "C" void lammpsplugin_init(void *lmp, void *handle, void *regfunc)
{
  lammpsplugin_t plugin;
  lammpsplugin_regfunc register_plugin = (lammpsplugin_regfunc) regfunc;
  plugin.version = LAMMPS_VERSION;
  plugin.author = "obmd_b200";
  plugin.handle = handle;

  plugin.style = "pair";
  plugin.name = "dpd";  // same name as the built-in: overrides it (PLUGIN/plugin.cpp:169-174)
  plugin.info = "pair_style dpd on B200 (libobmd_b200)";
  plugin.creator.v1 = (lammpsplugin_factory1 *) &pair_creator;
  (*register_plugin)(&plugin, lmp);

  plugin.style = "bond";
  plugin.name = "harmonic";  // overrides MOLECULE's bond_style harmonic (PLUGIN/plugin.cpp:176-183)
  plugin.info = "bond_style harmonic on B200 (libobmd_b200)";
  plugin.creator.v1 = (lammpsplugin_factory1 *) &bond_creator;
  (*register_plugin)(&plugin, lmp);

  plugin.style = "fix";
  plugin.name = "obmd";  // overrides the built-in fix obmd (PLUGIN/plugin.cpp:228-233)
  plugin.info = "fix obmd on B200 (libobmd_b200)";
  plugin.creator.v2 = (lammpsplugin_factory2 *) &fix_creator;
  (*register_plugin)(&plugin, lmp);

  plugin.style = "command";
  plugin.name = "run_obmd";  // device-resident `run` (stage B): see RunObmd above
  plugin.info = "run N timesteps of fix nve + fix obmd + pair dpd resident on B200 (libobmd_b200)";
  plugin.creator.v1 = (lammpsplugin_factory1 *) &run_creator;
  (*register_plugin)(&plugin, lmp);
}
