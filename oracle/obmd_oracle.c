/* TEST INFRASTRUCTURE -- CPU oracle for the OBMD-DPD hot path.  NOT the product.
 * See obmd_oracle.h for scope, pinning status and the citation convention.
 * Compile with -ffp-contract=off: the reference is an x86-64 -O3 build without
 * FMA contraction, and bit-exact parity depends on it.
 */
#include "obmd_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846 /* MathConst::MY_PI (math_const.h) */
#define EPSILON_FIX 1.0e-6        /* fix_obmd_merged.cpp:62 */
#define EPSILON_DPD 1.0e-10       /* pair_dpd.cpp:36 */

/* ------------------------------------------------------------------ RNGs */

/* random_park.cpp:24-50 -- Park-Miller minimal standard with Schrage's trick */
double orc_ranpark_uniform(int *seed)
{
  int k = *seed / 127773;
  *seed = 16807 * (*seed - k * 127773) - 2836 * k;
  if (*seed < 0) *seed += 2147483647;
  return (1.0 / 2147483647) * (*seed);
}

/* n successive uniform() calls leave seed = 16807^n * seed mod (2^31-1); the
   Schrage recurrence is exactly that product, so a modular power skips ahead.
   Used for the deletion draws of fix_obmd_merged.cpp:689-701 (Q7). */
int orc_ranpark_skip(int seed, long n)
{
  unsigned long long m = 2147483647ULL, a = 16807ULL, r = 1ULL;
  while (n > 0) {
    if (n & 1) r = (r * a) % m;
    a = (a * a) % m;
    n >>= 1;
  }
  return (int) (((unsigned long long) seed * r) % m);
}

/* random_mars.cpp:30-69 */
void orc_ranmars_init(orc_ranmars *r, int seed)
{
  int ij, kl, i, j, k, l, ii, jj, m;
  double s, t;
  memset(r, 0, sizeof(*r));
  ij = (seed - 1) / 30082;
  kl = (seed - 1) - 30082 * ij;
  i = (ij / 177) % 177 + 2;
  j = ij % 177 + 2;
  k = (kl / 169) % 178 + 1;
  l = kl % 169;
  for (ii = 1; ii <= 97; ii++) {
    s = 0.0;
    t = 0.5;
    for (jj = 1; jj <= 24; jj++) {
      m = ((i * j) % 179) * k % 179;
      i = j;
      j = k;
      k = m;
      l = (53 * l + 1) % 169;
      if ((l * m) % 64 >= 32) s = s + t;
      t = 0.5 * t;
    }
    r->u[ii] = s;
  }
  r->c = 362436.0 / 16777216.0;
  r->cd = 7654321.0 / 16777216.0;
  r->cm = 16777213.0 / 16777216.0;
  r->i97 = 97;
  r->j97 = 33;
  orc_ranmars_uniform(r);
}

/* random_mars.cpp:82-98 */
double orc_ranmars_uniform(orc_ranmars *r)
{
  double uni = r->u[r->i97] - r->u[r->j97];
  if (uni < 0.0) uni += 1.0;
  r->u[r->i97] = uni;
  r->i97--;
  if (r->i97 == 0) r->i97 = 97;
  r->j97--;
  if (r->j97 == 0) r->j97 = 97;
  r->c -= r->cd;
  if (r->c < 0.0) r->c += r->cm;
  uni -= r->c;
  if (uni < 0.0) uni += 1.0;
  return uni;
}

/* random_mars.cpp:103-122 -- polar Box-Muller with a saved second variate */
double orc_ranmars_gaussian(orc_ranmars *r)
{
  double first, v1, v2, rsq, fac;
  if (!r->save) {
    do {
      v1 = 2.0 * orc_ranmars_uniform(r) - 1.0;
      v2 = 2.0 * orc_ranmars_uniform(r) - 1.0;
      rsq = v1 * v1 + v2 * v2;
    } while ((rsq >= 1.0) || (rsq == 0.0));
    fac = sqrt(-2.0 * log(rsq) / rsq);
    r->second = v1 * fac;
    first = v2 * fac;
    r->save = 1;
  } else {
    first = r->second;
    r->save = 0;
  }
  return first;
}

/* ------------------------------------------------------------- geometry */

/* region_block.cpp:289-293: closed interval on all three axes (Q9) */
static int block_inside(const orc_block *b, double x, double y, double z)
{
  return (x >= b->lo[0] && x <= b->hi[0] && y >= b->lo[1] && y <= b->hi[1] && z >= b->lo[2] &&
          z <= b->hi[2]);
}

/* domain.cpp:980-1008 for boundary f p p (x is not periodic) */
static void minimum_image_yz(const orc_fix *p, double *dy, double *dz)
{
  double yprd = p->boxhi[1] - p->boxlo[1], zprd = p->boxhi[2] - p->boxlo[2];
  double yh = 0.5 * yprd, zh = 0.5 * zprd;
  while (fabs(*dy) > yh) {
    if (*dy < 0.0) *dy += yprd;
    else *dy -= yprd;
  }
  while (fabs(*dz) > zh) {
    if (*dz < 0.0) *dz += zprd;
    else *dz -= zprd;
  }
}

/* ---------------------------------------------------------------- cells
   A plain linked-list cell grid (binsize >= largest cutoff).  y,z indices wrap,
   x index clamps.  Used by the pair loop and by the cell variant of energy(). */
typedef struct {
  int n[3];
  double lo[3], inv[3];
  int *head, *next;
  int cap;
} cells_t;

static double max_cut(const orc_fix *p)
{
  double c = 0.0;
  for (int i = 1; i <= p->ntypes; i++)
    for (int j = 1; j <= p->ntypes; j++)
      if (p->cut[i][j] > c) c = p->cut[i][j];
  return c;
}

static int cell_coord(const cells_t *c, int d, double x)
{
  int k = (int) floor((x - c->lo[d]) * c->inv[d]);
  if (d == 0) {
    if (k < 0) k = 0;
    if (k >= c->n[0]) k = c->n[0] - 1;
  } else {
    k %= c->n[d];
    if (k < 0) k += c->n[d];
  }
  return k;
}

static int cell_of(const cells_t *c, const double *x)
{
  return (cell_coord(c, 0, x[0]) * c->n[1] + cell_coord(c, 1, x[1])) * c->n[2] +
      cell_coord(c, 2, x[2]);
}

static void cells_push(cells_t *c, int i, const double *x)
{
  if (i >= c->cap) {
    c->cap = 2 * i + 16;
    c->next = (int *) realloc(c->next, sizeof(int) * c->cap);
  }
  int b = cell_of(c, x);
  c->next[i] = c->head[b];
  c->head[b] = i;
}

static int cells_build(cells_t *c, const orc_fix *p, const orc_atoms *a)
{
  double rc = max_cut(p);
  for (int d = 0; d < 3; d++) {
    double L = p->boxhi[d] - p->boxlo[d];
    c->n[d] = (int) floor(L / rc);
    if (c->n[d] < 1) c->n[d] = 1;
    c->lo[d] = p->boxlo[d];
    c->inv[d] = c->n[d] / L;
  }
  if (c->n[1] < 3 || c->n[2] < 3) return -1;
  int nc = c->n[0] * c->n[1] * c->n[2];
  c->head = (int *) malloc(sizeof(int) * nc);
  c->cap = a->nlocal + 16;
  c->next = (int *) malloc(sizeof(int) * c->cap);
  for (int b = 0; b < nc; b++) c->head[b] = -1;
  /* reverse insertion keeps each list in ascending index order */
  for (int i = a->nlocal - 1; i >= 0; i--) cells_push(c, i, &a->x[3 * i]);
  return 0;
}

static void cells_free(cells_t *c)
{
  free(c->head);
  free(c->next);
  c->head = c->next = NULL;
}

/* ------------------------------------------------------------- pair dpd */

/* pair_dpd.cpp:279-291 (init_one) + Pair::init cutsq */
void orc_pair_init(orc_fix *p)
{
  for (int i = 1; i <= p->ntypes; i++)
    for (int j = i; j <= p->ntypes; j++) {
      p->sigma[i][j] = sqrt(2.0 * p->boltz * p->temperature * p->gamma[i][j]);
      p->cut[j][i] = p->cut[i][j];
      p->a0[j][i] = p->a0[i][j];
      p->gamma[j][i] = p->gamma[i][j];
      p->sigma[j][i] = p->sigma[i][j];
      p->cutsq[i][j] = p->cutsq[j][i] = p->cut[i][j] * p->cut[i][j];
    }
}

/* pair_dpd.cpp:401-418 with factor_dpd = 1 (fix_obmd_merged.cpp:1786-1801) */
double orc_pair_single(const orc_fix *p, int itype, int jtype, double rsq, double *fforce)
{
  double r, rinv, wd, phi;
  const double factor_dpd = 1.0;
  r = sqrt(rsq);
  if (r < EPSILON_DPD) {
    *fforce = 0.0;
    return 0.0;
  }
  rinv = 1.0 / r;
  wd = 1.0 - r / p->cut[itype][jtype];
  *fforce = p->a0[itype][jtype] * wd * factor_dpd * rinv;
  phi = 0.5 * p->a0[itype][jtype] * p->cut[itype][jtype] * wd * wd;
  return factor_dpd * phi;
}

/* Design substitution for RanMars (see header).  splitmix64 finaliser over a
   symmetric (min tag, max tag) key and a (seed, step) counter, then a
   single-precision Box-Muller.  The CUDA kernel implements the same function
   with __logf/__cosf, so agreement is to ~1e-6, not bitwise. */
double orc_pair_gaussian(int tagi, int tagj, unsigned step, unsigned seed)
{
  unsigned lo = (unsigned) (tagi < tagj ? tagi : tagj), hi = (unsigned) (tagi < tagj ? tagj : tagi);
  unsigned long long key = ((unsigned long long) lo << 32) | hi;
  unsigned long long ctr = ((unsigned long long) seed << 32) | step;
  unsigned long long z = key * 0x9E3779B97F4A7C15ULL + ctr * 0xD1B54A32D192ED03ULL;
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  unsigned h = (unsigned) (z >> 32), l = (unsigned) z;
  float u1 = ((float) h + 1.0f) * 2.3283064365386963e-10f; /* (0,1] */
  float u2 = (float) l * 2.3283064365386963e-10f;          /* [0,1] */
  float r = sqrtf(-2.0f * logf(u1));
  return (double) (r * cosf(6.283185307179586f * u2));
}

/* pair_dpd.cpp:66-163 evaluated as a full list: every owned atom i visits all
   j != i within the cutoff and accumulates only f[i].  The per-pair expression
   and its evaluation order follow :117-146 literally. */
void orc_pair_compute(const orc_fix *p, orc_atoms *a, int rng_mode, unsigned step, double *eng,
                      double *virial)
{
  cells_t c;
  double e = 0.0, vir[6] = {0, 0, 0, 0, 0, 0};
  double dtinvsqrt = 1.0 / sqrt(p->dt);
  if (cells_build(&c, p, a) != 0) {
    fprintf(stderr, "orc_pair_compute: box needs >= 3 cells in y and z\n");
    abort();
  }
  for (int i = 0; i < a->nlocal; i++) {
    const double *xi = &a->x[3 * i], *vi = &a->v[3 * i];
    int itype = a->type[i];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    int cx = cell_coord(&c, 0, xi[0]), cy = cell_coord(&c, 1, xi[1]), cz = cell_coord(&c, 2, xi[2]);
    for (int ox = -1; ox <= 1; ox++) {
      int bx = cx + ox;
      if (bx < 0 || bx >= c.n[0]) continue;
      for (int oy = -1; oy <= 1; oy++) {
        int by = (cy + oy + c.n[1]) % c.n[1];
        for (int oz = -1; oz <= 1; oz++) {
          int bz = (cz + oz + c.n[2]) % c.n[2];
          for (int j = c.head[(bx * c.n[1] + by) * c.n[2] + bz]; j >= 0; j = c.next[j]) {
            if (j == i) continue;
            int jtype = a->type[j];
            double delx = xi[0] - a->x[3 * j], dely = xi[1] - a->x[3 * j + 1],
                   delz = xi[2] - a->x[3 * j + 2];
            minimum_image_yz(p, &dely, &delz);
            double rsq = delx * delx + dely * dely + delz * delz;
            if (rsq < p->cutsq[itype][jtype]) {
              double r = sqrt(rsq);
              if (r < EPSILON_DPD) continue;
              double rinv = 1.0 / r;
              double delvx = vi[0] - a->v[3 * j], delvy = vi[1] - a->v[3 * j + 1],
                     delvz = vi[2] - a->v[3 * j + 2];
              double dot = delx * delvx + dely * delvy + delz * delvz;
              double wd = 1.0 - r / p->cut[itype][jtype];
              double randnum = 0.0;
              if (rng_mode == 1)
                randnum = orc_pair_gaussian(a->tag[i], a->tag[j], step, (unsigned) p->pair_seed);
              double fpair = p->a0[itype][jtype] * wd;
              fpair -= p->gamma[itype][jtype] * wd * wd * dot * rinv;
              fpair *= 1.0;
              fpair += 1.0 * p->sigma[itype][jtype] * wd * randnum * dtinvsqrt;
              fpair *= rinv;
              fx += delx * fpair;
              fy += dely * fpair;
              fz += delz * fpair;
              e += 0.5 * (0.5 * p->a0[itype][jtype] * p->cut[itype][jtype] * wd * wd);
              vir[0] += 0.5 * delx * delx * fpair;
              vir[1] += 0.5 * dely * dely * fpair;
              vir[2] += 0.5 * delz * delz * fpair;
              vir[3] += 0.5 * delx * dely * fpair;
              vir[4] += 0.5 * delx * delz * fpair;
              vir[5] += 0.5 * dely * delz * fpair;
            }
          }
        }
      }
    }
    a->f[3 * i] = fx;
    a->f[3 * i + 1] = fy;
    a->f[3 * i + 2] = fz;
  }
  cells_free(&c);
  if (eng) *eng = e;
  if (virial) memcpy(virial, vir, sizeof(vir));
}

/* -------------------------------------------------------------- fix obmd */

/* fix_obmd_merged.cpp:275-276: RanPark(seed) warmed by 30 draws; Q3: time and
   forces start at exactly 0 (what the -O3 reference build shows, Appendix B5). */
void orc_fix_init(orc_fix *p)
{
  p->rng_seed = p->seed;
  for (int ii = 0; ii < 30; ii++) orc_ranpark_uniform(&p->rng_seed);
  p->simulation_time = 0.0;
  for (int d = 0; d < 3; d++) {
    p->momentumForce_left[d] = p->momentumForce_right[d] = 0.0;
    p->shearForce_left[d] = p->shearForce_right[d] = 0.0;
    p->vnewl[d] = p->vnewr[d] = 0.0;
  }
  p->ninserted = p->ndeleted = 0;
  p->maxtag_all = 0;
}

/* AtomVec::copy (atom_vec.cpp:266-281): tag,type,mask,image,x,v -- not f */
static void atom_copy(orc_atoms *a, int i, int j)
{
  a->tag[j] = a->tag[i];
  a->type[j] = a->type[i];
  a->image[j] = a->image[i];
  for (int d = 0; d < 3; d++) {
    a->x[3 * j + d] = a->x[3 * i + d];
    a->v[3 * j + d] = a->v[3 * i + d];
  }
}

/* fix_obmd_merged.cpp:639-865, ATOM mode, one rank (nbefore = 0).  The region
   argument is only prematch()ed (Q6): every atom outside [boxlo,boxhi] in x goes. */
static int try_deleting(orc_fix *p, orc_atoms *a)
{
  int nlocal = a->nlocal;
  double boxl = p->boxlo[0], boxh = p->boxhi[0];
  int *list = (int *) malloc(sizeof(int) * (nlocal + 1));
  char *mark = (char *) calloc(nlocal + 1, 1);
  int ncount = 0;
  for (int i = 0; i < nlocal; i++)
    if (a->x[3 * i] < boxl || a->x[3 * i] > boxh) list[ncount++] = i;

  int nall = ncount, ndel = 0;
  while (nall) { /* :689-701, one uniform() per candidate (Q7) */
    int iwhich = (int) (nall * orc_ranpark_uniform(&p->rng_seed));
    if (iwhich < ncount) {
      mark[list[iwhich]] = 1;
      list[iwhich] = list[ncount - 1];
      ncount--;
    }
    ndel++;
    nall--;
  }
  /* :827-841 reverse loop, momentum side test on the pre-copy position */
  for (int i = nlocal - 1; i >= 0; i--) {
    if (mark[i]) {
      double m = p->mass[a->type[i]];
      double *vn = (a->x[3 * i] < 0.5 * (boxh + boxl)) ? p->vnewl : p->vnewr;
      vn[0] += m * a->v[3 * i];
      vn[1] += m * a->v[3 * i + 1];
      vn[2] += m * a->v[3 * i + 2];
      atom_copy(a, a->nlocal - 1, i);
      a->nlocal--;
    }
  }
  a->natoms -= ndel;
  p->ndeleted += ndel;
  free(list);
  free(mark);
  return ndel;
}

/* group.cpp:793-811 for group all */
static int group_count(const orc_atoms *a, const orc_block *reg)
{
  int n = 0;
  for (int i = 0; i < a->nlocal; i++)
    if (block_inside(reg, a->x[3 * i], a->x[3 * i + 1], a->x[3 * i + 2])) n++;
  return n;
}

/* fix_obmd_merged.cpp:1774-1810: brute-force scan over owned atoms in local order */
double orc_energy(const orc_fix *p, const orc_atoms *a, int itype, const double *coord,
                  double *fusher)
{
  double total_energy = 0.0, fpair = 0.0;
  for (int j = 0; j < a->nlocal; j++) {
    double delx = coord[0] - a->x[3 * j];
    double dely = coord[1] - a->x[3 * j + 1];
    double delz = coord[2] - a->x[3 * j + 2];
    minimum_image_yz(p, &dely, &delz);
    double rsq = delx * delx + dely * dely + delz * delz;
    int jtype = a->type[j];
    if (rsq < p->cutsq[itype][jtype]) {
      total_energy += orc_pair_single(p, itype, jtype, rsq, &fpair);
      fusher[0] += fpair * delx;
      fusher[1] += fpair * dely;
      fusher[2] += fpair * delz;
    }
  }
  return total_energy;
}

static int cmp_int(const void *a, const void *b)
{
  return (*(const int *) a > *(const int *) b) - (*(const int *) a < *(const int *) b);
}

/* Same sums in the same (ascending local index) order, candidates from 27 cells. */
static cells_t *g_cells = NULL;
static double energy_cells(const orc_fix *p, const orc_atoms *a, int itype, const double *coord,
                           double *fusher)
{
  const cells_t *c = g_cells;
  int cand[4096], nc = 0;
  int cx = cell_coord(c, 0, coord[0]), cy = cell_coord(c, 1, coord[1]),
      cz = cell_coord(c, 2, coord[2]);
  for (int ox = -1; ox <= 1; ox++) {
    int bx = cx + ox;
    if (bx < 0 || bx >= c->n[0]) continue;
    for (int oy = -1; oy <= 1; oy++) {
      int by = (cy + oy + c->n[1]) % c->n[1];
      for (int oz = -1; oz <= 1; oz++) {
        int bz = (cz + oz + c->n[2]) % c->n[2];
        for (int j = c->head[(bx * c->n[1] + by) * c->n[2] + bz]; j >= 0; j = c->next[j])
          if (nc < 4096) cand[nc++] = j;
      }
    }
  }
  qsort(cand, nc, sizeof(int), cmp_int);
  double total_energy = 0.0, fpair = 0.0;
  for (int k = 0; k < nc; k++) {
    int j = cand[k];
    double delx = coord[0] - a->x[3 * j];
    double dely = coord[1] - a->x[3 * j + 1];
    double delz = coord[2] - a->x[3 * j + 2];
    minimum_image_yz(p, &dely, &delz);
    double rsq = delx * delx + dely * dely + delz * delz;
    int jtype = a->type[j];
    if (rsq < p->cutsq[itype][jtype]) {
      total_energy += orc_pair_single(p, itype, jtype, rsq, &fpair);
      fusher[0] += fpair * delx;
      fusher[1] += fpair * dely;
      fusher[2] += fpair * delz;
    }
  }
  return total_energy;
}

/* fix_obmd_merged.cpp:1518-1616, ATOM mode (natom = 1, no torque terms) */
double orc_usher(const orc_fix *p, const orc_atoms *a, const orc_block *reg, double *coord,
                 int *iter)
{
  int i = 0;
  double entmp_all = 0.0, fusher[3], fabs_, ds;
  while (i < p->nattempt) {
    fusher[0] = fusher[1] = fusher[2] = 0.0;
    if (p->use_cells && g_cells) entmp_all = energy_cells(p, a, p->ntype, coord, fusher);
    else entmp_all = orc_energy(p, a, p->ntype, coord, fusher);

    if (entmp_all < p->etarget + EPSILON_FIX) break;
    else if (entmp_all > p->uovlp) {
      fabs_ = sqrt(fusher[0] * fusher[0] + fusher[1] * fusher[1] + fusher[2] * fusher[2]);
      if (fabs_ < EPSILON_FIX) break; /* the reference `continue`s forever here (Q10) */
      ds = p->dsovlp - pow(4 * p->eps / entmp_all, 1.0 / 12.0);
      coord[0] += fusher[0] / fabs_ * ds;
      coord[1] += fusher[1] / fabs_ * ds;
      coord[2] += fusher[2] / fabs_ * ds;
      if (!block_inside(reg, coord[0], coord[1], coord[2])) break;
    } else {
      fabs_ = sqrt(fusher[0] * fusher[0] + fusher[1] * fusher[1] + fusher[2] * fusher[2]);
      if (fabs_ < EPSILON_FIX) break; /* idem */
      ds = fmin((entmp_all - p->etarget) / fabs_, p->ds0);
      coord[0] += fusher[0] / fabs_ * ds;
      coord[1] += fusher[1] / fabs_ * ds;
      coord[2] += fusher[2] / fabs_ * ds;
      if (!block_inside(reg, coord[0], coord[1], coord[2])) break;
    }
    i += 1;
  }
  *iter = i;
  return entmp_all;
}

/* fix_obmd_merged.cpp:1645-1672 on one rank: sub-box = box, half-open; the top
   z face is owned by the (only) top processor. */
static int check_proc(const orc_fix *p, const double *c)
{
  const double *lo = p->boxlo, *hi = p->boxhi;
  if (c[0] >= lo[0] && c[0] < hi[0] && c[1] >= lo[1] && c[1] < hi[1] && c[2] >= lo[2] &&
      c[2] < hi[2])
    return 1;
  if (c[2] >= hi[2] && c[0] >= lo[0] && c[0] < hi[0] && c[1] >= lo[1] && c[1] < hi[1]) return 1;
  return 0;
}

/* fix_obmd_merged.cpp:869-1225, ATOM mode, DIST_UNIFORM, usher */
static void try_inserting(orc_fix *p, orc_atoms *a, int side, int stev, orc_trial *trials,
                          int maxtrials)
{
  const orc_block *reg = &p->region[4 + side];
  /* find_maxid (:1860-1875), idnext == 0 (Q12) */
  int mx = 0;
  for (int i = 0; i < a->nlocal; i++)
    if (a->tag[i] > mx) mx = a->tag[i];
  p->maxtag_all = mx;
  if (stev <= 0) return;

  for (int ninsert = 0; ninsert != stev; ninsert++) {
    int attempt = 0;
    while (attempt < p->maxattempt) {
      attempt++;
      double coord[3];
      do {
        coord[0] = reg->lo[0] + orc_ranpark_uniform(&p->rng_seed) * (reg->hi[0] - reg->lo[0]);
        coord[1] = reg->lo[1] + orc_ranpark_uniform(&p->rng_seed) * (reg->hi[1] - reg->lo[1]);
        coord[2] = reg->lo[2] + orc_ranpark_uniform(&p->rng_seed) * (reg->hi[2] - reg->lo[2]);
      } while (!block_inside(reg, coord[0], coord[1], coord[2]));

      orc_trial t;
      memset(&t, 0, sizeof(t));
      t.side = side;
      memcpy(t.start, coord, sizeof(coord));
      int iter = 0, ok = 1;
      double entmp = orc_usher(p, a, reg, coord, &iter);
      if (!(entmp < p->etarget + EPSILON_FIX)) ok = 0;
      if (ok && !check_proc(p, coord)) ok = 0;               /* :1096-1103 */
      if (ok && !block_inside(reg, coord[0], coord[1], coord[2])) ok = 0; /* :1107-1108 */
      t.energy = entmp;
      t.iter = iter;
      memcpy(t.end, coord, sizeof(coord));
      if (ok) {
        /* create_atom (atom_vec.cpp:1592-1612) + :1129-1150 */
        int n = a->nlocal;
        if (n >= a->nmax) {
          fprintf(stderr, "orc: atom arrays full (nmax=%d)\n", a->nmax);
          abort();
        }
        a->type[n] = p->ntype;
        a->x[3 * n] = coord[0];
        a->x[3 * n + 1] = coord[1];
        a->x[3 * n + 2] = coord[2];
        a->v[3 * n] = a->v[3 * n + 1] = a->v[3 * n + 2] = 0.0;
        a->image[n] = (ORC_IMGMAX << ORC_IMG2BITS) | (ORC_IMGMAX << ORC_IMGBITS) | ORC_IMGMAX;
        a->tag[n] = p->maxtag_all + 1;
        a->nlocal++;
        a->natoms += 1;
        p->maxtag_all += 1;
        p->ninserted++;
        t.accepted = 1;
        t.tag = a->tag[n];
        if (g_cells) cells_push(g_cells, n, coord);
        p->last_naccept++;
      }
      if (trials && p->last_ntrials < maxtrials) trials[p->last_ntrials] = t;
      p->last_ntrials++;
      if (ok) break;
    }
  }
}

/* fix_obmd_merged.cpp:550-633 */
void orc_fix_pre_exchange(orc_fix *p, orc_atoms *a, orc_trial *trials, int maxtrials)
{
  double ly = p->boxhi[1] - p->boxlo[1];
  double lz = p->boxhi[2] - p->boxlo[2];
  const double mol_len = 1.0;
  p->vnewl[0] = p->vnewl[1] = p->vnewl[2] = 0.0;
  p->vnewr[0] = p->vnewr[1] = p->vnewr[2] = 0.0;
  p->last_ntrials = p->last_naccept = 0;

  p->last_ndel = try_deleting(p, a);
  p->last_ndel += try_deleting(p, a);

  int cnt_left = group_count(a, &p->region[0]);
  int cnt_right = group_count(a, &p->region[1]);
  double ninsert_left =
      -(int) (((double) cnt_left / mol_len - p->alpha * p->nbuf) * p->dt / p->tau);
  double ninsert_right =
      -(int) (((double) cnt_right / mol_len - p->alpha * p->nbuf) * p->dt / p->tau);
  p->last_cnt_left = cnt_left;
  p->last_cnt_right = cnt_right;
  p->last_ninsert_left = (int) ninsert_left;
  p->last_ninsert_right = (int) ninsert_right;

  cells_t c;
  g_cells = NULL;
  if (p->use_cells && ((int) ninsert_left > 0 || (int) ninsert_right > 0)) {
    if (cells_build(&c, p, a) == 0) g_cells = &c;
  }
  try_inserting(p, a, 0, (int) ninsert_left, trials, maxtrials);
  try_inserting(p, a, 1, (int) ninsert_right, trials, maxtrials);
  if (g_cells) {
    cells_free(&c);
    g_cells = NULL;
  }

  p->last_ndel += try_deleting(p, a);
  p->last_ndel += try_deleting(p, a);

  double area = ly * lz;
  p->simulation_time += p->dt;
  double factor = p->pxx + p->dpxx * sin(2.0 * PI * p->freq * p->simulation_time);
  p->momentumForce_left[0] = p->vnewl[0] / (p->dt) + factor * area;
  p->momentumForce_left[1] = p->vnewl[1] / (p->dt);
  p->momentumForce_left[2] = p->vnewl[2] / (p->dt);
  p->shearForce_left[0] = 0.0;
  p->shearForce_left[1] = p->pxy * area;
  p->shearForce_left[2] = p->pxz * area;
  p->momentumForce_right[0] = p->vnewr[0] / (p->dt) - p->pxx * area;
  p->momentumForce_right[1] = p->vnewr[1] / (p->dt);
  p->momentumForce_right[2] = p->vnewr[2] / (p->dt);
  p->shearForce_right[0] = 0.0;
  p->shearForce_right[1] = -p->pxy * area;
  p->shearForce_right[2] = -p->pxz * area;
}

/* fix_obmd_merged.cpp:1312-1340.  The fall-through (inside the region but in
   neither buffer, Q13) has no return value in the reference; parity decks keep
   that set empty.  Here it returns 0. */
double orc_g_par_local(const orc_fix *p, double mass, double xv)
{
  double upper_x = p->boxhi[0], lower_x = p->boxlo[0];
  double buffer_size = p->buffer_size, g_fac = p->g_fac, carg;
  if (xv < lower_x + buffer_size) {
    if (xv < (lower_x + (1.0 - g_fac) * buffer_size)) return mass;
    carg = 1.0 / g_fac * PI * (xv - buffer_size - lower_x) / (-buffer_size) - PI;
    return 0.5 * (1.0 + cos(carg)) * mass;
  }
  if (xv > upper_x - buffer_size) {
    if (xv > (upper_x - (1.0 - g_fac) * buffer_size)) return mass;
    carg = 1.0 / g_fac * PI * (xv - upper_x + buffer_size) / (buffer_size) -PI;
    return 0.5 * (1.0 + cos(carg)) * mass;
  }
  return 0.0;
}

/* fix_obmd_merged.cpp:1244-1308: note both buffer tests are evaluated (no else) */
static double g_par_global(const orc_fix *p, const orc_atoms *a, const orc_block *reg)
{
  double upper_x = p->boxhi[0], lower_x = p->boxlo[0];
  double buffer_size = p->buffer_size, g_fac = p->g_fac, carg, g_par_all = 0.0;
  for (int i = 0; i < a->nlocal; i++) {
    double mass_temp = p->mass[a->type[i]];
    const double *x = &a->x[3 * i];
    if (!block_inside(reg, x[0], x[1], x[2])) continue;
    if (x[0] < lower_x + buffer_size) {
      if (x[0] < (lower_x + (1.0 - g_fac) * buffer_size)) g_par_all += mass_temp;
      else {
        carg = 1.0 / g_fac * PI * (x[0] - buffer_size - lower_x) / (-buffer_size) - PI;
        g_par_all += 0.5 * (1.0 + cos(carg)) * mass_temp;
      }
    }
    if (x[0] > upper_x - buffer_size) {
      if (x[0] > (upper_x - (1.0 - g_fac) * buffer_size)) g_par_all += mass_temp;
      else {
        carg = 1.0 / g_fac * PI * (x[0] - upper_x + buffer_size) / (buffer_size) -PI;
        g_par_all += 0.5 * (1.0 + cos(carg)) * mass_temp;
      }
    }
  }
  return g_par_all;
}

/* fix_obmd_merged.cpp:1385-1448 */
static void reg_force(const orc_fix *p, orc_atoms *a, const orc_block *reg, const double *F)
{
  double gtmp = g_par_global(p, a, reg);
  for (int i = 0; i < a->nlocal; i++) {
    const double *x = &a->x[3 * i];
    if (!block_inside(reg, x[0], x[1], x[2])) continue;
    double gloctmp = orc_g_par_local(p, p->mass[a->type[i]], x[0]);
    a->f[3 * i] += F[0] * gloctmp / gtmp;
    a->f[3 * i + 1] += F[1] * gloctmp / gtmp;
    a->f[3 * i + 2] += F[2] * gloctmp / gtmp;
  }
}

/* fix_obmd_merged.cpp:1344-1381 and 1452-1514 */
static void reg_force_perp(const orc_fix *p, orc_atoms *a, const orc_block *reg, const double *F)
{
  double gtmp = 0.0;
  for (int i = 0; i < a->nlocal; i++) {
    const double *x = &a->x[3 * i];
    if (!block_inside(reg, x[0], x[1], x[2])) continue;
    gtmp += p->mass[a->type[i]];
  }
  for (int i = 0; i < a->nlocal; i++) {
    const double *x = &a->x[3 * i];
    if (!block_inside(reg, x[0], x[1], x[2])) continue;
    double gloctmp = p->mass[a->type[i]];
    a->f[3 * i] += F[0] * gloctmp / gtmp;
    a->f[3 * i + 1] += F[1] * gloctmp / gtmp;
    a->f[3 * i + 2] += F[2] * gloctmp / gtmp;
  }
}

/* fix_obmd_merged.cpp:1229-1240 */
void orc_fix_post_force(const orc_fix *p, orc_atoms *a)
{
  reg_force(p, a, &p->region[0], p->momentumForce_left);
  reg_force(p, a, &p->region[1], p->momentumForce_right);
  reg_force_perp(p, a, &p->region[2], p->shearForce_left);
  reg_force_perp(p, a, &p->region[3], p->shearForce_right);
}

/* ------------------------------------------------------------ integrator */

/* fix_nve.cpp:134-145 (per-type mass branch), ftm2v = 1 in lj units */
void orc_nve_initial(const orc_fix *p, orc_atoms *a)
{
  double dtv = p->dt, dtf = 0.5 * p->dt * 1.0;
  for (int i = 0; i < a->nlocal; i++) {
    double dtfm = dtf / p->mass[a->type[i]];
    for (int d = 0; d < 3; d++) {
      a->v[3 * i + d] += dtfm * a->f[3 * i + d];
      a->x[3 * i + d] += dtv * a->v[3 * i + d];
    }
  }
}

/* fix_nve.cpp:236-244 */
void orc_nve_final(const orc_fix *p, orc_atoms *a)
{
  double dtf = 0.5 * p->dt * 1.0;
  for (int i = 0; i < a->nlocal; i++) {
    double dtfm = dtf / p->mass[a->type[i]];
    for (int d = 0; d < 3; d++) a->v[3 * i + d] += dtfm * a->f[3 * i + d];
  }
}

/* domain.cpp:529-640 for boundary f p p: only y and z wrap */
void orc_pbc(const orc_fix *p, orc_atoms *a)
{
  for (int i = 0; i < a->nlocal; i++) {
    for (int d = 1; d < 3; d++) {
      double lo = p->boxlo[d], hi = p->boxhi[d], period = hi - lo;
      int shift = (d == 1) ? ORC_IMGBITS : ORC_IMG2BITS;
      double *x = &a->x[3 * i + d];
      if (*x < lo) {
        *x += period;
        int idim = (a->image[i] >> shift) & ORC_IMGMASK;
        int other = a->image[i] ^ (idim << shift);
        idim = (idim - 1) & ORC_IMGMASK;
        a->image[i] = other | (idim << shift);
      }
      if (*x >= hi) {
        *x -= period;
        if (*x < lo) *x = lo; /* MAX(x,lo) */
        int idim = (a->image[i] >> shift) & ORC_IMGMASK;
        int other = a->image[i] ^ (idim << shift);
        idim = (idim + 1) & ORC_IMGMASK;
        a->image[i] = other | (idim << shift);
      }
    }
  }
}

/* verlet.cpp:229-363 with fix nve + fix obmd + pair dpd, one rank, re-neighbour
   every step (Q5): initial_integrate, pre_exchange, pbc, force_clear + pair,
   post_force, final_integrate. */
void orc_step(orc_fix *p, orc_atoms *a, int rng_mode, unsigned step, orc_trial *trials,
              int maxtrials)
{
  orc_nve_initial(p, a);
  orc_fix_pre_exchange(p, a, trials, maxtrials);
  orc_pbc(p, a);
  orc_pair_compute(p, a, rng_mode, step, NULL, NULL);
  orc_fix_post_force(p, a);
  orc_nve_final(p, a);
}

/* ---- bonded forces ------------------------------------------------------------------------------ */

/* MOLECULE/bond_harmonic.cpp:48-103.  The reference walks neighbor->bondlist, which NTopoBondAll::build
   (ntopo_bond_all.cpp:36-77) fills in local-index order with the partner's closest image; with
   newton_bond on every bond is stored once, on the atom the data file lists first. */
double orc_bond_harmonic(const orc_fix *p, orc_atoms *a, int nbonds, const int *btype, const int *tag1,
                         const int *tag2, const double *k, const double *r0)
{
  int maxtag = 0;
  for (int i = 0; i < a->nlocal; i++)
    if (a->tag[i] > maxtag) maxtag = a->tag[i];
  int *map = (int *) malloc(sizeof(int) * (size_t) (maxtag + 1));
  for (int t = 0; t <= maxtag; t++) map[t] = -1;
  for (int i = 0; i < a->nlocal; i++) map[a->tag[i]] = i;
  const double prd[3] = {0.0, p->boxhi[1] - p->boxlo[1], p->boxhi[2] - p->boxlo[2]};
  double etot = 0.0;
  for (int n = 0; n < nbonds; n++) {
    if (tag1[n] > maxtag || tag2[n] > maxtag) continue;
    const int i1 = map[tag1[n]], i2 = map[tag2[n]];
    if (i1 < 0 || i2 < 0) continue;
    double del[3];
    for (int d = 0; d < 3; d++) del[d] = a->x[3 * i1 + d] - a->x[3 * i2 + d];
    for (int d = 1; d < 3; d++) { /* Domain::closest_image (domain.cpp:1148-1172) */
      if (del[d] > 0.5 * prd[d]) del[d] -= prd[d];
      else if (del[d] < -0.5 * prd[d]) del[d] += prd[d];
    }
    const double rsq = del[0] * del[0] + del[1] * del[1] + del[2] * del[2];
    const double r = sqrt(rsq);
    const double dr = r - r0[btype[n]];
    const double rk = k[btype[n]] * dr;
    const double fbond = (r > 0.0) ? -2.0 * rk / r : 0.0;
    etot += rk * dr;
    for (int d = 0; d < 3; d++) {
      a->f[3 * i1 + d] += del[d] * fbond;
      a->f[3 * i2 + d] -= del[d] * fbond;
    }
  }
  free(map);
  return etot;
}

/* MOLECULE/bond_fene.cpp:53-135 */
double orc_bond_fene(const orc_fix *p, orc_atoms *a, int nbonds, const int *btype, const int *tag1,
                     const int *tag2, const double *k, const double *r0, const double *epsilon,
                     const double *sigma)
{
  const double CUBEROOT2 = 1.25992104989487316476; /* math_const.h MY_CUBEROOT2 */
  int maxtag = 0;
  for (int i = 0; i < a->nlocal; i++)
    if (a->tag[i] > maxtag) maxtag = a->tag[i];
  int *map = (int *) malloc(sizeof(int) * (size_t) (maxtag + 1));
  for (int t = 0; t <= maxtag; t++) map[t] = -1;
  for (int i = 0; i < a->nlocal; i++) map[a->tag[i]] = i;
  const double prd[3] = {0.0, p->boxhi[1] - p->boxlo[1], p->boxhi[2] - p->boxlo[2]};
  double etot = 0.0;
  for (int n = 0; n < nbonds; n++) {
    if (tag1[n] > maxtag || tag2[n] > maxtag) continue;
    const int i1 = map[tag1[n]], i2 = map[tag2[n]], ty = btype[n];
    if (i1 < 0 || i2 < 0) continue;
    double del[3];
    for (int d = 0; d < 3; d++) del[d] = a->x[3 * i1 + d] - a->x[3 * i2 + d];
    for (int d = 1; d < 3; d++) {
      if (del[d] > 0.5 * prd[d]) del[d] -= prd[d];
      else if (del[d] < -0.5 * prd[d]) del[d] += prd[d];
    }
    const double rsq = del[0] * del[0] + del[1] * del[1] + del[2] * del[2];
    const double r0sq = r0[ty] * r0[ty];
    double rlogarg = 1.0 - rsq / r0sq;
    if (rlogarg < 0.1) { /* :91-96 */
      if (rlogarg <= -3.0) { free(map); return -1.0e300; }
      rlogarg = 0.1;
    }
    double fbond = -k[ty] / rlogarg, sr6 = 0.0;
    const int lj = rsq < CUBEROOT2 * sigma[ty] * sigma[ty];
    if (lj) {
      const double sr2 = sigma[ty] * sigma[ty] / rsq;
      sr6 = sr2 * sr2 * sr2;
      fbond += 48.0 * epsilon[ty] * sr6 * (sr6 - 0.5) / rsq;
    }
    double ebond = -0.5 * k[ty] * r0sq * log(rlogarg);
    if (lj) ebond += 4.0 * epsilon[ty] * sr6 * (sr6 - 1.0) + epsilon[ty];
    etot += ebond;
    for (int d = 0; d < 3; d++) {
      a->f[3 * i1 + d] += del[d] * fbond;
      a->f[3 * i2 + d] -= del[d] * fbond;
    }
  }
  free(map);
  return etot;
}
