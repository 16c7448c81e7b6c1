// Bonded forces: BondHarmonic::compute (MOLECULE/bond_harmonic.cpp:48-103) and BondFENE::compute
// (MOLECULE/bond_fene.cpp:53-135) over the bond list
// NTopoBondAll::build (ntopo_bond_all.cpp:36-77) would make -- the bonded path of BASELINE config 4
// (SURVEY 8f N3).  Molecule-aware insertion / deletion is NOT here: the topology is static.
//
// The reference stores every bond once (newton_bond on), walks the list and scatters +-f to both atoms.
// Here every atom owns its force: the topology is kept in CSR form indexed by atom TAG (both directions,
// partners sorted by tag), the sort does not have to move it, and one thread per atom sums the terms of
// its own bonds in that fixed order -- no atomics, deterministic, like the pair kernels.  A tag -> slot
// map (Atom::map) is rebuilt for the current storage order by every call.
// Compiled with -fmad=false: r, dr, rk and fbond are evaluated exactly as the reference evaluates them.
#include <algorithm>
#include <vector>

#include "ctx.h"

struct BondTab {
  int ntypes, style;  // style 0: harmonic (k, r0); 1: fene (k, r0, eps, sig)
  double k[MAXT + 1], r0[MAXT + 1], eps[MAXT + 1], sig[MAXT + 1];
};

__global__ void k_tag_map_clear(int *__restrict__ map, int n)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) map[t] = -1;
}

// Atom::map_set (atom_map.cpp:141-208, array style): slot of every live tag (sorted atoms, then the atoms
// appended after the sort; parked atoms -- outside the box, about to be deleted -- are not mapped)
__global__ void k_tag_map_set(Planes p, int nsorted, Extras ex, int maxtag, int *__restrict__ map)
{
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  int i = q < nsorted ? q : extras_slot(ex, 2, q - nsorted);
  if (i < 0) return;
  if (q >= nsorted && q - nsorted >= ex.narr && q - nsorted < ex.narr + ex.nghost) return;  // ghosts are not owned
  const int t = p.tag[i];
  if (t >= 0 && t <= maxtag) map[t] = i;
}

__global__ void __launch_bounds__(256)
    k_bond_harmonic(Planes p, int nsorted, Extras ex, Grid g, BondTab bt, int maxtag, const int *__restrict__ off,
                    const int *__restrict__ partner, const int *__restrict__ ptype, const int *__restrict__ map,
                    double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
                    double *__restrict__ partials, StepScalars *sc)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  int i = q < nsorted ? q : extras_slot(ex, 1, q - nsorted);  // owned atoms: sorted + arrivals + insertions
  double e = 0.0, v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, v4 = 0.0, v5 = 0.0;
  if (i >= 0) {
    const int t = p.tag[i];
    if (t >= 0 && t <= maxtag) {
      const int b0 = off[t], b1 = off[t + 1];
      if (b1 > b0) {
        const double xi = p.x[i], yi = p.y[i], zi = p.z[i];
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int b = b0; b < b1; b++) {
          const int j = map[partner[b]];
          if (j < 0) {  // the partner was deleted: molecule-aware deletion is not implemented
            sc->usher_error = OBMD_ERR_STATE;
            continue;
          }
          double delx = xi - p.x[j], dely = yi - p.y[j], delz = zi - p.z[j];
          // Domain::closest_image (domain.cpp:1148-1172) for boundary f p p
          if (dely > g.half[1]) dely -= g.prd[1];
          else if (dely < -g.half[1]) dely += g.prd[1];
          if (delz > g.half[2]) delz -= g.prd[2];
          else if (delz < -g.half[2]) delz += g.prd[2];
          const double rsq = delx * delx + dely * dely + delz * delz;
          const int ty = ptype[b];
          double fbond, ebond;
          if (bt.style == 0) {  // bond_harmonic.cpp:73-85
            const double r = sqrt(rsq);
            const double dr = r - bt.r0[ty];
            const double rk = bt.k[ty] * dr;
            fbond = (r > 0.0) ? -2.0 * rk / r : 0.0;
            ebond = rk * dr;
          } else {  // bond_fene.cpp:81-117
            const double r0sq = bt.r0[ty] * bt.r0[ty];
            double rlogarg = 1.0 - rsq / r0sq;
            if (rlogarg < 0.1) {  // "FENE bond too long": the reference warns and clamps, aborts beyond 2 r0
              if (rlogarg <= -3.0) sc->usher_error = OBMD_ERR_NONFINITE;
              rlogarg = 0.1;
            }
            fbond = -bt.k[ty] / rlogarg;
            double sr6 = 0.0;
            const bool lj = rsq < 1.25992104989487316476 * bt.sig[ty] * bt.sig[ty];  // MY_CUBEROOT2
            if (lj) {
              const double sr2 = bt.sig[ty] * bt.sig[ty] / rsq;
              sr6 = sr2 * sr2 * sr2;
              fbond += 48.0 * bt.eps[ty] * sr6 * (sr6 - 0.5) / rsq;
            }
            ebond = -0.5 * bt.k[ty] * r0sq * log(rlogarg);
            if (lj) ebond += 4.0 * bt.eps[ty] * sr6 * (sr6 - 1.0) + bt.eps[ty];
          }
          ax += delx * fbond;
          ay += dely * fbond;
          az += delz * fbond;
          e += 0.5 * ebond;  // each bond is seen from both of its atoms
          if (partials) {        // Bond::ev_tally (bond.cpp:174-240), half per atom
            const double h = 0.5 * fbond;
            v0 += h * delx * delx;
            v1 += h * dely * dely;
            v2 += h * delz * delz;
            v3 += h * delx * dely;
            v4 += h * delx * delz;
            v5 += h * dely * delz;
          }
        }
        fx[i] += ax;
        fy[i] += ay;
        fz[i] += az;
      }
    }
  }
  if (partials) {  // deterministic block reduction of the energy and the virial
    __shared__ double sh[7][8];
    double v[7] = {e, v0, v1, v2, v3, v4, v5};
#pragma unroll
    for (int k = 0; k < 7; k++) {
      double s = v[k];
      for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
      if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
      double s = 0.0;
      for (int w = 0; w < 8; w++) s += sh[threadIdx.x][w];
      partials[blockIdx.x * 7 + threadIdx.x] = s;
    }
  }
}

// fixed-order final sums: out[0] = energy, out[1..6] = virial
__global__ void k_bond_energy_final(const double *__restrict__ partials, int nb, double *__restrict__ out)
{
  __shared__ double sh[256];
  for (int k = 0; k < 7; k++) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) s += partials[b * 7 + k];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = blockDim.x / 2; d > 0; d >>= 1) {
      if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = sh[0];
    __syncthreads();
  }
}

void bond_free(obmd_ctx *c)
{
  cudaFree(c->bond_off); cudaFree(c->bond_partner); cudaFree(c->bond_ptype); cudaFree(c->tag_map);
  cudaFree(c->bond_e);
  c->bond_off = c->bond_partner = c->bond_ptype = c->tag_map = nullptr;
  c->bond_e = nullptr;
  c->nbonds = 0;
}

extern "C" int obmd_bond_style_harmonic(obmd_ctx *c, int nbondtypes, const double *k, const double *r0)
{
  NEED(c && k && r0, OBMD_ERR_ARG, "obmd_bond_style_harmonic: NULL argument");
  NEED(nbondtypes >= 1 && nbondtypes <= MAXT, OBMD_ERR_ARG, "obmd_bond_style_harmonic: 1 <= nbondtypes <= %d", MAXT);
  c->nbondtypes = nbondtypes;
  c->bond_style = 0;
  for (int t = 1; t <= nbondtypes; t++) {
    c->bond_k[t] = k[t];
    c->bond_r0[t] = r0[t];
  }
  return OBMD_OK;
}

extern "C" int obmd_bond_style_fene(obmd_ctx *c, int nbondtypes, const double *k, const double *r0,
                                    const double *epsilon, const double *sigma)
{
  NEED(c && k && r0 && epsilon && sigma, OBMD_ERR_ARG, "obmd_bond_style_fene: NULL argument");
  NEED(nbondtypes >= 1 && nbondtypes <= MAXT, OBMD_ERR_ARG, "obmd_bond_style_fene: 1 <= nbondtypes <= %d", MAXT);
  c->nbondtypes = nbondtypes;
  c->bond_style = 1;
  for (int t = 1; t <= nbondtypes; t++) {
    c->bond_k[t] = k[t];
    c->bond_r0[t] = r0[t];
    c->bond_eps[t] = epsilon[t];
    c->bond_sig[t] = sigma[t];
  }
  return OBMD_OK;
}

extern "C" int obmd_upload_bonds(obmd_ctx *c, int nbonds, const int *type, const int *tag1, const int *tag2)
{
  NEED(c && (nbonds == 0 || (type && tag1 && tag2)), OBMD_ERR_ARG, "obmd_upload_bonds: NULL argument");
  NEED(c->nbondtypes > 0, OBMD_ERR_STATE, "obmd_upload_bonds: call obmd_bond_style_harmonic first");
  bond_free(c);
  if (nbonds == 0) return OBMD_OK;
  int maxtag = 0;
  for (int n = 0; n < nbonds; n++) {
    NEED(type[n] >= 1 && type[n] <= c->nbondtypes, OBMD_ERR_ARG, "Bond type %d of bond %d is invalid", type[n], n + 1);
    NEED(tag1[n] > 0 && tag2[n] > 0 && tag1[n] != tag2[n], OBMD_ERR_ARG, "Bond atoms of bond %d are invalid", n + 1);
    maxtag = std::max(maxtag, std::max(tag1[n], tag2[n]));
  }
  // CSR by tag, both directions, partners of an atom in ascending tag (a fixed summation order)
  std::vector<int> off(maxtag + 2, 0);
  for (int n = 0; n < nbonds; n++) {
    off[tag1[n] + 1]++;
    off[tag2[n] + 1]++;
  }
  for (int t = 0; t <= maxtag; t++) off[t + 1] += off[t];
  std::vector<std::pair<int, int>> ent(2 * (size_t) nbonds);
  std::vector<int> fill(off.begin(), off.end() - 1);
  for (int n = 0; n < nbonds; n++) {
    ent[fill[tag1[n]]++] = {tag2[n], type[n]};
    ent[fill[tag2[n]]++] = {tag1[n], type[n]};
  }
  std::vector<int> partner(2 * (size_t) nbonds), ptype(2 * (size_t) nbonds);
  for (int t = 0; t <= maxtag; t++) {
    std::sort(ent.begin() + off[t], ent.begin() + off[t + 1]);
    for (int b = off[t]; b < off[t + 1]; b++) {
      partner[b] = ent[b].first;
      ptype[b] = ent[b].second;
    }
  }
  CK(cudaMalloc(&c->bond_off, sizeof(int) * off.size()));
  CK(cudaMalloc(&c->bond_partner, sizeof(int) * partner.size()));
  CK(cudaMalloc(&c->bond_ptype, sizeof(int) * ptype.size()));
  CK(cudaMalloc(&c->tag_map, sizeof(int) * (size_t) (maxtag + 1)));
  CK(cudaMalloc(&c->bond_e, 7 * sizeof(double)));
  CK(cudaMemcpyAsync(c->bond_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->bond_partner, partner.data(), sizeof(int) * partner.size(), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->bond_ptype, ptype.data(), sizeof(int) * ptype.size(), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(c->bond_e, 0, 7 * sizeof(double), c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->nbonds = nbonds;
  c->bond_maxtag = maxtag;
  return OBMD_OK;
}

extern "C" int obmd_bond_compute(obmd_ctx *c, int eflag)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  NEED(c->n > 0, OBMD_ERR_STATE, "obmd_bond_compute: no atoms uploaded");
  if (c->nbonds == 0) return OBMD_OK;
  NEED(c->nparked == 0 || c->cells_valid, OBMD_ERR_STATE, "obmd_bond_compute: cell list is stale");
  ScopedTimer t(c, "bond");
  // before the first cell build every atom is "sorted" storage: [0, n)
  const int nsorted = c->cells_valid ? c->nsorted : c->n;
  const Extras ex = c->cells_valid ? extras_of(c) : Extras{c->n, 0, 0, 0};
  const int nall = nsorted + ex.narr + ex.nghost + ex.nins, nown = nsorted + ex.narr + ex.nins;
  BondTab bt;
  bt.ntypes = c->nbondtypes;
  bt.style = c->bond_style;
  for (int k = 0; k <= MAXT; k++) {
    const bool on = k >= 1 && k <= c->nbondtypes;
    bt.k[k] = on ? c->bond_k[k] : 0.0;
    bt.r0[k] = on ? c->bond_r0[k] : 0.0;
    bt.eps[k] = (on && c->bond_style == 1) ? c->bond_eps[k] : 0.0;
    bt.sig[k] = (on && c->bond_style == 1) ? c->bond_sig[k] : 0.0;
  }
  const int nb = cdiv(nown, 256);
  if (eflag && nb * 7 > c->partial_cap) {
    if (c->partials) cudaFree(c->partials);
    c->partial_cap = nb * 7 + 1024;
    CK(cudaMalloc(&c->partials, sizeof(double) * c->partial_cap));
  }
  CK(cudaMemsetAsync(&c->sc_d->usher_error, 0, sizeof(int), c->stream));
  LAUNCH(k_tag_map_clear, cdiv(c->bond_maxtag + 1, 256), 256, 0, c->tag_map, c->bond_maxtag + 1);
  LAUNCH(k_tag_map_set, cdiv(nall, 256), 256, 0, c->cur, nsorted, ex, c->bond_maxtag, c->tag_map);
  LAUNCH(k_bond_harmonic, nb, 256, 0, c->cur, nsorted, ex, c->grid, bt, c->bond_maxtag, c->bond_off, c->bond_partner,
         c->bond_ptype, c->tag_map, c->fx, c->fy, c->fz, eflag ? c->partials : nullptr, c->sc_d);
  if (eflag) LAUNCH(k_bond_energy_final, 1, 256, 0, c->partials, nb, c->bond_e);
  CK(cudaGetLastError());
  int err = 0;
  CK(cudaMemcpyAsync(&err, &c->sc_d->usher_error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  NEED(err != OBMD_ERR_NONFINITE, OBMD_ERR_NONFINITE, "Bad FENE bond");  // bond_fene.cpp:94
  NEED(err == 0, OBMD_ERR_STATE,
       "Bond atoms missing: a bonded atom was deleted (molecule-aware deletion, fix_obmd_merged.cpp:709-821, is not implemented)");
  return OBMD_OK;
}

extern "C" int obmd_bond_tallies(obmd_ctx *c, double *ebond, double virial[6])
{
  NEED(c && ebond, OBMD_ERR_ARG, "obmd_bond_tallies: NULL argument");
  double h[7] = {0, 0, 0, 0, 0, 0, 0};
  if (c->nbonds > 0) {
    CK(cudaMemcpyAsync(h, c->bond_e, 7 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  *ebond = h[0];
  if (virial)
    for (int k = 0; k < 6; k++) virial[k] = h[1 + k];
  return OBMD_OK;
}
