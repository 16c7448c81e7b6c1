// Slab decomposition along the open axis: replaces CommBrick::exchange / borders
// (comm_brick.cpp:652-978) and the fix's MPI collectives with NCCL over NVLink.
//
// Rank r owns a contiguous range of cell columns (x is the slowest cell index, so that range is one
// contiguous slot range).  Per step, after the cell sort:
//   1. counts   : {migrants, boundary-column population} to each x-neighbour      (2 ints each way)
//   2. migration: atoms whose column left the slab -> neighbour; arrivals are appended as "extras"
//   3. halo     : boundary column (already contiguous and cell-sorted) + the arrivals that came
//                 from that side -> neighbour, which appends them as ghost extras and threads them
//                 into the per-cell lists the pair kernel and USHER walk.
// No reverse (force) halo: the pair kernel is a full-neighbour kernel with a symmetric RNG.
// NCCL is loaded with dlopen so the library still loads on hosts without it (single-GPU use).
#include <dlfcn.h>
#include <nccl.h>

#include "ctx.h"

namespace {
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

bool nccl_load(obmd_ctx *c)
{
  if (g_nccl.handle) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names)
    if ((g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!g_nccl.handle) {
    obmd_fail(c, OBMD_ERR_STATE, "cannot load libnccl.so.2: %s", dlerror());
    return false;
  }
#define SYM(field, name) *(void **) (&g_nccl.field) = dlsym(g_nccl.handle, name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.AllReduce) {
    obmd_fail(c, OBMD_ERR_STATE, "libnccl lacks a required symbol");
    return false;
  }
  return true;
}
}    // namespace

#define NK(call)                                                                                  \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != ncclSuccess)                                                                        \
      return obmd_fail(c, OBMD_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,     \
                       g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error");         \
  } while (0)

extern "C" int obmd_comm_unique_id(char *out128)
{
  obmd_ctx *c = nullptr;
  if (!out128) return obmd_fail(c, OBMD_ERR_ARG, "obmd_comm_unique_id: NULL buffer");
  if (!nccl_load(c)) return OBMD_ERR_STATE;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NK(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return OBMD_OK;
}

int setup_grid(obmd_ctx *c);  // api.cu

extern "C" int obmd_comm_init(obmd_ctx *c, int rank, int nranks, const char *id128)
{
  NEED(c && id128, OBMD_ERR_ARG, "obmd_comm_init: NULL argument");
  NEED(nranks >= 1 && rank >= 0 && rank < nranks, OBMD_ERR_ARG, "obmd_comm_init: bad rank %d of %d", rank, nranks);
  NEED(c->n == 0, OBMD_ERR_STATE, "obmd_comm_init: call it before obmd_upload");
  if (nranks > 1) {
    if (!nccl_load(c)) return OBMD_ERR_STATE;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    NK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->nccl = comm;
  }
  c->myrank = rank;
  c->nranks = nranks;
  if (c->pair_ready) return setup_grid(c);
  return OBMD_OK;
}

extern "C" int obmd_comm_finalize(obmd_ctx *c)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  if (c->nccl) {
    cudaStreamSynchronize(c->stream);
    g_nccl.CommDestroy((ncclComm_t) c->nccl);
    c->nccl = nullptr;
  }
  return OBMD_OK;
}

// first nsum doubles summed, the next nmax doubles max-reduced, in place on the device
int comm_allreduce(obmd_ctx *c, double *dev, int nsum, int nmax)
{
  if (c->nranks == 1) return OBMD_OK;
  ncclComm_t comm = (ncclComm_t) c->nccl;
  if (nsum > 0) NK(g_nccl.AllReduce(dev, dev, nsum, ncclDouble, ncclSum, comm, c->stream));
  if (nmax > 0) NK(g_nccl.AllReduce(dev + nsum, dev + nsum, nmax, ncclDouble, ncclMax, comm, c->stream));
  c->launches += (nsum > 0) + (nmax > 0);
  return OBMD_OK;
}

// ---- pack / unpack ---------------------------------------------------------------------------

#define MIG_W 8   // doubles per migrating atom: x y z vx vy vz (tag,type) image
#define HALO_W 7  // doubles per ghost: x y z vx vy vz (tag,type)

// atoms [src, src+n) -> records of width W
__global__ void k_pack(Planes p, int src, int n, int W, double *__restrict__ buf)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int s = src + k;
  double *r = buf + (size_t) k * W;
  r[0] = p.x[s]; r[1] = p.y[s]; r[2] = p.z[s];
  r[3] = p.vx[s]; r[4] = p.vy[s]; r[5] = p.vz[s];
  r[6] = __hiloint2double(p.tag[s], p.type[s]);
  if (W > 7) r[7] = __hiloint2double(0, p.image[s]);
}

__global__ void k_unpack(Planes p, double *fx, double *fy, double *fz, int dst, int n, int W,
                         const double *__restrict__ buf)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int s = dst + k;
  const double *r = buf + (size_t) k * W;
  p.x[s] = r[0]; p.y[s] = r[1]; p.z[s] = r[2];
  p.vx[s] = r[3]; p.vy[s] = r[4]; p.vz[s] = r[5];
  p.tag[s] = __double2hiint(r[6]);
  p.type[s] = __double2loint(r[6]);
  p.image[s] = (W > 7) ? __double2loint(r[7]) : IMG0;
  p.lidx[s] = s;
  fx[s] = fy[s] = fz[s] = 0.0;
}

static int ensure_comm(obmd_ctx *c, int natoms)
{
  if (natoms <= c->comm_cap) return OBMD_OK;
  cudaStreamSynchronize(c->stream);
  cudaFree(c->sendL); cudaFree(c->sendR); cudaFree(c->recvL); cudaFree(c->recvR);
  c->comm_cap = natoms + natoms / 4 + 4096;
  size_t bytes = sizeof(double) * MIG_W * (size_t) c->comm_cap;
  CK(cudaMalloc(&c->sendL, bytes));
  CK(cudaMalloc(&c->sendR, bytes));
  CK(cudaMalloc(&c->recvL, bytes));
  CK(cudaMalloc(&c->recvR, bytes));
  return OBMD_OK;
}

// Migration + halo for one step.  Must run after cells_build() and fix_pre_begin() (the arrivals are
// stored where the parked atoms were).
int comm_exchange(obmd_ctx *c)
{
  if (c->nranks == 1) return OBMD_OK;
  ScopedTimer t(c, "comm");
  ncclComm_t comm = (ncclComm_t) c->nccl;
  const int left = c->myrank - 1, right = c->myrank + 1;
  const bool hasL = left >= 0, hasR = right < c->nranks;
  const int base = c->nsorted;
  NEED(hasL || c->nmigL == 0, OBMD_ERR_STATE, "slab run: atoms migrating left of rank 0");
  NEED(hasR || c->nmigR == 0, OBMD_ERR_STATE, "slab run: atoms migrating right of the last rank");
  const int ncolA = c->col_ab[1] - c->col_ab[0], ncolB = c->col_ab[3] - c->col_ab[2];
  int rc;

  // 1. counts -------------------------------------------------------------------------------
  int *cnt_h = reinterpret_cast<int *>(c->red_h) + 32;  // [0..1] to left, [2..3] to right, [4..7] received
  int *cnt_d = reinterpret_cast<int *>(c->red_d) + 32;
  cnt_h[0] = c->nmigL; cnt_h[1] = ncolA; cnt_h[2] = c->nmigR; cnt_h[3] = ncolB;
  cnt_h[4] = cnt_h[5] = cnt_h[6] = cnt_h[7] = 0;
  CK(cudaMemcpyAsync(cnt_d, cnt_h, 8 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NK(g_nccl.GroupStart());
  if (hasL) {
    NK(g_nccl.Send(cnt_d + 0, 2, ncclInt, left, comm, c->stream));
    NK(g_nccl.Recv(cnt_d + 4, 2, ncclInt, left, comm, c->stream));
  }
  if (hasR) {
    NK(g_nccl.Send(cnt_d + 2, 2, ncclInt, right, comm, c->stream));
    NK(g_nccl.Recv(cnt_d + 6, 2, ncclInt, right, comm, c->stream));
  }
  NK(g_nccl.GroupEnd());
  CK(cudaMemcpyAsync(cnt_h + 4, cnt_d + 4, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const int narrL = hasL ? cnt_h[4] : 0, nbrcolL = hasL ? cnt_h[5] : 0;  // from the left neighbour
  const int narrR = hasR ? cnt_h[6] : 0, nbrcolR = hasR ? cnt_h[7] : 0;
  const int ngL = hasL ? nbrcolL + c->nmigL : 0;  // its boundary column + what I just sent it
  const int ngR = hasR ? nbrcolR + c->nmigR : 0;
  const int hsendL = ncolA + narrL, hsendR = ncolB + narrR;
  int need = c->nmigL;
  for (int v : {c->nmigR, narrL, narrR, ngL, ngR, hsendL, hsendR}) need = v > need ? v : need;
  if ((rc = ensure_comm(c, need))) return rc;
  if ((rc = obmd_ensure_capacity(c, base + narrL + narrR + ngL + ngR + 4096))) return rc;

  // 2. migration ----------------------------------------------------------------------------
  if (c->nmigL) LAUNCH(k_pack, cdiv(c->nmigL, 128), 128, 0, c->cur, c->seg_migL, c->nmigL, MIG_W, c->sendL);
  if (c->nmigR) LAUNCH(k_pack, cdiv(c->nmigR, 128), 128, 0, c->cur, c->seg_migR, c->nmigR, MIG_W, c->sendR);
  NK(g_nccl.GroupStart());
  if (hasL) {
    if (c->nmigL) NK(g_nccl.Send(c->sendL, (size_t) c->nmigL * MIG_W, ncclDouble, left, comm, c->stream));
    if (narrL) NK(g_nccl.Recv(c->recvL, (size_t) narrL * MIG_W, ncclDouble, left, comm, c->stream));
  }
  if (hasR) {
    if (c->nmigR) NK(g_nccl.Send(c->sendR, (size_t) c->nmigR * MIG_W, ncclDouble, right, comm, c->stream));
    if (narrR) NK(g_nccl.Recv(c->recvR, (size_t) narrR * MIG_W, ncclDouble, right, comm, c->stream));
  }
  NK(g_nccl.GroupEnd());
  if (narrL) LAUNCH(k_unpack, cdiv(narrL, 128), 128, 0, c->cur, c->fx, c->fy, c->fz, base, narrL, MIG_W, c->recvL);
  if (narrR)
    LAUNCH(k_unpack, cdiv(narrR, 128), 128, 0, c->cur, c->fx, c->fy, c->fz, base + narrL, narrR, MIG_W, c->recvR);
  c->narr = narrL + narrR;

  // 3. halo -----------------------------------------------------------------------------------
  if (hasL) {
    if (ncolA) LAUNCH(k_pack, cdiv(ncolA, 128), 128, 0, c->cur, c->col_ab[0], ncolA, HALO_W, c->sendL);
    if (narrL)
      LAUNCH(k_pack, cdiv(narrL, 128), 128, 0, c->cur, base, narrL, HALO_W, c->sendL + (size_t) ncolA * HALO_W);
  }
  if (hasR) {
    if (ncolB) LAUNCH(k_pack, cdiv(ncolB, 128), 128, 0, c->cur, c->col_ab[2], ncolB, HALO_W, c->sendR);
    if (narrR)
      LAUNCH(k_pack, cdiv(narrR, 128), 128, 0, c->cur, base + narrL, narrR, HALO_W,
             c->sendR + (size_t) ncolB * HALO_W);
  }
  NK(g_nccl.GroupStart());
  if (hasL) {
    if (hsendL) NK(g_nccl.Send(c->sendL, (size_t) hsendL * HALO_W, ncclDouble, left, comm, c->stream));
    if (ngL) NK(g_nccl.Recv(c->recvL, (size_t) ngL * HALO_W, ncclDouble, left, comm, c->stream));
  }
  if (hasR) {
    if (hsendR) NK(g_nccl.Send(c->sendR, (size_t) hsendR * HALO_W, ncclDouble, right, comm, c->stream));
    if (ngR) NK(g_nccl.Recv(c->recvR, (size_t) ngR * HALO_W, ncclDouble, right, comm, c->stream));
  }
  NK(g_nccl.GroupEnd());
  c->launches += 6;
  const int g0 = base + c->narr;
  if (ngL) LAUNCH(k_unpack, cdiv(ngL, 128), 128, 0, c->cur, c->fx, c->fy, c->fz, g0, ngL, HALO_W, c->recvL);
  if (ngR) LAUNCH(k_unpack, cdiv(ngR, 128), 128, 0, c->cur, c->fx, c->fy, c->fz, g0 + ngL, ngR, HALO_W, c->recvR);
  c->nghost = ngL + ngR;
  c->nextra = c->narr + c->nghost;
  c->n = base + c->nextra;

  // 4. per-cell lists: cell-sorted ghost columns first, then the small unsorted groups in front
  if ((rc = fix_link_extras(c, c->narr, nbrcolL, true))) return rc;
  if ((rc = fix_link_extras(c, c->narr + ngL, nbrcolR, true))) return rc;
  if ((rc = fix_link_extras(c, 0, c->narr, false))) return rc;
  if ((rc = fix_link_extras(c, c->narr + nbrcolL, ngL - nbrcolL, false))) return rc;
  if ((rc = fix_link_extras(c, c->narr + ngL + nbrcolR, ngR - nbrcolR, false))) return rc;
  CK(cudaGetLastError());
  return OBMD_OK;
}
