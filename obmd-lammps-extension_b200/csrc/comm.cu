// Slab-decomposition pack/unpack kernels (replace CommBrick::exchange/borders,
// comm_brick.cpp:652-978).  The host moves the packed device buffers with NCCL.
#include "ctx.h"

extern "C" int obmd_migrate_pack(obmd_ctx *c, double *, double *, int, int *, int *)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  return obmd_fail(c, OBMD_ERR_STATE, "obmd_migrate_pack: slab decomposition not initialised");
}
extern "C" int obmd_migrate_unpack(obmd_ctx *c, const double *, int)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  return obmd_fail(c, OBMD_ERR_STATE, "obmd_migrate_unpack: slab decomposition not initialised");
}
extern "C" int obmd_halo_pack(obmd_ctx *c, double *, double *, int, int *, int *)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  return obmd_fail(c, OBMD_ERR_STATE, "obmd_halo_pack: slab decomposition not initialised");
}
extern "C" int obmd_halo_unpack(obmd_ctx *c, const double *, int, const double *, int)
{
  NEED(c, OBMD_ERR_ARG, "NULL context");
  return obmd_fail(c, OBMD_ERR_STATE, "obmd_halo_unpack: slab decomposition not initialised");
}
