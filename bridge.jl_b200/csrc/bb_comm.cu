/*
 * bb_comm.cu -- the one collective of the path: the all-reduce(sum) of the acceptance counter that the reference's
 * `acc += 1` bookkeeping (test/partialbridgenuH.jl:189) turns into when chains are sharded over GPUs.
 *
 * Chains never interact, so there is no data-path collective.  The counter is 8 bytes: nothing to fuse; what matters is
 * that the exchange never sits on the compute stream.  bb_allreduce_acc snapshots the counter with an 8-byte copy on
 * the compute stream (ordered after the iteration's kernel), and runs ncclAllReduce on the communicator's own side
 * stream behind an event; the next iteration's kernel starts immediately.  A ring of snapshot slots lets the side
 * stream lag several iterations.
 *
 * NCCL is bound at run time (dlopen "libnccl.so.2"; BB_NCCL_LIB overrides): a single-GPU host never needs it, and a
 * process that has already loaded NCCL (e.g. through torch) shares that copy.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "bb_host.h"

namespace {
typedef struct ncclComm* nccl_comm_t;
typedef struct { char internal[BB_NCCL_ID_BYTES]; } nccl_uid; /* ncclUniqueId: 128 opaque bytes */
enum { NCCL_UINT64 = 5, NCCL_SUM = 0 };                        /* ncclDataType_t / ncclRedOp_t values of nccl.h */

struct nccl_api {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
nccl_api g_nccl;
char g_nccl_err[256] = "";

int nccl_load() {
  if (g_nccl.handle) return BB_OK;
  const char* names[] = {getenv("BB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    snprintf(g_nccl_err, sizeof(g_nccl_err), "NCCL not found: %s", dlerror());
    return BB_ERR_UNSUPPORTED;
  }
  nccl_api a;
  a.handle = h;
  a.GetUniqueId = (int (*)(nccl_uid*))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_uid, int))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
  a.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
  a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  a.GetVersion = (int (*)(int*))dlsym(h, "ncclGetVersion");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce) {
    snprintf(g_nccl_err, sizeof(g_nccl_err), "NCCL library lacks a required symbol");
    return BB_ERR_UNSUPPORTED;
  }
  g_nccl = a;
  return BB_OK;
}
int nccl_check(int rc, const char* where) {
  if (rc == 0) return BB_OK;
  snprintf(g_nccl_err, sizeof(g_nccl_err), "%s: %s", where, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
  return BB_ERR_COMM;
}
}  // namespace

#define BB_COMM_SLOTS 8
struct bb_comm {
  bb_ctx* ctx = nullptr;
  nccl_comm_t comm = nullptr;
  bool own_comm = false;
  int nranks = 1, rank = 0;
  cudaStream_t side = nullptr;
  unsigned long long* slots = nullptr; /* [BB_COMM_SLOTS][2]: snapshot of the local counter, its global sum */
  cudaEvent_t snap[BB_COMM_SLOTS] = {nullptr}, done[BB_COMM_SLOTS] = {nullptr};
  int64_t issued = 0;                  /* all-reduces issued so far */
  bool counted = false;                /* registered in ctx->live */
};

extern "C" const char* bb_comm_last_error(void) { return g_nccl_err; }

extern "C" int bb_comm_unique_id(uint8_t* id) {
  if (!id) return BB_ERR_ARG;
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid u;
  rc = nccl_check(g_nccl.GetUniqueId(&u), "ncclGetUniqueId");
  if (rc) return rc;
  memcpy(id, u.internal, BB_NCCL_ID_BYTES);
  return BB_OK;
}

static int comm_finish(bb_ctx* ctx, bb_comm* c, bb_comm** out) {
  cudaError_t e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&c->slots, sizeof(unsigned long long) * 2 * BB_COMM_SLOTS);
  if (e == cudaSuccess) e = cudaMemset(c->slots, 0, sizeof(unsigned long long) * 2 * BB_COMM_SLOTS);
  for (int i = 0; i < BB_COMM_SLOTS && e == cudaSuccess; i++) {
    e = cudaEventCreateWithFlags(&c->snap[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    bb_set_cuda_error(e, "bb_comm_create");
    bb_comm_destroy(c);
    return e == cudaErrorMemoryAllocation ? BB_ERR_NOMEM : BB_ERR_CUDA;
  }
  ctx->live++;
  c->counted = true;
  *out = c;
  return BB_OK;
}

extern "C" int bb_comm_create(bb_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id, bb_comm** out) {
  if (!out) return BB_ERR_ARG;
  *out = nullptr;
  if (!ctx) return BB_ERR_NODEVICE;
  if (!id || nranks < 1 || rank < 0 || rank >= nranks) return BB_ERR_ARG;
  int rc = nccl_load();
  if (rc) return rc;
  BB_CUDA(cudaSetDevice(ctx->device));
  bb_comm* c = new (std::nothrow) bb_comm();
  if (!c) return BB_ERR_NOMEM;
  c->ctx = ctx; c->nranks = nranks; c->rank = rank; c->own_comm = true;
  nccl_uid u;
  memcpy(u.internal, id, BB_NCCL_ID_BYTES);
  rc = nccl_check(g_nccl.CommInitRank(&c->comm, nranks, u, rank), "ncclCommInitRank");
  if (rc) { delete c; return rc; }
  return comm_finish(ctx, c, out);
}

extern "C" int bb_comm_adopt(bb_ctx* ctx, void* nccl_comm, int32_t nranks, int32_t rank, bb_comm** out) {
  if (!out) return BB_ERR_ARG;
  *out = nullptr;
  if (!ctx) return BB_ERR_NODEVICE;
  if (!nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) return BB_ERR_ARG;
  int rc = nccl_load();
  if (rc) return rc;
  BB_CUDA(cudaSetDevice(ctx->device));
  bb_comm* c = new (std::nothrow) bb_comm();
  if (!c) return BB_ERR_NOMEM;
  c->ctx = ctx; c->nranks = nranks; c->rank = rank; c->own_comm = false;
  c->comm = (nccl_comm_t)nccl_comm;
  return comm_finish(ctx, c, out);
}

extern "C" int bb_comm_destroy(bb_comm* c) {
  if (!c) return BB_ERR_ARG;
  cudaSetDevice(c->ctx->device);
  if (c->side) cudaStreamSynchronize(c->side);
  if (c->own_comm && c->comm) g_nccl.CommDestroy(c->comm);
  for (int i = 0; i < BB_COMM_SLOTS; i++) {
    if (c->snap[i]) cudaEventDestroy(c->snap[i]);
    if (c->done[i]) cudaEventDestroy(c->done[i]);
  }
  if (c->slots) cudaFree(c->slots);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->counted) c->ctx->live--;
  delete c;
  return BB_OK;
}

/* Snapshot the ensemble's acceptance counter (ordered after everything issued on the context's stream so far) and
 * start its all-reduce on the side stream.  Returns immediately; nothing is added to the compute stream but an 8-byte
 * device-to-device copy and an event. */
static int allreduce_counter(bb_comm* c, const unsigned long long* counter) {
  if (!c || !counter) return BB_ERR_ARG;
  bb_ctx* ctx = c->ctx;
  BB_CUDA(cudaSetDevice(ctx->device));
  const int s = (int)(c->issued % BB_COMM_SLOTS);
  unsigned long long* slot = c->slots + 2 * s;
  /* the slot's previous all-reduce (BB_COMM_SLOTS iterations ago) must have finished before it is overwritten */
  if (c->issued >= BB_COMM_SLOTS) BB_CUDA(cudaStreamWaitEvent(ctx->stream, c->done[s], 0));
  BB_CUDA(cudaMemcpyAsync(slot, counter, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, ctx->stream));
  BB_CUDA(cudaEventRecord(c->snap[s], ctx->stream));
  BB_CUDA(cudaStreamWaitEvent(c->side, c->snap[s], 0));
  if (c->nranks > 1) {
    int rc = nccl_check(g_nccl.AllReduce(slot, slot + 1, 1, NCCL_UINT64, NCCL_SUM, c->comm, c->side), "ncclAllReduce");
    if (rc) return rc;
  } else {
    BB_CUDA(cudaMemcpyAsync(slot + 1, slot, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->side));
  }
  BB_CUDA(cudaEventRecord(c->done[s], c->side));
  c->issued++;
  return BB_OK;
}

extern "C" int bb_allreduce_acc(bb_ens* e, bb_comm* c) {
  if (!e || !c || e->ctx != c->ctx) return BB_ERR_ARG;
  return allreduce_counter(c, e->acc);
}
extern "C" int bb_allreduce_theta_acc(bb_ens* e, bb_comm* c) {
  if (!e || !c || e->ctx != c->ctx) return BB_ERR_ARG;
  return allreduce_counter(c, (const unsigned long long*)bb_theta_acc_device_ptr(e));
}

/* global sum of the most recent bb_allreduce_acc (waits for the side stream only) */
extern "C" int bb_comm_get_acc(bb_comm* c, int64_t* acc_global) {
  if (!c || !acc_global) return BB_ERR_ARG;
  if (c->issued == 0) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(c->ctx->device));
  const int s = (int)((c->issued - 1) % BB_COMM_SLOTS);
  unsigned long long v = 0;
  BB_CUDA(cudaMemcpyAsync(&v, c->slots + 2 * s + 1, sizeof(v), cudaMemcpyDeviceToHost, c->side));
  BB_CUDA(cudaStreamSynchronize(c->side));
  *acc_global = (int64_t)v;
  return BB_OK;
}

extern "C" int bb_comm_synchronize(bb_comm* c) {
  if (!c) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(c->ctx->device));
  BB_CUDA(cudaStreamSynchronize(c->side));
  return BB_OK;
}
extern "C" int bb_comm_rank(bb_comm* c) { return c ? c->rank : -1; }
extern "C" int bb_comm_size(bb_comm* c) { return c ? c->nranks : -1; }
