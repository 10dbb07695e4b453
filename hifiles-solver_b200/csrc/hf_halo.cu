// Halo exchange of partition ("mpi") interfaces over NCCL send/recv on a dedicated communication stream.
// Replaces the reference's pack -> MPI_Isend/Irecv -> MPI_Waitall on host buffers (reference
// src/mpi_inters.cpp:218-336) and the MPI_Allreduce(MIN) of the time step (reference src/solver.cpp:511,543).
// libnccl is bound at run time (dlopen) so the library still loads on a machine without NCCL; any attempt to use a
// multi-rank context without it fails loudly.
#include "hf_device.h"
#include <dlfcn.h>
#include <cstring>
#include <mutex>

namespace
{
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm_t;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_comm_init)(nccl_comm_t *, int, nccl_uid, int);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef int (*fn_send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef const char *(*fn_errstr)(int);

struct nccl_api
{
  void *lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_comm_init comm_init = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_send send = nullptr;
  fn_recv recv = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_errstr errstr = nullptr;
};
nccl_api g_nccl;
const int k_nccl_float64 = 8; // ncclFloat64
const int k_nccl_min = 3;     // ncclMin

int load_nccl_once()
{
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names)
  {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) { hf_set_error(std::string("cannot load libnccl: ") + dlerror()); return 1; }
#define HF_SYM(field, type, name)                                                           \
  g_nccl.field = (type)dlsym(g_nccl.lib, name);                                             \
  if (!g_nccl.field) { hf_set_error(std::string("libnccl lacks ") + name); return 1; }
  HF_SYM(get_uid, fn_get_uid, "ncclGetUniqueId");
  HF_SYM(comm_init, fn_comm_init, "ncclCommInitRank");
  HF_SYM(comm_destroy, fn_comm_destroy, "ncclCommDestroy");
  HF_SYM(send, fn_send, "ncclSend");
  HF_SYM(recv, fn_recv, "ncclRecv");
  HF_SYM(group_start, fn_group, "ncclGroupStart");
  HF_SYM(group_end, fn_group, "ncclGroupEnd");
  HF_SYM(allreduce, fn_allreduce, "ncclAllReduce");
  HF_SYM(errstr, fn_errstr, "ncclGetErrorString");
#undef HF_SYM
  return 0;
}
int load_nccl()
{
  // contexts may be created from several threads: bind the library once
  static std::once_flag once;
  static int status = 1;
  static std::string err;
  std::call_once(once, [] { status = load_nccl_once(); if (status) err = hf_get_error(); });
  if (status) hf_set_error(err);
  return status;
}
#define HF_NCCL(call)                                                                                      \
  do {                                                                                                     \
    int r_ = (call);                                                                                       \
    if (r_ != 0) { hf_set_error(std::string(#call) + ": " + g_nccl.errstr(r_)); return 1; }                \
  } while (0)
} // namespace

extern "C" int hf_dev_nccl_unique_id(void *unique_id_128_bytes)
{
  if (load_nccl()) return 1;
  nccl_uid id;
  HF_NCCL(g_nccl.get_uid(&id));
  memcpy(unique_id_128_bytes, &id, 128);
  return 0;
}

extern "C" int hf_dev_nccl_init(hf_ctx *c, const void *unique_id_128_bytes)
{
  if (load_nccl()) return 1;
  HF_CUDA(cudaSetDevice(c->device));
  nccl_uid id;
  memcpy(&id, unique_id_128_bytes, 128);
  nccl_comm_t comm = nullptr;
  HF_NCCL(g_nccl.comm_init(&comm, c->nproc, id, c->rank));
  c->nccl_comm = comm;
  // the cross-rank agreement of the fused path needs the finalized setup: when the communicator comes first it runs at the
  // end of hf_dev_finalize_setup instead
  return c->finalized ? hf_fused_after_nccl(c) : 0;
}

__global__ void k_bump_flag(unsigned *flag, unsigned value)
{
  __threadfence_system();
  *(volatile unsigned *)flag = value;
}

int hf_halo_post(hf_ctx *c, hf_mpi_inters_dev &I, const double *out, double *in, size_t per_inter, int which)
{
  if (!c->nccl_comm) { hf_set_error("partition interfaces present but hf_dev_nccl_init was not called"); return 1; }
  // the pack kernel ran on the compute stream: order the exchange after it, without blocking the compute stream
  HF_CUDA(cudaEventRecord(c->ev_a, c->stream));
  HF_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_a, 0));
  hf_tl_mark(c, c->tl_xmark, true);
  HF_NCCL(g_nccl.group_start());
  size_t off = 0;
  int first_err = 0; // a failing send / recv must not leave the communicator inside an open group
  for (size_t p = 0; p < I.nb_rank.size() && !first_err; p++)
  {
    size_t cnt = (size_t)I.nb_count[p] * per_inter;
    first_err = g_nccl.send(out + off, cnt, k_nccl_float64, I.nb_rank[p], (nccl_comm_t)c->nccl_comm, c->comm_stream);
    if (!first_err) first_err = g_nccl.recv(in + off, cnt, k_nccl_float64, I.nb_rank[p], (nccl_comm_t)c->nccl_comm, c->comm_stream);
    off += cnt;
  }
  const int end_err = g_nccl.group_end();
  if (first_err || end_err) { hf_set_error(std::string("ncclSend/ncclRecv of the halo exchange: ") + g_nccl.errstr(first_err ? first_err : end_err)); return 1; }
  hf_tl_mark(c, c->tl_xmark + 1, true);
  if (which >= 0)
  {
    // the received blocks are complete once this tiny kernel runs (stream order behind the NCCL kernels): kernels already running on
    // the compute stream see the counter move
    if (!c->d_xflag && hf_alloc_zero(c, &c->d_xflag, 2)) return 1;
    k_bump_flag<<<1, 1, 0, c->comm_stream>>>(c->d_xflag + which, ++c->x_posted[which]);
  }
  HF_CUDA(cudaEventRecord(c->ev_b, c->comm_stream));
  c->halo_pending = true;
  return 0;
}

int hf_halo_wait(hf_ctx *c)
{
  if (!c->halo_pending) return 0;
  HF_CUDA(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
  c->halo_pending = false;
  return 0;
}

int hf_halo_allreduce_min(hf_ctx *c, double *v)
{
  if (!c->nccl_comm) { hf_set_error("multi-rank context but hf_dev_nccl_init was not called"); return 1; }
  HF_CUDA(cudaMemcpyAsync(c->scratch, v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HF_NCCL(g_nccl.allreduce(c->scratch, c->scratch, 1, k_nccl_float64, k_nccl_min, (nccl_comm_t)c->nccl_comm, c->stream));
  HF_CUDA(cudaMemcpyAsync(v, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

/* sum of v[n] over the ranks, result on every rank (the reference reduces its diagnostics to rank 0 with MPI_Reduce(SUM),
 * src/output.cpp:2030-2037) */
static int allreduce_host(hf_ctx *c, double *v, int n, int op);
extern "C" int hf_dev_allreduce_sum(hf_ctx *c, double *v, int n) { return allreduce_host(c, v, n, 0 /* ncclSum */); }
/* max over the ranks (MPI_MAX of the infinity-norm residual, reference src/output.cpp:2216-2221) */
extern "C" int hf_dev_allreduce_max(hf_ctx *c, double *v, int n) { return allreduce_host(c, v, n, 2 /* ncclMax */); }
static int allreduce_host(hf_ctx *c, double *v, int n, int op)
{
  if (c->nproc < 2) return 0;
  if (!c->nccl_comm) { hf_set_error("multi-rank context but hf_dev_nccl_init was not called"); return 1; }
  if ((size_t)n * sizeof(double) > c->scratch_bytes) { hf_set_error("hf_dev_allreduce: too many values"); return 1; }
  HF_CUDA(cudaSetDevice(c->device));
  HF_CUDA(cudaMemcpyAsync(c->scratch, v, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HF_NCCL(g_nccl.allreduce(c->scratch, c->scratch, (size_t)n, k_nccl_float64, op, (nccl_comm_t)c->nccl_comm, c->stream));
  HF_CUDA(cudaMemcpyAsync(v, c->scratch, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

void hf_halo_destroy(hf_ctx *c)
{
  if (c->nccl_comm && g_nccl.comm_destroy) g_nccl.comm_destroy((nccl_comm_t)c->nccl_comm);
  c->nccl_comm = nullptr;
}
