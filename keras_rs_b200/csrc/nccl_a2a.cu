// nccl_a2a.cu — NCCL plumbing for the row-sharded path (separate library: libkrs_b200_nccl.so, so that
// the single-GPU library carries no NCCL dependency).  This is the BASELINE exchange for config C5:
// an all-to-all-v built from grouped ncclSend/ncclRecv for ids / rows / row-gradients, plus the
// all-reduce of the dense (cross + MLP) gradients.  The product path for rows is the peer-memory
// gather in gather.cu; this file exists so that path can be measured against plain NCCL.
#include <nccl.h>
#include <string.h>

#include "common.cuh"

#define KRS_NCCL(expr)                                                                 \
  do {                                                                                 \
    ncclResult_t _r = (expr);                                                          \
    if (_r != ncclSuccess) {                                                           \
      krs_nccl_set_error("NCCL error %d (%s): %s", (int)_r, ncclGetErrorString(_r), #expr); \
      return KRS_ENCCL;                                                                \
    }                                                                                  \
  } while (0)

static thread_local char g_nccl_err[512] = "";
static void krs_nccl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_nccl_err, sizeof(g_nccl_err), fmt, ap);
  va_end(ap);
}

extern "C" {
const char* krs_nccl_last_error(void) { return g_nccl_err; }

int krs_nccl_unique_id(void* id_out_128B) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  KRS_NCCL(ncclGetUniqueId(&id));
  memcpy(id_out_128B, &id, 128);
  return KRS_OK;
}
int krs_nccl_init(void** comm_out, const void* id_128B, int nranks, int rank) {
  ncclUniqueId id;
  memcpy(&id, id_128B, 128);
  ncclComm_t comm;
  KRS_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
  *comm_out = comm;
  return KRS_OK;
}
int krs_nccl_destroy(void* comm) {
  KRS_NCCL(ncclCommDestroy((ncclComm_t)comm));
  return KRS_OK;
}
int krs_nccl_all_to_all_v(void* comm, const void* sendbuf, const int64_t* send_bytes, const int64_t* send_displs,
                          void* recvbuf, const int64_t* recv_bytes, const int64_t* recv_displs, int nranks,
                          void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  KRS_NCCL(ncclGroupStart());
  for (int r = 0; r < nranks; ++r) {
    if (send_bytes[r] > 0)
      KRS_NCCL(ncclSend((const char*)sendbuf + send_displs[r], (size_t)send_bytes[r], ncclChar, r, (ncclComm_t)comm, s));
    if (recv_bytes[r] > 0)
      KRS_NCCL(ncclRecv((char*)recvbuf + recv_displs[r], (size_t)recv_bytes[r], ncclChar, r, (ncclComm_t)comm, s));
  }
  KRS_NCCL(ncclGroupEnd());
  return KRS_OK;
}
int krs_nccl_all_reduce_sum_f32(void* comm, float* buf, int64_t n, void* stream) {
  KRS_NCCL(ncclAllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, (ncclComm_t)comm, (cudaStream_t)stream));
  return KRS_OK;
}
}
