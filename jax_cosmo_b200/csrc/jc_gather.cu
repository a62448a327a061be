// jc_gather.cu -- the path's one multi-GPU exchange: the gather of the per-rank [B/G, P, L] blocks,
// overlapped with the compute of the following cosmologies.
//
// Cosmologies are independent, so the G ranks never exchange anything while K1..K4 run.  What remains is that every
// rank wants the full [B, P, L] result.  Each GPU must RECEIVE (G-1)/G of it (9.6 GB at config 5 on 8 GPUs = 10.7 ms
// of NVLink ingress at 900 GB/s) whatever the algorithm, so the design goal is to hide that time under the 17 ms of
// compute rather than to shave it:
//   * every rank owns a full-size result buffer in device memory; the buffers are mapped into all ranks
//     (cudaIpc handles between processes, plain peer pointers inside one process);
//   * the rank computes its rows in sub-chunks straight into its own buffer (K4 stores the final layout, no pad copy);
//   * as soon as a sub-chunk's K4 is done (event), the copy engines push that slab into the same rows of every
//     peer's buffer over NVLink (cudaMemcpyAsync on side streams: no SM is taken from the FP64 kernels, which NCCL's
//     copy kernels would do; peer order (rank+1, rank+2, ...) so that at any time a GPU receives from one sender);
//   * only the last sub-chunk's push is exposed.
// A K4 epilogue that stored to the peers directly would put all 9.6 GB of egress into K4's 6 ms (> the link rate)
// and stall the DMMA pipe on remote stores, so the push is deliberately a separate, asynchronous copy.
// Completion: on return `stream` is ordered after this rank's OUTGOING pushes; rows written by the peers are
// complete once every rank has reached that point -- the caller closes the step with any stream-ordered cross-rank
// barrier (bench.py / distributed.py: a one-element NCCL all-reduce).
#include <cstdlib>
#include <cstring>

#include "jc_internal.cuh"

#define JC_GATHER_MAX_STREAMS 4

struct jc_gather {
  int rank, world, device;
  size_t bytes;
  void* local;                      // this rank's full-size buffer (cudaMalloc)
  void* peer[JC_MAX_RANKS];         // every rank's buffer as seen from this process (peer[rank] == local)
  bool ipc_opened[JC_MAX_RANKS];
  bool connected;
  int n_streams;
  cudaStream_t copy[JC_GATHER_MAX_STREAMS];
  cudaEvent_t ev_chunk, ev_tail[JC_GATHER_MAX_STREAMS];
};

extern "C" int jc_gather_create(int32_t rank, int32_t world, int32_t device, size_t bytes, jc_gather** out,
                                unsigned char* handle_out) {
  if (!out || world < 1 || world > JC_MAX_RANKS || rank < 0 || rank >= world || bytes == 0) return JC_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == JC_IPC_HANDLE_BYTES, "handle size");
  JcDeviceGuard guard(device);
  JC_CUDA_TRY(guard.status);
  jc_gather* g = new jc_gather();
  memset(g, 0, sizeof(*g));
  g->rank = rank; g->world = world; g->device = device; g->bytes = bytes;
  cudaError_t e = cudaMalloc(&g->local, bytes);
  if (e != cudaSuccess) { delete g; jc_set_cuda_error(e, "cudaMalloc(gather buffer)"); return JC_ERR_CUDA; }
  g->peer[rank] = g->local;
  const char* env = getenv("JC_GATHER_STREAMS");  // tuning knob: copy streams the peers are dealt over
  g->n_streams = env ? atoi(env) : 2;
  if (g->n_streams < 1) g->n_streams = 1;
  if (g->n_streams > JC_GATHER_MAX_STREAMS) g->n_streams = JC_GATHER_MAX_STREAMS;
  for (int i = 0; i < g->n_streams; ++i) {
    e = cudaStreamCreateWithFlags(&g->copy[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_tail[i], cudaEventDisableTiming);
    if (e != cudaSuccess) { jc_set_cuda_error(e, "gather streams"); jc_gather_destroy(g); return JC_ERR_CUDA; }
  }
  e = cudaEventCreateWithFlags(&g->ev_chunk, cudaEventDisableTiming);
  if (e != cudaSuccess) { jc_set_cuda_error(e, "gather event"); jc_gather_destroy(g); return JC_ERR_CUDA; }
  if (handle_out) {
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, g->local);
    if (e != cudaSuccess) { jc_set_cuda_error(e, "cudaIpcGetMemHandle"); jc_gather_destroy(g); return JC_ERR_CUDA; }
    memcpy(handle_out, &h, sizeof(h));
  }
  *out = g;
  return JC_OK;
}

extern "C" void* jc_gather_buffer(const jc_gather* g) { return g ? g->local : nullptr; }

// handles: [world][JC_IPC_HANDLE_BYTES], the handle_out of every rank's jc_gather_create in rank order
extern "C" int jc_gather_connect_ipc(jc_gather* g, const unsigned char* handles) {
  if (!g || !handles) return JC_ERR_INVALID;
  JcDeviceGuard guard(g->device);
  JC_CUDA_TRY(guard.status);
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank || g->peer[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * JC_IPC_HANDLE_BYTES, sizeof(h));
    JC_CUDA_TRY(cudaIpcOpenMemHandle(&g->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    g->ipc_opened[r] = true;
  }
  g->connected = true;
  return JC_OK;
}

// one process driving several devices: ptrs[r] = jc_gather_buffer of rank r's object, devices[r] its device
extern "C" int jc_gather_connect_local(jc_gather* g, void* const* ptrs, const int32_t* devices) {
  if (!g || !ptrs || !devices) return JC_ERR_INVALID;
  JcDeviceGuard guard(g->device);
  JC_CUDA_TRY(guard.status);
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank) continue;
    if (!ptrs[r]) return JC_ERR_INVALID;
    if (devices[r] != g->device) {
      int can = 0;
      JC_CUDA_TRY(cudaDeviceCanAccessPeer(&can, g->device, devices[r]));
      if (!can) return JC_ERR_UNSUPPORTED;
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else JC_CUDA_TRY(e);
    }
    g->peer[r] = ptrs[r];
  }
  g->connected = true;
  return JC_OK;
}

extern "C" int jc_gather_destroy(jc_gather* g) {
  if (!g) return JC_OK;
  JcDeviceGuard guard(g->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < g->world; ++r)
    if (g->ipc_opened[r] && g->peer[r]) cudaIpcCloseMemHandle(g->peer[r]);
  for (int i = 0; i < JC_GATHER_MAX_STREAMS; ++i) {
    if (g->copy[i]) cudaStreamDestroy(g->copy[i]);
    if (g->ev_tail[i]) cudaEventDestroy(g->ev_tail[i]);
  }
  if (g->ev_chunk) cudaEventDestroy(g->ev_chunk);
  if (g->local) cudaFree(g->local);
  delete g;
  return JC_OK;
}

// Push rows [row0, row0 + rows) of the local buffer (row_bytes each) to every peer, ordered after `ev` on the copy streams.
static int push_rows(jc_gather* g, size_t row_bytes, int64_t row0, int64_t rows, cudaEvent_t ev) {
  for (int i = 0; i < g->n_streams; ++i) JC_CUDA_TRY(cudaStreamWaitEvent(g->copy[i], ev, 0));
  const size_t off = (size_t)row0 * row_bytes, n = (size_t)rows * row_bytes;
  for (int i = 1; i < g->world; ++i) {
    const int r = (g->rank + i) % g->world;
    JC_CUDA_TRY(cudaMemcpyAsync((char*)g->peer[r] + off, (const char*)g->local + off, n, cudaMemcpyDeviceToDevice,
                                g->copy[(i - 1) % g->n_streams]));
  }
  return JC_OK;
}

// angular_cl of this rank's n_cosmo cosmologies into rows [row_offset, row_offset + n_cosmo) of the gather buffer
// (layout [rows_total, P, L]) and, slice by slice, into the same rows of every peer's buffer.  K1..K3 run on compute
// chunks of `sub_chunk` cosmologies (full waves); the contraction of a chunk is launched per `push_rows` cosmologies and
// each finished slice is pushed while the following slices / the next chunk compute.
namespace {
struct PushCtx {
  jc_gather* g;
  size_t row_bytes;
  int64_t row_offset;
  cudaStream_t s;
};
int push_cb(void* p, int64_t first_row, int64_t rows) {
  PushCtx* c = (PushCtx*)p;
  if (c->g->world < 2) return JC_OK;
  JC_CUDA_TRY(cudaEventRecord(c->g->ev_chunk, c->s));
  return push_rows(c->g, c->row_bytes, c->row_offset + first_row, rows, c->g->ev_chunk);
}
}  // namespace

extern "C" int jc_angular_cl_gather_f64(const jc_plan* plan, jc_gather* g, const double* cosmo_dev, int64_t n_cosmo,
                                        int64_t row_offset, int64_t sub_chunk, int64_t push_rows_n, void* ws_dev,
                                        size_t ws_bytes, void* stream) {
  if (!plan || !g || plan->d.grid_mode || n_cosmo < 0 || row_offset < 0 || (n_cosmo > 0 && (!cosmo_dev || !ws_dev)))
    return JC_ERR_INVALID;
  if (g->world > 1 && !g->connected) return JC_ERR_INVALID;
  if (plan->device != g->device) return JC_ERR_INVALID;
  const size_t row_bytes = (size_t)plan->d.P * plan->d.L * sizeof(double);
  if ((size_t)(row_offset + n_cosmo) * row_bytes > g->bytes) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);
  JC_CUDA_TRY(guard.status);
  cudaStream_t s = (cudaStream_t)stream;
  if (n_cosmo > 0) {
    PushCtx ctx{g, row_bytes, row_offset, s};
    double* out = (double*)g->local + (size_t)row_offset * plan->d.P * plan->d.L;
    int st = jc_run_pipeline(plan, cosmo_dev, n_cosmo, out, ws_dev, ws_bytes, s, sub_chunk, push_rows_n, push_cb, &ctx);
    if (st != JC_OK) return st;
  }
  if (g->world > 1)
    for (int i = 0; i < g->n_streams; ++i) {  // `stream` continues after this rank's pushes
      JC_CUDA_TRY(cudaEventRecord(g->ev_tail[i], g->copy[i]));
      JC_CUDA_TRY(cudaStreamWaitEvent(s, g->ev_tail[i], 0));
    }
  return JC_OK;
}

// The exchange alone (rows already in the local buffer): used to time the NVLink leg by itself.
extern "C" int jc_gather_push_f64(jc_gather* g, size_t row_bytes, int64_t row_offset, int64_t rows, void* stream) {
  if (!g || rows < 0 || row_offset < 0 || (size_t)(row_offset + rows) * row_bytes > g->bytes) return JC_ERR_INVALID;
  if (g->world > 1 && !g->connected) return JC_ERR_INVALID;
  JcDeviceGuard guard(g->device);
  JC_CUDA_TRY(guard.status);
  cudaStream_t s = (cudaStream_t)stream;
  if (g->world > 1 && rows > 0) {
    JC_CUDA_TRY(cudaEventRecord(g->ev_chunk, s));
    int st = push_rows(g, row_bytes, row_offset, rows, g->ev_chunk);
    if (st != JC_OK) return st;
    for (int i = 0; i < g->n_streams; ++i) {
      JC_CUDA_TRY(cudaEventRecord(g->ev_tail[i], g->copy[i]));
      JC_CUDA_TRY(cudaStreamWaitEvent(s, g->ev_tail[i], 0));
    }
  }
  return JC_OK;
}
