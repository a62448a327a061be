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
// Two transports (jc_gather_create's `push_sms`):
//   push_sms = 0  copy engines: per finished slice one cudaMemcpyAsync per peer on a side stream.  Hidden completely at 2
//                 and 4 GPUs; at 8 GPUs the 98 copies of ~100 MB per step reach only 446 GB/s of the 767 GB/s one 1.4 GB copy
//                 per peer gets, with or without kernels running beside them (scripts/gather_probe.py): ranks drift apart,
//                 several senders hit one receiver while another idles, and the step grows by 6-11 ms.
//   push_sms > 0  a persistent PUSHER KERNEL on `push_sms` SMs of every GPU (default 12): it waits on a device flag that the
//                 compute stream bumps after each slice's contraction, reads the slice once from local HBM (ld.global.cg) and
//                 stores it to all peers' mapped buffers (st.global over NVLink), 16 bytes per thread, the 7 destinations
//                 interleaved in every warp -- every GPU talks to every other GPU at the same rate all the time, the uniform
//                 all-to-all pattern NVSwitch carries at full bandwidth.  The FP64 kernels that fill an SM's register file
//                 (lens, power, contraction) cannot share an SM with it, so the persistent contraction sizes its grid to the
//                 remaining SMs (JcDevPlan::reserved_sms).
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "jc_internal.cuh"
#include "jc_tma.cuh"

#define JC_GATHER_MAX_STREAMS 4
#define JC_PUSH_MAX_SLICES 96
#define JC_PUSH_THREADS 512

int jc_pipeline_slices(const jc_plan* plan, int64_t n_cosmo, size_t ws_bytes, int64_t chunk_cap, int64_t slice,
                       std::vector<std::pair<int64_t, int64_t>>* out);

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
  int push_sms;        // > 0: pusher-kernel transport on that many SMs
  unsigned* flag_dev;  // [0] slices whose contraction has finished, [1] pusher gave up waiting (error)
  bool warmed;         // the pipeline's kernels are loaded on this device (see jc_angular_cl_gather_f64)
  // copy-engine transport in lockstep: per slice a flag barrier over all ranks (stream memory operations, no kernel)
  bool lockstep_ok;    // stream memops available and enabled (JC_GATHER_LOCKSTEP, default 1)
  int lockstep_min_world;
  unsigned epoch;      // slices signalled so far (identical on every rank: equal shards, same schedule)
  unsigned* stage_dev; // [JC_STAGE_WORDS] local words the signal values are copied from
};

#define JC_SYNC_BYTES 4096  // flag words behind the result buffer: flags[src rank] = slices src has finished computing
#define JC_STAGE_WORDS 256

typedef int (*jc_memop32_fn)(cudaStream_t, unsigned long long /*CUdeviceptr*/, unsigned, unsigned);
static jc_memop32_fn g_wait32 = nullptr, g_write32 = nullptr;
static bool load_memops() {
  static int state = 0;  // 0 unknown, 1 ok, -1 unavailable
  if (state == 0) {
    void *w = nullptr, *r = nullptr;
    cudaDriverEntryPointQueryResult q1, q2;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &w, cudaEnableDefault, &q1) == cudaSuccess && w &&
        cudaGetDriverEntryPoint("cuStreamWriteValue32", &r, cudaEnableDefault, &q2) == cudaSuccess && r &&
        q1 == cudaDriverEntryPointSuccess && q2 == cudaDriverEntryPointSuccess) {
      g_wait32 = (jc_memop32_fn)w;
      g_write32 = (jc_memop32_fn)r;
      state = 1;
    } else {
      cudaGetLastError();
      state = -1;
    }
  }
  return state == 1;
}

extern "C" int jc_gather_create(int32_t rank, int32_t world, int32_t device, size_t bytes, int32_t push_sms, jc_gather** out,
                                unsigned char* handle_out) {
  if (!out || world < 1 || world > JC_MAX_RANKS || rank < 0 || rank >= world || bytes == 0 || push_sms < 0 || push_sms > 64)
    return JC_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == JC_IPC_HANDLE_BYTES, "handle size");
  JcDeviceGuard guard(device);
  JC_CUDA_TRY(guard.status);
  jc_gather* g = new jc_gather();
  memset(g, 0, sizeof(*g));
  g->rank = rank; g->world = world; g->device = device; g->bytes = bytes;
  // One allocation = result buffer + flag words, rounded up to 2 MiB: with an odd size (bytes + 4096) the peer copies of the
  // whole buffer ran at 544 instead of 770 GB/s (the mapping then falls back to small pages).
  const size_t alloc = (bytes + JC_SYNC_BYTES + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  cudaError_t e = cudaMalloc(&g->local, alloc);
  if (e != cudaSuccess) { delete g; jc_set_cuda_error(e, "cudaMalloc(gather buffer)"); return JC_ERR_CUDA; }
  e = cudaMemset((char*)g->local + bytes, 0, JC_SYNC_BYTES);
  if (e == cudaSuccess) e = cudaMalloc((void**)&g->stage_dev, JC_STAGE_WORDS * sizeof(unsigned));
  if (e != cudaSuccess) { jc_set_cuda_error(e, "gather sync words"); jc_gather_destroy(g); return JC_ERR_CUDA; }
  {
    const char* ls = getenv("JC_GATHER_LOCKSTEP");
    g->lockstep_ok = (ls ? atoi(ls) != 0 : true) && load_memops();
    g->lockstep_min_world = (ls && atoi(ls) == 2) ? 2 : 3;  // 2: also with a single peer (tests)
  }
  g->peer[rank] = g->local;
  g->push_sms = push_sms;
  e = cudaMalloc((void**)&g->flag_dev, 2 * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(g->flag_dev, 0, 2 * sizeof(unsigned));
  if (e != cudaSuccess) { jc_set_cuda_error(e, "gather flag"); jc_gather_destroy(g); return JC_ERR_CUDA; }
  // copy streams the peers are dealt over (copy-engine transport); 1: at 8 GPUs two concurrent copies per GPU already
  // cost 5 ms per step against one (25.4 vs 20.3 ms), four cost 11 ms
  const char* env = getenv("JC_GATHER_STREAMS");
  g->n_streams = env ? atoi(env) : 1;
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
  if (g->flag_dev) cudaFree(g->flag_dev);
  if (g->stage_dev) cudaFree(g->stage_dev);
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
  int n_done;
  bool lockstep;
};
// Flag barrier over all ranks on copy stream 0, without a kernel: the value travels as a 4-byte copy-engine copy from a
// local word (written by cuStreamWriteValue32) into flags[my rank] of every peer, cuStreamWaitValue32 then holds the
// stream until every peer's word has arrived here.  Equal shards: slice k is the same rows on every rank.
int lockstep_barrier(jc_gather* g, cudaStream_t st) {
  const unsigned v = ++g->epoch;
  unsigned* stage = g->stage_dev + (v % JC_STAGE_WORDS);
  if (g_write32(st, (unsigned long long)(uintptr_t)stage, v, 0) != 0) return JC_ERR_CUDA;
  for (int i = 1; i < g->world; ++i) {
    const int r = (g->rank + i) % g->world;
    JC_CUDA_TRY(cudaMemcpyAsync((char*)g->peer[r] + g->bytes + 4 * g->rank, stage, 4, cudaMemcpyDeviceToDevice, st));
  }
  unsigned* mine = (unsigned*)((char*)g->local + g->bytes);
  for (int i = 1; i < g->world; ++i) {
    const int r = (g->rank + i) % g->world;
    if (g_wait32(st, (unsigned long long)(uintptr_t)(mine + r), v, 1 /* CU_STREAM_WAIT_VALUE_GEQ */) != 0) return JC_ERR_CUDA;
  }
  return JC_OK;
}
int push_cb(void* p, int64_t first_row, int64_t rows) {
  PushCtx* c = (PushCtx*)p;
  if (c->g->world < 2) return JC_OK;
  JC_CUDA_TRY(cudaEventRecord(c->g->ev_chunk, c->s));
  if (c->lockstep) {
    // every rank has finished this slice before anyone pushes it: the 7 copies of the slice then run as the permutations
    // r -> r+1, r -> r+2, ... on all ranks at once (one sender per receiver at any time) instead of drifting into each other
    JC_CUDA_TRY(cudaStreamWaitEvent(c->g->copy[0], c->g->ev_chunk, 0));
    int st = lockstep_barrier(c->g, c->g->copy[0]);
    if (st != JC_OK) return st;
    if (c->g->n_streams > 1) {
      JC_CUDA_TRY(cudaEventRecord(c->g->ev_tail[0], c->g->copy[0]));
      return push_rows(c->g, c->row_bytes, c->row_offset + first_row, rows, c->g->ev_tail[0]);
    }
  }
  return push_rows(c->g, c->row_bytes, c->row_offset + first_row, rows, c->g->ev_chunk);
}

// ---- pusher-kernel transport ------------------------------------------------------------------------------------------
struct PushArgs {
  const double* local;
  double* peer[JC_MAX_RANKS - 1];
  int n_peers, n_slices;
  long long first[JC_PUSH_MAX_SLICES];  // first double of the slice (same offset in every buffer)
  long long count[JC_PUSH_MAX_SLICES];  // doubles (a multiple of 2)
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void jc_flag_kernel(unsigned* flag, unsigned value) {
  __threadfence();
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// One CTA per reserved SM (the dynamic shared-memory request keeps a second one away).  Slice by slice: wait for the
// flag, then copy [first, first + count) to every peer.  A thread moves 4 x 16 bytes per iteration: 4 independent
// L2 loads (ld.global.cg: the pusher read the same addresses one step earlier, its L1 must not answer), then the 4 x n_peers
// remote stores, destinations innermost so that a warp's traffic is spread over all NVLink peers at any moment.
__global__ void __launch_bounds__(JC_PUSH_THREADS, 1) jc_push_kernel(PushArgs a, unsigned* flag) {
  __shared__ int s_abort;
  const long long nthreads = (long long)gridDim.x * JC_PUSH_THREADS;
  const long long tid = (long long)blockIdx.x * JC_PUSH_THREADS + threadIdx.x;
  for (int j = 0; j < a.n_slices; ++j) {
    if (threadIdx.x == 0) {
      int abort_ = 0;
      const long long t0 = clock64();
      while (ld_acquire(flag) < (unsigned)(j + 1)) {
        __nanosleep(256);
        if (clock64() - t0 > 6000000000ll) { abort_ = 1; flag[1] = 1u; break; }  // ~3 s: the compute stream died; do not hang the GPU
      }
      s_abort = abort_;
    }
    __syncthreads();
    if (s_abort) return;
    const double2* src = reinterpret_cast<const double2*>(a.local + a.first[j]);
    const long long n2 = a.count[j] >> 1;
    for (long long i = tid; i < n2; i += 4 * nthreads) {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long k = i + u * nthreads;
        if (k < n2) v[u] = __ldcg(src + k);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long k = i + u * nthreads;
        if (k < n2) {
          for (int p = 0; p < a.n_peers; ++p) reinterpret_cast<double2*>(a.peer[p] + a.first[j])[k] = v[u];
        }
      }
    }
    __syncthreads();  // s_abort is rewritten for the next slice
  }
}

// TMA form of the pusher: one elected thread per CTA drives a ring of NBUF x 32 KB shared-memory buffers -- bulk loads
// (cp.async.bulk global -> shared, completion on an mbarrier) run D = NBUF - 1 chunks ahead, every landed chunk leaves as
// n_peers bulk stores (cp.async.bulk shared -> the peers' mapped global memory, one bulk group per chunk); a buffer is
// reloaded once cp.async.bulk.wait_group.read says its stores have read it.  The LSU version above tops out at ~31 GB/s
// per SM (stores in flight per SM), the TMA engine keeps hundreds of KB in flight from one thread.
constexpr int PUSH_CHUNK = 32768;  // bytes
constexpr int PUSH_NBUF = 6;

__device__ __forceinline__ void tma_store_g(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(__cvta_generic_to_global(gdst)),
               "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(32, 1) jc_push_tma_kernel(PushArgs a, unsigned* flag) {
  extern __shared__ __align__(128) unsigned char push_smem[];
  __shared__ uint64_t full[PUSH_NBUF];
  if (threadIdx.x != 0) return;
  for (int b = 0; b < PUSH_NBUF; ++b) mbar_init(full + b, 1);
  mbar_fence_init();
  long long issued = 0, done = 0;  // chunks loaded / stored by this CTA since the start (ring position, mbarrier parity)
  for (int j = 0; j < a.n_slices; ++j) {
    const long long t0 = clock64();
    while (ld_acquire(flag) < (unsigned)(j + 1)) {
      __nanosleep(256);
      if (clock64() - t0 > 6000000000ll) { flag[1] = 1u; return; }  // ~3 s: the compute stream died; do not hang the GPU
    }
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.local + a.first[j]);
    const long long bytes = a.count[j] * 8;
    const long long n_all = (bytes + PUSH_CHUNK - 1) / PUSH_CHUNK;
    // this CTA's chunks: blockIdx.x, blockIdx.x + gridDim.x, ...
    const long long n_mine = n_all > blockIdx.x ? (n_all - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto chunk_off = [&](long long i) { return (blockIdx.x + i * gridDim.x) * (long long)PUSH_CHUNK; };
    auto chunk_len = [&](long long i) { const long long o = chunk_off(i); return (unsigned)((bytes - o) < PUSH_CHUNK ? (bytes - o) : PUSH_CHUNK); };
    auto issue_load = [&](long long i) {
      const int b = (int)(issued % PUSH_NBUF);
      mbar_expect_tx(full + b, chunk_len(i));
      tma_load(push_smem + (size_t)b * PUSH_CHUNK, src + chunk_off(i), chunk_len(i), full + b);
      ++issued;
    };
    const long long ahead = n_mine < PUSH_NBUF - 1 ? n_mine : PUSH_NBUF - 1;
    // the buffers the prologue reloads were last read by stores of the previous slice: all but none may still be reading
    asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
    for (long long i = 0; i < ahead; ++i) issue_load(i);
    for (long long i = 0; i < n_mine; ++i) {
      const int b = (int)(done % PUSH_NBUF);
      mbar_wait(full + b, (unsigned)((done / PUSH_NBUF) & 1));
      const long long o = (long long)a.first[j] * 8 + chunk_off(i);
      const unsigned len = chunk_len(i);
      for (int p = 0; p < a.n_peers; ++p)
        tma_store_g(reinterpret_cast<unsigned char*>(a.peer[p]) + o, push_smem + (size_t)b * PUSH_CHUNK, len);
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      ++done;
      if (i + ahead < n_mine) {
        // chunk i + ahead reuses the buffer of chunk i - 1: every group but the newest has finished reading shared memory
        asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
        issue_load(i + ahead);
      }
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");  // all stores complete before the kernel ends
}

int flag_cb(void* p, int64_t, int64_t) {
  PushCtx* c = (PushCtx*)p;
  ++c->n_done;
  jc_flag_kernel<<<1, 1, 0, c->s>>>(c->g->flag_dev, (unsigned)c->n_done);
  return JC_OK;
}
}  // namespace

static int launch_pusher(jc_gather* g, const PushArgs& a, cudaStream_t st) {
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("JC_PUSH_TMA"); use_tma = e ? atoi(e) : 1; }  // 0: the LSU (ld/st) pusher
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, {
    cudaFuncSetAttribute(jc_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    cudaFuncSetAttribute(jc_push_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PUSH_NBUF * PUSH_CHUNK);
  });
  if (use_tma) jc_push_tma_kernel<<<g->push_sms, 32, PUSH_NBUF * PUSH_CHUNK, st>>>(a, g->flag_dev);
  else jc_push_kernel<<<g->push_sms, JC_PUSH_THREADS, 120 * 1024, st>>>(a, g->flag_dev);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

extern "C" int jc_angular_cl_gather_f64(const jc_plan* plan, jc_gather* g, const double* cosmo_dev, int64_t n_cosmo,
                                        int64_t row_offset, int64_t sub_chunk, int64_t push_rows_n, int32_t equal_shards,
                                        void* ws_dev, size_t ws_bytes, void* stream) {
  if (!plan || !g || plan->d.grid_mode || n_cosmo < 0 || row_offset < 0 || (n_cosmo > 0 && (!cosmo_dev || !ws_dev)))
    return JC_ERR_INVALID;
  if (g->world > 1 && !g->connected) return JC_ERR_INVALID;
  if (plan->device != g->device) return JC_ERR_INVALID;
  const size_t row_doubles = (size_t)plan->d.P * plan->d.L;
  const size_t row_bytes = row_doubles * sizeof(double);
  if ((size_t)(row_offset + n_cosmo) * row_bytes > g->bytes) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);
  JC_CUDA_TRY(guard.status);
  cudaStream_t s = (cudaStream_t)stream;
  double* out = (double*)g->local + (size_t)row_offset * row_doubles;
  bool sm_push = g->push_sms > 0 && g->world > 1 && n_cosmo > 0 && (row_doubles & 1) == 0;
  std::vector<std::pair<int64_t, int64_t>> slices;
  if (sm_push) {
    int st = jc_pipeline_slices(plan, n_cosmo, ws_bytes, sub_chunk, push_rows_n, &slices);
    if (st != JC_OK) return st;
    if (slices.size() > JC_PUSH_MAX_SLICES) sm_push = false;  // more slices than the kernel's table: copy engines
  }
  if (sm_push) {
    if (!g->warmed) {
      // CUDA loads kernels lazily, and a module load can wait for running kernels to finish: a pusher spinning on its flag
      // while the compute stream's first launch is still loading would dead-lock until the pusher's time-out.  Load every
      // kernel of the step before the first pusher starts.
      int st = jc_run_pipeline(plan, cosmo_dev, 1, out, ws_dev, ws_bytes, s, sub_chunk, push_rows_n, nullptr, nullptr, g->push_sms);
      if (st != JC_OK) return st;
      jc_flag_kernel<<<1, 1, 0, s>>>(g->flag_dev, 0u);
      JC_CUDA_TRY(cudaStreamSynchronize(s));
      g->warmed = true;
    }
    PushArgs a;
    memset(&a, 0, sizeof(a));
    a.local = (const double*)g->local;
    for (int i = 1; i < g->world; ++i) a.peer[a.n_peers++] = (double*)g->peer[(g->rank + i) % g->world];
    a.n_slices = (int)slices.size();
    for (size_t k = 0; k < slices.size(); ++k) {
      a.first[k] = (long long)((row_offset + slices[k].first) * (int64_t)row_doubles);
      a.count[k] = (long long)(slices[k].second * (int64_t)row_doubles);
    }
    // the pusher starts after everything queued on `stream` so far (previous step included) and runs beside this step
    JC_CUDA_TRY(cudaMemsetAsync(g->flag_dev, 0, 2 * sizeof(unsigned), s));
    JC_CUDA_TRY(cudaEventRecord(g->ev_chunk, s));
    JC_CUDA_TRY(cudaStreamWaitEvent(g->copy[0], g->ev_chunk, 0));
    int stp = launch_pusher(g, a, g->copy[0]);
    if (stp != JC_OK) return stp;
    PushCtx ctx{g, row_bytes, row_offset, s, 0, false};
    int st = jc_run_pipeline(plan, cosmo_dev, n_cosmo, out, ws_dev, ws_bytes, s, sub_chunk, push_rows_n, flag_cb, &ctx, g->push_sms);
    if (st != JC_OK) return st;
    JC_CUDA_TRY(cudaEventRecord(g->ev_tail[0], g->copy[0]));
    JC_CUDA_TRY(cudaStreamWaitEvent(s, g->ev_tail[0], 0));
    return JC_OK;
  }
  if (n_cosmo > 0) {
    PushCtx ctx{g, row_bytes, row_offset, s, 0, equal_shards != 0 && g->lockstep_ok && g->world >= g->lockstep_min_world};
    int st = jc_run_pipeline(plan, cosmo_dev, n_cosmo, out, ws_dev, ws_bytes, s, sub_chunk, push_rows_n, push_cb, &ctx, 0);
    if (st != JC_OK) return st;
  }
  if (g->world > 1)
    for (int i = 0; i < g->n_streams; ++i) {  // `stream` continues after this rank's pushes
      JC_CUDA_TRY(cudaEventRecord(g->ev_tail[i], g->copy[i]));
      JC_CUDA_TRY(cudaStreamWaitEvent(s, g->ev_tail[i], 0));
    }
  return JC_OK;
}

// 1 if the pusher kernel of the last step gave up waiting for the compute stream (it then left rows unsent)
extern "C" int jc_gather_status(jc_gather* g, int32_t* aborted_out) {
  if (!g || !aborted_out) return JC_ERR_INVALID;
  JcDeviceGuard guard(g->device);
  unsigned f[2] = {0, 0};
  JC_CUDA_TRY(cudaMemcpy(f, g->flag_dev, sizeof(f), cudaMemcpyDeviceToHost));
  *aborted_out = (int32_t)f[1];
  return JC_OK;
}

// The exchange alone (rows already in the local buffer): used to time the NVLink leg by itself.
extern "C" int jc_gather_push_f64(jc_gather* g, size_t row_bytes, int64_t row_offset, int64_t rows, void* stream) {
  if (!g || rows < 0 || row_offset < 0 || (size_t)(row_offset + rows) * row_bytes > g->bytes) return JC_ERR_INVALID;
  if (g->world > 1 && !g->connected) return JC_ERR_INVALID;
  JcDeviceGuard guard(g->device);
  JC_CUDA_TRY(guard.status);
  cudaStream_t s = (cudaStream_t)stream;
  if (g->world > 1 && rows > 0 && g->push_sms > 0 && (row_bytes & 15) == 0) {  // the pusher kernel alone, one slice, flag preset
    PushArgs a;
    memset(&a, 0, sizeof(a));
    a.local = (const double*)g->local;
    for (int i = 1; i < g->world; ++i) a.peer[a.n_peers++] = (double*)g->peer[(g->rank + i) % g->world];
    a.n_slices = 1;
    a.first[0] = (long long)((size_t)row_offset * (row_bytes / 8));
    a.count[0] = (long long)((size_t)rows * (row_bytes / 8));
    jc_flag_kernel<<<1, 1, 0, s>>>(g->flag_dev, 1u);
    return launch_pusher(g, a, s);
  }
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
