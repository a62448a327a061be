// jc_pipeline.cu -- workspace layout and the stream-ordered launcher of the K1..K4 pipeline
// (kernels: jc_setup.cu, jc_tracers.cu, jc_power.cu, jc_contract.cu; overview in jc_internal.cuh).
#include <utility>
#include <vector>

#include "jc_internal.cuh"

namespace {

size_t per_cosmo_doubles(const JcDevPlan& pl) {
  return (size_t)JC_NCHI + JC_NGROW + JC_SCAL_FIELDS + JC_NHFR + (size_t)JC_NODE_FIELDS * JC_NA_PAD +
         (size_t)JC_NA_PAD * pl.TS + (size_t)JC_NA * pl.Lpad + (size_t)pl.Lpad;
}

void layout_for(const JcDevPlan& pl, int64_t chunk, jc_ws_layout* lo) {
  int64_t o = 0;
  lo->chunk = chunk;
  lo->node_stride = JC_NA_PAD;
  lo->ell_stride = pl.Lpad;
  lo->tracer_stride = pl.TS;
  lo->chitab = o; o += chunk * JC_NCHI;
  lo->gtab = o; o += chunk * JC_NGROW;
  lo->scal = o; o += chunk * JC_SCAL_FIELDS;
  lo->stab = o; o += chunk * JC_NHFR;
  lo->node = o; o += chunk * (int64_t)JC_NODE_FIELDS * JC_NA_PAD;
  lo->rker = o; o += chunk * (int64_t)JC_NA_PAD * pl.TS;
  lo->vtab = o; o += chunk * (int64_t)JC_NA * pl.Lpad;
  lo->ellpow = o; o += chunk * (int64_t)pl.Lpad;
  lo->total = o;
}

}  // namespace

int jc_pipeline_init() {
  int st = jc_contract_init();
  if (st != JC_OK) return st;
  return jc_setup_init();
}

static void resolve(const jc_ws_layout& lo, double* base, ptrdiff_t doff, Ws* ws) {
  ws->chitab = base + lo.chitab; ws->gtab = base + lo.gtab; ws->scal = base + lo.scal;
  ws->stab = base + lo.stab; ws->node = base + lo.node; ws->rker = base + lo.rker; ws->vtab = base + lo.vtab;
  ws->ellpow = base + lo.ellpow;
  ws->doff = doff;
}

extern "C" int jc_workspace_bytes_jvp(const jc_plan* plan, int64_t n_cosmo, size_t* bytes_out) {
  int st = jc_workspace_bytes(plan, n_cosmo, bytes_out);
  if (st == JC_OK) *bytes_out *= 2;  // value plane + tangent plane
  return st;
}

// Workspace of the throughput path for min(n_cosmo, JC_MAX_CHUNK) entries: a value plane and the tangent planes of the mode
// jc_angular_cl_jvp_f64 will pick (reverse-sweep K3: jc_jvp_planes(K) planes, all K directions resident; tangent groups: 1 + g).
static bool use_adjoint(const JcDevPlan& pl, int n_tangents) {
  return g_jc_jvp_adjoint && g_jc_jvp_group == JC_JVP_MAX_GROUP && jc_power_adj_supported(pl, n_tangents);
}
extern "C" int jc_workspace_bytes_jvp_group(const jc_plan* plan, int64_t n_cosmo, int32_t n_tangents, size_t* bytes_out) {
  if (!plan || n_tangents < 1) return JC_ERR_INVALID;
  int st = jc_workspace_bytes(plan, n_cosmo < 1024 ? n_cosmo : 1024, bytes_out);  // 1024 entries per pass fill the GPU 7 times
  const int g = n_tangents < g_jc_jvp_group ? n_tangents : g_jc_jvp_group;
  if (st == JC_OK) *bytes_out = (*bytes_out + 16) * (size_t)(use_adjoint(plan->d, n_tangents) ? jc_jvp_planes(n_tangents) : 1 + g);
  return st;
}

// Forward-mode derivatives: passes of the DualN-instantiated kernels K1..K3 and one tangent contraction per direction.
extern "C" int jc_angular_cl_jvp_f64(const jc_plan* plan, const double* cosmo_dev, const double* tangents_dev,
                                     int32_t n_tangents, int64_t n_cosmo, double* cl_dev, double* dcl_dev,
                                     void* ws_dev, size_t ws_bytes, void* stream) {
  if (!plan || plan->d.grid_mode || !cosmo_dev || !tangents_dev || !dcl_dev || !ws_dev || n_cosmo < 1 || n_tangents < 1)
    return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);  // a null stream handle means the CURRENT device's default stream
  jc_ws_layout lo;
  int st = jc_workspace_layout(plan, ws_bytes / 2, &lo);
  if (st != JC_OK) return st;
  const JcDevPlan& pl = plan->d;
  cudaStream_t s = (cudaStream_t)stream;
  Ws ws;
  resolve(lo, (double*)ws_dev, (ptrdiff_t)((lo.total + 1) & ~(int64_t)1), &ws);
  const int64_t PL = (int64_t)pl.P * pl.L;
  if (n_tangents > 1 && n_cosmo * n_tangents <= lo.chunk && n_cosmo * n_tangents <= JC_JVP_FUSED_MAX) {
    // Small batches (a Fisher forecast at one cosmology): all K directions of every cosmology in ONE pass over B*K workspace
    // entries (entry b*K + k = cosmology b, direction k, which is also the layout of dcl) instead of K latency-bound passes.
    // Needs a workspace for B*K entries (jc_workspace_bytes_jvp(plan, B*K)); per-entry arithmetic is the same, results are
    // bitwise those of the pass-per-direction path.  The value plane of entry b*K is C_l of cosmology b: it is contracted into the
    // dcl buffer first (same size), its rows b*K copied out to cl, then the tangent contraction overwrites dcl.
    const int entries = (int)(n_cosmo * n_tangents);
    jc_launch_setup_jvp(pl, cosmo_dev, tangents_dev, ws, entries, n_tangents, 1, s);
    jc_launch_tracers_jvp(pl, ws, entries, 1, s);
    jc_launch_finish_jvp(pl, ws, entries, 1, s);
    jc_launch_power_jvp(pl, ws, entries, 1, s);
    if (cl_dev) {
      jc_launch_contract(pl, ws, dcl_dev, entries, s);
      JC_CUDA_TRY(cudaMemcpy2DAsync(cl_dev, (size_t)PL * sizeof(double), dcl_dev, (size_t)n_tangents * PL * sizeof(double),
                                    (size_t)PL * sizeof(double), (size_t)n_cosmo, cudaMemcpyDeviceToDevice, s));
    }
    jc_launch_contract_jvp(pl, ws, dcl_dev, PL, entries, s);
    JC_CUDA_TRY(cudaGetLastError());
    return JC_OK;
  }
  // Throughput path: every pass carries a GROUP of g <= jvp_group directions through one evaluation of the value (DualN<g>:
  // the transcendentals of K1..K3 are computed once per group; 7 parameters = groups of 4 + 3).  The workspace is cut into
  // 1 + g planes of `chunk` entries; the group shrinks when the planes would otherwise hold fewer entries than the batch
  // needs to fill the GPU.  The tangent contraction runs per direction on (value plane, tangent plane k).
  const int64_t want = n_cosmo < 148 ? n_cosmo : 148;  // entries per pass that keep every SM busy
  // Reverse-sweep K3 (>= 3 directions, Eisenstein-Hu with wiggles): K1 / K2 run per group of 4 directions on the workspace shifted
  // by 5 planes per group, so that the tangents of EVERY direction are resident; K3 then produces V and all K directional
  // derivatives from one reverse sweep of the point function (jc_power_adj.cu) -- ~2 x the value + one multiply-add per input and
  // direction instead of ~225 FP64 instructions per direction.
  if (use_adjoint(pl, n_tangents)) {
    const int planes = jc_jvp_planes(n_tangents);
    jc_ws_layout la;
    if (jc_workspace_layout(plan, ws_bytes / (size_t)planes, &la) == JC_OK && la.chunk >= want) {
      const ptrdiff_t doff = (ptrdiff_t)((la.total + 1) & ~(int64_t)1);
      resolve(la, (double*)ws_dev, doff, &ws);
      for (int64_t c0 = 0; c0 < n_cosmo; c0 += la.chunk) {
        const int chunk = (int)((n_cosmo - c0) < la.chunk ? (n_cosmo - c0) : la.chunk);
        for (int k0 = 0; k0 < n_tangents; k0 += JC_JVP_MAX_GROUP) {
          const int g = (n_tangents - k0) < JC_JVP_MAX_GROUP ? (n_tangents - k0) : JC_JVP_MAX_GROUP;
          Ws wg;  // this group's value plane = plane jc_jvp_plane(k0) - 1 of the workspace
          resolve(la, (double*)ws_dev + (ptrdiff_t)(jc_jvp_plane(k0) - 1) * doff, doff, &wg);
          jc_launch_setup_jvp(pl, cosmo_dev + c0 * pl.ncp, tangents_dev + (size_t)k0 * pl.ncp, wg, chunk, 1, g, s);
          // only group 0's value tables (R, chi) are read downstream: a later group whose directions all leave the tracer
          // kernels alone (h, n_s, sigma8) has no K2 work at all -- callers that order such directions last save it
          jc_launch_tracers_jvp(pl, wg, chunk, g, s, k0 > 0);
          jc_launch_finish_jvp(pl, wg, chunk, g, s, k0 > 0);
        }
        jc_launch_power_adj(pl, ws, chunk, n_tangents, s);
        if (cl_dev) jc_launch_contract(pl, ws, cl_dev + (size_t)c0 * PL, chunk, s);  // value plane
        for (int k = 0; k < n_tangents; ++k) {
          Ws wk = ws;
          wk.doff = (ptrdiff_t)jc_jvp_plane(k) * doff;  // the contraction pairs the value plane with direction k's plane
          jc_launch_contract_jvp(pl, wk, dcl_dev + ((size_t)c0 * n_tangents + k) * PL, (int64_t)n_tangents * PL, chunk, s);
        }
      }
      JC_CUDA_TRY(cudaGetLastError());
      return JC_OK;
    }
  }
  int group = n_tangents < g_jc_jvp_group ? n_tangents : g_jc_jvp_group;
  while (group > 1) {
    jc_ws_layout lg;
    if (jc_workspace_layout(plan, ws_bytes / (size_t)(1 + group), &lg) == JC_OK && lg.chunk >= want) break;
    --group;
  }
  if ((st = jc_workspace_layout(plan, ws_bytes / (size_t)(1 + group), &lo)) != JC_OK) return st;
  resolve(lo, (double*)ws_dev, (ptrdiff_t)((lo.total + 1) & ~(int64_t)1), &ws);
  for (int64_t c0 = 0; c0 < n_cosmo; c0 += lo.chunk) {
    const int chunk = (int)((n_cosmo - c0) < lo.chunk ? (n_cosmo - c0) : lo.chunk);
    for (int k0 = 0; k0 < n_tangents; k0 += group) {
      const int g = (n_tangents - k0) < group ? (n_tangents - k0) : group;
      jc_launch_setup_jvp(pl, cosmo_dev + c0 * pl.ncp, tangents_dev + (size_t)k0 * pl.ncp, ws, chunk, 1, g, s);
      jc_launch_tracers_jvp(pl, ws, chunk, g, s);
      jc_launch_finish_jvp(pl, ws, chunk, g, s);
      jc_launch_power_jvp(pl, ws, chunk, g, s);
      if (k0 == 0 && cl_dev) jc_launch_contract(pl, ws, cl_dev + (size_t)c0 * PL, chunk, s);  // value plane
      for (int j = 0; j < g; ++j) {
        Ws wk = ws;
        wk.doff = (ptrdiff_t)(j + 1) * ws.doff;  // the contraction pairs the value plane with tangent plane j
        jc_launch_contract_jvp(pl, wk, dcl_dev + ((size_t)c0 * n_tangents + k0 + j) * PL, (int64_t)n_tangents * PL, chunk, s);
      }
    }
  }
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

extern "C" int jc_workspace_bytes(const jc_plan* plan, int64_t n_cosmo, size_t* bytes_out) {
  if (!plan || !bytes_out || n_cosmo < 1) return JC_ERR_INVALID;
  int64_t chunk = n_cosmo < JC_MAX_CHUNK ? n_cosmo : JC_MAX_CHUNK;
  *bytes_out = per_cosmo_doubles(plan->d) * (size_t)chunk * sizeof(double);
  return JC_OK;
}

extern "C" int jc_workspace_layout(const jc_plan* plan, size_t ws_bytes, jc_ws_layout* lo) {
  if (!plan || !lo) return JC_ERR_INVALID;
  size_t per = per_cosmo_doubles(plan->d) * sizeof(double);
  int64_t chunk = (int64_t)(ws_bytes / per);
  if (chunk < 1) return JC_ERR_WORKSPACE;
  if (chunk > JC_MAX_CHUNK) chunk = JC_MAX_CHUNK;
  layout_for(plan->d, chunk, lo);
  return JC_OK;
}

// K1..K4 over n_cosmo cosmologies in chunks of <= chunk_cap (and <= what the workspace holds).  With slice > 0 the
// contraction of a chunk is launched per slice of cosmologies and `cb(ctx, first_row, rows)` is called after each slice's
// launch (the gather records an event there and starts the peer copies of those rows): the exchange granularity is
// then independent of the compute chunk, whose K1..K3 keep full waves.  The last slice of the batch is halved so that
// the copy left exposed at the end of the step is short.
// With slicing (the gather) the first compute chunk is one slice long: the first push can start after half the
// K1..K3 time of a full chunk (the exchange needs 89 % of the step at 8 GPUs, so its start-up delay is exposed 1:1).
static inline int64_t jc_first_chunk(int64_t cap, int64_t slice, int64_t c0) {
  return (slice > 0 && c0 == 0 && slice < cap) ? slice : cap;
}

// The (first row, rows) slices jc_run_pipeline will launch the contraction for, in order (same rule as below).
int jc_pipeline_slices(const jc_plan* plan, int64_t n_cosmo, size_t ws_bytes, int64_t chunk_cap, int64_t slice,
                       std::vector<std::pair<int64_t, int64_t>>* out) {
  jc_ws_layout lo;
  int st = jc_workspace_layout(plan, ws_bytes, &lo);
  if (st != JC_OK) return st;
  const int64_t cap = (chunk_cap > 0 && chunk_cap < lo.chunk) ? chunk_cap : lo.chunk;
  int64_t next = 0;
  for (int64_t c0 = 0; c0 < n_cosmo; c0 = next) {
    const int64_t this_cap = jc_first_chunk(cap, slice, c0);
    const int64_t chunk = (n_cosmo - c0) < this_cap ? (n_cosmo - c0) : this_cap;
    next = c0 + chunk;
    if (slice <= 0) { out->push_back({c0, chunk}); continue; }
    const bool last_chunk = c0 + chunk >= n_cosmo;
    for (int64_t s0 = 0; s0 < chunk;) {
      int64_t ns = (chunk - s0) < slice ? (chunk - s0) : slice;
      if (last_chunk && s0 + ns >= chunk && ns > slice / 2 && ns >= 2) ns = (ns + 1) / 2;
      out->push_back({c0 + s0, ns});
      s0 += ns;
    }
  }
  return JC_OK;
}

int jc_run_pipeline(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, double* cl_dev, void* ws_dev,
                    size_t ws_bytes, cudaStream_t s, int64_t chunk_cap, int64_t slice, jc_slice_cb cb, void* ctx,
                    int reserved_sms) {
  if (!plan || plan->d.grid_mode || !cosmo_dev || !cl_dev || !ws_dev || n_cosmo < 1) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);  // a null stream handle means the CURRENT device's default stream
  jc_ws_layout lo;
  int st = jc_workspace_layout(plan, ws_bytes, &lo);
  if (st != JC_OK) return st;
  JcDevPlan pl = plan->d;
  pl.reserved_sms = reserved_sms;  // SMs held by a concurrently running kernel (the gather's pusher): persistent grids leave them out
  double* base = (double*)ws_dev;
  Ws ws;
  resolve(lo, base, 0, &ws);
  const int64_t cap = (chunk_cap > 0 && chunk_cap < lo.chunk) ? chunk_cap : lo.chunk;
  const size_t PL = (size_t)pl.P * pl.L;

  JcProf* prof = (plan->prof && plan->prof->enabled) ? plan->prof : nullptr;
  int64_t c_next = 0;
  for (int64_t c0 = 0; c0 < n_cosmo; c0 = c_next) {
    const int64_t this_cap = jc_first_chunk(cap, slice, c0);
    const int chunk = (int)((n_cosmo - c0) < this_cap ? (n_cosmo - c0) : this_cap);
    c_next = c0 + chunk;
    cudaEvent_t* ev = nullptr;
    int* nl = nullptr;
    if (prof && prof->used < JC_PROF_SLOTS) { ev = prof->ev[prof->used]; nl = prof->launches[prof->used]; ++prof->used; }
#define JC_MARK(i) do { if (ev) cudaEventRecord(ev[i], s); } while (0)
    JC_MARK(0);
    jc_launch_setup(pl, cosmo_dev + c0 * pl.ncp, ws, chunk, s);
    JC_MARK(1);
    const int n_lens = jc_launch_tracers(pl, ws, chunk, s);
    JC_MARK(2);
    jc_launch_finish(pl, ws, chunk, s);
    JC_MARK(3);
    jc_launch_power(pl, ws, chunk, s);
    JC_MARK(4);
    int n_contract = 0;
    if (slice <= 0) {
      jc_launch_contract(pl, ws, cl_dev + (size_t)c0 * PL, chunk, s);
      n_contract = 1;
      if (cb && (st = cb(ctx, c0, chunk)) != JC_OK) return st;
    } else {
      const bool last_chunk = c0 + chunk >= n_cosmo;
      for (int s0 = 0; s0 < chunk;) {
        int ns = (int)((chunk - s0) < slice ? (chunk - s0) : slice);
        if (last_chunk && s0 + ns >= chunk && ns > slice / 2 && ns >= 2) ns = (ns + 1) / 2;  // halve the batch's final slice
        Ws w = ws;  // the contraction reads R and V of the slice's cosmologies only
        w.rker += (size_t)s0 * JC_NA_PAD * pl.TS;
        w.vtab += (size_t)s0 * JC_NA * pl.Lpad;
        jc_launch_contract(pl, w, cl_dev + (size_t)(c0 + s0) * PL, ns, s);
        ++n_contract;
        if (cb && (st = cb(ctx, c0 + s0, ns)) != JC_OK) return st;
        s0 += ns;
      }
    }
    JC_MARK(5);
#undef JC_MARK
    if (nl) { nl[0] = 1; nl[1] = n_lens; nl[2] = 1; nl[3] = 1; nl[4] = n_contract; }
  }
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

extern "C" int jc_angular_cl_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo,
                                 double* cl_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  return jc_run_pipeline(plan, cosmo_dev, n_cosmo, cl_dev, ws_dev, ws_bytes, (cudaStream_t)stream, 0, 0, nullptr, nullptr, 0);
}

// ---------------------------------------------------------------------------------------------------
// Grid plans: the reference's stand-alone background.* / power.* functions on the caller's (a, k) grid.
// K1 (+ K3 when a power spectrum is requested) run unchanged; this kernel gathers the first n_a nodes.
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) jc_grid_gather_kernel(JcDevPlan pl, Ws ws, const double* __restrict__ cosmo, int chunk,
                                                             double* __restrict__ pk, double* __restrict__ chi,
                                                             double* __restrict__ chi_t, double* __restrict__ growth,
                                                             double* __restrict__ hubble, double* __restrict__ kernels) {
  const int c = blockIdx.y, na = pl.grid_na;
  const int tid = blockIdx.x * 256 + threadIdx.x;
  if (pk && tid < na * pl.L) {
    const int n = tid / pl.L, l = tid - n * pl.L;
    pk[((size_t)c * na + n) * pl.L + l] = ws.vtab[((size_t)c * JC_NA + n) * pl.Lpad + l];
  }
  if (kernels && tid < na * pl.T) {  // [T, n_a] like the reference's (nbins, nz)
    const int t = tid / na, n = tid - t * na;
    kernels[((size_t)c * pl.T + t) * na + n] = ws.rker[((size_t)c * JC_NA_PAD + n) * pl.TS + t];
  }
  if (tid < na) {
    const double x = node_ptr(ws, c, JC_NODE_CHI)[tid];
    if (chi) chi[(size_t)c * na + tid] = x;
    if (chi_t) {  // transverse comoving distance, background.py:297-344 (cosmo.k = -sign(Omega_k), core.py:79-84)
      const double Ok = cosmo[(size_t)c * pl.ncp + 5];
      const double sk = sqrt(fabs(Ok));
      double v = x;
      if (Ok > 0.0) v = JC_RH / sk * sinh(sk * x / JC_RH);
      else if (Ok < 0.0) v = JC_RH / sk * sin(sk * x / JC_RH);
      chi_t[(size_t)c * na + tid] = v;
    }
    if (growth) growth[(size_t)c * na + tid] = node_ptr(ws, c, JC_NODE_GROWTH)[tid];
    if (hubble) hubble[(size_t)c * na + tid] = node_ptr(ws, c, JC_NODE_HUBBLE)[tid];
  }
}
}  // namespace

extern "C" int jc_grid_eval_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, double* pk_dev,
                                double* chi_dev, double* chi_transverse_dev, double* growth_dev, double* hubble_dev,
                                double* transfer_dev, double* kernels_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!plan || !plan->d.grid_mode || !cosmo_dev || !ws_dev || n_cosmo < 1) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);  // a null stream handle means the CURRENT device's default stream
  jc_ws_layout lo;
  int st = jc_workspace_layout(plan, ws_bytes, &lo);
  if (st != JC_OK) return st;
  const JcDevPlan& pl = plan->d;
  cudaStream_t s = (cudaStream_t)stream;
  Ws ws;
  resolve(lo, (double*)ws_dev, 0, &ws);
  const int na = pl.grid_na;
  int per = na * pl.L > na ? na * pl.L : na;
  if (kernels_dev && na * pl.T > per) per = na * pl.T;
  for (int64_t c0 = 0; c0 < n_cosmo; c0 += lo.chunk) {
    const int chunk = (int)((n_cosmo - c0) < lo.chunk ? (n_cosmo - c0) : lo.chunk);
    jc_launch_setup(pl, cosmo_dev + c0 * pl.ncp, ws, chunk, s);
    if (pk_dev) jc_launch_power(pl, ws, chunk, s);
    if (kernels_dev) {
      jc_launch_tracers(pl, ws, chunk, s);
      jc_launch_finish(pl, ws, chunk, s);
    }
    if (transfer_dev) jc_launch_transfer(pl, ws, chunk, transfer_dev + (size_t)c0 * pl.L, s);
    jc_grid_gather_kernel<<<dim3((per + 255) / 256, chunk), 256, 0, s>>>(
        pl, ws, cosmo_dev + c0 * pl.ncp, chunk, pk_dev ? pk_dev + (size_t)c0 * na * pl.L : nullptr,
        chi_dev ? chi_dev + (size_t)c0 * na : nullptr, chi_transverse_dev ? chi_transverse_dev + (size_t)c0 * na : nullptr,
        growth_dev ? growth_dev + (size_t)c0 * na : nullptr, hubble_dev ? hubble_dev + (size_t)c0 * na : nullptr,
        kernels_dev ? kernels_dev + (size_t)c0 * pl.T * na : nullptr);
  }
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

// The remaining background.py functions on a grid plan's scale factors: aux[c][f][n], f = JC_BG_* (jc_b200.h).
// growth_rate comes from K1 (node slot JC_NODE_GEOM of grid plans), the others are closed forms of the cosmology row and
// K1's E^2(a_n) (background.py:52, 90, 168, 196, 294).
namespace {
__global__ void __launch_bounds__(256) jc_grid_background_kernel(JcDevPlan pl, Ws ws, const double* __restrict__ cosmo,
                                                                 double* __restrict__ aux) {
  const int c = blockIdx.y, na = pl.grid_na, n = blockIdx.x * 256 + threadIdx.x;
  if (n >= na) return;
  const double* row = cosmo + (size_t)c * pl.ncp;
  const double Om = row[0] + row[1], Ok = row[5], w0 = row[6], wa = row[7];
  const double Ode = (1.0 - Ok) - Om;                               // core.py:95-96
  const double a = pl.limb_a[n];
  const double hub = node_ptr(ws, c, JC_NODE_HUBBLE)[n] / JC_H0;     // sqrt(E^2)
  const double e2 = hub * hub;
  const double fde = -3.0 * (1.0 + w0 + wa) * pl.limb_lna[n] + 3.0 * wa * (a - 1.0);
  double* o = aux + (size_t)c * JC_BG_FIELDS * na + n;
  o[(size_t)JC_BG_GROWTH_RATE * na] = node_ptr(ws, c, JC_NODE_GEOM)[n];
  o[(size_t)JC_BG_OMEGA_M_A * na] = Om / (a * a * a) / e2;
  o[(size_t)JC_BG_OMEGA_DE_A * na] = Ode * exp(fde) / e2;
  o[(size_t)JC_BG_DCHIOVERDA * na] = JC_RH / (a * a * hub);
  o[(size_t)JC_BG_W * na] = w0 + (1.0 - a) * wa;
  o[(size_t)JC_BG_F_DE * na] = fde;
}

// background.a_of_chi (background.py:245-267): interp(chi, chitab, atab) on K1's DECREASING chi table with the
// reference's rule as written (scipy/interpolate.py:25-37): nearest node by squared distance (first minimum), clipped to
// [1, n-2]; neighbour = sign(clip(x, xp[1], xp[-2]) - xp[ind]) with clip(x, lo, hi) = min(max(x, lo), hi) and lo > hi
// here, i.e. always xp[-2] - xp[ind] <= 0: the left neighbour, except ind = n-2 (sign 0 -> +1).
__global__ void __launch_bounds__(256) jc_a_of_chi_kernel(JcDevPlan pl, Ws ws, const double* __restrict__ chi, int n_chi,
                                                          double* __restrict__ out) {
  const int c = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_chi) return;
  const double* xp = ws.chitab + (size_t)c * JC_NCHI;
  const double x = chi[i];
  int lo = 0, hi = JC_NCHI - 1;  // xp[lo] >= x >= xp[hi] once inside the table
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] >= x) lo = mid; else hi = mid;
  }
  const double dl = (x - xp[lo]) * (x - xp[lo]), dr = (x - xp[hi]) * (x - xp[hi]);
  int ind = dl <= dr ? lo : hi;
  while (ind > 0 && (x - xp[ind - 1]) * (x - xp[ind - 1]) <= (x - xp[ind]) * (x - xp[ind])) --ind;  // ties / plateaus: first minimum
  ind = ind < 1 ? 1 : (ind > JC_NCHI - 2 ? JC_NCHI - 2 : ind);
  const double sgn = xp[JC_NCHI - 2] - xp[ind];
  const int nb = sgn < 0.0 ? ind - 1 : ind + 1;
  const double slope = (pl.chi_pt_a[2 * nb] - pl.chi_pt_a[2 * ind]) / (xp[nb] - xp[ind]);
  const double b = pl.chi_pt_a[2 * ind] - slope * xp[ind];
  out[(size_t)c * n_chi + i] = slope * x + b;
}
}  // namespace

extern "C" int jc_grid_background_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, double* aux_dev,
                                      void* ws_dev, size_t ws_bytes, void* stream) {
  if (!plan || !plan->d.grid_mode || !cosmo_dev || !aux_dev || !ws_dev || n_cosmo < 1) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);  // a null stream handle means the CURRENT device's default stream
  jc_ws_layout lo;
  int st = jc_workspace_layout(plan, ws_bytes, &lo);
  if (st != JC_OK) return st;
  const JcDevPlan& pl = plan->d;
  cudaStream_t s = (cudaStream_t)stream;
  Ws ws;
  resolve(lo, (double*)ws_dev, 0, &ws);
  const int na = pl.grid_na;
  for (int64_t c0 = 0; c0 < n_cosmo; c0 += lo.chunk) {
    const int chunk = (int)((n_cosmo - c0) < lo.chunk ? (n_cosmo - c0) : lo.chunk);
    jc_launch_setup(pl, cosmo_dev + c0 * pl.ncp, ws, chunk, s);
    jc_grid_background_kernel<<<dim3((na + 255) / 256, chunk), 256, 0, s>>>(pl, ws, cosmo_dev + c0 * pl.ncp,
                                                                           aux_dev + (size_t)c0 * JC_BG_FIELDS * na);
  }
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

extern "C" int jc_a_of_chi_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, const double* chi_dev,
                               int64_t n_chi, double* a_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!plan || !cosmo_dev || !chi_dev || !a_dev || !ws_dev || n_cosmo < 1 || n_chi < 1 || n_chi > (1 << 30)) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);  // a null stream handle means the CURRENT device's default stream
  jc_ws_layout lo;
  int st = jc_workspace_layout(plan, ws_bytes, &lo);
  if (st != JC_OK) return st;
  const JcDevPlan& pl = plan->d;
  cudaStream_t s = (cudaStream_t)stream;
  Ws ws;
  resolve(lo, (double*)ws_dev, 0, &ws);
  for (int64_t c0 = 0; c0 < n_cosmo; c0 += lo.chunk) {
    const int chunk = (int)((n_cosmo - c0) < lo.chunk ? (n_cosmo - c0) : lo.chunk);
    jc_launch_setup(pl, cosmo_dev + c0 * pl.ncp, ws, chunk, s);
    jc_a_of_chi_kernel<<<dim3((unsigned)((n_chi + 255) / 256), chunk), 256, 0, s>>>(pl, ws, chi_dev, (int)n_chi,
                                                                                     a_dev + (size_t)c0 * n_chi);
  }
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

extern "C" int jc_sigmasqr_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, const double* R_dev, int32_t n_R,
                               double kmin, double kmax, double* out_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!plan || !cosmo_dev || !R_dev || !out_dev || !ws_dev || n_cosmo < 1 || n_R < 1) return JC_ERR_INVALID;
  if (kmin != 0.0001 || kmax != 1000.0) return JC_ERR_UNSUPPORTED;  // the plan tabulates the reference's default limits
  JcDeviceGuard guard(plan->device);  // a null stream handle means the CURRENT device's default stream
  jc_ws_layout lo;
  int st = jc_workspace_layout(plan, ws_bytes, &lo);
  if (st != JC_OK) return st;
  const JcDevPlan& pl = plan->d;
  cudaStream_t s = (cudaStream_t)stream;
  Ws ws;
  resolve(lo, (double*)ws_dev, 0, &ws);
  for (int64_t c0 = 0; c0 < n_cosmo; c0 += lo.chunk) {
    const int chunk = (int)((n_cosmo - c0) < lo.chunk ? (n_cosmo - c0) : lo.chunk);
    jc_launch_setup(pl, cosmo_dev + c0 * pl.ncp, ws, chunk, s);
    jc_launch_sigmasqr(pl, ws, cosmo_dev + c0 * pl.ncp, chunk, R_dev, n_R, out_dev + (size_t)c0 * n_R, s);
  }
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

// Experiment support: run a subset of the pipeline stages of ONE chunk on `stream` (bit 0 setup, 1 lens, 2 finish, 3 power,
// 4 contraction (persistent TMA kernel), 5 contraction (8-warp cp.async kernel, one CTA per SM)).  scripts/overlap_probe.py
// uses it to time the power kernel of one chunk against the contraction of another on two streams.
extern "C" int jc_debug_stages_f64(const jc_plan* plan, int32_t stage_mask, const double* cosmo_dev, int64_t n_cosmo,
                                   double* cl_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!plan || plan->d.grid_mode || !cosmo_dev || !cl_dev || !ws_dev || n_cosmo < 1) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);  // a null stream handle means the CURRENT device's default stream
  jc_ws_layout lo;
  int st = jc_workspace_layout(plan, ws_bytes, &lo);
  if (st != JC_OK) return st;
  if (n_cosmo > lo.chunk) return JC_ERR_WORKSPACE;
  const JcDevPlan& pl = plan->d;
  cudaStream_t s = (cudaStream_t)stream;
  Ws ws;
  resolve(lo, (double*)ws_dev, 0, &ws);
  const int chunk = (int)n_cosmo;
  if (stage_mask & 1) jc_launch_setup(pl, cosmo_dev, ws, chunk, s);
  if (stage_mask & 2) jc_launch_tracers(pl, ws, chunk, s);
  if (stage_mask & 4) jc_launch_finish(pl, ws, chunk, s);
  if (stage_mask & 8) jc_launch_power(pl, ws, chunk, s);
  if (stage_mask & 16) jc_launch_contract(pl, ws, cl_dev, chunk, s);
  if (stage_mask & 32) jc_launch_contract_1cta(pl, ws, cl_dev, chunk, s);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}
