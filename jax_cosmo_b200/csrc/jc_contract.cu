// jc_contract.cu -- K4: tracer-pair contraction on the FP64 tensor-core path.
//
//   C[(i,j), l] = e_i(l) e_j(l) * sum_n R_i[n] R_j[n] V[n, l]          (angular_cl.py:82-96)
//
// is a dense GEMM  C[P x L] = KK[P x 513] . V[513 x L]  per cosmology (210 x 513 x 100 at 10+10 bins),
// where KK[p, n] = R_i(p)[n] R_j(p)[n] is never materialised: each A fragment element is one product
// of two shared-memory loads.  mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) runs at the DFMA peak rate on
// B200 (measured: jc_fp64_peak_tflops) but needs 1/16 of the issue slots and pads P,L only to
// multiples of 8 (216 x 104) instead of the 4x4-block x 32-lane tiling of a scalar kernel (240 x 128).
//
// CTA = (cosmology, group of <= 7 ell-tiles); warp = 2 pair-tiles (16 pairs) x 7 ell-tiles, 28 FP64
// accumulators per lane.  R and V stream through a 4-stage cp.async pipeline of 12 Limber nodes per
// stage (43 stages = 516 nodes, the 3 padding nodes are zero rows).  Row strides TS and 60 are
// 4 or 12 (mod 16) so every fragment load is bank-conflict free.
#include "jc_internal.cuh"

namespace {

constexpr int KC = 12;       // Limber nodes per pipeline stage (3 k-steps of 4)
constexpr int STAGES = 4;
constexpr int NTW = 7;       // ell-tiles (of 8) per warp
constexpr int NCOLS = NTW * 8;
constexpr int LSV = 60;      // shared-memory row stride of a V stage (>= 56, = 12 mod 16)
constexpr int NKC = (JC_NA + KC - 1) / KC;  // 43
constexpr int MAX_WARPS = 16;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(MAX_WARPS * 32)
jc_contract_kernel(JcDevPlan pl, Ws ws, double* __restrict__ cl) {
  extern __shared__ __align__(16) double smem[];
  const int c = blockIdx.y;
  const int l0 = blockIdx.x * NCOLS;
  const int ncols = min(NCOLS, pl.Lpad - l0);  // multiple of 4 (Lpad is)
  const int ntw = (min(pl.L - l0, NCOLS) + 7) >> 3;
  const int TS = pl.TS;
  const int stage_doubles = KC * (TS + LSV);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int mtiles = (pl.P + 7) >> 3;
  const int ntask = (mtiles + 1) >> 1;
  const double* Rg = ws.rker + (size_t)c * JC_NA_PAD * TS;
  const double* Vg = ws.vtab + (size_t)c * JC_NA * pl.Lpad + l0;

  auto load_stage = [&](int kc) {
    double* Rs = smem + (size_t)(kc % STAGES) * stage_doubles;
    double* Vs = Rs + KC * TS;
    const int n0 = kc * KC;
    const int rp = TS >> 1;  // 16-byte pieces per R row
    for (int q = threadIdx.x; q < KC * rp; q += blockDim.x) {
      const int r = q / rp, p2 = q - r * rp;
      double* dst = Rs + r * TS + 2 * p2;
      if (n0 + r < JC_NA) cp_async16(dst, Rg + (size_t)(n0 + r) * TS + 2 * p2);
      else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
    }
    const int vp = ncols >> 1;
    for (int q = threadIdx.x; q < KC * vp; q += blockDim.x) {
      const int r = q / vp, p2 = q - r * vp;
      double* dst = Vs + r * LSV + 2 * p2;
      if (n0 + r < JC_NA) cp_async16(dst, Vg + (size_t)(n0 + r) * pl.Lpad + 2 * p2);
      else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
    }
  };

  for (int round = 0; round * nwarps < ntask; ++round) {
    const int task = round * nwarps + warp;
    const bool active = task < ntask;
    // this lane's pair rows in the two m-tiles -> tracer indices (clamped rows are never stored)
    int ti[2], tj[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int p = min((2 * task + mt) * 8 + g, pl.P - 1);
      ti[mt] = active ? pl.pair_i[p] : 0;
      tj[mt] = active ? pl.pair_j[p] : 0;
    }
    double acc[2][NTW][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    __syncthreads();  // previous round's stages are no longer read
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      load_stage(s);
      cp_async_commit();
    }
    for (int kc = 0; kc < NKC; ++kc) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      if (kc + STAGES - 1 < NKC) load_stage(kc + STAGES - 1);
      cp_async_commit();
      if (active) {
        const double* Rs = smem + (size_t)(kc % STAGES) * stage_doubles;
        const double* Vs = Rs + KC * TS;
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
          const double* rr = Rs + (ks * 4 + tig) * TS;
          const double a0 = rr[ti[0]] * rr[tj[0]];
          const double a1 = rr[ti[1]] * rr[tj[1]];
          const double* vr = Vs + (ks * 4 + tig) * LSV + g;
#pragma unroll
          for (int nt = 0; nt < NTW; ++nt) {
            if (nt < ntw) {
              const double b = vr[nt * 8];
              dmma(acc[0][nt][0], acc[0][nt][1], a0, b);
              dmma(acc[1][nt][0], acc[1][nt][1], a1, b);
            }
          }
        }
      }
    }
    cp_async_wait<0>();
    if (active) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int p = (2 * task + mt) * 8 + g;
        if (p >= pl.P) continue;
        const bool wi = pl.tr_kind[ti[mt]] == JC_TRACER_WEAK_LENSING;
        const bool wj = pl.tr_kind[tj[mt]] == JC_TRACER_WEAK_LENSING;
        double* out = cl + ((size_t)c * pl.P + p) * pl.L;
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int l = l0 + nt * 8 + 2 * tig + h;
            if (l < pl.L) {
              const double ef = pl.ellfac[l];
              out[l] = acc[mt][nt][h] * ((wi ? ef : 1.0) * (wj ? ef : 1.0));  // probes.py:73
            }
          }
        }
      }
    }
  }
}

}  // namespace

int jc_contract_init() {
  JC_CUDA_TRY(cudaFuncSetAttribute(jc_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  return JC_OK;
}

void jc_launch_contract(const JcDevPlan& pl, const Ws& ws, double* cl, int chunk, cudaStream_t s) {
  const int mtiles = (pl.P + 7) / 8;
  const int ntask = (mtiles + 1) / 2;
  const int warps = ntask < MAX_WARPS ? ntask : MAX_WARPS;
  const int ngroups = (pl.L + NCOLS - 1) / NCOLS;
  const size_t smem = (size_t)STAGES * KC * (pl.TS + LSV) * sizeof(double);
  jc_contract_kernel<<<dim3(ngroups, chunk), warps * 32, smem, s>>>(pl, ws, cl);
}
