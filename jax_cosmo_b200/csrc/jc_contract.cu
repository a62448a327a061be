// jc_contract.cu -- K4: tracer-pair contraction on the FP64 tensor-core path.
//
//   C[(i,j), l] = e_i(l) e_j(l) * sum_n R_i[n] R_j[n] V[n, l]          (angular_cl.py:82-96)
//
// is a dense GEMM  C[P x L] = KK[P x 513] . V[513 x L]  per cosmology (210 x 513 x 100 at 10+10 bins),
// where KK[p, n] = R_i(p)[n] R_j(p)[n] is never materialised: each A fragment element is one product
// of two shared-memory loads.  mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) runs at the DFMA peak rate on
// B200 and shares its datapath (measured: jc_fp64_peak_tflops modes 0/1/2 = 36.5/37.2/34.1 TFLOP/s),
// but needs 1/16 of the issue slots and pads P, L only to multiples of 8 (216 x 104) instead of the
// 4x4-block x 32-lane tiling of a scalar kernel (240 x 128).
//
// CTA = (cosmology, group of <= 7 ell-tiles, share of the pair tiles).  The pair tiles (8 pairs each)
// are dealt to the warps 2-or-1 each (balanced over the four SMSPs); a warp holds <= 2 x 7 accumulator
// tiles (28 FP64 accumulators per lane).  R and V stream through a 4-stage cp.async pipeline of KC
// Limber nodes per stage (padding nodes are zero rows); every thread owns fixed copy slots and issues
// the next stage's copies after its MMA burst.  The stage body is branch-free (templated on the tile
// counts) so the fragment loads of a k-step are issued ahead of its MMAs.  Row strides TS and 60 are 4
// or 12 (mod 16): every fragment load is bank-conflict free.
//
// JVP mode (template flag): the tangent  dC = (dR_i R_j + R_i dR_j) . V + (R_i R_j) . dV  with the value
// planes and the tangent planes (ws.doff) of R and V staged side by side; two MMAs per tile and k-step.
#include <cstdlib>

#include "jc_internal.cuh"

namespace {

constexpr int STAGES = 4;
constexpr int NTW = 7;       // ell-tiles (of 8) per warp
constexpr int NCOLS = NTW * 8;
constexpr int LSV = 60;      // shared-memory row stride of a V stage (>= 56, = 12 mod 16)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One pipeline stage (KC/4 k-steps) of a warp's CNT x NT accumulator tiles.  `half` = doubles from the
// value image [R | V] of a stage to its tangent image (JVP only).
template <int KC, int CNT, int NT, bool JVP>
__device__ __forceinline__ void mma_stage(const double* __restrict__ Rs, const double* __restrict__ Vs,
                                          int TS, int half, int g, int tig, const int (&ti)[2], const int (&tj)[2],
                                          double (&acc)[2][NTW][2]) {
#pragma unroll
  for (int ks = 0; ks < KC / 4; ++ks) {
    const double* rr = Rs + (ks * 4 + tig) * TS;
    const double* vr = Vs + (ks * 4 + tig) * LSV + g;
    double a[CNT], b[NT];
#pragma unroll
    for (int mt = 0; mt < CNT; ++mt) a[mt] = rr[ti[mt]] * rr[tj[mt]];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b[nt] = vr[nt * 8];
    if (!JVP) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < CNT; ++mt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    } else {
      double ad[CNT], bd[NT];
#pragma unroll
      for (int mt = 0; mt < CNT; ++mt)
        ad[mt] = fma(rr[half + ti[mt]], rr[tj[mt]], rr[ti[mt]] * rr[half + tj[mt]]);  // dR_i R_j + R_i dR_j
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) bd[nt] = vr[half + nt * 8];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < CNT; ++mt) {
          dmma(acc[mt][nt][0], acc[mt][nt][1], ad[mt], b[nt]);
          dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], bd[nt]);
        }
    }
  }
}

template <int KC, int CNT, bool JVP>
__device__ __forceinline__ void mma_stage_nt(int ntw, const double* Rs, const double* Vs, int TS, int half, int g,
                                             int tig, const int (&ti)[2], const int (&tj)[2],
                                             double (&acc)[2][NTW][2]) {
  switch (ntw) {  // warp-uniform
    case 7: mma_stage<KC, CNT, 7, JVP>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 6: mma_stage<KC, CNT, 6, JVP>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 5: mma_stage<KC, CNT, 5, JVP>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 4: mma_stage<KC, CNT, 4, JVP>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 3: mma_stage<KC, CNT, 3, JVP>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 2: mma_stage<KC, CNT, 2, JVP>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    default: mma_stage<KC, CNT, 1, JVP>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
  }
}

// KC: Limber nodes per pipeline stage (KC/4 k-steps); WARPS per CTA; MINB CTAs per SM; the pair tiles
// are split over gridDim.z CTAs.  out_cosmo_stride: doubles between consecutive cosmologies of `out`.
template <int KC, int WARPS, int MINB, bool JVP>
__global__ void __launch_bounds__(WARPS * 32, MINB)
jc_contract_kernel(JcDevPlan pl, Ws ws, double* __restrict__ out_base, int64_t out_cosmo_stride, int msplit) {
  constexpr int NKC = (JC_NA + KC - 1) / KC;
  constexpr int NIMG = JVP ? 2 : 1;
  constexpr int MAX_SLOTS = (NIMG * KC * (18 + NCOLS / 2) + WARPS * 32 - 1) / (WARPS * 32);  // TS <= 36
  extern __shared__ __align__(16) double smem[];
  const int c = blockIdx.y;
  // blockIdx.x = pair-share z (fastest, so the shares of one cosmology are launched back to back and
  // co-reside on an SM) + msplit * ell-group
  const int zsplit = blockIdx.x % msplit, zgroup = blockIdx.x / msplit;
  const int l0 = zgroup * NCOLS;
  const int ncols = min(NCOLS, pl.Lpad - l0);  // multiple of 4 (Lpad is)
  const int ntw = (min(pl.L - l0, NCOLS) + 7) >> 3;
  const int TS = pl.TS;
  const int half = KC * (TS + LSV);            // one [R | V] image
  const int stage_doubles = NIMG * half;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int mtiles_all = (pl.P + 7) >> 3;
  const int m_per_cta = (mtiles_all + msplit - 1) / msplit;
  const int m_lo = zsplit * m_per_cta;
  const int mtiles = min(mtiles_all, m_lo + m_per_cta);  // this CTA owns pair tiles [m_lo, mtiles)
  const double* Rg = ws.rker + (size_t)c * JC_NA_PAD * TS;
  const double* Vg = ws.vtab + (size_t)c * JC_NA * pl.Lpad + l0;

  // fixed copy slots of this thread: piece q = tid + j*blockDim of the stage image(s) [R | V] (| [dR | dV])
  const double* slot_src[MAX_SLOTS];
  int slot_dst[MAX_SLOTS], slot_row[MAX_SLOTS], slot_step[MAX_SLOTS];
  {
    const int rp = TS >> 1, vp = ncols >> 1;
    const int nR = KC * rp, nV = KC * vp;
#pragma unroll
    for (int j = 0; j < MAX_SLOTS; ++j) {
      int q = threadIdx.x + j * blockDim.x;
      const int img = q / (nR + nV);  // 0 value image, 1 tangent image
      q -= img * (nR + nV);
      const ptrdiff_t goff = img ? ws.doff : 0;
      if (img >= NIMG) {
        slot_src[j] = nullptr; slot_dst[j] = 0; slot_row[j] = 0; slot_step[j] = 0;
      } else if (q < nR) {
        const int r = q / rp, p2 = q - r * rp;
        slot_src[j] = Rg + goff + r * TS + 2 * p2; slot_dst[j] = img * half + r * TS + 2 * p2;
        slot_row[j] = r; slot_step[j] = KC * TS;
      } else {
        const int qv = q - nR;
        const int r = qv / vp, p2 = qv - r * vp;
        slot_src[j] = Vg + goff + (size_t)r * pl.Lpad + 2 * p2; slot_dst[j] = img * half + KC * TS + r * LSV + 2 * p2;
        slot_row[j] = r; slot_step[j] = KC * pl.Lpad;
      }
    }
  }
  auto load_stage = [&](int kc) {
    double* st = smem + (size_t)(kc % STAGES) * stage_doubles;
#pragma unroll
    for (int j = 0; j < MAX_SLOTS; ++j) {
      if (slot_src[j]) {
        if (kc * KC + slot_row[j] < JC_NA) cp_async16(st + slot_dst[j], slot_src[j] + (size_t)kc * slot_step[j]);
        else *reinterpret_cast<double2*>(st + slot_dst[j]) = make_double2(0.0, 0.0);
      }
    }
  };

  for (int m_base = m_lo; m_base < mtiles; m_base += 2 * nwarps) {
    // deal the round's pair tiles to the warps: base or base+1 each, the extras to the lowest warps
    const int m_round = min(2 * nwarps, mtiles - m_base);
    const int base = m_round / nwarps, extra = m_round - base * nwarps;
    // odd shares hand their extras to the highest warps instead, so that the SMSP (= warp % 4) loads of two
    // co-resident shares add up evenly (27 tiles at P = 210: 4,4,3,3 + 3,3,3,4 = 7,7,6,7)
    const bool rev = zsplit & 1;
    const int cnt = base + ((rev ? nwarps - 1 - warp : warp) < extra ? 1 : 0);
    const int m_first = m_base + warp * base + (rev ? max(0, warp - (nwarps - extra)) : min(warp, extra));
    int ti[2], tj[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int p = min((m_first + mt) * 8 + g, pl.P - 1);  // clamped rows are never stored
      ti[mt] = pl.pair_i[p];
      tj[mt] = pl.pair_j[p];
    }
    double acc[2][NTW][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    __syncthreads();  // the previous round's stages are no longer read
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      load_stage(s);
      cp_async_commit();
    }
    for (int kc = 0; kc < NKC; ++kc) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();  // stage kc landed; stage kc-1 (refilled below) is no longer read by anyone
      const double* Rs = smem + (size_t)(kc % STAGES) * stage_doubles;
      const double* Vs = Rs + KC * TS;
      if (cnt == 2) mma_stage_nt<KC, 2, JVP>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
      else if (cnt == 1) mma_stage_nt<KC, 1, JVP>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
      if (kc + STAGES - 1 < NKC) load_stage(kc + STAGES - 1);
      cp_async_commit();
    }
    cp_async_wait<0>();

    const bool vec2 = (pl.L & 1) == 0 && (out_cosmo_stride & 1) == 0;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int p = (m_first + mt) * 8 + g;
      if (mt >= cnt || p >= pl.P) continue;
      const bool wi = pl.tr_kind[ti[mt]] == JC_TRACER_WEAK_LENSING;
      const bool wj = pl.tr_kind[tj[mt]] == JC_TRACER_WEAK_LENSING;
      double* out = out_base + (size_t)c * out_cosmo_stride + (size_t)p * pl.L;
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        const int l = l0 + nt * 8 + 2 * tig;
        double v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double ef = (l + h < pl.L) ? pl.ellfac[l + h] : 1.0;
          v[h] = acc[mt][nt][h] * ((wi ? ef : 1.0) * (wj ? ef : 1.0));  // probes.py:73
        }
        if (vec2 && l + 1 < pl.L) {
          *reinterpret_cast<double2*>(out + l) = make_double2(v[0], v[1]);
        } else {
          if (l < pl.L) out[l] = v[0];
          if (l + 1 < pl.L) out[l + 1] = v[1];
        }
      }
    }
  }
}

template <int KC, int WARPS, int MINB, bool JVP>
void launch_cfg(const JcDevPlan& pl, const Ws& ws, double* out, int64_t stride, int chunk, int msplit, cudaStream_t s) {
  const int ngroups = (pl.L + NCOLS - 1) / NCOLS;
  const size_t smem = (size_t)(JVP ? 2 : 1) * STAGES * KC * (pl.TS + LSV) * sizeof(double);
  static bool attr_done = false;  // idempotent attribute; racing writers set the same value
  if (!attr_done) {
    cudaFuncSetAttribute(jc_contract_kernel<KC, WARPS, MINB, JVP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done = true;
  }
  jc_contract_kernel<KC, WARPS, MINB, JVP><<<dim3(ngroups * msplit, chunk), WARPS * 32, smem, s>>>(pl, ws, out, stride, msplit);
}

int g_contract_cfg = -1;

}  // namespace

int jc_contract_init() {
  if (g_contract_cfg < 0) {
    const char* e = getenv("JC_CONTRACT_CFG");  // tuning knob (profiles/r01_tuning.md)
    g_contract_cfg = e ? atoi(e) : 0;
  }
  return JC_OK;
}

void jc_launch_contract(const JcDevPlan& pl, const Ws& ws, double* cl, int chunk, cudaStream_t s) {
  const int mtiles = (pl.P + 7) / 8;
  const int64_t stride = (int64_t)pl.P * pl.L;
  switch (g_contract_cfg) {
    case 1: launch_cfg<12, 16, 1, false>(pl, ws, cl, stride, chunk, 1, s); break;
    case 2: launch_cfg<24, 16, 1, false>(pl, ws, cl, stride, chunk, 1, s); break;
    // fastest (profiles/r01_tuning.md): 8 warps, 2 CTAs per SM, pair tiles split over 2 CTAs
    default: launch_cfg<12, 8, 2, false>(pl, ws, cl, stride, chunk, mtiles > 16 ? 2 : 1, s); break;
  }
}

void jc_launch_contract_jvp(const JcDevPlan& pl, const Ws& ws, double* dcl, int64_t dcl_cosmo_stride, int chunk,
                            cudaStream_t s) {
  const int mtiles = (pl.P + 7) / 8;
  launch_cfg<12, 8, 2, true>(pl, ws, dcl, dcl_cosmo_stride, chunk, mtiles > 16 ? 2 : 1, s);
}
