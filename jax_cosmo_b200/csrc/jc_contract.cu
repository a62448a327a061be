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
// Two kernels share the stage body (mma_stage): jc_contract_tma_kernel (>= 17 pair tiles, the bench configuration:
// persistent CTAs, TMA bulk copies + mbarriers, no CTA-wide barrier -- described at its definition below) and
// jc_contract_kernel (fewer pair tiles, and the baseline the TMA kernel was measured against):
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
#include "jc_tma.cuh"

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
// value image [R | V] of a stage to its tangent image (JVP only).  JVP: 0 = value, 1 = tangent (both products),
// 2 = tangent of a direction with dR = 0 (h, n_s, sigma8: JC_SCAL_MOVES_R): (R_i R_j) . dV only.
template <int KC, int CNT, int NT, int JVP, int M0 = 0>
__device__ __forceinline__ void mma_stage(const double* __restrict__ Rs, const double* __restrict__ Vs,
                                          int TS, int half, int g, int tig, const int (&ti)[2], const int (&tj)[2],
                                          double (&acc)[2][NTW][2]) {
#pragma unroll
  for (int ks = 0; ks < KC / 4; ++ks) {
    const double* rr = Rs + (ks * 4 + tig) * TS;
    const double* vr = Vs + (ks * 4 + tig) * LSV + g;
    double a[CNT], b[NT];
#pragma unroll
    for (int mt = 0; mt < CNT; ++mt) a[mt] = rr[ti[M0 + mt]] * rr[tj[M0 + mt]];
    if (JVP == 0) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b[nt] = vr[nt * 8];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < CNT; ++mt) dmma(acc[M0 + mt][nt][0], acc[M0 + mt][nt][1], a[mt], b[nt]);
    } else if (JVP == 1) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b[nt] = vr[nt * 8];
      double ad[CNT], bd[NT];
#pragma unroll
      for (int mt = 0; mt < CNT; ++mt)
        ad[mt] = fma(rr[half + ti[M0 + mt]], rr[tj[M0 + mt]], rr[ti[M0 + mt]] * rr[half + tj[M0 + mt]]);  // dR_i R_j + R_i dR_j
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) bd[nt] = vr[half + nt * 8];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < CNT; ++mt) {
          dmma(acc[M0 + mt][nt][0], acc[M0 + mt][nt][1], ad[mt], b[nt]);
          dmma(acc[M0 + mt][nt][0], acc[M0 + mt][nt][1], a[mt], bd[nt]);
        }
    } else {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b[nt] = vr[half + nt * 8];  // dV
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < CNT; ++mt) dmma(acc[M0 + mt][nt][0], acc[M0 + mt][nt][1], a[mt], b[nt]);
    }
  }
}

template <int KC, int CNT, int JVP, int M0 = 0>
__device__ __forceinline__ void mma_stage_nt(int ntw, const double* Rs, const double* Vs, int TS, int half, int g,
                                             int tig, const int (&ti)[2], const int (&tj)[2],
                                             double (&acc)[2][NTW][2]) {
  switch (ntw) {  // warp-uniform
    case 7: mma_stage<KC, CNT, 7, JVP, M0>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 6: mma_stage<KC, CNT, 6, JVP, M0>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 5: mma_stage<KC, CNT, 5, JVP, M0>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 4: mma_stage<KC, CNT, 4, JVP, M0>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 3: mma_stage<KC, CNT, 3, JVP, M0>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    case 2: mma_stage<KC, CNT, 2, JVP, M0>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
    default: mma_stage<KC, CNT, 1, JVP, M0>(Rs, Vs, TS, half, g, tig, ti, tj, acc); break;
  }
}

// tangent kernels: does the direction staged in the tangent planes move the tracer kernels of cosmology c?  (setup kernel)
__device__ __forceinline__ bool tangent_moves_r(const Ws& ws, int c) {
  return ws.scal[(size_t)c * JC_SCAL_FIELDS + JC_SCAL_MOVES_R + ws.doff] != 0.0;
}

// Epilogue of the TMA kernel: tiles follow the plan's contraction order (pl.cpair_*): sorted pair q = tile * 8 + g is row
// cpair_out[q] of the output.
__device__ __forceinline__ void store_tiles_ord(const JcDevPlan& pl, const double (&acc)[2][NTW][2], const int (&mtile)[2],
                                                const bool (&has)[2], const int (&ti)[2], const int (&tj)[2], int l0, int g,
                                                int tig, bool vec2, double* __restrict__ out_cosmo) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int q = mtile[mt] * 8 + g;
    if (!has[mt] || q >= pl.P) continue;
    const bool wi = pl.tr_kind[ti[mt]] == JC_TRACER_WEAK_LENSING;
    const bool wj = pl.tr_kind[tj[mt]] == JC_TRACER_WEAK_LENSING;
    double* out = out_cosmo + (size_t)pl.cpair_out[q] * pl.L;
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

// Epilogue of a warp's <= 2 pair tiles: the weak-lensing ell factors (probes.py:73) and the store of C[p, l0:].
__device__ __forceinline__ void store_tiles(const JcDevPlan& pl, const double (&acc)[2][NTW][2], int cnt, int m_first,
                                            const int (&ti)[2], const int (&tj)[2], int l0, int g, int tig, bool vec2,
                                            double* __restrict__ out_cosmo) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int p = (m_first + mt) * 8 + g;
    if (mt >= cnt || p >= pl.P) continue;
    const bool wi = pl.tr_kind[ti[mt]] == JC_TRACER_WEAK_LENSING;
    const bool wj = pl.tr_kind[tj[mt]] == JC_TRACER_WEAK_LENSING;
    double* out = out_cosmo + (size_t)p * pl.L;
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
  const bool with_dr = JVP && tangent_moves_r(ws, c);  // CTA-uniform

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
      if (!JVP) {
        if (cnt == 2) mma_stage_nt<KC, 2, 0>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
        else if (cnt == 1) mma_stage_nt<KC, 1, 0>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
      } else if (with_dr) {
        if (cnt == 2) mma_stage_nt<KC, 2, 1>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
        else if (cnt == 1) mma_stage_nt<KC, 1, 1>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
      } else {
        if (cnt == 2) mma_stage_nt<KC, 2, 2>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
        else if (cnt == 1) mma_stage_nt<KC, 1, 2>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
      }
      if (kc + STAGES - 1 < NKC) load_stage(kc + STAGES - 1);
      cp_async_commit();
    }
    cp_async_wait<0>();

    const bool vec2 = (pl.L & 1) == 0 && (out_cosmo_stride & 1) == 0;
    store_tiles(pl, acc, cnt, m_first, ti, tj, l0, g, tig, vec2, out_base + (size_t)c * out_cosmo_stride);
  }
}

// ---------------------------------------------------------------------------------------------------
// Persistent TMA variant (>= 17 pair tiles, i.e. T >= 16).  ncu on the kernel above (profiles/r01_ncu_summary.md):
// DMMA pipe 73 % busy although one warp per SM sub-partition saturates it from registers
// (scripts/dmma_occupancy.py) -- the idle quarter is warps parked at the per-stage __syncthreads (24 % of
// stall samples) behind the warps that own 2 pair tiles instead of 1, and 14 + 13 tiles never balance over
// the 4 sub-partitions of two independently scheduled CTAs.  Here one CTA of 16 warps (4 per sub-partition,
// 128 registers) per SM walks over cosmologies:
//   * the [R | V] stages arrive by TMA bulk copies (cp.async.bulk, mbarrier complete_tx) in an 8-stage ring that
//     runs on across cosmology boundaries;
//   * there is no CTA-wide barrier and no producer warp (a 17th warp would cost every warp 32 registers): a
//     warp waits on the stage's `full` mbarrier, runs its DMMA burst and arrives on the stage's `empty`
//     mbarrier; the warp holding the fewest tiles of the item waits for `empty` and issues the copies of the
//     stage 8 ahead into that buffer (letting the LAST warp out do it put the issue on the critical path:
//     7.0 ms against 6.4 ms with a separate producer warp at 96 registers);
//   * a warp owns <= 2 pair tiles x <= 7 ell tiles; 27 tiles are dealt 7,7,7,6 over the sub-partitions and the
//     light one rotates from item to item; a warp with one tile runs ahead instead of idling at a barrier.
// R is read KC rows at a time up to row 516 > 513: the finish kernel zeroes the pad rows of R, so the stale
// (finite) V rows of the last stage meet A = 0.
// ---------------------------------------------------------------------------------------------------
constexpr int TMA_CW = 16;      // warps per CTA

// KC nodes per stage (12 or 24), TMA_STAGES ring depth (power of two)
template <int KC, int TMA_STAGES, bool JVP>
__global__ void __launch_bounds__(TMA_CW * 32, 1)
jc_contract_tma_kernel(JcDevPlan pl, Ws ws, double* __restrict__ out_base, int64_t out_cosmo_stride, int chunk, int pairing) {
  constexpr int NKC = (JC_NA + KC - 1) / KC;
  constexpr int NIMG = JVP ? 2 : 1;
  // the last stage holds <= 12 valid nodes: it is consumed as a 12-node stage (3 k-steps) and only the R rows inside
  // the padded table are copied
  constexpr int TAIL_NODES = JC_NA - (NKC - 1) * KC;
  static_assert(KC == 12 && TAIL_NODES <= 12 && (NKC - 1) * KC + 12 <= JC_NA_PAD, "12-node stages: the plan's tile ranges (ctile_lo / ctile_hi) are in units of 12 nodes");
  static_assert(KC < 31, "one lane per V row + lane 31 for R");
  static_assert((TMA_STAGES & (TMA_STAGES - 1)) == 0, "ring index by mask");
  extern __shared__ __align__(16) double smem[];
  const int TS = pl.TS;
  const int half = KC * (TS + LSV);  // one [R | V] image
  const int stage_doubles = NIMG * half;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)TMA_STAGES * stage_doubles);
  uint64_t* empty = full + TMA_STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ngroups = (pl.L + NCOLS - 1) / NCOLS;
  const int mtiles_all = (pl.P + 7) >> 3;
  const int nrounds = (mtiles_all + 2 * TMA_CW - 1) / (2 * TMA_CW);
  const int my_cosmo = (chunk - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_stages = my_cosmo * ngroups * nrounds * NKC;  // stage q = ((ci * ngroups + grp) * nrounds + r) * NKC + kc

  // warp-collective: arm full[q % STAGES] and issue the TMA copies of stage q into that buffer
  auto issue_stage = [&](int q) {
    const int sb = q & (TMA_STAGES - 1);
    const int item = q / NKC, kc = q - item * NKC;
    const int t = item / nrounds;
    const int ci = t / ngroups, grp = t - ci * ngroups;
    const int c = blockIdx.x + ci * gridDim.x;
    const int l0 = grp * NCOLS;
    const int ncols = min(NCOLS, pl.Lpad - l0);  // multiple of 4 doubles: 32-byte pieces
    const int rows = min(KC, JC_NA - kc * KC);
    const int rrows = min(KC, JC_NA_PAD - kc * KC);  // R rows of the stage that exist in the padded table
    double* st = smem + (size_t)sb * stage_doubles;
    if (lane == 0) mbar_expect_tx(full + sb, (unsigned)(NIMG * (rrows * TS + rows * ncols) * sizeof(double)));
    __syncwarp();
#pragma unroll
    for (int img = 0; img < NIMG; ++img) {
      const ptrdiff_t goff = img ? ws.doff : 0;
      if (lane < rows)
        tma_load(st + img * half + KC * TS + lane * LSV,
                 ws.vtab + goff + ((size_t)c * JC_NA + kc * KC + lane) * pl.Lpad + l0, (unsigned)(ncols * sizeof(double)),
                 full + sb);
      else if (lane == 31)
        tma_load(st + img * half, ws.rker + goff + ((size_t)c * JC_NA_PAD + kc * KC) * TS,
                 (unsigned)(rrows * TS * sizeof(double)), full + sb);
    }
  };

  // stale rows of a partially filled last stage must be finite the first time round the ring
  for (int i = threadIdx.x; i < TMA_STAGES * stage_doubles; i += blockDim.x) smem[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int sb = 0; sb < TMA_STAGES; ++sb) {
      mbar_init(full + sb, 1);
      mbar_init(empty + sb, TMA_CW);
    }
    mbar_fence_init();
  }
  fence_proxy_async();  // generic zero fill before async-proxy writes
  __syncthreads();
  if (warp == 0)
    for (int q = 0; q < min(TMA_STAGES, total_stages); ++q) issue_stage(q);

  const int g = lane >> 2, tig = lane & 3;
  const bool vec2 = (pl.L & 1) == 0 && (out_cosmo_stride & 1) == 0;
  int q = 0;
  int rot = blockIdx.x;
  for (int c = blockIdx.x; c < chunk; c += gridDim.x) {
    const bool with_dr = JVP && tangent_moves_r(ws, c);
    for (int grp = 0; grp < ngroups; ++grp) {
      const int l0 = grp * NCOLS;
      const int ntw = (min(pl.L - l0, NCOLS) + 7) >> 3;
      for (int m_base = 0; m_base < mtiles_all; m_base += 2 * TMA_CW, ++rot) {
        // deal the round's pair tiles: the tiles are sorted by the first stage of their node range (plan: ctile_lo),
        // virtual warp v takes tiles v and v + 16, so every prefix of the sorted list -- the tiles active at a given
        // stage -- is spread evenly over the four sub-partitions (= warp % 4); v rotates from item to item so that the
        // sub-partition holding the fewest tiles moves on
        const int m_round = min(2 * TMA_CW, mtiles_all - m_base);
        const int v = (warp + rot) & (TMA_CW - 1);
        // pairing 0: tiles (v, v + 16); 1: (s + 8r, s + 8r + 4) with s = v % 4, r = v / 4 -- the same tiles per
        // sub-partition, neighbours in the sorted order paired in one warp; 2: (2v, 2v + 1)
        const int t0 = pairing == 0 ? v : (pairing == 1 ? (v & 3) + 8 * (v >> 2) : 2 * v);
        const int t1 = pairing == 0 ? v + TMA_CW : (pairing == 1 ? t0 + 4 : t0 + 1);
        const int mtile[2] = {m_base + t0, m_base + t1};
        const bool has[2] = {t0 < m_round, t1 < m_round};
        int ti[2], tj[2], s_lo[2], s_hi[2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int qp = min(mtile[mt] * 8 + g, ((pl.P + 7) & ~7) - 1);  // pad rows repeat the last pair, never stored
          ti[mt] = pl.cpair_i[qp];
          tj[mt] = pl.cpair_j[qp];
          s_lo[mt] = has[mt] ? pl.ctile_lo[mtile[mt]] : 1;
          s_hi[mt] = has[mt] ? pl.ctile_hi[mtile[mt]] : 0;
        }
        double acc[2][NTW][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTW; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        for (int kc = 0; kc < NKC; ++kc, ++q) {
          const int sb = q & (TMA_STAGES - 1);
          mbar_wait(full + sb, (q / TMA_STAGES) & 1);
          const double* Rs = smem + (size_t)sb * stage_doubles;
          const double* Vs = Rs + KC * TS;
          // stages outside a tile's range hold exact zeros (or values below the plan's threshold) of R_i R_j: skipped
          const bool a0 = kc >= s_lo[0] && kc <= s_hi[0], a1 = kc >= s_lo[1] && kc <= s_hi[1];
          if (!JVP) {
            if (a0 && a1) mma_stage_nt<KC, 2, 0, 0>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
            else if (a0) mma_stage_nt<KC, 1, 0, 0>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
            else if (a1) mma_stage_nt<KC, 1, 0, 1>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
          } else if (with_dr) {
            if (a0 && a1) mma_stage_nt<KC, 2, 1, 0>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
            else if (a0) mma_stage_nt<KC, 1, 1, 0>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
            else if (a1) mma_stage_nt<KC, 1, 1, 1>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
          } else {
            if (a0 && a1) mma_stage_nt<KC, 2, 2, 0>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
            else if (a0) mma_stage_nt<KC, 1, 2, 0>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
            else if (a1) mma_stage_nt<KC, 1, 2, 1>(ntw, Rs, Vs, TS, half, g, tig, ti, tj, acc);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty + sb);
          // the lightest-loaded warp of the item (v = 15 holds one tile at most) doubles as the producer: once all 16
          // warps have left the stage it refills the buffer with the stage 8 ahead -- off the critical path
          if (v == TMA_CW - 1 && q + TMA_STAGES < total_stages) {
            mbar_wait(empty + sb, (q / TMA_STAGES) & 1);
            issue_stage(q + TMA_STAGES);
          }
        }
        store_tiles_ord(pl, acc, mtile, has, ti, tj, l0, g, tig, vec2, out_base + (size_t)c * out_cosmo_stride);
      }
    }
  }
}

template <int KC, int TMA_STAGES, bool JVP>
void launch_tma(const JcDevPlan& pl, const Ws& ws, double* out, int64_t stride, int chunk, cudaStream_t s) {
  const size_t smem = (size_t)(JVP ? 2 : 1) * TMA_STAGES * KC * (pl.TS + LSV) * sizeof(double) + 2 * TMA_STAGES * sizeof(uint64_t);
  static int sms = 0;  // idempotent; racing writers set the same values (every device of a box is the same part)
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(jc_contract_tma_kernel<KC, TMA_STAGES, JVP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    sms = n > 0 ? n : 148;
  });
  static int pairing = -1;
  if (pairing < 0) { const char* e = getenv("JC_CONTRACT_PAIRING"); pairing = e ? atoi(e) : 1; }  // tuning knob; measured 5.67 / 5.17 / 5.55 ms for 0 / 1 / 2
  const int avail = sms - pl.reserved_sms > 8 ? sms - pl.reserved_sms : 8;
  jc_contract_tma_kernel<KC, TMA_STAGES, JVP><<<chunk < avail ? chunk : avail, TMA_CW * 32, smem, s>>>(pl, ws, out, stride, chunk, pairing);
}

// TMA bulk copies need 16-byte aligned rows on both sides
bool tma_ok(const JcDevPlan& pl, const Ws& ws) {
  return (pl.P + 7) / 8 > TMA_CW && ((reinterpret_cast<uintptr_t>(ws.rker) | reinterpret_cast<uintptr_t>(ws.vtab)) & 15) == 0 &&
         (ws.doff & 1) == 0 && (pl.TS & 1) == 0 && (pl.Lpad & 3) == 0;
}

template <int KC, int WARPS, int MINB, bool JVP>
void launch_cfg(const JcDevPlan& pl, const Ws& ws, double* out, int64_t stride, int chunk, int msplit, cudaStream_t s) {
  const int ngroups = (pl.L + NCOLS - 1) / NCOLS;
  const size_t smem = (size_t)(JVP ? 2 : 1) * STAGES * KC * (pl.TS + LSV) * sizeof(double);
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, cudaFuncSetAttribute(jc_contract_kernel<KC, WARPS, MINB, JVP>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  jc_contract_kernel<KC, WARPS, MINB, JVP><<<dim3(ngroups * msplit, chunk), WARPS * 32, smem, s>>>(pl, ws, out, stride, msplit);
}

}  // namespace

// jc_set_option("contract_kernel") / env JC_CONTRACT_CFG: 0 = persistent TMA kernel where it applies (default),
// 3 = the 8-warp cp.async kernel everywhere, 1 / 2 = its 16-warp variants (profiles/r01_tuning.md)
int g_contract_cfg = -1;

int jc_contract_init() {
  if (g_contract_cfg < 0) {
    const char* e = getenv("JC_CONTRACT_CFG");
    g_contract_cfg = e ? atoi(e) : 0;
  }
  return JC_OK;
}

void jc_launch_contract(const JcDevPlan& pl, const Ws& ws, double* cl, int chunk, cudaStream_t s) {
  const int mtiles = (pl.P + 7) / 8;
  const int64_t stride = (int64_t)pl.P * pl.L;
  switch (g_contract_cfg) {
    case 1: launch_cfg<12, 16, 1, false>(pl, ws, cl, stride, chunk, 1, s); break;
    // 8 warps, 2 CTAs per SM, pair tiles split over 2 CTAs, cp.async staging (the default up to 16 pair tiles)
    case 3: launch_cfg<12, 8, 2, false>(pl, ws, cl, stride, chunk, mtiles > 16 ? 2 : 1, s); break;
    default:
      if (!tma_ok(pl, ws)) launch_cfg<12, 8, 2, false>(pl, ws, cl, stride, chunk, mtiles > 16 ? 2 : 1, s);
      else if (g_contract_cfg == 2) launch_tma<12, 16, false>(pl, ws, cl, stride, chunk, s);  // 16-stage ring
      else launch_tma<12, 8, false>(pl, ws, cl, stride, chunk, s);
      break;
  }
}

// Experiment support (scripts/overlap_probe.py): the 8-warp cp.async kernel with a shared-memory request that admits one CTA
// per SM next to other kernels' CTAs (co-residency of the contraction with the power kernel).
void jc_launch_contract_1cta(const JcDevPlan& pl, const Ws& ws, double* cl, int chunk, cudaStream_t s) {
  const int mtiles = (pl.P + 7) / 8;
  const int msplit = mtiles > 16 ? 2 : 1;
  const int ngroups = (pl.L + NCOLS - 1) / NCOLS;
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, cudaFuncSetAttribute(jc_contract_kernel<12, 8, 2, false>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  // <.., 2, ..>: the 128-register build; the 120 KB request keeps it at one CTA per SM
  jc_contract_kernel<12, 8, 2, false><<<dim3(ngroups * msplit, chunk), 256, 120 * 1024, s>>>(pl, ws, cl, (int64_t)pl.P * pl.L, msplit);
}

void jc_launch_contract_jvp(const JcDevPlan& pl, const Ws& ws, double* dcl, int64_t dcl_cosmo_stride, int chunk,
                            cudaStream_t s) {
  const int mtiles = (pl.P + 7) / 8;
  if (g_contract_cfg != 3 && tma_ok(pl, ws)) launch_tma<12, 8, true>(pl, ws, dcl, dcl_cosmo_stride, chunk, s);
  else launch_cfg<12, 8, 2, true>(pl, ws, dcl, dcl_cosmo_stride, chunk, mtiles > 16 ? 2 : 1, s);
}
