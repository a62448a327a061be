// jc_tracers.cu -- K2: radial tracer kernels R_i(a_n), node-major [n][TS].
//
// K2a  lensing efficiency  q_s(z_n) = int_{z_n}^{zmax} n_s(z') max(chi'-chi_n,0)/max(chi',1) dz'
//      (probes.py:44-51).  The z' grid (linspace(z_n, zmax, 257) per node), its chi-table brackets and
//      the normalised n_s(z') are cosmology independent (plan tables, L2 resident); per cosmology only
//      the 256-entry chi table differs.  CTA = 128 nodes x 16 cosmologies: warp = (32 nodes, 4
//      cosmologies); the 4 warps that share a node block read the same table lines (L1 hits), so L2
//      traffic per cosmology is 1/16 of the table set.  g is computed once for all sources.
//      Epilogue: WL = (q (1+z) chi 3 H0^2 Om/(2c) + NLA) (1+m)    probes.py:51,71-74,102-129,201-207
// K2b  finish: NC = n_i(z) b_i(z) H(a)                            probes.py:77-99
//              + delta_nz source planes and node 512 of the extended sources
// Both are templates on the scalar type (double / DualN<K>, see jc_dual.cuh).
#include <cstdlib>

#include "jc_internal.cuh"
#include "jc_dual.cuh"

namespace {

constexpr int LENS_NODES = 128;  // nodes per CTA
constexpr int LENS_MR = 4;       // z' rows per pipeline stage
constexpr int LENS_STAGES = 3;

// NCOS: cosmologies per thread; LENS_CGROUPS: cosmology groups (of 4 warps) per CTA
// skip_static (tangent passes whose VALUE tables nobody reads -- the second and later tangent groups of the reverse-sweep path):
// when none of the pass's directions can move the tracer kernels (JC_SCAL_MOVES_R, same for every cosmology of a batch), dR = 0
// identically, the tangent contraction never reads it, and the pass has nothing to produce.
template <class T>
__device__ __forceinline__ bool pass_is_static(const Ws& ws, int c) {
  bool any = false;
  if constexpr (JxTangents<T>::N > 0) {
#pragma unroll
    for (int k = 0; k < JxTangents<T>::N; ++k)
      any = any || ws.scal[(size_t)c * JC_SCAL_FIELDS + JC_SCAL_MOVES_R + (ptrdiff_t)(k + 1) * ws.doff] != 0.0;
  } else {
    any = true;
  }
  return !any;
}

template <class T, int NS, int NCOS, int LENS_CGROUPS>
__global__ void __launch_bounds__(LENS_NODES * LENS_CGROUPS)
jc_lens_kernel(JcDevPlan pl, Ws ws, int n_cosmo, int s0, int skip_static) {
  constexpr int CTA_COSMO = NCOS * LENS_CGROUPS;
  __shared__ T s_chit[CTA_COSMO][JC_NCHI];
  if (skip_static && pass_is_static<T>(ws, 0)) return;  // uniform over the grid: before any barrier or copy
  const ptrdiff_t doff = ws.doff;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * LENS_NODES + (warp & 3) * 32 + lane;  // node
  const int cg = warp >> 2;                                         // cosmology group in CTA
  const int cta_c0 = blockIdx.y * CTA_COSMO;
  for (int i = threadIdx.x; i < CTA_COSMO * JC_NCHI; i += blockDim.x) {
    const int cc = min(cta_c0 + i / JC_NCHI, n_cosmo - 1);
    s_chit[i / JC_NCHI][i % JC_NCHI] = JxMem<T>::ld(ws.chitab + (size_t)cc * JC_NCHI + (i % JC_NCHI), doff);
  }
  const int c0 = cta_c0 + cg * NCOS;
  T chin[NCOS];
#pragma unroll
  for (int c = 0; c < NCOS; ++c) chin[c] = JxMem<T>::ld(node_ptr(ws, min(c0 + c, n_cosmo - 1), JC_NODE_CHI) + n, doff);
  __syncthreads();
  const T(*chit)[JC_NCHI] = s_chit + cg * NCOS;
  const size_t NL = (size_t)JC_NLENS * JC_NLENS_COLS;
  T acc[NCOS][NS];
#pragma unroll
  for (int c = 0; c < NCOS; ++c)
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[c][s] = T(0.0);

  // The [257 x 128] slices of lens_t / lens_ix / lens_nw (sources s0..s0+NS-1, consecutive slabs) of this
  // CTA's node block stream through a LENS_STAGES-deep cp.async pipeline of LENS_MR z'-rows per stage;
  // all 16 warps read them from shared memory (the kernel was L2-latency bound with direct loads:
  // FP64 pipe 26 % at 16 warps/SM, profiles/r01_ncu_summary.md).
  extern __shared__ __align__(16) unsigned char lens_smem[];
  constexpr int ROW_T = LENS_NODES * 8, ROW_IX = LENS_NODES * 2;          // bytes per z'-row
  constexpr int STAGE_BYTES = LENS_MR * (ROW_T * (1 + NS) + ROW_IX);
  constexpr int OFF_NW = LENS_MR * ROW_T, OFF_IX = LENS_MR * ROW_T * (1 + NS);
  constexpr int PIECES = STAGE_BYTES / 16;
  constexpr int SLOTS = (PIECES + LENS_NODES * LENS_CGROUPS - 1) / (LENS_NODES * LENS_CGROUPS);
  constexpr int NSTAGE = (JC_NLENS + LENS_MR - 1) / LENS_MR;
  const int node0 = blockIdx.x * LENS_NODES;
  // Copy slots: a running global pointer and a 32-bit shared-memory offset per slot, validity and row kind as bit
  // masks.  load_stage is called with st = 0, 1, 2, ... so the pointers simply advance by one stage per call (ncu: the
  // indexed form -- row test, 64-bit multiply-add and generic-to-shared conversion per slot -- was 160 of the 754
  // instructions a warp executes per stage).
  constexpr int ROWS_LAST = JC_NLENS - (NSTAGE - 1) * LENS_MR;
  constexpr int STEP_T = LENS_MR * JC_NLENS_COLS * 8, STEP_IX = LENS_MR * JC_NLENS_COLS * 2;
  const unsigned char* slot_src[SLOTS];
  unsigned slot_dst[SLOTS];
  unsigned valid = 0, tail_ok = 0, is_ix = 0;
#pragma unroll
  for (int j = 0; j < SLOTS; ++j) {
    const int q = threadIdx.x + j * (LENS_NODES * LENS_CGROUPS);
    const int b = q * 16;  // byte offset inside the stage image
    int r = 0;
    slot_dst[j] = (unsigned)b;
    if (q >= PIECES) {
      slot_src[j] = nullptr;
    } else if (b < OFF_NW) {  // t rows
      r = b / ROW_T;
      slot_src[j] = (const unsigned char*)(pl.lens_t + (size_t)r * JC_NLENS_COLS + node0) + (b - r * ROW_T);
    } else if (b < OFF_IX) {  // nw rows: [row][source][node]
      const int bb = b - OFF_NW;
      r = bb / (ROW_T * NS);
      const int s = (bb - r * ROW_T * NS) / ROW_T, o = bb - (r * NS + s) * ROW_T;
      slot_src[j] = (const unsigned char*)(pl.lens_nw + (size_t)(s0 + s) * NL + (size_t)r * JC_NLENS_COLS + node0) + o;
    } else {  // ix rows (uint16)
      const int bb = b - OFF_IX;
      r = bb / ROW_IX;
      slot_src[j] = (const unsigned char*)(pl.lens_ix + (size_t)r * JC_NLENS_COLS + node0) + (bb - r * ROW_IX);
      is_ix |= 1u << j;
    }
    if (q < PIECES) valid |= 1u << j;
    if (q < PIECES && r < ROWS_LAST) tail_ok |= 1u << j;
  }
  const unsigned smem0 = (unsigned)__cvta_generic_to_shared(lens_smem);
  auto load_stage = [&](int st) {
    const unsigned sbase = smem0 + (unsigned)(st % LENS_STAGES) * STAGE_BYTES;
    const unsigned ok = st == NSTAGE - 1 ? tail_ok : valid;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      if ((ok >> j) & 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sbase + slot_dst[j]), "l"(slot_src[j]));
      slot_src[j] += ((is_ix >> j) & 1) ? STEP_IX : STEP_T;
    }
  };
#pragma unroll
  for (int st = 0; st < LENS_STAGES - 1; ++st) {
    load_stage(st);
    asm volatile("cp.async.commit_group;\n" ::);
  }
  const int nl = (warp & 3) * 32 + lane;  // node within the CTA's block
  for (int st = 0; st < NSTAGE; ++st) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(LENS_STAGES - 2));
    __syncthreads();  // stage st landed; stage st-1 (refilled below) is no longer read
    if (st + LENS_STAGES - 1 < NSTAGE) load_stage(st + LENS_STAGES - 1);
    asm volatile("cp.async.commit_group;\n" ::);
    const unsigned char* base = lens_smem + (size_t)(st % LENS_STAGES) * STAGE_BYTES;
    const double* st_t = (const double*)base + nl;
    const double* st_nw = (const double*)(base + OFF_NW) + nl;
    const unsigned short* st_ix = (const unsigned short*)(base + OFF_IX) + nl;
    const int rows = min(LENS_MR, JC_NLENS - st * LENS_MR);
    // interpolation weights and brackets of all rows of the stage up front: the chi-table loads of a row depend on its
    // bracket, so loading them row by row put two shared-memory latencies in front of every row (ncu: 30 % of the stall
    // samples on the short scoreboard)
    double tr[LENS_MR];
    int ixr[LENS_MR];
#pragma unroll
    for (int r = 0; r < LENS_MR; ++r) {
      tr[r] = st_t[r * LENS_NODES];
      ixr[r] = st_ix[r * LENS_NODES];
    }
#pragma unroll
    for (int r = 0; r < LENS_MR; ++r) {
      if (r < rows) {
        const double t = tr[r];
        const int i0 = ixr[r] & 255, i1 = ixr[r] >> 8;
        // the NCOS lensing weights first, then one n(z') load per source feeding NCOS accumulators: NCOS live values
        // instead of NS (the kernel sits at the 128-register cap with NCOS x NS accumulators)
        T g[NCOS];
#pragma unroll
        for (int c = 0; c < NCOS; ++c) {
          const T f0 = chit[c][i0], f1 = chit[c][i1];
          const T chip = jx_clip0(f0 + (f1 - f0) * t);                          // background.py:242
          g[c] = jx_clip0(chip - chin[c]) * jx_rcp(jx_floor1(chip));            // probes.py:49
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const double w = st_nw[(r * NS + s) * LENS_NODES];
#pragma unroll
          for (int c = 0; c < NCOS; ++c) acc[c][s] = acc[c][s] + w * g[c];
        }
      }
    }
  }
  // Epilogue = the lensing term of the tracer kernel, written straight into R[n][t] (the finish kernel then only
  // touches number counts, delta planes, node 512 and -- in place -- the NLA term of IA-enabled sources):
  //   q (1+z) chi 3 H0^2 Om / (2c) (1+m)                                                   probes.py:51,71-74,201-207
  const double zfac = (pl.lens_zmax - pl.limb_z[n]) /* simps: dx * N (probes.py:51) */ * (1.0 + pl.limb_z[n]) *
                      (3.0 * JC_H0 * JC_H0 / 2.0 / JC_C_LIGHT);
#pragma unroll
  for (int c = 0; c < NCOS; ++c) {
    const int cc = c0 + c;
    if (cc >= n_cosmo) break;
    double* row = ws.rker + ((size_t)cc * JC_NA_PAD + n) * pl.TS;
    const T Om = JxMem<T>::ld(ws.scal + (size_t)cc * JC_SCAL_FIELDS + JC_SCAL_OMEGA_M, doff);
    const T amp = (zfac * chin[c]) * Om;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int t = pl.src_tracer[s0 + s];
      JxMem<T>::st(row + t, doff, acc[c][s] * amp * pl.tr_m1[t]);  // the NLA term is added by the finish kernel
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K2a on the FP64 tensor-core instruction (double path, >= 8 sources per launch).
//
// ncu on jc_lens_kernel (profiles/r02_ncu_summary.md): FP64 pipe 44 %, issue slots 50 %, one eligible warp per scheduler -- a
// thread spends 41 issue slots per (node, z', cosmology) item, 18 of them on the FP64 pipe: 8 for the lensing weight
// g = max(chi' - chi_n, 0) / max(chi', 1) and 10 multiply-adds q_s += w_s g, fed by 10 weight loads.  The source sums are a
// matrix product per node,  q[s][c] = sum_z' w[s][z'] g[z'][c]  (sources x cosmologies, z' as the k dimension), with the
// weights independent of the cosmology.  Here a warp owns TWO nodes and 32 cosmologies: per k-step of four z' every lane
// computes g for (z' = lane % 4, cosmology = lane / 4 + 8 q, q = 0..3) of both nodes -- exactly the B fragment of
// mma.m8n8k4.f64 -- and loads its A-fragment element w[source = lane / 4][z' = lane % 4] of both nodes with one LDS.128; eight
// DMMAs replace 64 DFMAs and 8 x 8 weight loads.  Sources beyond the eighth stay scalar (per-lane partial sums over the lane's
// z' residue, reduced over the four lanes of a group at the end): a second m-tile for two rows would cost 6 x their pipe time.
// DMMA shares the FP64 datapath, so the pipe time per item is unchanged (18 units); the issue slots fall from 41 to ~24 and a
// lane carries 8 independent items per k-step.
// CTA = 16 warps = 32 nodes x 32 cosmologies; the chi tables of the 32 cosmologies sit in shared memory with a row stride of
// 258 doubles (the 8 cosmologies of a quarter-warp hit different banks); t / ix / w stream through a 3-stage cp.async ring of
// 12 z'-rows, rows padded so that every fragment load is conflict free (units of 16 bytes: source stride 20, row stride
// = 1 mod 8).  z' rows 257..263 of the last stage are zero-filled (w = 0, bracket 0: g finite).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LM_NODES = 32;       // nodes per CTA (2 per warp)
constexpr int LM_COSMO = 32;       // cosmologies per CTA (4 n-tiles)
constexpr int LM_ZS = 12;          // z' rows per stage (3 k-steps)
constexpr int LM_STAGES = 3;
constexpr int LM_CHS = JC_NCHI + 2;  // chi-table row stride

__device__ __forceinline__ void lens_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NS> struct LmLayout {  // byte offsets inside one stage image
  static constexpr int RS_S = 20 * 16;                                         // source stride: 32 nodes (256 B) + 64 B
  static constexpr int RS_J = (NS * 20 + ((1 - NS * 20) % 8 + 8) % 8) * 16;      // z'-row stride of w: = 1 (mod 8) units
  static constexpr int RS_T = 17 * 16, RS_IX = 5 * 16;
  static constexpr int OFF_W = 0, OFF_T = LM_ZS * RS_J, OFF_IX = OFF_T + LM_ZS * RS_T;
  static constexpr int STAGE = OFF_IX + LM_ZS * RS_IX;
  static constexpr int PIECES_W = LM_ZS * NS * 16, PIECES_T = LM_ZS * 16, PIECES_IX = LM_ZS * 4;
  static constexpr int PIECES = PIECES_W + PIECES_T + PIECES_IX;
};

template <int MT, int NX>
__global__ void __launch_bounds__(512, 1) jc_lens_mma_kernel(JcDevPlan pl, Ws ws, int n_cosmo, int s0) {
  constexpr int NS = 8 * MT + NX;
  typedef LmLayout<NS> LY;
  constexpr int NQ = LM_COSMO / 8;
  constexpr int SLOTS = (LY::PIECES + 511) / 512;
  constexpr int NSTAGE = (JC_NLENS + LM_ZS - 1) / LM_ZS;
  constexpr int ROWS_LAST = JC_NLENS - (NSTAGE - 1) * LM_ZS;
  extern __shared__ __align__(16) unsigned char lm_smem[];
  double* s_chit = reinterpret_cast<double*>(lm_smem);                       // [LM_COSMO][LM_CHS]
  unsigned char* s_stage = lm_smem + (size_t)LM_COSMO * LM_CHS * sizeof(double);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const int node0 = blockIdx.x * LM_NODES, na = node0 + 2 * warp;           // the warp's nodes: na, na + 1
  const int cta_c0 = blockIdx.y * LM_COSMO;
  for (int i = threadIdx.x; i < LM_COSMO * JC_NCHI; i += 512) {
    const int c = i >> 8, k = i & 255;
    s_chit[c * LM_CHS + k] = ws.chitab[(size_t)min(cta_c0 + c, n_cosmo - 1) * JC_NCHI + k];
  }
  double chin[2][NQ];  // chi at the warp's nodes for this lane's B columns (cosmology g + 8 q)
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const double* cn = node_ptr(ws, min(cta_c0 + q * 8 + g, n_cosmo - 1), JC_NODE_CHI) + na;
    chin[0][q] = cn[0];
    chin[1][q] = cn[1];
  }
  // ---- copy slots (16-byte pieces of a stage image; the global pointers advance by one stage per call) ----
  const size_t NL = (size_t)JC_NLENS * JC_NLENS_COLS;
  const unsigned char* slot_src[SLOTS];
  unsigned slot_dst[SLOTS], slot_step[SLOTS];
  unsigned valid = 0, tail_ok = 0;
#pragma unroll
  for (int j = 0; j < SLOTS; ++j) {
    const int q = threadIdx.x + j * 512;
    int r = 0;
    slot_src[j] = nullptr; slot_dst[j] = 0; slot_step[j] = 0;
    if (q < LY::PIECES_W) {                       // w: piece = (row r, source s, 16-byte unit u of the 32-node row)
      r = q / (NS * 16);
      const int s = (q - r * NS * 16) >> 4, u = q & 15;
      slot_src[j] = (const unsigned char*)(pl.lens_nw + (size_t)(s0 + s) * NL + (size_t)r * JC_NLENS_COLS + node0) + u * 16;
      slot_dst[j] = LY::OFF_W + r * LY::RS_J + s * LY::RS_S + u * 16;
      slot_step[j] = LM_ZS * JC_NLENS_COLS * 8;
    } else if (q < LY::PIECES_W + LY::PIECES_T) {
      const int qq = q - LY::PIECES_W;
      r = qq >> 4;
      slot_src[j] = (const unsigned char*)(pl.lens_t + (size_t)r * JC_NLENS_COLS + node0) + (qq & 15) * 16;
      slot_dst[j] = LY::OFF_T + r * LY::RS_T + (qq & 15) * 16;
      slot_step[j] = LM_ZS * JC_NLENS_COLS * 8;
    } else if (q < LY::PIECES) {
      const int qq = q - LY::PIECES_W - LY::PIECES_T;
      r = qq >> 2;
      slot_src[j] = (const unsigned char*)(pl.lens_ix + (size_t)r * JC_NLENS_COLS + node0) + (qq & 3) * 16;
      slot_dst[j] = LY::OFF_IX + r * LY::RS_IX + (qq & 3) * 16;
      slot_step[j] = LM_ZS * JC_NLENS_COLS * 2;
    }
    if (q < LY::PIECES) valid |= 1u << j;
    if (q < LY::PIECES && r < ROWS_LAST) tail_ok |= 1u << j;
  }
  const unsigned smem0 = (unsigned)__cvta_generic_to_shared(s_stage);
  auto load_stage = [&](int st) {
    const unsigned sbase = smem0 + (unsigned)(st % LM_STAGES) * LY::STAGE;
    const bool last = st == NSTAGE - 1;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      if ((valid >> j) & 1) {
        if (!last || ((tail_ok >> j) & 1))
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sbase + slot_dst[j]), "l"(slot_src[j]));
        else  // z' rows past the rule's last node: zero weights, bracket (0, 0), t = 0
          asm volatile("st.shared.v2.f64 [%0], {%1, %1};\n" ::"r"(sbase + slot_dst[j]), "d"(0.0));
      }
      slot_src[j] += slot_step[j];
    }
  };
#pragma unroll
  for (int st = 0; st < LM_STAGES - 1; ++st) {
    load_stage(st);
    asm volatile("cp.async.commit_group;\n" ::);
  }
  double acc[2][NQ][MT][2];   // D fragments: [node][n-tile][m-tile]{cosmology 2 tig, 2 tig + 1} of source g + 8 mt
  double accx[2][NQ][NX > 0 ? NX : 1];  // scalar sources: partial sums over z' = tig (mod 4) for cosmology g + 8 q
#pragma unroll
  for (int nd = 0; nd < 2; ++nd)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) acc[nd][q][mt][0] = acc[nd][q][mt][1] = 0.0;
#pragma unroll
      for (int x = 0; x < (NX > 0 ? NX : 1); ++x) accx[nd][q][x] = 0.0;
    }
  const unsigned pair_off = (unsigned)(2 * warp) * 8;  // byte offset of the warp's node pair inside a 32-node row
  const double* chq = s_chit + g * LM_CHS;

  for (int st = 0; st < NSTAGE; ++st) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(LM_STAGES - 2));
    __syncthreads();  // stage st landed (and its zero fill is visible); stage st - 1, refilled below, is no longer read
    if (st + LM_STAGES - 1 < NSTAGE) load_stage(st + LM_STAGES - 1);
    asm volatile("cp.async.commit_group;\n" ::);
    const unsigned char* base = s_stage + (size_t)(st % LM_STAGES) * LY::STAGE;
#pragma unroll
    for (int ks = 0; ks < LM_ZS / 4; ++ks) {
      const int j = ks * 4 + tig;
      const double2 t2 = *reinterpret_cast<const double2*>(base + LY::OFF_T + j * LY::RS_T + pair_off);
      const unsigned ix2 = *reinterpret_cast<const unsigned*>(base + LY::OFF_IX + j * LY::RS_IX + 4 * warp);
      double2 a2[MT], wx2[NX > 0 ? NX : 1];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        a2[mt] = *reinterpret_cast<const double2*>(base + LY::OFF_W + j * LY::RS_J + (mt * 8 + g) * LY::RS_S + pair_off);
#pragma unroll
      for (int x = 0; x < NX; ++x)
        wx2[x] = *reinterpret_cast<const double2*>(base + LY::OFF_W + j * LY::RS_J + (MT * 8 + x) * LY::RS_S + pair_off);
#pragma unroll
      for (int nd = 0; nd < 2; ++nd) {
        const double t = nd ? t2.y : t2.x;
        const unsigned ix = nd ? (ix2 >> 16) : (ix2 & 0xffffu);
        const int i0 = ix & 255, i1 = ix >> 8;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const double f0 = chq[q * 8 * LM_CHS + i0], f1 = chq[q * 8 * LM_CHS + i1];
          const double chip = jx_clip0(f0 + (f1 - f0) * t);                                   // background.py:242
          const double gv = jx_clip0(chip - chin[nd][q]) * jx_rcp(jx_floor1(chip));          // probes.py:49
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) lens_dmma(acc[nd][q][mt][0], acc[nd][q][mt][1], nd ? a2[mt].y : a2[mt].x, gv);
#pragma unroll
          for (int x = 0; x < NX; ++x) accx[nd][q][x] = fma(nd ? wx2[x].y : wx2[x].x, gv, accx[nd][q][x]);
        }
      }
    }
  }
  // ---- epilogue: q (1+z) chi 3 H0^2 Om / (2c) (1+m) straight into R[n][t]                          probes.py:51,71-74,201-207
#pragma unroll
  for (int nd = 0; nd < 2; ++nd) {
    const int n = na + nd;
    const double zfac = (pl.lens_zmax - pl.limb_z[n]) * (1.0 + pl.limb_z[n]) * (3.0 * JC_H0 * JC_H0 / 2.0 / JC_C_LIGHT);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      // tensor-core sources: this lane holds source g + 8 mt for cosmologies 8 q + 2 tig, + 1
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = cta_c0 + q * 8 + 2 * tig + h;
        if (cc < n_cosmo) {
          const double amp = (zfac * node_ptr(ws, cc, JC_NODE_CHI)[n]) * ws.scal[(size_t)cc * JC_SCAL_FIELDS + JC_SCAL_OMEGA_M];
          double* row = ws.rker + ((size_t)cc * JC_NA_PAD + n) * pl.TS;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int tr = pl.src_tracer[s0 + mt * 8 + g];
            row[tr] = acc[nd][q][mt][h] * amp * pl.tr_m1[tr];
          }
        }
      }
      // scalar sources: sum the four z' residues of the group; lane tig = x % 4 of group g stores source 8 MT + x
#pragma unroll
      for (int x = 0; x < NX; ++x) {
        double v = accx[nd][q][x];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const int cc = cta_c0 + q * 8 + g;
        if (tig == (x & 3) && cc < n_cosmo) {
          const double amp = (zfac * chin[nd][q]) * ws.scal[(size_t)cc * JC_SCAL_FIELDS + JC_SCAL_OMEGA_M];
          const int tr = pl.src_tracer[s0 + MT * 8 + x];
          ws.rker[((size_t)cc * JC_NA_PAD + n) * pl.TS + tr] = v * amp * pl.tr_m1[tr];
        }
      }
    }
  }
}

template <int MT, int NX>
void launch_lens_mma(const JcDevPlan& pl, const Ws& ws, int chunk, int s0, cudaStream_t st) {
  typedef LmLayout<8 * MT + NX> LY;
  constexpr int smem = LM_COSMO * LM_CHS * (int)sizeof(double) + LM_STAGES * LY::STAGE;
  static_assert(smem <= 227 * 1024, "lens stage ring + chi tables must fit one SM");
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, cudaFuncSetAttribute(jc_lens_mma_kernel<MT, NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(JC_NLENS_COLS / LM_NODES, (chunk + LM_COSMO - 1) / LM_COSMO);
  jc_lens_mma_kernel<MT, NX><<<grid, 512, smem, st>>>(pl, ws, chunk, s0);
}

template <class T>
__global__ void __launch_bounds__(512) jc_tracer_finish_kernel(JcDevPlan pl, Ws ws, int skip_static) {
  if (skip_static && pass_is_static<T>(ws, 0)) return;
  // One CTA per cosmology; blockDim = rows x n_fin with every thread bound to ONE of the n_fin tracers that need all
  // nodes (number counts, delta planes, IA-enabled sources), so the per-tracer switches are loop invariant and there is no
  // index division: ncu had 193 warp instructions per element (issue bound, 15 % FP64 pipe) with a flat (node, tracer)
  // index.  Plain extended sources are complete after the lens kernel's epilogue except node 512 (a = 1: chi = 0 -> 0).
  const int c = blockIdx.x;
  const ptrdiff_t doff = ws.doff;
  for (int t2 = threadIdx.x; t2 < pl.T; t2 += blockDim.x)
    if (pl.tr_kind[t2] == JC_TRACER_WEAK_LENSING && pl.tr_delta_ix[t2] < 0 && !pl.tr_ia[t2])
      JxMem<T>::st(ws.rker + ((size_t)c * JC_NA_PAD + JC_NLENS_COLS) * pl.TS + t2, doff, T(0.0));
  const int nf = pl.n_fin;
  if (nf > 0 && (int)threadIdx.x < (int)(blockDim.x / nf) * nf) {
  const T Om = JxMem<T>::ld(ws.scal + (size_t)c * JC_SCAL_FIELDS + JC_SCAL_OMEGA_M, doff);
  const int t = pl.fin_idx[threadIdx.x % nf], rows = blockDim.x / nf;
  const bool is_wl = pl.tr_kind[t] == JC_TRACER_WEAK_LENSING, inv_growth = pl.tr_inv_growth[t], ia = pl.tr_ia[t];
  const int dix = pl.tr_delta_ix[t];
  const double m1 = pl.tr_m1[t];
  const double* Hn = node_ptr(ws, c, JC_NODE_HUBBLE);
  const double* Dn = node_ptr(ws, c, JC_NODE_GROWTH);
  const double* Cn = node_ptr(ws, c, JC_NODE_CHI);
  T chis = T(0.0), inv_chis = T(0.0);
  if (is_wl && dix >= 0) {  // delta_nz source plane (probes.py:53-64): clip(chi_s - chi, 0) / clip(chi_s, 1)
    const double* ct = ws.chitab + (size_t)c * JC_NCHI;
    const T f0 = JxMem<T>::ld(ct + (dix & 255), doff), f1 = JxMem<T>::ld(ct + (dix >> 8), doff);
    chis = jx_clip0(f0 + (f1 - f0) * pl.tr_delta_t[t]);
    inv_chis = 1.0 / jx_floor1(chis);
  }
  const T wl_amp = (3.0 * JC_H0 * JC_H0 / 2.0 / JC_C_LIGHT) * Om * m1;
  const bool staged = is_wl && dix < 0;  // IA source: lensing term in place for nodes 0..511, NLA term added here
  for (int n = threadIdx.x / nf; n < JC_NA; n += rows) {
    double* out = ws.rker + ((size_t)c * JC_NA_PAD + n) * pl.TS + t;
    const T H = JxMem<T>::ld(Hn + n, doff);
    T r;
    if (is_wl) {
      if (staged) {  // node 512: chi = 0
        r = n < JC_NLENS_COLS ? JxMem<T>::ld(out, doff) : T(0.0);
      } else {
        const T chi = JxMem<T>::ld(Cn + n, doff);
        r = jx_clip0(chis - chi) * inv_chis * (1.0 + pl.limb_z[n]) * chi * wl_amp;
      }
      if (ia) {  // probes.py:119-123
        const T D = JxMem<T>::ld(Dn + n, doff);
        T b = T(pl.bias_node[(size_t)n * pl.TS + t]);
        if (inv_growth) b = b / D;  // bias.py:37-39
        r = r + pl.nz_node[(size_t)n * pl.TS + t] * b * H * (-(JC_C1_RHOCRIT)*Om / D) * m1;
      }
    } else {
      T b = T(pl.bias_node[(size_t)n * pl.TS + t]);
      if (inv_growth) b = b / JxMem<T>::ld(Dn + n, doff);  // bias.py:37-39
      r = pl.nz_node[(size_t)n * pl.TS + t] * b * H;
    }
    JxMem<T>::st(out, doff, r);
  }
  }
  // pad rows 513..519: the TMA contraction reads R in whole 12-row blocks (up to row 515) and relies on zeros
  for (int idx = threadIdx.x; idx < (JC_NA_PAD - JC_NA) * pl.TS; idx += blockDim.x)
    JxMem<T>::st(ws.rker + ((size_t)c * JC_NA_PAD + JC_NA) * pl.TS + idx, doff, T(0.0));
}

template <class T, int NS, int NCOS, int LENS_CGROUPS>
void launch_lens(const JcDevPlan& pl, const Ws& ws, int chunk, int s0, cudaStream_t st, int skip_static = 0) {
  constexpr int CTA_COSMO = NCOS * LENS_CGROUPS;
  dim3 grid(JC_NLENS_COLS / LENS_NODES, (chunk + CTA_COSMO - 1) / CTA_COSMO);
  constexpr int smem = LENS_STAGES * LENS_MR * (LENS_NODES * 8 * (1 + NS) + LENS_NODES * 2);
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, cudaFuncSetAttribute(jc_lens_kernel<T, NS, NCOS, LENS_CGROUPS>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  jc_lens_kernel<T, NS, NCOS, LENS_CGROUPS><<<grid, LENS_NODES * LENS_CGROUPS, smem, st>>>(pl, ws, chunk, s0, skip_static);
}

template <class T, int NCOS, int CG>
int launch_all_lens(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s, int first_src = 0, int skip_static = 0) {
  int n_launch = 0;
  for (int s0 = first_src; s0 < pl.n_src; ++n_launch) {
    const int rem = pl.n_src - s0;
    constexpr bool WIDE = sizeof(T) > 4 * sizeof(double);  // DualN<4>: 8 or 10 sources per launch would spill the accumulators
    if (rem >= 10 && !WIDE) { launch_lens<T, 10, NCOS, CG>(pl, ws, chunk, s0, s, skip_static); s0 += 10; }
    else if (rem >= 8 && !WIDE) { launch_lens<T, 8, NCOS, CG>(pl, ws, chunk, s0, s, skip_static); s0 += 8; }
    else if (rem >= 6) { launch_lens<T, 6, NCOS, CG>(pl, ws, chunk, s0, s, skip_static); s0 += 6; }
    else if (rem >= 5) { launch_lens<T, 5, NCOS, CG>(pl, ws, chunk, s0, s, skip_static); s0 += 5; }
    else if (rem == 4) { launch_lens<T, 4, NCOS, CG>(pl, ws, chunk, s0, s, skip_static); s0 += 4; }
    else if (rem == 3) { launch_lens<T, 3, NCOS, CG>(pl, ws, chunk, s0, s, skip_static); s0 += 3; }
    else if (rem == 2) { launch_lens<T, 2, NCOS, CG>(pl, ws, chunk, s0, s, skip_static); s0 += 2; }
    else { launch_lens<T, 1, NCOS, CG>(pl, ws, chunk, s0, s, skip_static); s0 += 1; }
  }
  return n_launch;
}

}  // namespace

// Scalar kernel: 4 cosmologies per thread, 4 cosmology groups per CTA (2 x 8 and 2 x 4 measured slower, profiles/r01_tuning.md).
// jc_set_option("lens_mma", 1): launches of >= 8 sources run on jc_lens_mma_kernel (10 / 9 / 8 sources per launch: one m-tile plus
// up to two scalar sources; the stage ring of 16 sources would not fit beside the chi tables).
int jc_launch_tracers(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  if (!g_jc_lens_mma) return launch_all_lens<double, 4, 4>(pl, ws, chunk, s);
  int n_launch = 0, s0 = 0;
  for (; pl.n_src - s0 >= 8; ++n_launch) {
    const int rem = pl.n_src - s0;
    if (rem >= 10) { launch_lens_mma<1, 2>(pl, ws, chunk, s0, s); s0 += 10; }
    else if (rem == 9) { launch_lens_mma<1, 1>(pl, ws, chunk, s0, s); s0 += 9; }
    else { launch_lens_mma<1, 0>(pl, ws, chunk, s0, s); s0 += 8; }
  }
  return n_launch + launch_all_lens<double, 4, 4>(pl, ws, chunk, s, s0);
}
// tangent groups: one cosmology per thread (the accumulators are NS x (1 + ntan) doubles)
int jc_launch_tracers_jvp(const JcDevPlan& pl, const Ws& ws, int chunk, int ntan, cudaStream_t s, int skip_static) {
  switch (ntan) {
    case 2: return launch_all_lens<DualN<2>, 1, 4>(pl, ws, chunk, s, 0, skip_static);
    case 3: return launch_all_lens<DualN<3>, 1, 4>(pl, ws, chunk, s, 0, skip_static);
    case 4: return launch_all_lens<DualN<4>, 1, 4>(pl, ws, chunk, s, 0, skip_static);
    default: return launch_all_lens<Dual, 2, 4>(pl, ws, chunk, s, 0, skip_static);
  }
}
static int finish_threads(const JcDevPlan& pl) {  // rows x n_fin threads, at least one thread per tracer
  const int nf = pl.n_fin > 0 ? pl.n_fin : 1;
  const int n = (512 / nf) * nf;
  return n < pl.T ? ((pl.T + 31) / 32) * 32 : n;
}
void jc_launch_finish(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  jc_tracer_finish_kernel<double><<<chunk, finish_threads(pl), 0, s>>>(pl, ws, 0);
}
void jc_launch_finish_jvp(const JcDevPlan& pl, const Ws& ws, int chunk, int ntan, cudaStream_t s, int skip_static) {
  switch (ntan) {
    case 2: jc_tracer_finish_kernel<DualN<2>><<<chunk, finish_threads(pl), 0, s>>>(pl, ws, skip_static); break;
    case 3: jc_tracer_finish_kernel<DualN<3>><<<chunk, finish_threads(pl), 0, s>>>(pl, ws, skip_static); break;
    case 4: jc_tracer_finish_kernel<DualN<4>><<<chunk, finish_threads(pl), 0, s>>>(pl, ws, skip_static); break;
    default: jc_tracer_finish_kernel<Dual><<<chunk, finish_threads(pl), 0, s>>>(pl, ws, skip_static); break;
  }
}
