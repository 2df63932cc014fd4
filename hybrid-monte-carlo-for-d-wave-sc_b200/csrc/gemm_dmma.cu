// gemm_dmma.cu -- batched complex GEMM on the FP64 tensor cores (DMMA m8n8k4).
//   C = beta*C + alpha * sum_seg opA(A_seg) * opB(B_seg)       (column-major, interleaved complex)
// Used for the dense contractions of the eigensolver: the rank-2k trailing update of the
// tridiagonalisation (her2k as one two-segment GEMM, lower tiles only) and the block-reflector
// products of the eigenvector back-transformation.  One complex 8x8x4 product = 4 DMMA.
#include <cstdlib>

#include "dwhmc.h"
#include "gemm_dmma.cuh"
#include "internal.h"

using namespace dwg;

namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

#ifndef GEMM_CN_STAGES
#define GEMM_CN_STAGES 2
#endif

template <int BM, int BN, int OPA, int OPB>
struct Tile {
  static constexpr int A_ELEMS = (OPA == 0) ? BK * (BM + 2) : BM * (BK + 4);
  static constexpr int B_ELEMS = (OPB == 0) ? BN * (BK + 4) : BK * (BN + 2);
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  // three cp.async stages, except for the 64x64 (C, N) tile whose 40 KB stages would leave only one CTA per SM
  static constexpr int STAGES = (OPA == 1 && OPB == 0 && BM == 64 && BN == 64) ? GEMM_CN_STAGES : 3;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * sizeof(cplx);
};

struct KArgs {
  int M, N, K, nseg;
  const cplx* A0; const cplx* A1; const cplx* B0; const cplx* B1;
  int lda, ldb, ldc;
  long long sA, sB, sC;
  cplx* C;
  double alpha, beta;
  int lower;
  int b0;
  Mask mask;
  const int* skip_flag;
  int skip_cols;
  int stairA;
  int ksplit;
  long long sCk;
};

// two CTAs per SM whenever the tile's shared memory allows it (<= 128 registers per thread)
template <int BM, int BN, int WM, int WN, int OPA, int OPB>
__global__ void __launch_bounds__(NTHREADS, (Tile<BM, BN, OPA, OPB>::SMEM <= 113 * 1024) ? 2 : 1)
zgemm_dmma_kernel(KArgs g) {
  using T = Tile<BM, BN, OPA, OPB>;
  constexpr int STAGES = T::STAGES;
  constexpr int WTM = BM / WM, WTN = BN / WN;   // warp tile
  constexpr int MI = WTM / 8, NI = WTN / 8;
  static_assert(WM * WN * 32 == NTHREADS, "8 warps");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* smem = reinterpret_cast<cplx*>(smem_raw);

  const int b = g.b0 + blockIdx.z / g.ksplit;
  if (!g.mask.on(b)) return;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (g.lower && n0 > m0 + BM - 1) return;
  if (g.skip_flag != nullptr && n0 + BN <= g.skip_cols && g.skip_flag[b] != 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp % WM) * WTM, wn0 = (warp / WM) * WTN;

  const cplx* Aseg0 = g.A0 + (size_t)b * g.sA;
  const cplx* Aseg1 = g.A1 ? g.A1 + (size_t)b * g.sA : Aseg0;
  const cplx* Bseg0 = g.B0 + (size_t)b * g.sB;
  const cplx* Bseg1 = g.B1 ? g.B1 + (size_t)b * g.sB : Bseg0;
  cplx* C = g.C + (size_t)b * g.sC;
  if (g.ksplit > 1) {          // this CTA's piece of the K range (single segment only)
    const int ks = blockIdx.z % g.ksplit;
    const int kc = ((g.K + g.ksplit - 1) / g.ksplit + BK - 1) / BK * BK;
    const int kbeg = min(g.K, ks * kc);
    Aseg0 += (OPA == 0) ? (size_t)kbeg * g.lda : (size_t)kbeg;
    Bseg0 += (OPB == 0) ? (size_t)kbeg : (size_t)kbeg * g.ldb;
    g.K = min(kc, g.K - kbeg);
    C += (size_t)ks * g.sCk;
  }

  const int KTS = (g.K + BK - 1) / BK;   // k-tiles per segment
  const int KT = KTS * g.nseg;

  auto load_tile = [&](int kt, int stage) {
    const int seg = kt / KTS;
    const int k0 = (kt - seg * KTS) * BK;
    const cplx* A = seg ? Aseg1 : Aseg0;
    const cplx* Bm = seg ? Bseg1 : Bseg0;
    cplx* As = smem + (size_t)stage * T::STAGE_ELEMS;
    cplx* Bs = As + T::A_ELEMS;
    // ---- A tile
    if (OPA == 0) {   // A is M x K column-major, tile stored [k][m]
#pragma unroll
      for (int i = 0; i < (BM * BK) / NTHREADS; ++i) {
        int idx = tid + i * NTHREADS;
        int m = idx % BM, kk = idx / BM;
        bool p = (m0 + m < g.M) && (k0 + kk < g.K);
        if (g.stairA > 0) { const int df = (m0 + m) - (k0 + kk); p = p && df >= 0 && df < g.stairA; }
        const cplx* src = p ? A + (size_t)(k0 + kk) * g.lda + (m0 + m) : A;
        cp_async16(As + kk * (BM + 2) + m, src, p);
      }
    } else {          // A source is K x M column-major (A_eff = src^H), tile stored [m][k]
#pragma unroll
      for (int i = 0; i < (BM * BK) / NTHREADS; ++i) {
        int idx = tid + i * NTHREADS;
        int kk = idx % BK, m = idx / BK;
        bool p = (m0 + m < g.M) && (k0 + kk < g.K);
        if (g.stairA > 0) { const int df = (k0 + kk) - (m0 + m); p = p && df >= 0 && df < g.stairA; }
        const cplx* src = p ? A + (size_t)(m0 + m) * g.lda + (k0 + kk) : A;
        cp_async16(As + m * (BK + 4) + kk, src, p);
      }
    }
    // ---- B tile
    if (OPB == 0) {   // B is K x N column-major, tile stored [n][k]
#pragma unroll
      for (int i = 0; i < (BN * BK) / NTHREADS; ++i) {
        int idx = tid + i * NTHREADS;
        int kk = idx % BK, nn = idx / BK;
        bool p = (n0 + nn < g.N) && (k0 + kk < g.K);
        const cplx* src = p ? Bm + (size_t)(n0 + nn) * g.ldb + (k0 + kk) : Bm;
        cp_async16(Bs + nn * (BK + 4) + kk, src, p);
      }
    } else {          // B source is N x K column-major (B_eff = src^H), tile stored [k][n]
#pragma unroll
      for (int i = 0; i < (BN * BK) / NTHREADS; ++i) {
        int idx = tid + i * NTHREADS;
        int nn = idx % BN, kk = idx / BN;
        bool p = (n0 + nn < g.N) && (k0 + kk < g.K);
        const cplx* src = p ? Bm + (size_t)(k0 + kk) * g.ldb + (n0 + nn) : Bm;
        cp_async16(Bs + kk * (BN + 2) + nn, src, p);
      }
    }
  };

  const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_tile(s, s);
    cp_async_commit();
  }

  // The accumulators start from (beta / alpha) * C, so the read of C overlaps the pipeline prologue
  // instead of stalling the epilogue (alpha = -1, beta = 1 for the rank-k updates: exact).
  double cr[MI][NI][2], ci[MI][NI][2];
  {
    const double scale = (g.beta != 0.0) ? g.beta / g.alpha : 0.0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int row = m0 + wm0 + i * 8 + fr;
          const int col = n0 + wn0 + j * 8 + 2 * fk + q;
          cplx old = make_double2(0.0, 0.0);
          if (g.beta != 0.0 && row < g.M && col < g.N) old = C[(size_t)col * g.ldc + row];
          cr[i][j][q] = scale * old.x;
          ci[i][j][q] = scale * old.y;
        }
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) load_tile(nk, nk % STAGES);
      cp_async_commit();
    }
    const cplx* As = smem + (size_t)(kt % STAGES) * T::STAGE_ELEMS;
    const cplx* Bs = As + T::A_ELEMS;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double ar[MI], ai[MI], nai[MI], br[NI], bi[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        int m = wm0 + i * 8 + fr, k = k4 * 4 + fk;
        cplx v = (OPA == 0) ? As[k * (BM + 2) + m] : As[m * (BK + 4) + k];
        ar[i] = v.x;
        ai[i] = (OPA == 0) ? v.y : -v.y;
        nai[i] = -ai[i];
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        int nn = wn0 + j * 8 + fr, k = k4 * 4 + fk;
        cplx v = (OPB == 0) ? Bs[nn * (BK + 4) + k] : Bs[k * (BN + 2) + nn];
        br[j] = v.x;
        bi[j] = (OPB == 0) ? v.y : -v.y;
      }
      // four passes of MI*NI independent DMMAs: the two updates of one accumulator are 2*MI*NI
      // instructions apart, so no DMMA waits on the latency of its predecessor
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(cr[i][j][0], cr[i][j][1], nai[i], bi[j]);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(ci[i][j][0], ci[i][j][1], ai[i], br[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        int row = m0 + wm0 + i * 8 + fr;
        int col = n0 + wn0 + j * 8 + 2 * fk + q;
        if (row < g.M && col < g.N) {
          cplx* p = C + (size_t)col * g.ldc + row;
          cplx out;
          out.x = g.alpha * cr[i][j][q];
          out.y = g.alpha * ci[i][j][q];
          *p = out;
        }
      }
}

template <int BM, int BN, int WM, int WN, int OPA, int OPB>
int launch(Handle* h, const ZgemmArgs& a) {
  using T = Tile<BM, BN, OPA, OPB>;
  static bool attr_set[64] = {false};
  auto kern = zgemm_dmma_kernel<BM, BN, WM, WN, OPA, OPB>;
  if (!attr_set[h->device & 63]) {
    DW_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
    attr_set[h->device & 63] = true;
  }
  KArgs g;
  g.M = a.M; g.N = a.N; g.K = a.K; g.nseg = a.nseg;
  g.A0 = a.A[0]; g.A1 = a.nseg > 1 ? a.A[1] : nullptr;
  g.B0 = a.Bm[0]; g.B1 = a.nseg > 1 ? a.Bm[1] : nullptr;
  g.lda = a.lda; g.ldb = a.ldb; g.ldc = a.ldc;
  g.sA = a.sA; g.sB = a.sB; g.sC = a.sC;
  g.C = a.C; g.alpha = a.alpha; g.beta = a.beta; g.lower = a.lower; g.mask = a.mask; g.b0 = a.b0;
  g.skip_flag = a.skip_flag; g.skip_cols = a.skip_cols; g.stairA = a.stairA;
  g.ksplit = a.ksplit > 1 ? a.ksplit : 1; g.sCk = a.sCk;
  dim3 grid((a.M + BM - 1) / BM, (a.N + BN - 1) / BN, a.batch * g.ksplit);
  kern<<<grid, NTHREADS, T::SMEM, a.stream ? a.stream : h->stream>>>(g);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

}  // namespace

int dw_zgemm(Handle* h, const ZgemmArgs& a) {
  if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return DWHMC_OK;
  if (a.K <= 0 && a.beta == 1.0) return DWHMC_OK;
  const bool small_m = a.M <= 32;
  if (a.opA == 0 && a.opB == 1) {
    return launch<64, 64, 2, 4, 0, 1>(h, a);
  } else if (a.opA == 1 && a.opB == 0) {
    if (small_m) return launch<32, 128, 1, 8, 1, 0>(h, a);
    return launch<64, 64, 2, 4, 1, 0>(h, a);
  } else if (a.opA == 0 && a.opB == 0) {
    if (small_m) return launch<32, 128, 1, 8, 0, 0>(h, a);
    return launch<64, 64, 2, 4, 0, 0>(h, a);
  }
  h->err = "dw_zgemm: unsupported op combination";
  return DWHMC_E_BADARG;
}
