// band_apply.cu -- back-transformation of the band route: U = Q2 Z (/root/reference src/Hamiltonian.jl:106-111,
// the eigenvectors eigen! returns).  Q2 is the product of the Householder reflectors of the bulge chase (band.cu);
// the reflectors of g consecutive sweeps at the same step form one staircase block reflector I - Vb T Vb^H.
//   band_tfactor_kernel   T and Vb T per block (runs beside the D&C stage)
//   band_apply2_kernel    the block reflectors on the FP64 tensor cores, Z strips held in registers, three real
//                         products per complex one
//   band_unpermute_kernel rows back to the reference's site order
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dwhmc.h"
#include "gemm_dmma.cuh"
#include "internal.h"

namespace {

__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {       // acc += a b
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfmac(cplx& acc, cplx a, cplx b) {      // acc += conj(a) b
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}

// ---- block reflectors of the back-transformation ---------------------------------------------------
// block (s0, k): columns s0 .. s0+g-1 of V, rows rlo = s0+1+kb .. ; column c is non-zero on rows
// [c, c+b) of the block.  T (forward, columnwise) from the Gram matrix of the masked block.
constexpr int A2_MT = 4;                 // tiles of 8 reflectors: up to 32 reflectors per block
constexpr int A2_G = 8 * A2_MT;
constexpr int A2_NRT = 17;               // tiles of 8 rows: blocks of up to 136 rows (b + g - 1)
constexpr int A2_ROWS = 8 * A2_NRT;
constexpr int A2_WARPS = 8;              // 255 registers per thread: the whole row range of a strip lives in registers
constexpr int A2_TH = 32 * A2_WARPS;
// Complex products are formed from three real ones (a + ib)(c + id): k1 = (a + b) c, k2 = a (d - c), k3 = b (c + d),
// re = k1 - k3, im = k1 + k2; the shared-memory operands are stored as (c, d - c) pairs plus a plane of c + d.
constexpr int A2_LDT = A2_G + 8;         // row length of the transposed c + d plane of -Vb T (= 8 mod 16: conflict-free pairs)
// per block in global memory (band_tfactor_kernel -> band_apply2_kernel), in doubles
constexpr size_t A2_OFF_CVX = 0;                                      // [A2_G][A2_ROWS] (c, d - c) of conj(Vb)
constexpr size_t A2_OFF_NTX = A2_OFF_CVX + 2 * (size_t)A2_G * A2_ROWS;  // [A2_G][A2_ROWS] (c, d - c) of -Vb T
constexpr size_t A2_OFF_CVS = A2_OFF_NTX + 2 * (size_t)A2_G * A2_ROWS;  // [A2_G][A2_ROWS] c + d of conj(Vb)
constexpr size_t A2_OFF_NTS = A2_OFF_CVS + (size_t)A2_G * A2_ROWS;      // [A2_ROWS][A2_LDT] c + d of -Vb T, transposed
constexpr size_t A2_BLOCK_DOUBLES = A2_OFF_NTS + (size_t)A2_ROWS * A2_LDT;
static_assert(A2_BLOCK_DOUBLES == DW_APPLY_BLOCK_DOUBLES, "internal.h");
// shared memory: pair planes with odd leading dimension 8 NRT + 1 (both operand read patterns conflict-free), the
// c + d plane of conj(Vb) with rows of 8 NRT (+ 8 if that is a multiple of 16) doubles, the one of -Vb T as in global memory
__host__ __device__ constexpr int apply2_lds(int nrt) { return (8 * nrt) % 16 == 8 ? 8 * nrt : 8 * nrt + 8; }
constexpr size_t apply2_smem(int nrt) {
  return sizeof(cplx) * 2 * (size_t)A2_G * (8 * nrt + 1) + sizeof(double) * ((size_t)A2_G * apply2_lds(nrt) + (size_t)8 * nrt * A2_LDT);
}

static_assert(A2_G == DW_APPLY_G && A2_ROWS == DW_APPLY_ROWS, "internal.h");

// ---- T factors ------------------------------------------------------------------------------------------
// block (s0, k): columns s0 .. s0+g-1 of V, rows rlo = s0+1+kb .. ; column c is non-zero on rows
// [c, c+b) of the block.  T (forward, columnwise) from the Gram matrix of the masked block; the output is what
// band_apply2_kernel keeps in shared memory, ready for bulk copies: conj(Vb) and -Vb T, each [A2_G reflectors]
// [A2_ROWS rows] per block, zero padded.
__global__ void __launch_bounds__(256) band_tfactor_kernel(const cplx* __restrict__ Vall, const cplx* __restrict__ tau2,
                                                           cplx* __restrict__ NVTall, const int* __restrict__ blk_s0,
                                                           const int* __restrict__ blk_k, int n, int b, int g, int KT,
                                                           int nblk, Mask mask) {
  constexpr int TG = A2_G;
  const int blk = blockIdx.x, ch = blockIdx.y;
  if (!mask.on(ch)) return;
  extern __shared__ __align__(16) unsigned char smem_tf[];
  cplx* Vs = reinterpret_cast<cplx*>(smem_tf);   // [32][TG + 1] row chunk [32 rows][g columns], column index fastest
  cplx* G = Vs + 32 * (TG + 1);                   // [g][g] column-major
  cplx* T = G + TG * TG;                          // [g][g+1] row-major rows
  const int s0 = blk_s0[blk], k = blk_k[blk];
  const int gg = min(g, n - 1 - s0);                     // sweeps s0 .. s0+gg-1 exist
  const int rlo = s0 + 1 + k * b;
  const int rows = min(n - rlo, b + gg - 1);
  const cplx* V = Vall + (size_t)ch * n * n;
  const int tid = threadIdx.x;
  const cplx zero = make_double2(0.0, 0.0);
  auto load_chunk = [&](int rc) {
    for (int idx = tid; idx < 32 * gg; idx += 256) {
      const int r = idx & 31, c = idx >> 5;
      const int rr = rc + r;
      const bool ok = rr < rows && rr - c >= 0 && rr - c < b;
      Vs[r * (TG + 1) + c] = ok ? V[(size_t)(s0 + c) * n + rlo + rr] : zero;
    }
  };
  // Gram matrix, accumulated over row chunks; only its upper triangle G[c1, c2] = v_c1^H v_c2, c1 < c2, is read below.
  // Thread (ti, tj) = (tid mod 16, tid / 16) owns the pair (ti, tj + 16) and, if ti <= tj, the pairs (ti, tj) and
  // (ti + 16, tj + 16): four loads (two of consecutive lanes, two broadcast) for up to three products per row
  // (g = 32 = TG: dw_band_setup).
  const int ti = tid & 15, tj = tid >> 4;
  const bool upper = ti <= tj;
  cplx gA = zero, gB = zero, gC = zero;
  for (int rc = 0; rc < rows; rc += 32) {
    __syncthreads();
    load_chunk(rc);
    __syncthreads();
    for (int r = 0; r < 32; ++r) {
      const cplx a0 = Vs[r * (TG + 1) + ti], b1 = Vs[r * (TG + 1) + tj + 16];
      cfmac(gA, a0, b1);
      if (upper) {
        cfmac(gB, a0, Vs[r * (TG + 1) + tj]);
        cfmac(gC, Vs[r * (TG + 1) + ti + 16], b1);
      }
    }
  }
  for (int idx = tid; idx < g * g; idx += 256) G[idx] = zero;
  __syncthreads();
  {
    const bool okA = ti < gg && tj + 16 < gg, okB = upper && tj < gg, okC = upper && tj + 16 < gg;
    if (okA) G[(tj + 16) * g + ti] = gA;
    if (okB) G[tj * g + ti] = gB;
    if (okC) G[(tj + 16) * g + ti + 16] = gC;
  }
  for (int idx = tid; idx < g * (g + 1); idx += 256) T[idx] = zero;
  __syncthreads();
  // T[i,i] = tau_i ; T[0:i, i] = -tau_i T[0:i,0:i] G[0:i, i] ; eight lanes share row r (the rows are independent of
  // each other, so a warp barrier per column is all the synchronisation the recurrence needs)
  {
    const int r = tid >> 3, part = tid & 7;
    cplx* Tr = T + r * (g + 1);
    for (int i = 0; i < gg; ++i) {
      const cplx t = tau2[((size_t)ch * n + s0 + i) * KT + k];
      cplx s = zero;
      if (r < i)
        for (int l = r + part; l < i; l += 8) cfma(s, Tr[l], G[i * g + l]);
      s.x += __shfl_xor_sync(0xffffffffu, s.x, 1); s.y += __shfl_xor_sync(0xffffffffu, s.y, 1);
      s.x += __shfl_xor_sync(0xffffffffu, s.x, 2); s.y += __shfl_xor_sync(0xffffffffu, s.y, 2);
      s.x += __shfl_xor_sync(0xffffffffu, s.x, 4); s.y += __shfl_xor_sync(0xffffffffu, s.y, 4);
      if (part == 0) {
        if (r < i) Tr[i] = make_double2(-(t.x * s.x - t.y * s.y), -(t.x * s.y + t.y * s.x));
        else if (r == i) Tr[i] = t;
      }
      __syncwarp();
    }
  }
  // conj(Vb) and -Vb T as (c, d - c) pairs and c + d planes
  double* ob = reinterpret_cast<double*>(NVTall) + ((size_t)ch * nblk + blk) * A2_BLOCK_DOUBLES;
  cplx* cvx = reinterpret_cast<cplx*>(ob + A2_OFF_CVX);
  cplx* ntx = reinterpret_cast<cplx*>(ob + A2_OFF_NTX);
  double* cvs = ob + A2_OFF_CVS;
  double* nts = ob + A2_OFF_NTS;
  for (int rc = 0; rc < A2_ROWS; rc += 32) {
    __syncthreads();
    load_chunk(rc);
    if (tid < TG) for (int c = gg; c < TG; ++c) Vs[tid * (TG + 1) + c] = zero;     // columns of a short last group
    __syncthreads();
    double* stage = reinterpret_cast<double*>(G);      // [32 rows][33]: the transposed plane goes out row by row (G is dead)
    for (int idx = tid; idx < 32 * A2_G; idx += 256) {
      const int r = idx & 31, m = idx >> 5;
      if (rc + r >= A2_ROWS) continue;
      cplx a = zero;
      if (m < gg && rc + r < rows)
        for (int j = 0; j <= m; ++j) cfma(a, Vs[r * (TG + 1) + j], T[j * (g + 1) + m]);     // T upper triangular
      ntx[m * A2_ROWS + rc + r] = make_double2(-a.x, a.x - a.y);
      stage[r * 33 + m] = -a.x - a.y;
      const cplx v = Vs[r * (TG + 1) + m];
      cvx[m * A2_ROWS + rc + r] = make_double2(v.x, -v.y - v.x);
      cvs[m * A2_ROWS + rc + r] = v.x - v.y;
    }
    __syncthreads();
    for (int idx = tid; idx < 32 * A2_G; idx += 256) {
      const int m = idx & 31, r = idx >> 5;
      if (rc + r < A2_ROWS) nts[(rc + r) * A2_LDT + m] = stage[r * 33 + m];
    }
  }
}

// ---- staircase block reflectors on the FP64 tensor cores: Z[R, :] -= (Vb T) (Vb^H Z[R, :]) -------------------
// Work item = (block, column part, chain).  conj(Vb) and -Vb T (<= 136 rows x 32 reflectors, zero outside the
// staircase) sit in shared memory; every warp owns strips of 8 columns of Z and keeps the whole row range of
// a strip in registers, transposed: the DMMA m8n8k4 tile Z^T[8 columns][8 rows] is at once
//   * the A operand of the first product  W^T (8 columns x 32 reflectors) = Z^T conj(Vb)   (k runs over rows), and
//   * the accumulator of the second one   Z^T -= W^T (Vb T)^T                               (k runs over reflectors),
// because the order of the k index inside a dot product is free: lane (column, c) holds rows 2c and 2c+1 of a
// row tile as accumulator elements and uses exactly these two rows as its k entries of two k-steps (the shared-
// memory operand is read with the same permutation); likewise W^T comes out of the first product in accumulator
// layout and goes into the second as the A operand unchanged.  So Z is read from global memory once, updated in
// registers and written once, there is no shared-memory staging of Z, no transposition and no block barrier
// inside an item; four real DMMAs per complex product, k-steps outside the staircase are skipped.
// Items are handed out by a ticket counter in wavefront order (t = (Gmax - G) + k: blocks of one wavefront touch
// disjoint rows and only depend on smaller t); an item waits until the items of the previous wavefront of its
// (chain, column part) have been published, so one launch runs the whole back-transformation.
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ cplx ldcg(const cplx* p) {   // L2 only: Z is shared between CTAs across wavefronts
  cplx v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Apply2Args {
  cplx* Z; const cplx* V; const cplx* NVT;
  const int* blk_s0; const int* blk_k;
  const int4* items;             // wavefront order, chains one after the other inside a wavefront: (block, column part |
                                 // parts << 16, wavefront | chain << 16, items of the chain in the previous wavefront)
  int nitems;
  int* ticket;                   // [1]
  int* done;                     // [B][nwave] published items
  int* status;                   // [>= 3]: [2] set if a wait timed out
  const int* halfflag;
  int n, b, g, nblk, B, nwave, c_lo, use_half;
  Mask mask;
};

// NRT: row tiles of a block, 8 NRT >= b + 31.  Reflector tile mt (reflectors 8 mt .. 8 mt + 7) is non-zero on rows
// 8 mt .. 8 mt + 6 + b, columns 8 mt .. of -Vb T on rows 0 .. 8 mt + 6 + b: row tiles rt with rt - mt > SK = NRT - 4
// (>= (b + 6) / 8) or rt < mt are skipped at compile time; partial blocks at the matrix end are zero padded.
template <int NRT>
__global__ void __launch_bounds__(A2_TH, 1) band_apply2_kernel(Apply2Args a) {
  constexpr int ROWS = 8 * NRT, LDV = ROWS + 1, LDS = apply2_lds(NRT), SK = NRT - 4;
  extern __shared__ __align__(16) unsigned char smem_apply[];
  cplx* CV = reinterpret_cast<cplx*>(smem_apply);    // [A2_G][LDV]  (c, d - c) of conj(Vb)[r][m] at m * LDV + r
  cplx* NVT = CV + A2_G * LDV;                        // [A2_G][LDV]  (c, d - c) of -(Vb T)[r][m]
  double* CVS = reinterpret_cast<double*>(NVT + A2_G * LDV);   // [A2_G][LDS]    c + d of conj(Vb)[r][m] at m * LDS + r
  double* NTS = CVS + A2_G * LDS;                               // [ROWS][A2_LDT] c + d of -(Vb T)[r][m] at r * A2_LDT + m
  __shared__ int s_q;
  __shared__ unsigned long long s_bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int n = a.n, b = a.b;
  const cplx zero = make_double2(0.0, 0.0);
  if (tid == 0) mbar_init(&s_bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  unsigned phase = 0;
  for (;;) {
    __syncthreads();                                  // everyone is done with shared memory and s_q
    if (tid == 0) s_q = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int q = s_q;
    if (q >= a.nitems) return;
    const int4 it = a.items[q];
    const int blk = it.x, part = it.y & 0xffff, nparts = it.y >> 16, wave = it.z & 0xffff, chain = it.z >> 16;
    if (!a.mask.on(chain)) continue;
    const int s0 = a.blk_s0[blk], k = a.blk_k[blk];
    const int gg = min(a.g, n - 1 - s0);
    const int rlo = s0 + 1 + k * b;
    const int rows = min(n - rlo, b + gg - 1);
    int* done = a.done + (size_t)chain * a.nwave;
    const int cstart = (a.use_half && a.halfflag[chain] != 0) ? a.c_lo : 0;
    const int nstrip = (n - cstart + 7) >> 3;
    const int per = (nstrip + nparts - 1) / nparts;
    const int st0 = part * per, st1 = min(nstrip, st0 + per);
    if (rows > 0 && gg > 0 && st0 < st1) {
      if (tid == 0) {
        // operands by the bulk-copy engine, one copy per reflector (the rows of shared memory are padded)
        const double* src = reinterpret_cast<const double*>(a.NVT) + ((size_t)chain * a.nblk + blk) * A2_BLOCK_DOUBLES;
        const cplx* scv = reinterpret_cast<const cplx*>(src + A2_OFF_CVX);
        const cplx* snt = reinterpret_cast<const cplx*>(src + A2_OFF_NTX);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // reads of the previous item -> bulk writes
        mbar_expect_tx(&s_bar, (unsigned)(2 * A2_G * ROWS * sizeof(cplx) + (A2_G * ROWS + ROWS * A2_LDT) * sizeof(double)));
        for (int m = 0; m < A2_G; ++m) {
          bulk_g2s(CV + m * LDV, scv + m * A2_ROWS, (unsigned)(ROWS * sizeof(cplx)), &s_bar);
          bulk_g2s(NVT + m * LDV, snt + m * A2_ROWS, (unsigned)(ROWS * sizeof(cplx)), &s_bar);
          bulk_g2s(CVS + m * LDS, src + A2_OFF_CVS + m * A2_ROWS, (unsigned)(ROWS * sizeof(double)), &s_bar);
        }
        bulk_g2s(NTS, src + A2_OFF_NTS, (unsigned)(ROWS * A2_LDT * sizeof(double)), &s_bar);
        // the blocks of the previous wavefront of this chain have to be in global memory
        if (wave > 0 && it.w > 0) {
          int spins = 0;
          while (ld_acquire(done + wave - 1) < it.w) {
            __nanosleep(64);
            if (++spins > (1 << 22)) { atomicExch(a.status + 2, 1); break; }
          }
        }
      }
      __syncthreads();
      mbar_wait(&s_bar, phase);
      phase ^= 1;
      // Strips of this warp.  The row tiles of the next strip are loaded into the registers of the current one as
      // soon as their final values have been stored (inside the second product), so the loads and stores of a strip
      // are spread over the tensor work instead of forming a burst before and after it.
      double zr[NRT][2], zi[NRT][2];
      int strip = st0 + warp;
      auto strip_ptr = [&](int st, bool& cok) -> cplx* {
        const int col = cstart + st * 8 + fr;
        cok = col < n;
        return a.Z + ((size_t)chain * n + (cok ? col : 0)) * n + rlo + 2 * fk;
      };
      auto prefetch_strip = [&](int st) {              // -> L2 (lane = 128-byte lines of a column)
        const int col = cstart + st * 8 + fr;
        if (st < st1 && col < n) {
          const char* pz = reinterpret_cast<const char*>(a.Z + ((size_t)chain * n + col) * n + rlo);
          for (int off = fk * 128; off < rows * (int)sizeof(cplx); off += 4 * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pz + off));
        }
      };
      bool colok = false;
      cplx* zc = nullptr;
      if (strip < st1) {
        zc = strip_ptr(strip, colok);
        prefetch_strip(strip + A2_WARPS);
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const bool ok = colok && 8 * rt + 2 * fk + e < rows;
            const cplx v = ok ? ldcg(zc + 8 * rt + e) : zero;
            zr[rt][e] = v.x; zi[rt][e] = v.y;
          }
        }
      }
      while (strip < st1) {
        const int nxt = strip + A2_WARPS;
        const bool more = nxt < st1;
        bool colok_n = false;
        cplx* zn = more ? strip_ptr(nxt, colok_n) : zc;
        prefetch_strip(nxt + A2_WARPS);
        // ---- W^T = Z^T conj(Vb): accumulators (column fr, reflectors 8 mt + 2 fk + {0, 1}) of the three real products
        double k1[A2_MT][2], k2[A2_MT][2], k3[A2_MT][2];
#pragma unroll
        for (int mt = 0; mt < A2_MT; ++mt) { k1[mt][0] = k1[mt][1] = k2[mt][0] = k2[mt][1] = k3[mt][0] = k3[mt][1] = 0.0; }
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt) {
          const double as0 = zr[rt][0] + zi[rt][0], as1 = zr[rt][1] + zi[rt][1];
#pragma unroll
          for (int mt = 0; mt < A2_MT; ++mt) {
            if (rt < mt || rt - mt > SK) continue;
            const cplx x0 = CV[(8 * mt + fr) * LDV + 8 * rt + 2 * fk];
            const cplx x1 = CV[(8 * mt + fr) * LDV + 8 * rt + 2 * fk + 1];
            const double2 sp = *reinterpret_cast<const double2*>(CVS + (8 * mt + fr) * LDS + 8 * rt + 2 * fk);
            dmma(k1[mt][0], k1[mt][1], as0, x0.x);
            dmma(k2[mt][0], k2[mt][1], zr[rt][0], x0.y);
            dmma(k3[mt][0], k3[mt][1], zi[rt][0], sp.x);
            dmma(k1[mt][0], k1[mt][1], as1, x1.x);
            dmma(k2[mt][0], k2[mt][1], zr[rt][1], x1.y);
            dmma(k3[mt][0], k3[mt][1], zi[rt][1], sp.y);
          }
        }
        // W^T = (k1 - k3) + i (k1 + k2); operands of the second product: real part, imaginary part and their sum
#pragma unroll
        for (int mt = 0; mt < A2_MT; ++mt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double re = k1[mt][e] - k3[mt][e], im = k1[mt][e] + k2[mt][e];
            k1[mt][e] = re + im; k2[mt][e] = re; k3[mt][e] = im;
          }
        // ---- Z^T += W^T (-Vb T)^T: columns of -Vb T 8 mt .. 8 mt + 7 are non-zero on rows 0 .. 8 mt + 6 + b
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt) {
          double p1[2] = {0.0, 0.0}, p2[2] = {0.0, 0.0}, p3[2] = {0.0, 0.0};
#pragma unroll
          for (int mt = 0; mt < A2_MT; ++mt) {
            if (rt - mt > SK) continue;
            const cplx x0 = NVT[(8 * mt + 2 * fk) * LDV + 8 * rt + fr];
            const cplx x1 = NVT[(8 * mt + 2 * fk + 1) * LDV + 8 * rt + fr];
            const double2 sp = *reinterpret_cast<const double2*>(NTS + (8 * rt + fr) * A2_LDT + 8 * mt + 2 * fk);
            dmma(p1[0], p1[1], k1[mt][0], x0.x);
            dmma(p2[0], p2[1], k2[mt][0], x0.y);
            dmma(p3[0], p3[1], k3[mt][0], sp.x);
            dmma(p1[0], p1[1], k1[mt][1], x1.x);
            dmma(p2[0], p2[1], k2[mt][1], x1.y);
            dmma(p3[0], p3[1], k3[mt][1], sp.y);
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const bool rok = 8 * rt + 2 * fk + e < rows;
            if (colok && rok) zc[8 * rt + e] = make_double2(zr[rt][e] + (p1[e] - p3[e]), zi[rt][e] + (p1[e] + p2[e]));
            const cplx v = (more && colok_n && rok) ? ldcg(zn + 8 * rt + e) : zero;     // same tile of the next strip
            zr[rt][e] = v.x; zi[rt][e] = v.y;
          }
        }
        strip = nxt; zc = zn; colok = colok_n;
      }
    }
    // publish: every thread's stores are ordered before the counter update
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(done + wave, 1);
  }
}

// rows back to the reference's order: U[r, c] = Zb[pos[r], c]
__global__ void __launch_bounds__(256) band_unpermute_kernel(const cplx* __restrict__ Zall, cplx* __restrict__ Uall,
                                                             const int* __restrict__ pos, const int* __restrict__ halfflag,
                                                             int c_lo, int n, Mask mask) {
  const int b = blockIdx.y, c = blockIdx.x;
  if (!mask.on(b)) return;
  if (c < c_lo && halfflag[b] != 0) return;
  const cplx* src = Zall + (size_t)b * n * n + (size_t)c * n;
  cplx* dst = Uall + (size_t)b * n * n + (size_t)c * n;
  for (int r = threadIdx.x; r < n; r += blockDim.x) dst[r] = src[pos[r]];
}

}  // namespace

// -Vb T of all staircase blocks (needs only the chase output, so it can run beside the D&C stage)
int dw_band_tfactors(Handle* h, Mask mask, cudaStream_t stream) {
  const int n = h->n, B = h->B, bw = h->band_b, g = h->band_g;
  const int nblk = (int)h->band_blk_s0.size();
  dim3 grid(nblk, B);
  constexpr size_t smem = sizeof(cplx) * (32 * (A2_G + 1) + A2_G * A2_G + A2_G * (A2_G + 1));
  static bool attr_set[64] = {false};
  if (!attr_set[h->device & 63]) {
    DW_CUDA(h, cudaFuncSetAttribute(band_tfactor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[h->device & 63] = true;
  }
  band_tfactor_kernel<<<grid, 256, smem, stream>>>(h->V, h->band_tau, h->band_VT, h->band_blk_s0_dev, h->band_blk_k_dev, n, bw,
                                                g, h->band_KT, nblk, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

// item list of the back-transformation: wavefront order, inside a wavefront the chains one after the other (the items
// a chain's next wavefront waits for are then the ones handed out longest ago); called once from dwhmc_create after
// dw_band_setup.  Column parts (strips of 8 columns, at least one per warp): one part per block wherever the blocks of a
// wavefront fill the GPU on their own, more parts in the thin wavefronts at both ends.
void dw_band_apply_items(Handle* h, std::vector<int>& items4) {
  const int n = h->n;
  const int nwave = (int)h->band_wave_start.size() - 1;
  const int nstrip = (n - (h->N / 16) * 16 + 7) / 8;
  const int maxparts = std::max(1, (nstrip + A2_WARPS - 1) / A2_WARPS);
  items4.clear();
  int prev = 0;
  for (int t = 0; t < nwave; ++t) {
    const int w0 = h->band_wave_start[t], w1 = h->band_wave_start[t + 1];
    const int nsub = w1 - w0;
    if (nsub <= 0) continue;
    int np = (13 * h->nsm + 10 * nsub * h->B - 1) / (10 * nsub * h->B);
    np = std::max(1, std::min(np, maxparts));
    for (int c = 0; c < h->B; ++c)
      for (int i = w0; i < w1; ++i)
        for (int p = 0; p < np; ++p) {
          items4.push_back(h->band_wave_blk[i]); items4.push_back(p | (np << 16)); items4.push_back(t | (c << 16));
          items4.push_back(prev);
        }
    prev = nsub * np;
  }
  h->band_nitems = (int)items4.size() / 4;
}

// Zb (n x n complex, band row order, in h->A) <- Q2 Zb, then rows back to site order into U
int dw_band_backtransform(Handle* h, cplx* U, Mask mask, bool ph) {
  const int n = h->n, B = h->B;
  const bool half = ph && h->ph_mode;
  cplx* Z = h->A;
  // kernel instance by the number of row tiles of a block, 8 NRT >= b + 31
  const int nrt = std::max(8, (h->band_b + A2_G - 1 + 7) / 8);
  void (*kern)(Apply2Args) = nullptr;
  switch (nrt) {
    case 8: kern = band_apply2_kernel<8>; break;    case 9: kern = band_apply2_kernel<9>; break;
    case 10: kern = band_apply2_kernel<10>; break;  case 11: kern = band_apply2_kernel<11>; break;
    case 12: kern = band_apply2_kernel<12>; break;  case 13: kern = band_apply2_kernel<13>; break;
    case 14: kern = band_apply2_kernel<14>; break;  case 15: kern = band_apply2_kernel<15>; break;
    case 16: kern = band_apply2_kernel<16>; break;  case 17: kern = band_apply2_kernel<17>; break;
    default: h->err = "dw_band_backtransform: half-bandwidth not supported"; return DWHMC_E_BADARG;
  }
  const size_t smem = apply2_smem(nrt);
  if (h->band_apply_attr != nrt) {
    DW_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->band_apply_attr = nrt;
  }
  const int nwave = (int)h->band_wave_start.size() - 1;
  const int c_lo = half ? (h->N / 16) * 16 : 0;
  Apply2Args a;
  a.Z = Z; a.V = h->V; a.NVT = h->band_VT; a.blk_s0 = h->band_blk_s0_dev; a.blk_k = h->band_blk_k_dev;
  a.items = reinterpret_cast<const int4*>(h->band_items_dev);
  a.ticket = h->band_sync; a.done = h->band_sync + 1; a.status = h->status; a.halfflag = h->halfflag;
  a.n = n; a.b = h->band_b; a.g = h->band_g; a.nblk = (int)h->band_blk_s0.size(); a.B = B;
  a.nwave = nwave; a.c_lo = c_lo; a.use_half = half ? 1 : 0; a.mask = mask;
  const size_t nsync = 1 + (size_t)B * nwave;
  DW_CUDA(h, cudaMemsetAsync(h->band_sync, 0, sizeof(int) * nsync, h->stream));
  a.nitems = h->band_nitems;
  const int ctas = std::min(h->nsm, std::max(1, a.nitems));
  kern<<<ctas, A2_TH, smem, h->stream>>>(a);
  DW_LAUNCH_CHECK(h);
  {
    dim3 grid(n, B);
    band_unpermute_kernel<<<grid, 256, 0, h->stream>>>(Z, U, h->band_pos, h->halfflag, c_lo, n, mask);
    DW_LAUNCH_CHECK(h);
  }
  return DWHMC_OK;
}
