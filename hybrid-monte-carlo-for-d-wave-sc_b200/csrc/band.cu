// band.cu -- band route of diagonalize_H_BdG! (/root/reference src/Hamiltonian.jl:96-114).
//
// The BdG matrix of a periodic Lx x Ly lattice is sparse (13 entries per row).  With the sites of
// each ring folded (0, L-1, 1, L-2, ...) and particle / hole components interleaved it is a band
// matrix of half-bandwidth b = 4 min(Lx, Ly) + 4 (100 of n = 1152 at L = 24), so the dense -> band
// stage of a two-stage eigensolver is free.  This file holds the rest:
//   assemble_band   H straight into lower band storage  AB[d + j LD] = H[j + d, j], LD = 2b
//   chase           band -> real tridiagonal by Householder bulge chasing (Lang's algorithm):
//                   sweep s, step k: reflector on rows s+1+kb .. s+(k+1)b; persistent CTAs, one sweep
//                   each at a time, sweeps of a chain pipelined through release/acquire progress
//                   counters (two steps apart);
//                   the block pushed out by a step stays in shared memory for the next one (3 b^2
//                   elements of global traffic per step).  chase_tmah_kernel<b, ...> (compile-time b,
//                   the default): TMA tensor copies in column pieces, a helper warp for everything that
//                   waits on the memory system, sweeps handed out by ticket counters to one CTA per SM,
//                   the last b sweeps in chase_tail_kernel.  This sweep-owning kernel is the band route's
//                   fallback (DWHMC_CHASE=sweep); the default is the position-owning kernel of band_systolic.cu.
// The back-transformation U = Q2 Z (T factors, block reflectors on the FP64 tensor cores, rows back to the
// reference's site order) is in band_apply.cu.
// The numerics (LAPACK-style zlarfg / zhetd2 updates, application order of the blocks) are the
// ones prototyped against LAPACK in tests/algo_proto_band.py.
#include <cooperative_groups.h>
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dwhmc.h"
#include "gemm_dmma.cuh"
#include "internal.h"

#ifndef DWHMC_CHASE_PCC
#define DWHMC_CHASE_PCC 2
#endif

namespace {

constexpr int CT = 512;             // threads per CTA of the chase kernel
constexpr int CW = CT / 32;         // warps

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {       // acc += a b
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfms(cplx& acc, cplx a, cplx b) {       // acc -= a b
  acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfmac(cplx& acc, cplx a, cplx b) {      // acc += conj(a) b
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}
__device__ __forceinline__ cplx block_sum(cplx v, cplx* red) {          // result in every thread, fixed order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  cplx t = make_double2(0.0, 0.0);
  for (int i = 0; i < CW; ++i) t = cadd(t, red[i]);
  return t;
}
// the band is shared between the CTAs of a chain: all accesses go to L2 (no stale L1 lines)
__device__ __forceinline__ cplx ldg2(const cplx* p) {
  cplx v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg2(cplx* p, cplx v) {
  asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// polling: relaxed loads while spinning (an acquire load invalidates the SM's L1 on every iteration and holds up the
// load/store pipe the compute warps are using), one acquire fence once the value is there
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// ---- assembly into band storage ------------------------------------------------------------------
// pos[r]: band index of row r of the reference's matrix (r < N particle of site r, r >= N hole)
__global__ void band_scatter_kernel(cplx* __restrict__ ABall, const double* __restrict__ w, const double* __restrict__ par3,
                                    const cplx* __restrict__ delta, const int* __restrict__ nn, const int* __restrict__ nnn,
                                    const int* __restrict__ pos, int N, int B, int LD, Mask mask) {
  const int b = blockIdx.y;
  if (!mask.on(b)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int n = 2 * N;
  cplx* AB = ABall + (size_t)b * n * LD;
  const double t = par3[b], tp = par3[B + b], mu = par3[2 * B + b];
  const double term = w[(size_t)b * N + i] - mu;
  auto put = [&](int r, int c, double re, double im) {     // entry (r, c) of the Hermitian matrix
    const int pr = pos[r], pc = pos[c];
    if (pr >= pc) AB[(size_t)pc * LD + (pr - pc)] = make_double2(re, im);
    else AB[(size_t)pr * LD + (pc - pr)] = make_double2(re, -im);
  };
  put(i, i, term, 0.0);
  put(i + N, i + N, -term, 0.0);
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int j = nn[d * N + i];
    if (j != i) { put(i, j, -t, 0.0); put(i + N, j + N, t, 0.0); }
  }
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int j = nnn[d * N + i];
    if (j != i) { put(i, j, -tp, 0.0); put(i + N, j + N, tp, 0.0); }
  }
  const cplx* dl = delta + (size_t)b * 2 * N;
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    const int j = nn[dir * N + i];
    const cplx v = dl[dir * N + i];
    put(i, j + N, 0.5 * v.x, 0.5 * v.y);
    put(j, i + N, 0.5 * v.x, 0.5 * v.y);
  }
}

// ---- bulge chasing ---------------------------------------------------------------------------------
struct ChaseArgs {
  cplx* AB; cplx* V; cplx* tau2; int* prog;
  int n, b, LD, KT, P, c0;       // c0: first chain of this launch
  int* next; int B, stride;      // helper-warp kernel: per-chain sweep tickets, chains, rotation stride of the spare CTAs
  int nsweep;                    // sweeps handed out by tickets (the rest is left to chase_tail_kernel)
  int* status;                   // device status words: [2] set when a progress wait timed out
  Mask mask;
  long long* clk;                // optional [8] phase clock accumulators of CTA 0 (profiling experiments)
};

// ---- TMA chase kernel for a compile-time half-bandwidth ----------------------------------------------------
// The b x b blocks are covered by a TR x TC thread grid with an
// RB x CB sub-block per thread, so every offset is a compile-time constant off one per-thread base and the
// products run on register operands.  The two b x b block transfers of a step go through the TMA engine instead
// of the load/store units, as tensor copies over a 3-D tensor map of the skewed band storage (rows beyond the
// matrix are clipped by the hardware, so partial blocks need no special path): the carried block is updated in
// place in shared memory and written back (shared -> global, bulk group), the next block is fetched
// (global -> shared, mbarrier complete_tx).  The diagonal block itself stays in registers; its Hermitian product
// needs a row and a column reduction.
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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// shared-memory writes only (the block a tensor store is about to read): does not wait for global stores in flight
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA chase kernel with a helper warp -------------------------------------------------------------------
// Everything that waits on the memory system -- the progress poll of the
// previous sweep, the tensor copies of the carried block (store, wait, fetch of the next block), the re-read of
// the corner element, the fence + release that publishes a step -- is done by lane 0 of a 17th warp, so the 512
// compute threads only ever wait on named barriers the helper has usually reached already:
//   barrier 1  compute threads only (the former __syncthreads)
//   barrier 2  "ready":      helper arrives (poll passed, block landed, corner patched) -> compute threads wait
//   barrier 3  "Bc written": compute threads arrive -> helper waits, then stores the block and fetches the next
//   barrier 4  "D stored":   compute threads arrive -> helper waits, then fences and publishes the step
template <int NC> __device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); }
template <int NC> __device__ __forceinline__ void hbar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NC + 32) : "memory"); }
template <int NC> __device__ __forceinline__ void hbar_arrive(int id) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(NC + 32) : "memory");
}
__host__ __device__ constexpr int chase_nc(int tr, int tc) { return (tr * tc + 31) / 32 * 32; }
// thread-columns per piece of the carried block (helper-warp kernel)
__host__ __device__ constexpr int chase_pcc(int cb) { return DWHMC_CHASE_PCC; }

template <int TB, int TR, int TC, int RB, int CB>
__global__ void __launch_bounds__(chase_nc(TR, TC) + 32, 1) chase_tmah_kernel(ChaseArgs g, const __grid_constant__ CUtensorMap tmapA,
                                                                              const __grid_constant__ CUtensorMap tmapB) {
  constexpr int NC = chase_nc(TR, TC);           // compute threads (whole warps); the helper warp follows
  static_assert(TR * RB == TB && TC * CB >= TB && TC * (CB - 1) < TB && TB <= NC && NC + 32 <= 512, "cover, at most 16 warps");
  constexpr bool XC = TC * CB == TB;             // exact column cover; otherwise the last column of a thread may not exist
#define JV(cc) (XC || (cc) < CB - 1 || cj + (cc) * TC < TB)
  // The carried block travels in column pieces of PCC thread-columns each (PCC TC matrix columns; the last piece takes
  // the rest): a piece is written back as soon as it is updated, and its place is refilled from the next block as
  // soon as the write-back has read it.  tmapA: boxes of PCC TC columns, tmapB: box of the last piece.
  constexpr int PCC = chase_pcc(CB);
  constexpr int NPIECE = (CB + PCC - 1) / PCC;
  constexpr int PW = PCC * TC;
  static_assert(5 + NPIECE <= 16, "named barriers");
  static_assert(NPIECE == 1 || (PW * TB * sizeof(cplx)) % 128 == 0, "tensor copies need 128-byte aligned shared memory");
  constexpr int LDB = TB;        // dense box layout of the tensor copies
  constexpr int LDP = TB + 1;    // partial sums: odd leading dimension, conflict-free in both directions
  constexpr int LD = 2 * TB;
  constexpr int NP = (TR > TC) ? TR : TC;
  // Sweeps are handed out by a ticket counter per chain: a CTA takes the next sweep of its home chain (CTA i < 2 B:
  // chain i mod B, for good; the spare CTAs move on by `stride` chains after every sweep), and of the following
  // chains once that one is exhausted.  Tickets are taken in sweep order and a taken sweep is run at once, so the
  // sweep a CTA waits for is always in flight on a co-resident CTA (cooperative launch).
  extern __shared__ __align__(128) unsigned char smem_tma[];
  const int n = g.n;
  cplx* Bc = reinterpret_cast<cplx*>(smem_tma);      // [TB][LDB]
  cplx* vs = Bc + LDB * TB;
  cplx* vp = vs + TB;
  cplx* us = vp + TB;
  cplx* xs = us + TB;
  cplx* tu = xs + TB;
  cplx* wc = tu + TB;
  cplx* part = wc + TB;                              // [NP][LDP]
  cplx* red = part + NP * LDP;                       // [32]
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(red + 32);
  volatile int* sw = reinterpret_cast<volatile int*>(bar + 1);   // [3] chain and sweep of the ticket just taken (-1: none left), spin count
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  unsigned ephase = 0;                               // parity of the next-block barrier

  if (tid >= NC) {
    // ================= helper warp =================
    const bool l0 = tid == NC;
    bool store_pending = false, publish_pending = false;
    int pub_s = 0, pub_v = 0;
    const bool spare = (int)blockIdx.x >= 2 * g.B;
    int home = (spare ? (int)blockIdx.x - 2 * g.B : (int)blockIdx.x / 2) % g.B;   // neighbouring CTAs share a chain
    const int hop = 1 + (int)((blockIdx.x * 2654435761u >> 8) % (unsigned)g.B);      // where to look once home is exhausted
    for (;;) {
      int s = -1, chain = 0;
      if (l0) {
        for (int tries = 0; tries <= g.B; ++tries) {     // home first, then every chain starting from a CTA-specific one
          int c = (tries == 0) ? home : (home + hop + tries - 1) % g.B;
          if (!g.mask.on(c)) continue;
          const int v = atomicAdd(g.next + c, 1);
          if (v < g.nsweep) { s = v; chain = c; break; }
        }
        sw[0] = chain; sw[1] = s;
      }
      s = __shfl_sync(0xffffffffu, s, 0);
      chain = __shfl_sync(0xffffffffu, chain, 0);
      if (s < 0) { hbar_arrive<NC>(2); return; }       // nothing left anywhere
      home = spare ? (chain + g.stride) % g.B : chain;
      cplx* AB = g.AB + (size_t)chain * n * LD;
      int* prog = g.prog + (size_t)chain * n;
      int k = 0, r0 = s + 1, lcar = 0;
      while (true) {
        const int ln = min(TB, n - r0);
        if (l0) {
          if (s > 0) {
            const int need = k + 2;
            sw[2] = 0;                                 // spin count (in shared memory: the kernel has no register to spare);
            while (ld_relaxed(prog + s - 1) < need) {  // bounded: a lost neighbour ends in an error code, not a hang
              __nanosleep(20);
              if (++sw[2] > (1 << 24)) { atomicExch(g.status + 2, 1); break; }
            }
            fence_acq();
          }
          if (k > 0) {
            // the corner element is fetched while the last piece of the block is still landing
            const cplx corner = (lcar == TB) ? ldg2(AB + (size_t)(r0 - 1) * LD + TB) : make_double2(0.0, 0.0);
            mbar_wait(bar, ephase);
            if (lcar == TB) Bc[(TB - 1) * LDB + TB - 1] = corner;
          }
        }
        if (k > 0) ephase ^= 1;
        __syncwarp();
        if (publish_pending) hbar_sync<NC>(4);             // the diagonal block of the previous step is stored
        hbar_arrive<NC>(2);                                // this step may start
        if (publish_pending) {
          if (l0) {
            if (store_pending) bulk_wait_all();        // write-back performed in global memory
            fence_async();
            __threadfence();
            st_release(prog + pub_s, pub_v);
          }
          store_pending = false;
          publish_pending = false;
        }
        if (k > 0 && ln <= 1) break;
        const int r1 = r0 + ln;
        const int l2 = (r1 < n) ? min(TB, n - r1) : 0;
        auto store_piece = [&](int pc) {               // rows r0 .. (clipped at n), columns r0-TB+pc PW ..
          const CUtensorMap* tm = (pc < NPIECE - 1) ? &tmapA : &tmapB;
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                       ::"l"(tm), "r"(2 * r0), "r"(r0 - TB + pc * PW), "r"(chain), "r"(smem_u32(Bc + pc * PW * LDB)) : "memory");
          bulk_commit();
        };
        auto load_piece = [&](int pc) {                // next block: rows r1 .., columns r0+pc PW ..
          const CUtensorMap* tm = (pc < NPIECE - 1) ? &tmapA : &tmapB;
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       ::"r"(smem_u32(Bc + pc * PW * LDB)), "l"(tm), "r"(2 * r1), "r"(r0 + pc * PW), "r"(chain), "r"(smem_u32(bar)) : "memory");
        };
        if (k > 0) {
#pragma unroll
          for (int pc = 0; pc < NPIECE; ++pc) {
            hbar_sync<NC>(5 + pc);                     // piece pc of the carried block updated in shared memory
            if (l0) {
              store_piece(pc);
              if (l2 > 0) {                            // refill the piece as soon as its write-back has read it
                bulk_wait_read();
                if (pc == 0) { fence_async(); mbar_expect_tx(bar, (unsigned)(TB * TB * sizeof(cplx))); }
                load_piece(pc);
              }
            }
          }
          store_pending = true;
        } else if (l0 && l2 > 0) {
          fence_async();
          mbar_expect_tx(bar, (unsigned)(TB * TB * sizeof(cplx)));
#pragma unroll
          for (int pc = 0; pc < NPIECE; ++pc) load_piece(pc);
        }
        if (l2 == 0) break;
        publish_pending = true; pub_s = s; pub_v = k + 1;
        lcar = l2;
        r0 = r1;
        ++k;
      }
      hbar_sync<NC>(4);                                    // all writes of the sweep issued
      if (l0) {
        if (store_pending) bulk_wait_all();
        fence_async();
        __threadfence();
        st_release(prog + s, 1 << 30);
      }
      store_pending = false;
      publish_pending = false;
    }
  }

  // ================= compute threads =================
  const bool act = tid < TR * TC;
  const int ri = act ? tid % TR : 0, cj = act ? tid / TR : 0;
  const int soff = cj * LDB + ri;
  const int goff = cj * (LD - 1) + ri;
  const cplx zero = make_double2(0.0, 0.0);
  cplx* AB = nullptr;
  // partial sums are added up by four neighbouring lanes per row / column (fixed order), then two shuffles
  static_assert(NC >= 4 * TB, "four lanes per entry");
  const int ei = tid >> 2, esl = tid & 3;
  auto sum4 = [&](int col, int nq, bool valid) -> cplx {     // sum_q part[q][col], q < nq; result in all four lanes
    cplx a = zero;
    if (valid)
      for (int q = esl; q < nq; q += 4) a = cadd(a, part[q * LDP + col]);
    a.x += __shfl_xor_sync(0xffffffffu, a.x, 1); a.y += __shfl_xor_sync(0xffffffffu, a.y, 1);
    a.x += __shfl_xor_sync(0xffffffffu, a.x, 2); a.y += __shfl_xor_sync(0xffffffffu, a.y, 2);
    return a;
  };
  auto red_sum = [&]() -> cplx {                             // sum of the per-warp partials, four independent chains
    cplx a0 = zero, a1 = zero, a2 = zero, a3 = zero;
#pragma unroll
    for (int w = 0; w < NC / 32; w += 4) {
      a0 = cadd(a0, red[w]);
      if (w + 1 < NC / 32) a1 = cadd(a1, red[w + 1]);
      if (w + 2 < NC / 32) a2 = cadd(a2, red[w + 2]);
      if (w + 3 < NC / 32) a3 = cadd(a3, red[w + 3]);
    }
    return cadd(cadd(a0, a1), cadd(a2, a3));
  };
#define VCOL (g.V + ((size_t)sw[0] * n + s) * n)      /* reflector column of this sweep (rarely needed: not kept in registers) */
#ifdef DWHMC_CHASE_PROF                              // phase clocks of CTA 0 (experiments): accumulators in shared memory
  __shared__ long long tph[8];
  long long tlast = 0;
  const bool prof = g.clk != nullptr && blockIdx.x == 0 && tid == 0;
  if (prof) for (int i = 0; i < 8; ++i) tph[i] = 0;
#define PH(i) do { if (prof) { const long long t_ = clock64(); tph[i] += t_ - tlast; tlast = t_; } } while (0)
#else
  constexpr bool prof = false;
  long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
#define PH(i) do { } while (0)
#endif
  for (;;) {
    int s = 0, k = 0, r0 = 0;
    cplx taup = zero;
    int lcar = 0;                                    // rows of the carried block in Bc
    while (true) {
      if (prof) tlast = clock64();
      hbar_sync<NC>(2);                                  // sweep s-1 two steps ahead, carried block landed and patched
      PH(7);
      if (k == 0) {                                  // a new sweep: which one, of which chain
        s = sw[1];
        if (s < 0) break;
        const int chain = sw[0];
        r0 = s + 1;
        AB = g.AB + (size_t)chain * n * LD;
      }
      const int ln = min(TB, n - r0);
      double nrm2 = 0.0;                               // |x[1:]|^2, summed where x is produced
      if (k > 0) {
        mbar_wait(bar, ephase);                      // completed already: makes the bulk copy visible to this thread
        ephase ^= 1;
        if (act) {
          cplx acc[RB];
#pragma unroll
          for (int q = 0; q < RB; ++q) acc[q] = zero;
#pragma unroll
          for (int cc = 0; cc < CB; ++cc) {
            if (!JV(cc)) continue;
            const cplx vj = vp[cj + cc * TC];
#pragma unroll
            for (int q = 0; q < RB; ++q)
              if (ri + q * TR < lcar) cfma(acc[q], Bc[soff + cc * TC * LDB + q * TR], vj);
          }
#pragma unroll
          for (int q = 0; q < RB; ++q) part[cj * LDP + ri + q * TR] = acc[q];
        }
        csync<NC>();
        {
          const cplx u = sum4(ei, TC, ei < lcar);
          if (esl == 0 && ei < lcar) {
            us[ei] = u;
            // the column to annihilate comes out of the same pass: x = Bc[:, 0] - taup u (vp[0] = 1)
            const cplx t = cmul(taup, u);
            tu[ei] = t;
            const cplx x = csub(Bc[ei], t);
            xs[ei] = x;
            if (ei > 0) nrm2 += x.x * x.x + x.y * x.y;
          }
        }
      }
      PH(0);
      if (k > 0 && ln <= 1) {
        csync<NC>();
        for (int idx = tid; idx < ln * TB; idx += NC) {
          const int i = idx % ln, j = idx / ln;
          cplx a = Bc[j * LDB + i];
          cfms(a, cmul(taup, us[i]), cconj(vp[j]));
          stg2(AB + (size_t)(r0 - TB + j) * LD + (TB + i - j), a);
        }
        for (int i = tid; i < ln; i += NC) VCOL[r0 + i] = zero;
        break;
      }
      // ---- prefetch the lower triangle of the diagonal block into registers
      cplx* baseD = AB + (size_t)r0 * LD + goff;
      cplx dreg[RB][CB];
      auto prefetch_D = [&]() {
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
          for (int q = 0; q < RB; ++q) {
            const int i = ri + q * TR, j = cj + c * TC;
            if (q * TR + TR - 1 < c * TC) continue;    // sub-block wholly above the diagonal: never referenced
            dreg[q][c] = (act && i >= j && i < ln) ? ldg2(baseD + c * TC * (LD - 1) + q * TR) : zero;
          }
      };
      prefetch_D();
      // ---- A. column to annihilate
      if (k == 0) {
        for (int i = tid; i < ln; i += NC) {
          const cplx x = ldg2(AB + (size_t)s * LD + 1 + i);
          xs[i] = x;
          if (i > 0) nrm2 += x.x * x.x + x.y * x.y;
        }
      }
      nrm2 = warp_sum(make_double2(nrm2, 0.0)).x;
      if ((tid & 31) == 0) red[tid >> 5] = make_double2(nrm2, 0.0);
      PH(5);
      csync<NC>();
      // ---- B. reflector (LAPACK zlarfg; every thread computes tau, beta and the scale)
      cplx tau; double beta;
      {
        const double nrm = red_sum().x;
        const cplx alpha = xs[0];
        cplx scale;
        if (nrm == 0.0 && alpha.y == 0.0) {
          beta = alpha.x;
          tau = zero;
          scale = zero;
        } else {
          beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + nrm), alpha.x);
          const double dr = alpha.x - beta, di = alpha.y;
          const double ibeta = 1.0 / beta, iden = 1.0 / (dr * dr + di * di);   // two independent divisions
          tau = make_double2(-dr * ibeta, -di * ibeta);
          scale = make_double2(dr * iden, -di * iden);
        }
        for (int i = tid; i < ln; i += NC) vs[i] = (i == 0) ? make_double2(1.0, 0.0) : cmul(xs[i], scale);
        csync<NC>();
      }
      PH(2);
      for (int i = tid; i < ln; i += NC) VCOL[r0 + i] = vs[i];
      if (tid == 0) g.tau2[((size_t)sw[0] * n + s) * g.KT + k] = tau;
      cplx vr[RB];
#pragma unroll
      for (int q = 0; q < RB; ++q) vr[q] = (ri + q * TR < ln) ? vs[ri + q * TR] : zero;
      if (k == 0) {
        for (int i = tid; i < ln; i += NC) stg2(AB + (size_t)s * LD + 1 + i, (i == 0) ? make_double2(beta, 0.0) : zero);
      } else {
        // ---- C. carried block, updated in place; the helper writes it back with the bulk-copy engine
        // (v^H tu rides along as column TB of the partial sums: LDP = TB + 1)
        if (act) {
#pragma unroll
          for (int cc = 0; cc < CB; ++cc) {
            if (!JV(cc)) continue;
            cplx acc = zero;
#pragma unroll
            for (int q = 0; q < RB; ++q) cfmac(acc, vr[q], Bc[soff + cc * TC * LDB + q * TR]);
            part[ri * LDP + cj + cc * TC] = acc;
          }
          if (cj == 0) {
            cplx acc = zero;
#pragma unroll
            for (int q = 0; q < RB; ++q)
              if (ri + q * TR < ln) cfmac(acc, vr[q], tu[ri + q * TR]);
            part[ri * LDP + TB] = acc;
          }
        }
        csync<NC>();
        PH(1);
        const cplx ctau = cconj(tau);
        {
          cplx z = sum4(ei, TR, ei < TB);
          const cplx c = sum4(TB, TR, true);
          if (esl == 0 && ei < TB) {
            cfms(z, c, cconj(vp[ei]));
            wc[ei] = cmul(ctau, z);
          }
        }
        csync<NC>();
        {
          cplx tur[RB];
#pragma unroll
          for (int q = 0; q < RB; ++q) tur[q] = tu[min(ri + q * TR, TB - 1)];
#pragma unroll
          for (int cc = 0; cc < CB; ++cc) {
            if (act && JV(cc)) {
              const int j = cj + cc * TC;
              const cplx cvp = cconj(vp[j]), wj = wc[j];
#pragma unroll
              for (int q = 0; q < RB; ++q) {
                const int i = ri + q * TR;
                if (i < ln) {
                  cplx o = Bc[soff + cc * TC * LDB + q * TR];
                  cfms(o, tur[q], cvp);
                  cfms(o, vr[q], wj);
                  if (j == 0) o = (i == 0) ? make_double2(beta, 0.0) : zero;
                  Bc[soff + cc * TC * LDB + q * TR] = o;
                }
              }
            }
            if ((cc + 1) % PCC == 0 || cc == CB - 1) {
              fence_async_smem();                       // generic-proxy writes of Bc -> visible to the bulk engine
              hbar_arrive<NC>(5 + cc / PCC);            // the helper writes this piece back
            }
          }
        }
      }
      PH(3);
      // ---- D. diagonal block from registers: x = tau D v, D Hermitian (lower part held)
      {
        // row part: sum_{j <= i} D[i,j] v[j]
        if (act) {
          cplx acc[RB];
#pragma unroll
          for (int q = 0; q < RB; ++q) acc[q] = zero;
#pragma unroll
          for (int cc = 0; cc < CB; ++cc) {
            const int j = cj + cc * TC;
            const cplx vj = (j < ln) ? vs[j] : zero;
#pragma unroll
            for (int q = 0; q < RB; ++q) {
              if (q * TR + TR - 1 < cc * TC) continue;
              cplx a = dreg[q][cc];
              if (ri + q * TR == j) a.y = 0.0;
              cfma(acc[q], a, vj);
            }
          }
#pragma unroll
          for (int q = 0; q < RB; ++q) part[cj * LDP + ri + q * TR] = acc[q];
        }
        csync<NC>();
        {
          const cplx wv = sum4(ei, TC, ei < ln);
          if (esl == 0 && ei < ln) xs[ei] = wv;
        }
        csync<NC>();
        // column part: sum_{i > j} conj(D[i,j]) v[i]
        if (act) {
#pragma unroll
          for (int cc = 0; cc < CB; ++cc) {
            if (!JV(cc)) continue;
            const int j = cj + cc * TC;
            cplx acc = zero;
#pragma unroll
            for (int q = 0; q < RB; ++q)
              if (q * TR + TR - 1 >= cc * TC && ri + q * TR > j) cfmac(acc, dreg[q][cc], vr[q]);
            part[ri * LDP + j] = acc;
          }
        }
        csync<NC>();
        cplx dot = zero;
        {
          cplx wv = sum4(ei, TR, ei < ln);
          if (esl == 0 && ei < ln) {
            wv = cmul(tau, cadd(xs[ei], wv));
            xs[ei] = wv;
            cfmac(dot, wv, vs[ei]);
          }
        }
        dot = warp_sum(dot);
        if ((tid & 31) == 0) red[tid >> 5] = dot;       // red was last read before the barrier that closed the reflector
        csync<NC>();
      }
      cplx alpha;                                      // w = y + alpha v, formed where it is used
      {
        alpha = cmul(tau, red_sum());
        alpha.x *= -0.5; alpha.y *= -0.5;
      }
      PH(4);
      const int r1 = r0 + ln;
      const int l2 = (r1 < n) ? min(TB, n - r1) : 0;
      // ---- D update and store (lower part, from registers)
      if (act) {
        cplx wr[RB];
#pragma unroll
        for (int q = 0; q < RB; ++q) { wr[q] = xs[min(ri + q * TR, TB - 1)]; cfma(wr[q], alpha, vr[q]); }
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) {
          if (!JV(cc)) continue;
          const int j = cj + cc * TC;
          const cplx vj = vs[j];
          cplx wj = xs[j];
          cfma(wj, alpha, vj);
          const cplx cwj = cconj(wj), cvj = cconj(vj);
#pragma unroll
          for (int q = 0; q < RB; ++q) {
            const int i = ri + q * TR;
            if (q * TR + TR - 1 >= cc * TC && i >= j && i < ln) {
              cplx a = dreg[q][cc];
              if (i == j) a.y = 0.0;
              cfms(a, vr[q], cwj);
              cfms(a, wr[q], cvj);
              if (i == j) a.y = 0.0;
              stg2(baseD + cc * TC * (LD - 1) + q * TR, a);
            }
          }
        }
      }
      if (l2 == 0) break;
      for (int i = tid; i < ln; i += NC) vp[i] = vs[i];
      taup = tau;
      lcar = l2;
      PH(6);
      hbar_arrive<NC>(4);
      r0 = r1;
      ++k;
    }
    if (s < 0) break;
    hbar_arrive<NC>(4);
  }
  if (prof) for (int i = 0; i < 8; ++i) g.clk[i] = tph[i];
#undef PH
#undef JV
#undef VCOL
}

template <int TB, int TR, int TC>
constexpr size_t chase_tma_smem() {
  return sizeof(cplx) * ((size_t)TB * TB + 6 * TB + (size_t)((TR > TC) ? TR : TC) * (TB + 1) + 32) + 32;
}

// Tensor map of the band storage of all chains as the skewed view T[chain][c][r] = AB[chain][c LD + (r - c)]
// = base + chain n LD + c (LD - 1) + r (in complex elements; FP64 element type, so the inner extent is 2 n).
static int make_band_tensor_map(Handle* h, CUtensorMap* out, int box_cols) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    h->err = "cuTensorMapEncodeTiled not available";
    return DWHMC_E_CUDA;
  }
  const cuuint64_t n = (cuuint64_t)h->n, LD = (cuuint64_t)h->band_LD, b = (cuuint64_t)h->band_b;
  const cuuint64_t gdim[3] = {2 * n, n, (cuuint64_t)h->B};
  const cuuint64_t gstr[2] = {(LD - 1) * sizeof(cplx), n * LD * sizeof(cplx)};
  const cuuint32_t box[3] = {(cuuint32_t)(2 * b), (cuuint32_t)box_cols, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult rc = reinterpret_cast<EncodeFn>(fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, h->A, gdim, gstr, box, estr,
                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")"; return DWHMC_E_CUDA; }
  return DWHMC_OK;
}

// ---- the last b sweeps: a dense (b+1) x (b+1) Hermitian block, one CTA per chain -------------------------
// Sweeps s >= n-1-b have a single step each and would run strictly one after the other through the progress
// counters.  Here the trailing block is copied to shared memory (full storage) and tridiagonalised in place with
// the same reflector and update formulas as the chase; reflectors and tau go where the chase would have put them.
constexpr int TAIL_T = 512;
__global__ void __launch_bounds__(TAIL_T, 1) chase_tail_kernel(ChaseArgs g) {
  const int chain = blockIdx.x;
  if (!g.mask.on(chain)) return;
  extern __shared__ __align__(16) unsigned char smem_tail[];
  const int n = g.n, b = g.b, LD = g.LD;
  const int m0 = b + 1, ldf = m0 | 1;                // odd leading dimension
  cplx* F = reinterpret_cast<cplx*>(smem_tail);      // [m0][ldf] column-major, both triangles
  cplx* vs = F + (size_t)ldf * m0;                   // [m0]
  cplx* ys = vs + m0;                                // [m0]
  cplx* red = ys + m0;                               // [32]
  const int tid = threadIdx.x;
  const int s0 = n - 1 - b;                          // global index of local row / column 0
  cplx* AB = g.AB + (size_t)chain * n * LD;
  cplx* V = g.V + (size_t)chain * n * n;
  cplx* tau2 = g.tau2 + (size_t)chain * n * g.KT;
  const cplx zero = make_double2(0.0, 0.0);
  for (int idx = tid; idx < m0 * m0; idx += TAIL_T) {
    const int r = idx % m0, c = idx / m0;
    if (r >= c) {
      cplx a = ldg2(AB + (size_t)(s0 + c) * LD + (r - c));
      if (r == c) a.y = 0.0;
      F[(size_t)c * ldf + r] = a;
      if (r > c) F[(size_t)r * ldf + c] = cconj(a);
    }
  }
  __syncthreads();
  const int ei = tid >> 2, esl = tid & 3;            // four lanes per row of the matrix-vector product
  for (int j = 0; j < b; ++j) {
    const int m = b - j;                             // order of the trailing block A22 = F[j+1.., j+1..]
    const int s = s0 + j;
    const cplx* x = F + (size_t)j * ldf + j + 1;     // column to annihilate
    // reflector (zlarfg)
    double nrm2 = 0.0;
    for (int i = 1 + tid; i < m; i += TAIL_T) { const cplx a = x[i]; nrm2 += a.x * a.x + a.y * a.y; }
    const cplx nr = block_sum(make_double2(nrm2, 0.0), red);
    const cplx alpha0 = x[0];
    cplx tau, scale; double beta;
    if (nr.x == 0.0 && alpha0.y == 0.0) {
      beta = alpha0.x; tau = zero; scale = zero;
    } else {
      beta = -copysign(sqrt(alpha0.x * alpha0.x + alpha0.y * alpha0.y + nr.x), alpha0.x);
      const double dr = alpha0.x - beta, di = alpha0.y;
      const double ibeta = 1.0 / beta, iden = 1.0 / (dr * dr + di * di);
      tau = make_double2(-dr * ibeta, -di * ibeta);
      scale = make_double2(dr * iden, -di * iden);
    }
    for (int i = tid; i < m; i += TAIL_T) {
      const cplx v = (i == 0) ? make_double2(1.0, 0.0) : cmul(x[i], scale);
      vs[i] = v;
      V[(size_t)s * n + s + 1 + i] = v;
    }
    if (tid == 0) tau2[(size_t)s * g.KT] = tau;
    __syncthreads();
    // y = tau A22 v
    {
      cplx acc = zero;
      if (ei < m) {
        const cplx* row = F + (size_t)(j + 1) * ldf + (j + 1 + ei);
        for (int c = esl; c < m; c += 4) cfma(acc, row[(size_t)c * ldf], vs[c]);
      }
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
      cplx dot = zero;
      if (esl == 0 && ei < m) {
        const cplx wv = cmul(tau, acc);
        ys[ei] = wv;
        cfmac(dot, wv, vs[ei]);
      }
      dot = block_sum(dot, red);
      cplx al = cmul(tau, dot);
      al.x *= -0.5; al.y *= -0.5;
      // A22 -= v w^H + w v^H with w = y + alpha v (both triangles are kept)
      for (int idx = tid; idx < m * m; idx += TAIL_T) {
        const int r = idx % m, c = idx / m;
        cplx wr = ys[r], wc = ys[c];
        const cplx vr = vs[r], vc = vs[c];
        cfma(wr, al, vr);
        cfma(wc, al, vc);
        cplx a = F[(size_t)(j + 1 + c) * ldf + (j + 1 + r)];
        cfms(a, vr, cconj(wc));
        cfms(a, wr, cconj(vc));
        if (r == c) a.y = 0.0;
        F[(size_t)(j + 1 + c) * ldf + (j + 1 + r)] = a;
      }
    }
    if (tid == 0) F[(size_t)j * ldf + j + 1] = make_double2(beta, 0.0);
    __syncthreads();
  }
  // diagonal and sub-diagonal back to the band storage (all that band_de_kernel reads)
  for (int c = tid; c < m0; c += TAIL_T) {
    stg2(AB + (size_t)(s0 + c) * LD, make_double2(F[(size_t)c * ldf + c].x, 0.0));
    if (c + 1 < m0) stg2(AB + (size_t)(s0 + c) * LD + 1, make_double2(F[(size_t)c * ldf + c + 1].x, 0.0));
  }
}

__global__ void band_de_kernel(const cplx* __restrict__ ABall, double* __restrict__ d, double* __restrict__ e, int n,
                               int LD, Mask mask) {
  const int b = blockIdx.y;
  if (!mask.on(b)) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const cplx* col = ABall + ((size_t)b * n + j) * LD;
  d[(size_t)b * n + j] = col[0].x;
  e[(size_t)b * n + j] = (j + 1 < n) ? col[1].x : 0.0;
}


}  // namespace

bool dw_band_has_tma_kernel(int bw);
bool dw_band_has_systolic_kernel(int bw);
int dw_band_chase_systolic(Handle* h, Mask mask);

// ring fold: positions 0, L-1, 1, L-2, ... -> consecutive indices
static std::vector<int> fold_positions(int L) {
  std::vector<int> pos(L);
  int lo = 0, hi = L - 1, idx = 0;
  while (lo <= hi) {
    pos[lo] = idx++;
    if (hi != lo) pos[hi] = idx++;
    ++lo; --hi;
  }
  return pos;
}

// Decide whether the band route applies and prepare its index tables.  nn / nnn: 0-based [dir * N + site].
int dw_band_setup(Handle* h, const std::vector<int>& nn, const std::vector<int>& nnn) {
  const int Lx = h->Lx, Ly = h->Ly, N = h->N, n = h->n;
  h->band_b = 0;
  std::vector<int> px = fold_positions(Lx), py = fold_positions(Ly);
  std::vector<int> pos(n);
  for (int y = 0; y < Ly; ++y)
    for (int x = 0; x < Lx; ++x) {
      const int i = y * Lx + x;
      const int sidx = (Lx <= Ly) ? py[y] * Lx + px[x] : px[x] * Ly + py[y];   // short ring fastest
      pos[i] = 2 * sidx;
      pos[i + N] = 2 * sidx + 1;
    }
  int bw = 0;
  for (int i = 0; i < N; ++i) {
    for (int d = 0; d < 4; ++d) {
      const int j = nn[d * N + i], j2 = nnn[d * N + i];
      bw = std::max(bw, std::abs(pos[i] - pos[j]));
      bw = std::max(bw, std::abs(pos[i] - pos[j2]));
    }
    for (int dir = 0; dir < 2; ++dir) {
      const int j = nn[dir * N + i];
      bw = std::max(bw, std::abs(pos[i] - pos[j + N]));
      bw = std::max(bw, std::abs(pos[j] - pos[i + N]));
    }
  }
  bw = std::max(bw, 2);
  // A band of half-width bw is also a band of any larger half-width: round up to the next width with a compile-time
  // chase kernel (odd sides, L = 22, ...) as long as the matrix is still clearly wider than the band.
  for (int cand = bw; cand <= 100; ++cand)
    if (dw_band_has_tma_kernel(cand)) {
      if (3 * cand <= n) bw = cand;
      break;
    }
  // Default wherever the chase kernels have an instance for this bandwidth (faster than the dense route at every
  // size measured, 6 <= L <= 24); DWHMC_BAND=0 forces the dense route.
  int want = dw_band_has_tma_kernel(bw) ? 1 : 0;
  if (const char* e = getenv("DWHMC_BAND")) want = want && atoi(e);
  if (!want || 3 * bw > n) return DWHMC_OK;   // dense route
  // reflectors per block of the back-transformation (band_apply.cu): 32, block height b + 31 <= 136 rows
  if (bw + DW_APPLY_G - 1 > DW_APPLY_ROWS) return DWHMC_OK;                                // dense route
  const int g = DW_APPLY_G;
  h->band_b = bw;
  h->band_LD = 2 * bw;
  h->band_KT = (n + bw - 1) / bw + 1;
  h->band_g = g;
  // block list, application order: sweep groups last to first, steps ascending
  std::vector<int> bs0, bk;
  const int ngrp = (n - 1 + g - 1) / g;
  for (int G = ngrp - 1; G >= 0; --G) {
    const int s0 = G * g;
    for (int k = 0;; ++k) {
      const int r0 = s0 + 1 + k * bw;                 // first row of the first sweep of the group
      const int len = n - r0;
      if (len < 1 || (k > 0 && len < 2)) break;
      bs0.push_back(s0);
      bk.push_back(k);
    }
  }
  h->band_blk_s0 = bs0;
  h->band_blk_k = bk;
  // wavefronts of mutually independent blocks: t = (ngrp - 1 - G) + k.  Block (G, k) has to come after (G+1, k),
  // (G+1, k-1), (G+2, k-1), ... (later sweep groups whose rows overlap) and after (G, k-1): all have a smaller t;
  // blocks with equal t, (G, k) and (G-j, k-j), touch disjoint rows.
  {
    const int nb2 = (int)bs0.size();
    std::vector<int> tval(nb2), order(nb2);
    int tmax = 0;
    for (int i = 0; i < nb2; ++i) {
      tval[i] = (ngrp - 1 - bs0[i] / g) + bk[i];
      order[i] = i;
      tmax = std::max(tmax, tval[i]);
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return tval[a] < tval[c]; });
    h->band_wave_blk = order;
    h->band_wave_start.assign(tmax + 2, 0);
    for (int i = 0; i < nb2; ++i) h->band_wave_start[tval[i] + 1]++;
    for (int t = 0; t <= tmax; ++t) h->band_wave_start[t + 1] += h->band_wave_start[t];
  }
  h->band_pos_host = pos;
  return DWHMC_OK;
}

int dw_band_assemble(Handle* h, const double* w, const double* par3, const cplx* delta, Mask mask) {
  const int n = h->n, N = h->N, B = h->B, LD = h->band_LD;
  DW_CUDA(h, cudaMemsetAsync(h->A, 0, sizeof(cplx) * (size_t)n * LD * B, h->stream));
  dim3 grid((N + 127) / 128, B);
  band_scatter_kernel<<<grid, 128, 0, h->stream>>>(h->A, w, par3, delta, h->nn, h->nnn, h->band_pos, N, B, LD, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

// The CTAs of a chase launch wait on each other, so two chase launches must never share the GPU half-resident
// (two handles driven from two host threads would otherwise be able to deadlock).  Launches of one process on
// one device are chained through an event: each waits for the previous one, whichever handle issued it.
#include <mutex>
static std::mutex g_chase_mutex;
static cudaEvent_t g_chase_done[64] = {};

// one guarded cooperative launch (band_systolic.cu)
int dw_chase_launch_guarded(Handle* h, const void* kern, int nctas, int nthreads, void** args, size_t smem) {
  std::lock_guard<std::mutex> lock(g_chase_mutex);
  cudaEvent_t& done = g_chase_done[h->device & 63];
  if (!done) DW_CUDA(h, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  else DW_CUDA(h, cudaStreamWaitEvent(h->stream, done, 0));
  struct Rec { cudaEvent_t e; cudaStream_t s; ~Rec() { cudaEventRecord(e, s); } } rec{done, h->stream};
  DW_CUDA(h, cudaLaunchCooperativeKernel(kern, dim3(nctas), dim3(nthreads), args, smem, h->stream));
  h->launches++;
  return DWHMC_OK;
}

// launch helper: P persistent CTAs per chain, all co-resident (cooperative launch), chains in slices if needed
template <class Launch>
static int chase_launch_loop(Handle* h, Mask mask, const void* kern, int nthreads, size_t smem, bool tickets, Launch launch) {
  const int n = h->n, B = h->B, bw = h->band_b;
  int per_sm = 0;
  DW_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthreads, smem));
  const int cap = h->nsm * per_sm;
  if (cap < 1) { h->err = "dw_band_chase: kernel does not fit"; return DWHMC_E_CUDA; }
  const int P = std::max(1, std::min(4, cap / B));
  // ticket kernel: one launch over all chains with every CTA that fits (at most 4 per chain)
  const int per_launch = tickets ? B : std::max(1, cap / P);      // chains per launch
  std::lock_guard<std::mutex> lock(g_chase_mutex);
  cudaEvent_t& done = g_chase_done[h->device & 63];
  if (!done) DW_CUDA(h, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  else DW_CUDA(h, cudaStreamWaitEvent(h->stream, done, 0));
  struct Rec { cudaEvent_t e; cudaStream_t s; ~Rec() { cudaEventRecord(e, s); } } rec{done, h->stream};
  for (int c0 = 0; c0 < B; c0 += per_launch) {
    ChaseArgs a;
    a.AB = h->A; a.V = h->V; a.tau2 = h->band_tau; a.prog = h->band_prog;
    a.n = n; a.b = bw; a.LD = h->band_LD; a.KT = h->band_KT; a.P = P; a.c0 = c0; a.mask = mask;
    a.next = h->band_prog + (size_t)n * B; a.B = B; a.stride = 1; a.status = h->status;
    const bool tail = tickets && n - 1 - bw >= 1;
    a.nsweep = tail ? n - 1 - bw : n - 1;
#ifdef DWHMC_CHASE_PROF                                  // phase clocks of CTA 0 (experiments; needs the kernels built with the flag)
    static long long* clk_dev = nullptr;
    if (!clk_dev) cudaMalloc(&clk_dev, 8 * sizeof(long long));
    a.clk = (c0 == 0) ? clk_dev : nullptr;
#else
    a.clk = nullptr;
#endif
    const int nch = std::min(per_launch, B - c0);
    int nctas = nch * P;
    if (tickets) {
      nctas = std::min(cap, 4 * B);
      auto gcd = [](int x, int y) { while (y) { const int t = x % y; x = y; y = t; } return x; };
      a.stride = std::max(1, nctas - 2 * B);                     // spare CTAs visit every chain in turn
      while (gcd(a.stride, B) != 1) ++a.stride;
    }
    DW_TRY(launch(a, nctas));
    h->launches++;
    if (tail) {
      const size_t tsm = sizeof(cplx) * ((size_t)((bw + 1) | 1) * (bw + 1) + 2 * (size_t)(bw + 1) + 32);
      static bool tattr[64] = {false};
      if (!tattr[h->device & 63]) {
        DW_CUDA(h, cudaFuncSetAttribute(chase_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        tattr[h->device & 63] = true;
      }
      chase_tail_kernel<<<B, TAIL_T, tsm, h->stream>>>(a);
      DW_LAUNCH_CHECK(h);
      h->launches++;
    }
    if (a.clk) {
      long long c[8];
      cudaStreamSynchronize(h->stream);
      cudaMemcpy(c, a.clk, sizeof(c), cudaMemcpyDeviceToHost);
      fprintf(stderr, "chase phase clocks (Mclk) [0..7]: %.1f %.1f %.1f %.1f %.1f %.1f %.1f %.1f\n",
              c[0] / 1e6, c[1] / 1e6, c[2] / 1e6, c[3] / 1e6, c[4] / 1e6, c[5] / 1e6, c[6] / 1e6, c[7] / 1e6);
    }
  }
  return DWHMC_OK;
}

template <int TB, int TR, int TC, int RB, int CB>
static int chase_tma_dispatch(Handle* h, Mask mask) {
  const int nthreads = chase_nc(TR, TC) + 32;
  const size_t smem = chase_tma_smem<TB, TR, TC>();
  const void* kern = (const void*)chase_tmah_kernel<TB, TR, TC, RB, CB>;
  static bool attr[64] = {false};
  if (!attr[h->device & 63]) {
    DW_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[h->device & 63] = true;
  }
  static_assert(sizeof(CUtensorMap) <= sizeof(h->band_tmap_a), "tensor map storage");
  if (!h->band_tmap_set) {
    constexpr int pcc = chase_pcc(CB), npiece = (CB + pcc - 1) / pcc, pw = pcc * TC;   // column pieces of the carried block
    DW_TRY(make_band_tensor_map(h, reinterpret_cast<CUtensorMap*>(h->band_tmap_a), std::min(pw, TB)));
    DW_TRY(make_band_tensor_map(h, reinterpret_cast<CUtensorMap*>(h->band_tmap_b), TB - pw * (npiece - 1)));
    h->band_tmap_set = true;
  }
  return chase_launch_loop(h, mask, kern, nthreads, smem, true, [&](ChaseArgs& a, int ctas) -> int {
    void* targs[] = {&a, h->band_tmap_a, h->band_tmap_b};
    DW_CUDA(h, cudaLaunchCooperativeKernel(kern, dim3(ctas), dim3(nthreads), targs, smem, h->stream));
    return DWHMC_OK;
  });
}

// half-bandwidths with a compile-time TMA kernel: b = 4 L + 4 for the short side L = 6, 8, ..., 20, 24
// (92 = 4 x 23 has no exact TR x TC cover within 512 threads)
bool dw_band_has_tma_kernel(int bw) { return bw >= 28 && bw <= 100 && (bw - 28) % 8 == 0 && bw != 92; }

// h->A (band) -> h->d, h->e, h->V (reflectors, column s = sweep s), h->band_tau
int dw_band_chase(Handle* h, Mask mask) {
  const int n = h->n, B = h->B, bw = h->band_b;
  DW_CUDA(h, cudaMemsetAsync(h->band_prog, 0, sizeof(int) * ((size_t)n + 1) * B, h->stream));
  DW_CUDA(h, cudaMemsetAsync(h->band_tau, 0, sizeof(cplx) * (size_t)n * h->band_KT * B, h->stream));
  DW_CUDA(h, cudaMemsetAsync(h->band_bbox, 0, sizeof(cplx) * 2 * (size_t)h->band_KT * B, h->stream));
  // Default wherever it has an instance: the position-owning kernel of band_systolic.cu (one chain finishes in n step
  // times instead of 2 n, and the blocks never leave the SM): at L = 24 10.0 instead of 30.5 ms for up to 12 chains and
  // 30.2 instead of 46 ms for 64; L = 20: 15.8 against 18.6 ms for 64 chains, L = 16: 7.0 against 8.3.
  // DWHMC_CHASE=sweep forces the sweep-owning TMA kernel below, the band route's second kernel.
  static const char* which = getenv("DWHMC_CHASE");
  bool use_sys = dw_band_has_systolic_kernel(bw);
  if (which && which[0] == 's' && which[1] == 'w') use_sys = false;
  if (use_sys) DW_TRY(dw_band_chase_systolic(h, mask));
  else if (bw == 100) DW_TRY((chase_tma_dispatch<100, 25, 19, 4, 6>(h, mask)));
  else if (bw == 84) DW_TRY((chase_tma_dispatch<84, 21, 21, 4, 4>(h, mask)));
  else if (bw == 76) DW_TRY((chase_tma_dispatch<76, 19, 19, 4, 4>(h, mask)));
  else if (bw == 68) DW_TRY((chase_tma_dispatch<68, 17, 17, 4, 4>(h, mask)));
  else if (bw == 60) DW_TRY((chase_tma_dispatch<60, 15, 20, 4, 3>(h, mask)));
  else if (bw == 52) DW_TRY((chase_tma_dispatch<52, 13, 26, 4, 2>(h, mask)));
  else if (bw == 44) DW_TRY((chase_tma_dispatch<44, 22, 11, 2, 4>(h, mask)));
  else if (bw == 36) DW_TRY((chase_tma_dispatch<36, 18, 18, 2, 2>(h, mask)));
  else if (bw == 28) DW_TRY((chase_tma_dispatch<28, 14, 14, 2, 2>(h, mask)));
  else { h->err = "dw_band_chase: no kernel for this half-bandwidth"; return DWHMC_E_BADARG; }
  dim3 grid((n + 255) / 256, B);
  band_de_kernel<<<grid, 256, 0, h->stream>>>(h->A, h->d, h->e, n, h->band_LD, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

