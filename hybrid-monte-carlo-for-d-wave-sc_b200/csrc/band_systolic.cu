// band_systolic.cu -- position-owning ("systolic") bulge chase of the band route of diagonalize_H_BdG!
// (/root/reference src/Hamiltonian.jl:96-114; the reference calls LAPACK zheevr there).
//
// Same algorithm, reflectors, tau and tridiagonal matrix as chase_tmah_kernel in band.cu (Lang's bulge chasing:
// sweep s, step k, reflector on rows r0 .. r0+b-1, r0 = s + 1 + k b), but the work is cut the other way: a CTA owns
// step k ("position" k) of EVERY sweep of a chain instead of every step of one sweep.  The two b x b blocks of a
// step -- the carried block Bc (rows r0.., columns r0-b..) and the diagonal block D -- then only slide down the
// diagonal by one row and one column per sweep, so they never leave the SM:
//   * element (global row g, global column c) of a window lives at the physical slot (g mod b, c mod b); a slide
//     overwrites one physical row and one physical column, nothing else moves.  Vectors are indexed the same way.
//   * Bc lives in registers (RB x CB elements per thread, compile-time indices), D in shared memory with both
//     triangles (a torus has no fixed lower triangle), rows / columns beyond the matrix are kept at zero, so the
//     block operations carry no length masks.
//   * per sweep a position exchanges O(b) numbers with its neighbours through global memory:
//       k-1 -> k   the reflector v(s, k-1) and tau (the V / tau outputs the back-transformation needs anyway)
//       k+1 -> k   beta(s-1, k+1), the corner of the new last row of Bc(s, k) -- the one number the next reflector of
//                  position k depends on: published right after the reflector, in the 32-byte sector of its own counter;
//                  row 0 of Bc(s-1, k+1) after its update (the new last row of D(s, k)), formed right after z = v^H Bc
//                  and needed before y = D v; the corner D(s-1, k+1)[0, 0], last.  Row and corner go to a mailbox of
//                  two slots (sweep parity) per (chain, position).
//     guarded by release / acquire counters per (chain, position): sweeps published for v + beta, row, corner.
//     The column that leaves D on a slide becomes the new last column of Bc (position 0: the next column to
//     annihilate).  Position 0 also delivers the tridiagonal matrix (d, e) into AB where band_de_kernel reads it.
//   * a helper warp does everything that waits on the memory system: it polls the neighbours' counters (acquire
//     loads), brings their messages into (double-buffered) shared memory ahead of time (L2 loads) and publishes this
//     position's counters (release stores); the compute warps meet it at named barriers only.
// Global traffic per step drops from 3 b^2 elements (two tensor copies and the diagonal block through L2) to ~4 b.
// Tasks (epoch, chain, position) are handed out by one ticket counter in that order: after Q sweeps a position writes
// its windows back to the band storage and whichever CTA is free continues it (so the CTAs the short positions of
// early chains set free go to later chains at once and all chains end together).  Whoever waits for a neighbour waits
// for a task with a smaller ticket or for the next untaken ones, which the CTAs of finished tasks pick up, so the
// cooperative launch cannot deadlock.
// NumPy prototype of exactly this organisation: tests/algo_proto_systolic.py (checked against the sweep-owning
// prototype and LAPACK in tests/test_algo_proto.py).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "dwhmc.h"
#include "internal.h"

namespace {

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
// everything the positions of a chain exchange goes through L2 (no stale L1 lines)
__device__ __forceinline__ cplx ldg2(const cplx* p) {
  cplx v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg2(cplx* p, cplx v) {
  asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#ifndef LDS2_VOLATILE
#define LDS2_VOLATILE volatile
#endif
// shared-memory load the compiler may not hoist or merge (keeps streamed vector entries out of the register file)
__device__ __forceinline__ cplx lds2(const cplx* p) {
  cplx v;
  asm LDS2_VOLATILE("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

struct SysArgs {
  cplx* AB; cplx* V; cplx* tau2;
  cplx* bbox;                    // [B][KT][2]: beta(s, k) and, 16 bytes on, the number of sweeps it is published for (one 32-byte sector)
  cplx* rowbox;                  // [B][KT][2][TB+2]: row message of (chain, position), slot = sweep & 1: row 0 of the updated Bc, then the corner of D
  int* flags;                    // [B][n]: vflag[k] at k, rflag[k] at KT + k, dflag[k] at 2 KT + k, sflag[k] (epochs saved) at 3 KT + k; then [1] task ticket
  int n, LD, KT, KP, B;          // KP: positions per chain
  int Q, NE;                     // sweeps per epoch, epochs: a task = (epoch, chain, position), handed out in this order
  int* status;                   // device status words: [2] set when a wait timed out
  Mask mask;
  long long* clk;                // optional [8] phase clocks of one position (experiments, -DDWHMC_CHASE_PROF)
};

// named barriers: 1 compute threads only; the others are shared with the helper warp (NC + 32 threads)
enum { BAR_VP = 2, BAR_ROW = 3, BAR_CORNER = 4, BAR_VW = 5, BAR_RW = 6, BAR_DW = 7, BAR_TASK = 8, BAR_SAVE = 9, BAR_BETA = 10 };
template <int NC> __device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); }
template <int NC> __device__ __forceinline__ void hbar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NC + 32) : "memory"); }
template <int NC> __device__ __forceinline__ void hbar_arrive(int id) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(NC + 32) : "memory");
}
__host__ __device__ constexpr int sys_nc(int tr, int tc) { return (tr * tc + 31) / 32 * 32; }
// rows of the partial sums of z = v^H Bc: neighbouring lanes (rows ri, ri+1 of one thread column) are added by a
// shuffle first when TR is even
__host__ __device__ constexpr int sys_zrows(int tr) { return (tr % 2 == 0) ? tr / 2 : tr; }

template <int TB, int TR, int TC>
constexpr size_t sys_smem() {
  // D [TB][TB], partial sums [zrows + TC][TB+1], 10 vectors (vp x2, rowmsg x2, vs, xs, tu, wc, ys, xcol), red[32], 8 scalars, control
  return sizeof(cplx) * ((size_t)TB * TB + (size_t)sys_zrows(TR) * (TB + 1) + (size_t)TC * (TB + 8) + 10 * TB + 2 + 32 + 11) + 64;
}

// One step of a position (sweep s), compute warps:
//   P1  u = Bc vp (partial sums per thread column)                                  [needs v of position k-1]
//   P2  u, x = Bc[:,0] - taup u, |x|^2; the row / column that entered D last sweep   [needs the row message of k+1]
//   P3  reflector: tau, beta, v                                  -> V, tau2, (e) to global, vflag published
//   P4  z = v^H Bc (registers) and y = D v (shared memory), partial sums
//   P5  wc = conj(tau) z, y = tau D v, y^H v
//   P6  w = y + alpha v; the first row of the updated Bc (the row message) ahead of the update itself
//   P7  row message and corner to global, rflag published; Bc -= tu vp^H + v wc, D -= v w^H + w v^H
//   P8  slide
template <int TB, int TR, int TC, int RB, int CB>
__global__ void __launch_bounds__(sys_nc(TR, TC) + 32, 1) chase_sys_kernel(SysArgs g) {
  constexpr int NC = sys_nc(TR, TC);
  static_assert(TR * RB == TB && TC * CB >= TB && TC * (CB - 1) < TB, "cover of the b x b windows (rows exactly)");
  constexpr bool XC = TC * CB == TB;             // exact column cover; otherwise the last column of a thread may not exist
#define JV(cc) (XC || (cc) < CB - 1 || cj + (cc) * TC < TB)
  constexpr int LPE = (NC >= 4 * TB) ? 4 : ((NC >= 2 * TB) ? 2 : 1);   // lanes per entry of a matrix-vector product
  static_assert(NC >= TB && NC + 32 <= 1024, "one thread per vector entry at least");
  constexpr bool ZPAIR = TR % 2 == 0;
  constexpr int ZR = sys_zrows(TR);
  constexpr int LDD = TB;
  constexpr int LDP = TB + 1;    // partial sums of z: odd leading dimension (its rows are written by every second lane)
  // partial sums of u and y: read back by LPE lanes per entry, row q = esl + LPE i -- conflict-free when LPE LDY = 8 mod 16
  constexpr int LDY = (LPE == 1) ? TB + 1 : TB + ((8 / LPE - TB % 8) + 8) % 8;
  extern __shared__ __align__(16) unsigned char smem_sys[];
  cplx* D = reinterpret_cast<cplx*>(smem_sys);       // [TB][LDD] physical column-major, both triangles
  cplx* partz = D + TB * LDD;                        // [ZR][LDP] partial sums of z (column TB: v^H tu)
  cplx* party = partz + ZR * LDP;                    // [TC][LDY] partial sums of u, then of y
  cplx* vpbuf = party + TC * LDY;                    // [2][TB] previous reflector (physical column index)
  cplx* rowbuf = vpbuf + 2 * TB;                     // [2][TB+1] row message of position k+1: row, corner
  cplx* vs = rowbuf + 2 * (TB + 1);
  cplx* xs = vs + TB;
  cplx* tu = xs + TB;
  cplx* wc = tu + TB;
  cplx* ys = xs;                                     // (x is dead after the reflector, y is born in P5)
  cplx* xcol = wc + TB;
  cplx* xrow = xcol + TB;                            // row 0 of Bc before the update (for the row message)
  cplx* red = xrow + TB;                             // [32]
  cplx* scal = red + 32;                             // [11]: taup[2], -, -, tau, beta, beta of k+1 [2]
  volatile int* sw = reinterpret_cast<volatile int*>(scal + 11);   // [3] chain, position (-1: none left), epoch of the task just taken
  const int tid = threadIdx.x;
  const int n = g.n, LD = g.LD, KT = g.KT;
  const cplx zero = make_double2(0.0, 0.0);
  int* const ticket = g.flags + (size_t)n * g.B;

  if (tid >= NC) {
    // ================= helper warp =================
    const int lane = tid - NC;
    const bool l0 = lane == 0;
    bool dead = false;                               // a wait timed out: stop waiting, let the launch drain
#ifdef DWHMC_CHASE_PROF
    long long hw[6] = {0, 0, 0, 0, 0, 0};            // helper waits: v, beta, row, corner, epoch, whole tasks
    int hkind = 0;
#define HW_BEGIN(kind) long long hw_t0_ = clock64(); hkind = kind;
#define HW_END() hw[hkind] += clock64() - hw_t0_;
#define HW_TASK_BEGIN() const long long hw_task0 = clock64();
#define HW_TASK_END() do { hw[5] += clock64() - hw_task0; if (l0) for (int i_ = 0; i_ < 6; ++i_) { atomicAdd((unsigned long long*)g.clk + 16 + i_, (unsigned long long)hw[i_]); hw[i_] = 0; } } while (0)
#else
#define HW_BEGIN(kind)
#define HW_END()
#define HW_TASK_BEGIN()
#define HW_TASK_END() do { } while (0)
#endif
    // wait until *flag >= need (acquire loads by lane 0; the payload is then read by the whole warp with L2 loads)
    auto wait_for = [&](const int* flag, int need) {
      if (l0 && !dead) {
        int spins = 0;
        while (ld_acquire(flag) < need) {
          if (++spins > (1 << 22) || ((spins & 1023) == 0 && *((volatile int*)g.status + 2) != 0)) {
            atomicExch(g.status + 2, 1);
            dead = true;
            break;
          }
        }
      }
      __syncwarp();
    };
    for (;;) {
      int chain = 0, k = -1, ep = 0;
      if (l0) {
        for (;;) {
          int t = atomicAdd(ticket, 1);
          int e = 0, kpe = 0;
          for (; e < g.NE; ++e) {                    // tasks of epoch e: B x (positions alive at sweep e Q)
            kpe = (n - 2 - e * g.Q) / TB + 1;
            if (t < g.B * kpe) break;
            t -= g.B * kpe;
          }
          if (e >= g.NE) break;
          const int c = t / kpe;
          if (!g.mask.on(c)) continue;
          chain = c; k = t - c * kpe; ep = e;
          break;
        }
        sw[0] = chain; sw[1] = k; sw[2] = ep;
      }
      chain = __shfl_sync(0xffffffffu, chain, 0);
      k = __shfl_sync(0xffffffffu, k, 0);
      ep = __shfl_sync(0xffffffffu, ep, 0);
      cplx* AB = g.AB + (size_t)chain * n * LD;
      int* fl = g.flags + (size_t)chain * n;
      HW_TASK_BEGIN()
      if (k >= 0 && ep > 0) { HW_BEGIN(4) wait_for(fl + 3 * KT + k, ep); HW_END() }   // the windows of this position as the previous epoch left them
      hbar_arrive<NC>(BAR_TASK);
      if (k < 0) return;
      const int s0 = ep * g.Q, s1 = min((ep + 1) * g.Q, n - 1 - k * TB);
      int r0 = s0 + 1 + k * TB, o = r0 % TB;
      for (int s = s0; s < s1; ++s) {
        const int buf = s & 1;
        // ---- inputs of step s: v and tau of position k-1 (the columns of Bc always exist) ...
        if (k > 0) {
          { HW_BEGIN(0) wait_for(fl + (k - 1), s + 1); HW_END() }
          {
            const cplx* Vseg = g.V + ((size_t)chain * n + s) * n + (r0 - TB);
            for (int p = lane; p < TB; p += 32) vpbuf[buf * TB + p] = ldg2(Vseg + (p - o + (p < o ? TB : 0)));
            if (l0) scal[buf] = ldg2(g.tau2 + ((size_t)chain * n + s) * KT + (k - 1));
          }
        }
        hbar_arrive<NC>(BAR_VP);
        // ... beta of step (s-1, k+1), the corner of the new last row of Bc: the one number of the neighbour's step
        //     the reflector of this step depends on.  It travels in the 32-byte sector of its own counter, so the
        //     poll that sees the counter has it already.
        cplx* rdst = rowbuf + buf * (TB + 1);
        const bool newrow = r0 + TB - 1 < n;         // the row / column that entered the windows exists
        const cplx* box_in = g.rowbox + (((size_t)chain * KT + (k + 1)) * 2 + ((s - 1) & 1)) * (TB + 2);
        if (l0) {
          cplx bt = zero;
          if (newrow) {
            if (s > 0) {
              const cplx* slot = g.bbox + ((size_t)chain * KT + (k + 1)) * 2;
              int spins = 0;
              HW_BEGIN(1)
              const int* bflag = reinterpret_cast<const int*>(slot + 1);
              while (!dead && ld_acquire(bflag) < s) {
                if (++spins > (1 << 22) || ((spins & 1023) == 0 && *((volatile int*)g.status + 2) != 0)) {
                  atomicExch(g.status + 2, 1);
                  dead = true;
                }
              }
              HW_END()
              bt = ldg2(slot);
            } else {
              bt = ldg2(AB + (size_t)(r0 - 1) * LD + TB);   // first sweep: still in the band storage
            }
          }
          scal[6 + buf] = bt;
        }
        __syncwarp();
        hbar_arrive<NC>(BAR_BETA);
        // ---- v and beta of this step out
        hbar_sync<NC>(BAR_VW);
        if (l0) {
          __threadfence();                             // (cumulative over what the barrier ordered)
          asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(fl + k), "r"(s + 1) : "memory");
          asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(reinterpret_cast<int*>(g.bbox + ((size_t)chain * KT + k) * 2 + 1)), "r"(s + 1) : "memory");
        }
        // ---- the rest of the neighbour's messages: row 0 of its updated Bc (the new last row of D), needed before
        //      y = D v, and the corner of its D, needed after; one look at each before this step's own row message
        //      goes out, so that a late neighbour does not hold it up
        auto fetch_row = [&]() {
          for (int j = lane; j < TB; j += 32) rdst[j] = ldg2(box_in + j);
          __syncwarp();
        };
        auto fetch_corner = [&]() {
          if (l0) rdst[TB] = ldg2(box_in + TB);
          __syncwarp();
        };
        bool have_row = false, have_corner = false;
        if (newrow && s > 0) {
          int ready = 0;
          if (l0) ready = ld_acquire(fl + KT + (k + 1)) >= s;
          ready = __shfl_sync(0xffffffffu, ready, 0);
          if (ready) { fetch_row(); have_row = true; }
        } else {
          if (newrow) {                               // first sweep: still in the band storage
            for (int j = lane; j < TB; j += 32) rdst[j] = ldg2(AB + (size_t)(r0 - 1 + j) * LD + (TB - j));
            if (l0) rdst[TB] = ldg2(AB + (size_t)(r0 + TB - 1) * LD);
          } else {
            for (int j = lane; j <= TB; j += 32) rdst[j] = zero;
          }
          have_row = have_corner = true;
        }
        if (have_row) hbar_arrive<NC>(BAR_ROW);
        if (have_row && !have_corner) {
          int ready = 0;
          if (l0) ready = ld_acquire(fl + 2 * KT + (k + 1)) >= s;
          ready = __shfl_sync(0xffffffffu, ready, 0);
          if (ready) { fetch_corner(); have_corner = true; }
        }
        if (have_corner) hbar_arrive<NC>(BAR_CORNER);
        hbar_sync<NC>(BAR_RW);
        if (l0) { st_release(fl + KT + k, s + 1); }
        if (!have_row) {
          { HW_BEGIN(2) wait_for(fl + KT + (k + 1), s); HW_END() }
          fetch_row();
          hbar_arrive<NC>(BAR_ROW);
        }
        if (!have_corner) {
          { HW_BEGIN(3) wait_for(fl + 2 * KT + (k + 1), s); HW_END() }
          fetch_corner();
          hbar_arrive<NC>(BAR_CORNER);
        }
        hbar_sync<NC>(BAR_DW);
        if (l0) { st_release(fl + 2 * KT + k, s + 1); }
        ++r0;
        o = (o + 1 == TB) ? 0 : o + 1;
      }
      hbar_sync<NC>(BAR_SAVE);                         // windows written back (if the position goes on in the next epoch)
      HW_TASK_END();
      if (l0 && s1 < n - 1 - k * TB) { st_release(fl + 3 * KT + k, ep + 1); }
    }
  }

  // ================= compute threads =================
  const bool act = tid < TR * TC;
  const int ri = act ? tid % TR : 0, cj = act ? tid / TR : 0;
  const int ei = tid / LPE, esl = tid % LPE;         // LPE lanes per entry of a matrix-vector product
  auto sumL = [&](const cplx* part, int ld, int col, int nq, bool valid) -> cplx {   // sum_q part[q][col], q < nq; result in all LPE lanes
    cplx a = zero;
    if (valid)
      for (int q = esl; q < nq; q += LPE) a = cadd(a, part[q * ld + col]);
    if (LPE >= 2) { a.x += __shfl_xor_sync(0xffffffffu, a.x, 1); a.y += __shfl_xor_sync(0xffffffffu, a.y, 1); }
    if (LPE >= 4) { a.x += __shfl_xor_sync(0xffffffffu, a.x, 2); a.y += __shfl_xor_sync(0xffffffffu, a.y, 2); }
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
#ifdef DWHMC_CHASE_PROF
  long long tph[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
  bool prof = false;
// time between barrier RELEASES (a clock read right after BAR.SYNC would capture the issue of the barrier, not its
// release: the read is made to depend on a shared-memory load behind the barrier)
#define PH(i) do { if (prof) { const int d_ = *reinterpret_cast<volatile int*>(sw); long long t_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "r"(d_) : "memory"); tph[i] += t_ - tlast; tlast = t_; } } while (0)
  long long cw[6] = {0, 0, 0, 0, 0, 0};            // time blocked at the helper's barriers: v, beta, row, corner, task; whole tasks
#define CW_SYNC(id, slot) do { const long long c0_ = clock64(); hbar_sync<NC>(id); if (tid == 0) { const int d_ = *reinterpret_cast<volatile int*>(sw); long long t_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "r"(d_) : "memory"); cw[slot] += t_ - c0_; } } while (0)
#else
#define PH(i) do { } while (0)
#define CW_SYNC(id, slot) hbar_sync<NC>(id)
#endif
  for (;;) {
    CW_SYNC(BAR_TASK, 4);
    const int chain = sw[0], k = sw[1];
    if (k < 0) break;
#ifdef DWHMC_CHASE_PROF
    const long long ctask0 = clock64();
#endif
#ifdef DWHMC_CHASE_PROF
    if (prof) { for (int i = 0; i < 16; ++i) g.clk[i] = tph[i]; }
    prof = g.clk != nullptr && tid == 0 && chain == 0 && k == DWHMC_CHASE_PROF;
    if (prof) tlast = clock64();
#endif
    cplx* AB = g.AB + (size_t)chain * n * LD;
    const int s0 = sw[2] * g.Q, s1 = min((sw[2] + 1) * g.Q, n - 1 - k * TB);
    int r0 = s0 + 1 + k * TB, o = r0 % TB;
    cplx Bc[RB][CB];
#pragma unroll
    for (int cc = 0; cc < CB; ++cc)
#pragma unroll
      for (int q = 0; q < RB; ++q) Bc[q][cc] = zero;
    // ---- windows of sweep 0 from the band storage (the physical row / column that "just entered" is patched
    //      from the messages of the first step like in every later one)
    {
      const int po = (o == 0) ? TB - 1 : o - 1;
#pragma unroll
      for (int cc = 0; cc < CB; ++cc) {
        if (!JV(cc)) continue;
        const int pc = cj + cc * TC;
        const int lc = pc - o + (pc < o ? TB : 0);
#pragma unroll
        for (int q = 0; q < RB; ++q) {
          const int pr = ri + q * TR;
          const int gr = r0 + pr - o + (pr < o ? TB : 0);
          const bool rowok = act && gr < n && pr != po;
          {
            const int gc = r0 - TB + lc;
            Bc[q][cc] = (rowok && k > 0) ? ldg2(AB + (size_t)gc * LD + (gr - gc)) : zero;
          }
          if (act) {
            const int gc = r0 + lc;
            cplx a = zero;
            if (rowok && gc < n && pc != po) {
              if (gr >= gc) a = ldg2(AB + (size_t)gc * LD + (gr - gc));
              else a = cconj(ldg2(AB + (size_t)gr * LD + (gc - gr)));
              if (gr == gc) a.y = 0.0;
            }
            D[pc * LDD + pr] = a;
          }
        }
      }
      if (k == 0)
        for (int p = tid; p < TB; p += NC) {
          const int gr = r0 + p - o + (p < o ? TB : 0);
          xcol[p] = (gr < n && p != po) ? ldg2(AB + (size_t)(r0 - 1) * LD + (gr - (r0 - 1))) : zero;   // column s0 below the diagonal
        }
    }
    csync<NC>();
    for (int s = s0; s < s1; ++s) {
      const int po = (o == 0) ? TB - 1 : o - 1;      // physical index of the row / column that entered last
      const int buf = s & 1;
      const cplx* vp = vpbuf + buf * TB;
      const cplx* rowm = rowbuf + buf * (TB + 1);
      // ---- P1: u = Bc vp (the new last row of Bc is still zero; its corner comes with the row message)
      CW_SYNC(BAR_VP, 0);
      PH(0);
      if (k > 0 && act) {
        cplx acc[RB];
#pragma unroll
        for (int q = 0; q < RB; ++q) acc[q] = zero;
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) {
          if (!JV(cc)) continue;
          const int pc = cj + cc * TC;
          const cplx vj = vp[pc];
#pragma unroll
          for (int q = 0; q < RB; ++q) cfma(acc[q], Bc[q][cc], vj);
          if (pc == o) {                              // logical column 0: the column to annihilate (before the update)
#pragma unroll
            for (int q = 0; q < RB; ++q) xcol[ri + q * TR] = Bc[q][cc];
          }
        }
#pragma unroll
        for (int q = 0; q < RB; ++q) party[cj * LDY + ri + q * TR] = acc[q];
#pragma unroll
        for (int q = 0; q < RB; ++q)
          if (ri + q * TR == o) {                     // logical row 0 as it is now: the row message is formed from it in P5a
#pragma unroll
            for (int cc = 0; cc < CB; ++cc)
              if (JV(cc)) xrow[cj + cc * TC] = Bc[q][cc];
          }
      }
      csync<NC>();
      PH(1);
      CW_SYNC(BAR_BETA, 1);
      PH(2);
      // ---- P2: u, x and its norm (beta of the neighbour's step: the corner of the new last row of Bc)
      {
        const cplx betain = scal[6 + buf];
        double nrm2 = 0.0;
        if (k > 0) {
          cplx u = sumL(party, LDY, ei, TC, ei < TB);
          if (esl == 0 && ei < TB) {
            if (ei == po) cfma(u, betain, vp[po]);
            const cplx t = cmul(scal[buf], u);
            tu[ei] = t;
            const cplx x = csub(xcol[ei], t);
            xs[ei] = x;
            if (ei != o) nrm2 = x.x * x.x + x.y * x.y;
          }
          if (act) {
#pragma unroll
            for (int q = 0; q < RB; ++q)
              if (ri + q * TR == po) {
#pragma unroll
                for (int cc = 0; cc < CB; ++cc)
                  if (cj + cc * TC == po) Bc[q][cc] = betain;
              }
          }
        } else {
          for (int p = tid; p < TB; p += NC) {
            const cplx x = (p == po) ? betain : xcol[p];
            xs[p] = x;
            if (p != o) nrm2 += x.x * x.x + x.y * x.y;
          }
        }
        nrm2 = warp_sum(make_double2(nrm2, 0.0)).x;
        if ((tid & 31) == 0) red[tid >> 5] = make_double2(nrm2, 0.0);
      }
      csync<NC>();
      PH(3);
      // ---- P3: reflector (LAPACK zlarfg; every thread computes tau, beta and the scale)
      const bool flush = k > 0 && n - r0 <= 1;       // nothing to annihilate: only the pending right-application
      {
        cplx tau, betac;
        const double nrm = red_sum().x;
        const cplx alpha = xs[o];
        cplx scale;
        if (flush) {
          betac = alpha; tau = zero; scale = zero;
        } else if (nrm == 0.0 && alpha.y == 0.0) {
          betac = make_double2(alpha.x, 0.0); tau = zero; scale = zero;
        } else {
          const double beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + nrm), alpha.x);
          const double dr = alpha.x - beta, di = alpha.y;
          const double ibeta = 1.0 / beta, iden = 1.0 / (dr * dr + di * di);   // two independent divisions
          tau = make_double2(-dr * ibeta, -di * ibeta);
          scale = make_double2(dr * iden, -di * iden);
          betac = make_double2(beta, 0.0);
        }
        // v, tau and beta leave for the neighbours from the registers they are computed in (they are on the cycle
        // between neighbouring positions twice per sweep: no barrier and no second pass in front of them)
        cplx* Vcol = g.V + ((size_t)chain * n + s) * n + r0;
        for (int p = tid; p < TB; p += NC) {
          const cplx v = (p == o) ? make_double2(1.0, 0.0) : cmul(xs[p], scale);
          vs[p] = v;
          const int lg = p - o + (p < o ? TB : 0);
          if (r0 + lg < n) stg2(Vcol + lg, flush ? zero : v);
        }
        if (tid == 0) {
          scal[4] = tau; scal[5] = betac;                 // later phases re-read them (registers are scarce)
          stg2(g.tau2 + ((size_t)chain * n + s) * KT + k, tau);
          if (k == 0) stg2(AB + (size_t)s * LD + 1, betac);   // e[s]
          else stg2(g.bbox + ((size_t)chain * KT + k) * 2, betac);   // beta: position k-1 needs it before anything else
        }
      }
      hbar_arrive<NC>(BAR_VW);
      csync<NC>();
      PH(4);
      // ---- P4a-P6a (positions k > 0): z = v^H Bc, wc = conj(tau) z, and row r0 of the updated Bc -- the row message
      //      position k-1 is waiting for -- ahead of everything that does not feed it
      if (k > 0) {
        // (the shuffles are executed by whole warps: idle threads of the last warp hold a zero block)
        cplx vr[RB];
#pragma unroll
        for (int q = 0; q < RB; ++q) vr[q] = vs[ri + q * TR];
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) {
          cplx acc = zero;
#pragma unroll
          for (int q = 0; q < RB; ++q) cfmac(acc, vr[q], Bc[q][cc]);
          if (ZPAIR) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
            if (act && JV(cc) && (ri & 1) == 0) partz[(ri >> 1) * LDP + cj + cc * TC] = acc;
          } else if (act && JV(cc)) {
            partz[ri * LDP + cj + cc * TC] = acc;
          }
        }
        {
          cplx acc = zero;                             // v^H tu rides along as column TB (thread column 0 only)
#pragma unroll
          for (int q = 0; q < RB; ++q) cfmac(acc, vr[q], tu[ri + q * TR]);
          if (ZPAIR) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
            if (act && cj == 0 && (ri & 1) == 0) partz[(ri >> 1) * LDP + TB] = acc;
          } else if (act && cj == 0) {
            partz[ri * LDP + TB] = acc;
          }
        }
        csync<NC>();
        PH(5);
        {
          // wc = conj(tau) z, and with it row r0 of the updated Bc -- the row message -- straight to its mailbox:
          // Bc[0, j] - tu[0] conj(vp[j]) - v[0] wc[j] with v[0] = 1; the annihilated column holds beta
          cplx z = sumL(partz, LDP, ei, ZR, ei < TB);
          const cplx c = sumL(partz, LDP, TB, ZR, true);
          if (esl == 0 && ei < TB) {
            const cplx cvp = cconj(vp[ei]);
            cfms(z, c, cvp);
            const cplx w = cmul(cconj(lds2(scal + 4)), z);
            wc[ei] = w;
            cplx a = xrow[ei];
            cfms(a, tu[o], cvp);
            a = csub(a, w);
            if (ei == o) a = lds2(scal + 5);
            cplx* box = g.rowbox + (((size_t)chain * KT + k) * 2 + buf) * (TB + 2);
            stg2(box + (ei - o + (ei < o ? TB : 0)), a);   // logical column order
          }
        }
      }
      hbar_arrive<NC>(BAR_RW);
      // ---- the row / column that entered D (the rest of the neighbour's row message; its corner is added in P5b)
      CW_SYNC(BAR_ROW, 2);
      for (int p = tid; p < TB; p += NC) {
        if (p != po) {
          const int j = p - po + (p < po ? TB : 0);
          const cplx val = rowm[j];
          D[p * LDD + po] = val;                        // row po, column p
          D[po * LDD + p] = cconj(val);
        } else {
          D[po * LDD + po] = zero;
        }
      }
      csync<NC>();
      // ---- P4b: y = D v (shared memory), partial sums
      if (act) {
        cplx acc[RB];
#pragma unroll
        for (int q = 0; q < RB; ++q) acc[q] = zero;
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) {
          if (!JV(cc)) continue;
          const int pc = cj + cc * TC;
          const cplx vj = vs[pc];
#pragma unroll
          for (int q = 0; q < RB; ++q) cfma(acc[q], D[pc * LDD + ri + q * TR], vj);
        }
#pragma unroll
        for (int q = 0; q < RB; ++q) party[cj * LDY + ri + q * TR] = acc[q];
      }
      csync<NC>();
      PH(8);
      CW_SYNC(BAR_CORNER, 3);
      PH(9);
      // ---- P5b: y = tau D v (with the corner that entered), y^H v
      {
        cplx dot = zero;
        cplx y = sumL(party, LDY, ei, TC, ei < TB);
        if (esl == 0 && ei < TB) {
          if (ei == po) {
            const double corner = rowm[TB].x;
            const cplx vpo = vs[po];
            y.x = fma(corner, vpo.x, y.x); y.y = fma(corner, vpo.y, y.y);
            D[po * LDD + po] = make_double2(corner, 0.0);
          }
          y = cmul(lds2(scal + 4), y);
          ys[ei] = y;
          cfmac(dot, y, vs[ei]);
        }
        dot = warp_sum(dot);
        if ((tid & 31) == 0) red[tid >> 5] = dot;       // red was last read before the barrier that closed the reflector
      }
      csync<NC>();
      PH(10);
      // ---- P6b: w = y + alpha v (zhetd2), in place
      {
        cplx alpha2 = cmul(lds2(scal + 4), red_sum());
        alpha2.x *= -0.5; alpha2.y *= -0.5;
        for (int p = tid; p < TB; p += NC) {
          cplx w = ys[p]; cfma(w, alpha2, vs[p]); ys[p] = w;
          if (p == o) {
            // the corner message: D[0,0] after the update (v[0] = 1); position 0: d[s+1], straight into the band
            // storage.  It is formed here, by the thread that has w[0], because D is not stable once the barrier
            // below has opened: the first warps of the update phase rewrite D[0,0] at once.
            const cplx c = make_double2(D[o * LDD + o].x - 2.0 * w.x, 0.0);
            if (k > 0) stg2(g.rowbox + (((size_t)chain * KT + k) * 2 + buf) * (TB + 2) + TB, c);
            else stg2(AB + (size_t)r0 * LD, c);
          }
        }
      }
      csync<NC>();
      PH(11);
      // ---- P7: the corner message is on its way; the two block updates
      hbar_arrive<NC>(BAR_DW);
      if (act) {
        // Two independent updates: Bc in registers (pure FP64 work; row operands v, tu held, column operands by broadcast
        // loads) and D in shared memory (a read and a write per element: shared-memory bound).  Warps alternate in
        // which one they do first, so that at any time half of them load the FP64 pipe and half the shared-memory
        // pipe.  Inside a loop the operands of column c+1 are requested before the arithmetic of column c.
        auto bc_update = [&]() {
          if (k == 0) return;
          cplx vq[RB], tq[RB];
#pragma unroll
          for (int q = 0; q < RB; ++q) { vq[q] = vs[ri + q * TR]; tq[q] = tu[ri + q * TR]; }
          cplx cvpn = lds2(vp + cj), wjn = lds2(wc + cj);
#pragma unroll
          for (int cc = 0; cc < CB; ++cc) {
            if (!JV(cc)) continue;
            const cplx cvp = cconj(cvpn), wj = wjn;
            if (cc + 1 < CB && JV(cc + 1)) { cvpn = lds2(vp + cj + (cc + 1) * TC); wjn = lds2(wc + cj + (cc + 1) * TC); }
#pragma unroll
            for (int q = 0; q < RB; ++q) { cfms(Bc[q][cc], tq[q], cvp); cfms(Bc[q][cc], vq[q], wj); }
          }
        };
        auto d_update = [&]() {
          cplx vq[RB], wq[RB];
#pragma unroll
          for (int q = 0; q < RB; ++q) { vq[q] = vs[ri + q * TR]; wq[q] = ys[ri + q * TR]; }
          cplx dn[RB];
#pragma unroll
          for (int q = 0; q < RB; ++q) dn[q] = lds2(D + cj * LDD + ri + q * TR);
#pragma unroll
          for (int cc = 0; cc < CB; ++cc) {
            if (!JV(cc)) continue;
            const int j = cj + cc * TC;
            cplx dq[RB];
#pragma unroll
            for (int q = 0; q < RB; ++q) dq[q] = dn[q];
            const cplx cwj = cconj(lds2(ys + j)), cvj = cconj(lds2(vs + j));
            if (cc + 1 < CB && JV(cc + 1)) {
#pragma unroll
              for (int q = 0; q < RB; ++q) dn[q] = lds2(D + (j + TC) * LDD + ri + q * TR);
            }
#pragma unroll
            for (int q = 0; q < RB; ++q) {
              cfms(dq[q], vq[q], cwj);
              cfms(dq[q], wq[q], cvj);
              if (ri + q * TR == j) dq[q].y = 0.0;
              D[j * LDD + ri + q * TR] = dq[q];
            }
          }
        };
        if ((tid >> 5) & 1) { bc_update(); d_update(); }
        else { d_update(); bc_update(); }
      }
      csync<NC>();
      PH(12);
      // ---- P8 slide: the column that leaves D is the new last column of Bc (position 0: the next column to
      //      annihilate); the row that leaves makes room for the new last row (zero until its message arrives).
      //      (Logical column 0 of Bc, annihilated by this step, is the one that is overwritten.)
      if (k > 0) {
        if (act) {
#pragma unroll
          for (int cc = 0; cc < CB; ++cc)
            if (cj + cc * TC == o) {
#pragma unroll
              for (int q = 0; q < RB; ++q) Bc[q][cc] = D[o * LDD + ri + q * TR];
            }
#pragma unroll
          for (int q = 0; q < RB; ++q)
            if (ri + q * TR == o) {
#pragma unroll
              for (int cc = 0; cc < CB; ++cc) Bc[q][cc] = zero;
            }
        }
      } else {
        for (int p = tid; p < TB; p += NC) xcol[p] = (p == o) ? zero : D[o * LDD + p];
      }
      ++r0;
      o = (o + 1 == TB) ? 0 : o + 1;
    }
    // ---- end of the epoch: if the position goes on, its windows go back to the band storage for whichever CTA takes
    //      the next epoch (the row / column that entered last is still on its way as a message and is not written)
    if (s1 < n - 1 - k * TB) {
      const int po = (o == 0) ? TB - 1 : o - 1;
      if (k == 0) {
        for (int p = tid; p < TB; p += NC) {
          const int gr = r0 + p - o + (p < o ? TB : 0);
          if (gr < n && p != po) stg2(AB + (size_t)(r0 - 1) * LD + (gr - (r0 - 1)), xcol[p]);
        }
      }
      if (act) {
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) {
          if (!JV(cc)) continue;
          const int pc = cj + cc * TC;
          const int lc = pc - o + (pc < o ? TB : 0);
#pragma unroll
          for (int q = 0; q < RB; ++q) {
            const int pr = ri + q * TR;
            const int gr = r0 + pr - o + (pr < o ? TB : 0);
            if (gr >= n || pr == po) continue;
            if (k > 0) {
              const int gc = r0 - TB + lc;
              stg2(AB + (size_t)gc * LD + (gr - gc), Bc[q][cc]);
            }
            const int gc = r0 + lc;
            if (gc <= gr && pc != po) stg2(AB + (size_t)gc * LD + (gr - gc), D[pc * LDD + pr]);
          }
        }
      }
    }
    hbar_arrive<NC>(BAR_SAVE);
#ifdef DWHMC_CHASE_PROF
    if (tid == 0 && g.clk) { cw[5] = clock64() - ctask0; for (int i_ = 0; i_ < 6; ++i_) { atomicAdd((unsigned long long*)g.clk + 24 + i_, (unsigned long long)cw[i_]); cw[i_] = 0; } }
#endif
  }
#ifdef DWHMC_CHASE_PROF
  if (prof) { for (int i = 0; i < 16; ++i) g.clk[i] = tph[i]; }
#endif
#undef PH
#undef JV
}

}  // namespace

// launch serialisation shared with the sweep-owning kernels (band.cu): chase launches wait on each other
int dw_chase_launch_guarded(Handle* h, const void* kern, int nctas, int nthreads, void** args, size_t smem);

template <int TB, int TR, int TC, int RB, int CB>
static int sys_dispatch(Handle* h, Mask mask) {
  const int nthreads = sys_nc(TR, TC) + 32;
  const size_t smem = sys_smem<TB, TR, TC>();
  const void* kern = (const void*)chase_sys_kernel<TB, TR, TC, RB, CB>;
  static bool attr[64] = {false};
  if (!attr[h->device & 63]) {
    DW_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[h->device & 63] = true;
  }
  int per_sm = 0;
  DW_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthreads, smem));
  if (per_sm < 1) { h->err = "dw_band_chase_systolic: kernel does not fit"; return DWHMC_E_CUDA; }
  SysArgs a;
  a.AB = h->A; a.V = h->V; a.tau2 = h->band_tau; a.flags = h->band_prog; a.rowbox = h->band_rowbox; a.bbox = h->band_bbox;
  a.n = h->n; a.LD = h->band_LD; a.KT = h->band_KT; a.KP = (h->n - 2) / TB + 1; a.B = h->B;
  a.status = h->status; a.mask = mask; a.clk = nullptr;
  // Epochs: with more position tasks than CTAs a chain is cut into epochs of Q sweeps, so that the CTAs the short
  // positions of early chains set free go to later chains at once and every chain ends at about the same time (one
  // epoch per chain would leave the last chain running alone for n step times).  Small batches: a single epoch.
  a.Q = (h->B * a.KP <= h->nsm * per_sm) ? h->n : std::max(32, 5 * TB / 4);
  if (const char* e = getenv("DWHMC_CHASE_Q")) a.Q = std::max(8, atoi(e));
  a.NE = (h->n - 1 + a.Q - 1) / a.Q;
  const int nctas = std::min(h->nsm * per_sm, h->B * a.KP);
  void* args[] = {&a};
#ifdef DWHMC_CHASE_PROF                                  // phase clocks of position DWHMC_CHASE_PROF of chain 0 (experiments)
  static long long* clk_dev = nullptr;
  if (!clk_dev) { cudaMalloc(&clk_dev, 32 * sizeof(long long)); }
  cudaMemsetAsync(clk_dev, 0, 32 * sizeof(long long), h->stream);
  a.clk = clk_dev;
  DW_TRY(dw_chase_launch_guarded(h, kern, nctas, nthreads, args, smem));
  long long c[16];
  cudaStreamSynchronize(h->stream);
  cudaMemcpy(c, clk_dev, sizeof(c), cudaMemcpyDeviceToHost);
  const double st = 1e3 * (h->n - 1 - DWHMC_CHASE_PROF * TB);
  fprintf(stderr, "systolic kclk/step between barrier releases (k > 0: VP | u | ROW | P2 | P3 | z | wc | row | y | CORNER | ysum | w | upd; k = 0: three fewer):");
  double tot = 0;
  for (int i = 0; i < 16; ++i) { fprintf(stderr, " %.2f", c[i] / st); tot += c[i] / st; }
  fprintf(stderr, "  total %.2f\n", tot);
  long long hwv[6];
  cudaMemcpy(hwv, clk_dev + 16, sizeof(hwv), cudaMemcpyDeviceToHost);
  fprintf(stderr, "helper waits, share of all task time: v %.3f beta %.3f row %.3f corner %.3f epoch %.3f (tasks: %.1f Mclk on %d CTAs)\n",
          (double)hwv[0] / hwv[5], (double)hwv[1] / hwv[5], (double)hwv[2] / hwv[5], (double)hwv[3] / hwv[5], (double)hwv[4] / hwv[5],
          hwv[5] / 1e6, nctas);
  cudaMemcpy(hwv, clk_dev + 24, sizeof(hwv), cudaMemcpyDeviceToHost);
  fprintf(stderr, "compute warps blocked at the helper's barriers, share of all task time: v %.3f beta %.3f row %.3f corner %.3f next task %.3f\n",
          (double)hwv[0] / hwv[5], (double)hwv[1] / hwv[5], (double)hwv[2] / hwv[5], (double)hwv[3] / hwv[5], (double)hwv[4] / hwv[5]);
  return DWHMC_OK;
#else
  return dw_chase_launch_guarded(h, kern, nctas, nthreads, args, smem);
#endif
}

bool dw_band_has_systolic_kernel(int bw) {
  switch (bw) {
    case 100: case 84: case 76: case 68: case 60: case 52: case 44: case 36: case 28: return true;
    default: return false;
  }
}

// h->A (band) -> reflectors h->V, h->band_tau; the tridiagonal matrix is left in the band storage (band_de_kernel)
int dw_band_chase_systolic(Handle* h, Mask mask) {
  switch (h->band_b) {
    case 100: return sys_dispatch<100, 50, 7, 2, 15>(h, mask);   // 2 x 15 elements per thread: the row operands of an update fit next to the block
    case 84: return sys_dispatch<84, 42, 8, 2, 11>(h, mask);     // (two rows per thread everywhere: 8-10 % ahead of 4 x 4 ... 4 x 6 blocks)
    case 76: return sys_dispatch<76, 38, 10, 2, 8>(h, mask);
    case 68: return sys_dispatch<68, 34, 10, 2, 7>(h, mask);
    case 60: return sys_dispatch<60, 30, 10, 2, 6>(h, mask);
    case 52: return sys_dispatch<52, 26, 13, 2, 4>(h, mask);
    case 44: return sys_dispatch<44, 22, 11, 2, 4>(h, mask);
    case 36: return sys_dispatch<36, 18, 18, 2, 2>(h, mask);
    case 28: return sys_dispatch<28, 14, 14, 2, 2>(h, mask);
    default: h->err = "dw_band_chase_systolic: no kernel for this half-bandwidth"; return DWHMC_E_BADARG;
  }
}
