"""Discrete-event model of the sweep pipeline of the bulge chase (csrc/band.cu, chase_tmah_kernel): chains, CTAs, tickets,
the two-step lag between consecutive sweeps of a chain, publish latency.  One time unit = one full step.  It reproduces
the measured scaling with CTAs per chain (16 chains at L = 24: 2 / 3 / 4 / 6 CTAs per chain -> 49.1 / 37.7 / 33.1 / 31.4 ms
measured, 1 : 0.77 : 0.67 : 0.64; model 1 : 0.75 : 0.66 : 0.63) and the gain of the spare CTAs (64 chains, 128 -> 148
CTAs: measured 0.915, model 0.90), so scheduling policies can be compared without GPU time:

    python tools/chase_pipeline_sim.py

Findings (64 chains, 148 CTAs): utilisation 0.90; keeping the spare CTAs longer with a chain or letting a CTA prefer a
chain whose next sweep can start at once is no better (0.89 ... 0.72); the loss is the region of short sweeps (K <= 4
steps: at most K/2 CTAs per chain are useful), about 600 of the 3060 time units, which all chains cross together."""
import heapq
import math

def simulate(B=64, ncta=148, n=1152, b=100, policy="ticket", T=1.0, delta=0.12, t_ticket=0.08, stay=1, k0cost=0.55,
             dense_tail=True, verbose=False):
    S = n - 1 - b if dense_tail else n - 1
    def K(s): return (n - 1 - s + b - 1) // b
    def cost(s, k):
        Ks = K(s)
        if k == Ks - 1:
            f = ((n - 1 - s) - (Ks - 1) * b) / b
        else:
            f = 1.0
        base = k0cost if k == 0 else 1.0
        return T * base * (0.45 + 0.55 * f)
    nxt = [0] * B                      # tickets
    pub = {}                           # (c, s) -> list of publish times per completed step; done flag time
    done = {}
    waiters = {}                       # (c, s) -> list of (need, cta)
    # CTA state
    home = [0] * ncta; spare = [False] * ncta; nstay = [0] * ncta
    stride = max(1, ncta - 2 * B)
    while math.gcd(stride, B) != 1: stride += 1
    for i in range(ncta):
        if policy == "static":
            home[i] = (i // 2) % B
        else:
            spare[i] = i >= 2 * B
            home[i] = ((i - 2 * B) if spare[i] else i // 2) % B
    hop = [1 + ((i * 2654435761 >> 8) % B) for i in range(ncta)]
    state = [None] * ncta              # (c, s, k)
    heap = []
    busy = [0.0] * ncta
    wait = [0.0] * ncta
    finish = 0.0
    def can_start(c, s, k, t):
        """earliest time step k of sweep s may start given what is known; None if unknown yet"""
        if s == 0: return t
        if (c, s - 1) in done: 
            return max(t, done[(c, s - 1)])
        p = pub.get((c, s - 1), [])
        if len(p) >= k + 2: return max(t, p[k + 1])
        return None
    def take_ticket(i, t):
        c0 = home[i]
        order = [c0] + [(c0 + hop[i] + j) % B for j in range(B)]
        if policy == "prefer":
            # first pass: a chain whose next sweep can start at once (peek), only if the home chain would wait
            for c in order[:1 + 6]:
                s = nxt[c]
                if s >= S: continue
                if s == 0 or (c, s - 1) in done or len(pub.get((c, s - 1), [])) >= 2:
                    nxt[c] += 1
                    return c, s
        for c in order:
            if nxt[c] < S:
                s = nxt[c]; nxt[c] += 1
                return c, s
        return None
    for i in range(ncta):
        heapq.heappush(heap, (0.0, i, "ticket"))
    while heap:
        t, i, what = heapq.heappop(heap)
        if what == "ticket":
            r = take_ticket(i, t)
            if r is None:
                finish = max(finish, t); continue
            c, s = r
            if spare[i]:
                nstay[i] += 1
                if nstay[i] >= stay:
                    nstay[i] = 0; home[i] = (c + stride) % B
                else:
                    home[i] = c
            else:
                home[i] = c
            state[i] = (c, s, 0)
            pub[(c, s)] = []
            heapq.heappush(heap, (t + t_ticket, i, "step"))
        elif what == "step":
            c, s, k = state[i]
            ts = can_start(c, s, k, t)
            if ts is None:
                waiters.setdefault((c, s - 1), []).append((k + 2, i, t))
                continue
            wait[i] += ts - t
            te = ts + cost(s, k)
            busy[i] += te - ts
            heapq.heappush(heap, (te, i, "end"))
        else:  # end of a step
            c, s, k = state[i]
            pt = t + delta
            pub[(c, s)].append(pt)
            last = (k == K(s) - 1)
            if last: done[(c, s)] = pt
            # wake waiters
            ws = waiters.get((c, s), [])
            keep = []
            for need, j, tj in ws:
                if last or len(pub[(c, s)]) >= need:
                    heapq.heappush(heap, (max(tj, pt), j, "step"))
                else:
                    keep.append((need, j, tj))
            waiters[(c, s)] = keep
            if last:
                heapq.heappush(heap, (t, i, "ticket"))
            else:
                state[i] = (c, s, k + 1)
                heapq.heappush(heap, (t, i, "step"))
    tot_busy = sum(busy)
    return finish, tot_busy / (ncta * finish)

if __name__ == "__main__":
    for B, ncta in ((16, 32), (16, 48), (16, 64), (16, 96), (64, 128), (64, 148)):
        f, u = simulate(B=B, ncta=ncta)
        print(f"{B} chains, {ncta} CTAs: makespan {f:.0f} step times, utilisation {u:.2f}")
    for st in (4, 16, 10 ** 9):
        f, u = simulate(stay=st)
        print(f"64 chains, 148 CTAs, spare CTAs stay {st} sweeps: makespan {f:.0f}, utilisation {u:.2f}")
    f, u = simulate(policy="prefer")
    print(f"64 chains, 148 CTAs, prefer a chain that can start at once: makespan {f:.0f}, utilisation {u:.2f}")
