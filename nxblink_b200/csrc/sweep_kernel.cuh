// The Rao-Teh uniformization Gibbs sweep of the blink model: device functions of the primary
// update (one warp per unit) and of the tolerance updates (one lane per (unit, track)).
//
// Replaces, for a batch of independent (site, chain) units, the body of
// nxblink/raoteh.py:392-435 (_gen_samples) and everything it calls:
//   nxblink/poisson.py:53-131        candidate ("Poisson") events
//   nxblink/navigation.py:49-148     background-constant context segments
//   nxblink/chunking.py:39-261       chunk trees
//   nxblink/chunking.py:264-289      resample_using_chunk_tree -> sample_history (FFBS)
//   nxblink/summary.py:190-243       Summary.on_sample
// and nxblink/raoteh.py:187-361 (init_tracks) in `init` mode.
//
// B200 mapping (the kernels themselves are in nxb_api.cu):
//   * persistent grids, one CTA per SM; a unit's trajectory is staged in shared memory for the
//     phase (split form: primary-only sweep_kernel, then blink_kernel) or for all requested
//     sweeps (fused sweep_kernel: capacity tiers, split_phases < 0);
//   * primary update: lanes over edges for candidate generation/thinning, lanes over
//     the (<=64) primary states for the chunk-tree FFBS mat-vecs;
//   * tolerance updates: the K tracks are conditionally independent given the primary
//     track (chunking.py:39-107 reads only the primary track), so all K are updated
//     concurrently, every lane walking one track of one unit at its own pace;
//   * candidate events are drawn by thinning one dominating Poisson process per
//     (track, edge): tabulated exp(-lambda) and uniform times for the primary track,
//     exponential gaps inside the walk for the tolerance tracks; the uniformization rate
//     2*max_s(rate) of poisson.py:88-90 is a lookup in an L2-resident 2^K-entry table;
//   * sufficient statistics are reduced on device (shared-memory partials per CTA,
//     fp64/u64 atomics to L2 for the big tables).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "nxb_types.h"
#include "philox.cuh"
#include "nxb_portable.h"
#include "tol_update.cuh"

#define FULL 0xffffffffu

namespace nxb {

enum { RC_OK = 0, RC_DEFER = 1, RC_ERROR = 2,
       RC_STATECAP = 3 };   // the new blink list does not fit the unit's HBM slab; the statistics of the update ARE booked

__constant__ double c_inv_n[64];

struct Ctx {
    const NxbSweepParams* p;
    int N, E, P, K, lane;
    // model tables (shared memory)
    SPtr<const int> parent, slot;
    SPtr<const unsigned short> eperm;   // edge served by lane slot i of the edge-parallel loops (0xffff: none)
    SPtr<const double> tma, len, rate, lam_pri, p0_pri, lam_blk, p0_blk, qval, pi, qm;
    SPtr<const float> gsc_blk;
    SPtr<const unsigned short> qinfo, qptr;
    SPtr<const unsigned char> cls;
    SPtr<const unsigned long long> clsmask;
    unsigned ell_info, ell_q;          // shared-memory offsets of the ELL tables
    // CTA-level statistic partials (shared memory)
    SPtr<unsigned long long> s_off_on, s_on_off, s_root_pri, s_root_misc, s_root_on_cls;
    // per-warp persistent state (shared memory)
    unsigned wbase;           // shared-memory offset of this warp's region
    unsigned char* gbase;     // this warp slot's scratch in global memory
    SPtr<unsigned> node_tol; SPtr<unsigned char> node_pri;
    SPtr<double> ptm; SPtr<unsigned> pcode; SPtr<double> btm; SPtr<unsigned> bcode;
    int nP, nB;
    SPtr<int> boff, poff, tmpa, tmpb, tmpc;
    // site data
    SPtr<const unsigned long long> pri_ok; SPtr<const unsigned> tt_ok, tf_ok;
    // rng
    unsigned k0, k1, unit32, sweep;
    long long unit;       // local unit index
    // phase timers (cycles, lane 0)
    long long t_last;
    // phase lockstep: the warps of a group run the same sub-phase at the same time so that
    // its code is fetched once per SM instead of once per warp (the kernel is several times
    // larger than the instruction cache and instruction fetch was the top limiter)
    int bar_id, bar_threads, nbar;
    int wing;                 // warp index within its lockstep group
    unsigned gwbase;          // shared-memory offset of the group's first warp region
    unsigned char* ggbase;    // global scratch of the group's first warp slot
    unsigned sync_mask;       // which of the sync points actually wait (bit = nbar index)
};

#define NXB_NBAR_PRIMARY 6
#define NXB_NBAR_BLINK 4   /* start, after B0 (setup), after B1, after B2 */

__device__ __forceinline__ void phase_sync(Ctx& c) {
    if (c.bar_threads > 32 && ((c.sync_mask >> c.nbar) & 1u))
        asm volatile("bar.sync %0, %1;" :: "r"(c.bar_id), "r"(c.bar_threads) : "memory");
    ++c.nbar;
}
// after a phase function returned (possibly early), arrive at the barriers it skipped
__device__ __forceinline__ void phase_drain(Ctx& c, int expected) {
    while (c.nbar < expected) phase_sync(c);
}

enum { T_LOAD = 0, T_P0, T_P1, T_P2, T_P3, T_P4, T_P5, T_B0, T_B1, T_B2, T_BW, T_STORE };

__device__ __forceinline__ void tick(Ctx& c, int which) {
    if (c.p->perf) {
        long long now = clock64();
        if (c.lane == 0) atomicAdd(c.p->perf + which, (unsigned long long)(now - c.t_last));
        c.t_last = now;
    }
}

__device__ __forceinline__ NxbU4 rng(const Ctx& c, unsigned stream, unsigned block) {
    return nxb_philox_call(c.unit32, c.sweep, stream, block, c.k0, c.k1);
}

__device__ __noinline__ int warp_excl_scan(int v, int lane, int& total) {
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double y = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += y;
    }
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Draw an index from 64 non-negative weights, two per lane (items lane, lane+32).
// Returns -1 if all weights are zero.  u in (0,1).
__device__ __noinline__ int warp_sample64(double w_lo, double w_hi, double u, int lane) {
    double s_lo = warp_incl_scan(w_lo, lane);
    double tot_lo = __shfl_sync(FULL, s_lo, 31);
    double s_hi = warp_incl_scan(w_hi, lane) + tot_lo;
    double total = __shfl_sync(FULL, s_hi, 31);
    if (!(total > 0.0)) return -1;
    double target = u * total;
    unsigned b_lo = __ballot_sync(FULL, s_lo > target);
    if (b_lo) return __ffs(b_lo) - 1;
    unsigned b_hi = __ballot_sync(FULL, s_hi > target);
    if (b_hi) return 32 + __ffs(b_hi) - 1;
    // rounding: target == total; take the last item with positive weight
    unsigned p_hi = __ballot_sync(FULL, w_hi > 0.0);
    if (p_hi) return 32 + (31 - __clz(p_hi));
    unsigned p_lo = __ballot_sync(FULL, w_lo > 0.0);
    return 31 - __clz(p_lo);
}

__device__ __noinline__ int poisson_inv(double u, double p0, double lam) {
    int n = 0;
    double pr = p0, cdf = p0;
    while (u > cdf && n < 1000) {
        ++n;
        pr *= lam * (n < 64 ? c_inv_n[n] : 1.0 / (double)n);
        cdf += pr;
    }
    return n;
}

// sum_{b : class(b) in mask} q[s][b]   (util.py:24-48 restricted as in poisson.py:83-86)
__device__ __forceinline__ double rsum_masked(const Ctx& c, int s, unsigned mask) {
    double acc = 0.0;
    for (int idx = c.qptr[s]; idx < c.qptr[s + 1]; ++idx) {
        unsigned info = c.qinfo[idx];
        if ((mask >> (info >> 8)) & 1u) acc += c.qval[idx];
    }
    return acc;
}

// fallback when K is too large for the 2^K table
__device__ __noinline__ double rmax_compute(const unsigned short* qptr, const unsigned short* qinfo,
                                            const double* qval, int P, unsigned mask) {
    double m = 0.0;
    for (int s = 0; s < P; ++s) {
        double acc = 0.0;
        for (int idx = qptr[s]; idx < qptr[s + 1]; ++idx)
            if ((mask >> (qinfo[idx] >> 8)) & 1u) acc += qval[idx];
        m = fmax(m, acc);
    }
    return m;
}

// max over ALL source states (poisson.py:88-90 / util.py:13-17)
__device__ __forceinline__ double rmax_lookup(const Ctx& c, unsigned mask) {
    if (c.p->rmax_tab) return __ldg(c.p->rmax_tab + mask);
    return rmax_compute(c.qptr.ptr(), c.qinfo.ptr(), c.qval.ptr(), c.P, mask);
}

__device__ __forceinline__ void fail(const Ctx& c, int code, int detail) {
    if (c.lane == 0) {
        if (atomicCAS(c.p->status, 0, code) == 0) {
            c.p->status[1] = (int)c.unit;
            c.p->status[2] = detail;
            c.p->status[3] = (int)c.sweep;
        }
    }
}

// ---------------------------------------------------------------------------
// index helpers shared by both phases
// ---------------------------------------------------------------------------

// poff[e] = first retained primary event on edge e (events are sorted by edge, time)
__device__ __forceinline__ void build_poff(Ctx& c) {
    const int E = c.E, lane = c.lane;
    for (int e = lane; e <= E; e += 32) c.tmpa[e] = 0;
    __syncwarp();
    for (int i = lane; i < c.nP; i += 32) atomicAdd(&c.tmpa[c.pcode[i] & 0xffffu], 1);
    __syncwarp();
    int run = 0;
    for (int b = 0; b <= E; b += 32) {
        int e = b + lane;
        int v = (e < E) ? c.tmpa[e] : 0;
        int tot;
        int ex = warp_excl_scan(v, lane, tot);
        if (e <= E) c.poff[e] = run + ex;
        run += tot;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// PRIMARY TRACK UPDATE  (raoteh.py:394-411; init: raoteh.py:350-361)
// ---------------------------------------------------------------------------
template <typename MSG>
__device__ __forceinline__ int primary_phase(Ctx& c, bool init, bool accumulate) {
    const NxbSweepParams& p = *c.p;
    const NxbWarpLayout& wl = p.wl;
    const int E = c.E, N = c.N, P = c.P, K = c.K, lane = c.lane;
    const unsigned ALLK = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);

    const SPtr<unsigned short> bord = {c.wbase + wl.off_bord};
    const SPtr<double> rtm = {c.wbase + wl.off_rtm};
    const SPtr<unsigned> rmask = {c.wbase + wl.off_rmask};
    const SPtr<double> ftm = {c.wbase + wl.off_ftm};
    const SPtr<unsigned> fmask = {c.wbase + wl.off_fmask};
    const SPtr<unsigned short> fedge = {c.wbase + wl.off_fedge};
    const SPtr<int> nchunk = {c.wbase + wl.off_nchunk};
    const SPtr<int> jump = {c.wbase + wl.off_jump};
    const SPtr<unsigned short> par = {c.wbase + wl.off_par};
    const SPtr<unsigned> toland = {c.wbase + wl.off_toland};
    const SPtr<unsigned long long> callow = {c.wbase + wl.off_callow};
    const SPtr<unsigned char> cstate = {c.wbase + wl.off_cstate};
    MSG* msg = (MSG*)(c.gbase + p.g_off_msg);          // chunk messages: global (L2)
    const SPtr<MSG> srow = {c.wbase + wl.off_msg};           // staging row of the chunk being pushed up
    const SPtr<int> boff = c.boff, poff = c.poff;
    const SPtr<int> mE = c.tmpa, rbase = c.tmpb, foff = c.tmpc;
    // Edge-parallel loops serve edges in a balanced order (longest edges alone on a lane, short
    // ones paired): the work of an edge is proportional to rate x length, which spans 2 decades.
    const int Epad = (E + 31) & ~31;

    phase_sync(c);
    // ---- P0: edge-major, time-sorted view of the retained blink events ----------
    for (int e = lane; e <= E; e += 32) { c.tmpa[e] = 0; c.tmpb[e] = 0; }
    __syncwarp();
    for (int i = lane; i < c.nB; i += 32) atomicAdd(&c.tmpa[c.bcode[i] & 0xffffu], 1);
    __syncwarp();
    {
        int run = 0;
        for (int b = 0; b <= E; b += 32) {
            int e = b + lane;
            int v = (e < E) ? c.tmpa[e] : 0;
            int tot;
            int ex = warp_excl_scan(v, lane, tot);
            if (e <= E) boff[e] = run + ex;
            run += tot;
        }
    }
    __syncwarp();
    for (int i = lane; i < c.nB; i += 32) {
        int e = c.bcode[i] & 0xffffu;
        int pos = atomicAdd(&c.tmpb[e], 1);
        bord[boff[e] + pos] = (unsigned short)i;
    }
    __syncwarp();
    for (int q = lane; q < Epad; q += 32) {
        const int e = c.eperm[q];
        if (e >= E) continue;
        int a = boff[e], b = boff[e + 1];
        for (int i = a + 1; i < b; ++i) {
            unsigned short id = bord[i];
            double t = c.btm[id];
            int j = i - 1;
            while (j >= a && c.btm[bord[j]] > t) { bord[j + 1] = bord[j]; --j; }
            bord[j + 1] = id;
        }
    }
    __syncwarp();
    build_poff(c);   // uses tmpa
    tick(c, T_P0);
    phase_sync(c);

    // ---- P1: candidates by thinning, merged walk, foreground event regions ------
    int run = 0;
    for (int b = 0; b < Epad; b += 32) {
        const int e = c.eperm[b + lane];
        const bool valid = e < E;
        int n_ret = 0, n_cand = 0;
        bool active = false;
        if (valid) {
            n_ret = poff[e + 1] - poff[e];
            active = (c.len[e] > 0.0) && (c.rate[e] > 0.0);
            if (active) {
                if (init) n_cand = p.ml.diameter;
                else {
                    NxbU4 r = rng(c, NXB_STREAM_PRI_EDGE | (unsigned)e, 0u);
                    n_cand = poisson_inv(nxb_u32(r.x), c.p0_pri[e], c.lam_pri[e]);
                }
            }
        }
        int size = n_ret + n_cand, tot;
        int base = run + warp_excl_scan(size, lane, tot);
        run += tot;
        if (run > wl.capR) return RC_DEFER;          // uniform
        int m = 0;
        if (valid) {
            const int c0 = base + n_ret;
            const double tma = c.tma[e], len = c.len[e];
            for (int i = 0; i < n_cand; ++i) {
                NxbU4 r = rng(c, NXB_STREAM_PRI_EDGE | (unsigned)e, 1u + (unsigned)i);
                double u = nxb_u53(r.x, r.y);
                double t = init ? tma + len * ((1.0 + u) * (1.0 / 3.0)) : tma + len * u;
                // insertion sort, carrying the thinning uniform
                int j = c0 + i - 1;
                while (j >= c0 && rtm[j] > t) { rtm[j + 1] = rtm[j]; rmask[j + 1] = rmask[j]; --j; }
                rtm[j + 1] = t; rmask[j + 1] = r.z;
            }
            // merged walk: blink events (background), retained + candidate events (foreground)
            const int va = c.parent[e];
            unsigned mask = c.node_tol[va];
            int s = init ? 0 : c.node_pri[va];
            int ib = boff[e], ibe = boff[e + 1];
            int ip = poff[e], ipe = poff[e + 1];
            int ic = c0, ice = base + size;
            int w = base;
            while (true) {
                double tb = (ib < ibe) ? c.btm[bord[ib]] : NXB_INF;
                double tp = (ip < ipe) ? c.ptm[ip] : NXB_INF;
                double tc = (ic < ice) ? rtm[ic] : NXB_INF;
                if (tb == NXB_INF && tp == NXB_INF && tc == NXB_INF) break;
                if (tb < tp && tb < tc) {
                    mask ^= 1u << ((c.bcode[bord[ib]] >> 16) & 0xffu);
                    ++ib;
                } else if (tp < tc) {
                    // retained foreground transition: always a chunk boundary
                    rtm[w] = tp; rmask[w] = mask; ++w;
                    s = (c.pcode[ip] >> 24) & 0xffu;
                    ++ip;
                } else {
                    unsigned uth = rmask[ic];
                    ++ic;
                    bool keep = true;
                    unsigned pm = mask;
                    if (init) pm = ALLK;    // raoteh.py:271-276: full-Q uniformization
                    else {
                        // poisson.py:111-116: local rate = omega_ctx - r_s, thinned from the
                        // dominating rate 2*Rmax(all on)
                        double a = (2.0 * rmax_lookup(c, mask) - rsum_masked(c, s, mask))
                                   * p.ml.inv_omega_pri;
                        keep = nxb_u32(uth) < a;
                    }
                    if (keep) { rtm[w] = tc; rmask[w] = pm; ++w; }
                }
            }
            m = w - base;
            mE[e] = m; rbase[e] = base;
        }
    }
    __syncwarp();
    int F = 0;
    {
        int acc = 0;
        for (int b = 0; b <= E; b += 32) {
            int e = b + lane;
            int v = (e < E) ? mE[e] : 0;
            int tot;
            int ex = warp_excl_scan(v, lane, tot);
            if (e <= E) foff[e] = acc + ex;
            acc += tot;
        }
        F = acc;
    }
    if (F > wl.capF || F > p.gcapF) return RC_DEFER;
    __syncwarp();
    tick(c, T_P1);
    phase_sync(c);
    for (int q = lane; q < Epad; q += 32) {
        const int e = c.eperm[q];
        if (e >= E) continue;
        int src = rbase[e], dst = foff[e], m = mE[e];
        for (int i = 0; i < m; ++i) {
            ftm[dst + i] = rtm[src + i];
            fmask[dst + i] = rmask[src + i];
            fedge[dst + i] = (unsigned short)e;
        }
    }
    __syncwarp();       // (the lean layout reuses the candidate arrays from here on: jump[], later the FFBS arrays)
    // ---- P2: chunk structure (chunking.py:158-261) --------------------------------
    // chunk 0 holds the root; foreground event j opens chunk j+1.
    for (int v = lane; v < N; v += 32) {
        int own = -1;
        if (v == 0) own = 0;
        else if (mE[v - 1] > 0) own = foff[v - 1] + mE[v - 1];   // id of the last event on the edge
        nchunk[v] = own;
        jump[v] = (v == 0) ? 0 : c.parent[v - 1];
    }
    __syncwarp();
    for (int it = 0; it < 32; ++it) {
        bool pending = false;
        for (int v = lane; v < N; v += 32) {
            if (nchunk[v] < 0) {
                int a = jump[v];
                int ca = nchunk[a];
                if (ca >= 0) nchunk[v] = ca;
                else { jump[v] = jump[a]; pending = true; }
            }
        }
        __syncwarp();
        if (!__any_sync(FULL, pending)) break;
    }
    for (int j = lane; j <= F; j += 32) { toland[j] = ALLK; callow[j] = ~0ull; }
    for (int j = lane; j < F; j += 32) {
        int e = fedge[j];
        par[j + 1] = (unsigned short)((j == foff[e]) ? nchunk[c.parent[e]] : j);
    }
    __syncwarp();
    for (int q = lane; q < Epad; q += 32) {
        const int e = c.eperm[q];
        if (e >= E) continue;
        const int va = c.parent[e];
        int cur = nchunk[va];
        unsigned mask = c.node_tol[va];
        unsigned acc = mask;
        int ib = boff[e], ibe = boff[e + 1];
        int j = foff[e], je = foff[e] + mE[e];
        while (ib < ibe || j < je) {
            double tb = (ib < ibe) ? c.btm[bord[ib]] : NXB_INF;
            double tf = (j < je) ? ftm[j] : NXB_INF;
            if (tb < tf) {
                mask ^= 1u << ((c.bcode[bord[ib]] >> 16) & 0xffu);
                acc &= mask;
                ++ib;
            } else {
                atomicAnd(&toland[cur], acc);
                cur = j + 1; acc = mask; ++j;
            }
        }
        atomicAnd(&toland[cur], acc);
        atomicAnd(&callow[cur], c.pri_ok[e + 1]);
    }
    if (lane == 0) atomicAnd(&callow[0], c.pri_ok[0]);
    __syncwarp();
    for (int j = lane; j <= F; j += 32) {
        unsigned tm = toland[j];
        unsigned long long ok = 0ull;
        for (int k = 0; k < K; ++k) if ((tm >> k) & 1u) ok |= c.clsmask[k];
        callow[j] &= ok;
    }
    tick(c, T_P2);
    phase_sync(c);
    // ---- P3: upward pass over the chunk tree ------------------------------------------
    // Hoisted, lane-parallel: per-event 1/(2*Rmax(mask)) and the uniforms of the backward pass.
    const SPtr<MSG> finv = {c.wbase + wl.off_finv};
    const SPtr<double> fu = {c.wbase + wl.off_fu};
    {
        int bad = 0;
        for (int j = lane; j <= F; j += 32) {
            if (j < F) {
                double rm = rmax_lookup(c, fmask[j]);
                if (!(rm > 0.0)) bad = 1;
                finv[j] = (MSG)(0.5 / rm);
            }
            NxbU4 r = rng(c, NXB_STREAM_PRI_FFBS, (unsigned)j);
            fu[j] = nxb_u53(r.x, r.y);
        }
        if (__any_sync(FULL, bad)) { fail(c, NXB_ERR_ZERO_RATE, 0); return RC_ERROR; }
    }
    // A chunk message is the product of what its child chunks push up; cstate[] (free until the
    // downward pass) flags the chunks that have received a factor, so that the first child
    // assigns instead of multiplying and an untouched chunk reads as all ones -- no
    // initialisation pass over the L2 scratch and no read of never-written messages.
    const SPtr<const unsigned short> ell_info = {c.ell_info};
    const SPtr<const MSG> ell_q = {c.ell_q};
    const SPtr<const MSG> qdt = {(unsigned)p.ml.off_qdt};
    const int W = p.ml.ell_w;
    const bool two = P > 32;
    // A parent's message is the raw product of its children's factors (each <= 1) until the parent
    // is pushed up itself: a chunk with very many child chunks can underflow to all zeros in f32 on
    // feasible data.  The pass is therefore run again, `careful`, when it meets an all-zero
    // message: then every product is rescaled by a power of two as it is folded (zeros stay
    // exact).  Only a second failure is an infeasible unit (raoteh.py:178-180).
    bool careful = false;
p3_again:
    for (int j = lane; j <= F; j += 32) cstate[j] = 0;
    __syncwarp();
    for (int j = F - 1; j >= 0; --j) {
        const int ci = j + 1, pi_ = par[ci];
        const unsigned fm = fmask[j];
        const unsigned long long allow = callow[ci];
        const MSG inv = finv[j];
        MSG* mc = msg + ci * NXB_PPAD;
        MSG* mp = msg + pi_ * NXB_PPAD;
        const bool cinit = cstate[ci] != 0, pinit = cstate[pi_] != 0;
        if (__popcll(allow) == 1) {
            // One feasible state o in the child chunk (it holds an observed leaf): the message is
            // one-hot, so P . L is column o of the uniformized matrix -- a lookup, not a mat-vec.
            const int o = __ffsll((long long)allow) - 1;
            if (cinit && !((MSG)mc[o] > (MSG)0)) {
                if (!careful) { careful = true; __syncwarp(); goto p3_again; }
                fail(c, NXB_ERR_INFEASIBLE, j); return RC_ERROR;
            }
            MSG qo = 0;
            if (lane < W) {
                const unsigned info = ell_info[lane * NXB_PPAD + o];
                if ((fm >> (info >> 8)) & 1u) qo = ell_q[lane * NXB_PPAD + o];
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) qo += __shfl_xor_sync(FULL, qo, d);     // R_o(mask)
            const MSG c0 = (lane == o) ? (MSG)1 - inv * qo : inv * qdt[o * NXB_PPAD + lane];
            mp[lane] = pinit ? mp[lane] * c0 : c0;
            if (two) {
                const MSG c1 = (lane + 32 == o) ? (MSG)1 - inv * qo : inv * qdt[o * NXB_PPAD + lane + 32];
                mp[lane + 32] = pinit ? mp[lane + 32] * c1 : c1;
            }
            if (__builtin_expect(careful, 0)) {
                MSG n0 = mp[lane], n1 = two ? mp[lane + 32] : (MSG)0, mx = n0 > n1 ? n0 : n1;
                for (int d = 16; d > 0; d >>= 1) { const MSG y = __shfl_xor_sync(FULL, mx, d); mx = y > mx ? y : mx; }
                if (mx > (MSG)0) { const MSG r = TolMath<MSG>::pow2_rescale(mx); mp[lane] = n0 * r; if (two) mp[lane + 32] = n1 * r; }
            }
            if (lane == 0) cstate[pi_] = 1;
            __syncwarp();
            continue;
        }
        // each lane reads back exactly the two entries it wrote itself (program order)
        MSG v0 = ((allow >> lane) & 1ull) ? (cinit ? mc[lane] : (MSG)1) : (MSG)0;
        MSG v1 = (two && ((allow >> (lane + 32)) & 1ull)) ? (cinit ? mc[lane + 32] : (MSG)1) : (MSG)0;
        MSG mx = v0 > v1 ? v0 : v1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { MSG y = __shfl_xor_sync(FULL, mx, o); mx = y > mx ? y : mx; }
        if (!(mx > (MSG)0)) {
            if (!careful) { careful = true; __syncwarp(); goto p3_again; }
            fail(c, NXB_ERR_INFEASIBLE, j); return RC_ERROR;
        }
        const MSG sc = (MSG)1 / mx;
        v0 *= sc; v1 *= sc;
        mc[lane] = v0;                   // kept for the backward pass
        mc[lane + 32] = v1;
        srow[lane] = v0;                 // shared-memory copy for the gather below
        srow[lane + 32] = v1;
        __syncwarp();
        // branch-free sparse mat-vec from the transposed ELL copy of Q: row a of the
        // uniformized matrix is  P[a][a] = 1 - inv*R_a(mask),  P[a][b] = inv*q_ab [class(b) on]
        MSG aq0 = 0, am0 = 0, aq1 = 0, am1 = 0;
#pragma unroll 3
        for (int w = 0; w < W; ++w) {
            const unsigned i0 = ell_info[w * NXB_PPAD + lane];
            const MSG q0 = ((fm >> (i0 >> 8)) & 1u) ? ell_q[w * NXB_PPAD + lane] : (MSG)0;
            aq0 += q0;
            am0 += q0 * srow[i0 & 0xffu];
            if (two) {
                const unsigned i1 = ell_info[w * NXB_PPAD + lane + 32];
                const MSG q1 = ((fm >> (i1 >> 8)) & 1u) ? ell_q[w * NXB_PPAD + lane + 32] : (MSG)0;
                aq1 += q1;
                am1 += q1 * srow[i1 & 0xffu];
            }
        }
        {
            const MSG f0 = v0 + inv * (am0 - aq0 * v0);
            mp[lane] = pinit ? mp[lane] * f0 : f0;
            if (two) {
                const MSG f1 = v1 + inv * (am1 - aq1 * v1);
                mp[lane + 32] = pinit ? mp[lane + 32] * f1 : f1;
            }
        }
        if (__builtin_expect(careful, 0)) {
            MSG n0 = mp[lane], n1 = two ? mp[lane + 32] : (MSG)0, mx2 = n0 > n1 ? n0 : n1;
            for (int d = 16; d > 0; d >>= 1) { const MSG y = __shfl_xor_sync(FULL, mx2, d); mx2 = y > mx2 ? y : mx2; }
            if (mx2 > (MSG)0) { const MSG r = TolMath<MSG>::pow2_rescale(mx2); mp[lane] = n0 * r; if (two) mp[lane + 32] = n1 * r; }
        }
        if (lane == 0) cstate[pi_] = 1;
        __syncwarp();
    }
    if (!careful && cstate[0] != 0) {
        // the root chunk is never pushed up: look at its raw product here (before the lockstep barrier)
        const unsigned long long allow0 = callow[0];
        const bool pos = (lane < P && ((allow0 >> lane) & 1ull) && msg[lane] > (MSG)0) ||
                         (lane + 32 < P && ((allow0 >> (lane + 32)) & 1ull) && msg[lane + 32] > (MSG)0);
        if (!__any_sync(FULL, pos)) { careful = true; goto p3_again; }
    }
    tick(c, T_P3);
    phase_sync(c);
    // ---- P4: root draw and downward sampling ----------------------------------------------
    {
        const unsigned long long allow = callow[0];
        const bool rinit = cstate[0] != 0;        // false only when there is no event at all
        double w0 = (lane < P && ((allow >> lane) & 1ull)) ? (rinit ? (double)msg[lane] : 1.0) * c.pi[lane] : 0.0;
        double w1 = (lane + 32 < P && ((allow >> (lane + 32)) & 1ull))
                    ? (rinit ? (double)msg[lane + 32] : 1.0) * c.pi[lane + 32] : 0.0;
        __syncwarp();
        int s0 = warp_sample64(w0, w1, fu[0], lane);
        if (s0 < 0) { fail(c, NXB_ERR_INFEASIBLE, -1); return RC_ERROR; }
        if (lane == 0) cstate[0] = (unsigned char)s0;
    }
    __syncwarp();
    for (int j = 0; j < F; ++j) {
        const int ci = j + 1;
        {
            const unsigned long long allow = callow[ci];
            if (__popcll(allow) == 1) {      // one feasible state: nothing to sample
                if (lane == 0) cstate[ci] = (unsigned char)(__ffsll((long long)allow) - 1);
                __syncwarp();
                continue;
            }
        }
        const int a = cstate[par[ci]];
        const unsigned fm = fmask[j];
        const MSG inv = finv[j];
        const MSG* mc = msg + ci * NXB_PPAD;
        // lane 0: stay in a (diagonal of the uniformized matrix); lane w+1: ELL entry w of row a
        MSG q = 0;
        int b = a;
        if (lane >= 1 && lane <= W) {
            const unsigned info = ell_info[(lane - 1) * NXB_PPAD + a];
            b = info & 0xffu;
            if ((fm >> (info >> 8)) & 1u) q = ell_q[(lane - 1) * NXB_PPAD + a];
        }
        const MSG wgt = inv * q * mc[b];
        MSG sq = q, sw = wgt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            MSG y = __shfl_up_sync(FULL, sq, d), z = __shfl_up_sync(FULL, sw, d);
            if (lane >= d) { sq += y; sw += z; }
        }
        const MSG totq = __shfl_sync(FULL, sq, 31), totw = __shfl_sync(FULL, sw, 31);
        const MSG w_stay = ((MSG)1 - inv * totq) * mc[a];
        const MSG total = w_stay + totw;
        if (!(total > (MSG)0)) { fail(c, NXB_ERR_INFEASIBLE, j); return RC_ERROR; }
        const MSG target = (MSG)fu[ci] * total;
        int bsel = a;
        if (!(target < w_stay)) {
            unsigned bal = __ballot_sync(FULL, wgt > (MSG)0 && (w_stay + sw) > target);
            if (bal) bsel = __shfl_sync(FULL, b, __ffs(bal) - 1);
            else {
                unsigned pos = __ballot_sync(FULL, wgt > (MSG)0);
                if (pos) bsel = __shfl_sync(FULL, b, 31 - __clz(pos));
            }
        }
        if (lane == 0) cstate[ci] = (unsigned char)bsel;
        __syncwarp();
    }
    tick(c, T_P4);
    phase_sync(c);
    // ---- P5: write back, dropping self transitions (raoteh.py:411) ------------------------------
    int nkeep = 0;
    for (int b = 0; b < F; b += 32) {
        int j = b + lane;
        bool keep = (j < F) && (cstate[par[j + 1]] != cstate[j + 1]);
        nkeep += __popc(__ballot_sync(FULL, keep));
    }
    if (nkeep > wl.capP) return RC_DEFER;     // nothing persistent has been modified yet
    nkeep = 0;
    for (int b = 0; b < F; b += 32) {
        int j = b + lane;
        int sa = 0, sb = 0;
        bool keep = false;
        if (j < F) { sa = cstate[par[j + 1]]; sb = cstate[j + 1]; keep = sa != sb; }
        unsigned bal = __ballot_sync(FULL, keep);
        if (keep) {
            int pos = nkeep + __popc(bal & ((1u << lane) - 1u));
            c.ptm[pos] = ftm[j];
            c.pcode[pos] = (unsigned)fedge[j] | ((unsigned)sa << 16) | ((unsigned)sb << 24);
        }
        nkeep += __popc(bal);
    }
    c.nP = nkeep;
    for (int v = lane; v < N; v += 32) c.node_pri[v] = cstate[nchunk[v]];
    __syncwarp();

    // ---- primary-track statistics (summary.py:214-232 and the primary half of :111-121) ----------
    if (accumulate) {
        build_poff(c);
        const NxbSummaryLayout& sl = p.sl;
        for (int q = lane; q < Epad; q += 32) {
            const int e = c.eperm[q];
            if (e >= E) continue;
            double t = c.tma[e];
            int s = c.node_pri[c.parent[e]];
            for (int i = poff[e]; i < poff[e + 1]; ++i) {
                double te = c.ptm[i];
                atomicAdd(p.dwell + sl.d_T + e * P + s, te - t);
                int sb = (c.pcode[i] >> 24) & 0xffu;
                int idx = c.qptr[s];
                while ((c.qinfo[idx] & 0xffu) != (unsigned)sb) ++idx;
                atomicAdd(p.counts + sl.c_pri_trans + (long long)e * p.ml.nnz + idx, 1ull);
                s = sb; t = te;
            }
            atomicAdd(p.dwell + sl.d_T + e * P + s, c.tma[e] + c.len[e] - t);
        }
        if (lane == 0) atomicAdd(&c.s_root_pri[c.node_pri[0]], 1ull);
    }
    tick(c, T_P5);
    return RC_OK;
}

// ---------------------------------------------------------------------------
// TOLERANCE TRACK UPDATES  (raoteh.py:413-432), all K tracks at once, one lane per (unit, track):
// the per-lane code is in tol_update.cuh; here the warp-level set-up and write-back of a unit.
// ---------------------------------------------------------------------------

// ---- S0 (the warp that owns the unit): index of the primary events, offsets of the retained
// events per track, and the unit header that the lane jobs of other warps read
//   uhdr: [0] active [2] nB [3] unit32 [4] bad [5] sweep [6] accumulate [8] / [9] data masks of the
//   root (tolerance may be ON / OFF), [10 + k] survivors | live events << 16 of track k
// SPLIT (blink_kernel): the data masks and the codes of the retained events are read straight from
// global memory (tt, tf, g_bcode); shared memory keeps only the edge (u16) of every retained event.
template <bool SPLIT>
__device__ __forceinline__ void blink_setup(Ctx& c, bool accumulate, bool active,
        const unsigned* tt, const unsigned* tf, const unsigned* g_bcode) {
    const NxbWarpLayout& wl = c.p->wl;
    const int N = c.N, K = c.K, lane = c.lane;
    const SPtr<int> uhdr_own = {c.wbase + wl.off_uhdr};
    const SPtr<unsigned> newtol_own = {c.wbase + wl.off_newtol};
    const SPtr<int> toff_own = {c.wbase + wl.off_toff};
    if (active) {
        build_poff(c);
        for (int v = lane; v < N; v += 32) newtol_own[v] = 0u;
        // toff[k] = number of retained events on tracks < k (K <= 32: one scan)
        if (lane <= K) toff_own[lane] = 0;
        if (K == 32 && lane == 0) toff_own[32] = 0;
        __syncwarp();
        if (SPLIT) {
            unsigned short* bedge = sm_ptr<unsigned short>(c.bcode.off);
            for (int i = lane; i < c.nB; i += 32) {
                const unsigned code = __ldcs(g_bcode + i);
                bedge[i] = (unsigned short)(code & 0xffffu);
                atomicAdd(&toff_own[(code >> 16) & 0xffu], 1);
            }
        } else {
            for (int i = lane; i < c.nB; i += 32) atomicAdd(&toff_own[(c.bcode[i] >> 16) & 0xffu], 1);
        }
        __syncwarp();
        int total;
        const int cnt = lane < K ? toff_own[lane] : 0;
        const int ex = warp_excl_scan(cnt, lane, total);
        __syncwarp();
        if (lane < K) toff_own[lane] = ex;
        if (lane == 0) toff_own[K] = total;
    }
    if (active) {
        // one 16-byte record per edge for the upward walk: data masks, old tolerance states and
        // primary state of the lower node, and the index of the edge's first primary event
        const SPtr<uint4> nrec_own = {c.wbase + wl.off_nrec};
        for (int e = lane; e < c.E; e += 32)
            nrec_own[e] = make_uint4(tt[e + 1], tf[e + 1], c.node_tol[e + 1],
                                     (unsigned)c.node_pri[e + 1] | ((unsigned)c.poff[e] << 8));
    }
    if (lane == 0) {
        uhdr_own[0] = active ? 1 : 0; uhdr_own[2] = c.nB; uhdr_own[3] = (int)c.unit32; uhdr_own[4] = 0;
        uhdr_own[5] = (int)c.sweep; uhdr_own[6] = accumulate ? 1 : 0;
        if (active) { uhdr_own[8] = (int)tt[0]; uhdr_own[9] = (int)tf[0]; }
    }
    __syncwarp();
}

// ---- lane job: the whole update of ONE tolerance track of ONE unit.  Job j works on track j % K
// of unit j / K of a pool of units resident in this CTA's shared memory (regions ub0 + u * wl.bytes,
// global scratch gb0 + u * gscratch_stride), so that every lane of the busy warps carries a track;
// with lane = track only K of 32 lanes would (20 of 32 at the p53 shape).
template <typename MSG, bool SPLIT>
__device__ __forceinline__ void blink_job(Ctx& c, int job, int n_jobs, unsigned ub0, unsigned char* gb0) {
    TolCtx tc;
    tc.p = c.p; tc.E = c.E; tc.P = c.P; tc.K = c.K;
    tc.cls = c.cls; tc.qm = c.qm; tc.slot = c.slot; tc.s_off_on = c.s_off_on; tc.k0 = c.k0; tc.k1 = c.k1;
    TolLane<MSG> L;
    TolUnit<MSG, SPLIT> u;
    tol_begin<MSG, SPLIT>(tc, L, u, job, n_jobs, ub0, gb0);
    const unsigned x_root = tol_up<MSG, SPLIT>(tc, L, u);
    __syncwarp();
    tick(c, T_B1);
    phase_sync(c);
    tol_down<MSG, SPLIT>(tc, L, u, x_root);
    tick(c, T_B2);
}

// ---- S3 (the warp that owns the unit): write back the new retained blink events
// SPLIT: the new list goes straight to the unit's slab (the shared-memory copy is not needed again).
template <typename MSG, bool SPLIT>
__device__ __forceinline__ int blink_finish(Ctx& c, bool accumulate) {
    const NxbSweepParams& p = *c.p;
    const NxbWarpLayout& wl = p.wl;
    const int N = c.N, K = c.K, lane = c.lane;
    const int TCAP = p.gcapC;
    const SPtr<int> uhdr_own = {c.wbase + wl.off_uhdr};
    const SPtr<unsigned> newtol_own = {c.wbase + wl.off_newtol};
    // (track, edge, time) order
    {
        const int flags = uhdr_own[4];
        if (flags & 2) { fail(c, NXB_ERR_SCRATCH_CAP, 0); return RC_ERROR; }
        if (flags & 1) { fail(c, NXB_ERR_INFEASIBLE, -3); return RC_ERROR; }
    }
    LiveEv<MSG>* cev_own = (LiveEv<MSG>*)(c.gbase + p.g_off_ctm) + (size_t)lane * TCAP;
    const int packed = (lane < K) ? uhdr_own[10 + lane] : 0;
    const int n_mine = packed & 0xffff, n_live = packed >> 16;
    int total;
    int dst = warp_excl_scan(n_mine, lane, total);
    if (tol_event_driven(c.E)) {
        // The event-driven downward pass (tol_down_ev) sampled the node states only below the edges
        // on which a track has live events.  hasev[e] = those tracks, from the evm words the lanes
        // left in their dead accumulator records; an event-free edge hands the upper node's state
        // down, so every node takes, per track, the bits of its nearest ancestor that has them
        // (lane = node, at most depth steps); then the OFF dwell of the tracks that are OFF
        // throughout an event-free edge (summary.py:123-131), lane = edge.  (tol_fill_scalar is the
        // scalar statement of the same, checked by the host simulation.)
        const int E = c.E, P = c.P;
        const unsigned ALLK = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);
        const SPtr<unsigned> hasev = {c.wbase + wl.off_nrec};          // (the per-edge records of the upward walk are dead)
        const SPtr<const AccRec<MSG> > acc = {c.wbase + wl.off_acc};
        const SPtr<const NxbEdgeRec> erec = {(unsigned)p.ml.off_edgerec};
        for (int e = lane; e < E; e += 32) hasev[e] = 0u;
        __syncwarp();
        if (lane < K) {
            unsigned long long m = *reinterpret_cast<const unsigned long long*>(&acc[lane]);
            while (m) {
                const int e = __ffsll((long long)m) - 1;
                m &= m - 1ull;
                atomicOr(&hasev[e], 1u << lane);
            }
        }
        __syncwarp();
        for (int v = lane; v < N; v += 32) {
            unsigned need = (v == 0) ? 0u : (ALLK & ~hasev[v - 1]);
            unsigned bits = newtol_own[v] & ~need;
            int cur = v;
            while (need) {
                cur = erec[cur - 1].va;
                const unsigned known = (cur == 0) ? ALLK : hasev[cur - 1];
                bits |= newtol_own[cur] & need & known;
                need &= ~known;
            }
            c.node_tol[v] = bits;
        }
        __syncwarp();
        if (accumulate) {
            const NxbSummaryLayout& sl = p.sl;
            for (int e = lane; e < E; e += 32) {
                const unsigned offm = ~c.node_tol[e + 1] & ~hasev[e] & ALLK;
                if (!offm) continue;
                const NxbEdgeRec er = erec[e];
                double t = er.tma;
                int s = c.node_pri[er.va];
                const int i1 = c.poff[e + 1];
                for (int i = c.poff[e]; i <= i1; ++i) {
                    const bool last = i == i1;
                    const double te = last ? er.tmb : c.ptm[i];
                    const double dt = te - t;
                    if (dt > 0.0) {
                        double* const cell = p.dwell + (sl.d_off + (e * P + s) * K);
                        for (unsigned m = offm; m; m &= m - 1u) atomicAdd(cell + (__ffs(m) - 1), dt);
                    }
                    if (!last) { s = (int)((c.pcode[i] >> 24) & 0xffu); t = te; }
                }
            }
        }
    } else {
        for (int v = lane; v < N; v += 32) c.node_tol[v] = newtol_own[v];
    }
    if (accumulate && lane == 0) {
        int on = __popc(newtol_own[0]);
        atomicAdd(&c.s_root_misc[0], (unsigned long long)on);
        atomicAdd(&c.s_root_misc[1], (unsigned long long)(K - on));
        atomicAdd(&c.s_root_on_cls[c.cls[c.node_pri[0]]], (unsigned long long)on);
        atomicAdd(&c.s_root_misc[2], 1ull);
    }
    double* obtm = c.btm.ptr(); unsigned* obcode = c.bcode.ptr();
    const bool spill = SPLIT || total > wl.capB;
    // SPLIT: the new list is assembled in shared memory -- times over the old ones, codes in the dead
    // accumulator records -- and leaves as two bulk-async copies (cp.async.bulk, shared -> global) that
    // run while the warp finishes its next unit; blink_kernel waits for them before the group's
    // shared memory is reused.  Lists that do not fit are stored directly.
    bool bulk = false;
    if (SPLIT) {
        const int acc_bytes = (int)sizeof(AccRec<MSG>) * K * p.ml.nslots;
        bulk = p.bulk_store && total > 0 && total <= wl.capB && 4 * ((total + 3) & ~3) <= acc_bytes;
    }
    unsigned char* const slab_out = p.state + c.unit * p.state_stride;
    if (spill) {
        // The sweep is complete and its statistics are booked; only the shared-memory
        // arrays are too small.  Spill the new blink list straight to the HBM slab and
        // let the next capacity tier pick the unit up at the next sweep boundary.
        if (total > p.gcapB) { fail(c, NXB_ERR_STATE_CAP, total); return RC_STATECAP; }
        if (bulk) obcode = sm_ptr<unsigned>(c.wbase + wl.off_acc);
        else {
            obtm = (double*)(slab_out + p.so_btm);
            obcode = (unsigned*)(slab_out + p.so_bcode);
        }
    }
    // (two pool records in flight per lane: the copy is bound by the L2 latency of its loads)
    for (int q0 = 0; q0 < n_mine; q0 += 2) {
        LiveEv<MSG> sv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (q0 + j < n_mine) sv[j] = nxb_ev_load(cev_own + (n_live - 1 - (q0 + j)));      // first survivor (rootmost, earliest) on top
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (q0 + j < n_mine) {
                const unsigned ce = sv[j].e;
                obtm[dst + q0 + j] = sv[j].t;
                obcode[dst + q0 + j] = (unsigned)(ce & CE_EDGE) | ((unsigned)lane << 16) | ((ce & CE_NEW1) ? (1u << 31) : 0u);
            }
    }
    __syncwarp();
    if (bulk && lane == 0) {
        // generic-proxy writes to shared memory -> visible to the async proxy, then the two copies
        const unsigned s_tm = (unsigned)__cvta_generic_to_shared(obtm);
        const unsigned s_cd = (unsigned)__cvta_generic_to_shared(obcode);
        const unsigned b_tm = (unsigned)((8 * total + 15) & ~15), b_cd = (unsigned)((4 * total + 15) & ~15);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(slab_out + p.so_btm), "r"(s_tm), "r"(b_tm) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(slab_out + p.so_bcode), "r"(s_cd), "r"(b_cd) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // The pool records of this sweep are dead now.  They live in the L2-resident scratch; tell L2
    // that their lines need no write-back (discard.global.L2), or every sweep writes ~3 KB per
    // unit of dead records to HBM when the lines are evicted.
    if (lane < K && p.discard_pool) {
        const char* b = (const char*)cev_own;
        for (int off = 0; off < n_live * (int)sizeof(LiveEv<MSG>); off += 128)
            asm volatile("discard.global.L2 [%0], 128;" :: "l"(b + off) : "memory");
    }
    tick(c, T_BW);
    if (spill && !SPLIT) { c.nB = -total; return RC_DEFER; }
    c.nB = total;
    return RC_OK;
}


// TOLERANCE UPDATE of the fused kernel: the warps of a lockstep group set up their own units,
// then carry the group's (unit, track) jobs, then write their own units back.
template <typename MSG>
__device__ __forceinline__ int blink_phase(Ctx& c, bool accumulate, bool active) {
    phase_sync(c);
    blink_setup<false>(c, accumulate, active, sm_ptr<const unsigned>(c.tt_ok.off), sm_ptr<const unsigned>(c.tf_ok.off), nullptr);
    tick(c, T_B0);
    phase_sync(c);            // (a named barrier orders shared memory across the group)
    const int GW = c.bar_threads >> 5;                 // 1 when the warps run independently
    blink_job<MSG, false>(c, c.wing * 32 + c.lane, GW * c.K, c.gwbase, c.ggbase);
    phase_sync(c);
    if (!active) return RC_OK;
    return blink_finish<MSG, false>(c, accumulate);
}

// ---------------------------------------------------------------------------
// initial tolerance histories  (raoteh.py:187-229)
// ---------------------------------------------------------------------------
__device__ __forceinline__ int init_blink(Ctx& c) {
    const int E = c.E, N = c.N, K = c.K, lane = c.lane;
    const unsigned ALLK = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);
    int bad = 0;
    for (int v = lane; v < N; v += 32) {
        unsigned t = c.tt_ok[v] & ALLK, f = c.tf_ok[v] & ALLK;
        if ((t | f) != ALLK) bad = 1;          // raoteh.py:201
        c.node_tol[v] = t;                     // ON wherever the data allow it
    }
    if (__any_sync(FULL, bad)) { fail(c, NXB_ERR_INFEASIBLE, -4); return RC_ERROR; }
    __syncwarp();
    const bool trk = lane < K;
    int cnt = 0;
    if (trk) for (int e = 0; e < E; ++e) {
        if (!(c.len[e] > 0.0 && c.rate[e] > 0.0)) continue;
        cnt += !((c.node_tol[c.parent[e]] >> lane) & 1u);
        cnt += !((c.node_tol[e + 1] >> lane) & 1u);
    }
    int total;
    int dst = warp_excl_scan(cnt, lane, total);
    if (total > c.p->wl.capB) return RC_DEFER;
    if (trk) for (int e = 0; e < E; ++e) {
        if (!(c.len[e] > 0.0 && c.rate[e] > 0.0)) continue;
        unsigned sa = (c.node_tol[c.parent[e]] >> lane) & 1u;
        unsigned sb = (c.node_tol[e + 1] >> lane) & 1u;
        if (sa && sb) continue;
        NxbU4 r = rng(c, NXB_STREAM_INIT_BLK | ((unsigned)lane << 16) | (unsigned)e, 0u);
        if (!sa) {
            c.btm[dst] = c.tma[e] + c.len[e] * (nxb_u53(r.x, r.y) * (1.0 / 3.0));
            c.bcode[dst] = (unsigned)e | ((unsigned)lane << 16) | (1u << 31);
            ++dst;
        }
        if (!sb) {
            c.btm[dst] = c.tma[e] + c.len[e] * ((2.0 + nxb_u53(r.z, r.w)) * (1.0 / 3.0));
            c.bcode[dst] = (unsigned)e | ((unsigned)lane << 16);
            ++dst;
        }
    }
    c.nB = total;
    c.nP = 0;
    __syncwarp();
    return RC_OK;
}

// ---------------------------------------------------------------------------
// trajectory invariants (the checks of navigation.py:39-44,121-135, chunking.py:96-100,
// summary.py:243, plus the model constraints of nxblink/compound.py:23-40)
// ---------------------------------------------------------------------------
__device__ __forceinline__ int validate_unit(Ctx& c) {
    const int E = c.E, N = c.N, K = c.K, lane = c.lane;
    const SPtr<unsigned short> bord = {c.wbase + c.p->wl.off_bord};
    int bad = 0;
    // primary events sorted by (edge, time); blink events by (track, edge, time)
    for (int i = lane; i + 1 < c.nP; i += 32) {
        unsigned ea = c.pcode[i] & 0xffffu, eb = c.pcode[i + 1] & 0xffffu;
        if (ea > eb || (ea == eb && !(c.ptm[i] < c.ptm[i + 1]))) bad = 1;
    }
    for (int i = lane; i + 1 < c.nB; i += 32) {
        unsigned ka = c.bcode[i] & 0xffffffu, kb = c.bcode[i + 1] & 0xffffffu;
        if (ka > kb || (ka == kb && !(c.btm[i] < c.btm[i + 1]))) bad = 2;
    }
    if (__any_sync(FULL, bad)) { fail(c, NXB_ERR_INVARIANT, __reduce_max_sync(FULL, bad)); return RC_ERROR; }
    // edge-major view of the blink events
    for (int e = lane; e <= E; e += 32) { c.tmpa[e] = 0; c.tmpb[e] = 0; }
    __syncwarp();
    for (int i = lane; i < c.nB; i += 32) atomicAdd(&c.tmpa[c.bcode[i] & 0xffffu], 1);
    __syncwarp();
    {
        int run = 0;
        for (int b = 0; b <= E; b += 32) {
            int e = b + lane, tot;
            int v = (e < E) ? c.tmpa[e] : 0;
            int ex = warp_excl_scan(v, lane, tot);
            if (e <= E) c.boff[e] = run + ex;
            run += tot;
        }
    }
    __syncwarp();
    for (int i = lane; i < c.nB; i += 32) {
        int e = c.bcode[i] & 0xffffu;
        bord[c.boff[e] + atomicAdd(&c.tmpb[e], 1)] = (unsigned short)i;
    }
    __syncwarp();
    for (int e = lane; e < E; e += 32) {
        int a = c.boff[e], b = c.boff[e + 1];
        for (int i = a + 1; i < b; ++i) {
            unsigned short id = bord[i];
            double t = c.btm[id];
            int j = i - 1;
            while (j >= a && c.btm[bord[j]] > t) { bord[j + 1] = bord[j]; --j; }
            bord[j + 1] = id;
        }
    }
    __syncwarp();
    build_poff(c);
    for (int v = lane; v < N; v += 32) {
        int s = c.node_pri[v];
        unsigned m = c.node_tol[v];
        if (!((c.pri_ok[v] >> s) & 1ull)) bad = 3;
        if ((m & ~c.tt_ok[v]) || (~m & ((K == 32) ? ~0u : ((1u << K) - 1u)) & ~c.tf_ok[v])) bad = 4;
        if (!((m >> c.cls[s]) & 1u)) bad = 5;
    }
    for (int e = lane; e < E; e += 32) {
        const int va = c.parent[e];
        const double t0 = c.tma[e], t1 = t0 + c.len[e];
        unsigned mask = c.node_tol[va];
        int s = c.node_pri[va];
        int ib = c.boff[e], ibe = c.boff[e + 1], ip = c.poff[e], ipe = c.poff[e + 1];
        if ((ib < ibe || ip < ipe) && !(c.len[e] > 0.0 && c.rate[e] > 0.0)) bad = 6;
        while (ib < ibe || ip < ipe) {
            double tb = (ib < ibe) ? c.btm[bord[ib]] : NXB_INF;
            double tp = (ip < ipe) ? c.ptm[ip] : NXB_INF;
            double t = fmin(tb, tp);
            if (!(t > t0 && t < t1)) bad = 7;
            if (tb < tp) {
                unsigned code = c.bcode[bord[ib]];
                unsigned k = (code >> 16) & 0xffu, ns = code >> 31;
                if (((mask >> k) & 1u) == ns) bad = 8;            // self / incompatible transition
                if (!ns && c.cls[s] == k) bad = 9;                 // own class may not switch off
                mask ^= 1u << k;
                ++ib;
            } else {
                unsigned code = c.pcode[ip];
                int sa = (code >> 16) & 0xffu, sb = (code >> 24) & 0xffu;
                if (sa != s || sa == sb) bad = 10;
                if (!((mask >> c.cls[sb]) & 1u)) bad = 11;         // target class must be ON
                bool found = false;
                for (int idx = c.qptr[sa]; idx < c.qptr[sa + 1]; ++idx)
                    if ((c.qinfo[idx] & 0xffu) == (unsigned)sb) found = true;
                if (!found) bad = 12;
                s = sb;
                ++ip;
            }
        }
        if (s != c.node_pri[e + 1]) bad = 13;
        if (mask != c.node_tol[e + 1]) bad = 14;
    }
    if (__any_sync(FULL, bad)) { fail(c, NXB_ERR_INVARIANT, __reduce_max_sync(FULL, bad)); return RC_ERROR; }
    return RC_OK;
}

}  // namespace nxb
