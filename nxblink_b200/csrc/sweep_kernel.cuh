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

#define FULL 0xffffffffu
#define NXB_INF (__longlong_as_double(0x7ff0000000000000LL))

namespace nxb {

enum { RC_OK = 0, RC_DEFER = 1, RC_ERROR = 2 };

__constant__ double c_inv_n[64];

// Everything a warp touches in shared memory is addressed as a 32-bit offset from the one dynamic
// shared array, so that the compiler keeps shared-space (LDS/STS) addressing with 32-bit
// arithmetic instead of rebuilding 64-bit generic pointers inside the hot loops.
extern __shared__ __align__(16) unsigned char nxb_smem[];
template <typename T> struct SPtr {
    unsigned off;
    __device__ __forceinline__ T& operator[](int i) const { return reinterpret_cast<T*>(nxb_smem + off)[i]; }
    __device__ __forceinline__ T* ptr() const { return reinterpret_cast<T*>(nxb_smem + off); }
};
template <typename T> __device__ __forceinline__ T* sm_ptr(unsigned off) {
    return reinterpret_cast<T*>(nxb_smem + off);
}

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

// One out-of-line copy: the sweep kernel is far larger than the instruction cache, so
// code size (not call overhead) is what matters for the helpers below.
__device__ __noinline__ NxbU4 philox_call(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                          unsigned k0, unsigned k1) {
    return nxb_philox4x32_10(c0, c1, c2, c3, k0, k1);
}
__device__ __forceinline__ NxbU4 rng(const Ctx& c, unsigned stream, unsigned block) {
    return philox_call(c.unit32, c.sweep, stream, block, c.k0, c.k1);
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
    for (int j = lane; j <= F; j += 32) cstate[j] = 0;
    __syncwarp();
    const SPtr<const unsigned short> ell_info = {c.ell_info};
    const SPtr<const MSG> ell_q = {c.ell_q};
    const SPtr<const MSG> qdt = {(unsigned)p.ml.off_qdt};
    const int W = p.ml.ell_w;
    const bool two = P > 32;
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
            if (cinit && !((MSG)mc[o] > (MSG)0)) { fail(c, NXB_ERR_INFEASIBLE, j); return RC_ERROR; }
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
        if (!(mx > (MSG)0)) { fail(c, NXB_ERR_INFEASIBLE, j); return RC_ERROR; }
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
        if (lane == 0) cstate[pi_] = 1;
        __syncwarp();
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
// TOLERANCE TRACK UPDATES  (raoteh.py:413-432), all K tracks at once, one lane per (unit, track)
// ---------------------------------------------------------------------------
// Exponential gap of the dominating Poisson process, mean gsc / ln 2 (time units).  The logarithm
// is taken in f32 (23-bit uniform).  As an f64 the value has 29 zero low mantissa bits: the unused
// 9 bits of the random word and the track id go there, which makes equal event times rare (they
// are harmless between tracks; a candidate that ties with a retained event of its own track is
// dropped by the walk).
__device__ __forceinline__ double exp_gap(NxbXo& xo, float gsc, int tag) {
    const unsigned w = nxb_xo_next(xo);
    const float f = ((float)(w >> 9) + 0.5f) * (1.0f / 8388608.0f);   // in (0,1), exact
    // MUFU.LG2 has ~2^-22 absolute error, which only matters for f within 1e-6 of 1: clamp so
    // that the gap stays strictly positive (affects one draw in a million by < 1e-7 of a mean gap)
    const double g = (double)(fmaxf(-__log2f(f), 5.9604645e-8f) * gsc);
    // low mantissa bits: the 9 bits of the word that the uniform did not use, then the track id
    return __longlong_as_double(__double_as_longlong(g) | (long long)(((w & 0x1ffu) << 6) | (unsigned)tag));
}

// max / min without the NaN handling of fmax / fmin (event times are never NaN)
__device__ __forceinline__ double dmax2(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin2(double a, double b) { return a < b ? a : b; }

#define CE_EDGE 0x3fffu
#define CE_NEW1 0x4000u     // (survivors only) new state of the surviving event

// One 16-byte record per live event: time, q = l1 / (l0 + l1) of the (OFF, ON) message below the
// event (the backward draw only needs the ratio), edge.  The downward pass reuses the records it
// has consumed for the survivors (e then also carries CE_NEW1), so a track touches only
// ceil(16 * live / 128) lines of the L2-resident pool per sweep.
struct __align__(16) LiveEv { double t; float q; unsigned short e, pad; };
// Accumulator of a node for one track: product of the (OFF, ON) messages of the children folded
// so far (f32, rescaled by powers of two) and the penalty exponent still to be applied to ON.
struct __align__(16) AccRec { double pen; float u0, u1; };

// 2^-(floor(log2 m) + 1) for a positive finite m: exact power-of-two rescaling into [0.5, 1)
// without a division; preserves exact zeros of the other component.
__device__ __forceinline__ float pow2_rescale_f(float m) {
    return __int_as_float((0xfd << 23) - (__float_as_int(m) & (0xff << 23)));
}

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

// ---- lane job: the whole update of ONE tolerance track of ONE unit (upward pass, root draw,
// downward pass).  Job j works on track j % K of unit j / K of a pool of units resident in this
// CTA's shared memory (regions ub0 + u * wl.bytes, global scratch gb0 + u * gscratch_stride), so
// that every lane of the busy warps walks a track; with lane = track only K of 32 lanes would
// (20 of 32 at the p53 shape).
template <bool SPLIT>
__device__ __forceinline__ void blink_job(Ctx& c, int job, int n_jobs, unsigned ub0, unsigned char* gb0) {
    const NxbSweepParams& p = *c.p;
    const NxbWarpLayout& wl = p.wl;
    const NxbModelLayout& ml = p.ml;
    const int E = c.E, P = c.P, K = c.K;
    const int TCAP = p.gcapC;
    const SPtr<const NxbEdgeRec> erec = {(unsigned)ml.off_edgerec};
    const bool has_job = job < n_jobs;
    const int ju = has_job ? job / K : 0;
    const int k = has_job ? job - ju * K : 0;
    const unsigned ub = ub0 + (unsigned)(ju * wl.bytes);             // unit ju's shared-memory region
    unsigned char* const ugb = gb0 + (long long)ju * p.gscratch_stride;
    // live foreground events of this sweep: global (L2-resident) scratch, one fixed region per
    // (unit, track), filled by the upward pass and consumed backwards by the downward pass
    LiveEv* cev = (LiveEv*)(ugb + p.g_off_ctm) + (size_t)k * TCAP;
    const SPtr<int> uhdr = {ub + wl.off_uhdr};
    const SPtr<AccRec> acc = {ub + wl.off_acc};                 // [slot][K]
    const SPtr<unsigned> newtol = {ub + wl.off_newtol};
    const SPtr<int> poff = {ub + wl.off_poff};
    const SPtr<const int> toff = {ub + wl.off_toff};
    const SPtr<const uint4> nrec = {ub + wl.off_nrec};
    const SPtr<unsigned> u_node_tol = {ub + wl.off_node_tol};
    const SPtr<unsigned char> u_node_pri = {ub + wl.off_node_pri};
    const SPtr<double> u_ptm = {ub + wl.off_ptm};
    const SPtr<unsigned> u_pcode = {ub + wl.off_pcode};
    const SPtr<double> u_btm = {ub + wl.off_btm};
    typedef typename std::conditional<SPLIT, unsigned short, unsigned>::type BCode;   // edge in the low 16 bits
    const SPtr<BCode> u_bcode = {ub + wl.off_bcode};

    const bool trk = has_job && uhdr[0] != 0;
    const bool accum_u = trk && uhdr[6] != 0;
    // retained events of track k are the contiguous run [olo, ohi) of the (track, edge, time)
    // sorted list (offsets per track: blink_setup)
    const int olo = trk ? toff[k] : 0, ohi = trk ? toff[k + 1] : 0;
    int bad = 0;
    // per-track sequential stream: candidate gaps, thinning and backward-sampling decisions
    NxbXo xo;
    {
        NxbU4 r = philox_call((unsigned)uhdr[3], (unsigned)uhdr[5], NXB_STREAM_BLK_FFBS | ((unsigned)k << 16), 0u, c.k0, c.k1);
        xo.s0 = r.x; xo.s1 = r.y; xo.s2 = r.z; xo.s3 = r.w | 1u;
    }

    // ---- B1: upward pass.  Every lane walks its own track through the edges in descending
    // index (children before parents), one boundary per iteration: a retained event of the old
    // trajectory, a candidate, a primary event, or the top of the edge.  Candidates are the
    // arrivals of the dominating Poisson process of the edge, generated backwards in time by
    // exponential gaps (poisson.py:22-50 draws count + uniform times; same process), then thinned
    // to the local rate (poisson.py:111-116).  Only accepted ("live") events are stored.
    int cnt = 0;                       // live events pushed by this lane
    {
        // thinning acceptance as integer thresholds: accept iff u32 <= thr (an acceptance
        // probability of exactly 1 becomes 1 - 2^-32)
        const unsigned thr_o0 = (unsigned)fmin(ml.acc_blk[0] * 4294967296.0, 4294967295.0);
        const unsigned thr_o1 = (unsigned)fmin(ml.acc_blk[1] * 4294967296.0, 4294967295.0);
        const unsigned thr_w0 = (unsigned)fmin(ml.acc_blk[2] * 4294967296.0, 4294967295.0);
        const unsigned thr_w1 = (unsigned)fmin(ml.acc_blk[3] * 4294967296.0, 4294967295.0);
        const float a_on = (float)ml.a_on, a_off = (float)ml.a_off;
        int e = trk ? E - 1 : -1;
        int io = ohi - 1;              // cursor into the retained events (descending)
        unsigned inited = 0u;          // live accumulator slots of this lane
        bool fresh = true;
        double tma = 0.0, rate = 0.0, tcur = 0.0, pen = 0.0;
        float u0 = 1.f, u1 = 1.f;      // (OFF, ON) message at the current point of the walk
        double tcand = -NXB_INF;
        float gsc = 0.f;
        int s = 0, ip = trk ? poff[E] - 1 : 0, ip0 = 0, sa_slot = 0, sb_slot = -1;
        unsigned xold = 0u;
        bool gap_due = false;
        while (__any_sync(FULL, e >= 0)) {
            if (e >= 0) {
                // one gap is due: at the bottom of a fresh edge, or after the candidate consumed by the
                // previous step (drawn here so that the walk has a single copy of the gap code)
                bool draw_gap = gap_due;
                gap_due = false;
                if (fresh) {
                    const NxbEdgeRec er = erec[e];
                    sb_slot = er.sb_slot; sa_slot = er.sa_slot;
                    tma = er.tma; rate = er.rate; tcur = er.tmb;
                    u0 = 1.f; u1 = 1.f; pen = 0.0;
                    if (sb_slot >= 0 && ((inited >> sb_slot) & 1u)) {
                        const AccRec a = acc[sb_slot * K + k];
                        u0 = a.u0; u1 = a.u1; pen = a.pen;
                    }
                    const uint4 nr = nrec[e];
                    if (!((nr.x >> k) & 1u)) u1 = 0.f;
                    if (!((nr.y >> k) & 1u)) u0 = 0.f;
                    s = (int)(nr.w & 0xffu);
                    xold = (nr.z >> k) & 1u;
                    // the cursor into the primary events carries over from the edge before
                    // (events are sorted by edge): only the lower bound is new
                    ip0 = (int)(nr.w >> 8);
                    // dominating rate of the edge: 2 * max(on, off) * rate_e per unit time
                    gsc = er.gsc;
                    draw_gap = gsc > 0.f;
                    tcand = draw_gap ? tcur : -NXB_INF;
                    fresh = false;
                }
                if (draw_gap) {
                    tcand -= exp_gap(xo, gsc, k + 1);
                    if (!(tcand > tma)) tcand = -NXB_INF;
                }
                const double tp = (ip >= ip0) ? u_ptm[ip] : -NXB_INF;
                const bool haso = (io >= olo) && ((u_bcode[io] & 0xffffu) == (unsigned)e);
                const double to = haso ? u_btm[io] : -NXB_INF;
                const double tf = dmax2(to, tcand);
                const double tnext = dmax2(dmax2(tp, tf), tma);
                const bool own = (c.cls[s] == k);
                // chunking.py:67-79: own class forces ON; ON pays the rate into class k
                pen += c.qm[s * K + k] * rate * (tcur - tnext);
                if (own) u0 = 0.f;
                tcur = tnext;
                if (tp == -NXB_INF && tf == -NXB_INF) {
                    // top of the edge: fold into the parent's accumulator (one normalisation)
                    const int ai = sa_slot * K + k;
                    AccRec n; n.u0 = u0; n.u1 = u1; n.pen = pen;
                    if ((inited >> sa_slot) & 1u) { const AccRec a = acc[ai]; n.u0 *= a.u0; n.u1 *= a.u1; n.pen += a.pen; }
                    float m2 = fmaxf(n.u0, n.u1);
                    if (!(m2 > 0.f)) { bad = 1; m2 = 1.f; }
                    const float r2 = pow2_rescale_f(m2);
                    n.u0 *= r2; n.u1 *= r2;
                    acc[ai] = n;
                    inited |= 1u << sa_slot;
                    if (sb_slot >= 0) inited &= ~(1u << sb_slot);
                    --e;
                    fresh = true;
                } else if (tf > tp) {
                    bool live = true;
                    if (to > tcand) { xold ^= 1u; --io; }
                    else {
                        const unsigned thr = own ? (xold ? thr_w1 : thr_w0) : (xold ? thr_o1 : thr_o0);
                        // (a candidate that ties with a retained event is dropped: probability ~2^-50)
                        live = (nxb_xo_next(xo) < thr) && (tcand != to);
                        // next arrival of the dominating process (rootward): drawn by the next step;
                        // until then the candidate slot is empty
                        gap_due = true;
                    }
                    if (live) {
                        // (messages are scale-free: with OFF impossible the factor is a pure rescaling
                        // and is dropped, so that a long own-class segment cannot underflow ON to 0)
                        if (pen != 0.0) { if (u0 != 0.f) u1 *= expf((float)(-pen)); pen = 0.0; }
                        float mx = fmaxf(u0, u1);
                        if (!(mx > 0.f)) { bad = 1; mx = 1.f; }
                        const float r = pow2_rescale_f(mx);
                        const float l1 = u1 * r, l0 = u0 * r;        // scale-free message below the event
                        if (cnt < TCAP) {
                            // Ratio by fast division (operands in [0, 2)) -- but a hard constraint must stay
                            // exact: x / x is not guaranteed to be 1 there, and a q of 1 - 2^-24 in an own-class
                            // segment (l0 == 0) let the downward pass switch the class off once in ~10^8
                            // unit-sweeps.  l1 == 0 gives an exact 0.
                            LiveEv ev; ev.t = tnext;
                            ev.q = (l0 == 0.f) ? 1.f : fminf(__fdividef(l1, l0 + l1), 1.f);
                            ev.e = (unsigned short)e; ev.pad = 0;
                            cev[cnt] = ev;
                        } else bad = 2;
                        ++cnt;
                        // uniformized 2x2 matrix of this context (poisson.py:92-96)
                        if (own) { u0 = 0.5f * l0 + 0.5f * l1; u1 = l1; }
                        else {
                            u0 = l0 + a_on * (l1 - l0);
                            u1 = l1 + a_off * (l0 - l1);
                        }
                    }
                } else {
                    s = (u_pcode[ip] >> 16) & 0xffu;     // rootward: the state before the event
                    --ip;
                }
            }
        }
    }
    __syncwarp();
    tick(c, T_B1);
    phase_sync(c);
    // ---- root draw (prior: toymodel.py:17-27 get_blink_distn) ----------------------------------------
    unsigned x = 0u;
    if (trk) {
        const AccRec a = acc[c.slot[0] * K + k];
        double w1 = (double)a.u1 * (a.u0 != 0.f ? exp(-a.pen) : 1.0) * ml.p_on;
        double w0 = (double)a.u0 * ml.p_off;
        if (!(((unsigned)uhdr[8] >> k) & 1u)) w1 = 0.0;
        if (!(((unsigned)uhdr[9] >> k) & 1u)) w0 = 0.0;
        if (c.cls[u_node_pri[0]] == k) w0 = 0.0;
        if (!(w0 + w1 > 0.0)) bad = 1;
        x = (nxb_u32(nxb_xo_next(xo)) * (w0 + w1) < w1) ? 1u : 0u;
    }
    if (trk && x) atomicOr(&newtol[0], 1u << k);
    __syncwarp();

    // ---- B2: downward sampling, edges in ascending index: the live events are read back in
    // reverse; survivors (real transitions) are written from the top of the region downwards.
    // The OFF-time and transition statistics are fused in (summary.py:123-131,234-243).
    const NxbSummaryLayout& sl = p.sl;
    int wtop = cnt - 1;            // survivors overwrite consumed records, from the last one down
    {
        const float a_on_f = (float)ml.a_on, a_off1_f = (float)(1.0 - ml.a_off);
        int e = trk ? 0 : E;
        int ir = cnt - 1;              // read cursor (descending)
        bool fresh = true;
        int s = 0, ip = 0, ipe = 0, vb = 0, n_on = 0, n_off = 0;
        double tmb = 0.0, tcur = 0.0, offtime = 0.0;
        LiveEv pf; pf.t = NXB_INF; pf.q = 0.f; pf.e = 0xffffu; pf.pad = 0;
        if (ir >= 0) pf = cev[ir];
        while (__any_sync(FULL, e < E)) {
            if (e < E) {
                if (fresh) {
                    const NxbEdgeRec er = erec[e];
                    vb = e + 1;
                    const int va = er.va;
                    x = (newtol[va] >> k) & 1u;
                    s = u_node_pri[va];
                    ip = poff[e]; ipe = poff[e + 1];
                    tcur = er.tma; tmb = er.tmb;
                    offtime = 0.0; n_on = 0; n_off = 0;
                    fresh = false;
                }
                const double tp = (ip < ipe) ? u_ptm[ip] : NXB_INF;
                const double tc = ((unsigned)pf.e == (unsigned)e) ? pf.t : NXB_INF;
                const double tnext = dmin2(dmin2(tp, tc), tmb);
                if (!x) offtime += tnext - tcur;
                tcur = tnext;
                if (tc < tp) {
                    const float l1 = pf.q, l0 = 1.f - pf.q;
                    --ir;
                    // (the record at ir is in registers before a survivor may overwrite it: wtop >= ir)
                    if (ir >= 0) pf = cev[ir];
                    else { pf.t = NXB_INF; pf.e = 0xffffu; }
                    const bool own = (c.cls[s] == k);
                    const float p1 = own ? (x ? 1.f : 0.5f) : (x ? a_off1_f : a_on_f);
                    const float w1 = p1 * l1, w0 = (1.f - p1) * l0;
                    if (!(w0 + w1 > 0.f)) bad = 1;
                    const float u = (float)(nxb_xo_next(xo) >> 8) * (1.0f / 16777216.0f);
                    const unsigned xn = (u * (w0 + w1) < w1) ? 1u : 0u;
                    if (xn != x) {
                        LiveEv sv; sv.t = tc; sv.q = 0.f; sv.e = (unsigned short)(e | (xn ? CE_NEW1 : 0u)); sv.pad = 0;
                        cev[wtop] = sv;
                        --wtop;
                        if (xn) ++n_on; else ++n_off;
                        x = xn;
                    }
                } else {
                    // primary event or the bottom of the edge: close the OFF-time segment
                    if (accum_u && offtime > 0.0)
                        atomicAdd(p.dwell + (sl.d_off + (e * P + s) * K + k), offtime);
                    offtime = 0.0;
                    if (tp != NXB_INF) {
                        s = (u_pcode[ip] >> 24) & 0xffu;
                        ++ip;
                    } else {
                        // packed word per edge (low: off->on, high: on->off), updated by halves with
                        // native 32-bit shared atomics; a CTA books far fewer than 2^32 per launch
                        if (accum_u && n_on) atomicAdd((unsigned*)&c.s_off_on[e], (unsigned)n_on);
                        if (accum_u && n_off) atomicAdd((unsigned*)&c.s_off_on[e] + 1, (unsigned)n_off);
                        if (x) atomicOr(&newtol[vb], 1u << k);
                        ++e;
                        fresh = true;
                    }
                }
            }
        }
    }
    if (trk) {
        uhdr[10 + k] = (cnt - 1 - wtop) | (cnt << 16);   // survivors of this track | live events
        if (bad) atomicOr(&uhdr[4], bad);
    }
    tick(c, T_B2);
}

// ---- S3 (the warp that owns the unit): write back the new retained blink events
// SPLIT: the new list goes straight to the unit's slab (the shared-memory copy is not needed again).
template <bool SPLIT>
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
    LiveEv* cev_own = (LiveEv*)(c.gbase + p.g_off_ctm) + (size_t)lane * TCAP;
    const int packed = (lane < K) ? uhdr_own[10 + lane] : 0;
    const int n_mine = packed & 0xffff, n_live = packed >> 16;
    int total;
    int dst = warp_excl_scan(n_mine, lane, total);
    for (int v = lane; v < N; v += 32) c.node_tol[v] = newtol_own[v];
    if (accumulate && lane == 0) {
        int on = __popc(newtol_own[0]);
        atomicAdd(&c.s_root_misc[0], (unsigned long long)on);
        atomicAdd(&c.s_root_misc[1], (unsigned long long)(K - on));
        atomicAdd(&c.s_root_on_cls[c.cls[c.node_pri[0]]], (unsigned long long)on);
        atomicAdd(&c.s_root_misc[2], 1ull);
    }
    double* obtm = c.btm.ptr(); unsigned* obcode = c.bcode.ptr();
    const bool spill = SPLIT || total > wl.capB;
    if (spill) {
        // The sweep is complete and its statistics are booked; only the shared-memory
        // arrays are too small.  Spill the new blink list straight to the HBM slab and
        // let the next capacity tier pick the unit up at the next sweep boundary.
        if (total > p.gcapB) { fail(c, NXB_ERR_STATE_CAP, total); return RC_ERROR; }
        unsigned char* slab = p.state + c.unit * p.state_stride;
        obtm = (double*)(slab + p.so_btm);
        obcode = (unsigned*)(slab + p.so_bcode);
    }
    for (int q = 0; q < n_mine; ++q) {
        const LiveEv sv = cev_own[n_live - 1 - q];      // first survivor (rootmost, earliest) on top
        const unsigned ce = sv.e;
        obtm[dst + q] = sv.t;
        obcode[dst + q] = (unsigned)(ce & CE_EDGE) | ((unsigned)lane << 16) | ((ce & CE_NEW1) ? (1u << 31) : 0u);
    }
    __syncwarp();
    tick(c, T_BW);
    if (spill && !SPLIT) { c.nB = -total; return RC_DEFER; }
    c.nB = total;
    return RC_OK;
}


// TOLERANCE UPDATE of the fused kernel: the warps of a lockstep group set up their own units,
// then carry the group's (unit, track) jobs, then write their own units back.
__device__ __forceinline__ int blink_phase(Ctx& c, bool accumulate, bool active) {
    phase_sync(c);
    blink_setup<false>(c, accumulate, active, sm_ptr<const unsigned>(c.tt_ok.off), sm_ptr<const unsigned>(c.tf_ok.off), nullptr);
    tick(c, T_B0);
    phase_sync(c);            // (a named barrier orders shared memory across the group)
    const int GW = c.bar_threads >> 5;                 // 1 when the warps run independently
    blink_job<false>(c, c.wing * 32 + c.lane, GW * c.K, c.gwbase, c.ggbase);
    phase_sync(c);
    if (!active) return RC_OK;
    return blink_finish<false>(c, accumulate);
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
