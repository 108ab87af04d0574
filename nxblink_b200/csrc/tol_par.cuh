// TOLERANCE TRACK UPDATES (raoteh.py:413-432) of one unit, event-parallel form: one lane per EVENT
// in every heavy phase, one lane per track only in two short tree passes.  Replaces, per track,
// poisson.py:176-199 (candidate events), chunking.py:39-155 (blinking chunk tree),
// chunking.py:264-289 -> nxmctree sample_history (FFBS) and the tolerance half of
// summary.py:190-243.
//
// The 2-state forward-filtering / backward-sampling recursion is associative along an edge:
// between two consecutive live events the upward message is multiplied by a diagonal matrix
// diag(OFF feasible?, exp(-penalty)) (chunking.py:67-79) and at a live event by the uniformized
// 2x2 matrix of its context (poisson.py:92-96); the backward draw at an event, with its uniform
// fixed in advance, is a map {0,1} -> {0,1} and maps compose.  So a sweep of the K tracks of a
// unit is:
//   G  candidates: the dominating Poisson processes of all K tracks are laid out on ONE axis
//      (track-major, edges in index order, rate-length units) cut into NXB_TP_NSLICE equal slices;
//      lane = slice: Poisson count by table inversion, then the sorted positions of the slice by
//      the descending order-statistics recursion (v *= u^(1/j)) -> candidates come out SORTED
//      by (track, edge, time) at known ranks, no sort and no atomics (poisson.py:22-50: count +
//      uniform times per segment; same process);
//   M  merge with the retained events (already sorted): one binary search per retained event;
//   T  thinning, lane = candidate (poisson.py:111-116): context (own class?, old state) by
//      direct lookup, accepted candidates and retained events go to their merged rank;
//   B  lane = live event: penalty integral to the previous live event, 2x2 matrix;
//   S  segmented suffix products of the matrices per (track, edge) bucket (warp shuffles);
//   U  lane = track: 48 uniform steps up the tree with one 2x2 mat-vec per non-empty bucket; root draw;
//   F  lane = live event: message below the event, the two possible outcomes of its backward draw
//      (a 2-bit map), segmented prefix composition of the maps per bucket;
//   D  one lane: new node states, bit-parallel over the tracks;
//   W  lane = live event: state above / below, survivors (real transitions) written to the slab in
//      (track, edge, time) order, statistics (summary.py:123-131,234-243).
// Every random number is a function of (seed, unit, sweep, slice | retained index | track): results
// do not depend on the launch geometry, the capacity tier or the number of warps per unit.
//
// Templated on the message type MSG (float: fast path; double: the reference's precision end to
// end, chunking.py:32-36,264-289) and on WPU, the warps that share one unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "nxb_types.h"
#include "philox.cuh"
#include "nxb_portable.h"
#include "tol_update.cuh"     // AccRec, TolMath

namespace nxb {

#define TPK_EDGE_MASK 0x1fffu        /* 13 bits: edge */
#define TPK_TRACK_SHIFT 13           /* 5 bits: track */
#define TPK_BKT_MASK 0x3ffffu        /* (track, edge) bucket */
#define TPK_OWN (1u << 18)           /* primary state at the event belongs to this track's class */
#define TPK_CAND (1u << 19)          /* candidate (else: retained event of the old trajectory) */
#define TPK_X (1u << 20)             /* state of the OLD trajectory just above the event */
#define TPK_LAST (1u << 21)          /* last event of its bucket */
#define TPK_AUX_SHIFT 22             /* 10 bits: primary-event index within the edge (T, B); last event of a bucket: distance to the first (S..) */
#define TPK_AUX_MAX 1023u
#define TPS_DEAD 0x80000000u         /* staging record: candidate outside its edge (rounding) */

template <typename MSG> struct TpT;
template <> struct TpT<float> {
    typedef float4 M4;
    typedef unsigned UA;                           // uniform of the backward draw: 24 bits; later the map bits
    struct __align__(16) Stg { double t; unsigned a, w; };
    static NXB_DEV float4 make(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
};
template <> struct TpT<double> {
    struct __align__(16) M4 { double x, y, z, w; };
    typedef unsigned long long UA;                 // 53 bits
    struct __align__(16) Stg { double t; unsigned a, w; unsigned w2, pad0; double pad1; };
    static NXB_DEV M4 make(double x, double y, double z, double w) { M4 m; m.x = x; m.y = y; m.z = z; m.w = w; return m; }
};

// the warps that share a unit
template <int WPU> struct TpGroup {
    int li, lane, w, bar;
    NXB_DEV void sync() const {
        if (WPU == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" :: "r"(bar), "n"(32 * WPU) : "memory");
    }
};

struct TpCtx {
    const NxbSweepParams* p;
    int E, P, K;
    SPtr<const unsigned char> cls;
    SPtr<const double> qm;
    SPtr<const int> slot;
    SPtr<const NxbEdgeRec> erec;
    SPtr<const NxbGenRec> grec;
    SPtr<const unsigned> pcdf;
    SPtr<const unsigned short> etab;
    SPtr<const unsigned short> lvl_off, lvl_node, ch_off, ch_node, dep_off, dep_edge;
    SPtr<const short> lslot;
    SPtr<unsigned long long> s_off_on, s_root_misc, s_root_on_cls;
    unsigned k0, k1;
};

template <typename MSG> struct TpUnit {
    typedef typename TpT<MSG>::M4 M4;
    typedef typename TpT<MSG>::UA UA;
    typedef typename TpT<MSG>::Stg Stg;
    SPtr<double> ptm; SPtr<unsigned> pcode; SPtr<double> btm; SPtr<unsigned> bcode;
    SPtr<unsigned> node_tol, newtol; SPtr<unsigned char> node_pri;
    SPtr<uint4> nrec;
    SPtr<unsigned short> ins;
    SPtr<int> toff;
    SPtr<unsigned> f0, f1, hasev;
    SPtr<AccRec<MSG> > acc;
    SPtr<double> T; SPtr<unsigned> key; SPtr<UA> aux; SPtr<M4> q; SPtr<Stg> stg;
    SPtr<int> x;
    int nP, nB;
    unsigned unit32, sweep;
    NXB_DEV void bind(unsigned ub, const NxbTolLayout& tl) {
        ptm.off = ub + tl.off_ptm; pcode.off = ub + tl.off_pcode; btm.off = ub + tl.off_btm; bcode.off = ub + tl.off_bcode;
        node_tol.off = ub + tl.off_node_tol; newtol.off = ub + tl.off_newtol; node_pri.off = ub + tl.off_node_pri;
        nrec.off = ub + tl.off_nrec; ins.off = ub + tl.off_ins; toff.off = ub + tl.off_toff;
        f0.off = ub + tl.off_f0; f1.off = ub + tl.off_f1; hasev.off = ub + tl.off_hasev;
        acc.off = ub + tl.off_acc; T.off = ub + tl.off_T; key.off = ub + tl.off_key; aux.off = ub + tl.off_aux;
        q.off = ub + tl.off_q; stg.off = ub + tl.off_q; x.off = ub + tl.off_x;
    }
};

// ---- small helpers ---------------------------------------------------------------------------
NXB_DEV int tp_warp_excl_scan(int v, int lane, int& total) {
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}
// exclusive scan over the lanes of the unit group (every lane calls); `slot` alternates between calls
template <int WPU>
NXB_DEV int tp_group_excl_scan(const TpGroup<WPU>& g, int v, int& total, SPtr<int> xch, int& slot) {
    int tw;
    const int ex = tp_warp_excl_scan(v, g.lane, tw);
    if (WPU == 1) { total = tw; return ex; }
    const int s0 = (slot & 1) * WPU;
    ++slot;
    if (g.lane == 0) xch[s0 + g.w] = tw;
    g.sync();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < WPU; ++i) { const int t = xch[s0 + i]; if (i < g.w) base += t; tot += t; }
    total = tot;
    return base + ex;
}
// same for a flag, by ballot
template <int WPU>
NXB_DEV int tp_group_excl_count(const TpGroup<WPU>& g, bool f, int& total, SPtr<int> xch, int& slot) {
    const unsigned b = __ballot_sync(0xffffffffu, f);
    const int ex = __popc(b & ((1u << g.lane) - 1u));
    const int tw = __popc(b);
    if (WPU == 1) { total = tw; return ex; }
    const int s0 = (slot & 1) * WPU;
    ++slot;
    if (g.lane == 0) xch[s0 + g.w] = tw;
    g.sync();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < WPU; ++i) { const int t = xch[s0 + i]; if (i < g.w) base += t; tot += t; }
    total = tot;
    return base + ex;
}

template <typename M4> NXB_DEV M4 tp_shfl_down(const M4& m, int d);
template <> NXB_DEV float4 tp_shfl_down<float4>(const float4& m, int d) {
    return make_float4(__shfl_down_sync(0xffffffffu, m.x, d), __shfl_down_sync(0xffffffffu, m.y, d),
                       __shfl_down_sync(0xffffffffu, m.z, d), __shfl_down_sync(0xffffffffu, m.w, d));
}
template <> NXB_DEV TpT<double>::M4 tp_shfl_down<TpT<double>::M4>(const TpT<double>::M4& m, int d) {
    TpT<double>::M4 r;
    r.x = __shfl_down_sync(0xffffffffu, m.x, d); r.y = __shfl_down_sync(0xffffffffu, m.y, d);
    r.z = __shfl_down_sync(0xffffffffu, m.z, d); r.w = __shfl_down_sync(0xffffffffu, m.w, d);
    return r;
}
// (a applied after b: a is the rootward factor)   rows: state above, columns: state below
template <typename MSG, typename M4> NXB_DEV M4 tp_mul(const M4& a, const M4& b) {
    return TpT<MSG>::make(a.x * b.x + a.y * b.z, a.x * b.y + a.y * b.w, a.z * b.x + a.w * b.z, a.z * b.y + a.w * b.w);
}
template <typename MSG, typename M4> NXB_DEV void tp_rescale(M4& m) {
    MSG mx = m.x > m.y ? m.x : m.y;
    const MSG my = m.z > m.w ? m.z : m.w;
    mx = mx > my ? mx : my;
    if (mx > (MSG)0) { const MSG r = TolMath<MSG>::pow2_rescale(mx); m.x *= r; m.y *= r; m.z *= r; m.w *= r; }
}

NXB_DEV unsigned tp_sb(unsigned pcode) { return (pcode >> 24) & 0xffu; }
NXB_DEV unsigned tp_sa(unsigned pcode) { return (pcode >> 16) & 0xffu; }

// first primary event of the edge (index range [ip0, ip1)) with time > t
template <typename MSG>
NXB_DEV int tp_locate(const TpUnit<MSG>& u, int ip0, int ip1, double t) {
    int ip = ip0;
    while (ip < ip1 && u.ptm[ip] < t) ++ip;
    return ip;
}
// Penalty exponent per unit edge rate (chunking.py:76-79) and "an own-class segment is touched"
// (chunking.py:69-73) over (t_lo, t_hi) on an edge whose primary events in the interval are
// [ipl, iph); s = primary state at t_lo.
template <typename MSG>
NXB_DEV void tp_interval(const TpCtx& c, const TpUnit<MSG>& u, int k, int ipl, int iph, double t_lo, double t_hi,
                         int s, double& pen, bool& own) {
    const int K = c.K;
    double acc = 0.0, tc = t_lo;
    bool o = false;
    for (int ip = ipl; ip < iph; ++ip) {
        const double tn = u.ptm[ip];
        acc += c.qm[s * K + k] * (tn - tc);
        o = o || (c.cls[s] == k);
        s = (int)tp_sb(u.pcode[ip]);
        tc = tn;
    }
    acc += c.qm[s * K + k] * (t_hi - tc);
    o = o || (c.cls[s] == k);
    pen = acc; own = o;
}
// OFF time of track k over (t_lo, t_hi) on edge e, booked per primary state (summary.py:123-131)
template <typename MSG>
NXB_DEV void tp_add_off(const TpCtx& c, const TpUnit<MSG>& u, int e, int k, int ip0, int ip1, int s_top,
                        double t_lo, double t_hi) {
    const NxbSweepParams& p = *c.p;
    int ip = tp_locate(u, ip0, ip1, t_lo);
    int s = (ip == ip0) ? s_top : (int)tp_sb(u.pcode[ip - 1]);
    double tc = t_lo;
    double* const base = p.dwell + p.sl.d_off + (long long)e * c.P * c.K + k;
    while (ip < ip1 && u.ptm[ip] < t_hi) {
        const double tn = u.ptm[ip];
        if (tn > tc) atomicAdd(base + s * c.K, tn - tc);
        tc = tn;
        s = (int)tp_sb(u.pcode[ip]);
        ++ip;
    }
    if (t_hi > tc) atomicAdd(base + s * c.K, t_hi - tc);
}

// uniform of the backward draw, stored with the event
template <typename MSG> struct TpU;
template <> struct TpU<float> {
    // accepted candidate: the thinning word is uniform on [0, thr)
    static NXB_DEV unsigned from_thin(unsigned w, unsigned, unsigned thr) {
        const float f = ((float)w + 0.5f) * __frcp_rn((float)thr) * 16777216.0f;
        const unsigned v = (unsigned)f;
        return v > 0xffffffu ? 0xffffffu : v;
    }
    static NXB_DEV unsigned from_words(unsigned a, unsigned) { return a >> 8; }
    // in [0, 1): (s + 0.5) * 2^-24 would round to exactly 1 for s = 2^24 - 1 and break "u * w1 < w1"
    static NXB_DEV float value(unsigned s) { return (float)(s & 0xffffffu) * (1.0f / 16777216.0f); }
};
template <> struct TpU<double> {
    static NXB_DEV unsigned long long from_thin(unsigned w, unsigned w2, unsigned thr) {
        const double f = ((double)w * 4294967296.0 + (double)w2 + 0.5) / ((double)thr * 4294967296.0) * 9007199254740992.0;
        const unsigned long long v = (unsigned long long)f;
        return v > 0x1fffffffffffffull ? 0x1fffffffffffffull : v;
    }
    static NXB_DEV unsigned long long from_words(unsigned a, unsigned b) { return ((unsigned long long)a << 21) ^ (unsigned long long)(b >> 11); }
    static NXB_DEV double value(unsigned long long s) { return (double)(s & 0x1fffffffffffffull) * (1.0 / 9007199254740992.0); }
};

// next position fraction of a slice, descending: v *= u^(1/j)
template <typename MSG> struct TpGen;
template <> struct TpGen<float> {
    float vf;
    NXB_DEV void init() { vf = 1.0f; }
    NXB_DEV double next(NxbXo& xo, int j) {
        const unsigned w = nxb_xo_next(xo);
        const float f = ((float)(w >> 9) + 0.5f) * (1.0f / 8388608.0f);
        vf *= nxb_fast_exp2f(nxb_fast_log2f(f) * __frcp_rn((float)j));
        // low mantissa bits of the f64 value: the 9 unused bits of the word, then j (see tol_gap)
        return nxb_ll_as_double(nxb_double_as_ll((double)vf) | (long long)(((w & 0x1ffu) << 6) | ((unsigned)j & 63u)));
    }
};
template <> struct TpGen<double> {
    double vd;
    NXB_DEV void init() { vd = 1.0; }
    NXB_DEV double next(NxbXo& xo, int j) {
        const uint32_t hi = nxb_xo_next(xo), lo = nxb_xo_next(xo);
        vd *= exp2(::log2(nxb_u53(hi, lo)) / (double)j);
        return vd;
    }
};

NXB_DEV void tp_stg_extra(TpT<float>::Stg&, NxbXo&) {}
NXB_DEV void tp_stg_extra(TpT<double>::Stg& r, NxbXo& x) { r.w2 = nxb_xo_next(x); r.pad0 = 0u; r.pad1 = 0.0; }
NXB_DEV unsigned tp_stg_w2(const TpT<float>::Stg&) { return 0u; }
NXB_DEV unsigned tp_stg_w2(const TpT<double>::Stg& r) { return r.w2; }
// backward-sampling uniform of retained event j (index in the (track, edge, time) sorted list)
template <typename MSG> NXB_DEV typename TpT<MSG>::UA tp_ret_uniform(const TpCtx& c, unsigned unit32, unsigned sweep, int j);
template <> NXB_DEV unsigned tp_ret_uniform<float>(const TpCtx& c, unsigned unit32, unsigned sweep, int j) {
    const NxbU4 r = nxb_philox_call(unit32, sweep, NXB_STREAM_TOL_RET, (unsigned)(j >> 2), c.k0, c.k1);
    const unsigned w = (j & 2) ? ((j & 1) ? r.w : r.z) : ((j & 1) ? r.y : r.x);
    return TpU<float>::from_words(w, 0u);
}
template <> NXB_DEV unsigned long long tp_ret_uniform<double>(const TpCtx& c, unsigned unit32, unsigned sweep, int j) {
    const NxbU4 r = nxb_philox_call(unit32, sweep, NXB_STREAM_TOL_RET, (unsigned)(j >> 1), c.k0, c.k1);
    return (j & 1) ? TpU<double>::from_words(r.z, r.w) : TpU<double>::from_words(r.x, r.y);
}

enum { TP_OK = 0, TP_DEFER = 1, TP_ERROR = 2 };
enum { TPT_LOAD = 0, TPT_GEN, TPT_MERGE, TPT_THIN, TPT_MAT, TPT_SCAN, TPT_UP, TPT_MAPS, TPT_DOWN, TPT_SURV, TPT_FREE, TPT_STORE };

struct TpTimer {
    unsigned long long* perf;
    long long t_last;
    bool on;
    NXB_DEV void tick(int which) {
        if (perf) {
            const long long now = clock64();
            if (on) atomicAdd(perf + which, (unsigned long long)(now - t_last));
            t_last = now;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// The update of one unit by the WPU warps of a group.  The unit's trajectory is in shared memory
// (u.ptm/pcode/btm/bcode/node_tol/node_pri, u.nrec set up); on TP_OK the new tolerance states are
// in u.newtol, the new blink list is in the unit's slab and *n_out holds its length.  TP_DEFER:
// nothing persistent was touched, the unit needs a larger capacity tier.
// ---------------------------------------------------------------------------------------------
template <typename MSG, int WPU>
NXB_DEV int tolpar_update(const TpCtx& c, TpUnit<MSG>& u, const TpGroup<WPU>& g, bool accum,
                          unsigned char* slab, int* n_out, int* err_detail, TpTimer& tm) {
    typedef TolMath<MSG> M;
    typedef typename TpT<MSG>::M4 M4;
    typedef typename TpT<MSG>::UA UA;
    typedef typename TpT<MSG>::Stg Stg;
    const NxbSweepParams& p = *c.p;
    const NxbModelLayout& ml = p.ml;
    const NxbTolLayout& tl = p.tl;
    const int E = c.E, K = c.K, G = 32 * WPU;
    const int li = g.li, lane = g.lane;
    const unsigned lt_mask = (1u << lane) - 1u;
    const SPtr<int> xch = {u.x.off};             // [0 .. 2 * WPU): scan exchange slots
    int xslot = 0;
    const int nB = u.nB;

    // ================================================================== G: candidates
    // slices owned by this lane: consecutive, so that the group scan gives rank offsets
    constexpr int SPL = NXB_TP_NSLICE / (32 * WPU);      // slices per lane (2 or 1)
    NxbXo xo[SPL];
    int ns[SPL];
    int mine = 0;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const int s = li * SPL + i;
        const NxbU4 r = nxb_philox_call(u.unit32, u.sweep, NXB_STREAM_TOL_SLICE | (unsigned)s, 0u, c.k0, c.k1);
        xo[i].s0 = r.x; xo[i].s1 = r.y; xo[i].s2 = r.z; xo[i].s3 = r.w | 1u;
        // Poisson(W) count by inversion: first j with word < pcdf[j]
        const unsigned w = nxb_xo_next(xo[i]);
        int lo = 0, hi = ml.n_pcdf - 1;
        while (lo < hi) { const int m = (lo + hi) >> 1; if (w < c.pcdf[m]) hi = m; else lo = m + 1; }
        ns[i] = ml.tp_lam > 0.0 ? lo : 0;
        mine += ns[i];
    }
    int NC;
    int off = tp_group_excl_scan<WPU>(g, mine, NC, xch, xslot);
    if (NC > tl.capQ || nB > tl.capB) return TP_DEFER;
    {
        const double W = ml.tp_W, lam = ml.tp_lam, inv_lam = ml.tp_inv_lam, inv_bin = ml.tp_inv_bin;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            const int s = li * SPL + i;
            TpGen<MSG> gen;
            gen.init();
            double vprev = 1.0;
            NxbXo x = xo[i];
            for (int j = ns[i]; j >= 1; --j) {
                double vd = gen.next(x, j);
                if (!(vd < vprev)) vd = nxb_ll_as_double(nxb_double_as_ll(vprev) - 1LL);
                vprev = vd;
                // position on the axis -> (track, edge, time): direct lookups, monotone in the position
                const double y = ((double)s + vd) * W;
                int k = (int)(y * inv_lam);
                k = k < K - 1 ? k : K - 1;
                double yl = y - (double)k * lam;
                if (yl < 0.0) { if (k > 0) { --k; yl += lam; } else yl = 0.0; }
                int b = (int)(yl * inv_bin);
                b = b < NXB_TP_NBIN - 1 ? b : NXB_TP_NBIN - 1;
                int e = (int)c.etab[b];
                NxbGenRec gr = c.grec[e];
                while (yl < gr.cum && e > 0) { --e; gr = c.grec[e]; }
                double t = gr.tma + (yl - gr.cum) * gr.inv_om;
                unsigned a = (unsigned)e | ((unsigned)k << TPK_TRACK_SHIFT);
                if (!(t > gr.tma && t < gr.tmb)) {          // rounding at the ends of the edge, inactive edge
                    a |= TPS_DEAD;
                    t = t < gr.tmb ? gr.tma : gr.tmb;
                }
                Stg rec;
                rec.t = t; rec.a = a; rec.w = nxb_xo_next(x);
                tp_stg_extra(rec, x);
                u.stg[off + (j - 1)] = rec;
            }
            off += ns[i];
        }
    }
    g.sync();
    tm.tick(TPT_GEN);

    // ================================================================== M: merge ranks
    // ins[j] = number of candidates that sort before retained event j
    for (int j = li; j < nB; j += G) {
        const unsigned code = u.bcode[j];
        const unsigned bk = (code & 0xffffu) | (((code >> 16) & 0xffu) << TPK_TRACK_SHIFT);
        const double tj = u.btm[j];
        int lo = 0, hi = NC;
        while (lo < hi) {
            const int m = (lo + hi) >> 1;
            const Stg r = u.stg[m];
            const unsigned ca = r.a & TPK_BKT_MASK;
            if (ca < bk || (ca == bk && r.t < tj)) lo = m + 1; else hi = m;
        }
        u.ins[j] = (unsigned short)lo;
    }
    g.sync();
    tm.tick(TPT_MERGE);

    // ================================================================== T: thinning + placement
    int bad = 0;
    int nLC = 0;                // live candidates so far
    for (int b = 0; b < NC; b += G) {
        const int cidx = b + li;
        bool live = false;
        int r = 0, ipf = 0;
        unsigned kf = 0u;
        double t = 0.0;
        unsigned wthin = 0u, w2 = 0u, thr = 1u;
        if (cidx < NC) {
            const Stg rec = u.stg[cidx];
            t = rec.t; wthin = rec.w; w2 = tp_stg_w2(rec);
            const unsigned bkt = rec.a & TPK_BKT_MASK;
            const int e = (int)(bkt & TPK_EDGE_MASK), k = (int)(bkt >> TPK_TRACK_SHIFT);
            bool dead = (rec.a & TPS_DEAD) != 0u;
            // retained events before this candidate: r = #{j : ins[j] <= cidx}
            {
                int lo = 0, hi = nB;
                while (lo < hi) { const int m = (lo + hi) >> 1; if ((int)u.ins[m] <= cidx) lo = m + 1; else hi = m; }
                r = lo;
            }
            const uint4 nr = u.nrec[e];
            const int ip0 = (int)(nr.w & 0xffffu), ip1 = (int)(nr.w >> 16);
            const int ip = tp_locate(u, ip0, ip1, t);
            const int s = (ip == ip0) ? (int)((nr.z >> 8) & 0xffu) : (int)tp_sb(u.pcode[ip - 1]);
            const bool own = (c.cls[s] == k);
            unsigned xold = (u.node_tol[c.erec[e].va] >> k) & 1u;
            if (r > 0) {
                const unsigned code = u.bcode[r - 1];
                if ((code & 0xffffffu) == ((unsigned)e | ((unsigned)k << 16))) {
                    xold = code >> 31;
                    if (u.btm[r - 1] == t) dead = true;       // tie with a retained event of the same track
                }
            }
            if (cidx > 0) {
                const Stg pr = u.stg[cidx - 1];
                if ((pr.a & TPK_BKT_MASK) == bkt && !(pr.t < t)) dead = true;   // tie between candidates (rounding)
            }
            thr = ml.tp_thr[(own ? 2 : 0) + (int)xold];
            live = !dead && wthin < thr;
            ipf = ip - ip0;
            if (ipf > (int)TPK_AUX_MAX) { bad = 2; ipf = TPK_AUX_MAX; }
            kf = bkt | (own ? TPK_OWN : 0u) | TPK_CAND | (xold ? TPK_X : 0u) | ((unsigned)ipf << TPK_AUX_SHIFT);
        }
        int tot;
        const int ex = tp_group_excl_count<WPU>(g, live, tot, xch, xslot);
        if (cidx < NC) u.stg[cidx].w = (unsigned)(nLC + ex);          // live candidates before this one
        if (live) {
            const int R = nLC + ex + r;
            if (R < tl.capL) {
                u.T[R] = t; u.key[R] = kf;
                u.aux[R] = TpU<MSG>::from_thin(wthin, w2, thr);
            }
        }
        nLC += tot;
    }
    const int nL = nLC + nB;
    if (nL > tl.capL) return TP_DEFER;
    g.sync();
    // retained events
    for (int j = li; j < nB; j += G) {
        const int ci = (int)u.ins[j];
        const int R = j + (ci < NC ? (int)u.stg[ci].w : nLC);
        const unsigned code = u.bcode[j];
        const int e = (int)(code & 0xffffu), k = (int)((code >> 16) & 0xffu);
        const double t = u.btm[j];
        const uint4 nr = u.nrec[e];
        const int ip0 = (int)(nr.w & 0xffffu), ip1 = (int)(nr.w >> 16);
        const int ip = tp_locate(u, ip0, ip1, t);
        const int s = (ip == ip0) ? (int)((nr.z >> 8) & 0xffu) : (int)tp_sb(u.pcode[ip - 1]);
        int ipf = ip - ip0;
        if (ipf > (int)TPK_AUX_MAX) { bad = 2; ipf = TPK_AUX_MAX; }
        u.T[R] = t;
        u.key[R] = (unsigned)e | ((unsigned)k << TPK_TRACK_SHIFT) | ((c.cls[s] == k) ? TPK_OWN : 0u) |
                   ((code >> 31) ? 0u : TPK_X) | ((unsigned)ipf << TPK_AUX_SHIFT);
        u.aux[R] = tp_ret_uniform<MSG>(c, u.unit32, u.sweep, j);
    }
    g.sync();
    tm.tick(TPT_THIN);

    // ================================================================== B: matrices
    // toff[k] = first live event of track >= k (binary search, lane = track)
    for (int k = li; k <= K; k += G) {
        int lo = 0, hi = nL;
        while (lo < hi) { const int m = (lo + hi) >> 1; if ((int)((u.key[m] >> TPK_TRACK_SHIFT) & 31u) < k) lo = m + 1; else hi = m; }
        u.toff[k] = lo;
    }
    {
        const MSG a_on = (MSG)ml.a_on, a_off = (MSG)ml.a_off;
        for (int r = li; r < nL; r += G) {
            const unsigned key = u.key[r];
            const unsigned bkt = key & TPK_BKT_MASK;
            const int e = (int)(bkt & TPK_EDGE_MASK), k = (int)(bkt >> TPK_TRACK_SHIFT);
            const uint4 nr = u.nrec[e];
            const int ip0 = (int)(nr.w & 0xffffu);
            const int ip = ip0 + (int)(key >> TPK_AUX_SHIFT);
            const NxbEdgeRec er = c.erec[e];
            double t_lo = er.tma;
            int ipl = ip0;
            if (r > 0) {
                const unsigned pk = u.key[r - 1];
                if ((pk & TPK_BKT_MASK) == bkt) { t_lo = u.T[r - 1]; ipl = ip0 + (int)(pk >> TPK_AUX_SHIFT); }
            }
            const int s_lo = (ipl == ip0) ? (int)((nr.z >> 8) & 0xffu) : (int)tp_sb(u.pcode[ipl - 1]);
            double pen; bool own_iv;
            tp_interval(c, u, k, ipl, ip, t_lo, u.T[r], s_lo, pen, own_iv);
            // diag(z, e1) * A; with OFF impossible (z == 0) the factor e1 is a pure rescaling and is
            // dropped, so that a long own-class segment cannot underflow ON to 0
            const MSG z = own_iv ? (MSG)0 : (MSG)1;
            const MSG e1 = own_iv ? (MSG)1 : M::expneg(pen * er.rate);
            M4 bm;
            if (key & TPK_OWN) bm = TpT<MSG>::make((MSG)0, (MSG)0, (MSG)0, (MSG)1);      // (own at the event => own_iv)
            else bm = TpT<MSG>::make(z * ((MSG)1 - a_on), z * a_on, e1 * a_off, e1 * ((MSG)1 - a_off));
            u.q[r] = bm;
        }
    }
    g.sync();
    tm.tick(TPT_MAT);

    // ================================================================== S: suffix products per bucket
    // each warp of the group takes a contiguous part of the list that ends at a bucket boundary
    int rlo = 0, rhi = nL;
    if (WPU == 2) {
        int mid = nL >> 1;
        while (mid > 0 && mid < nL && ((u.key[mid - 1] ^ u.key[mid]) & TPK_BKT_MASK) == 0u) ++mid;
        if (g.w == 0) rhi = mid; else rlo = mid;
    }
    {
        M4 carry = TpT<MSG>::make((MSG)1, (MSG)0, (MSG)0, (MSG)1);
        const int nblk = (rhi - rlo + 31) >> 5;
        for (int bi = nblk - 1; bi >= 0; --bi) {
            const int r = rlo + (bi << 5) + lane;
            const bool valid = r < rhi;
            unsigned key = 0u;
            M4 m = TpT<MSG>::make((MSG)1, (MSG)0, (MSG)0, (MSG)1);
            bool islast = true;
            if (valid) {
                key = u.key[r];
                m = u.q[r];
                islast = (r + 1 >= nL) || (((u.key[r + 1] ^ key) & TPK_BKT_MASK) != 0u);
            }
            bool isfirst = false;
            if (valid) isfirst = (r == 0) || (((u.key[r - 1] ^ key) & TPK_BKT_MASK) != 0u);
            const unsigned fmb = __ballot_sync(0xffffffffu, isfirst) & ((2u << lane) - 1u);   // first flags at or below this lane
            const unsigned lm = __ballot_sync(0xffffffffu, islast) >> lane;      // bit 0: this lane
            const bool cont = (lm == 0u);                 // the bucket continues in the block above
            const int rem = cont ? (31 - lane) : (__ffs(lm) - 1);
            const int maxrem = __reduce_max_sync(0xffffffffu, rem);
            for (int d = 1; d <= maxrem; d <<= 1) {
                const M4 o = tp_shfl_down<M4>(m, d);
                if (d <= rem) m = tp_mul<MSG, M4>(m, o);
            }
            if (cont) m = tp_mul<MSG, M4>(m, carry);
            tp_rescale<MSG, M4>(m);
            if (valid) {
                u.q[r] = m;
                // the last event of a bucket carries the distance to the bucket's first event
                // (sentinel when that one lies in a lower block)
                int dh = 0;
                if (islast) dh = fmb ? (lane - (31 - __clz(fmb))) : (int)TPK_AUX_MAX;
                u.key[r] = (key & ((1u << TPK_AUX_SHIFT) - 1u) & ~TPK_LAST) | (islast ? TPK_LAST : 0u) | ((unsigned)dh << TPK_AUX_SHIFT);
                if (islast) atomicOr(&u.hasev[(int)(key & TPK_EDGE_MASK)], 1u << ((key >> TPK_TRACK_SHIFT) & 31u));
            }
            carry.x = __shfl_sync(0xffffffffu, m.x, 0); carry.y = __shfl_sync(0xffffffffu, m.y, 0);
            carry.z = __shfl_sync(0xffffffffu, m.z, 0); carry.w = __shfl_sync(0xffffffffu, m.w, 0);
        }
    }
    g.sync();
    tm.tick(TPT_SCAN);

    // ================================================================== U: up the tree, level by level
    // lane = (internal node of the level, track): the messages of the node's children (one 2x2
    // mat-vec per non-empty bucket) are multiplied into the node's accumulator.  Levels are the
    // heights of the nodes, so every child is done before its parent.
    for (int h = 0; h < ml.n_levels; ++h) {
        const int n0 = (int)c.lvl_off[h], nn = (int)c.lvl_off[h + 1] - n0;
        for (int task = li; task < nn * K; task += G) {
            const int vi = task / K, k = task - vi * K;
            const int v = (int)c.lvl_node[n0 + vi];
            const int tlo = u.toff[k], thi = u.toff[k + 1];
            AccRec<MSG> n; n.u0 = (MSG)1; n.u1 = (MSG)1; n.pen = 0.0;
            for (int ci = (int)c.ch_off[v]; ci < (int)c.ch_off[v + 1]; ++ci) {
                const int cn = (int)c.ch_node[ci], e = cn - 1;
                const NxbEdgeRec er = c.erec[e];
                const uint4 nr = u.nrec[e];
                MSG u0 = 1, u1 = 1;
                double pen = 0.0;
                const int cs = (int)c.lslot[cn];
                if (cs >= 0) { const AccRec<MSG> a = u.acc[cs * K + k]; u0 = a.u0; u1 = a.u1; pen = a.pen; }
                if (!((nr.x >> k) & 1u)) u1 = 0;
                if (!((nr.y >> k) & 1u)) u0 = 0;
                const int ip0 = (int)(nr.w & 0xffffu), ip1 = (int)(nr.w >> 16);
                const bool has = ((u.hasev[e] >> k) & 1u) != 0u;
                int cur = -1;
                if (has) {          // last event of the bucket (k, e): events of a track are sorted by edge
                    int lo = tlo, hi = thi;
                    while (lo < hi) { const int m = (lo + hi) >> 1; if ((int)(u.key[m] & TPK_EDGE_MASK) <= e) lo = m + 1; else hi = m; }
                    cur = lo - 1;
                }
                const double t_lo = has ? u.T[cur] : er.tma;
                double pen_b; bool own_b;
                if (ip0 == ip1) {
                    // (the same for every track of the unit) no primary event on this edge: one segment
                    const int s = (int)(nr.z & 0xffu);
                    pen_b = c.qm[s * K + k] * (er.tmb - t_lo);
                    own_b = (c.cls[s] == k);
                } else {
                    const int ipl = has ? tp_locate(u, ip0, ip1, t_lo) : ip0;
                    const int s_lo = (ipl == ip0) ? (int)((nr.z >> 8) & 0xffu) : (int)tp_sb(u.pcode[ipl - 1]);
                    tp_interval(c, u, k, ipl, ip1, t_lo, er.tmb, s_lo, pen_b, own_b);
                }
                pen += pen_b * er.rate;
                if (own_b) u0 = 0;
                if (has) {
                    if (pen != 0.0) { if (u0 != (MSG)0) u1 *= M::expneg(pen); pen = 0.0; }
                    MSG mx = u0 > u1 ? u0 : u1;
                    if (!(mx > (MSG)0)) { bad |= 1; mx = 1; }
                    const MSG rs = M::pow2_rescale(mx);
                    const MSG l0 = u0 * rs, l1 = u1 * rs;
                    const unsigned ckey = u.key[cur];
                    // (the last event of a bucket carries the distance to its first one, if the block of
                    // the suffix scan saw it)
                    int head = cur - (int)(ckey >> TPK_AUX_SHIFT);
                    if ((ckey >> TPK_AUX_SHIFT) == TPK_AUX_MAX) {
                        head = cur;
                        while (head - 1 >= tlo && (int)(u.key[head - 1] & TPK_EDGE_MASK) == e) --head;
                    }
                    const M4 qh = u.q[head];
                    u0 = qh.x * l0 + qh.y * l1;
                    u1 = qh.z * l0 + qh.w * l1;
                    u.q[head] = TpT<MSG>::make(l0, l1, (MSG)0, (MSG)0);      // message at the bucket's last event
                }
                n.u0 *= u0; n.u1 *= u1; n.pen += pen;
            }
            MSG m2 = n.u0 > n.u1 ? n.u0 : n.u1;
            if (!(m2 > (MSG)0)) { bad |= 1; m2 = 1; }
            const MSG r2 = M::pow2_rescale(m2);
            n.u0 *= r2; n.u1 *= r2;
            nxb_acc_pad(n);
            u.acc[(int)c.lslot[v] * K + k] = n;
        }
        g.sync();
    }
    // root draw (prior: toymodel.py:17-27 get_blink_distn), lane = track
    if (g.w == 0) {
        const int k = lane;
        const bool trk = k < K;
        unsigned xroot = 0u;
        if (trk) {
            MSG a0 = 1, a1 = 1; double apen = 0.0;
            if ((int)c.lslot[0] >= 0) { const AccRec<MSG> a = u.acc[(int)c.lslot[0] * K + k]; a0 = a.u0; a1 = a.u1; apen = a.pen; }
            double w1 = (double)a1 * (a0 != (MSG)0 ? fmax(exp(-apen), 2.2250738585072014e-308) : 1.0) * ml.p_on;
            double w0 = (double)a0 * ml.p_off;
            if (!(((unsigned)u.x[8] >> k) & 1u)) w1 = 0.0;
            if (!(((unsigned)u.x[9] >> k) & 1u)) w0 = 0.0;
            if (c.cls[u.node_pri[0]] == k) w0 = 0.0;
            if (!(w0 + w1 > 0.0)) bad |= 1;
            const NxbU4 rr = nxb_philox_call(u.unit32, u.sweep, NXB_STREAM_TOL_ROOT, (unsigned)(k >> 2), c.k0, c.k1);
            const unsigned wsel = (k & 2) ? ((k & 1) ? rr.w : rr.z) : ((k & 1) ? rr.y : rr.x);
            xroot = (nxb_u32(wsel) * (w0 + w1) < w1) ? 1u : 0u;
        }
        const unsigned rootbits = __ballot_sync(0xffffffffu, trk && xroot);
        if (lane == 0) u.newtol[0] = rootbits;
    }
    g.sync();
    tm.tick(TPT_UP);

    // ================================================================== F: backward maps
    {
        const MSG a_on = (MSG)ml.a_on, a_off1 = (MSG)(1.0 - ml.a_off);
        int carry_head = rlo;
        unsigned carry_map = 2u;                    // identity: f(0) = 0, f(1) = 1
        const int nblk = (rhi - rlo + 31) >> 5;
        for (int bi = 0; bi < nblk; ++bi) {
            const int b = rlo + (bi << 5), r = b + lane;
            const bool valid = r < rhi;
            unsigned key = 0u;
            bool isfirst = false;
            if (valid) {
                key = u.key[r];
                isfirst = (r == 0) || (((u.key[r - 1] ^ key) & TPK_BKT_MASK) != 0u);
            }
            const unsigned fm = __ballot_sync(0xffffffffu, isfirst);
            const unsigned below = fm & ((2u << lane) - 1u);
            const int hp = below ? (31 - __clz(below)) : 0;          // start of the segment inside the block
            const int head = below ? (b + hp) : carry_head;
            unsigned map = 2u, zz = 0u;
            if (valid) {
                const M4 mbv = u.q[head];
                MSG l0 = mbv.x, l1 = mbv.y;
                if (!(key & TPK_LAST)) {
                    const M4 qn = u.q[r + 1];
                    const MSG t0 = qn.x * l0 + qn.y * l1, t1 = qn.z * l0 + qn.w * l1;
                    l0 = t0; l1 = t1;
                }
                const MSG uu = TpU<MSG>::value(u.aux[r]);
                const bool own = (key & TPK_OWN) != 0u;
                // state above = 0
                {
                    const MSG p1 = own ? (MSG)0.5 : a_on;
                    const MSG w1 = p1 * l1, w0 = ((MSG)1 - p1) * l0;
                    if (!(w0 + w1 > (MSG)0)) zz |= 1u;
                    if (uu * (w0 + w1) < w1) map |= 1u; else map &= ~1u;
                }
                // state above = 1
                {
                    const MSG p1 = own ? (MSG)1 : a_off1;
                    const MSG w1 = p1 * l1, w0 = ((MSG)1 - p1) * l0;
                    if (!(w0 + w1 > (MSG)0)) zz |= 2u;
                    if (uu * (w0 + w1) < w1) map |= 2u; else map &= ~2u;
                }
            }
            // inclusive prefix composition inside the bucket: gm = f_r o ... o f_head
            unsigned gm = map;
            const int back = lane - hp;
            const int maxback = __reduce_max_sync(0xffffffffu, valid ? back : 0);
            for (int d = 1; d <= maxback; d <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, gm, d);
                if (d <= back) gm = ((gm >> (o & 1u)) & 1u) | (((gm >> ((o >> 1) & 1u)) & 1u) << 1);
            }
            if (!below) gm = ((gm >> (carry_map & 1u)) & 1u) | (((gm >> ((carry_map >> 1) & 1u)) & 1u) << 1);
            if (valid) {
                u.aux[r] = (UA)(map | (zz << 2) | (gm << 4));
                if (key & TPK_LAST) {
                    const int e = (int)(key & TPK_EDGE_MASK);
                    const unsigned bit = 1u << ((key >> TPK_TRACK_SHIFT) & 31u);
                    if (gm & 1u) atomicOr(&u.f0[e], bit);
                    if (gm & 2u) atomicOr(&u.f1[e], bit);
                }
            }
            carry_map = __shfl_sync(0xffffffffu, gm, 31);
            carry_head = __shfl_sync(0xffffffffu, head, 31);
        }
    }
    g.sync();
    tm.tick(TPT_MAPS);

    // ================================================================== D: node states, all tracks at once
    // lane = edge of one depth level (parents before children); bit-parallel over the tracks
    for (int dl = 0; dl < ml.n_depths; ++dl) {
        const int e0 = (int)c.dep_off[dl], ne = (int)c.dep_off[dl + 1] - e0;
        for (int i = li; i < ne; i += G) {
            const int e = (int)c.dep_edge[e0 + i];
            const unsigned xa = u.newtol[c.erec[e].va];
            const unsigned hv = u.hasev[e];
            u.newtol[e + 1] = (xa & ~hv) | (hv & ((xa & u.f1[e]) | (~xa & u.f0[e])));
        }
        g.sync();
    }
    tm.tick(TPT_DOWN);

    // ================================================================== W: survivors
    // pass 1: state above / below every event, survivor counts
    int nsurv_w = 0;
    {
        const int nblk = (rhi - rlo + 31) >> 5;
        for (int bi = 0; bi < nblk; ++bi) {
            const int r = rlo + (bi << 5) + lane;
            bool surv = false;
            if (r < rhi) {
                const unsigned key = u.key[r];
                const int e = (int)(key & TPK_EDGE_MASK), k = (int)((key >> TPK_TRACK_SHIFT) & 31u);
                const unsigned xva = (u.newtol[c.erec[e].va] >> k) & 1u;
                const bool isfirst = (r == 0) || (((u.key[r - 1] ^ key) & TPK_BKT_MASK) != 0u);
                const unsigned gprev = isfirst ? 2u : (((unsigned)u.aux[r - 1] >> 4) & 3u);
                const unsigned a = (unsigned)u.aux[r];
                const unsigned xa = (gprev >> xva) & 1u;
                const unsigned xb = (a >> xa) & 1u;
                if ((a >> (2 + xa)) & 1u) bad |= 1;
                surv = xa != xb;
                u.key[r] = (key & ~TPK_X & ~TPK_CAND) | (xb ? TPK_X : 0u) | (surv ? TPK_CAND : 0u);   // X: new state below, CAND: survivor
            }
            nsurv_w += __popc(__ballot_sync(0xffffffffu, surv));
        }
    }
#if defined(NXB_TP_DEBUG)
    g.sync();
    if (li == 0 && (long long)u.unit32 == p.dbg_unit + p.unit_offset && (int)u.sweep == p.dbg_sweep) {
        printf("DBG unit %u sweep %u NC %d nB %d nL %d\n", u.unit32, u.sweep, NC, nB, nL);
        for (int r = 0; r < nL; ++r) {
            const unsigned key = u.key[r];
            printf("  r %3d k %2u e %2u own %u surv %u xb %u last %u de %3u t %.15g aux %x q %g %g %g %g\n", r,
                   (key >> TPK_TRACK_SHIFT) & 31u, key & TPK_EDGE_MASK, (key >> 18) & 1u, (key >> 19) & 1u, (key >> 20) & 1u,
                   (key >> 21) & 1u, key >> TPK_AUX_SHIFT, u.T[r], (unsigned)u.aux[r], (double)u.q[r].x, (double)u.q[r].y, (double)u.q[r].z, (double)u.q[r].w);
        }
        for (int e = 0; e < E; ++e) printf("  e %2d hasev %05x f0 %05x f1 %05x newtol %05x\n", e, u.hasev[e], u.f0[e], u.f1[e], u.newtol[e + 1]);
    }
    g.sync();
#endif
    int nsurv;
    int sbase;
    {
        // (lane 0 of each warp carries the warp's count)
        int tot;
        const int ex = tp_group_excl_scan<WPU>(g, lane == 0 ? nsurv_w : 0, tot, xch, xslot);
        sbase = __shfl_sync(0xffffffffu, ex, 0);
        nsurv = tot;
    }
    // all lanes of the group now agree on `bad`?  collect it
    {
        const int anybad = __reduce_or_sync(0xffffffffu, bad);
        if (lane == 0 && anybad) atomicOr(&u.x[4], anybad);
    }
    g.sync();
    {
        const int flags = u.x[4];
        if (flags) { *err_detail = flags; return TP_ERROR; }
    }
    if (nsurv > p.gcapB) { *err_detail = -nsurv; return TP_ERROR; }
    // pass 2: write the survivors (already in (track, edge, time) order), statistics
    {
        double* const obtm = (double*)(slab + p.so_btm);
        unsigned* const obcode = (unsigned*)(slab + p.so_bcode);
        int pos = sbase;
        const int nblk = (rhi - rlo + 31) >> 5;
        for (int bi = 0; bi < nblk; ++bi) {
            const int r = rlo + (bi << 5) + lane;
            bool surv = false;
            unsigned key = 0u;
            if (r < rhi) { key = u.key[r]; surv = (key & TPK_CAND) != 0u; }
            const unsigned sb = __ballot_sync(0xffffffffu, surv);
            if (r < rhi) {
                const int e = (int)(key & TPK_EDGE_MASK), k = (int)((key >> TPK_TRACK_SHIFT) & 31u);
                const unsigned xb = (key & TPK_X) ? 1u : 0u;
                const double t = u.T[r];
                if (surv) {
                    const int o = pos + __popc(sb & lt_mask);
                    __stcs(obtm + o, t);
                    __stcs(obcode + o, (unsigned)e | ((unsigned)k << 16) | (xb << 31));
                    // packed word per edge (low: off->on, high: on->off)
                    if (accum) atomicAdd((unsigned*)&c.s_off_on[e] + (xb ? 0 : 1), 1u);
                }
                if (accum) {
                    const uint4 nr = u.nrec[e];
                    const int ip0 = (int)(nr.w & 0xffffu), ip1 = (int)(nr.w >> 16);
                    const int s_top = (int)((nr.z >> 8) & 0xffu);
                    if (!xb) {
                        const double t_hi = (key & TPK_LAST) ? c.erec[e].tmb : u.T[r + 1];
                        tp_add_off(c, u, e, k, ip0, ip1, s_top, t, t_hi);
                    }
                    const bool isfirst = (r == 0) || (((u.key[r - 1] ^ key) & TPK_BKT_MASK) != 0u);
                    if (isfirst && !((u.newtol[c.erec[e].va] >> k) & 1u))
                        tp_add_off(c, u, e, k, ip0, ip1, s_top, c.erec[e].tma, t);
                }
            }
            pos += __popc(sb);
        }
    }
    tm.tick(TPT_SURV);
    // tracks that are OFF along a whole edge without events
    if (accum) {
        const unsigned ALLK = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);
        for (int e = li; e < E; e += G) {
            const NxbEdgeRec er = c.erec[e];
            if (!(er.tmb > er.tma)) continue;
            unsigned m = ~u.newtol[er.va] & ~u.hasev[e] & ALLK;
            if (!m) continue;
            const uint4 nr = u.nrec[e];
            const int ip0 = (int)(nr.w & 0xffffu), ip1 = (int)(nr.w >> 16);
            int s = (int)((nr.z >> 8) & 0xffu);
            double tc = er.tma;
            for (int ip = ip0; ip <= ip1; ++ip) {
                const double tn = (ip < ip1) ? u.ptm[ip] : er.tmb;
                if (tn > tc) {
                    double* const base = p.dwell + p.sl.d_off + ((long long)e * c.P + s) * K;
                    unsigned mm = m;
                    while (mm) { const int k = __ffs(mm) - 1; mm &= mm - 1u; atomicAdd(base + k, tn - tc); }
                }
                if (ip < ip1) { s = (int)tp_sb(u.pcode[ip]); tc = tn; }
            }
        }
    }
    tm.tick(TPT_FREE);
    *n_out = nsurv;
    return TP_OK;
}

}  // namespace nxb
