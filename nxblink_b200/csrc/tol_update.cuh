// TOLERANCE TRACK UPDATES (raoteh.py:413-432) of one unit: all K tracks at once, one lane per
// (unit, track).  Replaces, per track, poisson.py:176-199 (candidate events), chunking.py:39-155
// (blinking chunk tree), chunking.py:264-289 -> nxmctree sample_history (FFBS) and the tolerance
// half of summary.py:190-243.
//
// Round-2 formulation.  The per-track work is split by what it is proportional to:
//
//   * per POINT of the track (retained event, candidate, primary event): `tol_walk_up`, a walk over
//     the points only -- it never steps through event-free edges.  Candidates are the arrivals of one
//     homogeneous Poisson process per track on the rate-time line of the tree (edges in descending
//     index, each from its leafward to its rootward end, length rate x branch length), thinned to the
//     local rate of poisson.py:111-116 -- one exponential gap per arrival, none per edge.  Within an
//     edge that holds live (retained or accepted) events the walk multiplies out the 2x2 transfer
//     operator of the edge, T_e[x_top][x_bottom], column by column (one scale-free message per
//     bottom state), and leaves one 16-byte record per live event (time, and for either bottom state
//     the probability that the track is ON just below the event) and one per such edge (T_e);
//   * per EDGE of the tree: `tol_tree_up` / `tol_tree_down`, warp-uniform loops over the edges with
//     no inner loop: an edge without live events is diagonal (own class forces ON, ON pays
//     exp(-penalty), chunking.py:67-79), an edge with a record applies T_e; the downward loop draws
//     the state at every node;
//   * per LIVE event: `tol_walk_down`, which, given the states at both ends of an edge, draws the
//     states between its live events, keeps the real transitions and books the statistics.
//
// All of it is the exact FFBS of the reference on the chunk tree of the track (a chunk = a maximal
// region without live events); only the order of the arithmetic differs.
//
// The per-lane functions compile for the device and, for the CPU simulation under oracle/hostsim
// (test infrastructure), for the host.
#pragma once
#include <stdint.h>
#include "nxb_types.h"
#include "philox.cuh"
#include "nxb_portable.h"

#if defined(TOL_DEBUG)
#include <stdio.h>
#define TOL_BAD(var, id) do { var = 1; fprintf(stderr, "tol bad at site %d (line %d)\n", id, __LINE__); } while (0)
#else
#define TOL_BAD(var, id) var = 1
#endif

namespace nxb {

#define CE_EDGE 0x3fffu
#define CE_NEW1 0x4000u     // (survivors only) new state of the surviving event

// ---- records of the per-(unit, track) region of the L2-resident pool: live events from the bottom
// of the region upwards, one record per edge with live events from the top downwards ----
template <typename MSG> struct TolRec;
template <> struct __align__(16) TolRec<float> {
    union { double t; struct { float lscale; unsigned short e, n; } s; };
    union { float a; unsigned code; };
    float b;
};
template <> struct __align__(16) TolRec<double> {
    union { double t; struct { unsigned short e, n; unsigned pad; } s; };
    union { double a; unsigned code; };
    double b;
    double lscale;
};
//   live event : t = time; a = P(ON just below the event | bottom state OFF), sign bit set when the
//                primary state at the event is in this class; b = the same given bottom state ON
//   edge       : s.e = edge, s.n = live events on it; a, b, lscale = the two columns of T_e as
//                (P(top ON | column), log2 of the weight of column ON over column OFF); the upward
//                tree pass replaces a, b by P(bottom ON | top OFF), P(bottom ON | top ON)
//   survivor   : t = time, code = edge | CE_NEW1 (written over consumed live events)

NXB_DEV void nxb_set_lscale(TolRec<float>& r, float ls) { r.s.lscale = ls; }
NXB_DEV void nxb_set_lscale(TolRec<double>& r, double ls) { r.lscale = ls; }
NXB_DEV float nxb_get_lscale(const TolRec<float>& r) { return r.s.lscale; }
NXB_DEV double nxb_get_lscale(const TolRec<double>& r) { return r.lscale; }
NXB_DEV void nxb_store_ab(TolRec<float>* r, float a, float b) { r->a = a; r->b = b; }
NXB_DEV void nxb_store_ab(TolRec<double>* r, double a, double b) { r->a = a; r->b = b; }
// (live events and survivors use t as the time: only the padding of the f64 record is cleared)
NXB_DEV void tol_rec_pad(TolRec<float>&) {}
NXB_DEV void tol_rec_pad(TolRec<double>& r) { r.lscale = 0.0; }

// accumulator of a node for one track: product of the (OFF, ON) messages of the children folded so
// far (rescaled by powers of two) and the penalty exponent still to be applied to ON
template <typename MSG> struct AccRec;
template <> struct __align__(16) AccRec<float> { double pen; float u0, u1; };
template <> struct __align__(16) AccRec<double> { double pen; double u0, u1; double pad; };
NXB_DEV void nxb_acc_pad(AccRec<float>&) {}
NXB_DEV void nxb_acc_pad(AccRec<double>& n) { n.pad = 0.0; }

template <typename MSG> struct TolMath;
template <> struct TolMath<float> {
    static NXB_DEV float expneg(double pen) { return nxb_fast_expf((float)(-pen)); }
    static NXB_DEV float log2(float x) { return nxb_fast_log2f(x); }
    static NXB_DEV float exp2(float x) { return nxb_fast_exp2f(x); }
    static NXB_DEV float div(float a, float b) { return nxb_fast_divf(a, b); }
    static NXB_DEV float tiny() { return 5.421010862427522e-20f; }     // 2^-64
    static NXB_DEV float big() { return 1.8446744073709552e19f; }      // 2^64
    // 2^-(floor(log2 m) + 1) for a positive finite m: exact rescaling into [0.5, 1)
    static NXB_DEV float pow2_rescale(float m) {
        return nxb_int_as_float((0xfd << 23) - (nxb_float_as_int(m) & (0xff << 23)));
    }
    static NXB_DEV bool negative(float a) { return nxb_float_as_int(a) < 0; }
    static NXB_DEV float uniform(NxbXo& xo) { return (float)(nxb_xo_next(xo) >> 8) * (1.0f / 16777216.0f); }
};
template <> struct TolMath<double> {
    static NXB_DEV double expneg(double pen) { return exp(-pen); }
    static NXB_DEV double log2(double x) { return ::log2(x); }
    static NXB_DEV double exp2(double x) { return ::exp2(x); }
    static NXB_DEV double div(double a, double b) { return a / b; }
    static NXB_DEV double tiny() { return 5.421010862427522e-20; }
    static NXB_DEV double big() { return 1.8446744073709552e19; }
    static NXB_DEV double pow2_rescale(double m) {
        return nxb_ll_as_double((0x7fdLL << 52) - (nxb_double_as_ll(m) & (0x7ffLL << 52)));
    }
    static NXB_DEV bool negative(double a) { return nxb_double_as_ll(a) < 0; }
    static NXB_DEV double uniform(NxbXo& xo) {
        const uint32_t hi = nxb_xo_next(xo), lo = nxb_xo_next(xo);
        return (double)((((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11))) * (1.0 / 9007199254740992.0);
    }
};

// Exponential gap of the dominating Poisson process of a track, in rate-time units, mean
// gsc / ln 2.  f32 messages: the logarithm is taken in f32 (23-bit uniform, one random word per
// gap); as an f64 the value has 29 zero low mantissa bits: the unused 9 bits of the word and the
// track id go there, which makes equal event times rare (they are harmless between tracks; a
// candidate that ties with a live event of its own track is dropped by the walk).  f64 messages:
// 53-bit uniform and the f64 logarithm.
template <typename MSG> NXB_DEV double tol_gap(NxbXo& xo, float gsc, int tag);
template <> NXB_DEV double tol_gap<float>(NxbXo& xo, float gsc, int tag) {
    const unsigned w = nxb_xo_next(xo);
    const float f = ((float)(w >> 9) + 0.5f) * (1.0f / 8388608.0f);   // in (0,1), exact
    // the fast log2 has ~2^-22 absolute error, which only matters for f within 1e-6 of 1: clamp so
    // that the gap stays strictly positive (affects one draw in a million by < 1e-7 of a mean gap)
    const double g = (double)(fmaxf(-nxb_fast_log2f(f), 5.9604645e-8f) * gsc);
    return nxb_ll_as_double(nxb_double_as_ll(g) | (long long)(((w & 0x1ffu) << 6) | (unsigned)tag));
}
template <> NXB_DEV double tol_gap<double>(NxbXo& xo, float gsc, int) {
    const uint32_t hi = nxb_xo_next(xo), lo = nxb_xo_next(xo);
    return -::log2(nxb_u53(hi, lo)) * (double)gsc;
}

// what the lane jobs need of the kernel context
struct TolCtx {
    const NxbSweepParams* p;
    int E, P, K;
    SPtr<const unsigned char> cls;
    SPtr<const double> qm;
    SPtr<unsigned long long> s_off_on;
    unsigned k0, k1;
};

// shared-memory arrays of one unit
template <typename MSG, bool SPLIT> struct TolUnit {
    SPtr<int> uhdr;
    SPtr<AccRec<MSG> > acc;                 // [slot][K]
    SPtr<unsigned> newtol;
    SPtr<const int> poff, toff;
    SPtr<const uint4> nrec;                 // per edge: data masks (may be ON / OFF) and old tolerance states of
                                            // the lower node; its primary state | has primary events << 8 | poff << 9
    SPtr<const unsigned char> node_pri;
    SPtr<const double> ptm, btm;
    SPtr<const unsigned> pcode;
    unsigned bcode_off;
    NXB_DEV unsigned bedge(int i) const {
        if (SPLIT) return sm_ptr<const unsigned short>(bcode_off)[i];
        return sm_ptr<const unsigned>(bcode_off)[i] & 0xffffu;
    }
    NXB_DEV void bind(unsigned ub, const NxbWarpLayout& wl) {
        uhdr.off = ub + wl.off_uhdr; acc.off = ub + wl.off_acc; newtol.off = ub + wl.off_newtol;
        poff.off = ub + wl.off_poff; toff.off = ub + wl.off_toff; nrec.off = ub + wl.off_nrec;
        node_pri.off = ub + wl.off_node_pri; ptm.off = ub + wl.off_ptm; btm.off = ub + wl.off_btm;
        pcode.off = ub + wl.off_pcode; bcode_off = ub + wl.off_bcode;
    }
};

// state of a lane job between its phases
template <typename MSG> struct TolLane {
    int k;                  // track
    bool trk, accum;
    int bad;                // 1: infeasible, 2: pool overflow
    int cnt, nseg;          // live event records, edge records
    unsigned x_root;
    TolRec<MSG>* rec;       // this track's region of the pool
    NxbXo xo;               // per-track sequential stream
    MSG u0, u1; double pen; // message at the root after the upward tree pass
};

template <typename MSG> NXB_DEV MSG tol_ratio(MSG z, MSG o) {
    // P(ON) of a scale-free (OFF, ON) pair.  Hard constraints stay exact: z == 0 gives exactly 1
    // (a fast division does not guarantee x / x == 1), o == 0 exactly 0; (0, 0) reads as 0.
    if (z == (MSG)0) return (o > (MSG)0) ? (MSG)1 : (MSG)0;
    const MSG r = TolMath<MSG>::div(o, z + o);
    return r < (MSG)1 ? r : (MSG)1;
}

// ---------------------------------------------------------------------------------------------
// job prologue: which (unit, track), its pool region, its random stream
// ---------------------------------------------------------------------------------------------
template <typename MSG, bool SPLIT>
NXB_DEV void tol_begin(const TolCtx& c, TolLane<MSG>& L, TolUnit<MSG, SPLIT>& u, int job, int n_jobs,
                       unsigned ub0, unsigned char* gb0) {
    const NxbSweepParams& p = *c.p;
    const int K = c.K;
    const bool has_job = job < n_jobs;
    const int ju = has_job ? job / K : 0;
    L.k = has_job ? job - ju * K : 0;
    u.bind(ub0 + (unsigned)(ju * p.wl.bytes), p.wl);
    unsigned char* const ugb = gb0 + (long long)ju * p.gscratch_stride;
    L.rec = (TolRec<MSG>*)(ugb + p.g_off_ctm) + (size_t)L.k * p.gcapC;
    L.trk = has_job && u.uhdr[0] != 0;
    L.accum = L.trk && u.uhdr[6] != 0;
    L.bad = 0; L.cnt = 0; L.nseg = 0; L.x_root = 0u;
    L.u0 = (MSG)1; L.u1 = (MSG)1; L.pen = 0.0;
    NxbU4 r = nxb_philox_call((unsigned)u.uhdr[3], (unsigned)u.uhdr[5],
                              NXB_STREAM_BLK_FFBS | ((unsigned)L.k << 16), 0u, c.k0, c.k1);
    L.xo.s0 = r.x; L.xo.s1 = r.y; L.xo.s2 = r.z; L.xo.s3 = r.w | 1u;
}

// ---------------------------------------------------------------------------------------------
// W: walk over the points of the track, leafward to rootward.  Point order: edges in descending
// index, times descending within an edge; at equal times retained event, primary event, candidate.
// ---------------------------------------------------------------------------------------------
template <typename MSG, bool SPLIT>
NXB_DEV void tol_walk_up(const TolCtx& c, TolLane<MSG>& L, const TolUnit<MSG, SPLIT>& u) {
    typedef TolMath<MSG> M;
    const NxbSweepParams& p = *c.p;
    const NxbModelLayout& ml = p.ml;
    const int E = c.E, K = c.K, k = L.k;
    const int TCAP = p.gcapC;
    const SPtr<const NxbEdgeRec> erec = {(unsigned)ml.off_edgerec};
    const SPtr<const NxbEdgeLoc> eloc = {(unsigned)ml.off_edgeloc};
    const SPtr<const double> cumw = {(unsigned)ml.off_cumw};
    TolRec<MSG>* const rec = L.rec;
    // thinning acceptance as integer thresholds: accept iff u32 < thr (an acceptance probability of
    // exactly 1 becomes 1 - 2^-32)
    const unsigned thr_o0 = (unsigned)fmin(ml.acc_blk[0] * 4294967296.0, 4294967295.0);
    const unsigned thr_o1 = (unsigned)fmin(ml.acc_blk[1] * 4294967296.0, 4294967295.0);
    const unsigned thr_w0 = (unsigned)fmin(ml.acc_blk[2] * 4294967296.0, 4294967295.0);
    const unsigned thr_w1 = (unsigned)fmin(ml.acc_blk[3] * 4294967296.0, 4294967295.0);
    const MSG a_on = (MSG)ml.a_on, a_off = (MSG)ml.a_off;
    const float gsc = ml.gsc_sigma;
    const double sum_w = ml.sum_w;
    const int olo = L.trk ? u.toff[k] : 0, ohi = L.trk ? u.toff[k + 1] : 0;
    NxbXo xo = L.xo;
    int cnt = 0, nseg = 0, bad = 0;

    // the three streams of points: next retained event (descending through the (edge, time) sorted
    // run of this track), next primary event (descending through the unit's list), next candidate
    int io = ohi - 1, e_r = -1; double t_r = 0.0;
    if (io >= olo) { e_r = (int)u.bedge(io); t_r = u.btm[io]; }
    int ip = L.trk ? u.poff[E] - 1 : -1, e_p = -1; double t_p = 0.0;
    if (ip >= 0) { e_p = (int)(u.pcode[ip] & 0xffffu); t_p = u.ptm[ip]; }
    double sig = 0.0; int e_c = -1; double t_c = 0.0; bool c_void = false;
    // next arrival of the dominating process: position on the rate-time line -> (edge, time)
    auto next_cand = [&]() {
        sig += tol_gap<MSG>(xo, gsc, k + 1);
        e_c = -1; c_void = false;
        if (sig < sum_w) {
            int j = 0;                                  // number of edges (walk order) that end at or before sig
            for (int st = ml.loc_steps; st >= 1; st >>= 1) {
                const int jj = j + st;
                if (jj <= E && cumw[jj - 1] <= sig) j = jj;
            }
            if (j < E) {
                const int e = E - 1 - j;
                const NxbEdgeLoc el = eloc[e];
                const NxbEdgeRec er = erec[e];
                double t = er.tmb - (sig - el.cum_excl) * el.inv_rate;
                // rounding may put it on an end of the edge: such a point is kept in the order but is
                // not an event (probability ~2^-50)
                if (!(t > er.tma)) { t = er.tma; c_void = true; }
                if (!(t < er.tmb)) { t = er.tmb; c_void = true; }
                e_c = e; t_c = t;
            }
        }
    };
    if (L.trk) next_cand();

    // the open edge: columns (z_c, o_c) = (OFF, ON) message just above the last point given that
    // the bottom of the edge is in state c, each with its own log2 scale
    int ecur = -1, s = 0, nl = 0;
    unsigned xold = 0u;
    bool own = false;
    MSG z0 = 0, o0 = 0, z1 = 0, o1 = 0, L0 = 0, L1 = 0;
    double pen = 0.0, tcur = 0.0, qmr = 0.0, tma_cur = 0.0, rate_cur = 0.0, t_last = NXB_INF;

    // penalty exponent accumulated since the last live event -> the ON components.  Where OFF is
    // impossible in a column the factor is a pure rescaling of that column and goes to its log scale
    // (so a long own-class stretch cannot underflow ON to 0); where it is impossible in both columns
    // it cancels.
    auto apply_pen = [&]() {
        if (pen != 0.0) {
            if (z0 != (MSG)0 || z1 != (MSG)0) {
                const MSG f = M::expneg(pen);
                const MSG lf = (MSG)(-1.4426950408889634 * pen);
                if (z0 != (MSG)0) o0 *= f; else L0 += lf;
                if (z1 != (MSG)0) o1 *= f; else L1 += lf;
            }
            pen = 0.0;
        }
    };
    auto close_seg = [&]() {
        if (nl > 0) {
            pen += qmr * (tcur - tma_cur);
            if (own) { z0 = 0; z1 = 0; }
            apply_pen();
            const MSG S0 = z0 + o0, S1 = z1 + o1;
            MSG ls = 0;
            if (S0 > (MSG)0 && S1 > (MSG)0) ls = (L1 + M::log2(S1)) - (L0 + M::log2(S0));
            else if (S1 > (MSG)0) ls = (MSG)NXB_INF;
            else if (S0 > (MSG)0) ls = (MSG)(-NXB_INF);
            else TOL_BAD(bad, 1);
            if (cnt + nseg < TCAP) {
                TolRec<MSG> sg;
                sg.s.e = (unsigned short)ecur; sg.s.n = (unsigned short)nl;
                sg.a = tol_ratio<MSG>(z0, o0); sg.b = tol_ratio<MSG>(z1, o1);
                nxb_set_lscale(sg, ls);
                rec[TCAP - 1 - nseg] = sg;
            } else bad = 2;
            ++nseg;
        }
    };

    while (TOL_ANY((e_r & e_p & e_c) >= 0)) {
        const int e_n = e_r > e_p ? (e_r > e_c ? e_r : e_c) : (e_p > e_c ? e_p : e_c);
        if (e_n >= 0) {
            const double tr = (e_r == e_n) ? t_r : -NXB_INF;
            const double tp = (e_p == e_n) ? t_p : -NXB_INF;
            const double tc = (e_c == e_n) ? t_c : -NXB_INF;
            const bool is_r = tr >= tp && tr >= tc;
            const bool is_p = !is_r && tp >= tc;
            const double tn = is_r ? tr : (is_p ? tp : tc);
            if (e_n != ecur) {
                if (ecur >= 0) close_seg();
                const NxbEdgeRec er = erec[e_n];
                const uint4 nr = u.nrec[e_n];
                ecur = e_n; tcur = er.tmb; tma_cur = er.tma; rate_cur = er.rate;
                s = (int)(nr.w & 0xffu);
                xold = (nr.z >> k) & 1u;
                own = (c.cls[s] == k);
                qmr = c.qm[s * K + k] * rate_cur;
                z0 = 1; o0 = 0; z1 = 0; o1 = 1; L0 = 0; L1 = 0;
                pen = 0.0; nl = 0; t_last = NXB_INF;
            }
            // chunking.py:67-79 on the stretch just crossed: ON pays the rate into class k, own class forces ON
            pen += qmr * (tcur - tn);
            tcur = tn;
            if (own) { z0 = 0; z1 = 0; }
            if (is_p) {
                s = (int)((u.pcode[ip] >> 16) & 0xffu);      // rootward: the state before the event
                own = (c.cls[s] == k);
                qmr = c.qm[s * K + k] * rate_cur;
                --ip;
                if (ip >= 0) { e_p = (int)(u.pcode[ip] & 0xffffu); t_p = u.ptm[ip]; } else e_p = -1;
            } else {
                bool live = true;
                if (is_r) {
                    xold ^= 1u;
                    --io;
                    if (io >= olo) { e_r = (int)u.bedge(io); t_r = u.btm[io]; } else e_r = -1;
                } else {
                    // poisson.py:111-116: local rate = omega_ctx - r_x, thinned from the dominating rate
                    const unsigned thr = own ? (xold ? thr_w1 : thr_w0) : (xold ? thr_o1 : thr_o0);
                    live = (nxb_xo_next(xo) < thr) && !c_void && (tn != t_last);
                    next_cand();
                }
                if (live) {
                    apply_pen();
                    // the components only ever shrink: keep them away from underflow
                    {
                        const MSG m0 = z0 > o0 ? z0 : o0, m1 = z1 > o1 ? z1 : o1;
                        if (m0 < M::tiny() && m0 > (MSG)0) { z0 *= M::big(); o0 *= M::big(); L0 -= (MSG)64; }
                        if (m1 < M::tiny() && m1 > (MSG)0) { z1 *= M::big(); o1 *= M::big(); L1 -= (MSG)64; }
                    }
                    if (cnt + nseg < TCAP) {
                        TolRec<MSG> ev;
                        ev.t = tn;
                        const MSG r0 = tol_ratio<MSG>(z0, o0);
                        ev.a = own ? -r0 : r0;
                        ev.b = tol_ratio<MSG>(z1, o1);
                        tol_rec_pad(ev);
                        rec[cnt] = ev;
                    } else bad = 2;
                    ++cnt; ++nl; t_last = tn;
                    // uniformized 2x2 matrix of this context (poisson.py:92-96)
                    if (own) { z0 = (MSG)0.5 * z0 + (MSG)0.5 * o0; z1 = (MSG)0.5 * z1 + (MSG)0.5 * o1; }
                    else {
                        const MSG nz0 = z0 + a_on * (o0 - z0), nz1 = z1 + a_on * (o1 - z1);
                        o0 = o0 + a_off * (z0 - o0); o1 = o1 + a_off * (z1 - o1);
                        z0 = nz0; z1 = nz1;
                    }
                }
            }
        }
    }
    if (ecur >= 0) close_seg();
    L.xo = xo; L.cnt = cnt; L.nseg = nseg; L.bad |= bad;
}

// ---------------------------------------------------------------------------------------------
// TU: upward tree pass, edges in descending index (children before parents); node accumulators
// live in a handful of reusable slots (nodes are numbered depth-first), first-child chains stay in
// registers.  Then the root draw (prior: toymodel.py:17-27 get_blink_distn).
// ---------------------------------------------------------------------------------------------
template <typename MSG, bool SPLIT>
NXB_DEV void tol_tree_up(const TolCtx& c, TolLane<MSG>& L, const TolUnit<MSG, SPLIT>& u) {
    typedef TolMath<MSG> M;
    const NxbSweepParams& p = *c.p;
    const NxbModelLayout& ml = p.ml;
    const int E = c.E, K = c.K, k = L.k;
    const int TCAP = p.gcapC;
    const SPtr<const NxbEdgeRec> erec = {(unsigned)ml.off_edgerec};
    TolRec<MSG>* const rec = L.rec;
    const int nseg = (L.bad & 2) ? 0 : L.nseg;
    int is = 0, seg_e = -1, bad = 0;
    TolRec<MSG> sg;
    sg.a = 0; sg.b = 0; sg.s.e = 0; sg.s.n = 0; nxb_set_lscale(sg, (MSG)0);
    if (nseg > 0) { sg = rec[TCAP - 1]; seg_e = sg.s.e; }
    MSG u0 = 1, u1 = 1;
    double pen = 0.0;
    bool chained = false;
    for (int e = E - 1; e >= 0; --e) {
        const NxbEdgeRec er = erec[e];
        if (L.trk) {
            const uint4 nr = u.nrec[e];
            if (!chained) {
                u0 = 1; u1 = 1; pen = 0.0;
                if (er.sb_slot >= 0) { const AccRec<MSG> a = u.acc[er.sb_slot * K + k]; u0 = a.u0; u1 = a.u1; pen = a.pen; }
            }
            if (!((nr.x >> k) & 1u)) u1 = 0;
            if (!((nr.y >> k) & 1u)) u0 = 0;
            int s = (int)(nr.w & 0xffu);
            if (seg_e == e) {
                // message at the bottom, then T_e from the record: column c has weight 2^(+-lscale)
                MSG m1 = u1;
                if (pen != 0.0) { if (u0 != (MSG)0) m1 *= M::expneg(pen); pen = 0.0; }
                const MSG ls = nxb_get_lscale(sg);
                const MSG w0 = u0 * M::exp2(ls > (MSG)0 ? -ls : (MSG)0), w1 = m1 * M::exp2(ls < (MSG)0 ? ls : (MSG)0);
                const MSG t00 = ((MSG)1 - sg.a) * w0, t10 = sg.a * w0;      // top OFF / ON, bottom OFF
                const MSG t01 = ((MSG)1 - sg.b) * w1, t11 = sg.b * w1;      // top OFF / ON, bottom ON
                u0 = t00 + t01; u1 = t10 + t11;
                // P(bottom ON | top state), exact under hard constraints
                const MSG p0 = (t01 == (MSG)0) ? (MSG)0 : ((t00 == (MSG)0) ? (MSG)1 : fmin_msg(M::div(t01, u0), (MSG)1));
                const MSG p1 = (t11 == (MSG)0) ? (MSG)0 : ((t10 == (MSG)0) ? (MSG)1 : fmin_msg(M::div(t11, u1), (MSG)1));
                nxb_store_ab(&rec[TCAP - 1 - is], p0, p1);
                ++is;
                if (is < nseg) { sg = rec[TCAP - 1 - is]; seg_e = sg.s.e; } else seg_e = -1;
            } else if (!((nr.w >> 8) & 1u)) {
                // no point of this track on the edge: diagonal
                if (c.cls[s] == k) u0 = 0;
                pen += c.qm[s * K + k] * (er.rate * (er.tmb - er.tma));
            } else {
                // primary events but no live event of this track: diagonal, stretch by stretch
                double tcur = er.tmb;
                const int lo = (int)(nr.w >> 9);
                for (int i = u.poff[e + 1] - 1; i >= lo; --i) {
                    const double t = u.ptm[i];
                    if (c.cls[s] == k) u0 = 0;
                    pen += c.qm[s * K + k] * er.rate * (tcur - t);
                    tcur = t;
                    s = (int)((u.pcode[i] >> 16) & 0xffu);
                }
                if (c.cls[s] == k) u0 = 0;
                pen += c.qm[s * K + k] * er.rate * (tcur - er.tma);
            }
            // fold into the parent's accumulator (one normalisation)
            const int ai = er.sa_slot * K + k;
            if (!(er.flags & NXB_EF_FIRST)) { const AccRec<MSG> a = u.acc[ai]; u0 *= a.u0; u1 *= a.u1; pen += a.pen; }
            MSG m2 = u0 > u1 ? u0 : u1;
#if defined(TOL_DEBUG)
            if (!(m2 > (MSG)0)) fprintf(stderr, "  k=%d e=%d seg_e=%d is=%d nseg=%d nr=(%x %x %x %x) flags=%d\n", k, e, seg_e, is, nseg, nr.x, nr.y, nr.z, nr.w, er.flags);
#endif
            if (!(m2 > (MSG)0)) { TOL_BAD(bad, 2); m2 = 1; }
            const MSG r2 = M::pow2_rescale(m2);
            u0 *= r2; u1 *= r2;
            chained = (er.flags & NXB_EF_CHAIN_UP) != 0;
            if (!chained) { AccRec<MSG> n; n.pen = pen; n.u0 = u0; n.u1 = u1; nxb_acc_pad(n); u.acc[ai] = n; }
        }
    }
    // (edge 0 hangs from the root and is the last to fold into it: the root message is in registers)
    unsigned x = 0u;
    if (L.trk) {
        double w1 = (double)u1 * (u0 != (MSG)0 ? exp(-pen) : 1.0) * ml.p_on;
        double w0 = (double)u0 * ml.p_off;
        if (!(((unsigned)u.uhdr[8] >> k) & 1u)) w1 = 0.0;
        if (!(((unsigned)u.uhdr[9] >> k) & 1u)) w0 = 0.0;
        if (c.cls[u.node_pri[0]] == k) w0 = 0.0;
        if (!(w0 + w1 > 0.0)) TOL_BAD(bad, 3);
        x = (nxb_u32(nxb_xo_next(L.xo)) * (w0 + w1) < w1) ? 1u : 0u;
        if (x) nxb_atomic_or_shared(&u.newtol[0], 1u << k);
    }
    L.x_root = x;
    L.bad |= bad;
}

// ---------------------------------------------------------------------------------------------
// TD: downward tree pass, edges in ascending index: the state at every node.  An edge without live
// events keeps its state; its OFF time is booked here (summary.py:123-131).
// ---------------------------------------------------------------------------------------------
template <typename MSG, bool SPLIT>
NXB_DEV void tol_tree_down(const TolCtx& c, TolLane<MSG>& L, const TolUnit<MSG, SPLIT>& u) {
    typedef TolMath<MSG> M;
    const NxbSweepParams& p = *c.p;
    const NxbModelLayout& ml = p.ml;
    const NxbSummaryLayout& sl = p.sl;
    const int E = c.E, K = c.K, P = c.P, k = L.k;
    const int TCAP = p.gcapC;
    const SPtr<const NxbEdgeRec> erec = {(unsigned)ml.off_edgerec};
    TolRec<MSG>* const rec = L.rec;
    const int nseg = (L.bad & 2) ? 0 : L.nseg;
    int is = nseg - 1, seg_e = -1;
    MSG sa = 0, sb = 0;
    if (is >= 0) { const TolRec<MSG> sg = rec[TCAP - 1 - is]; seg_e = sg.s.e; sa = sg.a; sb = sg.b; }
    unsigned x = L.x_root;
    for (int e = 0; e < E; ++e) {
        const NxbEdgeRec er = erec[e];
        if (L.trk) {
            if (!(er.flags & NXB_EF_CHAIN_DOWN)) x = (u.newtol[er.va] >> k) & 1u;
            if (seg_e == e) {
                const MSG pp = x ? sb : sa;
                x = (M::uniform(L.xo) < pp) ? 1u : 0u;
                --is;
                if (is >= 0) { const TolRec<MSG> sg = rec[TCAP - 1 - is]; seg_e = sg.s.e; sa = sg.a; sb = sg.b; } else seg_e = -1;
            } else if (!x && L.accum) {
                const uint4 nr = u.nrec[e];
                if (!((nr.w >> 8) & 1u)) {
                    nxb_atomic_add_f64(p.dwell + (sl.d_off + (e * P + (int)(nr.w & 0xffu)) * K + k), er.tmb - er.tma);
                } else {
                    double tcur = er.tma;
                    int s = u.node_pri[er.va];
                    const int hi = u.poff[e + 1];
                    for (int i = (int)(nr.w >> 9); i < hi; ++i) {
                        const double t = u.ptm[i];
                        nxb_atomic_add_f64(p.dwell + (sl.d_off + (e * P + s) * K + k), t - tcur);
                        tcur = t;
                        s = (int)((u.pcode[i] >> 24) & 0xffu);
                    }
                    nxb_atomic_add_f64(p.dwell + (sl.d_off + (e * P + s) * K + k), er.tmb - tcur);
                }
            }
            if (x) nxb_atomic_or_shared(&u.newtol[e + 1], 1u << k);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// WD: the live events, edge by edge in ascending index and ascending time (the records are read
// back in reverse), given the states at both ends of the edge.  Survivors (real transitions) are
// written over the consumed records from the top of the live run downwards.  The OFF-time and
// transition statistics of these edges are fused in (summary.py:123-131,234-243).
// ---------------------------------------------------------------------------------------------
template <typename MSG, bool SPLIT>
NXB_DEV void tol_walk_down(const TolCtx& c, TolLane<MSG>& L, const TolUnit<MSG, SPLIT>& u) {
    typedef TolMath<MSG> M;
    const NxbSweepParams& p = *c.p;
    const NxbModelLayout& ml = p.ml;
    const NxbSummaryLayout& sl = p.sl;
    const int K = c.K, P = c.P, k = L.k;
    const int TCAP = p.gcapC;
    const SPtr<const NxbEdgeRec> erec = {(unsigned)ml.off_edgerec};
    TolRec<MSG>* const rec = L.rec;
    const MSG a_on = (MSG)ml.a_on, a_off1 = (MSG)(1.0 - ml.a_off);
    const bool ok = L.trk && !(L.bad & 2);
    const int cnt = ok ? L.cnt : 0;
    int ir = cnt - 1, wtop = cnt - 1, is = L.nseg - 1;
    int left = 0, e = 0, s = 0, ip = 0, ipe = 0, n_on = 0, n_off = 0, bad = 0;
    unsigned x = 0u, xc = 0u;
    double tcur = 0.0, tmb = 0.0, offtime = 0.0;
    TolRec<MSG> pf;
    pf.t = 0.0; pf.a = 0; pf.b = 0; tol_rec_pad(pf);
    if (ir >= 0) pf = rec[ir];
    while (TOL_ANY(ir >= 0)) {
        if (ir >= 0) {
            if (left == 0) {
                const TolRec<MSG> sg = rec[TCAP - 1 - is];
                --is;
                e = sg.s.e; left = sg.s.n;
                const NxbEdgeRec er = erec[e];
                x = (u.newtol[er.va] >> k) & 1u;
                xc = (u.newtol[e + 1] >> k) & 1u;
                s = u.node_pri[er.va];
                ip = u.poff[e]; ipe = u.poff[e + 1];
                tcur = er.tma; tmb = er.tmb;
                offtime = 0.0; n_on = 0; n_off = 0;
            }
            const TolRec<MSG> ev = pf;
            --ir; --left;
            // (the record at ir is in registers before a survivor may overwrite it: wtop >= ir)
            if (ir >= 0) pf = rec[ir];
            if (L.accum) {
                while (ip < ipe && u.ptm[ip] < ev.t) {       // primary events before this one: close the OFF stretch
                    const double t = u.ptm[ip];
                    if (!x) offtime += t - tcur;
                    if (offtime > 0.0) nxb_atomic_add_f64(p.dwell + (sl.d_off + (e * P + s) * K + k), offtime);
                    offtime = 0.0; tcur = t;
                    s = (int)((u.pcode[ip] >> 24) & 0xffu);
                    ++ip;
                }
                if (!x) offtime += ev.t - tcur;
                tcur = ev.t;
            }
            {
                const bool own = M::negative(ev.a);
                const MSG r1 = xc ? ev.b : (own ? -ev.a : ev.a);
                const MSG p1 = own ? (x ? (MSG)1 : (MSG)0.5) : (x ? a_off1 : a_on);
                const MSG w1 = p1 * r1, w0 = ((MSG)1 - p1) * ((MSG)1 - r1);
                if (!(w0 + w1 > (MSG)0)) TOL_BAD(bad, 4);
                const unsigned xn = (M::uniform(L.xo) * (w0 + w1) < w1) ? 1u : 0u;
                if (xn != x) {
                    TolRec<MSG> sv;
                    sv.t = ev.t; sv.b = 0; tol_rec_pad(sv);
                    sv.code = (unsigned)e | (xn ? CE_NEW1 : 0u);
                    rec[wtop] = sv;
                    --wtop;
                    if (xn) ++n_on; else ++n_off;
                    x = xn;
                }
            }
            if (left == 0) {
                // the rest of the edge
                if (x != xc) TOL_BAD(bad, 5);
                if (L.accum) {
                    while (ip < ipe) {
                        const double t = u.ptm[ip];
                        if (!x) offtime += t - tcur;
                        if (offtime > 0.0) nxb_atomic_add_f64(p.dwell + (sl.d_off + (e * P + s) * K + k), offtime);
                        offtime = 0.0; tcur = t;
                        s = (int)((u.pcode[ip] >> 24) & 0xffu);
                        ++ip;
                    }
                    if (!x) offtime += tmb - tcur;
                    if (offtime > 0.0) nxb_atomic_add_f64(p.dwell + (sl.d_off + (e * P + s) * K + k), offtime);
                    // packed word per edge (low: off->on, high: on->off), updated by halves with native
                    // 32-bit shared atomics; a CTA books far fewer than 2^32 per launch
                    if (n_on) nxb_atomic_add_shared((unsigned*)&c.s_off_on[e], (unsigned)n_on);
                    if (n_off) nxb_atomic_add_shared((unsigned*)&c.s_off_on[e] + 1, (unsigned)n_off);
                }
            }
        }
    }
    if (L.trk) {
        u.uhdr[10 + k] = (cnt - 1 - wtop) | (cnt << 16);   // survivors of this track | live events
        if (L.bad | bad) nxb_atomic_or_shared((unsigned*)&u.uhdr[4], (unsigned)(L.bad | bad));
    }
}

}  // namespace nxb
