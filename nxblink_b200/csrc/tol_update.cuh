// TOLERANCE TRACK UPDATES (raoteh.py:413-432) of one unit: all K tracks at once, one lane per
// (unit, track).  Replaces, per track, poisson.py:176-199 (candidate events), chunking.py:39-155
// (blinking chunk tree), chunking.py:264-289 -> nxmctree sample_history (FFBS) and the tolerance
// half of summary.py:190-243.
//
// The per-lane code below is templated on the message type MSG (float: the fast path; double: the
// reference's precision end to end -- messages, accumulators, exp(-penalty), gap logarithm,
// backward-sampling ratios and uniforms, chunking.py:32-36,264-289) and compiles for the device
// (the product) and, for the CPU simulation under oracle/hostsim (test infrastructure), for the host.
//
// (Round 2 also tried to split this walk into an events-only walk with per-edge 2x2 transfer
// records plus warp-uniform tree passes -- commit 15b8732.  It was exact (same tests) but executed
// 33.3 k warp-instructions per unit against 27.1 k here: profiles/r2_eventwalk_*)
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

// One record per live event in the per-(unit, track) region of the L2-resident pool: time,
// q = l1 / (l0 + l1) of the (OFF, ON) message below the event (the backward draw only needs the
// ratio), edge.  The downward pass reuses the records it has consumed for the survivors (e then
// also carries CE_NEW1), so a track touches only ceil(16 * live / 128) lines of the pool per sweep.
template <typename MSG> struct LiveEv;
template <> struct __align__(16) LiveEv<float> { double t; float q; unsigned short e, pad; };
template <> struct __align__(16) LiveEv<double> { double t; double q; unsigned short e, pad; unsigned pad2; double pad3; };
// Accumulator of a node for one track: product of the (OFF, ON) messages of the children folded
// so far (rescaled by powers of two) and the penalty exponent still to be applied to ON.
template <typename MSG> struct AccRec;
template <> struct __align__(16) AccRec<float> { double pen; float u0, u1; };
template <> struct __align__(16) AccRec<double> { double pen; double u0, u1; double pad; };
NXB_DEV void nxb_acc_pad(AccRec<float>&) {}
NXB_DEV void nxb_acc_pad(AccRec<double>& n) { n.pad = 0.0; }
// pool records live in L2 (global scratch): access them with the cache-global operators, which do
// not allocate in the small L1 that is left beside 212 KB of shared memory
#if defined(__CUDACC__)
NXB_DEV LiveEv<float> nxb_ev_load(const LiveEv<float>* p) {
    const int4 v = __ldcg(reinterpret_cast<const int4*>(p));
    LiveEv<float> e;
    e.t = __hiloint2double(v.y, v.x); e.q = __int_as_float(v.z);
    e.e = (unsigned short)(v.w & 0xffff); e.pad = (unsigned short)((unsigned)v.w >> 16);
    return e;
}
NXB_DEV void nxb_ev_store(LiveEv<float>* p, const LiveEv<float>& e) {
    int4 v;
    v.x = __double2loint(e.t); v.y = __double2hiint(e.t); v.z = __float_as_int(e.q);
    v.w = (int)((unsigned)e.e | ((unsigned)e.pad << 16));
    __stcg(reinterpret_cast<int4*>(p), v);
}
NXB_DEV LiveEv<double> nxb_ev_load(const LiveEv<double>* p) { return *p; }
NXB_DEV void nxb_ev_store(LiveEv<double>* p, const LiveEv<double>& e) { *p = e; }
#else
template <typename MSG> NXB_DEV LiveEv<MSG> nxb_ev_load(const LiveEv<MSG>* p) { return *p; }
template <typename MSG> NXB_DEV void nxb_ev_store(LiveEv<MSG>* p, const LiveEv<MSG>& e) { *p = e; }
#endif
NXB_DEV void nxb_ev_pad(LiveEv<float>&) {}
NXB_DEV void nxb_ev_pad(LiveEv<double>& e) { e.pad2 = 0u; e.pad3 = 0.0; }

template <typename MSG> struct TolMath;
template <> struct TolMath<float> {
    static NXB_DEV float a_on(const NxbModelLayout& ml) { return ml.a_on_f; }
    static NXB_DEV float a_off(const NxbModelLayout& ml) { return ml.a_off_f; }
    static NXB_DEV float a_off1(const NxbModelLayout& ml) { return ml.a_off1_f; }
    // exp(-pen), never exactly 0: zeros in a message are HARD constraints (data, own class); a
    // penalty beyond the f32 range (exponent > 87) must not create one, or a unit whose OFF state is
    // excluded by data would turn infeasible.  Messages are rescaled by powers of two at every fold,
    // so the smallest normal number is as good as the true 1e-40.
    static NXB_DEV float expneg(double pen) { return fmaxf(nxb_fast_expf((float)(-pen)), 1.17549435e-38f); }
    static NXB_DEV float div(float a, float b) { return nxb_fast_divf(a, b); }
    // 2^-(floor(log2 m) + 1) for a positive finite m: exact power-of-two rescaling into [0.5, 1)
    // without a division; preserves exact zeros of the other component.
    static NXB_DEV float pow2_rescale(float m) {
        return nxb_int_as_float((0xfd << 23) - (nxb_float_as_int(m) & (0xff << 23)));
    }
    static NXB_DEV float uniform(NxbXo& xo) { return (float)(nxb_xo_next(xo) >> 8) * (1.0f / 16777216.0f); }
};
template <> struct TolMath<double> {
    static NXB_DEV double a_on(const NxbModelLayout& ml) { return ml.a_on; }
    static NXB_DEV double a_off(const NxbModelLayout& ml) { return ml.a_off; }
    static NXB_DEV double a_off1(const NxbModelLayout& ml) { return 1.0 - ml.a_off; }
    static NXB_DEV double expneg(double pen) { return fmax(exp(-pen), 2.2250738585072014e-308); }
    static NXB_DEV double div(double a, double b) { return a / b; }
    static NXB_DEV double pow2_rescale(double m) {
        return nxb_ll_as_double((0x7fdLL << 52) - (nxb_double_as_ll(m) & (0x7ffLL << 52)));
    }
    static NXB_DEV double uniform(NxbXo& xo) {
        const uint32_t hi = nxb_xo_next(xo), lo = nxb_xo_next(xo);
        return (double)((((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11))) * (1.0 / 9007199254740992.0);
    }
};

// Exponential gap of the dominating Poisson process of an edge, mean gsc / ln 2 (time units).
// f32 messages: the logarithm is taken in f32 (23-bit uniform, one random word per gap); as an f64
// the value has 29 zero low mantissa bits: the unused 9 bits of the word and the track id go
// there, which makes equal event times rare (they are harmless between tracks; a candidate that
// ties with a retained event of its own track is dropped by the walk).  f64 messages: 53-bit
// uniform and the f64 logarithm.
template <typename MSG> NXB_DEV double tol_gap(NxbXo& xo, float gsc, int tag);
template <> NXB_DEV double tol_gap<float>(NxbXo& xo, float gsc, int tag) {
    const unsigned w = nxb_xo_next(xo);
    const float f = ((float)(w >> 9) + 0.5f) * (1.0f / 8388608.0f);   // in (0,1), exact
    // the fast log2 has ~2^-22 absolute error, which only matters for f within 1e-6 of 1: clamp so
    // that the gap stays strictly positive (affects one draw in a million by < 1e-7 of a mean gap)
    const double g = (double)(fmaxf(-nxb_fast_log2f(f), 5.9604645e-8f) * gsc);
    // low mantissa bits: the 9 bits of the word that the uniform did not use, then the track id
    return nxb_ll_as_double(nxb_double_as_ll(g) | (long long)(((w & 0x1ffu) << 6) | (unsigned)tag));
}
template <> NXB_DEV double tol_gap<double>(NxbXo& xo, float gsc, int) {
    const uint32_t hi = nxb_xo_next(xo), lo = nxb_xo_next(xo);
    return -::log2(nxb_u53(hi, lo)) * (double)gsc;
}

// max / min without the NaN handling of fmax / fmin (event times are never NaN)
NXB_DEV double dmax2(double a, double b) { return a > b ? a : b; }
NXB_DEV double dmin2(double a, double b) { return a < b ? a : b; }

// what the lane jobs need of the kernel context
struct TolCtx {
    const NxbSweepParams* p;
    int E, P, K;
    SPtr<const unsigned char> cls;
    SPtr<const double> qm;
    SPtr<const int> slot;
    SPtr<unsigned long long> s_off_on;
    unsigned k0, k1;
};

// shared-memory arrays of one unit
template <typename MSG, bool SPLIT> struct TolUnit {
    SPtr<int> uhdr;
    SPtr<AccRec<MSG> > acc;                 // [slot][K]
    SPtr<unsigned> newtol;
    SPtr<const int> poff, toff;
    SPtr<const uint4> nrec;                 // per edge: data masks (may be ON / OFF), old tolerance states and
                                            // primary state of the lower node | first primary event << 8
    SPtr<const unsigned char> node_pri;
    SPtr<const double> ptm, btm;
    SPtr<const unsigned> pcode;
    unsigned bcode_off;
    NXB_DEV unsigned bedge(int i) const {       // edge of retained event i (SPLIT: u16 edges only)
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
    int cnt;                // live events pushed by the upward pass
    unsigned long long evm; // edges on which this track has a live event (bit e; E <= 64 only)
    LiveEv<MSG>* cev;       // this track's region of the pool
    NxbXo xo;               // per-track sequential stream: candidate gaps, thinning and backward-sampling decisions
};

// ---------------------------------------------------------------------------------------------
// job prologue: which (unit, track), its pool region, its random stream.  Job j works on track
// j % K of unit j / K of a pool of units resident in this CTA's shared memory (regions
// ub0 + u * wl.bytes, global scratch gb0 + u * gscratch_stride).
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
    // live foreground events of this sweep: global (L2-resident) scratch, one fixed region per
    // (unit, track), filled by the upward pass and consumed backwards by the downward pass
    L.cev = (LiveEv<MSG>*)(ugb + p.g_off_ctm) + (size_t)L.k * p.gcapC;
    L.trk = has_job && u.uhdr[0] != 0;
    L.accum = L.trk && u.uhdr[6] != 0;
    L.bad = 0; L.cnt = 0; L.evm = 0ull;
    NxbU4 r = nxb_philox_call((unsigned)u.uhdr[3], (unsigned)u.uhdr[5],
                              NXB_STREAM_BLK_FFBS | ((unsigned)L.k << 16), 0u, c.k0, c.k1);
    L.xo.s0 = r.x; L.xo.s1 = r.y; L.xo.s2 = r.z; L.xo.s3 = r.w | 1u;
}

// ---------------------------------------------------------------------------------------------
// B1: upward pass.  Every lane walks its own track through the edges in descending index
// (children before parents; node accumulators live in a handful of reusable slots because nodes
// are numbered depth-first), one boundary per iteration: a retained event of the old trajectory, a
// candidate, a primary event, or the top of the edge.  Candidates are the arrivals of the
// dominating Poisson process of the edge, generated backwards in time by exponential gaps
// (poisson.py:22-50 draws count + uniform times; same process), then thinned to the local rate
// (poisson.py:111-116).  Only accepted ("live") events are stored.  Then the root draw.
// Returns the state drawn at the root.
// ---------------------------------------------------------------------------------------------
template <typename MSG, bool SPLIT>
NXB_DEV unsigned tol_up(const TolCtx& c, TolLane<MSG>& L, const TolUnit<MSG, SPLIT>& u) {
    typedef TolMath<MSG> M;
    const NxbSweepParams& p = *c.p;
    const NxbModelLayout& ml = p.ml;
    const int E = c.E, K = c.K, k = L.k;
    const int TCAP = p.gcapC;
    const bool trk = L.trk;
    const SPtr<const NxbEdgeRec> erec = {(unsigned)ml.off_edgerec};
    LiveEv<MSG>* const cev = L.cev;
    // retained events of track k are the contiguous run [olo, ohi) of the (track, edge, time)
    // sorted list (offsets per track: blink_setup)
    const int olo = trk ? u.toff[k] : 0, ohi = trk ? u.toff[k + 1] : 0;
    NxbXo xo = L.xo;
    int bad = 0, cnt = 0;
    unsigned long long evm = 0ull;
    {
        // Candidate rate = omega - (rate out of the OLD state), poisson.py:176-199: piecewise constant
        // along the old trajectory, 0 in own-class segments (inert events, tol_tables.h).  The
        // arrivals are exponential gaps at the rate of the current piece, restarted where the
        // piece changes (a retained event, a primary event that changes the own flag): no thinning.
        const float gf_off = ml.gf_off, gf_on = ml.gf_on;
        const MSG a_on = M::a_on(ml), a_off = M::a_off(ml);
        int e = trk ? E - 1 : -1;
        int io = ohi - 1;              // cursor into the retained events (descending)
        unsigned inited = 0u;          // live accumulator slots of this lane
        bool fresh = true;
        double tma = 0.0, rate = 0.0, tcur = 0.0, pen = 0.0;
        MSG u0 = 1, u1 = 1;            // (OFF, ON) message at the current point of the walk
        double tcand = -NXB_INF;
        float gsc = 0.f;
        int s = 0, ip = trk ? u.poff[E] - 1 : 0, ip0 = 0, sa_slot = 0, sb_slot = -1;
        unsigned xold = 0u;
        bool gap_due = false;
        while (TOL_ANY(e >= 0)) {
            if (e >= 0) {
                // one gap is due: at the bottom of a fresh edge, or after the candidate consumed by the
                // previous step (drawn here so that the walk has a single copy of the gap code)
                bool draw_gap = gap_due;
                gap_due = false;
                if (fresh) {
                    const NxbEdgeRec er = erec[e];
                    sb_slot = er.sb_slot; sa_slot = er.sa_slot;
                    tma = er.tma; rate = er.rate; tcur = er.tmb;
                    u0 = 1; u1 = 1; pen = 0.0;
                    if (sb_slot >= 0 && ((inited >> sb_slot) & 1u)) {
                        const AccRec<MSG> a = u.acc[sb_slot * K + k];
                        u0 = a.u0; u1 = a.u1; pen = a.pen;
                    }
                    const uint4 nr = u.nrec[e];
                    if (!((nr.x >> k) & 1u)) u1 = 0;
                    if (!((nr.y >> k) & 1u)) u0 = 0;
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
                const bool own = (c.cls[s] == k);
                if (draw_gap) {
                    tcand -= tol_gap<MSG>(xo, gsc * (xold ? gf_on : gf_off), k + 1);
                    if (own || !(tcand > tma)) tcand = -NXB_INF;
                }
                const double tp = (ip >= ip0) ? u.ptm[ip] : -NXB_INF;
                const bool haso = (io >= olo) && (u.bedge(io) == (unsigned)e);
                const double to = haso ? u.btm[io] : -NXB_INF;
                const double tf = dmax2(to, tcand);
                const double tnext = dmax2(dmax2(tp, tf), tma);
                // chunking.py:67-79: own class forces ON; ON pays the rate into class k
                pen += c.qm[s * K + k] * rate * (tcur - tnext);
                if (own) u0 = 0;
                tcur = tnext;
                if (tp == -NXB_INF && tf == -NXB_INF) {
                    // top of the edge: fold into the parent's accumulator (one normalisation)
                    const int ai = sa_slot * K + k;
                    AccRec<MSG> n; n.u0 = u0; n.u1 = u1; n.pen = pen;
                    if ((inited >> sa_slot) & 1u) { const AccRec<MSG> a = u.acc[ai]; n.u0 *= a.u0; n.u1 *= a.u1; n.pen += a.pen; }
                    MSG m2 = n.u0 > n.u1 ? n.u0 : n.u1;
                    if (!(m2 > (MSG)0)) { TOL_BAD(bad, 1); m2 = 1; }
                    const MSG r2 = M::pow2_rescale(m2);
                    n.u0 *= r2; n.u1 *= r2;
                    nxb_acc_pad(n);
                    u.acc[ai] = n;
                    inited |= 1u << sa_slot;
                    if (sb_slot >= 0) inited &= ~(1u << sb_slot);
                    --e;
                    fresh = true;
                } else if (tf > tp) {
                    bool live = true;
                    if (to > tcand) {
                        // the old trajectory changes state here: new piece, new candidate rate
                        // (a retained event on an edge without events -- only an initial trajectory
                        // from outside can hold one -- leaves the candidate slot empty)
                        xold ^= 1u; --io;
                        tcand = gsc > 0.f ? tnext : -NXB_INF;
                    } else {
                        // (a candidate that ties with a retained event is dropped: probability ~2^-50)
                        live = (tcand != to);
                    }
                    // next arrival (rootward): drawn by the next step, from this point on
                    gap_due = gsc > 0.f;
                    if (live) {
                        // (messages are scale-free: with OFF impossible the factor is a pure rescaling
                        // and is dropped, so that a long own-class segment cannot underflow ON to 0)
                        if (pen != 0.0) { if (u0 != (MSG)0) u1 *= M::expneg(pen); pen = 0.0; }
                        MSG mx = u0 > u1 ? u0 : u1;
                        if (!(mx > (MSG)0)) { TOL_BAD(bad, 2); mx = 1; }
                        const MSG r = M::pow2_rescale(mx);
                        const MSG l1 = u1 * r, l0 = u0 * r;        // scale-free message below the event
                        if (cnt < TCAP) {
                            // Ratio by fast division (operands in [0, 2)) -- but a hard constraint must stay
                            // exact: x / x is not guaranteed to be 1 there, and a q of 1 - 2^-24 in an own-class
                            // segment (l0 == 0) let the downward pass switch the class off once in ~10^8
                            // unit-sweeps.  l1 == 0 gives an exact 0.
                            LiveEv<MSG> ev;
                            ev.t = tnext;
                            const MSG q = M::div(l1, l0 + l1);
                            ev.q = (l0 == (MSG)0) ? (MSG)1 : (q < (MSG)1 ? q : (MSG)1);
                            // pad = 1: first record of this edge, i.e. the LAST live event the downward pass
                            // meets on it (it then closes the edge in the same step)
                            ev.e = (unsigned short)e; ev.pad = (unsigned short)(((evm >> (e & 63)) & 1ull) ? 0u : 1u);
                            nxb_ev_pad(ev);
                            nxb_ev_store(cev + cnt, ev);
                        } else bad = 2;
                        ++cnt;
                        evm |= 1ull << (e & 63);
                        // uniformized 2x2 matrix of this context (poisson.py:92-96)
                        if (own) { u0 = (MSG)0.5 * l0 + (MSG)0.5 * l1; u1 = l1; }
                        else {
                            u0 = l0 + a_on * (l1 - l0);
                            u1 = l1 + a_off * (l0 - l1);
                        }
                    }
                } else {
                    s = (int)((u.pcode[ip] >> 16) & 0xffu);     // rootward: the state before the event
                    --ip;
                    if ((c.cls[s] == k) != own) { tcand = tnext; gap_due = gsc > 0.f; }   // the own flag changes: new piece
                }
            }
        }
    }
    L.cnt = cnt;
    L.evm = evm;
    // ---- root draw (prior: toymodel.py:17-27 get_blink_distn) ----
    unsigned x = 0u;
    if (trk) {
        const AccRec<MSG> a = u.acc[c.slot[0] * K + k];
        double w1 = (double)a.u1 * (a.u0 != (MSG)0 ? fmax(exp(-a.pen), 2.2250738585072014e-308) : 1.0) * ml.p_on;
        double w0 = (double)a.u0 * ml.p_off;
        if (!(((unsigned)u.uhdr[8] >> k) & 1u)) w1 = 0.0;
        if (!(((unsigned)u.uhdr[9] >> k) & 1u)) w0 = 0.0;
        if (c.cls[u.node_pri[0]] == k) w0 = 0.0;
        if (!(w0 + w1 > 0.0)) TOL_BAD(bad, 3);
        x = (nxb_u32(nxb_xo_next(xo)) * (w0 + w1) < w1) ? 1u : 0u;
    }
    if (trk && x) nxb_atomic_or_shared(&u.newtol[0], 1u << k);
    L.xo = xo;
    L.bad |= bad;
    return x;
}

// ---------------------------------------------------------------------------------------------
// B2: downward sampling, edges in ascending index: the live events are read back in reverse;
// survivors (real transitions) are written from the top of the region downwards.  The OFF-time
// and transition statistics are fused in (summary.py:123-131,234-243).
// ---------------------------------------------------------------------------------------------
template <typename MSG, bool SPLIT>
NXB_DEV void tol_down_walk(const TolCtx& c, TolLane<MSG>& L, const TolUnit<MSG, SPLIT>& u, unsigned x) {
    typedef TolMath<MSG> M;
    const NxbSweepParams& p = *c.p;
    const NxbModelLayout& ml = p.ml;
    const NxbSummaryLayout& sl = p.sl;
    const int E = c.E, P = c.P, K = c.K, k = L.k;
    const bool trk = L.trk, accum_u = L.accum;
    const SPtr<const NxbEdgeRec> erec = {(unsigned)ml.off_edgerec};
    LiveEv<MSG>* const cev = L.cev;
    NxbXo xo = L.xo;
    const int cnt = (L.bad & 2) ? 0 : L.cnt;
    int bad = 0;
    int wtop = cnt - 1;            // survivors overwrite consumed records, from the last one down
    {
        const MSG a_on_f = M::a_on(ml), a_off1_f = M::a_off1(ml);
        int e = trk ? 0 : E;
        int ir = cnt - 1;              // read cursor (descending)
        bool fresh = true;
        int s = 0, ip = 0, ipe = 0, vb = 0, n_on = 0, n_off = 0;
        double tmb = 0.0, tcur = 0.0, offtime = 0.0;
        LiveEv<MSG> pf;
        pf.t = NXB_INF; pf.q = 0; pf.e = 0xffffu; pf.pad = 0;
        nxb_ev_pad(pf);
        if (ir >= 0) pf = nxb_ev_load(cev + ir);
        while (TOL_ANY(e < E)) {
            if (e < E) {
                if (fresh) {
                    const NxbEdgeRec er = erec[e];
                    vb = e + 1;
                    const int va = er.va;
                    x = (u.newtol[va] >> k) & 1u;
                    s = u.node_pri[va];
                    ip = u.poff[e]; ipe = u.poff[e + 1];
                    tcur = er.tma; tmb = er.tmb;
                    offtime = 0.0; n_on = 0; n_off = 0;
                    fresh = false;
                }
                const double tp = (ip < ipe) ? u.ptm[ip] : NXB_INF;
                const double tc = ((unsigned)pf.e == (unsigned)e) ? pf.t : NXB_INF;
                const double tnext = dmin2(dmin2(tp, tc), tmb);
                if (!x) offtime += tnext - tcur;
                tcur = tnext;
                if (tc < tp) {
                    const MSG l1 = pf.q, l0 = (MSG)1 - pf.q;
                    --ir;
                    // (the record at ir is in registers before a survivor may overwrite it: wtop >= ir)
                    if (ir >= 0) pf = nxb_ev_load(cev + ir);
                    else { pf.t = NXB_INF; pf.e = 0xffffu; }
                    const bool own = (c.cls[s] == k);
                    const MSG p1 = own ? (x ? (MSG)1 : (MSG)0.5) : (x ? a_off1_f : a_on_f);
                    const MSG w1 = p1 * l1, w0 = ((MSG)1 - p1) * l0;
                    if (!(w0 + w1 > (MSG)0)) TOL_BAD(bad, 4);
                    const MSG uu = M::uniform(xo);
                    const unsigned xn = (uu * (w0 + w1) < w1) ? 1u : 0u;
                    if (xn != x) {
                        LiveEv<MSG> sv;
                        sv.t = tc; sv.q = 0; sv.e = (unsigned short)(e | (xn ? CE_NEW1 : 0u)); sv.pad = 0;
                        nxb_ev_pad(sv);
                        nxb_ev_store(cev + wtop, sv);
                        --wtop;
                        if (xn) ++n_on; else ++n_off;
                        x = xn;
                    }
                } else {
                    // primary event or the bottom of the edge: close the OFF-time segment
                    if (accum_u && offtime > 0.0)
                        nxb_atomic_add_f64(p.dwell + (sl.d_off + (e * P + s) * K + k), offtime);
                    offtime = 0.0;
                    if (tp != NXB_INF) {
                        s = (int)((u.pcode[ip] >> 24) & 0xffu);
                        ++ip;
                    } else {
                        // packed word per edge (low: off->on, high: on->off), updated by halves with
                        // native 32-bit shared atomics; a CTA books far fewer than 2^32 per launch
                        if (accum_u && n_on) nxb_atomic_add_shared((unsigned*)&c.s_off_on[e], (unsigned)n_on);
                        if (accum_u && n_off) nxb_atomic_add_shared((unsigned*)&c.s_off_on[e] + 1, (unsigned)n_off);
                        if (x) nxb_atomic_or_shared(&u.newtol[vb], 1u << k);
                        ++e;
                        fresh = true;
                    }
                }
            }
        }
    }
    if (trk) {
        u.uhdr[10 + k] = (cnt - 1 - wtop) | (cnt << 16);   // survivors of this track | live events
        if (L.bad | bad) nxb_atomic_or_shared((unsigned*)&u.uhdr[4], (unsigned)(L.bad | bad));
    }
    L.xo = xo;
}

// ---------------------------------------------------------------------------------------------
// B2, event-driven form (trees with E <= 64; the walk above serves larger ones).  On an edge
// without a live event the track keeps the state of the upper node, so the lane only visits the
// edges on which it has live events (a third of the boundaries of the full walk at the p53 shape):
// the state at the top of such an edge is the state below the nearest ancestor edge with events --
// the highest set bit of evm & ancmask[va], already sampled because edges are taken in ascending
// order -- or the root's.  What the walk did on the edges that are skipped here is done once per
// unit, bit-parallel over the tracks, by tol_fill_* (node states below event-free edges, OFF dwell
// of the tracks that are OFF throughout such an edge).  Same draws in the same order as the walk:
// the sampled trajectory is identical.
// ---------------------------------------------------------------------------------------------
template <typename MSG, bool SPLIT>
NXB_DEV void tol_down_ev(const TolCtx& c, TolLane<MSG>& L, const TolUnit<MSG, SPLIT>& u) {
    typedef TolMath<MSG> M;
    const NxbSweepParams& p = *c.p;
    const NxbModelLayout& ml = p.ml;
    const NxbSummaryLayout& sl = p.sl;
    const int P = c.P, K = c.K, k = L.k;
    const bool trk = L.trk, accum_u = L.accum;
    const SPtr<const NxbEdgeRec> erec = {(unsigned)ml.off_edgerec};
    const SPtr<const unsigned long long> ancm = {(unsigned)ml.off_ancmask};
    LiveEv<MSG>* const cev = L.cev;
    NxbXo xo = L.xo;
    const int cnt = (L.bad & 2) ? 0 : L.cnt;
    const unsigned long long evm = (L.bad & 2) ? 0ull : L.evm;
    int bad = 0;
    int wtop = cnt - 1;            // survivors overwrite consumed records, from the last one down
    {
        const MSG a_on_f = M::a_on(ml), a_off1_f = M::a_off1(ml);
        int ir = cnt - 1;              // read cursor (descending)
        bool fresh = true;
        bool act = trk && ir >= 0;
        int e = 0, s = 0, ip = 0, ipe = 0, n_on = 0, n_off = 0;
        unsigned x = 0u;
        double tmb = 0.0, tcur = 0.0, offtime = 0.0;
        LiveEv<MSG> pf;
        pf.t = NXB_INF; pf.q = 0; pf.e = 0xffffu; pf.pad = 0;
        nxb_ev_pad(pf);
        if (ir >= 0) pf = nxb_ev_load(cev + ir);
        while (TOL_ANY(act)) {
            if (act) {
                if (fresh) {
                    e = (int)pf.e;
                    const NxbEdgeRec er = erec[e];
                    const int va = er.va;
                    const unsigned long long m = evm & ancm[va];
                    const int a = 64 - nxb_clzll(m);       // node below the nearest ancestor edge with events; 0: the root
                    x = (u.newtol[a] >> k) & 1u;
                    s = u.node_pri[va];
                    ip = u.poff[e]; ipe = u.poff[e + 1];
                    tcur = er.tma; tmb = er.tmb;
                    offtime = 0.0; n_on = 0; n_off = 0;
                    fresh = false;
                }
                double tp = (ip < ipe) ? u.ptm[ip] : NXB_INF;
                const double tc = ((int)pf.e == e) ? pf.t : NXB_INF;
                const double tnext = dmin2(dmin2(tp, tc), tmb);
                if (!x) offtime += tnext - tcur;
                tcur = tnext;
                bool close = true;             // a primary event or the bottom of the edge ends an OFF-time segment
                if (tc < tp) {
                    const MSG l1 = pf.q, l0 = (MSG)1 - pf.q;
                    const bool last_here = pf.pad != 0;
                    --ir;
                    // (the record at ir is in registers before a survivor may overwrite it: wtop >= ir)
                    if (ir >= 0) pf = nxb_ev_load(cev + ir);
                    else { pf.t = NXB_INF; pf.e = 0xffffu; }
                    const bool own = (c.cls[s] == k);
                    const MSG p1 = own ? (x ? (MSG)1 : (MSG)0.5) : (x ? a_off1_f : a_on_f);
                    const MSG w1 = p1 * l1, w0 = ((MSG)1 - p1) * l0;
                    if (!(w0 + w1 > (MSG)0)) TOL_BAD(bad, 4);
                    const MSG uu = M::uniform(xo);
                    const unsigned xn = (uu * (w0 + w1) < w1) ? 1u : 0u;
                    if (xn != x) {
                        LiveEv<MSG> sv;
                        sv.t = tc; sv.q = 0; sv.e = (unsigned short)(e | (xn ? CE_NEW1 : 0u)); sv.pad = 0;
                        nxb_ev_pad(sv);
                        nxb_ev_store(cev + wtop, sv);
                        --wtop;
                        if (xn) ++n_on; else ++n_off;
                        x = xn;
                    }
                    // nothing else on this edge (no further live event, no primary event): its bottom
                    // is reached in the same step
                    close = last_here && ip >= ipe;
                    if (close && !x) offtime += tmb - tc;
                }
                if (close) {
                    if (accum_u && offtime > 0.0)
                        nxb_atomic_add_f64(p.dwell + (sl.d_off + (e * P + s) * K + k), offtime);
                    offtime = 0.0;
                    if (tp != NXB_INF) {
                        s = (int)((u.pcode[ip] >> 24) & 0xffu);
                        ++ip;
                    } else {
                        if (accum_u && n_on) nxb_atomic_add_shared((unsigned*)&c.s_off_on[e], (unsigned)n_on);
                        if (accum_u && n_off) nxb_atomic_add_shared((unsigned*)&c.s_off_on[e] + 1, (unsigned)n_off);
                        if (x) nxb_atomic_or_shared(&u.newtol[e + 1], 1u << k);
                        fresh = true;
                        act = ir >= 0;
                    }
                }
            }
        }
    }
    if (trk) {
        u.uhdr[10 + k] = (cnt - 1 - wtop) | (cnt << 16);   // survivors of this track | live events
        if (L.bad | bad) nxb_atomic_or_shared((unsigned*)&u.uhdr[4], (unsigned)(L.bad | bad));
        // the track's accumulator records are dead: its slot-0 record carries evm to tol_fill_*
        *reinterpret_cast<unsigned long long*>(&u.acc[k]) = evm;
    }
    L.xo = xo;
}

// (TOL_FORCE_WALK: test builds of the host simulation that keep the full walk, to compare with)
NXB_DEV bool tol_event_driven(int E) {
#if defined(TOL_FORCE_WALK)
    (void)E; return false;
#else
    return E <= 64;
#endif
}

template <typename MSG, bool SPLIT>
NXB_DEV void tol_down(const TolCtx& c, TolLane<MSG>& L, const TolUnit<MSG, SPLIT>& u, unsigned x) {
    if (tol_event_driven(c.E)) tol_down_ev<MSG, SPLIT>(c, L, u);
    else tol_down_walk<MSG, SPLIT>(c, L, u, x);
}

// ---------------------------------------------------------------------------------------------
// What tol_down_ev leaves to do, for ONE unit after all of its tracks are done (E <= 64): scalar
// form for the host simulation; blink_finish (sweep_kernel.cuh) does the same with a warp.
//   hasev[e]   = tracks with a live event on edge e (from the evm words left in the accumulators)
//   node_out[v] = sampled tolerance states of node v: its own bits where hasev[v-1] is set (or v
//                is the root), else those of the nearest ancestor that has them
//   OFF dwell of the tracks that are OFF throughout an event-free edge (summary.py:123-131)
// ---------------------------------------------------------------------------------------------
template <typename MSG, bool SPLIT>
NXB_DEV void tol_fill_scalar(const TolCtx& c, const TolUnit<MSG, SPLIT>& u, bool accum, unsigned* node_out) {
    const NxbSweepParams& p = *c.p;
    const int E = c.E, P = c.P, K = c.K, N = E + 1;
    const unsigned ALLK = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);
    const SPtr<const NxbEdgeRec> erec = {(unsigned)p.ml.off_edgerec};
    node_out[0] = u.newtol[0];
    for (int v = 1; v < N; ++v) {
        const int e = v - 1;
        unsigned h = 0u;
        for (int k = 0; k < K; ++k)
            h |= (unsigned)((*reinterpret_cast<const unsigned long long*>(&u.acc[k]) >> e) & 1ull) << k;
        const int va = erec[e].va;
        node_out[v] = (u.newtol[v] & h) | (node_out[va] & ~h & ALLK);
        const unsigned offm = ~node_out[v] & ~h & ALLK;
        if (!accum || !offm) continue;
        double t = erec[e].tma;
        int s = u.node_pri[va];
        for (int i = u.poff[e]; i <= u.poff[e + 1]; ++i) {
            const bool last = i == u.poff[e + 1];
            const double te = last ? erec[e].tmb : u.ptm[i];
            if (te - t > 0.0)
                for (int k = 0; k < K; ++k)
                    if ((offm >> k) & 1u) nxb_atomic_add_f64(p.dwell + (p.sl.d_off + (e * P + s) * K + k), te - t);
            if (!last) { s = (int)((u.pcode[i] >> 24) & 0xffu); t = te; }
        }
    }
}

}  // namespace nxb
