// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// CPU simulation of the per-lane code of the tolerance update (nxblink_b200/csrc/tol_update.cuh):
// the same source, compiled for the host with the shims of nxb_portable.h, run one lane (one
// track) after the other on one unit.  It lets the tests check the device algorithm -- the
// restatement of poisson.py:176-199, chunking.py:39-155,264-289 and the tolerance half of
// summary.py:190-243 -- against exact two-state posteriors without a GPU
// (tests/test_hostsim_tol.py).  The warp-level set-up and write-back of a unit (blink_setup /
// blink_finish in sweep_kernel.cuh) are restated sequentially here.
//
// Build: oracle/hostsim/Makefile (g++), output oracle/_ref/libtol_hostsim.so.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../nxblink_b200/csrc/tol_tables.h"
#include "../../nxblink_b200/csrc/tol_update.cuh"

namespace nxb {
alignas(16) unsigned char nxb_smem[1 << 22];
}
using namespace nxb;

struct hs_unit {
    int32_t n_pri, n_blk;
    const uint8_t* node_pri;        // [N]
    uint32_t* node_tol;             // [N] in/out
    const double* ptm;              // primary events sorted by (edge, time)
    const uint32_t* pcode;          // edge | sa << 16 | sb << 24
    double* btm;                    // blink events sorted by (track, edge, time), in/out (capacity cap_blk)
    uint32_t* bcode;                // edge | track << 16 | new << 31
    int32_t cap_blk;
    const uint32_t* tt;             // [N] data masks: class may be ON
    const uint32_t* tf;             // [N] class may be OFF
};

template <typename MSG>
static int run(const nxb_model_desc* d, hs_unit* un, uint64_t seed, uint32_t unit, uint32_t sweep,
               int accumulate, double* doff, uint64_t* off_on, uint64_t* on_off, int32_t* n_live_out) {
    const int N = d->n_nodes, P = d->n_states, K = d->n_classes, E = N - 1;
    NxbSweepParams p;
    memset(&p, 0, sizeof(p));
    NxbModelLayout& ml = p.ml;
    ml.N = N; ml.E = E; ml.P = P; ml.K = K; ml.nnz = d->nnz;
    std::vector<unsigned char> blob;
    std::string err;
    if (nxb_build_tol_tables(d, ml, blob, err)) { fprintf(stderr, "hostsim: %s\n", err.c_str()); return -1; }
    auto align_up = [](int x, int a) { return (x + a - 1) / a * a; };
    memcpy(nxb_smem, blob.data(), blob.size());
    ml.bytes = ml.blink_bytes; ml.off_stats = ml.blink_bytes;
    const int stats_bytes = 8 * (2 * E + P + 4 + K + 16);
    memset(nxb_smem + ml.off_stats, 0, stats_bytes);
    // one unit region, laid out as build_blink_layout does (nxb_api.cu)
    NxbWarpLayout& wl = p.wl;
    const int capP = un->n_pri + 8, capB = un->n_blk + 8;
    {
        int o = 0;
        auto take = [&](int bytes, int al) { o = align_up(o, al); int r = o; o += bytes; return r; };
        wl.capP = capP; wl.capB = capB;
        wl.off_ptm = take(8 * capP, 8);
        wl.off_btm = take(8 * capB, 8);
        wl.off_node_tol = take(4 * N, 4);
        wl.off_node_pri = take(align_up(N, 4), 4);
        wl.off_pcode = take(4 * capP, 4);
        wl.off_bcode = take(2 * capB, 2);
        wl.off_poff = take(4 * (E + 1), 4);
        wl.off_uhdr = take(4 * (10 + 32), 4);
        wl.off_toff = take(4 * (K + 1), 4);
        wl.off_acc = take((int)sizeof(AccRec<MSG>) * K * ml.nslots, 16);
        wl.off_newtol = take(4 * N, 4);
        wl.off_nrec = take(16 * E, 16);
        wl.bytes = align_up(o, 16);
    }
    const unsigned ub = (unsigned)align_up(ml.off_stats + stats_bytes, 16);
    if ((size_t)ub + wl.bytes > sizeof(nxb_smem)) return -2;
    // pool: generous
    const int TCAP = 4096;
    p.gcapC = TCAP; p.g_off_ctm = 0;
    p.gscratch_stride = (long long)sizeof(LiveEv<MSG>) * TCAP * K;
    std::vector<unsigned char> pool((size_t)p.gscratch_stride + 64);
    unsigned char* gb0 = (unsigned char*)(((uintptr_t)pool.data() + 15) & ~(uintptr_t)15);
    p.sl.d_off = 0;
    p.dwell = doff;
    // ---- set-up (blink_setup) ----
    double* ptm = sm_ptr<double>(ub + wl.off_ptm);
    unsigned* pcode = sm_ptr<unsigned>(ub + wl.off_pcode);
    double* btm = sm_ptr<double>(ub + wl.off_btm);
    unsigned short* bedge = sm_ptr<unsigned short>(ub + wl.off_bcode);
    unsigned char* node_pri = sm_ptr<unsigned char>(ub + wl.off_node_pri);
    int* poff = sm_ptr<int>(ub + wl.off_poff);
    int* toff = sm_ptr<int>(ub + wl.off_toff);
    int* uhdr = sm_ptr<int>(ub + wl.off_uhdr);
    unsigned* newtol = sm_ptr<unsigned>(ub + wl.off_newtol);
    uint4* nrec = sm_ptr<uint4>(ub + wl.off_nrec);
    for (int i = 0; i < un->n_pri; ++i) { ptm[i] = un->ptm[i]; pcode[i] = un->pcode[i]; }
    for (int v = 0; v < N; ++v) { node_pri[v] = un->node_pri[v]; newtol[v] = 0u; }
    for (int e = 0; e <= E; ++e) poff[e] = 0;
    for (int i = 0; i < un->n_pri; ++i) poff[(pcode[i] & 0xffffu) + 1]++;
    for (int e = 0; e < E; ++e) poff[e + 1] += poff[e];
    for (int k = 0; k <= K; ++k) toff[k] = 0;
    for (int i = 0; i < un->n_blk; ++i) {
        btm[i] = un->btm[i];
        bedge[i] = (unsigned short)(un->bcode[i] & 0xffffu);
        toff[((un->bcode[i] >> 16) & 0xffu) + 1]++;
    }
    for (int k = 0; k < K; ++k) toff[k + 1] += toff[k];
    for (int e = 0; e < E; ++e) {
        nrec[e].x = un->tt[e + 1]; nrec[e].y = un->tf[e + 1]; nrec[e].z = un->node_tol[e + 1];
        nrec[e].w = (unsigned)node_pri[e + 1] | ((unsigned)poff[e] << 8);
    }
    memset(uhdr, 0, 4 * (10 + 32));
    uhdr[0] = 1; uhdr[2] = un->n_blk; uhdr[3] = (int)unit; uhdr[4] = 0; uhdr[5] = (int)sweep; uhdr[6] = accumulate ? 1 : 0;
    uhdr[8] = (int)un->tt[0]; uhdr[9] = (int)un->tf[0];
    // ---- lane jobs ----
    TolCtx tc;
    tc.p = &p; tc.E = E; tc.P = P; tc.K = K;
    tc.cls.off = (unsigned)ml.off_cls; tc.qm.off = (unsigned)ml.off_qm; tc.slot.off = (unsigned)ml.off_slot;
    tc.s_off_on.off = (unsigned)ml.off_stats;
    tc.k0 = (unsigned)(seed & 0xffffffffull); tc.k1 = (unsigned)(seed >> 32);
    int total_live = 0;
    for (int k = 0; k < K; ++k) {
        TolLane<MSG> L;
        TolUnit<MSG, true> u;
        tol_begin<MSG, true>(tc, L, u, k, K, ub, gb0);
        const unsigned x_root = tol_up<MSG, true>(tc, L, u);
        tol_down<MSG, true>(tc, L, u, x_root);
        total_live += L.cnt;
    }
    if (n_live_out) *n_live_out = total_live;
    // ---- write-back (blink_finish) ----
    if (uhdr[4]) return uhdr[4];
    int dst = 0;
    for (int k = 0; k < K; ++k) {
        const int packed = uhdr[10 + k], n_mine = packed & 0xffff, n_live = packed >> 16;
        const LiveEv<MSG>* cev = (const LiveEv<MSG>*)(gb0 + p.g_off_ctm) + (size_t)k * TCAP;
        for (int q = 0; q < n_mine; ++q) {
            const LiveEv<MSG> sv = cev[n_live - 1 - q];
            if (dst >= un->cap_blk) return 3;
            un->btm[dst] = sv.t;
            un->bcode[dst] = ((unsigned)sv.e & CE_EDGE) | ((unsigned)k << 16) | ((sv.e & CE_NEW1) ? (1u << 31) : 0u);
            ++dst;
        }
    }
    un->n_blk = dst;
    if (tol_event_driven(E)) {
        // the event-driven downward pass leaves the event-free edges to the per-unit fill-in
        TolUnit<MSG, true> uu;
        uu.bind(ub, wl);
        std::vector<unsigned> filled(N, 0u);
        tol_fill_scalar<MSG, true>(tc, uu, accumulate != 0, filled.data());
        for (int v = 0; v < N; ++v) un->node_tol[v] = filled[v];
    } else
        for (int v = 0; v < N; ++v) un->node_tol[v] = newtol[v];
    if (accumulate) {
        const unsigned long long* st = (const unsigned long long*)(nxb_smem + ml.off_stats);
        for (int e = 0; e < E; ++e) { off_on[e] += st[e] & 0xffffffffull; on_off[e] += st[e] >> 32; }
    }
    return 0;
}

extern "C" int hostsim_tol_update(const nxb_model_desc* d, hs_unit* un, uint64_t seed, uint32_t unit,
                                  uint32_t sweep, int msg_f64, int accumulate, double* doff,
                                  uint64_t* off_on, uint64_t* on_off, int32_t* n_live) {
    return msg_f64 ? run<double>(d, un, seed, unit, sweep, accumulate, doff, off_on, on_off, n_live)
                   : run<float>(d, un, seed, unit, sweep, accumulate, doff, off_on, on_off, n_live);
}
