// Host-side construction of the model tables of the tolerance update (plain C++, no CUDA): the
// prefix of the model blob that blink_kernel stages into shared memory.  Shared by nxb_api.cu (the
// product) and oracle/hostsim (the CPU simulation of the per-lane code, test infrastructure).
#pragma once
#include <math.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/nxblink_b200.h"
#include "nxb_types.h"

// Fills the tolerance-update fields of `ml` (nslots, a_on/a_off, acc_blk, p_on/p_off and the off_*
// of the tables below) and appends to `blob`:
//   NxbEdgeRec[E], slot int[N], cls u8[P], qm f64[P*K], ancmask u64[N]
// ml.N/E/P/K must be set.  Returns 0, or 1 with a message.
static inline int nxb_build_tol_tables(const nxb_model_desc* d, NxbModelLayout& ml,
                                       std::vector<unsigned char>& blob, std::string& err, int pcdf_cap = 0) {
    const int N = d->n_nodes, P = d->n_states, K = d->n_classes, E = N - 1;
    auto align_up = [](int x, int a) { return (x + a - 1) / a * a; };
    auto put = [&](const void* src, int bytes, int al) {
        int o = align_up((int)blob.size(), al);
        blob.resize(o + bytes);
        memcpy(blob.data() + o, src, bytes);
        return o;
    };
    std::vector<double> tm(N, 0.0);
    for (int v = 1; v < N; ++v) tm[v] = tm[d->parent[v]] + d->blen[v];
    // accumulator slots of the upward tree pass (edges in descending index)
    std::vector<int> slot(N, -1);
    {
        std::vector<int> free_slots;
        int next = 0;
        for (int e = E - 1; e >= 0; --e) {
            int vb = e + 1, va = d->parent[vb];
            if (slot[va] < 0) {
                if (!free_slots.empty()) {
                    auto it = std::min_element(free_slots.begin(), free_slots.end());
                    slot[va] = *it; free_slots.erase(it);
                } else slot[va] = next++;
            }
            if (slot[vb] >= 0) free_slots.push_back(slot[vb]);
        }
        ml.nslots = std::max(next, 1);
        if (ml.nslots > 32) {
            err = "tree needs more than 32 live accumulators; renumber nodes depth-first";
            return 1;
        }
    }
    const double M = std::max(d->rate_on, d->rate_off);
    ml.rate_on = d->rate_on; ml.rate_off = d->rate_off;
    ml.a_on = d->rate_on / (2.0 * M);
    ml.a_off = d->rate_off / (2.0 * M);
    ml.acc_blk[0] = (2.0 * M - d->rate_on) / (2.0 * M);   // other class, track OFF
    ml.acc_blk[1] = (2.0 * M - d->rate_off) / (2.0 * M);  // other class, track ON
    // Own-class context (the primary state belongs to this track's class): the reference draws
    // virtual events there too (poisson.py:176-199, rate 2 on - 0), but they are inert -- the track
    // is ON on both sides of such a point (chunking.py:69-73 forces it), the uniformized row of ON
    // is (0, 1) there, so the event multiplies the message by 1 and its backward draw returns ON
    // with probability 1.  Not generating them leaves the transition kernel of the sweep unchanged
    // and saves 1/K of the live events (5 % at K = 20; a third at the toy models' K = 3).
    ml.acc_blk[2] = 0.0;                                  // own class, OFF (infeasible state)
    ml.acc_blk[3] = 0.0;                                  // own class, ON: inert events, dropped
    ml.gf_off = (float)(1.0 / ml.acc_blk[0]); ml.gf_on = (float)(1.0 / ml.acc_blk[1]);
    ml.a_on_f = (float)ml.a_on; ml.a_off_f = (float)ml.a_off; ml.a_off1_f = (float)(1.0 - ml.a_off); ml.pad_f = 0.f;
    ml.p_on = d->rate_on / (d->rate_on + d->rate_off);
    ml.p_off = d->rate_off / (d->rate_on + d->rate_off);

    std::vector<NxbEdgeRec> recs(E);
    for (int e = 0; e < E; ++e) {
        const int va = d->parent[e + 1];
        const double len = d->blen[e + 1], rate = d->rate[e + 1];
        memset(&recs[e], 0, sizeof(NxbEdgeRec));
        recs[e].tma = tm[va]; recs[e].tmb = tm[va] + len; recs[e].rate = rate;
        // ln2 / (dominating rate 2 * max(on, off) * rate_e); 0: no events possible on this edge
        recs[e].gsc = (rate > 0.0 && len > 0.0) ? (float)(0.6931471805599453 / (2.0 * M * rate)) : 0.f;
        recs[e].va = (unsigned short)va;
        recs[e].sa_slot = (signed char)slot[va];
        recs[e].sb_slot = (signed char)slot[e + 1];
    }
    ml.off_edgerec = put(recs.data(), (int)(sizeof(NxbEdgeRec) * E), 16);
    ml.off_slot = put(slot.data(), 4 * N, 4);
    std::vector<unsigned char> cls(P);
    for (int s = 0; s < P; ++s) cls[s] = (unsigned char)d->state_class[s];
    ml.off_cls = put(cls.data(), P, 1);
    std::vector<double> qm((size_t)P * K, 0.0);
    for (int s = 0; s < P; ++s)
        for (int i = d->q_ptr[s]; i < d->q_ptr[s + 1]; ++i)
            qm[(size_t)s * K + d->state_class[d->q_col[i]]] += d->q_val[i];
    ml.off_qm = put(qm.data(), 8 * P * K, 8);
    // Edges on the path from the root to every node, as a bit mask (event-driven downward pass of
    // tol_update.cuh: the nearest ancestor edge on which a track has a live event is the highest
    // set bit of (edges with events) & ancmask, because parents are numbered before children).
    {
        std::vector<unsigned long long> anc(N, 0ull);
        for (int v = 1; v < N; ++v)
            anc[v] = anc[d->parent[v]] | ((v - 1 < 64) ? (1ull << (v - 1)) : 0ull);
        ml.off_ancmask = put(anc.data(), 8 * N, 8);
    }
    // ---- candidate generator of the event-parallel update (tol_par.cuh) ----
    {
        std::vector<NxbGenRec> g(E);
        double cum = 0.0;
        for (int e = 0; e < E; ++e) {
            const double len = d->blen[e + 1], rate = d->rate[e + 1];
            const bool active = rate > 0.0 && len > 0.0;
            const double om = 2.0 * M * rate;            // dominating rate per unit time
            g[e].cum = cum;
            g[e].inv_om = active ? 1.0 / om : 0.0;
            g[e].tma = recs[e].tma; g[e].tmb = recs[e].tmb;
            if (active) cum += om * len;
        }
        ml.tp_lam = cum;
        ml.tp_W = (double)K * cum / (double)NXB_TP_NSLICE;
        ml.off_grec = put(g.data(), (int)(sizeof(NxbGenRec) * E), 16);
        // Poisson(W) thresholds: count = first j with word < pcdf[j]; the last entry catches everything
        std::vector<unsigned> pcdf;
        {
            const double W = ml.tp_W;
            if (W > 2000.0) { err = "tolerance update: more than 2000 expected candidates per slice of the event axis"; return 1; }
            long double pr = expl(-(long double)W), cdf = pr;
            int j = 0;
            while (true) {
                const long double x = cdf * 4294967296.0L;
                if (x >= 4294967295.0L || j > 8000) { pcdf.push_back(0xffffffffu); break; }
                pcdf.push_back((unsigned)x);
                ++j;
                pr *= (long double)W / (long double)j;
                cdf += pr;
            }
        }
        ml.n_pcdf = (int)pcdf.size();
        // The table has a fixed capacity so that the blob keeps its size when nxb_update_model
        // changes the rates: room for about twice the rates the handle was created with.
        if (pcdf_cap <= 0) pcdf_cap = ((int)(2.0 * ml.tp_W + 12.0 * sqrt(2.0 * ml.tp_W + 1.0)) + 64 + 7) / 8 * 8;
        if (ml.n_pcdf > pcdf_cap) { err = "the blink rates grew beyond what the handle was created for; create a new handle"; return 1; }
        pcdf.resize(pcdf_cap, 0xffffffffu);
        ml.pcdf_cap = pcdf_cap;
        ml.off_pcdf = put(pcdf.data(), 4 * pcdf_cap, 4);
        // first guess of the edge: the last edge that starts before the upper end of the bin
        {
            std::vector<unsigned short> etab(NXB_TP_NBIN, 0);
            const double bw = cum / (double)NXB_TP_NBIN;
            for (int b = 0; b < NXB_TP_NBIN; ++b) {
                int e = E - 1;
                while (e > 0 && g[e].cum > (double)(b + 1) * bw) --e;
                etab[b] = (unsigned short)e;
            }
            ml.off_etab = put(etab.data(), 2 * NXB_TP_NBIN, 2);
            ml.tp_inv_lam = cum > 0.0 ? 1.0 / cum : 0.0;
            ml.tp_inv_bin = cum > 0.0 ? (double)NXB_TP_NBIN / cum : 0.0;
        }
        for (int i = 0; i < 4; ++i)
            ml.tp_thr[i] = (unsigned)fmin(ml.acc_blk[i] * 4294967296.0, 4294967295.0);
    }
    // ---- level tables of the level-parallel tree passes (tol_par.cuh) ----
    {
        std::vector<int> height(N, 0), depth(N, 0), nchild(N, 0);
        for (int v = N - 1; v >= 1; --v) height[d->parent[v]] = std::max(height[d->parent[v]], height[v] + 1);
        for (int v = 1; v < N; ++v) { depth[v] = depth[d->parent[v]] + 1; nchild[d->parent[v]]++; }
        const int H = height[0], D = *std::max_element(depth.begin(), depth.end());
        std::vector<unsigned short> lvl_off(H + 1, 0), lvl_node, ch_off(N + 1, 0), ch_node(E, 0), dep_off(D + 1, 0), dep_edge;
        for (int h = 1; h <= H; ++h) {
            lvl_off[h - 1] = (unsigned short)lvl_node.size();
            for (int v = 0; v < N; ++v) if (height[v] == h) lvl_node.push_back((unsigned short)v);
        }
        lvl_off[H] = (unsigned short)lvl_node.size();
        for (int v = 0; v < N; ++v) ch_off[v + 1] = (unsigned short)(ch_off[v] + nchild[v]);
        {
            std::vector<int> fill(N, 0);
            for (int v = 1; v < N; ++v) { const int pa = d->parent[v]; ch_node[ch_off[pa] + fill[pa]++] = (unsigned short)v; }
        }
        for (int dd = 1; dd <= D; ++dd) {
            dep_off[dd - 1] = (unsigned short)dep_edge.size();
            for (int v = 1; v < N; ++v) if (depth[v] == dd) dep_edge.push_back((unsigned short)(v - 1));
        }
        dep_off[D] = (unsigned short)dep_edge.size();
        // message slots in level order: taken at the node's level, released after the parent's level
        std::vector<short> lslot(N, -1);
        {
            std::vector<int> free_slots;
            int next = 0;
            for (int h = 1; h <= H; ++h) {
                for (int i = lvl_off[h - 1]; i < lvl_off[h]; ++i) {
                    const int v = lvl_node[i];
                    if (!free_slots.empty()) { lslot[v] = (short)free_slots.back(); free_slots.pop_back(); }
                    else lslot[v] = (short)next++;
                }
                // the children of this level's nodes are not needed again
                for (int i = lvl_off[h - 1]; i < lvl_off[h]; ++i) {
                    const int v = lvl_node[i];
                    for (int j = ch_off[v]; j < ch_off[v + 1]; ++j)
                        if (lslot[ch_node[j]] >= 0) free_slots.push_back(lslot[ch_node[j]]);
                }
            }
            ml.nslots_lvl = std::max(next, 1);
        }
        ml.n_levels = H; ml.n_depths = D;
        ml.off_lvl_off = put(lvl_off.data(), 2 * (H + 1), 2);
        ml.off_lvl_node = put(lvl_node.data(), 2 * (int)std::max<size_t>(lvl_node.size(), 1), 2);
        ml.off_ch_off = put(ch_off.data(), 2 * (N + 1), 2);
        ml.off_ch_node = put(ch_node.data(), 2 * std::max(E, 1), 2);
        ml.off_lslot = put(lslot.data(), 2 * N, 2);
        ml.off_dep_off = put(dep_off.data(), 2 * (D + 1), 2);
        ml.off_dep_edge = put(dep_edge.data(), 2 * std::max(E, 1), 2);
    }
    blob.resize(align_up((int)blob.size(), 16));
    ml.blink_bytes = (int)blob.size();
    return 0;
}
