// C-ABI of the sweep engine (include/nxblink_b200.h) and the persistent kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/nxblink_b200.h"
#include "sweep_kernel.cuh"
#include "tol_par.cuh"
#include "tol_tables.h"

#define NXB_MAX_THREADS 640
#define NXB_LEAN_THREADS 896       /* LEAN primary-only kernel with f32 messages: 28 warps (7 per SM sub-partition) at 72 registers, no spills */
#define NXB_TP_THREADS 640         /* tolpar_kernel: up to 20 warps (10 units x 2 warps at the p53 shape) */
#define NXB_BLINK_GROUP_WARPS 5   /* independent warp groups of blink_kernel: 160 lanes = 8 units x 20 tracks at the p53 shape */
#define NXB_BLINK_THREADS 800      /* 25 warps at 72 registers: 40 units x 20 tracks per CTA at the p53 shape */

namespace nxb {

// ---------------------------------------------------------------------------
// persistent sweep kernel: one CTA per SM, one warp per unit (primary update; fused form: both updates)
// ---------------------------------------------------------------------------
// model blob -> shared memory, statistic partials zeroed, model-level fields of the context
__device__ __forceinline__ void cta_prologue(Ctx& c, const NxbSweepParams& p, int tid, int lane) {
    unsigned char* const smem = nxb_smem;
    const NxbModelLayout& ml = p.ml;
    for (int i = tid * 16; i < ml.bytes; i += blockDim.x * 16)
        *(uint4*)(smem + i) = *(const uint4*)(p.blob + i);
    unsigned long long* stats = (unsigned long long*)(smem + ml.off_stats);
    for (int i = tid; i < ml.stats_words; i += blockDim.x) stats[i] = 0ull;
    __syncthreads();
    c.p = &p;
    c.N = ml.N; c.E = ml.E; c.P = ml.P; c.K = ml.K; c.lane = lane;
    c.parent.off = ml.off_parent;
    c.eperm.off = ml.off_eperm;
    c.slot.off = ml.off_slot;
    c.tma.off = ml.off_tma;
    c.len.off = ml.off_len;
    c.rate.off = ml.off_rate;
    c.lam_pri.off = ml.off_lam_pri;
    c.p0_pri.off = ml.off_p0_pri;
    c.lam_blk.off = ml.off_lam_blk;
    c.p0_blk.off = ml.off_p0_blk;
    c.gsc_blk.off = ml.off_gsc_blk;
    c.qval.off = ml.off_qval;
    c.qinfo.off = ml.off_qinfo;
    c.qptr.off = ml.off_qptr;
    c.cls.off = ml.off_cls;
    c.pi.off = ml.off_pi;
    c.qm.off = ml.off_qm;
    c.clsmask.off = ml.off_clsmask;
    c.ell_info = ml.off_ell_info;
    c.ell_q = ml.off_ell_q;
    c.s_off_on.off = ml.off_stats;
    c.s_on_off.off = ml.off_stats + 8 * ml.E;
    c.s_root_pri.off = ml.off_stats + 8 * (2 * ml.E);
    c.s_root_misc.off = ml.off_stats + 8 * (2 * ml.E + ml.P);
    c.s_root_on_cls.off = ml.off_stats + 8 * (2 * ml.E + ml.P + 4);
    c.k0 = (unsigned)(p.seed & 0xffffffffull);
    c.k1 = (unsigned)(p.seed >> 32);
    c.t_last = clock64();
}

// flush the CTA's statistic partials
__device__ __forceinline__ void cta_epilogue(Ctx& c, const NxbSweepParams& p, int tid) {
    const NxbModelLayout& ml = p.ml;
    const NxbSummaryLayout& sl = p.sl;
    const unsigned long long* stats = (const unsigned long long*)(nxb_smem + ml.off_stats);
    __syncthreads();
    // per edge one packed word: low 32 bits off->on transitions, high 32 bits on->off
    for (int i = tid; i < ml.E; i += blockDim.x) {
        const unsigned long long w = stats[i];
        if (w & 0xffffffffull) atomicAdd(p.counts + sl.c_off_on + i, w & 0xffffffffull);
        if (w >> 32) atomicAdd(p.counts + sl.c_on_off + i, w >> 32);
    }
    for (int i = tid; i < ml.P; i += blockDim.x)
        if (c.s_root_pri[i]) atomicAdd(p.counts + sl.c_root_pri + i, c.s_root_pri[i]);
    for (int i = tid; i < ml.K; i += blockDim.x)
        if (c.s_root_on_cls[i]) atomicAdd(p.counts + sl.c_root_on_cls + i, c.s_root_on_cls[i]);
    if (tid == 0) {
        if (c.s_root_misc[0]) atomicAdd(p.counts + sl.c_root_on, c.s_root_misc[0]);
        if (c.s_root_misc[1]) atomicAdd(p.counts + sl.c_root_off, c.s_root_misc[1]);
        if (c.s_root_misc[2]) atomicAdd(p.counts + sl.c_nsamples, c.s_root_misc[2]);
    }
}

// L2 prefetch of the live part of one unit's slab, 16 lines by the first 16 lanes of a warp:
// header + node states (3 lines), primary events (2), blink times (7), blink codes (4).  The
// persistent kernels issue it one full wave of work ahead, so that whichever CTA takes that unit
// later finds its slab in L2 instead of HBM (the event counts are not known yet: expected sizes).
__device__ __forceinline__ void prefetch_unit_slab(const NxbSweepParams& p, long long u, int lane) {
    if (lane >= 16 || u >= p.n_units) return;
    const unsigned char* slab = p.state + u * p.state_stride;
    const unsigned char* a;
    if (lane < 3) a = slab + 128 * lane;
    else if (lane == 3) a = slab + p.so_ptm;
    else if (lane == 4) a = slab + p.so_pcode;
    else if (lane < 12) a = slab + p.so_btm + 128 * (lane - 5);
    else a = slab + p.so_bcode + 128 * (lane - 12);
    asm volatile("prefetch.global.L2 [%0];" :: "l"(a));
}

// PRIMARY_ONLY: the kernel runs the primary update of one sweep and leaves the tolerance update
// to blink_kernel (the two alternate, one launch each per sweep); otherwise both are fused and a
// launch runs every unit up to p.target_sweeps.
// LEAN (primary-only launches of the asynchronous split path): no trajectory validation in the
// kernel.  The primary-only kernels never run the initial pass (p_init == 0 in every such launch),
// so `init` is a compile-time false there; both cut the code the hot kernel carries (11.7 k -> 8.5 k
// SASS instructions, 96 -> 88 registers).
template <typename MSG, bool PRIMARY_ONLY, bool LEAN>
__global__ void __launch_bounds__((LEAN && sizeof(MSG) == 4) ? NXB_LEAN_THREADS : NXB_MAX_THREADS, 1)
sweep_kernel(const __grid_constant__ NxbSweepParams p) {
    const bool p_init = !PRIMARY_ONLY && p.init;
    const bool p_validate = !LEAN && p.validate;
    unsigned char* const smem = nxb_smem;
    const NxbModelLayout& ml = p.ml;
    const NxbWarpLayout& wl = p.wl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // list length from the device (launch enqueued before the list was written): nothing to do?
    int n_list = p.n_list;
    if (p.n_list_dev) {
        n_list = *(volatile const int*)p.n_list_dev;
        if (n_list <= 0) return;
        if (p.defer_total && blockIdx.x == 0 && tid == 0) atomicAdd(p.defer_total, (unsigned long long)n_list);
    }
    Ctx c;
    cta_prologue(c, p, tid, lane);
    c.wbase = (unsigned)(p.smem_warp0 + warp * wl.bytes);
    c.gbase = p.gscratch + ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * p.gscratch_stride;
    c.node_tol.off = c.wbase + wl.off_node_tol;
    c.node_pri.off = c.wbase + wl.off_node_pri;
    c.ptm.off = c.wbase + wl.off_ptm;
    c.pcode.off = c.wbase + wl.off_pcode;
    c.btm.off = c.wbase + wl.off_btm;
    c.bcode.off = c.wbase + wl.off_bcode;
    c.boff.off = c.wbase + wl.off_boff;
    c.poff.off = c.wbase + wl.off_poff;
    c.tmpa.off = c.wbase + wl.off_tmpa;
    c.tmpb.off = c.wbase + wl.off_tmpb;
    c.tmpc.off = c.wbase + wl.off_tmpc;
    unsigned long long* d_pri = sm_ptr<unsigned long long>(c.wbase + wl.off_dpri);
    unsigned* d_tt = sm_ptr<unsigned>(c.wbase + wl.off_dtt);
    unsigned* d_tf = sm_ptr<unsigned>(c.wbase + wl.off_dtf);
    c.pri_ok.off = c.wbase + wl.off_dpri; c.tt_ok.off = c.wbase + wl.off_dtt; c.tf_ok.off = c.wbase + wl.off_dtf;
    const int N = ml.N;
    const long long n_work = p.unit_list ? (long long)n_list : p.n_units;
    // Lockstep groups: GW consecutive warps take a batch of GW units and run every sub-phase
    // together (named barrier per group).  GW == 1: independent warps (deferral tiers).
    const int GW = p.group_warps > 1 ? p.group_warps : 1;
    const int group = warp / GW, wing = warp - group * GW;
    const bool lockstep = GW > 1;
    c.bar_id = 1 + group;
    c.wing = wing;
    c.gwbase = (unsigned)(p.smem_warp0 + (group * GW) * wl.bytes);
    c.ggbase = p.gscratch + ((long long)blockIdx.x * (blockDim.x >> 5) + group * GW) * p.gscratch_stride;
    c.bar_threads = lockstep ? GW * 32 : 32;
    c.nbar = 0;
    unsigned long long* s_batch = (unsigned long long*)(smem + ml.off_stats) + ml.stats_words;
    const unsigned n_steps = p_init ? 1u : (p.target_sweeps - p.begin_sweeps);
    const unsigned n_iter = PRIMARY_ONLY ? (p.validate_only ? 0u : 1u) : 2u * n_steps;   // phases per unit and launch

    while (true) {
        unsigned long long widx = 0;
        if (lockstep) {
            c.sync_mask = 0xffffffffu; c.nbar = 0;
            if (wing == 0 && lane == 0) s_batch[group] = atomicAdd(p.work_counter, 1ull);
            phase_sync(c);
            widx = s_batch[group] * GW + wing;
            phase_sync(c);
            if ((long long)(s_batch[group] * GW) >= n_work) break;
        } else {
            if (lane == 0) widx = atomicAdd(p.work_counter, 1ull);
            widx = __shfl_sync(FULL, widx, 0);
            if ((long long)widx >= n_work) break;
        }
        bool has = (long long)widx < n_work;
        if (p.prefetch_wave && !p.unit_list)
            prefetch_unit_slab(p, (long long)widx + (long long)gridDim.x * (blockDim.x >> 5), lane);
        long long u = 0;
        unsigned char* slab = nullptr;
        NxbUnitHeader hdr;
        hdr.sweeps_done = 0; hdr.n_pri = 0; hdr.n_blk = 0; hdr.phase = NXB_PHASE_PRIMARY; hdr.pad = 0;
        int rc = RC_OK;
        if (has) {
            u = p.unit_list ? (long long)p.unit_list[widx] : (long long)widx;
            c.unit = u;
            c.unit32 = (unsigned)(p.unit_offset + u);
            slab = p.state + u * p.state_stride;
            hdr = *(const NxbUnitHeader*)slab;
            if (p_init) { hdr.sweeps_done = 0; hdr.n_pri = 0; hdr.n_blk = 0; hdr.phase = NXB_PHASE_INIT; hdr.pad = 0; }
            const bool off_schedule = (lockstep || PRIMARY_ONLY) && !p_init &&
                (hdr.phase != NXB_PHASE_PRIMARY || hdr.sweeps_done != p.begin_sweeps);
            if (PRIMARY_ONLY && p.unit_list && !p_init && hdr.phase == NXB_PHASE_BLINK && hdr.sweeps_done == p.begin_sweeps) {
                has = false;        // (deferred by the tolerance kernel of the tier before: that kernel's turn)
            } else if ((int)hdr.n_pri > wl.capP || (int)hdr.n_blk > wl.capB || off_schedule) {
                // does not fit this tier's shared-memory arrays: untouched, next tier
                if (lane == 0) p.defer_list[atomicAdd(p.defer_count, 1)] = (int)u;
                has = false;
            }
        }
        // ---- load the unit (coalesced; only the live part of the slab is read) ----
        if (has) {
            const long long site = u / p.chains;
            const unsigned long long* g_pri = p.pri_ok + site * N;
            const unsigned* g_tt = p.tol_true_ok + site * N;
            const unsigned* g_tf = p.tol_false_ok + site * N;
            if (LEAN) { for (int v = lane; v < N; v += 32) d_pri[v] = g_pri[v]; }      // (the lean layout has no tolerance data)
            else for (int v = lane; v < N; v += 32) { d_pri[v] = g_pri[v]; d_tt[v] = g_tt[v]; d_tf[v] = g_tf[v]; }
            c.nP = hdr.n_pri; c.nB = hdr.n_blk;
            if (hdr.phase != NXB_PHASE_INIT) {
                const unsigned* g_tol = (const unsigned*)(slab + p.so_tol);
                const unsigned* g_npri = (const unsigned*)(slab + p.so_pri);
                // streaming (evict-first) accesses: the slab is touched once per launch and must
                // not displace the L2-resident scratch of the resident warps
                for (int v = lane; v < N; v += 32) c.node_tol[v] = __ldcs(g_tol + v);
                for (int v = lane; v < (N + 3) / 4; v += 32) sm_ptr<unsigned>(c.node_pri.off)[v] = __ldcs(g_npri + v);
                const double* g_ptm = (const double*)(slab + p.so_ptm);
                const unsigned* g_pcode = (const unsigned*)(slab + p.so_pcode);
                for (int i = lane; i < c.nP; i += 32) { c.ptm[i] = __ldcs(g_ptm + i); c.pcode[i] = __ldcs(g_pcode + i); }
                const double* g_btm = (const double*)(slab + p.so_btm);
                const unsigned* g_bcode = (const unsigned*)(slab + p.so_bcode);
                for (int i = lane; i < c.nB; i += 32) { c.btm[i] = __ldcs(g_btm + i); c.bcode[i] = __ldcs(g_bcode + i); }
            }
            __syncwarp();
        }
        // split mode: the tolerance update ran in another kernel; check its result on arrival
        if (PRIMARY_ONLY && p_validate && has && hdr.phase != NXB_PHASE_INIT) rc = validate_unit(c);
        tick(c, T_LOAD);
        bool spilled = false;
        // phases from the slab state on whose statistics were booked by an earlier, failed launch
        // (the unit outgrew its slab: NXB_ERR_STATE_CAP), and phases completed by this launch
        const unsigned booked_before = has ? hdr.pad : 0u;
        unsigned done_phases = 0u;
        bool statecap = false;
        // One phase per iteration, so that every phase function has a single call site.  In
        // lockstep mode every warp of the group runs the same number of iterations and arrives
        // at every barrier, whether or not its unit is still active.
        unsigned step = 0;
        while (true) {
            bool active = has && rc == RC_OK && !spilled;
            const bool more = hdr.phase == NXB_PHASE_INIT || hdr.sweeps_done < p.target_sweeps;
            if (lockstep || PRIMARY_ONLY) { if (step >= n_iter) break; }
            else if (!(active && more)) break;
            active = active && more;
            const bool init = !PRIMARY_ONLY && hdr.phase == NXB_PHASE_INIT;
            const bool blink_turn = PRIMARY_ONLY ? false
                : (lockstep ? ((step & 1u) != 0u && !p_init) : (hdr.phase == NXB_PHASE_BLINK));
            ++step;
            if (lockstep && p_init && (step & 1u) == 0u) continue;     // init: one primary pass only
            if ((lockstep || PRIMARY_ONLY) && active && blink_turn != (hdr.phase == NXB_PHASE_BLINK)) active = false;
            c.sweep = init ? NXB_SWEEP_INIT : hdr.sweeps_done;
            const bool accum = !init && hdr.sweeps_done >= p.accum_from && done_phases >= booked_before;
            c.nbar = 0;
            // the barriers of the tolerance update are not optional: its lanes work on other warps' units
            c.sync_mask = blink_turn ? 0xffffffffu : p.sync_mask;
            if (!blink_turn) {
                if (active && init) rc = init_blink(c);
                if (active && rc == RC_OK) {
                    rc = primary_phase<MSG>(c, init, accum);
                    if (rc == RC_OK) { hdr.phase = init ? NXB_PHASE_PRIMARY : NXB_PHASE_BLINK; if (!init) ++done_phases; }
                }
                phase_drain(c, NXB_NBAR_PRIMARY);
            } else if (!PRIMARY_ONLY) {
                // every warp of a lockstep group takes part: its lanes carry tracks of other units
                if (active || lockstep) {
                    const int brc = blink_phase<MSG>(c, accum, active);
                    if (active) {
                        rc = brc;
                        if (rc == RC_STATECAP) { statecap = true; ++done_phases; rc = RC_ERROR; }
                        else if (rc == RC_DEFER && c.nB < 0) {     // sweep complete, blink list spilled to HBM
                            hdr.phase = NXB_PHASE_PRIMARY; hdr.sweeps_done++; ++done_phases;
                            spilled = true;
                        } else if (rc == RC_OK) { hdr.phase = NXB_PHASE_PRIMARY; hdr.sweeps_done++; ++done_phases; }
                    }
                }
                phase_drain(c, NXB_NBAR_BLINK);
            }
            if (p_validate && active && rc == RC_OK) rc = validate_unit(c);
        }
        // ---- store the unit ----
        if (has && rc != RC_ERROR && !(PRIMARY_ONLY && p.validate_only)) {
            if (c.nP > p.gcapP || (!spilled && c.nB > p.gcapB)) {
                fail(c, NXB_ERR_STATE_CAP, c.nP > p.gcapP ? c.nP : c.nB);
                statecap = true;
            } else {
                __syncwarp();
                // (a unit that is handed to the next tier in the middle of its already booked phases keeps the rest)
                hdr.pad = booked_before > done_phases ? booked_before - done_phases : 0u;
                hdr.n_pri = (uint16_t)c.nP;
                hdr.n_blk = (uint16_t)(spilled ? -c.nB : c.nB);
                if (lane == 0) *(NxbUnitHeader*)slab = hdr;
                if (hdr.phase != NXB_PHASE_INIT) {
                    unsigned* g_tol = (unsigned*)(slab + p.so_tol);
                    unsigned* g_npri = (unsigned*)(slab + p.so_pri);
                    // (the primary update leaves the tolerance states and blink events as they are)
                    if (!PRIMARY_ONLY) for (int v = lane; v < N; v += 32) __stcs(g_tol + v, c.node_tol[v]);
                    for (int v = lane; v < (N + 3) / 4; v += 32) __stcs(g_npri + v, sm_ptr<unsigned>(c.node_pri.off)[v]);
                    double* g_ptm = (double*)(slab + p.so_ptm);
                    unsigned* g_pcode = (unsigned*)(slab + p.so_pcode);
                    for (int i = lane; i < c.nP; i += 32) { __stcs(g_ptm + i, c.ptm[i]); __stcs(g_pcode + i, c.pcode[i]); }
                    if (!spilled && !PRIMARY_ONLY) {
                        double* g_btm = (double*)(slab + p.so_btm);
                        unsigned* g_bcode = (unsigned*)(slab + p.so_bcode);
                        for (int i = lane; i < c.nB; i += 32) { __stcs(g_btm + i, c.btm[i]); __stcs(g_bcode + i, c.bcode[i]); }
                    }
                }
                if (rc == RC_DEFER && lane == 0) p.defer_list[atomicAdd(p.defer_count, 1)] = (int)u;
                __syncwarp();
            }
        }
        // the slab keeps the state the unit was loaded with; record how many phases from there on
        // are booked, so that the rerun after the slabs have grown does not book them again
        if (statecap && has && lane == 0 && !p_init)
            ((NxbUnitHeader*)slab)->pad = booked_before > done_phases ? booked_before : done_phases;
        tick(c, T_STORE);
    }
    cta_epilogue(c, p, tid);
}

// ---------------------------------------------------------------------------
// TOLERANCE UPDATE as its own kernel (split mode).  A CTA takes batches of U = p.batch_units
// units (U * K <= 32 * warps) into shared memory: the warps set the units up (blink_setup),
// every lane of the CTA then carries one (unit, track) job (blink_job), and the warps write the
// units back (blink_finish).  Units are in phase BLINK at sweep p.begin_sweeps when it starts
// (the primary-only sweep_kernel ran just before) and in phase PRIMARY of the next sweep after.
// Units that are elsewhere are skipped (the primary launch listed them for the next tier);
// units whose events do not fit the shared-memory capacities are listed here.
// ---------------------------------------------------------------------------
// (25 warps are 7 on one SM sub-partition: 7 x 32 x 72 registers is what its 16 K file holds)
template <typename MSG, int GWT>
__global__ void __launch_bounds__(NXB_BLINK_THREADS, 1)
blink_kernel(const __grid_constant__ NxbSweepParams p) {
    const NxbModelLayout& ml = p.ml;
    const NxbWarpLayout& wl = p.wl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    Ctx c;
    cta_prologue(c, p, tid, lane);
    c.bar_id = 1; c.bar_threads = 32; c.nbar = 0; c.sync_mask = 0u;     // phase_sync: no-op
    c.gwbase = (unsigned)p.smem_warp0;
    const int N = ml.N, K = ml.K;
    // Independent warp groups (p.blink_groups > 1): G groups of W / G warps, each with its own
    // batches of p.batch_units / G units, its own part of the unit regions and of the event pool,
    // and a named barrier instead of __syncthreads: a group waits for the slowest of ITS walks
    // only, and its load / set-up / write-back overlap the walks of the other groups.
    // (GWT: warps per group at compile time, 0 = the whole CTA is one group)
    const int GW = GWT ? GWT : W, G = GWT ? W / GWT : 1, U = GWT ? p.batch_units / G : p.batch_units;
    const int grp = warp / GW, wing = warp - grp * GW, gtid = tid - grp * GW * 32;
    const unsigned ub0 = (unsigned)(p.smem_warp0 + grp * U * wl.bytes);
    unsigned char* const gb0 = p.gscratch + ((long long)blockIdx.x * p.batch_units + (long long)grp * U) * p.gscratch_stride;
    c.ggbase = gb0;
    c.wing = wing;
    auto group_sync = [&]() {
        if (GWT == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" :: "r"(1 + (int)(threadIdx.x >> 5) / GWT), "n"(GWT * 32) : "memory");
    };
    const long long n_work = p.unit_list ? (long long)p.n_list : p.n_units;
    unsigned long long* s_batch = (unsigned long long*)(nxb_smem + ml.off_stats) + ml.stats_words;
    auto bind = [&](int ui) {
        c.wbase = ub0 + (unsigned)(ui * wl.bytes);
        c.gbase = gb0 + (long long)ui * p.gscratch_stride;
        c.node_tol.off = c.wbase + wl.off_node_tol;
        c.node_pri.off = c.wbase + wl.off_node_pri;
        c.ptm.off = c.wbase + wl.off_ptm;
        c.pcode.off = c.wbase + wl.off_pcode;
        c.btm.off = c.wbase + wl.off_btm;
        c.bcode.off = c.wbase + wl.off_bcode;
        c.poff.off = c.wbase + wl.off_poff;
        c.tmpa.off = c.wbase + wl.off_tmpa;
    };
    while (true) {
        if (gtid == 0) s_batch[grp] = atomicAdd(p.work_counter, 1ull);
        group_sync();
        const long long w0 = (long long)s_batch[grp] * U;
        group_sync();
        if (w0 >= n_work) break;
        if (p.prefetch_wave && !p.unit_list)
            for (int ui = wing; ui < U; ui += GW) prefetch_unit_slab(p, w0 + (long long)gridDim.x * p.batch_units + ui, lane);
        // ---- load + set up the units of the batch ----
        for (int ui = wing; ui < U; ui += GW) {
            bind(ui);
            const long long widx = w0 + ui;
            bool active = false, accum = false;
            SPtr<int> uhdr = {c.wbase + wl.off_uhdr};
            const unsigned *g_tt = nullptr, *g_tf = nullptr, *g_bcode = nullptr;
            c.nP = 0; c.nB = 0; c.unit = 0; c.unit32 = 0u; c.sweep = 0u;
            if (widx < n_work) {
                const long long u = p.unit_list ? (long long)p.unit_list[widx] : widx;
                const unsigned char* slab = p.state + u * p.state_stride;
                const NxbUnitHeader hdr = *(const NxbUnitHeader*)slab;
                if (hdr.phase == NXB_PHASE_BLINK && hdr.sweeps_done == p.begin_sweeps) {
                    if ((int)hdr.n_pri > wl.capP || (int)hdr.n_blk > wl.capB) {
                        if (lane == 0) p.defer_list[atomicAdd(p.defer_count, 1)] = (int)u;
                    } else {
                        active = true;
                        c.unit = u; c.unit32 = (unsigned)(p.unit_offset + u);
                        c.nP = hdr.n_pri; c.nB = hdr.n_blk; c.sweep = hdr.sweeps_done;
                        accum = hdr.sweeps_done >= p.accum_from && hdr.pad == 0u;     // (pad: already booked, see sweep_kernel)
                        const long long site = u / p.chains;
                        g_tt = p.tol_true_ok + site * N;
                        g_tf = p.tol_false_ok + site * N;
                        g_bcode = (const unsigned*)(slab + p.so_bcode);
                        const unsigned* g_tol = (const unsigned*)(slab + p.so_tol);
                        const unsigned* g_npri = (const unsigned*)(slab + p.so_pri);
                        for (int v = lane; v < N; v += 32) c.node_tol[v] = __ldcs(g_tol + v);
                        for (int v = lane; v < (N + 3) / 4; v += 32) sm_ptr<unsigned>(c.node_pri.off)[v] = __ldcs(g_npri + v);
                        const double* g_ptm = (const double*)(slab + p.so_ptm);
                        const unsigned* g_pcode = (const unsigned*)(slab + p.so_pcode);
                        for (int i = lane; i < c.nP; i += 32) { c.ptm[i] = __ldcs(g_ptm + i); c.pcode[i] = __ldcs(g_pcode + i); }
                        const double* g_btm = (const double*)(slab + p.so_btm);
                        for (int i = lane; i < c.nB; i += 32) c.btm[i] = __ldcs(g_btm + i);
                        __syncwarp();
                    }
                }
            }
            tick(c, T_LOAD);
            blink_setup<true>(c, accum, active, g_tt, g_tf, g_bcode);
            if (lane == 0) uhdr[7] = (int)c.unit;
            tick(c, T_B0);
        }
        group_sync();
        // ---- one (unit, track) job per lane ----
        blink_job<MSG, true>(c, gtid, U * K, ub0, gb0);
        group_sync();
        tick(c, T_BW);          // (waiting for the slowest lane of the group is booked here)
        // ---- write the units back ----
        for (int ui = wing; ui < U; ui += GW) {
            bind(ui);
            SPtr<int> uhdr = {c.wbase + wl.off_uhdr};
            if (!uhdr[0]) continue;
            c.nB = uhdr[2]; c.unit32 = (unsigned)uhdr[3]; c.sweep = (unsigned)uhdr[5];
            c.unit = (long long)uhdr[7];
            const bool accum = uhdr[6] != 0;
            unsigned char* slab = p.state + c.unit * p.state_stride;
            NxbUnitHeader hdr = *(const NxbUnitHeader*)slab;     // (issued early: needed after blink_finish)
            int rc = blink_finish<MSG, true>(c, accum);          // new blink list written to the slab
            tick(c, T_BW);
            if (rc == RC_STATECAP) { if (lane == 0) ((NxbUnitHeader*)slab)->pad = 1u; continue; }
            if (rc == RC_ERROR) continue;
            hdr.phase = NXB_PHASE_PRIMARY; hdr.sweeps_done++; hdr.pad = hdr.pad > 0u ? hdr.pad - 1u : 0u;
            hdr.n_blk = (uint16_t)c.nB;
            __syncwarp();
            if (lane == 0) *(NxbUnitHeader*)slab = hdr;
            unsigned* g_tol = (unsigned*)(slab + p.so_tol);
            for (int v = lane; v < N; v += 32) __stcs(g_tol + v, c.node_tol[v]);
            __syncwarp();
            tick(c, T_STORE);
        }
        // the bulk copies of this warp's units (blink_finish) have left shared memory and are complete
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        group_sync();
    }
    cta_epilogue(c, p, tid);
}


// ---------------------------------------------------------------------------
// TOLERANCE UPDATE, event-parallel form (tol_par.cuh): a group of WPU warps owns one unit at a
// time; the unit's trajectory and its whole event pool live in shared memory.  Units are in phase
// BLINK at sweep p.begin_sweeps (the primary-only sweep_kernel ran just before) and in phase
// PRIMARY of the next sweep after.  Units that are elsewhere are skipped; units whose events do not
// fit the shared-memory capacities of this tier are listed for the next one, untouched.
// ---------------------------------------------------------------------------
template <typename MSG, int WPU>
__global__ void __launch_bounds__(NXB_TP_THREADS, 1)
tolpar_kernel(const __grid_constant__ NxbSweepParams p) {
    const NxbModelLayout& ml = p.ml;
    const NxbTolLayout& tl = p.tl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ctx c;
    cta_prologue(c, p, tid, lane);
    TpCtx tc;
    tc.p = &p; tc.E = ml.E; tc.P = ml.P; tc.K = ml.K;
    tc.cls.off = ml.off_cls; tc.qm.off = ml.off_qm; tc.slot.off = ml.off_slot;
    tc.erec.off = ml.off_edgerec; tc.grec.off = ml.off_grec; tc.etab.off = ml.off_etab; tc.pcdf.off = ml.off_pcdf;
    tc.lvl_off.off = ml.off_lvl_off; tc.lvl_node.off = ml.off_lvl_node; tc.ch_off.off = ml.off_ch_off; tc.ch_node.off = ml.off_ch_node;
    tc.dep_off.off = ml.off_dep_off; tc.dep_edge.off = ml.off_dep_edge; tc.lslot.off = ml.off_lslot;
    tc.s_off_on = c.s_off_on; tc.s_root_misc = c.s_root_misc; tc.s_root_on_cls = c.s_root_on_cls;
    tc.k0 = c.k0; tc.k1 = c.k1;
    TpGroup<WPU> g;
    g.lane = lane; g.w = warp % WPU; g.li = g.w * 32 + lane;
    const int grp = warp / WPU;
    g.bar = 1 + grp;
    constexpr int G = 32 * WPU;
    TpUnit<MSG> u;
    u.bind((unsigned)(p.smem_warp0 + grp * tl.bytes), tl);
    TpTimer tm;
    tm.perf = p.perf ? p.perf + 12 : nullptr; tm.on = (g.li == 0); tm.t_last = clock64();
    const int N = ml.N, E = ml.E, K = ml.K;
    const long long n_work = p.unit_list ? (long long)p.n_list : p.n_units;
    while (true) {
        long long widx = 0;
        if (WPU == 1) {
            if (lane == 0) widx = (long long)atomicAdd(p.work_counter, 1ull);
            widx = __shfl_sync(FULL, widx, 0);
        } else {
            if (g.li == 0) *(long long*)&u.x[6] = (long long)atomicAdd(p.work_counter, 1ull);
            g.sync();
            widx = *(const long long*)&u.x[6];
            g.sync();
        }
        if (widx >= n_work) break;
        const long long unit = p.unit_list ? (long long)p.unit_list[widx] : widx;
        unsigned char* const slab = p.state + unit * p.state_stride;
        const NxbUnitHeader hdr = *(const NxbUnitHeader*)slab;
        if (!(hdr.phase == NXB_PHASE_BLINK && hdr.sweeps_done == p.begin_sweeps)) continue;
        if ((int)hdr.n_pri > tl.capP || (int)hdr.n_blk > tl.capB) {
            if (g.li == 0) p.defer_list[atomicAdd(p.defer_count, 1)] = (int)unit;
            continue;
        }
        // ---- load (coalesced, streaming; only the live part of the slab is read) ----
        u.nP = hdr.n_pri; u.nB = hdr.n_blk;
        u.unit32 = (unsigned)(p.unit_offset + unit); u.sweep = hdr.sweeps_done;
        const bool accum = hdr.sweeps_done >= p.accum_from && hdr.pad == 0u;
        const long long site = unit / p.chains;
        const unsigned* const g_tt = p.tol_true_ok + site * N;
        const unsigned* const g_tf = p.tol_false_ok + site * N;
        {
            const unsigned* g_tol = (const unsigned*)(slab + p.so_tol);
            const unsigned* g_npri = (const unsigned*)(slab + p.so_pri);
            for (int v = g.li; v < N; v += G) u.node_tol[v] = __ldcs(g_tol + v);
            for (int v = g.li; v < (N + 3) / 4; v += G) sm_ptr<unsigned>(u.node_pri.off)[v] = __ldcs(g_npri + v);
            const double* g_ptm = (const double*)(slab + p.so_ptm);
            const unsigned* g_pcode = (const unsigned*)(slab + p.so_pcode);
            for (int i = g.li; i < u.nP; i += G) { u.ptm[i] = __ldcs(g_ptm + i); u.pcode[i] = __ldcs(g_pcode + i); }
            const double* g_btm = (const double*)(slab + p.so_btm);
            const unsigned* g_bcode = (const unsigned*)(slab + p.so_bcode);
            for (int i = g.li; i < u.nB; i += G) { u.btm[i] = __ldcs(g_btm + i); u.bcode[i] = __ldcs(g_bcode + i); }
            for (int e = g.li; e < E; e += G) { u.f0[e] = 0u; u.f1[e] = 0u; u.hasev[e] = 0u; }
            if (g.li == 0) { u.x[4] = 0; u.x[8] = (int)__ldg(g_tt); u.x[9] = (int)__ldg(g_tf); }
        }
        g.sync();
        // per-edge record: data masks of the lower node, primary states at both ends, range of the
        // edge's primary events (sorted by (edge, time): two lower bounds)
        for (int e = g.li; e < E; e += G) {
            int lo = 0, hi = u.nP;
            while (lo < hi) { const int m = (lo + hi) >> 1; if ((int)(u.pcode[m] & 0xffffu) < e) lo = m + 1; else hi = m; }
            const int a = lo;
            hi = u.nP;
            while (lo < hi) { const int m = (lo + hi) >> 1; if ((int)(u.pcode[m] & 0xffffu) <= e) lo = m + 1; else hi = m; }
            const int va = tc.erec[e].va;
            u.nrec[e] = make_uint4(__ldg(g_tt + e + 1), __ldg(g_tf + e + 1),
                                   (unsigned)u.node_pri[e + 1] | ((unsigned)u.node_pri[va] << 8),
                                   (unsigned)a | ((unsigned)lo << 16));
        }
        g.sync();
        tm.tick(TPT_LOAD);
        int n_new = 0, detail = 0;
        const int rc = tolpar_update<MSG, WPU>(tc, u, g, accum, slab, &n_new, &detail, tm);
        if (rc == TP_DEFER) {
            if (g.li == 0) p.defer_list[atomicAdd(p.defer_count, 1)] = (int)unit;
            g.sync();
            continue;
        }
        if (rc == TP_ERROR) {
            if (g.li == 0) {
                const int code = detail < 0 ? NXB_ERR_STATE_CAP : ((detail & 2) ? NXB_ERR_SCRATCH_CAP : NXB_ERR_INFEASIBLE);
                if (atomicCAS(p.status, 0, code) == 0) {
                    p.status[1] = (int)unit; p.status[2] = detail < 0 ? -detail : -3; p.status[3] = (int)hdr.sweeps_done;
                }
            }
            g.sync();
            continue;
        }
        // ---- store: header, node states (the new blink list is already in the slab) ----
        {
            NxbUnitHeader h2 = hdr;
            h2.phase = NXB_PHASE_PRIMARY; h2.sweeps_done = hdr.sweeps_done + 1; h2.n_blk = (uint16_t)n_new; h2.pad = hdr.pad > 0u ? hdr.pad - 1u : 0u;
            if (g.li == 0) *(NxbUnitHeader*)slab = h2;
            unsigned* g_tol = (unsigned*)(slab + p.so_tol);
            for (int v = g.li; v < N; v += G) __stcs(g_tol + v, u.newtol[v]);
            if (accum && g.li == 0) {
                const int on = __popc(u.newtol[0]);
                atomicAdd(&c.s_root_misc[0], (unsigned long long)on);
                atomicAdd(&c.s_root_misc[1], (unsigned long long)(K - on));
                atomicAdd(&c.s_root_on_cls[c.cls[u.node_pri[0]]], (unsigned long long)on);
                atomicAdd(&c.s_root_misc[2], 1ull);
            }
        }
        g.sync();
        tm.tick(TPT_STORE);
    }
    cta_epilogue(c, p, tid);
}

// ---------------------------------------------------------------------------
// independent (non-fused) statistics of the stored trajectories: one warp per unit,
// lanes over edges, straight from the HBM slabs.  Restates summary.py:190-243 directly
// (segment by segment, as navigation.py:14-46 gen_segments) and is used to cross-check
// the statistics fused into the sweep.
// ---------------------------------------------------------------------------
__global__ void summarize_kernel(const __grid_constant__ NxbSweepParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const NxbModelLayout& ml = p.ml;
    for (int i = threadIdx.x * 16; i < ml.bytes; i += blockDim.x * 16)
        *(uint4*)(smem + i) = *(const uint4*)(p.blob + i);
    __syncthreads();
    const int* parent = (const int*)(smem + ml.off_parent);
    const double* tma = (const double*)(smem + ml.off_tma);
    const double* len = (const double*)(smem + ml.off_len);
    const unsigned short* qinfo = (const unsigned short*)(smem + ml.off_qinfo);
    const unsigned short* qptr = (const unsigned short*)(smem + ml.off_qptr);
    const unsigned char* cls = (const unsigned char*)(smem + ml.off_cls);
    const NxbSummaryLayout& sl = p.sl;
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int E = ml.E, P = ml.P, K = ml.K;
    for (long long u = wid; u < p.n_units; u += nw) {
        const unsigned char* slab = p.state + u * p.state_stride;
        const NxbUnitHeader hdr = *(const NxbUnitHeader*)slab;
        const unsigned* g_tol = (const unsigned*)(slab + p.so_tol);
        const unsigned char* g_npri = (const unsigned char*)(slab + p.so_pri);
        const double* g_ptm = (const double*)(slab + p.so_ptm);
        const unsigned* g_pcode = (const unsigned*)(slab + p.so_pcode);
        const double* g_btm = (const double*)(slab + p.so_btm);
        const unsigned* g_bcode = (const unsigned*)(slab + p.so_bcode);
        if (lane == 0) {
            int s0 = g_npri[0];
            int on = __popc(g_tol[0]);
            atomicAdd(p.counts + sl.c_root_pri + s0, 1ull);
            atomicAdd(p.counts + sl.c_root_on, (unsigned long long)on);
            atomicAdd(p.counts + sl.c_root_off, (unsigned long long)(K - on));
            atomicAdd(p.counts + sl.c_root_on_cls + cls[s0], (unsigned long long)on);
            atomicAdd(p.counts + sl.c_nsamples, 1ull);
        }
        for (int e = lane; e < E; e += 32) {
            int s = g_npri[parent[e]];
            unsigned mask = g_tol[parent[e]];
            double t = tma[e];
            const double tend = tma[e] + len[e];
            int last_blk = -1;      // blink events already processed at time t (events of DIFFERENT
                                    // tracks may share a time: gap times sit on the f32 grid)
            while (true) {
                // next event on this edge after t, ties between blink events in index order
                double best = NXB_INF; int which = -1; bool is_pri = false;
                for (int i = 0; i < hdr.n_pri; ++i)
                    if ((int)(g_pcode[i] & 0xffffu) == e && g_ptm[i] > t && g_ptm[i] < best) {
                        best = g_ptm[i]; which = i; is_pri = true;
                    }
                for (int i = 0; i < hdr.n_blk; ++i)
                    if ((int)(g_bcode[i] & 0xffffu) == e &&
                        (g_btm[i] > t || (g_btm[i] == t && i > last_blk && last_blk >= 0)) && g_btm[i] < best) {
                        best = g_btm[i]; which = i; is_pri = false;
                    }
                double tn = (which < 0) ? tend : best;
                double dt = tn - t;
                atomicAdd(p.dwell + sl.d_T + e * P + s, dt);
                for (int k = 0; k < K; ++k)
                    if (!((mask >> k) & 1u))
                        atomicAdd(p.dwell + sl.d_off + ((long long)e * P + s) * K + k, dt);
                if (which < 0) break;
                if (is_pri) {
                    int sb = (g_pcode[which] >> 24) & 0xffu;
                    int idx = qptr[s];
                    while ((qinfo[idx] & 0xffu) != (unsigned)sb) ++idx;
                    atomicAdd(p.counts + sl.c_pri_trans + (long long)e * ml.nnz + idx, 1ull);
                    s = sb;
                } else {
                    unsigned code = g_bcode[which];
                    if (code >> 31) atomicAdd(p.counts + sl.c_off_on + e, 1ull);
                    else atomicAdd(p.counts + sl.c_on_off + e, 1ull);
                    mask ^= 1u << ((code >> 16) & 0xffu);
                }
                last_blk = is_pri ? -1 : which;
                t = tn;
            }
        }
    }
}


// ---------------------------------------------------------------------------
// Forward (prior) simulation of the blink process, root to leaves: the Gillespie loop of
// nxblink/fwdsample.py:67-262, one thread per unit, written straight into the unit slabs so that
// the sampled trajectories can be summarised, exported or used as chain states.
// Semantics: the primary state's own tolerance class cannot switch off (the reference's
// comment at fwdsample.py:57-59; its guard at :60-64 tests the wrong track and lets it through,
// SURVEY.md component 13).  Root: primary ~ prior, own class ON, other classes ~ blink prior.
// ---------------------------------------------------------------------------
__global__ void forward_kernel(const __grid_constant__ NxbSweepParams p, int cap_events,
                               unsigned short* __restrict__ tmp_code, double* __restrict__ tmp_time,
                               const int* __restrict__ root_pri, const unsigned* __restrict__ root_tol) {
    extern __shared__ __align__(16) unsigned char smem[];
    const NxbModelLayout& ml = p.ml;
    for (int i = threadIdx.x * 16; i < ml.bytes; i += blockDim.x * 16)
        *(uint4*)(smem + i) = *(const uint4*)(p.blob + i);
    __syncthreads();
    const int* parent = (const int*)(smem + ml.off_parent);
    const double* tma = (const double*)(smem + ml.off_tma);
    const double* len = (const double*)(smem + ml.off_len);
    const double* rate = (const double*)(smem + ml.off_rate);
    const double* qval = (const double*)(smem + ml.off_qval);
    const unsigned short* qinfo = (const unsigned short*)(smem + ml.off_qinfo);
    const unsigned short* qptr = (const unsigned short*)(smem + ml.off_qptr);
    const unsigned char* cls = (const unsigned char*)(smem + ml.off_cls);
    const double* pi = (const double*)(smem + ml.off_pi);
    const int E = ml.E, P = ml.P, K = ml.K;
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= p.n_units) return;
    unsigned char* slab = p.state + u * p.state_stride;
    unsigned* g_tol = (unsigned*)(slab + p.so_tol);
    unsigned char* g_npri = (unsigned char*)(slab + p.so_pri);
    double* g_ptm = (double*)(slab + p.so_ptm);
    unsigned* g_pcode = (unsigned*)(slab + p.so_pcode);
    double* g_btm = (double*)(slab + p.so_btm);
    unsigned* g_bcode = (unsigned*)(slab + p.so_bcode);
    unsigned short* bcode_tmp = tmp_code + u * (long long)cap_events;     // edge | track<<... staged unsorted
    double* btime_tmp = tmp_time + u * (long long)cap_events;
    NxbXo xo;
    {
        NxbU4 r = nxb_philox4x32_10((unsigned)(p.unit_offset + u), 0xFFFFFFFEu, 0x07000000u, 0u,
                                    (unsigned)(p.seed & 0xffffffffull), (unsigned)(p.seed >> 32));
        xo.s0 = r.x; xo.s1 = r.y; xo.s2 = r.z; xo.s3 = r.w | 1u;
    }
    auto unif = [&]() { return nxb_u53(nxb_xo_next(xo), nxb_xo_next(xo)); };
    // root
    int s0 = P - 1;
    {
        double target = unif(), acc = 0.0, tot = 0.0;
        for (int s = 0; s < P; ++s) tot += pi[s];
        target *= tot;
        for (int s = 0; s < P; ++s) { acc += pi[s]; if (acc > target) { s0 = s; break; } }
    }
    unsigned m0 = 0u;
    for (int k = 0; k < K; ++k)
        if (k == cls[s0] || unif() < ml.p_on) m0 |= 1u << k;
    if (root_pri) {           // caller-given root state (fwdsample.py:145-160)
        s0 = root_pri[u];
        m0 = root_tol[u] | (1u << cls[s0]);
    }
    g_npri[0] = (unsigned char)s0;
    g_tol[0] = m0;
    int nP = 0, nB = 0;
    bool overflow = false;
    for (int e = 0; e < E; ++e) {
        const int va = parent[e];
        int s = g_npri[va];
        unsigned m = g_tol[va];
        double t = tma[e];
        const double tend = tma[e] + len[e], rho = rate[e];
        if (rho > 0.0 && len[e] > 0.0) {
            while (true) {
                double rpri = 0.0;
                for (int i = qptr[s]; i < qptr[s + 1]; ++i)
                    if ((m >> (qinfo[i] >> 8)) & 1u) rpri += qval[i];
                const int n_on = __popc(m);
                const double r_on = ml.rate_on * (K - n_on), r_off = ml.rate_off * (n_on - 1);
                const double total = rho * (rpri + r_on + r_off);
                if (!(total > 0.0)) break;
                t += -log(unif()) / total;
                if (!(t < tend)) break;
                double x = unif() * (rpri + r_on + r_off);
                if (x < rpri) {
                    int sb = s;
                    for (int i = qptr[s]; i < qptr[s + 1]; ++i)
                        if ((m >> (qinfo[i] >> 8)) & 1u) { sb = qinfo[i] & 0xffu; x -= qval[i]; if (x < 0.0) break; }
                    if (nP < p.gcapP) {
                        g_ptm[nP] = t;
                        g_pcode[nP] = (unsigned)e | ((unsigned)s << 16) | ((unsigned)sb << 24);
                    } else overflow = true;
                    ++nP;
                    s = sb;
                } else {
                    x -= rpri;
                    int kk = -1;
                    if (x < r_on) {            // an OFF class switches on
                        int j = (int)(x / ml.rate_on);
                        for (int k = 0; k < K; ++k) if (!((m >> k) & 1u)) { kk = k; if (j-- == 0) break; }
                    } else {                   // an ON class other than the own one switches off
                        int j = (int)((x - r_on) / ml.rate_off);
                        for (int k = 0; k < K; ++k)
                            if (((m >> k) & 1u) && k != cls[s]) { kk = k; if (j-- == 0) break; }
                    }
                    if (kk >= 0) {
                        m ^= 1u << kk;
                        if (nB < cap_events) {
                            btime_tmp[nB] = t;
                            bcode_tmp[nB] = (unsigned short)(e | (((m >> kk) & 1u) ? 0x8000u : 0u));
                            // track kept in the low bits of the time's companion array below
                            g_bcode[nB] = (unsigned)kk;      // staged: track only (cap_events == gcapB)
                        } else overflow = true;
                        ++nB;
                    }
                }
            }
        }
        g_npri[e + 1] = (unsigned char)s;
        g_tol[e + 1] = m;
    }
    if (nB > p.gcapB || nB > cap_events) overflow = true;
    if (overflow) {
        if (atomicCAS(p.status, 0, NXB_ERR_STATE_CAP) == 0) { p.status[1] = (int)u; p.status[2] = nB; p.status[3] = nP; }
        nP = 0; nB = 0;
    }
    // blink events were generated in (edge, time) order; the slab wants (track, edge, time):
    // stable counting sort by track
    if (nB > 0) {
        // g_bcode[i] currently holds the track of staged event i
        int offs[NXB_MAX_CLASSES + 1];
        for (int k = 0; k <= K; ++k) offs[k] = 0;
        for (int i = 0; i < nB; ++i) offs[g_bcode[i] + 1]++;
        for (int k = 0; k < K; ++k) offs[k + 1] += offs[k];
        // tracks are needed after g_bcode is overwritten: stash them in the spare high bits of
        // the staged code (edge < 2^14, bit 15 = new state, bit 14 unused) -> use a second pass
        for (int i = 0; i < nB; ++i) {
            const unsigned k = g_bcode[i];
            const int dst = offs[k]++;
            g_btm[dst] = btime_tmp[i];
        }
        // rebuild offsets and write the codes
        for (int k = 0; k <= K; ++k) offs[k] = 0;
        for (int i = 0; i < nB; ++i) offs[g_bcode[i] + 1]++;
        for (int k = 0; k < K; ++k) offs[k + 1] += offs[k];
        // codes go to the temporary first (g_bcode still holds the tracks we are reading)
        unsigned* code_out = (unsigned*)btime_tmp;      // reuse: times are already copied out
        for (int i = 0; i < nB; ++i) {
            const unsigned k = g_bcode[i];
            const int dst = offs[k]++;
            const unsigned c16 = bcode_tmp[i];
            code_out[dst] = (c16 & 0x3fffu) | (k << 16) | ((c16 & 0x8000u) ? (1u << 31) : 0u);
        }
        for (int i = 0; i < nB; ++i) g_bcode[i] = code_out[i];
    }
    NxbUnitHeader hdr;
    hdr.sweeps_done = 0; hdr.n_pri = (uint16_t)nP; hdr.n_blk = (uint16_t)nB;
    hdr.phase = NXB_PHASE_PRIMARY; hdr.pad = 0;
    *(NxbUnitHeader*)slab = hdr;
}

__global__ void node_states_kernel(const unsigned char* state, long long stride, long long n, int N,
                                   int so_tol, int so_pri, int* out_pri, unsigned* out_tol) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * N) return;
    const long long u = i / N;
    const int v = (int)(i - u * N);
    const unsigned char* slab = state + u * stride;
    out_pri[i] = slab[so_pri + v];
    out_tol[i] = ((const unsigned*)(slab + so_tol))[v];
}

// rmax_tab[mask] = max_s sum_{b: class(b) in mask} q[s][b]
__global__ void rmax_table_kernel(double* tab, unsigned n_masks, int P, const int* q_ptr,
                                  const int* q_cls, const double* q_val) {
    unsigned m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_masks) return;
    double best = 0.0;
    for (int s = 0; s < P; ++s) {
        double acc = 0.0;
        for (int i = q_ptr[s]; i < q_ptr[s + 1]; ++i)
            if ((m >> q_cls[i]) & 1u) acc += q_val[i];
        best = fmax(best, acc);
    }
    tab[m] = best;
}

__global__ void unit_stats_kernel(const unsigned char* state, long long stride, long long n,
                                  unsigned long long* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long a = 0, b = 0;
    if (i < n) {
        const NxbUnitHeader* h = (const NxbUnitHeader*)(state + i * stride);
        a = h->n_pri; b = h->n_blk;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(FULL, a, o);
        b += __shfl_xor_sync(FULL, b, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
}

}  // namespace nxb

// ===========================================================================
// host side
// ===========================================================================
using namespace nxb;

struct Tier {
    NxbWarpLayout wl;
    int warps;
    int smem_warp0;
    size_t smem;
    // split mode (tier 0 only): shared-memory layout of one unit in blink_kernel
    NxbWarpLayout bwl;
    int blink_warps = 0, blink_units = 0, blink_warp0 = 0, blink_groups = 1;
    // tier 0: leaner layout and warp count of the LEAN primary-only kernel (asynchronous split path)
    NxbWarpLayout pwl;
    int pwarps = 0, pskip = 0, pwarp0 = 0;      // pskip: bytes of the model blob (the tolerance tables) the LEAN kernel does not stage
    size_t psmem = 0;
    size_t blink_smem = 0;
    // event-parallel tolerance update (every steady-state tier): one unit per group of tp_wpu warps
    NxbTolLayout tl;
    int tp_units = 0, tp_wpu = 1, tp_warp0 = 0;
    size_t tp_smem = 0;
};

// what the capacities are derived from: the options of nxb_create and the event-rate sums of the current model
struct CapCfg { double cap_scale = 1.0, rate_on = 1.0, rate_off = 1.0; int diameter = 1, max_warps = 0; double sums[3] = {0, 0, 0}; };

struct nxb_handle {
    CapCfg capcfg;
    int device = 0;
    int num_sms = 0;
    int max_smem = 0;
    bool msg_f64 = false;
    int validate = 0;
    NxbModelLayout ml;
    NxbSummaryLayout sl;
    std::vector<Tier> tiers;        // steady-state sweeps
    std::vector<Tier> init_tiers;   // initial trajectories (raoteh.py:337-361) hold far more events
    unsigned long long* d_perf = nullptr;
    unsigned char* d_gscratch = nullptr;
    long long gscratch_stride = 0;
    int g_off_msg = 0, g_off_ctm = 0, gcapF = 0, gcapC = 0;
    int profile = 0;
    int lockstep = 0;              // warps per lockstep group in tier 0 (0: independent warps)
    bool split = true;             // tier 0: primary and tolerance updates as alternating kernels
    bool tolpar = false;           // true (NXB_TOLPAR=1): tolerance update by tolpar_kernel (event-parallel, experimental: exact but slower); false: blink_kernel (lane per track)
    // shape of the model the handle was created for (nxb_update_model accepts only the same)
    std::vector<int> shape_parent, shape_qptr, shape_qcol, shape_cls;
    std::vector<char> shape_active;     // edge has rate > 0 and length > 0
    unsigned sync_mask = 0x3ffu;
    int gcapP = 0, gcapB = 0;
    int so_tol = 0, so_pri = 0, so_ptm = 0, so_pcode = 0, so_btm = 0, so_bcode = 0;
    long long stride = 0;
    unsigned char* d_blob = nullptr;
    double* d_rmax = nullptr;
    unsigned long long* d_pri_ok = nullptr;
    unsigned* d_tt = nullptr;
    unsigned* d_tf = nullptr;
    long long n_sites = 0;
    unsigned char* d_state = nullptr;
    long long n_units = 0;
    int chains = 0;
    long long unit_offset = 0;
    unsigned long long seed = 0;
    unsigned sweeps = 0;
    unsigned long long* d_counts = nullptr;
    double* d_dwell = nullptr;
    int* d_status = nullptr;
    unsigned long long* d_work = nullptr;
    int* d_defer[2] = {nullptr, nullptr};
    int* d_defer_count = nullptr;
    unsigned long long* d_ctl = nullptr;           // run_split_async: work counters u64[4] | list lengths int[4]
    unsigned long long* d_defer_total = nullptr;   // units listed for a higher tier during the current chunk
    unsigned reached = 0;                          // after NXB_ERR_STATE_CAP: the sweep count the other units have reached
    bool need_reconfig = false;                    // the model blob changed size: tiers must be rebuilt
    double cap_boost = 1.0;                        // doubled when a unit outgrows its slab (slab, pool and message-scratch capacities)
    double cap_means[2] = {0.0, 0.0};              // expected primary / blink events per unit the capacities were sized for
    int discard_pool = 1;                          // NXB_DISCARD_POOL
    int prefetch_wave = 1;                         // NXB_PREFETCH: L2 prefetch of the unit slabs one wave of work ahead
    int blink_ctas = 1;                            // co-resident blink_kernel CTAs per SM (NXB_BLINK_CTAS)
    int bulk_store = 1;                            // NXB_BULK_STORE=0: blink_finish stores the new blink list directly instead of cp.async.bulk
    int blink_groups = 1;                          // NXB_BLINK_GROUPS=0: blink_kernel CTAs as one group (default: groups of 5 warps when the shape allows)
    int smem_per_sm = 0;
    bool sync_sweeps = false;                      // NXB_SYNC_SWEEPS=1: the host-synchronised split path (one wait per sweep)
    std::vector<cudaEvent_t> evpool;
    unsigned long long* d_scratch = nullptr;
    cudaStream_t stream = nullptr;         // the stream every launch and copy goes to
    cudaStream_t own_stream = nullptr;     // created by nxb_create; `stream` may be a caller's (nxb_set_stream)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;     // a launch whose result is collected by the next one
    int pending_kind = -1;
    float last_ms = 0.f;
    int last_launches = 0;
    float kind_ms[4] = {0.f, 0.f, 0.f, 0.f};   // of the last call: fused, primary-only, blink, check
    int kind_launches[4] = {0, 0, 0, 0};
    long long deferred_total = 0, launches_total = 0;
    long long grown_total = 0;                     // times the capacities were doubled because a unit outgrew its slab
    std::string err;
};

extern "C" int nxb_destroy(nxb_handle* h);
static std::string g_create_err;

static int set_err(nxb_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

#define CUDA_TRY(h, call)                                                        \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess)                                                   \
            return set_err(h, NXB_ERR_CUDA, std::string(#call) + ": " +          \
                                                cudaGetErrorString(e_));         \
    } while (0)

static int align_up(int x, int a) { return (x + a - 1) / a * a; }

// device temporaries of one call, released on every exit path
struct DevFree {
    std::vector<void*> ptrs;
    ~DevFree() { for (void* q : ptrs) cudaFree(q); }
};

// lean: the layout of the LEAN primary-only kernel -- no tolerance data, no unit header, no
// tolerance-update scratch, and the arrays of the FFBS passes (staging row, 1/(2 Rmax), uniforms)
// alias the candidate arrays of P1, which are dead once the foreground events have been compacted.
static void build_warp_layout(const NxbModelLayout& ml, bool msg_f64, int capP, int capB,
                              int capR, int capF, int capC, NxbWarpLayout* wl, bool lean = false) {
    const int N = ml.N, E = ml.E, K = ml.K;
    int o = 0;
    auto take = [&](int bytes, int al) { o = align_up(o, al); int r = o; o += bytes; return r; };
    wl->capP = capP; wl->capB = capB; wl->capR = capR; wl->capF = capF; wl->capC = capC;
    wl->off_node_tol = take(4 * N, 4);
    wl->off_node_pri = take(align_up(N, 4), 4);
    wl->off_ptm = take(8 * capP, 8);
    wl->off_pcode = take(4 * capP, 4);
    wl->off_btm = take(8 * capB, 8);
    wl->off_bcode = take(4 * capB, 4);
    wl->off_dpri = take(8 * N, 8);
    wl->off_dtt = lean ? wl->off_dpri : take(4 * N, 4);
    wl->off_dtf = lean ? wl->off_dpri : take(4 * N, 4);
    wl->off_boff = take(4 * (E + 1), 4);
    wl->off_poff = take(4 * (E + 1), 4);
    wl->off_tmpa = take(4 * (E + 1), 4);
    wl->off_tmpb = take(4 * (E + 1), 4);
    wl->off_tmpc = take(4 * (E + 1), 4);
    wl->off_uhdr = lean ? 0 : take(4 * (10 + 32), 4);
    const int scratch0 = align_up(o, 16);
    // primary-phase scratch
    o = scratch0;
    wl->off_bord = take(2 * capB, 2);
    wl->off_rtm = take(8 * capR, 8);
    wl->off_rmask = take(4 * capR, 4);
    const int cand_end = o;
    // lean: the FFBS arrays (first written in P3) and the pointer-jumping scratch of P2 live in the
    // candidate arrays, which are dead once P1 has compacted the foreground events (primary_phase
    // separates the last read from the first write with a __syncwarp / the lockstep barrier)
    const int msz = msg_f64 ? 8 : 4;
    const int ffbs_bytes = align_up(msz * NXB_PPAD, 8) + align_up(msz * capF, 8) + 8 * (capF + 1);
    const int a0 = align_up(wl->off_rtm, 16);
    const bool alias_ffbs = lean && a0 + ffbs_bytes <= cand_end;
    const bool alias_jump = alias_ffbs && a0 + ffbs_bytes + 4 * N <= cand_end;
    wl->off_ftm = take(8 * capF, 8);
    wl->off_fmask = take(4 * capF, 4);
    wl->off_fedge = take(2 * capF, 2);
    wl->off_nchunk = take(4 * N, 4);
    wl->off_jump = alias_jump ? a0 + ffbs_bytes : take(4 * N, 4);
    wl->off_par = take(2 * (capF + 1), 2);
    wl->off_toland = take(4 * (capF + 1), 4);
    wl->off_callow = take(8 * (capF + 1), 8);
    wl->off_cstate = take(capF + 1, 1);
    if (alias_ffbs) {
        int oo = a0;
        wl->off_msg = oo; oo += align_up(msz * NXB_PPAD, 8);
        wl->off_finv = oo; oo += align_up(msz * capF, 8);
        wl->off_fu = oo;
    } else {
        wl->off_msg = take(msz * NXB_PPAD, 16);     // one staging row; the messages are global
        wl->off_finv = take(msz * capF, 8);
        wl->off_fu = take(8 * (capF + 1), 8);
    }
    const int end_pri = o;
    // blink-phase scratch (aliases the primary-phase scratch)
    o = scratch0;
    wl->off_toff = take(4 * (K + 1), 4);
    wl->off_acc = take((msg_f64 ? 32 : 16) * K * ml.nslots, 16);
    wl->off_newtol = take(4 * N, 4);
    wl->off_nrec = take(16 * E, 16);
    // validate_unit needs bord while the persistent state is live; it is only called
    // between phases, when either scratch is dead.
    const int end_blk = o;
    wl->bytes = align_up(lean ? end_pri : std::max(end_pri, end_blk), 16);
}

// Warps per lockstep group.  Several small groups per CTA drift out of phase, so the walks of
// one group's tolerance update (few, fully populated warps) overlap another group's primary
// update; measured for the FUSED kernel on the p53 shape with 20 warps: 5 -> 11.8, 10 -> 11.6, 4 -> 11.3, 20 -> 11.1
// M unit-sweeps/s.  Automatic choice: 5, 4 or 6, whichever leaves the fewest warps unused.
static int group_warps_for(const nxb_handle* h, int warps) {
    if (h->lockstep == 0 || warps < 2) return 1;
    if (h->lockstep > 0) return std::min(warps, h->lockstep);
    if (warps < 8) return warps;
    int best = 5;
    for (int g : {5, 4, 6})
        if (warps / g * g > warps / best * best) best = g;
    return best;
}

// One unit of blink_kernel: the trajectory, the tolerance data, the index arrays of the
// tolerance update and its accumulators -- none of the primary update's scratch.
static void build_blink_layout(const NxbModelLayout& ml, bool msg_f64, int capP, int capB, NxbWarpLayout* wl) {
    const int N = ml.N, E = ml.E, K = ml.K;
    memset(wl, 0, sizeof(*wl));
    int o = 0;
    auto take = [&](int bytes, int al) { o = align_up(o, al); int r = o; o += bytes; return r; };
    wl->capP = capP; wl->capB = capB;
    wl->off_ptm = take(8 * capP, 8);
    wl->off_btm = take(8 * capB, 8);
    wl->off_node_tol = take(4 * N, 4);
    wl->off_node_pri = take(align_up(N, 4), 4);
    wl->off_pcode = take(4 * capP, 4);
    wl->off_bcode = take(2 * capB, 2);               // edges only (u16)
    wl->off_poff = take(4 * (E + 1), 4);
    wl->off_uhdr = take(4 * (10 + 32), 4);
    wl->off_toff = take(4 * (K + 1), 4);
    wl->off_acc = take((msg_f64 ? 32 : 16) * K * ml.nslots, 16);
    wl->off_newtol = take(4 * N, 4);
    wl->off_tmpa = wl->off_newtol;                   // scratch of build_poff, dead before newtol is zeroed
    wl->off_nrec = take(16 * E, 16);
    wl->bytes = align_up(o, 16);
    // Experiment knob: unit stride with a chosen residue mod 128 bytes (the lanes of a warp belong to
    // two units; measured: no effect on the bank-conflict replays, 19.35-19.42 ms for every residue).
    int residue = -1;
    if (const char* ev = getenv("NXB_BLINK_STRIDE_MOD")) residue = atoi(ev);
    if (residue >= 0) while (wl->bytes % 128 != residue % 128) wl->bytes += 16;
    if (getenv("NXB_DEBUG_LAYOUT")) fprintf(stderr, "blink unit layout: %d bytes (payload %d)\n", wl->bytes, o);
}

// One unit of tolpar_kernel: the trajectory, the per-edge records, the event pool (time, key,
// uniform / map bits, 2x2 matrix per live event; the matrix slots first hold the candidates) and
// the node accumulators of the upward tree pass.
static void build_tol_layout(const NxbModelLayout& ml, bool msg_f64, int capP, int capB, int capQ, int capL,
                             NxbTolLayout* tl) {
    const int N = ml.N, E = ml.E, K = ml.K;
    memset(tl, 0, sizeof(*tl));
    int o = 0;
    auto take = [&](int bytes, int al) { o = align_up(o, al); int r = o; o += bytes; return r; };
    tl->capP = capP; tl->capB = capB; tl->capQ = capQ; tl->capL = capL;
    tl->off_q = take((msg_f64 ? 32 : 16) * std::max(capQ, capL), 16);
    tl->off_nrec = take(16 * E, 16);
    tl->off_acc = take((msg_f64 ? 32 : 16) * K * std::max(ml.nslots, ml.nslots_lvl), 16);
    tl->off_T = take(8 * capL, 8);
    tl->off_ptm = take(8 * capP, 8);
    tl->off_btm = take(8 * capB, 8);
    tl->off_aux = take((msg_f64 ? 8 : 4) * capL, 8);
    tl->off_key = take(4 * capL, 4);
    tl->off_pcode = take(4 * capP, 4);
    tl->off_bcode = take(4 * capB, 4);
    tl->off_node_tol = take(4 * N, 4);
    tl->off_newtol = take(4 * N, 4);
    tl->off_toff = take(4 * (K + 2), 4);
    tl->off_f0 = take(4 * E, 4);
    tl->off_f1 = take(4 * E, 4);
    tl->off_hasev = take(4 * E, 4);
    tl->off_x = take(4 * 16, 8);
    tl->off_node_pri = take(align_up(N, 4), 4);
    tl->off_ins = take(2 * capB, 2);
    tl->bytes = align_up(o, 16);
}

extern "C" int nxb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" const char* nxb_last_error(const nxb_handle* h) {
    return h ? h->err.c_str() : g_create_err.c_str();
}

extern "C" void nxb_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    NxbU4 r = nxb_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// Everything that depends on the VALUES of the model (rates, branch lengths, prior): the model
// blob, the scalar constants in h->ml and the table of uniformization rates.  Called by
// nxb_create and, for a model of the same shape, by nxb_update_model (device buffers are reused).
// sums: [0] sum of rate x length, [1] / [2] dominating Poisson means of the primary / tolerance tracks.
#define CUDA_TRY_M(call)                                                                  \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            set_err(nullptr, NXB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
            return NXB_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)
static int build_model_tables(nxb_handle* h, const nxb_model_desc* d, double sums[3]) {
    const int N = d->n_nodes, P = d->n_states, K = d->n_classes, nnz = d->nnz, E = N - 1;
    // ---- derived model quantities ----
    // (built aside and committed only after every check has passed: a refused nxb_update_model
    // leaves the handle as it was)
    NxbModelLayout ml;
    memset(&ml, 0, sizeof(ml));
    ml.N = N; ml.E = E; ml.P = P; ml.K = K; ml.nnz = nnz; ml.diameter = d->diameter;
    ml.rate_on = d->rate_on; ml.rate_off = d->rate_off;
    std::vector<double> tm(N, 0.0);
    for (int v = 1; v < N; ++v) tm[v] = tm[d->parent[v]] + d->blen[v];
    // rates: total out-rate per state with every class on, and its maximum
    double rmax_all = 0.0;
    for (int s = 0; s < P; ++s) {
        double acc = 0.0;
        for (int i = d->q_ptr[s]; i < d->q_ptr[s + 1]; ++i) acc += d->q_val[i];
        rmax_all = std::max(rmax_all, acc);
    }
    if (!(rmax_all > 0.0)) { set_err(nullptr, NXB_ERR_BAD_ARG, "empty rate matrix"); return NXB_ERR_BAD_ARG; }
    ml.inv_omega_pri = 1.0 / (2.0 * rmax_all);
    const double M = std::max(d->rate_on, d->rate_off);

    // ---- model blob ----
    std::vector<unsigned char> blob;
    {
        std::string err;
        // (an update keeps the size of the Poisson table while the new rates fit it; otherwise the
        // blob is rebuilt with a new size and the capacities are configured again)
        bool okb = false;
        if (h->d_blob) {
            std::vector<unsigned char> b2;
            NxbModelLayout ml2 = ml;
            std::string err2;
            if (nxb_build_tol_tables(d, ml2, b2, err2, h->ml.pcdf_cap) == 0) { ml = ml2; blob.swap(b2); okb = true; }
        }
        if (!okb && nxb_build_tol_tables(d, ml, blob, err, 0)) { set_err(nullptr, NXB_ERR_BAD_ARG, err); return NXB_ERR_BAD_ARG; }
    }
    auto put = [&](const void* src, int bytes, int al) {
        int o = align_up((int)blob.size(), al);
        blob.resize(o + bytes);
        memcpy(blob.data() + o, src, bytes);
        return o;
    };
    double& sum_rt = sums[0]; double& lam_pri_sum = sums[1]; double& lam_blk_sum = sums[2];
    sum_rt = 0.0; lam_pri_sum = 0.0; lam_blk_sum = 0.0;
    {
        std::vector<int> parent(E);
        std::vector<double> tma(E), len(E), rate(E), lam_pri(E), p0_pri(E), lam_blk(E), p0_blk(E);
        for (int e = 0; e < E; ++e) {
            parent[e] = d->parent[e + 1];
            tma[e] = tm[parent[e]];
            len[e] = d->blen[e + 1];
            rate[e] = d->rate[e + 1];
            lam_pri[e] = 2.0 * rmax_all * rate[e] * len[e];
            lam_blk[e] = 2.0 * M * rate[e] * len[e];
            p0_pri[e] = exp(-lam_pri[e]);
            p0_blk[e] = exp(-lam_blk[e]);
            sum_rt += rate[e] * len[e];
            lam_pri_sum += lam_pri[e];
            lam_blk_sum += lam_blk[e];
            if (lam_pri[e] > 600.0 || lam_blk[e] > 600.0) {
                set_err(nullptr, NXB_ERR_BAD_ARG, "edge rate * length too large for the tabulated Poisson inversion");
                return NXB_ERR_BAD_ARG;
            }
        }
        // ---- the rest: primary update (the tolerance tables, built by nxb_build_tol_tables above,
        // are the prefix of the blob: blink_kernel stages only that prefix) ----
        std::vector<float> gsc(E, 0.f);
        for (int e = 0; e < E; ++e)
            if (rate[e] > 0.0 && len[e] > 0.0) gsc[e] = (float)(0.6931471805599453 / (2.0 * M * rate[e]));
        ml.off_parent = put(parent.data(), 4 * E, 4);
        {
            // longest-processing-time assignment of the edges to 32 lane slots per round: round r
            // serves the r-th block of 32 edges by decreasing rate x length, odd rounds reversed,
            // so that the longest edges share a lane with the shortest ones
            const int Epad = (E + 31) / 32 * 32;
            std::vector<int> order(E);
            for (int e = 0; e < E; ++e) order[e] = e;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                return rate[a] * len[a] > rate[b] * len[b]; });
            std::vector<unsigned short> eperm(Epad, 0xffffu);
            for (int i = 0; i < Epad; ++i) {
                const int r = i / 32, l = i % 32;
                const int src = r * 32 + ((r & 1) ? 31 - l : l);
                if (src < E) eperm[i] = (unsigned short)order[src];
            }
            ml.off_eperm = put(eperm.data(), 2 * Epad, 2);
        }
        ml.off_tma = put(tma.data(), 8 * E, 8);
        ml.off_len = put(len.data(), 8 * E, 8);
        ml.off_rate = put(rate.data(), 8 * E, 8);
        ml.off_lam_pri = put(lam_pri.data(), 8 * E, 8);
        ml.off_p0_pri = put(p0_pri.data(), 8 * E, 8);
        ml.off_lam_blk = put(lam_blk.data(), 8 * E, 8);
        ml.off_p0_blk = put(p0_blk.data(), 8 * E, 8);
        ml.off_gsc_blk = put(gsc.data(), 4 * E, 4);
        ml.off_qval = put(d->q_val, 8 * nnz, 8);
        std::vector<unsigned short> qinfo(nnz), qptr(P + 1);
        for (int i = 0; i < nnz; ++i)
            qinfo[i] = (unsigned short)(d->q_col[i] | (d->state_class[d->q_col[i]] << 8));
        for (int s = 0; s <= P; ++s) qptr[s] = (unsigned short)d->q_ptr[s];
        ml.off_qinfo = put(qinfo.data(), 2 * nnz, 2);
        ml.off_qptr = put(qptr.data(), 2 * (P + 1), 2);
        ml.off_pi = put(d->prior, 8 * P, 8);
        std::vector<unsigned long long> clsmask(K, 0ull);
        for (int s = 0; s < P; ++s) clsmask[d->state_class[s]] |= 1ull << s;
        ml.off_clsmask = put(clsmask.data(), 8 * K, 8);
        // transposed ELL copy of Q for the branch-free mat-vec: [w][row], padded rows/entries 0
        int W = 1;
        for (int s = 0; s < P; ++s) W = std::max(W, d->q_ptr[s + 1] - d->q_ptr[s]);
        ml.ell_w = W;
        std::vector<unsigned short> ell_info((size_t)W * NXB_PPAD, 0);
        std::vector<float> ell_q32((size_t)W * NXB_PPAD, 0.f);
        std::vector<double> ell_q64((size_t)W * NXB_PPAD, 0.0);
        for (int s = 0; s < NXB_PPAD; ++s)
            for (int w = 0; w < W; ++w) {
                const size_t o = (size_t)w * NXB_PPAD + s;
                ell_info[o] = (unsigned short)(s < P ? s : 0);     // padding: self, rate 0
                if (s < P && d->q_ptr[s] + w < d->q_ptr[s + 1]) {
                    const int i = d->q_ptr[s] + w;
                    ell_info[o] = qinfo[i];
                    ell_q32[o] = (float)d->q_val[i];
                    ell_q64[o] = d->q_val[i];
                }
            }
        ml.off_ell_info = put(ell_info.data(), 2 * W * NXB_PPAD, 2);
        ml.off_ell_q = h->msg_f64 ? put(ell_q64.data(), 8 * W * NXB_PPAD, 8)
                                  : put(ell_q32.data(), 4 * W * NXB_PPAD, 8);
        // dense transposed copy for the one-hot shortcut of the upward pass
        {
            std::vector<float> qdt32((size_t)NXB_PPAD * NXB_PPAD, 0.f);
            std::vector<double> qdt64((size_t)NXB_PPAD * NXB_PPAD, 0.0);
            for (int s = 0; s < P; ++s)
                for (int i = d->q_ptr[s]; i < d->q_ptr[s + 1]; ++i) {
                    qdt32[(size_t)d->q_col[i] * NXB_PPAD + s] = (float)d->q_val[i];
                    qdt64[(size_t)d->q_col[i] * NXB_PPAD + s] = d->q_val[i];
                }
            ml.off_qdt = h->msg_f64 ? put(qdt64.data(), 8 * NXB_PPAD * NXB_PPAD, 8)
                                    : put(qdt32.data(), 4 * NXB_PPAD * NXB_PPAD, 8);
        }
        blob.resize(align_up((int)blob.size(), 16));
        ml.bytes = (int)blob.size();
        ml.off_stats = ml.bytes;
        ml.stats_words = 2 * E + P + 4 + K;
    }
    if (h->d_blob && (ml.bytes != h->ml.bytes || ml.blink_bytes != h->ml.blink_bytes)) {
        // the tables changed size: new device blob; the shared-memory tiers depend on it
        unsigned char* nb = nullptr;
        CUDA_TRY_M(cudaMalloc(&nb, blob.size()));
        cudaFree(h->d_blob);
        h->d_blob = nb;
        h->need_reconfig = true;
    }
    h->ml = ml;
    if (!h->d_blob) CUDA_TRY_M(cudaMalloc(&h->d_blob, blob.size()));
    CUDA_TRY_M(cudaMemcpy(h->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    {
        double inv_n[64];
        inv_n[0] = 0.0;
        for (int i = 1; i < 64; ++i) inv_n[i] = 1.0 / (double)i;
        CUDA_TRY_M(cudaMemcpyToSymbol(c_inv_n, inv_n, sizeof(inv_n)));
    }
    // ---- L2-resident table of uniformization rates, one entry per tolerance mask ----
    h->ml.table_bits = 0;
    if (K <= 22) {
        h->ml.table_bits = K;
        const unsigned n_masks = 1u << K;
        int *dq_ptr, *dq_cls; double* dq_val;
        std::vector<int> qcls(nnz);
        for (int i = 0; i < nnz; ++i) qcls[i] = d->state_class[d->q_col[i]];
        if (!h->d_rmax) CUDA_TRY_M(cudaMalloc(&h->d_rmax, sizeof(double) * n_masks));
        CUDA_TRY_M(cudaMalloc(&dq_ptr, 4 * (P + 1)));
        CUDA_TRY_M(cudaMalloc(&dq_cls, 4 * std::max(nnz, 1)));
        CUDA_TRY_M(cudaMalloc(&dq_val, 8 * std::max(nnz, 1)));
        CUDA_TRY_M(cudaMemcpy(dq_ptr, d->q_ptr, 4 * (P + 1), cudaMemcpyHostToDevice));
        CUDA_TRY_M(cudaMemcpy(dq_cls, qcls.data(), 4 * nnz, cudaMemcpyHostToDevice));
        CUDA_TRY_M(cudaMemcpy(dq_val, d->q_val, 8 * nnz, cudaMemcpyHostToDevice));
        rmax_table_kernel<<<(n_masks + 255) / 256, 256, 0, h->stream>>>(h->d_rmax, n_masks, P, dq_ptr, dq_cls, dq_val);
        CUDA_TRY_M(cudaGetLastError());
        CUDA_TRY_M(cudaStreamSynchronize(h->stream));
        cudaFree(dq_ptr); cudaFree(dq_cls); cudaFree(dq_val);
    }
    return NXB_OK;
}

static int check_desc(const nxb_model_desc* d) {
    const int N = d->n_nodes, P = d->n_states, K = d->n_classes, nnz = d->nnz;
    if (N < 2 || N > 16384) return set_err(nullptr, NXB_ERR_BAD_ARG, "need 2 <= n_nodes <= 16384");
    if (P < 1 || P > NXB_MAX_STATES) return set_err(nullptr, NXB_ERR_BAD_ARG, "need 1 <= n_states <= 64");
    if (K < 1 || K > NXB_MAX_CLASSES) return set_err(nullptr, NXB_ERR_BAD_ARG, "need 1 <= n_classes <= 32");
    if (!(d->rate_on > 0.0) || !(d->rate_off > 0.0))
        return set_err(nullptr, NXB_ERR_BAD_ARG, "blink rates must be positive");
    if (d->parent[0] != -1) return set_err(nullptr, NXB_ERR_BAD_ARG, "parent[0] must be -1 (root)");
    for (int v = 1; v < N; ++v)
        if (d->parent[v] < 0 || d->parent[v] >= v)
            return set_err(nullptr, NXB_ERR_BAD_ARG, "nodes must be numbered so that parent[v] < v");
    for (int v = 1; v < N; ++v)
        if (!(d->blen[v] >= 0.0) || !(d->rate[v] >= 0.0))
            return set_err(nullptr, NXB_ERR_BAD_ARG, "branch lengths and rates must be >= 0");
    if (d->q_ptr[0] != 0 || d->q_ptr[P] != nnz) return set_err(nullptr, NXB_ERR_BAD_ARG, "bad q_ptr");
    for (int s = 0; s < P; ++s) {
        if (d->state_class[s] < 0 || d->state_class[s] >= K)
            return set_err(nullptr, NXB_ERR_BAD_ARG, "state_class out of range");
        if (d->q_ptr[s + 1] - d->q_ptr[s] > 31)
            return set_err(nullptr, NXB_ERR_BAD_ARG, "at most 31 transitions per primary state");
        for (int i = d->q_ptr[s]; i < d->q_ptr[s + 1]; ++i) {
            if (d->q_col[i] < 0 || d->q_col[i] >= P || d->q_col[i] == s || !(d->q_val[i] > 0.0))
                return set_err(nullptr, NXB_ERR_BAD_ARG, "bad primary rate matrix entry");
        }
    }
    return NXB_OK;
}

// units that have completed fewer than `reached` sweeps (they outgrew their slab and were left behind)
__global__ void collect_behind_kernel(const unsigned char* state, long long stride, long long n_units,
                                      unsigned reached, int* list, int* count) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const NxbUnitHeader* h = (const NxbUnitHeader*)(state + u * stride);
    if (h->sweeps_done < reached) list[atomicAdd(count, 1)] = (int)u;
}

// move the live part of every unit slab to a new slab layout (larger event capacities)
struct ReslabArgs {
    const unsigned char* src; unsigned char* dst;
    long long src_stride, dst_stride, n_units;
    int so[6], dn[6];       // tol, pri, ptm, pcode, btm, bcode offsets: old / new
    int N;
};
__global__ void reslab_kernel(const ReslabArgs a) {
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long u = w0; u < a.n_units; u += nw) {
        const unsigned char* s = a.src + u * a.src_stride;
        unsigned char* d = a.dst + u * a.dst_stride;
        const NxbUnitHeader hdr = *(const NxbUnitHeader*)s;
        if (lane == 0) *(NxbUnitHeader*)d = hdr;
        for (int v = lane; v < a.N; v += 32) {
            ((unsigned*)(d + a.dn[0]))[v] = ((const unsigned*)(s + a.so[0]))[v];
            d[a.dn[1] + v] = s[a.so[1] + v];
        }
        for (int i = lane; i < (int)hdr.n_pri; i += 32) {
            ((double*)(d + a.dn[2]))[i] = ((const double*)(s + a.so[2]))[i];
            ((unsigned*)(d + a.dn[3]))[i] = ((const unsigned*)(s + a.so[3]))[i];
        }
        for (int i = lane; i < (int)hdr.n_blk; i += 32) {
            ((double*)(d + a.dn[4]))[i] = ((const double*)(s + a.so[4]))[i];
            ((unsigned*)(d + a.dn[5]))[i] = ((const unsigned*)(s + a.so[5]))[i];
        }
    }
}

// Capacity tiers (shared memory), slab layout and global scratch from the expected event counts
// of the current model; `h->cap_boost` scales the HBM-side capacities (doubled when a unit outgrows its
// slab or the largest tier).  Called by nxb_create, by nxb_update_model when the new rates need
// more room than the handle has, and by the sweep when a unit hits a capacity: existing unit
// states are moved to the new slab layout (reslab_kernel), so chains continue.
static int configure_capacities(nxb_handle* h, const CapCfg& cfg) {
    const NxbModelLayout& ml = h->ml;
    const int N = ml.N, E = ml.E, K = ml.K;
    const double sum_rt = cfg.sums[0], lam_pri_sum = cfg.sums[1], lam_blk_sum = cfg.sums[2];
    // the layout the existing unit states were written with
    const int old_off[6] = {h->so_tol, h->so_pri, h->so_ptm, h->so_pcode, h->so_btm, h->so_bcode};
    const long long old_stride = h->stride;
    // ---- capacity tiers ----
    {
        const double scale = (cfg.cap_scale > 0.0) ? cfg.cap_scale : 1.0;
        const double boost = h->cap_boost;      // HBM-side capacities only: the shared-memory tiers keep their sizes (and occupancy)
        const double on = cfg.rate_on, off = cfg.rate_off;
        const double mean_blk = K * sum_rt * 2.0 * on * off / (on + off);
        const double mean_pri = sum_rt;
        const double cand_pri = lam_pri_sum, cand_blk = K * lam_blk_sum;
        auto cap = [&](double mean, double nsig, double slack) {
            return align_up((int)ceil(scale * (mean + nsig * sqrt(std::max(mean, 1.0)) + slack)), 8);
        };
        int capP = cap(mean_pri, 4.0, 8.0);
        int capB = cap(mean_blk, 4.0, 2.0 * E < 32 ? 2.0 * E * K : 16.0);
        // the initial trajectory holds up to two blink events per (track, active edge)
        // for tracks that the data force off (raoteh.py:205-229)
        int capF = cap(mean_pri + 0.75 * cand_pri + (double)cfg.diameter * 0, 4.0, 8.0);
        int capR = cap(mean_pri + cand_pri, 4.0, 8.0) + capP;
        int capC = cap(mean_blk + cand_blk, 4.0, 16.0);
        h->gcapP = std::max(h->gcapP, std::min(65535, align_up((int)(boost * (4 * capP + 32)), 8)));      // (slabs never shrink)
        h->gcapB = std::max(h->gcapB, std::min(65535, align_up((int)(boost * (4 * capB + 64)), 8)));
        h->tiers.clear(); h->init_tiers.clear();
        const int stats_bytes = align_up(8 * (ml.stats_words + 16), 16);   // + lockstep batch slots
        const int fixed = ml.bytes + stats_bytes;
        const int max_warps = (cfg.max_warps > 0) ? std::min(cfg.max_warps, NXB_MAX_THREADS / 32)
                                                 : NXB_MAX_THREADS / 32;
        for (int pass = 0; pass < 2; ++pass) {
            for (int t = 0; t < 8; ++t) {
                Tier tier;
                const int f = 1 << t;
                int cP = capP * f, cB = capB * f, cR = capR * f, cF = capF * f, cC = capC * f;
                if (pass == 1) {
                    // init pass: `diameter` forced events on every edge, all of which may
                    // survive; up to two blink events per (track, edge)
                    cF = std::max(cF, E * cfg.diameter + 8);
                    cR = std::max(cR, E * cfg.diameter + 8);
                    cP = std::max(cP, std::min(E * cfg.diameter + 8, 4 * capP + 32));
                }
                build_warp_layout(ml, h->msg_f64, std::min(cP, h->gcapP), std::min(cB, h->gcapB),
                                  cR, cF, cC, &tier.wl);
                tier.smem_warp0 = fixed;
                int w = (h->max_smem - fixed) / tier.wl.bytes;
                if (w < 1) break;
                tier.warps = std::min(w, max_warps);
                if (getenv("NXB_DEBUG_LAYOUT"))
                    fprintf(stderr, "tier %d pass %d: fixed %d, per warp %d bytes (capP %d capB %d capR %d capF %d), fit %d warps, used %d\n",
                            t, pass, fixed, tier.wl.bytes, tier.wl.capP, tier.wl.capB, cR, cF, w, tier.warps);
                tier.smem = (size_t)fixed + (size_t)tier.warps * tier.wl.bytes;
                if (pass == 0 && t == 0) {
                    build_warp_layout(ml, h->msg_f64, tier.wl.capP, tier.wl.capB, cR, cF, cC, &tier.pwl, true);
                    {
                        // f32 messages: 80 registers per thread leave room for 24 warps (6 per SM sub-partition)
                        const int cap_w = (cfg.max_warps > 0) ? max_warps : (h->msg_f64 ? NXB_MAX_THREADS / 32 : NXB_LEAN_THREADS / 32);
                        // the primary update reads none of the tolerance tables (the prefix of the blob)
                        tier.pskip = ml.blink_bytes;
                        tier.pwarp0 = fixed - tier.pskip;
                        tier.pwarps = std::min((h->max_smem - tier.pwarp0) / tier.pwl.bytes, cap_w);
                        if (getenv("NXB_LEAN_WARPS")) tier.pwarps = std::min(tier.pwarps, std::max(1, atoi(getenv("NXB_LEAN_WARPS"))));
                        tier.psmem = (size_t)tier.pwarp0 + (size_t)tier.pwarps * tier.pwl.bytes;
                        if (getenv("NXB_DEBUG_LAYOUT"))
                            fprintf(stderr, "lean primary layout: fixed %d, %d bytes per warp, %d warps\n", tier.pwarp0, tier.pwl.bytes, tier.pwarps);
                    }
                    build_blink_layout(ml, h->msg_f64, tier.wl.capP, tier.wl.capB, &tier.bwl);
                    const int bfixed = ml.blink_bytes + stats_bytes;      // model prefix + statistic partials
                    tier.blink_warp0 = bfixed;
                    // NXB_BLINK_CTAS=2: two co-resident CTAs per SM, each with half of the shared memory, so
                    // that one CTA's load / set-up / write-back overlaps the other's walks
                    int smem_budget = h->max_smem;
                    if (h->blink_ctas > 1) smem_budget = std::min(h->max_smem, h->smem_per_sm / h->blink_ctas - 1024);
                    const int fit = (smem_budget - bfixed) / tier.bwl.bytes;
                    int bw = (cfg.max_warps > 0) ? std::min(cfg.max_warps, NXB_BLINK_THREADS / 32)
                                                : NXB_BLINK_THREADS / 32;
                    if (const char* ev = getenv("NXB_BLINK_WARPS")) bw = std::max(1, std::min(atoi(ev), NXB_BLINK_THREADS / 32));
                    // one (unit, track) job per lane of the CTA
                    int bu = std::min(fit, (32 * bw) / K);
                    // prefer a batch whose jobs fill whole warps (p53: 32 units x 20 tracks = 20 warps;
                    // measured 14.9 M unit-sweeps/s against 14.1-14.6 M with 21-23 ragged warps)
                    for (int cand = bu; cand >= 1 && 4 * cand >= 3 * bu; --cand)
                        if ((cand * K) % 32 == 0) { bu = cand; break; }
                    if (bu >= 1) {
                        tier.blink_units = bu;
                        tier.blink_warps = std::min(bw, (bu * K + 31) / 32);
                        tier.blink_smem = (size_t)bfixed + (size_t)bu * tier.bwl.bytes;
                        // independent warp groups (NXB_BLINK_GROUPS): whole warps and whole units per group
                        const int g = tier.blink_warps / NXB_BLINK_GROUP_WARPS;
                        tier.blink_groups = (h->blink_groups != 0 && g > 1 && g <= 15 && tier.blink_warps % NXB_BLINK_GROUP_WARPS == 0 &&
                                             bu % g == 0 && (bu / g) * K <= 32 * NXB_BLINK_GROUP_WARPS) ? g : 1;
                    }
                }
                if (pass == 0 && h->tolpar) {
                    // candidates of one sweep (all tracks) and live events (accepted candidates + retained):
                    // the acceptance probability is that of the stationary mix of ON and OFF
                    const double Mx = std::max(on, off);
                    const double acc_mean = (off / (on + off)) * (2.0 * Mx - on) / (2.0 * Mx)
                                          + (on / (on + off)) * (2.0 * Mx - off) / (2.0 * Mx);
                    const int capQ0 = cap(cand_blk, 3.5, 16.0);
                    const int capL0 = cap(mean_blk + acc_mean * cand_blk, 3.5, 16.0);
                    const int cQ = std::min(65528, capQ0 * f), cL = std::min(65528, capL0 * f);
                    build_tol_layout(ml, h->msg_f64, tier.wl.capP, tier.wl.capB, cQ, cL, &tier.tl);
                    const int bfixed = ml.blink_bytes + stats_bytes;
                    tier.tp_warp0 = bfixed;
                    const int fit = (h->max_smem - bfixed) / tier.tl.bytes;
                    int wpu = (cand_blk + mean_blk >= 128.0) ? 2 : 1;
                    if (const char* ev = getenv("NXB_TP_WPU")) wpu = atoi(ev) == 2 ? 2 : 1;
                    int maxw = NXB_TP_THREADS / 32;
                    if (cfg.max_warps > 0) maxw = std::max(wpu, std::min(maxw, cfg.max_warps));
                    int nu = std::min(fit, maxw / wpu);
                    if (wpu == 2) nu = std::min(nu, 15);          // one named barrier per unit group
                    if (nu >= 1) {
                        tier.tp_units = nu; tier.tp_wpu = wpu;
                        tier.tp_smem = (size_t)bfixed + (size_t)nu * tier.tl.bytes;
                    }
                }
                (pass == 0 ? h->tiers : h->init_tiers).push_back(tier);
            }
        }
        if (h->tiers.empty() || h->init_tiers.empty()) {
            set_err(nullptr, NXB_ERR_BAD_ARG, "model too large: one unit does not fit in shared memory");
            return NXB_ERR_BAD_ARG;
        }
        // global scratch per warp slot: chunk messages + candidate pool
        {
            int max_w = 1;
            for (auto& t : h->tiers) max_w = std::max(max_w, std::max(t.warps, t.pwarps));
            for (auto& t : h->init_tiers) max_w = std::max(max_w, t.warps);
            max_w = std::max(max_w, h->tiers[0].blink_units * h->blink_ctas);
            h->gcapF = std::max((int)(boost * 8 * capF), E * cfg.diameter + 8);
            // live events per track and sweep: retained + accepted candidates
            {
                const double per_track = (mean_blk + cand_blk) / K;
                h->gcapC = std::min(32760, align_up((int)(boost * (4.0 * per_track + 10.0 * sqrt(per_track + 1.0) + 64.0)), 8));
            }
            int o = 0;
            h->g_off_msg = o; o += (h->msg_f64 ? 8 : 4) * (h->gcapF + 1) * NXB_PPAD;
            o = align_up(o, 128);         // (whole 128-byte lines per track region: discard.global.L2)
            h->g_off_ctm = o; o += (h->msg_f64 ? 32 : 16) * h->gcapC * K;      // one region per track, 16 B (f32) / 32 B (f64) records
            h->gscratch_stride = align_up(o, 256);
            if (h->d_gscratch) { cudaFree(h->d_gscratch); h->d_gscratch = nullptr; }
            cudaError_t e1 = cudaMalloc(&h->d_gscratch, (size_t)h->gscratch_stride * max_w * h->num_sms);
            if (e1 != cudaSuccess) { set_err(nullptr, NXB_ERR_CUDA, cudaGetErrorString(e1)); return NXB_ERR_CUDA; }
        }
        // slab layout
        int o = (int)sizeof(NxbUnitHeader);
        h->so_tol = o; o += 4 * N;
        h->so_pri = o; o += align_up(N, 4);
        o = align_up(o, 16);           // (16-byte sections: bulk-async copies)
        h->so_ptm = o; o += 8 * h->gcapP;
        h->so_pcode = o; o += 4 * h->gcapP;
        o = align_up(o, 16);
        h->so_btm = o; o += 8 * h->gcapB;
        h->so_bcode = o; o += 4 * h->gcapB;
        h->stride = align_up(o, 16);
    }
    h->cap_means[0] = sum_rt; h->cap_means[1] = K * sum_rt * 2.0 * cfg.rate_on * cfg.rate_off / (cfg.rate_on + cfg.rate_off);
    if (h->d_state && h->n_units > 0 && old_stride > 0) {
        const int new_off[6] = {h->so_tol, h->so_pri, h->so_ptm, h->so_pcode, h->so_btm, h->so_bcode};
        bool same = old_stride == h->stride;
        for (int i = 0; i < 6; ++i) same = same && old_off[i] == new_off[i];
        if (!same) {
            unsigned char* ns = nullptr;
            cudaError_t e1 = cudaMalloc(&ns, (size_t)h->n_units * h->stride);
            if (e1 != cudaSuccess) { set_err(h, NXB_ERR_CUDA, cudaGetErrorString(e1)); return NXB_ERR_CUDA; }
            ReslabArgs ra;
            ra.src = h->d_state; ra.dst = ns; ra.src_stride = old_stride; ra.dst_stride = h->stride;
            for (int i = 0; i < 6; ++i) { ra.so[i] = old_off[i]; ra.dn[i] = new_off[i]; }
            ra.N = N; ra.n_units = h->n_units;
            reslab_kernel<<<(unsigned)std::min<long long>((h->n_units + 7) / 8, 148LL * 32), 256, 0, h->stream>>>(ra);
            e1 = cudaStreamSynchronize(h->stream);
            if (e1 != cudaSuccess) { cudaFree(ns); set_err(h, NXB_ERR_CUDA, cudaGetErrorString(e1)); return NXB_ERR_CUDA; }
            cudaFree(h->d_state);
            h->d_state = ns;
        }
    }
    return NXB_OK;
}

extern "C" int nxb_create(const nxb_model_desc* d, int device, nxb_handle** out) {
    if (!d || !out) return set_err(nullptr, NXB_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    const int N = d->n_nodes, P = d->n_states, K = d->n_classes, nnz = d->nnz, E = N - 1;
    { int rc = check_desc(d); if (rc) return rc; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return set_err(nullptr, NXB_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err(nullptr, NXB_ERR_BAD_ARG, "bad device index");

    nxb_handle* h = new nxb_handle();
    h->device = device;
    h->msg_f64 = d->msg_f64 != 0;
    h->validate = d->validate;
    if (getenv("NXB_SYNC_MASK")) h->sync_mask = (unsigned)strtoul(getenv("NXB_SYNC_MASK"), nullptr, 0);
    // 0: automatic group size (group_warps_for / run_split); negative switches lockstep off
    h->lockstep = d->lockstep_warps == 0 ? -1 : (d->lockstep_warps < 0 ? 0 : d->lockstep_warps);   // -1: automatic
    h->split = d->split_phases >= 0;
    if (const char* ev = getenv("NXB_TOLPAR")) h->tolpar = atoi(ev) != 0;
    if (const char* ev = getenv("NXB_SYNC_SWEEPS")) h->sync_sweeps = atoi(ev) != 0;
    if (E > 8191) return (delete h, set_err(nullptr, NXB_ERR_BAD_ARG, "more than 8191 edges"));
#define CUDA_TRY_C(call)                                                                  \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            set_err(nullptr, NXB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
            nxb_destroy(h);                                                               \
            return NXB_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)
    CUDA_TRY_C(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY_C(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    h->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    if (const char* ev = getenv("NXB_DISCARD_POOL")) h->discard_pool = atoi(ev) != 0;
    if (const char* ev = getenv("NXB_PREFETCH")) h->prefetch_wave = atoi(ev) != 0;
    if (const char* ev = getenv("NXB_BLINK_CTAS")) h->blink_ctas = std::max(1, std::min(4, atoi(ev)));
    if (const char* ev = getenv("NXB_BLINK_GROUPS")) h->blink_groups = atoi(ev) != 0;
    if (const char* ev = getenv("NXB_BULK_STORE")) h->bulk_store = atoi(ev) != 0;
    CUDA_TRY_C(cudaStreamCreate(&h->own_stream));
    h->stream = h->own_stream;
    CUDA_TRY_C(cudaEventCreate(&h->ev0));
    CUDA_TRY_C(cudaEventCreate(&h->ev1));
    CUDA_TRY_C(cudaEventCreate(&h->ev2));
    CUDA_TRY_C(cudaEventCreate(&h->ev3));

    // ---- model tables ----
    double sums[3];
    {
        int rc = build_model_tables(h, d, sums);
        if (rc) { nxb_destroy(h); return rc; }
    }
    NxbModelLayout& ml = h->ml;
    const double sum_rt = sums[0], lam_pri_sum = sums[1], lam_blk_sum = sums[2];
    // ---- statistics ----
    {
        NxbSummaryLayout& sl = h->sl;
        int o = 0;
        sl.c_root_pri = o; o += P;
        sl.c_root_on = o; o += 1;
        sl.c_root_off = o; o += 1;
        sl.c_root_on_cls = o; o += K;
        sl.c_pri_trans = o; o += E * nnz;
        sl.c_off_on = o; o += E;
        sl.c_on_off = o; o += E;
        sl.c_nsamples = o; o += 1;
        sl.n_counts = o;
        o = 0;
        sl.d_T = o; o += E * P;
        sl.d_off = o; o += E * P * K;
        sl.n_dwell = o;
        CUDA_TRY_C(cudaMalloc(&h->d_counts, 8 * (size_t)sl.n_counts));
        CUDA_TRY_C(cudaMalloc(&h->d_dwell, 8 * (size_t)sl.n_dwell));
        CUDA_TRY_C(cudaMemset(h->d_counts, 0, 8 * (size_t)sl.n_counts));
        CUDA_TRY_C(cudaMemset(h->d_dwell, 0, 8 * (size_t)sl.n_dwell));
    }
    CUDA_TRY_C(cudaMalloc(&h->d_status, 4 * sizeof(int)));
    CUDA_TRY_C(cudaMemset(h->d_status, 0, 4 * sizeof(int)));
    CUDA_TRY_C(cudaMalloc(&h->d_work, 8));
    CUDA_TRY_C(cudaMalloc(&h->d_defer_count, 4));
    CUDA_TRY_C(cudaMalloc(&h->d_ctl, 48));
    CUDA_TRY_C(cudaMalloc(&h->d_defer_total, 8));
    CUDA_TRY_C(cudaMalloc(&h->d_scratch, 64));
    CUDA_TRY_C(cudaMalloc(&h->d_perf, 32 * 8));
    CUDA_TRY_C(cudaMemset(h->d_perf, 0, 32 * 8));

    // ---- capacity tiers, slab layout, global scratch ----
    h->capcfg.cap_scale = d->cap_scale; h->capcfg.rate_on = d->rate_on; h->capcfg.rate_off = d->rate_off;
    h->capcfg.diameter = d->diameter; h->capcfg.max_warps = d->max_warps;
    h->capcfg.sums[0] = sum_rt; h->capcfg.sums[1] = lam_pri_sum; h->capcfg.sums[2] = lam_blk_sum;
    {
        int rc = configure_capacities(h, h->capcfg);
        if (rc) { nxb_destroy(h); return rc; }
    }
    {
        cudaError_t e1 = h->msg_f64
            ? cudaFuncSetAttribute(sweep_kernel<double, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem)
            : cudaFuncSetAttribute(sweep_kernel<float, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
        if (e1 == cudaSuccess) e1 = h->msg_f64
            ? cudaFuncSetAttribute(sweep_kernel<double, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem)
            : cudaFuncSetAttribute(sweep_kernel<float, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
        if (e1 == cudaSuccess) e1 = h->msg_f64
            ? cudaFuncSetAttribute(sweep_kernel<double, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem)
            : cudaFuncSetAttribute(sweep_kernel<float, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
        if (e1 == cudaSuccess) e1 = h->msg_f64
            ? cudaFuncSetAttribute(blink_kernel<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem)
            : cudaFuncSetAttribute(blink_kernel<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
        if (e1 == cudaSuccess) e1 = h->msg_f64
            ? cudaFuncSetAttribute(blink_kernel<double, NXB_BLINK_GROUP_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem)
            : cudaFuncSetAttribute(blink_kernel<float, NXB_BLINK_GROUP_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
        if (e1 != cudaSuccess) { set_err(nullptr, NXB_ERR_CUDA, cudaGetErrorString(e1)); nxb_destroy(h); return NXB_ERR_CUDA; }
    }
    if (h->msg_f64) {
        cudaFuncSetAttribute(tolpar_kernel<double, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
        cudaFuncSetAttribute(tolpar_kernel<double, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
    } else {
        cudaFuncSetAttribute(tolpar_kernel<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
        cudaFuncSetAttribute(tolpar_kernel<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
    }
    cudaFuncSetAttribute(summarize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
    h->shape_parent.assign(d->parent, d->parent + N);
    h->shape_qptr.assign(d->q_ptr, d->q_ptr + P + 1);
    h->shape_qcol.assign(d->q_col, d->q_col + nnz);
    h->shape_cls.assign(d->state_class, d->state_class + P);
    h->shape_active.resize(E);
    for (int e = 0; e < E; ++e) h->shape_active[e] = (d->blen[e + 1] > 0.0 && d->rate[e + 1] > 0.0);
    *out = h;
    return NXB_OK;
}

extern "C" int nxb_update_model(nxb_handle* h, const nxb_model_desc* d) {
    if (!h || !d) return set_err(h, NXB_ERR_BAD_ARG, "null argument");
    { int rc = check_desc(d); if (rc) { h->err = g_create_err; return rc; } }
    const int N = d->n_nodes, P = d->n_states, nnz = d->nnz, E = N - 1;
    const NxbModelLayout& ml = h->ml;
    if (N != ml.N || P != ml.P || d->n_classes != ml.K || nnz != ml.nnz || (d->msg_f64 != 0) != h->msg_f64 ||
        !std::equal(d->parent, d->parent + N, h->shape_parent.begin()) ||
        !std::equal(d->q_ptr, d->q_ptr + P + 1, h->shape_qptr.begin()) ||
        !std::equal(d->q_col, d->q_col + nnz, h->shape_qcol.begin()) ||
        !std::equal(d->state_class, d->state_class + P, h->shape_cls.begin()))
        return set_err(h, NXB_ERR_BAD_ARG, "nxb_update_model: the model must have the shape the handle was created for "
                                           "(tree, sparsity pattern of Q, classes, message type)");
    if (d->diameter != ml.diameter) return set_err(h, NXB_ERR_BAD_ARG, "nxb_update_model: diameter changed");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    double sums[3];
    int rc = build_model_tables(h, d, sums);
    if (rc) { h->err = g_create_err; return rc; }
    // Capacities follow the model: when the new rates mean more events per unit than the
    // shared-memory tiers and slabs were sized for (or far fewer), they are derived again and the
    // unit states move to the new slab layout; chains continue.
    {
        h->capcfg.rate_on = d->rate_on; h->capcfg.rate_off = d->rate_off;
        for (int i = 0; i < 3; ++i) h->capcfg.sums[i] = sums[i];
        const double mp = sums[0], mb = ml.K * sums[0] * 2.0 * d->rate_on * d->rate_off / (d->rate_on + d->rate_off);
        const bool grew = mp > 1.1 * h->cap_means[0] + 1.0 || mb > 1.1 * h->cap_means[1] + 1.0;
        const bool shrank = mp < 0.5 * h->cap_means[0] - 1.0 || mb < 0.5 * h->cap_means[1] - 1.0;
        if (h->need_reconfig || grew || shrank) {
            h->need_reconfig = false;
            rc = configure_capacities(h, h->capcfg);
            if (rc) return rc;
        }
    }
    // Trajectories stay valid under new rates unless an edge switched between active and
    // inactive (events on a zero-rate edge); then the chains have to be initialised again.
    bool same_active = true;
    for (int e = 0; e < E; ++e) {
        const char act = (d->blen[e + 1] > 0.0 && d->rate[e + 1] > 0.0);
        if (act != h->shape_active[e]) same_active = false;
        h->shape_active[e] = act;
    }
    if (!same_active && h->d_state) { cudaFree(h->d_state); h->d_state = nullptr; h->n_units = 0; }
    return NXB_OK;
}

extern "C" int nxb_destroy(nxb_handle* h) {
    if (!h) return NXB_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_blob); cudaFree(h->d_rmax); cudaFree(h->d_pri_ok); cudaFree(h->d_tt); cudaFree(h->d_tf);
    cudaFree(h->d_state); cudaFree(h->d_counts); cudaFree(h->d_dwell); cudaFree(h->d_status);
    cudaFree(h->d_work); cudaFree(h->d_defer[0]); cudaFree(h->d_defer[1]); cudaFree(h->d_defer_count);
    cudaFree(h->d_scratch); cudaFree(h->d_perf); cudaFree(h->d_gscratch);
    cudaFree(h->d_ctl); cudaFree(h->d_defer_total);
    for (auto& e : h->evpool) if (e) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    if (h->ev3) cudaEventDestroy(h->ev3);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return NXB_OK;
}

extern "C" int nxb_set_stream(nxb_handle* h, uint64_t cuda_stream) {
    if (!h) return NXB_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));      // nothing of ours is left in flight on the old stream
    h->stream = cuda_stream ? (cudaStream_t)(uintptr_t)cuda_stream : h->own_stream;
    return NXB_OK;
}

extern "C" int nxb_set_data(nxb_handle* h, int64_t n_sites, const uint64_t* pri_ok,
                            const uint32_t* tt, const uint32_t* tf) {
    if (!h || n_sites <= 0 || !pri_ok || !tt || !tf) return set_err(h, NXB_ERR_BAD_ARG, "bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)n_sites * h->ml.N;
    if (n_sites != h->n_sites) {
        // the unit states belong to the old sites (unit u is chain u % chains of site u / chains):
        // drop them, so that a sweep without a new nxb_init_chains fails instead of reading the
        // data of other sites or past the end of the new buffers
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (h->d_state) { cudaFree(h->d_state); h->d_state = nullptr; }
        h->n_units = 0;
        cudaFree(h->d_pri_ok); cudaFree(h->d_tt); cudaFree(h->d_tf);
        h->d_pri_ok = nullptr; h->d_tt = nullptr; h->d_tf = nullptr;
        h->n_sites = 0;
        CUDA_TRY(h, cudaMalloc(&h->d_pri_ok, 8 * n));
        CUDA_TRY(h, cudaMalloc(&h->d_tt, 4 * n));
        CUDA_TRY(h, cudaMalloc(&h->d_tf, 4 * n));
        h->n_sites = n_sites;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->d_pri_ok, pri_ok, 8 * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_tt, tt, 4 * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_tf, tf, 4 * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NXB_OK;
}

static const char* status_text(int code) {
    switch (code) {
        case NXB_ERR_INFEASIBLE: return "infeasible data: the forward-filtering pass met an all-zero message";
        case NXB_ERR_ZERO_RATE: return "zero max rate";
        case NXB_ERR_STATE_CAP: return "a unit outgrew its HBM slab; re-create the handle with a larger cap_scale";
        case NXB_ERR_INVARIANT: return "trajectory invariant violated";
        case NXB_ERR_SCRATCH_CAP: return "scratch overflow at the largest shared-memory tier; increase cap_scale";
        default: return "error";
    }
}

enum { KIND_FUSED = 0, KIND_PRIMARY = 1, KIND_BLINK = 2, KIND_CHECK = 3, KIND_PRIMARY_LEAN = 4 };   // (LEAN: tier 0, pwl / pwarps)

// Kernel parameters of one launch over `list` (or every unit).
static void fill_params(nxb_handle* h, NxbSweepParams& p, int kind, const Tier& tier, int group_warps,
                        const int* list, int n_list, int* out, unsigned begin, unsigned target,
                        unsigned accum_from, int init) {
    memset(&p, 0, sizeof(p));
    p.ml = h->ml; p.wl = (kind == KIND_BLINK) ? tier.bwl : (kind == KIND_PRIMARY_LEAN ? tier.pwl : tier.wl); p.sl = h->sl;
    p.smem_warp0 = tier.smem_warp0;
    const bool tolpar = kind == KIND_BLINK && h->tolpar && tier.tp_units > 0;
    if (kind == KIND_BLINK) {       // only the tolerance tables of the model are staged
        p.ml.bytes = h->ml.blink_bytes; p.ml.off_stats = h->ml.blink_bytes;
        p.smem_warp0 = tolpar ? tier.tp_warp0 : tier.blink_warp0;
        p.tl = tier.tl;
    }
    p.blob = h->d_blob; p.rmax_tab = h->d_rmax;
    if (kind == KIND_PRIMARY_LEAN && tier.pskip > 0) {
        // stage blob[pskip:] at shared offset 0: every table of the primary update moves down by
        // pskip; the tolerance tables (never touched by this kernel) get an offset that faults
        const int skip = tier.pskip;
        NxbModelLayout& m = p.ml;
        int* const moved[] = {&m.off_stats, &m.off_eperm, &m.off_parent, &m.off_tma, &m.off_len, &m.off_rate,
            &m.off_lam_pri, &m.off_p0_pri, &m.off_lam_blk, &m.off_p0_blk, &m.off_gsc_blk, &m.off_qval, &m.off_qinfo,
            &m.off_qptr, &m.off_pi, &m.off_clsmask, &m.off_ell_info, &m.off_ell_q, &m.off_qdt};
        for (int* q : moved) *q -= skip;
        m.bytes -= skip;
        m.off_slot = m.off_edgerec = m.off_cls = m.off_qm = m.off_ancmask = 0x7ffffff0;
        p.blob = h->d_blob + skip;
        p.smem_warp0 = tier.pwarp0;
    }
    p.pri_ok = h->d_pri_ok; p.tol_true_ok = h->d_tt; p.tol_false_ok = h->d_tf;
    p.state = h->d_state; p.state_stride = h->stride;
    p.gcapP = h->gcapP; p.gcapB = h->gcapB;
    p.so_tol = h->so_tol; p.so_pri = h->so_pri; p.so_ptm = h->so_ptm; p.so_pcode = h->so_pcode;
    p.so_btm = h->so_btm; p.so_bcode = h->so_bcode;
    p.n_units = h->n_units; p.unit_offset = h->unit_offset; p.chains = h->chains;
    p.unit_list = list; p.n_list = n_list;
    p.defer_list = out; p.defer_count = h->d_defer_count;
    p.seed = h->seed; p.target_sweeps = target; p.accum_from = accum_from;
    p.init = init;
    p.group_warps = group_warps;
    p.begin_sweeps = begin;
    p.sync_mask = h->sync_mask;
    p.validate = h->validate;
    p.validate_only = (kind == KIND_CHECK) ? 1 : 0;
    p.batch_units = tier.blink_units;
    p.blink_groups = tier.blink_groups;
    p.bulk_store = h->bulk_store;
    p.counts = h->d_counts; p.dwell = h->d_dwell; p.status = h->d_status;
    p.work_counter = h->d_work;
    p.perf = h->profile ? h->d_perf : nullptr;
    p.gscratch = h->d_gscratch; p.gscratch_stride = h->gscratch_stride;
    p.g_off_msg = h->g_off_msg; p.g_off_ctm = h->g_off_ctm; p.gcapF = h->gcapF; p.gcapC = h->gcapC;
    p.discard_pool = h->discard_pool;
    p.prefetch_wave = h->prefetch_wave;
    p.dbg_unit = -1; p.dbg_sweep = 0;
    if (const char* ev = getenv("NXB_DBG_UNIT")) { p.dbg_unit = atoll(ev); p.dbg_sweep = getenv("NXB_DBG_SWEEP") ? atoi(getenv("NXB_DBG_SWEEP")) : 0; }
}

// Enqueue the kernel of `kind` for `n_work` work items (grid sizing only) on the handle's stream.
static cudaError_t enqueue_kernel(nxb_handle* h, int kind, const Tier& tier, const NxbSweepParams& p, long long n_work) {
    const bool tolpar = kind == KIND_BLINK && h->tolpar && tier.tp_units > 0;
    if (tolpar) {
        const int grid = (int)std::max<long long>(1, std::min<long long>(h->num_sms, (n_work + tier.tp_units - 1) / tier.tp_units));
        const int threads = tier.tp_units * tier.tp_wpu * 32;
        if (h->msg_f64) {
            if (tier.tp_wpu == 2) tolpar_kernel<double, 2><<<grid, threads, tier.tp_smem, h->stream>>>(p);
            else tolpar_kernel<double, 1><<<grid, threads, tier.tp_smem, h->stream>>>(p);
        } else {
            if (tier.tp_wpu == 2) tolpar_kernel<float, 2><<<grid, threads, tier.tp_smem, h->stream>>>(p);
            else tolpar_kernel<float, 1><<<grid, threads, tier.tp_smem, h->stream>>>(p);
        }
    } else if (kind == KIND_BLINK) {
        int grid = (int)std::min<long long>((long long)h->num_sms * h->blink_ctas, (n_work + tier.blink_units - 1) / tier.blink_units);
        const bool grouped = tier.blink_groups > 1;          // groups of NXB_BLINK_GROUP_WARPS warps
        if (h->msg_f64) {
            if (grouped) blink_kernel<double, NXB_BLINK_GROUP_WARPS><<<std::max(grid, 1), tier.blink_warps * 32, tier.blink_smem, h->stream>>>(p);
            else blink_kernel<double, 0><<<std::max(grid, 1), tier.blink_warps * 32, tier.blink_smem, h->stream>>>(p);
        } else {
            if (grouped) blink_kernel<float, NXB_BLINK_GROUP_WARPS><<<std::max(grid, 1), tier.blink_warps * 32, tier.blink_smem, h->stream>>>(p);
            else blink_kernel<float, 0><<<std::max(grid, 1), tier.blink_warps * 32, tier.blink_smem, h->stream>>>(p);
        }
    } else {
        const int gw = std::max(p.group_warps, 1);
        const bool lean_layout = kind == KIND_PRIMARY_LEAN;
        const int launch_warps = (lean_layout ? tier.pwarps : tier.warps) / gw * gw;       // whole groups only
        const size_t smem = lean_layout ? tier.psmem : tier.smem;
        int grid = (int)std::min<long long>(h->num_sms, (n_work + launch_warps - 1) / launch_warps);
        grid = std::max(grid, 1);
        if (kind == KIND_FUSED) {
            if (h->msg_f64) sweep_kernel<double, false, false><<<grid, launch_warps * 32, tier.smem, h->stream>>>(p);
            else sweep_kernel<float, false, false><<<grid, launch_warps * 32, tier.smem, h->stream>>>(p);
        } else {
            const bool lean = !p.validate && !p.validate_only;
            if (h->msg_f64) {
                if (lean) sweep_kernel<double, true, true><<<grid, launch_warps * 32, smem, h->stream>>>(p);
                else sweep_kernel<double, true, false><<<grid, launch_warps * 32, smem, h->stream>>>(p);
            } else {
                if (lean) sweep_kernel<float, true, true><<<grid, launch_warps * 32, smem, h->stream>>>(p);
                else sweep_kernel<float, true, false><<<grid, launch_warps * 32, smem, h->stream>>>(p);
            }
        }
    }
    return cudaGetLastError();
}

// One kernel launch over `list` (or every unit): fills the defer list `out` (appending when
// `append`), adds the kernel time, returns the number of listed units through *n_deferred.
// `no_wait`: do not wait for the kernel; its time, status and listed units are collected by the
// next (waiting) launch on the same stream.
static int launch_one(nxb_handle* h, int kind, const Tier& tier, int group_warps, const int* list, int n_list,
                      int* out, bool append, unsigned begin, unsigned target, unsigned accum_from, int init,
                      int* n_deferred, bool no_wait = false) {
    NxbSweepParams p;
    fill_params(h, p, kind, tier, group_warps, list, n_list, out, begin, target, accum_from, init);
    CUDA_TRY(h, cudaMemsetAsync(h->d_work, 0, 8, h->stream));
    if (!append) CUDA_TRY(h, cudaMemsetAsync(h->d_defer_count, 0, 4, h->stream));
    const long long n_work = list ? n_list : h->n_units;
    CUDA_TRY(h, cudaEventRecord(no_wait ? h->ev2 : h->ev0, h->stream));
    CUDA_TRY(h, enqueue_kernel(h, kind, tier, p, n_work));
    CUDA_TRY(h, cudaEventRecord(no_wait ? h->ev3 : h->ev1, h->stream));
    if (no_wait) { h->pending_kind = kind; *n_deferred = 0; return NXB_OK; }
    int hs[5];
    CUDA_TRY(h, cudaMemcpyAsync(hs, h->d_status, 16, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(hs + 4, h->d_defer_count, 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->last_ms += ms; h->last_launches++; h->launches_total++;
    h->kind_ms[kind] += ms; h->kind_launches[kind]++;
    if (h->pending_kind >= 0) {
        float ms2 = 0.f;
        cudaEventElapsedTime(&ms2, h->ev2, h->ev3);
        h->last_ms += ms2; h->last_launches++; h->launches_total++;
        h->kind_ms[h->pending_kind] += ms2; h->kind_launches[h->pending_kind]++;
        h->pending_kind = -1;
    }
    if (hs[0] != 0) {
        char buf[256];
        snprintf(buf, sizeof(buf), "%s (code %d, unit %d, detail %d, sweep %d)", status_text(hs[0]),
                 hs[0], hs[1], hs[2], hs[3]);
        cudaMemsetAsync(h->d_status, 0, 16, h->stream);
        return set_err(h, hs[0], buf);
    }
    *n_deferred = hs[4];
    return NXB_OK;
}

// Run the units of `list` (all units when null) up to `target` sweeps with the fused kernel,
// escalating overflowing units through the tiers from `first_tier` on.
static int run_tiers(nxb_handle* h, unsigned target, unsigned accum_from, int init,
                     size_t first_tier = 0, const int* list = nullptr, int n_list = 0) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    const std::vector<Tier>& tiers = init ? h->init_tiers : h->tiers;
    for (size_t t = first_tier; t < tiers.size(); ++t) {
        const Tier& tier = tiers[t];
        int nd = 0;
        const int gw = (t == 0) ? group_warps_for(h, tier.warps) : 1;
        int rc = launch_one(h, KIND_FUSED, tier, gw, list, n_list, h->d_defer[t & 1], false, h->sweeps, target,
                            accum_from, (init && t == 0) ? 1 : 0, &nd);
        if (rc) return rc;
        if (nd == 0) return NXB_OK;
        h->deferred_total += nd;
        list = h->d_defer[t & 1];
        n_list = nd;
    }
    return set_err(h, NXB_ERR_SCRATCH_CAP, status_text(NXB_ERR_SCRATCH_CAP));
}

// Split mode: per sweep one primary-only launch and one blink_kernel launch over all units;
// whatever they could not take (events beyond the shared-memory capacities) is finished by
// the fused kernel of the next tiers before the following sweep starts.
static int run_split(nxb_handle* h, unsigned target, unsigned accum_from) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    const Tier& t0 = h->tiers[0];
    // the primary-only kernel keeps every warp busy in every sub-phase: larger groups
    const int gw = (h->lockstep < 0 && t0.warps >= 8) ? (t0.warps % 10 == 0 ? 10 : t0.warps / 2) : group_warps_for(h, t0.warps);
    while (h->sweeps < target) {
        int nd = 0;
        // (the tolerance launch below waits and collects the status and the listed units of both)
        int rc = launch_one(h, KIND_PRIMARY, t0, gw, nullptr, 0, h->d_defer[0], false, h->sweeps, h->sweeps + 1,
                            accum_from, 0, &nd, true);
        if (rc) return rc;
        h->reached = h->sweeps + 1;
        rc = launch_one(h, KIND_BLINK, t0, 1, nullptr, 0, h->d_defer[0], true, h->sweeps, h->sweeps + 1,
                        accum_from, 0, &nd);
        if (rc) return rc;
        if (nd > 0 && h->tolpar) {
            // Units that did not fit tier 0 of either kernel: the same two kernels with the
            // capacities of the next tiers (fewer units per CTA).  A unit deferred by the primary
            // kernel is still in phase PRIMARY and is skipped by the tolerance kernel of the tier,
            // and the other way round, so each tier takes every unit as far as its capacities allow.
            h->deferred_total += nd;
            const int* list = h->d_defer[0];
            int n_list = nd;
            size_t t = 1;
            for (; t < h->tiers.size() && n_list > 0; ++t) {
                const Tier& tt = h->tiers[t];
                if (tt.tp_units <= 0) break;
                int* out = h->d_defer[t & 1];
                int nd2 = 0;
                rc = launch_one(h, KIND_PRIMARY, tt, 1, list, n_list, out, false, h->sweeps, h->sweeps + 1,
                                accum_from, 0, &nd2, true);
                if (rc) return rc;
                rc = launch_one(h, KIND_BLINK, tt, 1, list, n_list, out, true, h->sweeps, h->sweeps + 1,
                                accum_from, 0, &nd2);
                if (rc) return rc;
                list = out; n_list = nd2;
            }
            if (n_list > 0) return set_err(h, NXB_ERR_SCRATCH_CAP, status_text(NXB_ERR_SCRATCH_CAP));
        } else if (nd > 0) {
            h->deferred_total += nd;
            rc = run_tiers(h, h->sweeps + 1, accum_from, 0, 1, h->d_defer[0], nd);
            if (rc) return rc;
        }
        h->sweeps++;
    }
    if (h->validate) {          // the last tolerance update has not been looked at yet
        int nd = 0;
        int rc = launch_one(h, KIND_CHECK, t0, gw, nullptr, 0, h->d_defer[0], false, h->sweeps, h->sweeps,
                            accum_from, 0, &nd);
        if (rc) return rc;
    }
    return NXB_OK;
}

// Split mode without a host round trip per sweep: the launches of up to NXB_ASYNC_CHUNK sweeps are
// enqueued back to back.  Per sweep: primary-only kernel, blink_kernel (both list the units that
// do not fit tier 0), then the fused kernel of the next (up to two) tiers over that list, whose
// length the kernels read on the device (NxbSweepParams::n_list_dev) -- an empty list costs an
// empty launch.  A unit that is still listed after those tiers stays behind, is listed again by
// the primary kernel of the following sweeps (it is off schedule) and, if it is still behind when
// the chunk ends, is finished by the host-driven escalation through the remaining tiers.
// Status, listed units and kernel times are collected once per chunk.
#define NXB_ASYNC_CHUNK 32
static int run_split_async(nxb_handle* h, unsigned target, unsigned accum_from) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    const Tier& t0 = h->tiers[0];
    // primary-only launches: the LEAN kernel with its own layout and warp count.  Lockstep groups,
    // measured at the p53 shape with 24 warps: 8 -> 8.98, 12 -> 9.11, 6 -> 9.95 ms per 4 sweeps of
    // 131 072 units (20 warps: 10 -> 10.15)
    const int pw = t0.pwarps;
    const int gw = (h->lockstep < 0 && pw >= 8) ? (pw % 8 == 0 ? 8 : (pw % 10 == 0 ? 10 : pw / 2)) : group_warps_for(h, pw);
    // (28 warps at 72 registers: 14 -> 8.84, 7 -> 9.38, 28 -> 9.47, 4 -> 11.9; 24 warps of the same build in groups of 8: 9.39)
    const int nblind = (int)std::min<size_t>(2, h->tiers.size() - 1);
    if (h->evpool.empty()) {
        h->evpool.resize(4 * NXB_ASYNC_CHUNK);
        for (auto& e : h->evpool) CUDA_TRY(h, cudaEventCreate(&e));
    }
    unsigned long long* const work = h->d_ctl;
    int* const cnt = (int*)(h->d_ctl + 4);
    int* const lists[2] = {h->d_defer[0], h->d_defer[1]};
    while (h->sweeps < target) {
        const int nsw = (int)std::min<unsigned>(NXB_ASYNC_CHUNK, target - h->sweeps);
        CUDA_TRY(h, cudaMemsetAsync(h->d_defer_total, 0, 8, h->stream));
        for (int i = 0; i < nsw; ++i) {
            const unsigned sw = h->sweeps + (unsigned)i;
            NxbSweepParams p;
            CUDA_TRY(h, cudaMemsetAsync(h->d_ctl, 0, 48, h->stream));
            CUDA_TRY(h, cudaEventRecord(h->evpool[4 * i + 0], h->stream));
            fill_params(h, p, KIND_PRIMARY_LEAN, t0, gw, nullptr, 0, lists[0], sw, sw + 1, accum_from, 0);
            p.work_counter = work + 0; p.defer_count = cnt + 0;
            CUDA_TRY(h, enqueue_kernel(h, KIND_PRIMARY_LEAN, t0, p, h->n_units));
            CUDA_TRY(h, cudaEventRecord(h->evpool[4 * i + 1], h->stream));
            fill_params(h, p, KIND_BLINK, t0, 1, nullptr, 0, lists[0], sw, sw + 1, accum_from, 0);
            p.work_counter = work + 1; p.defer_count = cnt + 0;
            CUDA_TRY(h, enqueue_kernel(h, KIND_BLINK, t0, p, h->n_units));
            CUDA_TRY(h, cudaEventRecord(h->evpool[4 * i + 2], h->stream));
            for (int b = 0; b < nblind; ++b) {
                const Tier& tt = h->tiers[1 + b];
                fill_params(h, p, KIND_FUSED, tt, 1, lists[b & 1], 0, lists[(b + 1) & 1], sw, sw + 1, accum_from, 0);
                p.n_list_dev = cnt + b; p.defer_count = cnt + b + 1; p.work_counter = work + 2 + b;
                if (b == 0) p.defer_total = h->d_defer_total;
                // (full grid: the list length is not known here; CTAs without work leave at once)
                CUDA_TRY(h, enqueue_kernel(h, KIND_FUSED, tt, p, h->n_units));
            }
            CUDA_TRY(h, cudaEventRecord(h->evpool[4 * i + 3], h->stream));
        }
        struct { int status[4]; int cnt[4]; unsigned long long total; } hs;
        CUDA_TRY(h, cudaMemcpyAsync(hs.status, h->d_status, 16, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(hs.cnt, cnt, 16, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(&hs.total, h->d_defer_total, 8, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (int i = 0; i < nsw; ++i) {
            float a = 0.f, b = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, h->evpool[4 * i + 0], h->evpool[4 * i + 1]);
            cudaEventElapsedTime(&b, h->evpool[4 * i + 1], h->evpool[4 * i + 2]);
            cudaEventElapsedTime(&c, h->evpool[4 * i + 2], h->evpool[4 * i + 3]);
            h->kind_ms[KIND_PRIMARY] += a; h->kind_ms[KIND_BLINK] += b; h->kind_ms[KIND_FUSED] += c;
            h->last_ms += a + b + c;
        }
        h->kind_launches[KIND_PRIMARY] += nsw; h->kind_launches[KIND_BLINK] += nsw; h->kind_launches[KIND_FUSED] += nsw * nblind;
        h->last_launches += nsw * (2 + nblind); h->launches_total += nsw * (2 + nblind);
        if (hs.status[0] != 0) {
            char buf[256];
            snprintf(buf, sizeof(buf), "%s (code %d, unit %d, detail %d, sweep %d)", status_text(hs.status[0]),
                     hs.status[0], hs.status[1], hs.status[2], hs.status[3]);
            cudaMemsetAsync(h->d_status, 0, 16, h->stream);
            h->reached = h->sweeps + (unsigned)nsw;        // where the units that did fit are now
            return set_err(h, hs.status[0], buf);
        }
        h->deferred_total += (long long)hs.total;
        h->sweeps += (unsigned)nsw;
        const int left = hs.cnt[nblind];            // still behind after the last sweep of the chunk
        if (left > 0) {
            h->deferred_total += left;
            int rc = run_tiers(h, h->sweeps, accum_from, 0, (size_t)nblind + 1, lists[nblind & 1], left);
            if (rc) return rc;
        }
    }
    return NXB_OK;
}

extern "C" int nxb_init_chains(nxb_handle* h, int32_t chains, uint64_t seed, int64_t unit_offset) {
    if (!h || chains <= 0) return set_err(h, NXB_ERR_BAD_ARG, "bad argument");
    if (!h->d_pri_ok) return set_err(h, NXB_ERR_BAD_ARG, "call nxb_set_data first");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const long long n_units = h->n_sites * chains;
    if (n_units > 0x7fffffffLL) return set_err(h, NXB_ERR_BAD_ARG, "too many units for one device");
    if (n_units != h->n_units) {
        cudaFree(h->d_state); cudaFree(h->d_defer[0]); cudaFree(h->d_defer[1]);
        h->d_state = nullptr; h->d_defer[0] = h->d_defer[1] = nullptr;
        CUDA_TRY(h, cudaMalloc(&h->d_state, (size_t)n_units * h->stride));
        CUDA_TRY(h, cudaMalloc(&h->d_defer[0], 4 * (size_t)n_units));
        CUDA_TRY(h, cudaMalloc(&h->d_defer[1], 4 * (size_t)n_units));
    }
    h->n_units = n_units; h->chains = chains; h->seed = seed; h->unit_offset = unit_offset;
    h->sweeps = 0;
    h->last_ms = 0.f; h->last_launches = 0;
    for (int i = 0; i < 4; ++i) { h->kind_ms[i] = 0.f; h->kind_launches[i] = 0; }
    int rc = run_tiers(h, 0u, 0xffffffffu, 1);
    for (int attempt = 0; rc == NXB_ERR_STATE_CAP && attempt < 6; ++attempt) {
        // an initial trajectory does not fit the slab: larger slabs, draw all of them again
        cudaFree(h->d_state); h->d_state = nullptr;
        h->cap_boost *= 2.0;
        h->grown_total++;
        rc = configure_capacities(h, h->capcfg);
        if (rc) return rc;
        CUDA_TRY(h, cudaMalloc(&h->d_state, (size_t)n_units * h->stride));
        rc = run_tiers(h, 0u, 0xffffffffu, 1);
    }
    return rc;
}

static int sweep_to(nxb_handle* h, unsigned target, unsigned accum_from) {
    if (h->sweeps >= target) return NXB_OK;
    if (h->split && !h->tolpar && !h->validate && !h->sync_sweeps && h->tiers[0].blink_units > 0) return run_split_async(h, target, accum_from);
    if (h->split && (h->tiers[0].blink_units > 0 || (h->tolpar && h->tiers[0].tp_units > 0))) return run_split(h, target, accum_from);
    h->reached = target;
    int rc = run_tiers(h, target, accum_from, 0);
    if (rc == NXB_OK) h->sweeps = target;
    return rc;
}

// A unit outgrew its HBM slab (NXB_ERR_STATE_CAP): double the HBM-side capacities, move the unit states to
// the larger slabs, and run the units that were left behind up to the sweep count the others have
// reached.  Their random streams depend only on (seed, unit, sweep), and the statistics that the
// failed attempt had already booked are skipped (NxbUnitHeader::pad), so the result is what slabs
// that were large enough from the start would have given.
static int recover_state_cap(nxb_handle* h, unsigned accum_from) {
    const unsigned reached = h->reached;
    h->cap_boost *= 2.0;
    int rc = configure_capacities(h, h->capcfg);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemsetAsync(h->d_defer_count, 0, 4, h->stream));
    collect_behind_kernel<<<(unsigned)((h->n_units + 255) / 256), 256, 0, h->stream>>>(
        h->d_state, h->stride, h->n_units, reached, h->d_defer[0], h->d_defer_count);
    int n = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&n, h->d_defer_count, 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->grown_total++;
    if (n > 0) {
        h->reached = reached;
        if (h->tiers.size() < 2) return set_err(h, NXB_ERR_SCRATCH_CAP, status_text(NXB_ERR_SCRATCH_CAP));
        rc = run_tiers(h, reached, accum_from, 0, 1, h->d_defer[0], n);
        if (rc) return rc;
    }
    h->sweeps = reached;
    return NXB_OK;
}

extern "C" int nxb_sweep(nxb_handle* h, int32_t n_sweeps, int32_t accumulate) {
    if (!h || n_sweeps < 0) return set_err(h, NXB_ERR_BAD_ARG, "bad argument");
    if (!h->d_state) return set_err(h, NXB_ERR_BAD_ARG, "call nxb_init_chains first");
    if (n_sweeps == 0) return NXB_OK;
    const unsigned target = h->sweeps + (unsigned)n_sweeps;
    const unsigned accum_from = accumulate ? h->sweeps : 0xffffffffu;
    h->last_ms = 0.f; h->last_launches = 0;
    for (int i = 0; i < 4; ++i) { h->kind_ms[i] = 0.f; h->kind_launches[i] = 0; }
    int rc = sweep_to(h, target, accum_from);
    for (int attempt = 0; rc == NXB_ERR_STATE_CAP && attempt < 6; ++attempt) {
        rc = recover_state_cap(h, accum_from);
        if (rc == NXB_OK) rc = sweep_to(h, target, accum_from);
    }
    return rc;
}

extern "C" int nxb_summary_layout_get(const nxb_handle* h, nxb_summary_layout* o) {
    if (!h || !o) return NXB_ERR_BAD_ARG;
    const NxbSummaryLayout& s = h->sl;
    o->n_counts = s.n_counts; o->n_dwell = s.n_dwell;
    o->c_root_pri = s.c_root_pri; o->c_root_on = s.c_root_on; o->c_root_off = s.c_root_off;
    o->c_root_on_cls = s.c_root_on_cls; o->c_pri_trans = s.c_pri_trans; o->c_off_on = s.c_off_on;
    o->c_on_off = s.c_on_off; o->c_nsamples = s.c_nsamples; o->d_T = s.d_T; o->d_off = s.d_off;
    return NXB_OK;
}

extern "C" int nxb_summary_reset(nxb_handle* h) {
    if (!h) return NXB_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemsetAsync(h->d_counts, 0, 8 * (size_t)h->sl.n_counts, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->d_dwell, 0, 8 * (size_t)h->sl.n_dwell, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NXB_OK;
}

extern "C" int nxb_summary_read(nxb_handle* h, uint64_t* counts, double* dwell) {
    if (!h || !counts || !dwell) return set_err(h, NXB_ERR_BAD_ARG, "bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(counts, h->d_counts, 8 * (size_t)h->sl.n_counts, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(dwell, h->d_dwell, 8 * (size_t)h->sl.n_dwell, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NXB_OK;
}

extern "C" int nxb_summary_device_ptrs(nxb_handle* h, void** c, void** d) {
    if (!h || !c || !d) return NXB_ERR_BAD_ARG;
    *c = h->d_counts; *d = h->d_dwell;
    return NXB_OK;
}

extern "C" int nxb_summarize_current(nxb_handle* h) {
    if (!h || !h->d_state) return set_err(h, NXB_ERR_BAD_ARG, "no chains");
    CUDA_TRY(h, cudaSetDevice(h->device));
    NxbSweepParams p;
    memset(&p, 0, sizeof(p));
    p.ml = h->ml; p.sl = h->sl; p.blob = h->d_blob;
    p.state = h->d_state; p.state_stride = h->stride;
    p.so_tol = h->so_tol; p.so_pri = h->so_pri; p.so_ptm = h->so_ptm; p.so_pcode = h->so_pcode;
    p.so_btm = h->so_btm; p.so_bcode = h->so_bcode;
    p.n_units = h->n_units; p.counts = h->d_counts; p.dwell = h->d_dwell;
    const int threads = 128;
    long long warps = h->n_units;
    int grid = (int)std::min<long long>((warps * 32 + threads - 1) / threads, 148LL * 16);
    summarize_kernel<<<grid, threads, h->ml.bytes, h->stream>>>(p);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NXB_OK;
}

extern "C" int nxb_export_unit(nxb_handle* h, int64_t unit, int32_t* node_pri, uint32_t* node_tol,
                               int32_t cap, int32_t* n_pri, double* pt, int32_t* pe, int32_t* psa,
                               int32_t* psb, int32_t* n_blk, double* bt, int32_t* be, int32_t* bk,
                               int32_t* bs) {
    if (!h || !h->d_state || unit < 0 || unit >= h->n_units) return set_err(h, NXB_ERR_BAD_ARG, "bad unit");
    CUDA_TRY(h, cudaSetDevice(h->device));
    std::vector<unsigned char> slab(h->stride);
    CUDA_TRY(h, cudaMemcpy(slab.data(), h->d_state + unit * h->stride, h->stride, cudaMemcpyDeviceToHost));
    const NxbUnitHeader* hdr = (const NxbUnitHeader*)slab.data();
    const int N = h->ml.N;
    const unsigned* tol = (const unsigned*)(slab.data() + h->so_tol);
    const unsigned char* pri = slab.data() + h->so_pri;
    for (int v = 0; v < N; ++v) { node_pri[v] = pri[v]; node_tol[v] = tol[v]; }
    *n_pri = hdr->n_pri; *n_blk = hdr->n_blk;
    if (hdr->n_pri > cap || hdr->n_blk > cap) return set_err(h, NXB_ERR_BAD_ARG, "export buffers too small");
    const double* g_ptm = (const double*)(slab.data() + h->so_ptm);
    const unsigned* g_pcode = (const unsigned*)(slab.data() + h->so_pcode);
    for (int i = 0; i < hdr->n_pri; ++i) {
        pt[i] = g_ptm[i]; pe[i] = g_pcode[i] & 0xffffu;
        psa[i] = (g_pcode[i] >> 16) & 0xffu; psb[i] = (g_pcode[i] >> 24) & 0xffu;
    }
    const double* g_btm = (const double*)(slab.data() + h->so_btm);
    const unsigned* g_bcode = (const unsigned*)(slab.data() + h->so_bcode);
    for (int i = 0; i < hdr->n_blk; ++i) {
        bt[i] = g_btm[i]; be[i] = g_bcode[i] & 0xffffu;
        bk[i] = (g_bcode[i] >> 16) & 0xffu; bs[i] = g_bcode[i] >> 31;
    }
    return NXB_OK;
}


extern "C" int nxb_forward_sample_rooted(nxb_handle* h, int64_t n_units, uint64_t seed, int64_t unit_offset,
                                         const int32_t* root_pri, const uint32_t* root_tol);

extern "C" int nxb_forward_sample(nxb_handle* h, int64_t n_units, uint64_t seed, int64_t unit_offset) {
    return nxb_forward_sample_rooted(h, n_units, seed, unit_offset, nullptr, nullptr);
}

extern "C" int nxb_forward_sample_rooted(nxb_handle* h, int64_t n_units, uint64_t seed, int64_t unit_offset,
                                         const int32_t* root_pri, const uint32_t* root_tol) {
    if (!h || n_units <= 0 || n_units > 0x7fffffffLL) return set_err(h, NXB_ERR_BAD_ARG, "bad argument");
    if ((root_pri == nullptr) != (root_tol == nullptr))
        return set_err(h, NXB_ERR_BAD_ARG, "give both root_pri and root_tol, or neither");
    if (root_pri)
        for (int64_t u = 0; u < n_units; ++u) {
            if (root_pri[u] < 0 || root_pri[u] >= h->ml.P)
                return set_err(h, NXB_ERR_BAD_ARG, "root primary state out of range");
            if (h->ml.K < 32 && (root_tol[u] >> h->ml.K))
                return set_err(h, NXB_ERR_BAD_ARG, "root tolerance mask has bits beyond n_classes");
        }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)n_units * h->ml.N;
    // unconstrained data: one "site" per unit, one chain (so that the units can be swept afterwards)
    {
        std::vector<uint64_t> pri(n, h->ml.P == 64 ? ~0ull : ((1ull << h->ml.P) - 1ull));
        std::vector<uint32_t> tt(n, h->ml.K == 32 ? ~0u : ((1u << h->ml.K) - 1u));
        int rc = nxb_set_data(h, n_units, pri.data(), tt.data(), tt.data());
        if (rc) return rc;
    }
    if (n_units != h->n_units) {
        cudaFree(h->d_state); cudaFree(h->d_defer[0]); cudaFree(h->d_defer[1]);
        h->d_state = nullptr; h->d_defer[0] = h->d_defer[1] = nullptr;
        CUDA_TRY(h, cudaMalloc(&h->d_state, (size_t)n_units * h->stride));
        CUDA_TRY(h, cudaMalloc(&h->d_defer[0], 4 * (size_t)n_units));
        CUDA_TRY(h, cudaMalloc(&h->d_defer[1], 4 * (size_t)n_units));
    }
    h->n_units = n_units; h->chains = 1; h->seed = seed; h->unit_offset = unit_offset; h->sweeps = 0;
    const int cap_events = h->gcapB;
    unsigned short* tmp_code = nullptr; double* tmp_time = nullptr;
    DevFree tmp;                        // temporaries are released on every exit path
    CUDA_TRY(h, cudaMalloc(&tmp_code, 2 * (size_t)n_units * cap_events));
    tmp.ptrs.push_back(tmp_code);
    CUDA_TRY(h, cudaMalloc(&tmp_time, 8 * (size_t)n_units * cap_events));
    tmp.ptrs.push_back(tmp_time);
    NxbSweepParams p;
    memset(&p, 0, sizeof(p));
    p.ml = h->ml; p.sl = h->sl; p.blob = h->d_blob;
    p.state = h->d_state; p.state_stride = h->stride;
    p.gcapP = h->gcapP; p.gcapB = h->gcapB;
    p.so_tol = h->so_tol; p.so_pri = h->so_pri; p.so_ptm = h->so_ptm; p.so_pcode = h->so_pcode;
    p.so_btm = h->so_btm; p.so_bcode = h->so_bcode;
    p.n_units = n_units; p.unit_offset = unit_offset; p.seed = seed; p.status = h->d_status;
    const int threads = 128;
    cudaFuncSetAttribute(forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->ml.bytes);
    int* d_rpri = nullptr; unsigned* d_rtol = nullptr;
    if (root_pri) {
        CUDA_TRY(h, cudaMalloc(&d_rpri, 4 * (size_t)n_units));
        tmp.ptrs.push_back(d_rpri);
        CUDA_TRY(h, cudaMalloc(&d_rtol, 4 * (size_t)n_units));
        tmp.ptrs.push_back(d_rtol);
        CUDA_TRY(h, cudaMemcpyAsync(d_rpri, root_pri, 4 * (size_t)n_units, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(d_rtol, root_tol, 4 * (size_t)n_units, cudaMemcpyHostToDevice, h->stream));
    }
    forward_kernel<<<(unsigned)((n_units + threads - 1) / threads), threads, h->ml.bytes, h->stream>>>(
        p, cap_events, tmp_code, tmp_time, d_rpri, d_rtol);
    CUDA_TRY(h, cudaGetLastError());
    int hs[4];
    CUDA_TRY(h, cudaMemcpyAsync(hs, h->d_status, 16, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (hs[0] != 0) {
        cudaMemsetAsync(h->d_status, 0, 16, h->stream);
        return set_err(h, hs[0], "forward sample outgrew the slab capacity; re-create the handle with a larger cap_scale");
    }
    return NXB_OK;
}

extern "C" int nxb_read_node_states(nxb_handle* h, int32_t* node_pri, uint32_t* node_tol) {
    if (!h || !h->d_state || !node_pri || !node_tol) return set_err(h, NXB_ERR_BAD_ARG, "bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_units * h->ml.N;
    int* d_pri = nullptr; unsigned* d_tol = nullptr;
    DevFree tmp;
    CUDA_TRY(h, cudaMalloc(&d_pri, 4 * n));
    tmp.ptrs.push_back(d_pri);
    CUDA_TRY(h, cudaMalloc(&d_tol, 4 * n));
    tmp.ptrs.push_back(d_tol);
    node_states_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
        h->d_state, h->stride, h->n_units, h->ml.N, h->so_tol, h->so_pri, d_pri, d_tol);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(node_pri, d_pri, 4 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(node_tol, d_tol, 4 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NXB_OK;
}

extern "C" int nxb_get_stats(nxb_handle* h, nxb_stats* o) {
    if (!h || !o) return NXB_ERR_BAD_ARG;
    memset(o, 0, sizeof(*o));
    CUDA_TRY(h, cudaSetDevice(h->device));
    o->n_units = h->n_units;
    if (h->d_state && h->n_units > 0) {
        CUDA_TRY(h, cudaMemsetAsync(h->d_scratch, 0, 16, h->stream));
        unit_stats_kernel<<<(unsigned)((h->n_units + 255) / 256), 256, 0, h->stream>>>(
            h->d_state, h->stride, h->n_units, h->d_scratch);
        unsigned long long r[2];
        CUDA_TRY(h, cudaMemcpyAsync(r, h->d_scratch, 16, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        o->n_pri_events = (int64_t)r[0]; o->n_blk_events = (int64_t)r[1];
        const int N = h->ml.N;
        o->state_bytes_used = h->n_units * (int64_t)(sizeof(NxbUnitHeader) + 4 * N + N) + 12 * (int64_t)(r[0] + r[1]);
    }
    o->data_bytes = h->n_sites * (int64_t)h->ml.N * 16;
    const Tier& t0 = h->tiers[0];
    // the kernel that a plain nxb_sweep launches for the primary update: the LEAN one with its own
    // layout on the asynchronous split path, else the tier-0 kernel
    const bool lean_path = h->split && !h->tolpar && !h->validate && !h->sync_sweeps && t0.blink_units > 0 && t0.pwarps > 0;
    o->warps_per_cta = lean_path ? t0.pwarps : t0.warps; o->n_ctas = h->num_sms;
    o->smem_bytes = (int32_t)(lean_path ? t0.psmem : t0.smem);
    o->n_tiers = (int32_t)h->tiers.size();
    o->capP = t0.wl.capP; o->capB = t0.wl.capB; o->capF = t0.wl.capF; o->capC = t0.wl.capC;
    o->deferred_units = h->deferred_total; o->kernel_launches = h->launches_total;
    o->capacity_doublings = h->grown_total;
    return NXB_OK;
}

extern "C" int nxb_last_kernel_ms_by_kind(nxb_handle* h, float ms[4], int32_t n[4]) {
    if (!h) return NXB_ERR_BAD_ARG;
    for (int i = 0; i < 4; ++i) {
        if (ms) ms[i] = h->kind_ms[i];
        if (n) n[i] = h->kind_launches[i];
    }
    return NXB_OK;
}

extern "C" int nxb_last_kernel_ms(nxb_handle* h, float* ms, int32_t* n) {
    if (!h) return NXB_ERR_BAD_ARG;
    if (ms) *ms = h->last_ms;
    if (n) *n = h->last_launches;
    return NXB_OK;
}

extern "C" int nxb_profile(nxb_handle* h, int32_t enable, uint64_t* cycles12) {
    if (!h) return NXB_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (cycles12) {
        CUDA_TRY(h, cudaMemcpy(cycles12, h->d_perf, 24 * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(h, cudaMemset(h->d_perf, 0, 32 * 8));
    }
    h->profile = enable;
    return NXB_OK;
}
