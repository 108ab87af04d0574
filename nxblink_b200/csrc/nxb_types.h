// Shared host/device types for the nxblink_b200 sweep engine.
//
// Vocabulary follows the reference (nxblink): a *unit* is one (site, chain);
// a *track* is the primary process or one tolerance ("blink") process; an
// *event* is a point on a tree edge where a track changes state; a *chunk* is
// a maximal region of the tree on which the foreground track cannot change
// (nxblink/chunking.py:1-12).
#pragma once
#include <stdint.h>
#if !defined(__CUDACC__) && !defined(__align__)
#define __align__(n) __attribute__((aligned(n)))   /* host build of the tolerance-update code (oracle/hostsim) */
#endif

#define NXB_MAX_STATES 64    // primary states P  (bitmask in one u64)
#define NXB_MAX_CLASSES 32   // tolerance classes K (bitmask in one u32; one lane per track)
#define NXB_PPAD 64          // padded row length of a chunk message

// unit status codes written to the per-handle status word
enum {
    NXB_OK = 0,
    NXB_ERR_INFEASIBLE = 1,      // FFBS met an all-zero message (raoteh.py:178-180,201)
    NXB_ERR_ZERO_RATE = 2,       // util.py:15-16 'zero max rate'
    NXB_ERR_STATE_CAP = 3,       // unit no longer fits its HBM slab
    NXB_ERR_INVARIANT = 4,       // navigation.py:40,122-142 style consistency check
    NXB_ERR_BAD_ARG = 5,
    NXB_ERR_CUDA = 6,
    NXB_ERR_SCRATCH_CAP = 7,     // scratch overflow even at the largest shared-memory tier
};

// where a deferred unit resumes (see DESIGN.md "capacity tiers")
enum { NXB_PHASE_PRIMARY = 0, NXB_PHASE_BLINK = 1, NXB_PHASE_INIT = 2 };

// Per-unit slab header in HBM (16 bytes).
struct NxbUnitHeader {
    uint32_t sweeps_done;   // absolute number of completed sweeps
    uint16_t n_pri;         // retained primary events
    uint16_t n_blk;         // retained blink events
    uint32_t phase;         // NXB_PHASE_*: PRIMARY = at a sweep boundary
    uint32_t pad;           // number of phases from this state on whose statistics are ALREADY booked (a unit that
                            // outgrew its slab is run again after the slabs have grown; see nxb_sweep); normally 0
};

// Offsets (bytes) of the model tables inside the model blob that every CTA
// stages into shared memory once.
struct NxbModelLayout {
    int N, E, P, K, nnz, nslots, diameter, table_bits;
    int off_stats;      // CTA-level statistic partials (u64), zeroed per launch
    int stats_words;
    int off_eperm;      // u16    [ceil32(E)] balanced lane assignment of the edges (0xffff: none)
    int off_parent;     // int32  [E]   parent node of node e+1
    int off_slot;       // int32  [N]   accumulator slot of internal nodes, -1 for leaves
    int off_tma;        // f64    [E]   time at the rootward end of edge e
    int off_len;        // f64    [E]   branch length (time units)
    int off_rate;       // f64    [E]   edge rate scaling factor
    int off_lam_pri;    // f64    [E]   dominating Poisson mean, primary track
    int off_p0_pri;     // f64    [E]   exp(-lam_pri)
    int off_lam_blk;    // f64    [E]   dominating Poisson mean, one blink track
    int off_p0_blk;     // f64    [E]   exp(-lam_blk)
    int off_gsc_blk;    // f32    [E]   ln2 / (dominating blink rate of the edge), 0 if inactive
    int off_edgerec;    // NxbEdgeRec [E]  the per-edge constants of the blink walks, one 32-byte record
    int off_qval;       // f64    [nnz] normalised primary rates (CSR values)
    int off_qinfo;      // u16    [nnz] column | class(column) << 8
    int off_qptr;       // u16    [P+1]
    int off_cls;        // u8     [P]
    int off_pi;         // f64    [P]
    int off_qm;         // f64    [P*K] dense Q_meta (model.py:93-106)
    int off_ancmask;    // u64    [N]   edges on the path root -> node (bit e = edge e; only used when E <= 64)
    int off_clsmask;    // u64    [K]   states belonging to class k
    int ell_w;          // padded row width of the ELL copy of Q (max out-degree)
    int off_ell_info;   // u16    [ell_w*64] column | class << 8, transposed: [w][row]
    int off_ell_q;      // f32 or f64 (message type) [ell_w*64] rates, 0 in the padding
    int off_qdt;        // message type [64*64] dense transposed rates: qdt[b*64 + a] = q[a -> b]
    int bytes;
    int blink_bytes;    // prefix of the blob that the tolerance update reads (edge records, slots, classes, Q_meta)
    double rate_on, rate_off;
    double inv_omega_pri;   // 1 / (2 * Rmax(all classes on))
    double acc_blk[4];      // thinning acceptance [own][fg state]
    double a_on, a_off;     // off-diagonal entries of the uniformized 2x2 blink matrix
    float gf_off, gf_on;    // 1 / acc_blk[0], 1 / acc_blk[1]: mean-gap factors of the candidate process (old state OFF / ON)
    float a_on_f, a_off_f, a_off1_f, pad_f;   // f32 copies of a_on, a_off, 1 - a_off (no conversions in the f32 walks)
    double p_on, p_off;     // blink prior at the root
    // event-parallel tolerance update (tol_par.cuh): the dominating Poisson process of all K
    // tracks laid out on one axis (track-major, edges in index order, rate-length units)
    int off_grec;           // NxbGenRec [E]   per-edge constants of the candidate generator
    int off_pcdf;           // u32 [pcdf_cap]  Poisson(tp_W) CDF thresholds (count = first j with word < pcdf[j]); n_pcdf used
    int n_pcdf, pcdf_cap;
    int off_etab;           // u16 [NXB_TP_NBIN]  first guess of the edge from the position on the track's axis
    double tp_inv_lam, tp_inv_bin;
    double tp_lam;          // length of one track on the axis: sum_e 2 max(on,off) rate_e len_e
    double tp_W;            // slice width = K * tp_lam / NXB_TP_NSLICE
    unsigned tp_thr[4];     // thinning thresholds [own][fg state] as u32 (accept iff word < thr)
    // level-parallel tree passes of tol_par.cuh: internal nodes by height (children before parents),
    // their children, message slots (a slot is live from a node's level to its parent's), and
    // edges by depth of the child node (parents before children)
    int n_levels, n_depths, nslots_lvl;
    int off_lvl_off;        // u16 [n_levels + 1]  offsets into lvl_node, heights 1 .. n_levels
    int off_lvl_node;       // u16 [internal nodes]
    int off_ch_off;         // u16 [N + 1]
    int off_ch_node;        // u16 [E]   children of every node (edge into child c is c - 1)
    int off_lslot;          // i16 [N]   message slot of an internal node in the level-ordered pass, -1 for leaves
    int off_dep_off;        // u16 [n_depths + 1] offsets into dep_edge, depths 1 .. n_depths
    int off_dep_edge;       // u16 [E]
};

#define NXB_TP_NSLICE 64
#define NXB_TP_NBIN 256
// Per-edge constants of the candidate generator of tol_par.cuh.
struct __align__(16) NxbGenRec {
    double cum;             // start of the edge on the track's axis (rate-length units)
    double inv_om;          // time per rate-length unit on this edge (0: inactive edge)
    double tma, tmb;
};

// Shared-memory layout of one unit in tolpar_kernel (bytes from the unit's base).
struct NxbTolLayout {
    int capP, capB;         // retained primary / blink events
    int capQ;               // candidates of one sweep (staging) = matrix slots
    int capL;               // live events (accepted candidates + retained)
    int off_ptm, off_pcode, off_btm, off_bcode;
    int off_node_tol, off_newtol, off_node_pri;
    int off_nrec;           // uint4[E]
    int off_ins;            // u16[capB]
    int off_toff;           // int[K + 2]
    int off_f0, off_f1, off_hasev;   // u32[E]
    int off_acc;            // AccRec<MSG>[nslots * K]
    int off_T, off_key, off_aux, off_q;
    int off_x;              // int[16] exchange slots / unit header
    int bytes;
};

// Per-edge constants read at every edge boundary of the tolerance walks.
struct __align__(16) NxbEdgeRec {
    double tma, tmb;        // times at the rootward / leafward end
    double rate;            // edge rate scaling factor
    float gsc;              // ln2 / dominating blink rate (0: no events possible on this edge)
    unsigned short va;      // parent node (the child node is e + 1)
    signed char sa_slot, sb_slot;   // accumulator slots of parent and child (-1: leaf)
};

// Per-warp shared-memory layout (bytes from the warp's base), chosen by the host
// for a capacity tier.
struct NxbWarpLayout {
    int capP, capB;         // retained primary / blink events held in shared memory
    int capR, capF;         // primary phase: pre-thinning region, foreground events
    int capC;               // blink phase: candidate pool
    // persistent
    int off_node_tol, off_node_pri, off_ptm, off_pcode, off_btm, off_bcode;
    int off_dpri, off_dtt, off_dtf;   // this unit's site constraints
    // small index arrays (live in both phases)
    int off_boff, off_poff, off_tmpa, off_tmpb, off_tmpc;
    // primary-phase scratch
    int off_bord, off_rtm, off_rmask, off_ftm, off_fmask, off_fedge,
        off_nchunk, off_jump, off_par, off_toland, off_callow, off_cstate, off_msg,
        off_finv, off_fu;
    // blink-phase scratch
    int off_toff, off_acc, off_newtol;
    int off_nrec;           // uint4[E]: per-edge record of the tolerance update's upward walk
    int off_uhdr;           // int[8 + 32]: unit header read by the lane jobs of other warps (persistent region)
    int bytes;
};

struct NxbSummaryLayout {
    // counts (u64)
    int c_root_pri;     // [P]
    int c_root_on;      // [1] total ON tolerance tracks at the root
    int c_root_off;     // [1]
    int c_root_on_cls;  // [K] ON tracks at the root, summed over samples whose root state is in class k
    int c_pri_trans;    // [E*nnz]
    int c_off_on;       // [E] off->on blink transitions
    int c_on_off;       // [E]
    int c_nsamples;     // [1]
    int n_counts;
    // dwell (f64)
    int d_T;            // [E*P]   time on edge e in primary state s
    int d_off;          // [E*P*K] time on edge e in state s with class k OFF
    int n_dwell;
};

struct NxbSweepParams {
    NxbModelLayout ml;
    NxbWarpLayout wl;
    NxbTolLayout tl;
    NxbSummaryLayout sl;
    int smem_warp0;                     // byte offset of warp 0's region in shared memory
    const unsigned char* blob;          // model blob (global)
    const double* rmax_tab;             // [2^K] max_s sum_{c in mask} Qmeta[s][c]
    // per-site data constraints
    const unsigned long long* pri_ok;   // [S][N] allowed primary states
    const unsigned* tol_true_ok;        // [S][N] bit k: class k may be ON
    const unsigned* tol_false_ok;       // [S][N] bit k: class k may be OFF
    // unit slabs
    unsigned char* state;
    long long state_stride;
    int gcapP, gcapB;                   // slab capacities
    int so_tol, so_pri, so_ptm, so_pcode, so_btm, so_bcode;   // byte offsets inside a slab
    long long n_units;                  // units on this device
    long long unit_offset;              // global id of local unit 0 (RNG key)
    int chains;                         // chains per site
    const int* unit_list;               // optional: deferred units to process
    int n_list;
    const int* n_list_dev;              // optional: length of unit_list, read on the device (written by the
                                        // kernels before it on the stream: no host round trip); 0 -> the launch does nothing
    unsigned long long* defer_total;    // optional: running total of listed units (block 0 adds *n_list_dev)
    int* defer_list;                    // out: units that overflowed this tier
    int* defer_count;
    unsigned long long seed;
    unsigned target_sweeps;             // run every unit up to this absolute sweep count
    unsigned accum_from;                // accumulate statistics for sweeps >= this
    int init;                           // 1: initialise trajectories (raoteh.py:337-361)
    int group_warps;                    // >1: warps run in phase lockstep in groups of this size
    unsigned sync_mask;                 // lockstep: which of the 6 primary sync points wait (the tolerance update's always do)
    unsigned begin_sweeps;              // lockstep: every unit starts the launch at this sweep count
    int validate;                       // 1: check trajectory invariants after every phase
    int validate_only;                  // primary-only kernel: load, check, do nothing else
    int batch_units;                    // blink_kernel: units per CTA batch
    int blink_groups;                   // blink_kernel: independent warp groups per CTA (each takes batch_units / groups units)
    int bulk_store;                     // blink_kernel: new blink lists leave shared memory as bulk-async copies (cp.async.bulk)
    unsigned long long* counts;
    double* dwell;
    int* status;                        // [4]: code, unit, detail, detail
    unsigned long long* work_counter;
    unsigned long long* perf;           // [12] optional per-phase cycle counters
    // per-warp-slot scratch in global memory (L2-resident): chunk messages of the primary
    // FFBS and the candidate pool of the tolerance update
    unsigned char* gscratch;
    long long gscratch_stride;
    int g_off_msg, g_off_ctm;          // global scratch: chunk messages, event pool
    int gcapF, gcapC;
    int prefetch_wave;                  // 1: the persistent kernels prefetch the unit slabs one wave of work ahead into L2
    int discard_pool;                   // 1: dead pool records are discarded from L2 instead of written back (NXB_DISCARD_POOL=0 switches it off)
    long long dbg_unit;                 // development: dump the event list of this unit (tolpar_kernel), -1: none
    int dbg_sweep;
};
