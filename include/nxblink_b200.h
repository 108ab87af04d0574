/* nxblink_b200 -- C-ABI of the B200 Rao-Teh blink-model sweep engine.
 *
 * The reference (argriffing/nxblink) is pure Python and has no FFI; its boundary
 * for this path is the Python call surface listed in SURVEY.md section 8(b).  The
 * entry points below are what a ctypes binding of that surface needs; each one
 * cites the reference code whose work it replaces (paths relative to the reference
 * root).  All pointers are plain host pointers unless the name says `dev`; the
 * caller owns every buffer; every function returns 0 on success or an NXB_ERR_*
 * code, with a message available from nxb_last_error().  A handle is bound to one
 * CUDA device and is not re-entrant.  There is no CPU fallback: every call fails
 * with NXB_ERR_CUDA if no device is usable.
 */
#ifndef NXBLINK_B200_H
#define NXBLINK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nxb_handle nxb_handle;

/* Integer-indexed model tables.
 * Replaces the `model` duck type read by nxblink/raoteh.py:59-119
 * (get_T_and_root, get_edge_to_blen, get_edge_to_rate, get_primary_to_tol,
 * get_Q_primary, get_primary_distn, get_rate_on, get_rate_off).
 * Nodes are numbered so that parent[v] < v and node 0 is the root; edge e is the
 * edge into node e+1.  q_* is the primary rate matrix in CSR form ALREADY divided
 * by its expected rate (raoteh.py:90-102). */
typedef struct {
    int32_t n_nodes;
    int32_t n_states;        /* P <= 64 */
    int32_t n_classes;       /* K <= 32 */
    int32_t nnz;
    const int32_t* parent;   /* [n_nodes], parent[0] = -1 */
    const double* blen;      /* [n_nodes], branch length of the edge into v (blen[0] unused) */
    const double* rate;      /* [n_nodes], rate scaling factor of the edge into v */
    const int32_t* q_ptr;    /* [n_states + 1] */
    const int32_t* q_col;    /* [nnz] */
    const double* q_val;     /* [nnz] */
    const int32_t* state_class; /* [n_states], tolerance class of each primary state */
    const double* prior;     /* [n_states], root prior of the primary process */
    double rate_on;          /* blink off -> on */
    double rate_off;         /* blink on -> off */
    int32_t diameter;        /* directed diameter of Q (raoteh.py:351) */
    int32_t msg_f64;         /* 1: chunk messages stored as f64, 0: f32 */
    double cap_scale;        /* scales every per-unit capacity; 0 or 1.0 = default */
    int32_t validate;        /* 1: check trajectory invariants after every phase */
    int32_t max_warps;       /* 0 = auto */
    int32_t lockstep_warps;  /* warps per phase-lockstep group (instruction-cache sharing);
                                0 = automatic, negative = independent warps */
    int32_t split_phases;    /* 0 = default: primary and tolerance updates run as two alternating
                                kernels; negative = one fused kernel.  Results are identical. */
} nxb_model_desc;

/* Sizes of the flat statistic buffers (see nxb_summary_read). */
typedef struct {
    int32_t n_counts, n_dwell;
    int32_t c_root_pri, c_root_on, c_root_off, c_root_on_cls, c_pri_trans,
            c_off_on, c_on_off, c_nsamples;
    int32_t d_T, d_off;
} nxb_summary_layout;

typedef struct {
    int64_t n_units;
    int64_t state_bytes_used;    /* sum over units of the live slab bytes */
    int64_t n_pri_events;        /* retained primary events, all units */
    int64_t n_blk_events;        /* retained blink events, all units */
    int64_t data_bytes;          /* per-site constraint bytes held on the device */
    int32_t warps_per_cta, n_ctas, smem_bytes, n_tiers;   /* warps / shared memory of the primary-update kernel that nxb_sweep launches */
    int32_t capP, capB, capF, capC;
    int64_t deferred_units;      /* units that needed a higher capacity tier so far */
    int64_t kernel_launches;     /* sweep-kernel launches so far */
    int64_t capacity_doublings;  /* times a unit outgrew its HBM slab and the slab capacities were doubled (see nxb_sweep) */
} nxb_stats;

int nxb_device_count(void);

/* Build device tables for one model on `device`. */
int nxb_create(const nxb_model_desc* desc, int device, nxb_handle** out);
int nxb_destroy(nxb_handle* h);
/* Replace the VALUES of the model (rates, branch lengths, prior, blink rates) by those of a
 * model of the same shape (tree, sparsity pattern of Q, classes): what an EM iteration changes
 * (examples/p53/run-stochastic.py:154-208 builds a new model object per iteration).  Device
 * buffers, capacity tiers, data and unit states are kept; chains may be continued (warm start)
 * or initialised again with nxb_init_chains.  If an edge switched between zero and positive
 * rate x length the unit states are dropped and nxb_init_chains must be called.  Capacities
 * (shared-memory tiers, slab sizes: 4x the expected event counts) follow the model: when the new
 * rates mean more events per unit than the handle was sized for, they are derived again and the
 * unit states are moved to the larger slabs.  The update is atomic: a refused model (e.g. edge
 * rate x length too large) leaves the handle exactly as it was. */
int nxb_update_model(nxb_handle* h, const nxb_model_desc* desc);
const char* nxb_last_error(const nxb_handle* h);   /* h may be NULL: last create error */

/* Per-site data constraints; replaces the `data` duck type of raoteh.py:45-52,61
 * after the zero-length pooling of raoteh.py:160-184 (done by the caller).
 * pri_ok[s*N+v]: bit i set iff primary state i is allowed at node v.
 * tol_true_ok / tol_false_ok: bit k set iff class k may be ON / OFF at node v.
 * Host buffers; copied to the device. */
int nxb_set_data(nxb_handle* h, int64_t n_sites, const uint64_t* pri_ok,
                 const uint32_t* tol_true_ok, const uint32_t* tol_false_ok);

/* Allocate `chains` independent chains per site and draw the initial feasible
 * trajectories (raoteh.py:337-361 init_tracks).  `unit_offset` is the global index
 * of this device's first unit: random streams are keyed by global unit index. */
int nxb_init_chains(nxb_handle* h, int32_t chains, uint64_t seed, int64_t unit_offset);

/* n_sweeps full Gibbs sweeps (raoteh.py:392-435) of every unit.  If `accumulate`
 * the sufficient statistics of every sweep are added to the device summary
 * (nxblink/summary.py:190-243 Summary.on_sample).  The launches of up to 32 sweeps are enqueued
 * back to back; status and timings are collected once per such chunk.  A unit whose trajectory
 * outgrows its HBM slab does not end the call: the HBM-side capacities are doubled, the states move to
 * the larger slabs and the unit is run again from where it stood, its already booked statistics
 * skipped -- results are identical to those of a handle that was large enough from the start. */
int nxb_sweep(nxb_handle* h, int32_t n_sweeps, int32_t accumulate);

int nxb_summary_layout_get(const nxb_handle* h, nxb_summary_layout* out);
int nxb_summary_reset(nxb_handle* h);
/* Copies the statistics to host buffers: counts[n_counts] (u64), dwell[n_dwell] (f64). */
int nxb_summary_read(nxb_handle* h, uint64_t* counts, double* dwell);
/* Device pointers of the same two buffers, for an in-place NCCL all-reduce. */
int nxb_summary_device_ptrs(nxb_handle* h, void** counts_dev, void** dwell_dev);
/* Recompute the statistics of the CURRENT trajectories with an independent
 * kernel (not fused with the sweep) and add them to the device summary. */
int nxb_summarize_current(nxb_handle* h);

/* Export one unit's trajectory (the tracks yielded by raoteh.py:435).
 * node_pri[N], node_tol[N]; events: up to cap entries each, returns counts.
 * Primary: time, edge, sa, sb.  Blink: time, edge, track, new state. */
int nxb_export_unit(nxb_handle* h, int64_t unit, int32_t* node_pri, uint32_t* node_tol,
                    int32_t cap, int32_t* n_pri, double* pri_time, int32_t* pri_edge,
                    int32_t* pri_sa, int32_t* pri_sb,
                    int32_t* n_blk, double* blk_time, int32_t* blk_edge,
                    int32_t* blk_track, int32_t* blk_state);

/* Forward (prior) simulation of n_units independent trajectories from the model, root to
 * leaves (nxblink/fwdsample.py:67-262 gen_forward_samples), written into the unit slabs: they can
 * then be summarised (nxb_summarize_current), exported (nxb_export_unit, nxb_read_node_states:
 * synthetic alignments) or swept.  Replaces any data previously set with unconstrained data. */
int nxb_forward_sample(nxb_handle* h, int64_t n_units, uint64_t seed, int64_t unit_offset);
/* Same, with the root state of every trajectory given by the caller, as
 * fwdsample.gen_forward_samples(model, seq_of_track_to_root_state) takes it (fwdsample.py:67,145-160):
 * root_pri[n_units] state indices, root_tol[n_units] bit masks of the classes that are ON
 * (the class of the root's primary state is forced ON).  Both NULL: draw roots from the prior. */
int nxb_forward_sample_rooted(nxb_handle* h, int64_t n_units, uint64_t seed, int64_t unit_offset,
                              const int32_t* root_pri, const uint32_t* root_tol);
/* Node states of every unit: node_pri[n_units*N], node_tol[n_units*N] (bit k: class k on). */
int nxb_read_node_states(nxb_handle* h, int32_t* node_pri, uint32_t* node_tol);

/* Run every later launch and copy of this handle on the caller's CUDA stream (a cudaStream_t
 * passed as an integer, e.g. torch.cuda.current_stream().cuda_stream) instead of the handle's
 * own stream, so that the caller's events and collectives order with the sweeps without extra
 * synchronisation.  0 restores the handle's own stream.  The handle never destroys a caller's
 * stream.  (The reference has no streams; this is what a torch-side caller of the E-step needs,
 * examples/p53/run-stochastic.py:172-195.) */
int nxb_set_stream(nxb_handle* h, uint64_t cuda_stream);

int nxb_get_stats(nxb_handle* h, nxb_stats* out);
/* Last sweep-kernel launch times (ms, CUDA events on the launching stream). */
int nxb_last_kernel_ms(nxb_handle* h, float* total_ms, int32_t* n_launches);
/* The same, per kernel: [0] fused sweep_kernel (deferral tiers / split_phases < 0),
 * [1] primary-only sweep_kernel, [2] blink_kernel, [3] validation-only pass. */
int nxb_last_kernel_ms_by_kind(nxb_handle* h, float ms[4], int32_t n_launches[4]);

/* Per-phase cycle counters of the sweep kernel (lane 0 of every warp, summed):
 * load, P0..P5 (primary update), B0..B2 + write-back (tolerance updates), store.
 * `enable` switches collection on/off; if cycles12 is non-NULL the 24 counters (12 of the sweep kernels, 12 of tolpar_kernel) are
 * copied out and cleared.  Diagnostic only. */
int nxb_profile(nxb_handle* h, int32_t enable, uint64_t* cycles12);

/* ---- deterministic dense path (FP64 tensor pipe) --------------------------------------
 * P_b = expm(Q * t_b), b < batch, for one n x n rate matrix (row-major, n <= 64).
 * Replaces scipy.linalg.expm per edge, examples/codon2x3/toymodel-analytic.py:112-121.
 * Host buffers: Q[n*n], t[batch], out[batch*n*n]. */
int nxb_expm_batched(int device, int32_t n, int32_t batch, const double* Q, const double* t,
                     double* out);
/* expm(Q t_b) and its Frechet derivative in direction E_b,
 *   L_b = int_0^1 exp(Q t_b s) E_b exp(Q t_b (1 - s)) ds,
 * for b < batch.  Replaces scipy.linalg.expm_frechet of nxblink/denseutil.py:148-176
 * (compute_edge_expectation / compute_dwell_times) and toymodel-analytic.py:159-226.
 * Host buffers: Q[n*n], t[batch], E[batch*n*n], outP[batch*n*n] (may be NULL), outL[batch*n*n]. */
int nxb_expm_frechet_batched(int device, int32_t n, int32_t batch, const double* Q, const double* t,
                             const double* E, double* outP, double* outL);
/* Pruning log-likelihood per site.  Replaces nxmctree.dynamic_fset_lhood.get_lhood,
 * toymodel-analytic.py:123-126.  parent[v] < v, root 0; P[e] is the n x n transition
 * matrix of the edge into node e+1; masks[s*n_nodes+v] bit i: state i feasible at node v
 * of site s.  Host buffers. */
int nxb_pruning_loglik(int device, int32_t n, int32_t n_nodes, const int32_t* parent,
                       const double* P, const double* prior, int64_t n_sites,
                       const uint64_t* masks, double* loglik);
const char* nxb_dense_last_error(void);
/* Duration (ms, CUDA events around the launch) of the kernel of the last dense call above. */
float nxb_dense_last_kernel_ms(void);

/* ---- M-step on the device ----------------------------------------------------
 * Structure of the MG94 codon parametrisation of examples/p53/p53em.py:188-229 (get_pre_Q): for
 * entry i of the CSR pattern (q_ptr, q_col) of the primary rate matrix,
 *   q_i = nt[nt_to[i]] * (is_ts[i] ? kappa : 1) * (is_non[i] ? omega : 1),
 * and the codon distribution is prod_j nt_j^nt_counts[s][j], normalised (p53em.py:232-257).
 * xon_weight[k] is 1 if an ON track of class k at the root counts towards `root_xon_count`
 * (summary.py:221), else 0. */
typedef struct {
    int32_t n_edges, n_states, n_classes, nnz;
    const int32_t* q_ptr;        /* [n_states + 1] */
    const int32_t* q_col;        /* [nnz] */
    const int32_t* state_class;  /* [n_states] */
    const int32_t* nt_to;        /* [nnz] target nucleotide 0..3 */
    const uint8_t* is_ts;        /* [nnz] transition? */
    const uint8_t* is_non;       /* [nnz] nonsynonymous? */
    const double* nt_counts;     /* [n_states * 4] */
    const double* xon_weight;    /* [n_classes] */
} nxb_mstep_desc;
/* Objective f(x) = -(em.get_ll_root + em.get_ll_dwell + em.get_ll_trans) / nsamples
 * (nxblink/em.py:19-198 as assembled by examples/p53/p53em.py:58-91) and its gradient in the
 * log-parameters x = log(kappa, omega, A, C, G, T, blink on, blink off, edge rates[n_edges])
 * (p53em.py:143-185), evaluated on the DEVICE-resident statistics (counts_dev / dwell_dev: the
 * buffers of nxb_summary_device_ptrs or all-reduced copies of them, laid out as `layout`).
 * eval_only != 0: evaluate at x (f_out, grad_out[8 + n_edges]).  Otherwise minimise from x by
 * L-BFGS inside one kernel launch (replaces scipy.optimize.minimize(..., 'L-BFGS-B') with algopy
 * gradients, p53em.py:94-140) and return the optimum in x, f_out, grad_out (may be NULL);
 * info[0] iterations, info[1] evaluations, info[2] 1: gradient below 1e-5, 2: relative decrease
 * below 2.2e-9 (scipy's defaults), 3: line search failed, 4: iteration limit.  Host x / outputs. */
int nxb_mstep(int device, const nxb_mstep_desc* d, const nxb_summary_layout* layout,
              const void* counts_dev, const void* dwell_dev, double* x, int32_t maxiter,
              int32_t eval_only, double* f_out, double* grad_out, int32_t* info);
const char* nxb_mstep_last_error(void);
float nxb_mstep_last_kernel_ms(void);

/* Known-answer hook for the counter-based generator (host code path). */
void nxb_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
