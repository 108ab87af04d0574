// M-STEP ON THE DEVICE.  The objective of examples/p53/p53em.py:58-91 --
// minus (em.get_ll_root + em.get_ll_dwell + em.get_ll_trans), nxblink/em.py:19-198, over the
// log-parameters (kappa, omega, A, C, G, T, blink on, blink off, one rate per edge;
// p53em.py:143-185) -- is evaluated together with its closed-form gradient straight from the
// device-resident sufficient statistics (the flat buffers of nxb_summary_layout, or their
// all-reduced copies), and minimised by an L-BFGS iteration that runs entirely inside ONE
// kernel launch: one CTA, the 8 + E parameters and the L-BFGS history in shared memory, the
// [E x nnz] dwell matrix read from L2.  The reference does this on the host with algopy
// forward-mode AD and scipy L-BFGS-B (p53em.py:94-140); nothing here leaves the GPU until the
// optimum is copied out.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/nxblink_b200.h"
#include "nxb_types.h"

namespace nxbm {

#define MS_THREADS 512
#define MS_MAXN 320          /* parameters: 8 + edges */
#define MS_HIST 10

struct MsArgs {
    int E, P, K, nnz, n;                 // n = 8 + E
    // model structure (device)
    const int* q_ptr; const int* q_col; const int* cls; const int* nt_to;
    const unsigned char* is_ts; const unsigned char* is_non;
    const double* nt_counts; const double* xon_w;
    // statistics (device)
    const unsigned long long* counts; const double* dwell;
    nxb_summary_layout L;
    // scratch (device): pri_dwell [E*nnz], ctrans [nnz], offx [E], xoff [E], tt [E]
    double* pri_dwell; double* ctrans; double* offx; double* xoff; double* tt;
    // in/out
    double* x;                           // [n] log-parameters
    double* out;                         // [0] f, [1] iterations, [2] evaluations, [3] status, [4] max |g|, [8 ..] gradient
    int maxiter; int eval_only;
    double gtol, ftol;
};

struct MsShared {
    double red[4][MS_THREADS / 32];
    double bc[8];
    double x[MS_MAXN], g[MS_MAXN], xn[MS_MAXN], gn[MS_MAXN], d[MS_MAXN];
    double S[MS_HIST][MS_MAXN], Y[MS_HIST][MS_MAXN];
    double De[MS_MAXN], rate[MS_MAXN];
    double scal[16];                     // root_off, root_xon, n, OffOn, OnOff, TT
};

// sum of up to 4 values over the CTA; every thread gets the totals
__device__ void block_sum4(MsShared& sh, double v[4], int nv) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < nv; ++k) {
        double a = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) sh.red[k][w] = a;
    }
    __syncthreads();
    for (int k = 0; k < nv; ++k) {
        double a = (lane < MS_THREADS / 32) ? sh.red[k][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        v[k] = a;
    }
    __syncthreads();
}

// n*LL and its gradient in the log-parameters at xs[]; dynamic shared arrays q/G [nnz],
// distn/rowsum/dld [P].  Returns f = -LL/n through sh.bc[0]; gradient of f in gout[].
// PROFILE: the edge rates are not read from xs but set to their closed-form stationary point
// given the other parameters, rate_e = ER * transitions_e / D_e (1e-12 for an edge without
// transitions, which the reference sets to zero: p53em.py:25-27); by the envelope theorem the
// gradient in the 8 global parameters is the same expression evaluated at those rates.
template <bool PROFILE>
__device__ void ms_eval(const MsArgs& a, MsShared& sh, const double* xs, double* gout, double* dyn) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, NW = MS_THREADS / 32;
    const int E = a.E, P = a.P, nnz = a.nnz;
    double* q = dyn;                 // [nnz]
    double* G = q + nnz;             // [nnz]
    double* distn = G + nnz;         // [P]
    double* rowsum = distn + P;      // [P]
    double* dld = rowsum + P;        // [P]
    const int* rowof = reinterpret_cast<const int*>(dld + P + 8);      // [nnz] row of entry i (set once by the kernel)
    const double kappa = exp(xs[0]), omega = exp(xs[1]);
    double nt[4];
    {
        double t = 0.0;
        for (int j = 0; j < 4; ++j) { nt[j] = exp(xs[2 + j]); t += nt[j]; }
        for (int j = 0; j < 4; ++j) nt[j] /= t;
    }
    const double on = exp(xs[6]), off = exp(xs[7]);
    const double p_off = off / (on + off), p_on = on / (on + off);
    for (int i = tid; i < nnz; i += MS_THREADS)
        q[i] = nt[a.nt_to[i]] * (a.is_ts[i] ? kappa : 1.0) * (a.is_non[i] ? omega : 1.0);
    if (!PROFILE) for (int e = tid; e < E; e += MS_THREADS) sh.rate[e] = exp(xs[8 + e]);
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int s = tid; s < P; s += MS_THREADS) {
        double wgt = 1.0;
        for (int j = 0; j < 4; ++j) wgt *= pow(nt[j], a.nt_counts[s * 4 + j]);
        distn[s] = wgt;
        v[0] += wgt;
    }
    __syncthreads();
    for (int s = tid; s < P; s += MS_THREADS) {
        double r = 0.0;
        for (int i = a.q_ptr[s]; i < a.q_ptr[s + 1]; ++i) r += q[i];
        rowsum[s] = r;
    }
    block_sum4(sh, v, 1);
    const double Z = v[0];
    v[0] = 0.0; v[1] = 0.0;
    for (int s = tid; s < P; s += MS_THREADS) {
        const double ds = distn[s] / Z;
        distn[s] = ds;
        v[0] += rowsum[s] * ds;                                                  // ER
        v[1] += (double)a.counts[a.L.c_root_pri + s] * log(ds);                  // root_pri . log distn
    }
    // D_e: one warp per edge
    for (int e = w; e < E; e += NW) {
        const double* row = a.pri_dwell + (size_t)e * nnz;
        double acc = 0.0;
        for (int i = lane; i < nnz; i += 32) acc += row[i] * q[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) sh.De[e] = acc + a.offx[e] * on + a.xoff[e] * off;
    }
    block_sum4(sh, v, 2);
    const double ER = v[0], ll_root_pri = v[1];
    if (PROFILE) {
        for (int e = tid; e < E; e += MS_THREADS)
            sh.rate[e] = (a.tt[e] > 0.0 && sh.De[e] > 0.0) ? ER * a.tt[e] / sh.De[e] : 1e-12;
        __syncthreads();
    }
    v[0] = 0.0; v[1] = 0.0; v[2] = 0.0; v[3] = 0.0;
    for (int e = tid; e < E; e += MS_THREADS) {
        const double r = sh.rate[e];
        v[0] += sh.De[e] * r;                                   // A
        if (a.tt[e] > 0.0) v[1] += a.tt[e] * (log(r) - log(ER));   // sum_e TT_e log(rate_e / ER)
        v[2] += r * a.offx[e];
        v[3] += r * a.xoff[e];
    }
    block_sum4(sh, v, 4);
    const double A = v[0], ll_rates = v[1], r_offx = v[2], r_xoff = v[3];
    // G_i = -sum_e rate_e dwell_ei q_i / ER + ctrans_i + dER * distn[row_i] * q_i ; sum_i ctrans_i log q_i
    const double TT = sh.scal[5];
    const double dER = A / (ER * ER) - TT / ER;
    v[0] = 0.0;
    for (int i = tid; i < nnz; i += MS_THREADS) {            // one column of the dwell matrix per thread (coalesced)
        double col = 0.0;
        for (int e = 0; e < E; ++e) col += sh.rate[e] * a.pri_dwell[(size_t)e * nnz + i];
        G[i] = -col * q[i] / ER + a.ctrans[i] + dER * distn[rowof[i]] * q[i];
        if (a.ctrans[i] != 0.0) v[0] += a.ctrans[i] * log(q[i]);
    }
    for (int s = tid; s < P; s += MS_THREADS)
        dld[s] = (double)a.counts[a.L.c_root_pri + s] + dER * rowsum[s] * distn[s];
    block_sum4(sh, v, 1);
    const double ll_trans_q = v[0];
    const double root_off = sh.scal[0], root_xon = sh.scal[1], nsamp = sh.scal[2], OffOn = sh.scal[3], OnOff = sh.scal[4];
    const double ll = ll_root_pri + root_off * log(p_off) + root_xon * log(p_on) - A / ER + ll_trans_q +
                      OffOn * log(on) + OnOff * log(off) + ll_rates;
    // gradient pieces that need sums over i / s
    double h[4] = {0.0, 0.0, 0.0, 0.0};
    double gk = 0.0, go = 0.0;
    for (int i = tid; i < nnz; i += MS_THREADS) {
        const double Gi = G[i];
        if (a.is_ts[i]) gk += Gi;
        if (a.is_non[i]) go += Gi;
        h[a.nt_to[i]] += Gi;
    }
    double cb[4] = {0.0, 0.0, 0.0, 0.0}, dsum = 0.0, dc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int s = tid; s < P; s += MS_THREADS) {
        dsum += dld[s];
        for (int j = 0; j < 4; ++j) { cb[j] += distn[s] * a.nt_counts[s * 4 + j]; dc[j] += dld[s] * a.nt_counts[s * 4 + j]; }
    }
    v[0] = gk; v[1] = go; v[2] = dsum; v[3] = 0.0;
    block_sum4(sh, v, 3);
    gk = v[0]; go = v[1]; dsum = v[2];
    block_sum4(sh, h, 4);
    block_sum4(sh, cb, 4);
    block_sum4(sh, dc, 4);
    if (tid == 0) {
        double hh[4], hs = 0.0;
        for (int j = 0; j < 4; ++j) { hh[j] = h[j] + dc[j] - cb[j] * dsum; hs += hh[j]; }
        gout[0] = -gk / nsamp;
        gout[1] = -go / nsamp;
        for (int j = 0; j < 4; ++j) gout[2 + j] = -(hh[j] - nt[j] * hs) / nsamp;
        gout[6] = -(-root_off * p_on + root_xon * p_off - r_offx * on / ER + OffOn) / nsamp;
        gout[7] = -(root_off * p_on - root_xon * p_off - r_xoff * off / ER + OnOff) / nsamp;
        sh.bc[0] = -ll / nsamp;
    }
    if (!PROFILE)
        for (int e = tid; e < E; e += MS_THREADS)
            gout[8 + e] = -(-sh.rate[e] * sh.De[e] / ER + a.tt[e]) / nsamp;
    __syncthreads();
}

__device__ double ms_dot(MsShared& sh, const double* u, const double* w, int n) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += MS_THREADS) v[0] += u[i] * w[i];
    block_sum4(sh, v, 1);
    return v[0];
}

__global__ void __launch_bounds__(MS_THREADS, 1) mstep_kernel(const MsArgs a) {
    extern __shared__ __align__(16) unsigned char ms_dyn[];
    MsShared& sh = *reinterpret_cast<MsShared*>(ms_dyn);
    double* dyn = reinterpret_cast<double*>(ms_dyn + ((sizeof(MsShared) + 15) / 16) * 16);
    const int tid = threadIdx.x, n = a.n, E = a.E, P = a.P, K = a.K, nnz = a.nnz;
    // ---- statistics -> the reduced forms the objective needs (summary.py:111-131 unfolding) ----
    const double* T = a.dwell + a.L.d_T;
    const double* Doff = a.dwell + a.L.d_off;
    int* rowof = reinterpret_cast<int*>(dyn + 2 * nnz + 3 * P + 8);
    for (int s = tid; s < P; s += MS_THREADS)
        for (int i = a.q_ptr[s]; i < a.q_ptr[s + 1]; ++i) {
            rowof[i] = s;
            const int c = a.cls[a.q_col[i]];
            double ct = 0.0;
            for (int e = 0; e < E; ++e) {
                const double t = T[e * P + s];
                double dw = t - Doff[((size_t)e * P + s) * K + c];
                // (T and Doff are sums of the same pieces in different orders: rounding residue)
                if (!(dw > 1e-11 * t)) dw = 0.0;
                a.pri_dwell[(size_t)e * nnz + i] = dw;
                ct += (double)a.counts[a.L.c_pri_trans + (size_t)e * nnz + i];
            }
            a.ctrans[i] = ct;
        }
    for (int e = tid; e < E; e += MS_THREADS) {
        double so = 0.0, st = 0.0;
        for (int j = 0; j < P * K; ++j) so += Doff[(size_t)e * P * K + j];
        for (int s = 0; s < P; ++s) st += T[e * P + s];
        a.offx[e] = so;
        a.xoff[e] = (double)(K - 1) * st - so;
        double tt = (double)a.counts[a.L.c_off_on + e] + (double)a.counts[a.L.c_on_off + e];
        for (int i = 0; i < nnz; ++i) tt += (double)a.counts[a.L.c_pri_trans + (size_t)e * nnz + i];
        a.tt[e] = tt;
    }
    __syncthreads();
    {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        for (int e = tid; e < E; e += MS_THREADS) {
            v[0] += (double)a.counts[a.L.c_off_on + e];
            v[1] += (double)a.counts[a.L.c_on_off + e];
            v[2] += a.tt[e];
        }
        for (int k = tid; k < K; k += MS_THREADS) v[3] += a.xon_w[k] * (double)a.counts[a.L.c_root_on_cls + k];
        block_sum4(sh, v, 4);
        if (tid == 0) {
            sh.scal[0] = (double)a.counts[a.L.c_root_off];
            sh.scal[1] = v[3];
            sh.scal[2] = (double)a.counts[a.L.c_nsamples];
            sh.scal[3] = v[0]; sh.scal[4] = v[1]; sh.scal[5] = v[2];
        }
    }
    for (int i = tid; i < n; i += MS_THREADS) sh.x[i] = a.x[i];
    __syncthreads();
    if (!(sh.scal[2] > 0.0)) {              // no sample accumulated: nothing to maximise
        if (tid == 0) { a.out[0] = 0.0; a.out[1] = 0.0; a.out[2] = 0.0; a.out[3] = -1.0; a.out[4] = 0.0; }
        return;
    }
    int evals = 1, iters = 0, status = 0, hist = 0, head = 0;
    double f;
    if (a.eval_only) {
        ms_eval<false>(a, sh, sh.x, sh.g, dyn);
        f = sh.bc[0];
    } else {
        // ---- L-BFGS over the 8 global parameters of the profile objective (two-loop recursion,
        // weak-Wolfe line search by bisection / doubling) ----
        const int m = 8;
        for (int i = tid; i < n; i += MS_THREADS) { sh.g[i] = 0.0; sh.gn[i] = 0.0; }
        __syncthreads();
        ms_eval<true>(a, sh, sh.x, sh.g, dyn);
        f = sh.bc[0];
        double rho[MS_HIST], alpha[MS_HIST];         // (every thread holds the same values)
        for (int j = 0; j < MS_HIST; ++j) { rho[j] = 0.0; alpha[j] = 0.0; }
        for (; iters < a.maxiter; ++iters) {
            double gmax = 0.0;
            for (int i = 0; i < m; ++i) gmax = fmax(gmax, fabs(sh.g[i]));
            if (gmax <= a.gtol) { status = 1; break; }
            // direction d = -H g
            for (int i = tid; i < m; i += MS_THREADS) sh.d[i] = sh.g[i];
            __syncthreads();
            for (int j = 0; j < hist; ++j) {
                const int k = (head - 1 - j + MS_HIST) % MS_HIST;
                const double al = rho[k] * ms_dot(sh, sh.S[k], sh.d, m);
                alpha[k] = al;
                for (int i = tid; i < m; i += MS_THREADS) sh.d[i] -= al * sh.Y[k][i];
                __syncthreads();
            }
            if (hist > 0) {
                const int k = (head - 1 + MS_HIST) % MS_HIST;
                const double yy = ms_dot(sh, sh.Y[k], sh.Y[k], m);
                const double gam = 1.0 / (rho[k] * yy);
                for (int i = tid; i < m; i += MS_THREADS) sh.d[i] *= gam;
                __syncthreads();
            }
            for (int j = hist - 1; j >= 0; --j) {
                const int k = (head - 1 - j + MS_HIST) % MS_HIST;
                const double be = rho[k] * ms_dot(sh, sh.Y[k], sh.d, m);
                const double al = alpha[k];
                for (int i = tid; i < m; i += MS_THREADS) sh.d[i] += (al - be) * sh.S[k][i];
                __syncthreads();
            }
            for (int i = tid; i < m; i += MS_THREADS) sh.d[i] = -sh.d[i];
            __syncthreads();
            double gd = ms_dot(sh, sh.g, sh.d, m);
            if (!(gd < 0.0)) {                  // not a descent direction: restart from steepest descent
                hist = 0;
                for (int i = tid; i < m; i += MS_THREADS) sh.d[i] = -sh.g[i];
                __syncthreads();
                gd = ms_dot(sh, sh.g, sh.d, m);
            }
            double step = 1.0;
            if (hist == 0) { const double gn2 = sqrt(ms_dot(sh, sh.g, sh.g, m)); step = fmin(1.0, 1.0 / gn2); }
            // weak Wolfe: f(x + t d) <= f + c1 t g.d  and  g(x + t d).d >= c2 g.d
            double lo = 0.0, hi = INFINITY, fn = f;
            bool ok = false;
            for (int ls = 0; ls < 40; ++ls) {
                for (int i = tid; i < m; i += MS_THREADS) sh.xn[i] = sh.x[i] + step * sh.d[i];
                __syncthreads();
                ms_eval<true>(a, sh, sh.xn, sh.gn, dyn);
                ++evals;
                fn = sh.bc[0];
                if (!(isfinite(fn) && fn <= f + 1e-4 * step * gd)) { hi = step; step = 0.5 * (lo + hi); continue; }
                const double gdn = ms_dot(sh, sh.gn, sh.d, m);
                if (gdn < 0.9 * gd) { lo = step; step = isinf(hi) ? 2.0 * step : 0.5 * (lo + hi); continue; }
                ok = true;
                break;
            }
            if (!ok && lo > 0.0) {              // no Wolfe point found: the largest step with sufficient decrease
                for (int i = tid; i < m; i += MS_THREADS) sh.xn[i] = sh.x[i] + lo * sh.d[i];
                __syncthreads();
                ms_eval<true>(a, sh, sh.xn, sh.gn, dyn);
                ++evals;
                fn = sh.bc[0];
                ok = true;
            }
            if (!ok) { status = 3; break; }
            const int k = head;
            for (int i = tid; i < m; i += MS_THREADS) { sh.S[k][i] = sh.xn[i] - sh.x[i]; sh.Y[k][i] = sh.gn[i] - sh.g[i]; }
            __syncthreads();
            const double sy = ms_dot(sh, sh.S[k], sh.Y[k], m);
            const double yy = ms_dot(sh, sh.Y[k], sh.Y[k], m);
            if (sy > 1e-10 * yy && sy > 0.0) {
                rho[k] = 1.0 / sy;
                head = (head + 1) % MS_HIST;
                if (hist < MS_HIST) ++hist;
            }
            const double fold = f;
            for (int i = tid; i < m; i += MS_THREADS) { sh.x[i] = sh.xn[i]; sh.g[i] = sh.gn[i]; }
            f = fn;
            __syncthreads();
            if ((fold - f) <= a.ftol * fmax(fmax(fabs(fold), fabs(f)), 1.0)) { ++iters; status = 2; break; }
        }
        if (status == 0) status = 4;            // iteration limit
        // the accepted point once more: its rates (the last evaluation may have been a rejected trial)
        ms_eval<true>(a, sh, sh.x, sh.g, dyn);
        ++evals;
        f = sh.bc[0];
        for (int e = tid; e < E; e += MS_THREADS) { sh.x[8 + e] = log(sh.rate[e]); sh.g[8 + e] = 0.0; }
        __syncthreads();
    }
    double gmax = 0.0;
    for (int i = 0; i < n; ++i) gmax = fmax(gmax, fabs(sh.g[i]));
    for (int i = tid; i < n; i += MS_THREADS) { a.x[i] = sh.x[i]; a.out[8 + i] = sh.g[i]; }
    if (tid == 0) { a.out[0] = f; a.out[1] = (double)iters; a.out[2] = (double)evals; a.out[3] = (double)status; a.out[4] = gmax; }
}

}  // namespace nxbm

static std::string g_ms_err;
extern "C" const char* nxb_mstep_last_error(void) { return g_ms_err.c_str(); }

#define MTRY(call)                                                                   \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            g_ms_err = std::string(#call) + ": " + cudaGetErrorString(e_);           \
            return NXB_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

namespace {
struct DevBufs {
    std::vector<void*> p;
    ~DevBufs() { for (void* q : p) cudaFree(q); }
    template <typename T> cudaError_t alloc(T** out, size_t bytes) {
        cudaError_t e = cudaMalloc((void**)out, bytes ? bytes : 16);
        if (e == cudaSuccess) p.push_back(*out);
        return e;
    }
};
}  // namespace

static float g_ms_kernel_ms = 0.f;
extern "C" float nxb_mstep_last_kernel_ms(void) { return g_ms_kernel_ms; }

extern "C" int nxb_mstep(int device, const nxb_mstep_desc* d, const nxb_summary_layout* L,
                         const void* counts_dev, const void* dwell_dev, double* x, int32_t maxiter,
                         int32_t eval_only, double* f_out, double* grad_out, int32_t* info) {
    if (!d || !L || !counts_dev || !dwell_dev || !x || !f_out) { g_ms_err = "null argument"; return NXB_ERR_BAD_ARG; }
    const int E = d->n_edges, P = d->n_states, K = d->n_classes, nnz = d->nnz, n = 8 + E;
    if (E < 1 || P < 1 || K < 1 || nnz < 1 || n > MS_MAXN) { g_ms_err = "bad sizes (need 8 + n_edges <= 320)"; return NXB_ERR_BAD_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { g_ms_err = "no CUDA device available (there is no CPU fallback)"; return NXB_ERR_CUDA; }
    MTRY(cudaSetDevice(device));
    DevBufs b;
    nxbm::MsArgs a;
    memset(&a, 0, sizeof(a));
    a.E = E; a.P = P; a.K = K; a.nnz = nnz; a.n = n;
    // one arena, one upload: [structure tables | x] then the kernel's scratch and outputs
    auto up8 = [](size_t v) { return (v + 7) / 8 * 8; };
    size_t o = 0;
    const size_t o_qptr = o; o = up8(o + 4 * (size_t)(P + 1));
    const size_t o_qcol = o; o = up8(o + 4 * (size_t)nnz);
    const size_t o_cls = o; o = up8(o + 4 * (size_t)P);
    const size_t o_ntto = o; o = up8(o + 4 * (size_t)nnz);
    const size_t o_ts = o; o = up8(o + (size_t)nnz);
    const size_t o_non = o; o = up8(o + (size_t)nnz);
    const size_t o_ntc = o; o += 8 * (size_t)P * 4;
    const size_t o_xw = o; o += 8 * (size_t)K;
    const size_t o_x = o; o += 8 * (size_t)n;
    const size_t n_up = o;
    const size_t o_out = o; o += 8 * (size_t)(8 + n);
    const size_t o_pd = o; o += 8 * (size_t)E * nnz;
    const size_t o_ct = o; o += 8 * (size_t)nnz;
    const size_t o_offx = o; o += 8 * (size_t)E;
    const size_t o_xoff = o; o += 8 * (size_t)E;
    const size_t o_tt = o; o += 8 * (size_t)E;
    unsigned char* arena = nullptr;
    MTRY(b.alloc(&arena, o));
    std::vector<unsigned char> hostbuf(n_up, 0);
    memcpy(hostbuf.data() + o_qptr, d->q_ptr, 4 * (size_t)(P + 1));
    memcpy(hostbuf.data() + o_qcol, d->q_col, 4 * (size_t)nnz);
    memcpy(hostbuf.data() + o_cls, d->state_class, 4 * (size_t)P);
    memcpy(hostbuf.data() + o_ntto, d->nt_to, 4 * (size_t)nnz);
    memcpy(hostbuf.data() + o_ts, d->is_ts, (size_t)nnz);
    memcpy(hostbuf.data() + o_non, d->is_non, (size_t)nnz);
    memcpy(hostbuf.data() + o_ntc, d->nt_counts, 8 * (size_t)P * 4);
    memcpy(hostbuf.data() + o_xw, d->xon_weight, 8 * (size_t)K);
    memcpy(hostbuf.data() + o_x, x, 8 * (size_t)n);
    MTRY(cudaMemcpy(arena, hostbuf.data(), n_up, cudaMemcpyHostToDevice));
    const int* q_ptr = (const int*)(arena + o_qptr); const int* q_col = (const int*)(arena + o_qcol);
    const int* cls = (const int*)(arena + o_cls); const int* nt_to = (const int*)(arena + o_ntto);
    const unsigned char* is_ts = arena + o_ts; const unsigned char* is_non = arena + o_non;
    const double* nt_counts = (const double*)(arena + o_ntc); const double* xon_w = (const double*)(arena + o_xw);
    double* dx = (double*)(arena + o_x); double* dout = (double*)(arena + o_out);
    a.pri_dwell = (double*)(arena + o_pd); a.ctrans = (double*)(arena + o_ct);
    a.offx = (double*)(arena + o_offx); a.xoff = (double*)(arena + o_xoff); a.tt = (double*)(arena + o_tt);
    a.q_ptr = q_ptr; a.q_col = q_col; a.cls = cls; a.nt_to = nt_to; a.is_ts = is_ts; a.is_non = is_non;
    a.nt_counts = nt_counts; a.xon_w = xon_w;
    a.counts = (const unsigned long long*)counts_dev; a.dwell = (const double*)dwell_dev; a.L = *L;
    a.x = dx; a.out = dout; a.maxiter = maxiter; a.eval_only = eval_only;
    a.gtol = 1e-5; a.ftol = 2.220446049250313e-9;          // scipy L-BFGS-B defaults (pgtol, factr * eps)
    const size_t smem = ((sizeof(nxbm::MsShared) + 15) / 16) * 16 + 8 * (2 * (size_t)nnz + 3 * (size_t)P + 8) + 4 * (size_t)nnz + 16;
    int max_smem = 0;
    MTRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (smem > (size_t)max_smem) { g_ms_err = "model too large for the on-device M-step"; return NXB_ERR_BAD_ARG; }
    MTRY(cudaFuncSetAttribute(nxbm::mstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    MTRY(cudaEventCreate(&e0)); MTRY(cudaEventCreate(&e1));
    cudaEventRecord(e0, 0);
    nxbm::mstep_kernel<<<1, MS_THREADS, smem>>>(a);
    cudaEventRecord(e1, 0);
    cudaError_t le = cudaGetLastError();
    if (le == cudaSuccess) le = cudaEventSynchronize(e1);
    if (le == cudaSuccess) cudaEventElapsedTime(&g_ms_kernel_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    MTRY(le);
    std::vector<double> out(8 + 2 * n);           // x and the outputs are adjacent in the arena
    MTRY(cudaMemcpy(out.data(), dx, 8 * (size_t)(8 + 2 * n), cudaMemcpyDeviceToHost));
    memcpy(x, out.data(), 8 * (size_t)n);
    out.erase(out.begin(), out.begin() + n);
    if (out[3] < 0.0) { g_ms_err = "the statistics hold no samples (accumulate sweeps before the M-step)"; return NXB_ERR_BAD_ARG; }
    *f_out = out[0];
    if (grad_out) for (int i = 0; i < n; ++i) grad_out[i] = out[8 + i];
    if (info) { info[0] = (int)out[1]; info[1] = (int)out[2]; info[2] = (int)out[3]; }
    return NXB_OK;
}
