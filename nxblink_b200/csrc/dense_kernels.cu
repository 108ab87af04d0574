// Deterministic dense kernels: batched matrix exponentials and the pruning
// (Felsenstein) likelihood over alignment sites, both FP64 on the tensor pipe
// (DMMA, mma.sync.m8n8k4.f64).
//
// Replaces, on the device, the dense cross-check path of the reference:
//   examples/codon2x3/toymodel-analytic.py:112-126   scipy.linalg.expm per edge +
//                                                    nxmctree.dynamic_fset_lhood.get_lhood
//   nxblink/denseutil.py:116-146                     dense rate / transition matrices
// and serves the per-iteration likelihood of the EM / ML loop (SURVEY.md 8d config 5:
// one 61x61 expm per edge, pruning over all sites: per edge a [n x n] . [n x sites]
// contraction).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/nxblink_b200.h"
#include "nxb_types.h"

namespace nxbd {

#define ND 64            // padded matrix order (n <= 64: feasibility masks are one u64)
#define ST 32            // sites per CTA tile in the pruning kernel

// D[8x8] += A[8x4] * B[4x8] on the FP64 tensor pipe.
// Fragments (lane T): a = A[T/4][T%4], b = B[T%4][T/4], c0/c1 = C[T/4][2*(T%4) + {0,1}].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C = A * B for ND x ND matrices in shared memory (row-major, leading dimension ND).
// 8 warps: warp w owns the 8-row strip w; C must not alias A or B.  The A fragments of the
// strip (16 k-steps) stay in registers across the 8 column tiles: one shared-memory load per
// DMMA instead of two.
__device__ void matmul64(const double* A, const double* B, double* C, int warp, int lane) {
    const int r = lane >> 2, q = lane & 3;
    double a[ND / 4];
#pragma unroll
    for (int ks = 0; ks < ND / 4; ++ks) a[ks] = A[(warp * 8 + r) * ND + ks * 4 + q];
    for (int jt = 0; jt < ND / 8; ++jt) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < ND / 4; ++ks)
            dmma(c0, c1, a[ks], B[(ks * 4 + q) * ND + jt * 8 + r]);
        C[(warp * 8 + r) * ND + jt * 8 + 2 * q] = c0;
        C[(warp * 8 + r) * ND + jt * 8 + 2 * q + 1] = c1;
    }
}

// One CTA (256 threads) per matrix: P_b = expm(Q * t_b) by scaling and squaring with a
// degree-13 Taylor polynomial evaluated by Horner's rule (|| Q t / 2^s ||_1 <= 1/2:
// truncation error 0.5^14 / 14! = 7e-16).
__global__ void __launch_bounds__(256) expm_kernel(int n, const double* __restrict__ Q,
                                                    const double* __restrict__ t, double* __restrict__ out) {
    extern __shared__ __align__(16) double sm[];
    double* A = sm;                 // Q t / 2^s
    double* X = sm + ND * ND;
    double* Y = sm + 2 * ND * ND;
    __shared__ double s_norm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double tb = t[blockIdx.x];
    if (tid == 0) s_norm = 0.0;
    __syncthreads();
    // 1-norm bound: max column sum of |Q t|
    for (int j = tid; j < n; j += blockDim.x) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += fabs(Q[i * n + j]);
        // non-negative doubles order like their bit patterns
        atomicMax((unsigned long long*)&s_norm, (unsigned long long)__double_as_longlong(acc * fabs(tb)));
    }
    __syncthreads();
    int s = 0;
    {
        double nrm = s_norm;
        while (nrm > 0.5 && s < 60) { nrm *= 0.5; ++s; }
    }
    const double scale = ldexp(tb, -s);
    for (int i = tid; i < ND * ND; i += blockDim.x) {
        const int r = i / ND, c = i % ND;
        A[i] = (r < n && c < n) ? Q[r * n + c] * scale : 0.0;
        X[i] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int k = 13; k >= 1; --k) {
        matmul64(A, X, Y, warp, lane);          // Y = A X
        __syncthreads();
        const double inv = 1.0 / (double)k;
        for (int i = tid; i < ND * ND; i += blockDim.x) {
            const int r = i / ND, c = i % ND;
            X[i] = Y[i] * inv + ((r == c) ? 1.0 : 0.0);
        }
        __syncthreads();
    }
    for (int q = 0; q < s; ++q) {
        matmul64(X, X, Y, warp, lane);
        __syncthreads();
        for (int i = tid; i < ND * ND; i += blockDim.x) X[i] = Y[i];
        __syncthreads();
    }
    double* o = out + (size_t)blockIdx.x * n * n;
    for (int i = tid; i < n * n; i += blockDim.x) o[i] = X[(i / n) * ND + (i % n)];
}

// C (+)= A * B, as matmul64; with ACC the product is added to C.
template <bool ACC>
__device__ void matmul64x(const double* A, const double* B, double* C, int warp, int lane) {
    const int r = lane >> 2, q = lane & 3;
    double a[ND / 4];
#pragma unroll
    for (int ks = 0; ks < ND / 4; ++ks) a[ks] = A[(warp * 8 + r) * ND + ks * 4 + q];
    for (int jt = 0; jt < ND / 8; ++jt) {
        double* cp = C + (warp * 8 + r) * ND + jt * 8 + 2 * q;
        double c0 = ACC ? cp[0] : 0.0, c1 = ACC ? cp[1] : 0.0;
#pragma unroll
        for (int ks = 0; ks < ND / 4; ++ks)
            dmma(c0, c1, a[ks], B[(ks * 4 + q) * ND + jt * 8 + r]);
        cp[0] = c0; cp[1] = c1;
    }
}

// One CTA per (matrix, direction): P_b = expm(A) and the Frechet derivative
//     L_b = int_0^1 exp(A s) E_b exp(A (1 - s)) ds,      A = Q t_b,
// i.e. the top-right block of expm([[A, E], [0, A]]), computed on the blocks: Horner's rule
// for the degree-13 Taylor polynomial of the scaled pair (X <- I + A X / k,
// Xl <- (A Xl + E X) / k), then s squarings (Xl <- X Xl + Xl X, X <- X X).
// Replaces scipy.linalg.expm_frechet of nxblink/denseutil.py:148-176.
__global__ void __launch_bounds__(256) frechet_kernel(int n, const double* __restrict__ Q,
        const double* __restrict__ t, const double* __restrict__ E,
        double* __restrict__ outP, double* __restrict__ outL) {
    extern __shared__ __align__(16) double sm[];
    double* A = sm;                       // Q t / 2^s
    double* D = sm + ND * ND;             // E / 2^s
    double* X = sm + 2 * ND * ND;
    double* Xl = sm + 3 * ND * ND;
    double* Y = sm + 4 * ND * ND;
    double* Yl = sm + 5 * ND * ND;
    __shared__ double s_norm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double tb = t[blockIdx.x];
    const double* Eb = E + (size_t)blockIdx.x * n * n;
    if (tid == 0) s_norm = 0.0;
    __syncthreads();
    for (int j = tid; j < n; j += blockDim.x) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += fabs(Q[i * n + j]);
        atomicMax((unsigned long long*)&s_norm, (unsigned long long)__double_as_longlong(acc * fabs(tb)));
    }
    __syncthreads();
    int s = 0;
    {
        double nrm = s_norm;
        while (nrm > 0.5 && s < 60) { nrm *= 0.5; ++s; }
    }
    const double scale = ldexp(tb, -s), escale = ldexp(1.0, -s);
    for (int i = tid; i < ND * ND; i += blockDim.x) {
        const int r = i / ND, c = i % ND;
        const bool in = r < n && c < n;
        A[i] = in ? Q[r * n + c] * scale : 0.0;
        D[i] = in ? Eb[r * n + c] * escale : 0.0;
        X[i] = (r == c) ? 1.0 : 0.0;
        Xl[i] = 0.0;
    }
    __syncthreads();
    for (int k = 13; k >= 1; --k) {
        matmul64x<false>(A, X, Y, warp, lane);
        matmul64x<false>(A, Xl, Yl, warp, lane);
        matmul64x<true>(D, X, Yl, warp, lane);        // each warp re-reads only its own rows of Yl
        __syncthreads();
        const double inv = 1.0 / (double)k;
        for (int i = tid; i < ND * ND; i += blockDim.x) {
            const int r = i / ND, c = i % ND;
            X[i] = Y[i] * inv + ((r == c) ? 1.0 : 0.0);
            Xl[i] = Yl[i] * inv;
        }
        __syncthreads();
    }
    for (int q = 0; q < s; ++q) {
        matmul64x<false>(X, Xl, Yl, warp, lane);
        matmul64x<true>(Xl, X, Yl, warp, lane);
        matmul64x<false>(X, X, Y, warp, lane);
        __syncthreads();
        for (int i = tid; i < ND * ND; i += blockDim.x) { X[i] = Y[i]; Xl[i] = Yl[i]; }
        __syncthreads();
    }
    if (outP) {
        double* o = outP + (size_t)blockIdx.x * n * n;
        for (int i = tid; i < n * n; i += blockDim.x) o[i] = X[(i / n) * ND + (i % n)];
    }
    double* o = outL + (size_t)blockIdx.x * n * n;
    for (int i = tid; i < n * n; i += blockDim.x) o[i] = Xl[(i / n) * ND + (i % n)];
}

// Pruning likelihood.  One CTA (256 threads) per tile of ST sites; the tree recursion runs
// inside the CTA with node messages [ND states x ST sites] in reusable shared-memory slots
// (edges in descending index: children before parents; parent[v] < v).  Per edge the
// contraction M = P_e . L_child is done on the FP64 tensor pipe.  Messages are rescaled per
// site after every node and the logarithms of the scales are accumulated.
__global__ void __launch_bounds__(256) pruning_kernel(
        int n, int n_nodes, int nslots, const int* __restrict__ parent, const int* __restrict__ slot,
        const int* __restrict__ last_child,   // 1 if edge e is the last one folded into its parent
        const double* __restrict__ P,         // [E][n][n]
        const double* __restrict__ prior, long long n_sites,
        const unsigned long long* __restrict__ masks,     // [S][N]
        double* __restrict__ loglik) {
    extern __shared__ __align__(16) double sm[];
    double* Pe = sm;                            // ND x ND
    double* M = sm + ND * ND;                   // ND x ST
    double* slots = M + ND * ST;                // nslots x ND x ST
    __shared__ double s_logscale[ST];
    __shared__ double s_colmax[ST];
    double* s_part = M;                         // M is free until the contraction below
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long site0 = (long long)blockIdx.x * ST;
    const int E = n_nodes - 1;
    if (tid < ST) s_logscale[tid] = 0.0;
    unsigned inited = 0u;                       // uniform across the CTA
    __syncthreads();
    for (int e = E - 1; e >= -1; --e) {
        // ---- finalise node v = e + 1 (or the root when e == -1): data mask, rescale ----
        const int v = e + 1;
        const int sv = slot[v];
        double* Lv = slots + (size_t)sv * ND * ST;
        const bool has_children = (inited >> sv) & 1u;
        for (int i = tid; i < ND * ST; i += blockDim.x) {
            const int st = i / ST, si = i % ST;
            const long long site = site0 + si;
            double val = has_children ? Lv[i] : 1.0;
            const bool ok = st < n && site < n_sites && ((masks[site * n_nodes + v] >> st) & 1ull);
            Lv[i] = ok ? val : 0.0;
        }
        __syncthreads();
        {
            // per-site maximum over the states: 8 groups of 8 states in parallel, then 8 partials
            const int si = tid % ST, g = tid / ST;          // blockDim.x == 8 * ST
            double mx = 0.0;
            for (int st = g * 8; st < g * 8 + 8; ++st) mx = fmax(mx, Lv[st * ST + si]);   // rows >= n are 0
            s_part[g * ST + si] = mx;
        }
        __syncthreads();
        if (tid < ST) {
            double mx = 0.0;
#pragma unroll
            for (int g = 0; g < 8; ++g) mx = fmax(mx, s_part[g * ST + tid]);
            s_colmax[tid] = mx;
            if (mx > 0.0) s_logscale[tid] += log(mx);
            else if (site0 + tid < n_sites) s_logscale[tid] = -INFINITY;
        }
        __syncthreads();
        for (int i = tid; i < ND * ST; i += blockDim.x) {
            const double mx = s_colmax[i % ST];
            Lv[i] = (mx > 0.0) ? Lv[i] / mx : 0.0;
        }
        __syncthreads();
        if (e < 0) break;
        // ---- push the message of v up edge e: M = P_e . L_v ----
        for (int i = tid; i < ND * ND; i += blockDim.x) {
            const int r = i / ND, c = i % ND;
            Pe[i] = (r < n && c < n) ? P[((size_t)e * n + r) * n + c] : 0.0;
        }
        __syncthreads();
        {
            const int r = lane >> 2, q = lane & 3;
            double a[ND / 4];           // this warp's strip of P_e, reused across the site tiles
#pragma unroll
            for (int ks = 0; ks < ND / 4; ++ks) a[ks] = Pe[(warp * 8 + r) * ND + ks * 4 + q];
            for (int jt = 0; jt < ST / 8; ++jt) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < ND / 4; ++ks)
                    dmma(c0, c1, a[ks], Lv[(ks * 4 + q) * ST + jt * 8 + r]);
                M[(warp * 8 + r) * ST + jt * 8 + 2 * q] = c0;
                M[(warp * 8 + r) * ST + jt * 8 + 2 * q + 1] = c1;
            }
        }
        __syncthreads();
        const int va = parent[v];
        const int sa = slot[va];
        double* La = slots + (size_t)sa * ND * ST;
        const bool first = !((inited >> sa) & 1u);
        for (int i = tid; i < ND * ST; i += blockDim.x) La[i] = first ? M[i] : La[i] * M[i];
        inited &= ~(1u << sv);
        inited |= 1u << sa;
        __syncthreads();
    }
    // ---- root: sum_s prior[s] L_root[s] ----
    const double* Lr = slots + (size_t)slot[0] * ND * ST;
    if (tid < ST && site0 + tid < n_sites) {
        double acc = 0.0;
        for (int st = 0; st < n; ++st) acc += prior[st] * Lr[st * ST + tid];
        loglik[site0 + tid] = (acc > 0.0) ? log(acc) + s_logscale[tid] : -INFINITY;
    }
}

}  // namespace nxbd

static std::string g_dense_err;

extern "C" const char* nxb_dense_last_error(void) { return g_dense_err.c_str(); }

// duration of the kernel of the last dense call (CUDA events around the launch, default stream)
static float g_dense_ms = 0.f;
extern "C" float nxb_dense_last_kernel_ms(void) { return g_dense_ms; }
namespace {
struct KernelTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    KernelTimer() { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, 0); }
    void stop() { cudaEventRecord(b, 0); cudaEventSynchronize(b); cudaEventElapsedTime(&g_dense_ms, a, b); }
    ~KernelTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};
}  // namespace


// Device workspace of the dense calls: grow-only chunks per device, handed out by a bump pointer and
// reset at the start of every call.  (cudaMalloc / cudaFree per call cost 5-12 ms per call on a
// device that already holds the multi-GB unit slabs of a sampler -- cudaFree synchronises the
// device -- against 1.6 ms of kernel; and an early error return no longer leaks.)  Like the rest of
// the C-ABI the dense calls are not re-entrant.
namespace {
struct WsChunk { unsigned char* base; size_t cap, used; };
std::vector<WsChunk> g_ws[64];
void ws_reset(int device) { for (auto& c : g_ws[device & 63]) c.used = 0; }
cudaError_t ws_take(int device, void** out, size_t bytes) {
    auto& chunks = g_ws[device & 63];
    const size_t need = (bytes + 255) & ~(size_t)255;
    for (auto& c : chunks)
        if (c.cap - c.used >= need) { *out = c.base + c.used; c.used += need; return cudaSuccess; }
    size_t cap = std::max<size_t>(need, (size_t)1 << 20);
    if (!chunks.empty()) cap = std::max(cap, 2 * chunks.back().cap);
    WsChunk c{nullptr, cap, need};
    cudaError_t e = cudaMalloc((void**)&c.base, cap);
    if (e != cudaSuccess) return e;
    chunks.push_back(c);
    *out = c.base;
    return cudaSuccess;
}
}  // namespace

#define DTRY(call)                                                                   \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            g_dense_err = std::string(#call) + ": " + cudaGetErrorString(e_);        \
            return NXB_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

extern "C" int nxb_expm_batched(int device, int32_t n, int32_t batch, const double* Q,
                                const double* t, double* out) {
    if (n < 1 || n > ND || batch < 1 || !Q || !t || !out) { g_dense_err = "bad argument (need 1 <= n <= 64)"; return NXB_ERR_BAD_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { g_dense_err = "no CUDA device available (there is no CPU fallback)"; return NXB_ERR_CUDA; }
    DTRY(cudaSetDevice(device));
    ws_reset(device);
    double *dQ = nullptr, *dt = nullptr, *dout = nullptr;
    DTRY(ws_take(device, (void**)&dQ, sizeof(double) * n * n));
    DTRY(ws_take(device, (void**)&dt, sizeof(double) * batch));
    DTRY(ws_take(device, (void**)&dout, sizeof(double) * (size_t)batch * n * n));
    DTRY(cudaMemcpy(dQ, Q, sizeof(double) * n * n, cudaMemcpyHostToDevice));
    DTRY(cudaMemcpy(dt, t, sizeof(double) * batch, cudaMemcpyHostToDevice));
    const size_t smem = sizeof(double) * 3 * ND * ND;
    DTRY(cudaFuncSetAttribute(nxbd::expm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelTimer kt;
        nxbd::expm_kernel<<<batch, 256, smem>>>(n, dQ, dt, dout);
        kt.stop();
    }
    DTRY(cudaGetLastError());
    DTRY(cudaMemcpy(out, dout, sizeof(double) * (size_t)batch * n * n, cudaMemcpyDeviceToHost));
    return NXB_OK;
}

extern "C" int nxb_expm_frechet_batched(int device, int32_t n, int32_t batch, const double* Q,
                                        const double* t, const double* E, double* outP, double* outL) {
    if (n < 1 || n > ND || batch < 1 || !Q || !t || !E || !outL) { g_dense_err = "bad argument (need 1 <= n <= 64)"; return NXB_ERR_BAD_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { g_dense_err = "no CUDA device available (there is no CPU fallback)"; return NXB_ERR_CUDA; }
    DTRY(cudaSetDevice(device));
    ws_reset(device);
    const size_t mat = sizeof(double) * (size_t)batch * n * n;
    double *dQ = nullptr, *dt = nullptr, *dE = nullptr, *dP = nullptr, *dL = nullptr;
    DTRY(ws_take(device, (void**)&dQ, sizeof(double) * n * n));
    DTRY(ws_take(device, (void**)&dt, sizeof(double) * batch));
    DTRY(ws_take(device, (void**)&dE, mat));
    DTRY(ws_take(device, (void**)&dL, mat));
    if (outP) DTRY(ws_take(device, (void**)&dP, mat));
    DTRY(cudaMemcpy(dQ, Q, sizeof(double) * n * n, cudaMemcpyHostToDevice));
    DTRY(cudaMemcpy(dt, t, sizeof(double) * batch, cudaMemcpyHostToDevice));
    DTRY(cudaMemcpy(dE, E, mat, cudaMemcpyHostToDevice));
    const size_t smem = sizeof(double) * 6 * ND * ND;
    DTRY(cudaFuncSetAttribute(nxbd::frechet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelTimer kt;
        nxbd::frechet_kernel<<<batch, 256, smem>>>(n, dQ, dt, dE, dP, dL);
        kt.stop();
    }
    DTRY(cudaGetLastError());
    DTRY(cudaMemcpy(outL, dL, mat, cudaMemcpyDeviceToHost));
    if (outP) DTRY(cudaMemcpy(outP, dP, mat, cudaMemcpyDeviceToHost));
    return NXB_OK;
}

extern "C" int nxb_pruning_loglik(int device, int32_t n, int32_t n_nodes, const int32_t* parent,
                                  const double* P, const double* prior, int64_t n_sites,
                                  const uint64_t* masks, double* loglik) {
    if (n < 1 || n > ND || n_nodes < 2 || n_sites < 1 || !parent || !P || !prior || !masks || !loglik) {
        g_dense_err = "bad argument (need 1 <= n <= 64, n_nodes >= 2)"; return NXB_ERR_BAD_ARG;
    }
    if (parent[0] != -1) { g_dense_err = "parent[0] must be -1"; return NXB_ERR_BAD_ARG; }
    for (int v = 1; v < n_nodes; ++v)
        if (parent[v] < 0 || parent[v] >= v) { g_dense_err = "nodes must be numbered so that parent[v] < v"; return NXB_ERR_BAD_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { g_dense_err = "no CUDA device available (there is no CPU fallback)"; return NXB_ERR_CUDA; }
    DTRY(cudaSetDevice(device));
    ws_reset(device);
    const int E = n_nodes - 1;
    // message slots: a node's slot is live from its first folded child until its own edge
    std::vector<int> slot(n_nodes, -1), last_child(std::max(E, 1), 0);
    int nslots = 0;
    {
        std::vector<int> free_slots;
        for (int e = E - 1; e >= 0; --e) {
            const int vb = e + 1, va = parent[vb];
            if (slot[vb] < 0) {           // leaf: needs a slot just for its own message
                if (!free_slots.empty()) { slot[vb] = free_slots.back(); free_slots.pop_back(); }
                else slot[vb] = nslots++;
            }
            if (slot[va] < 0) {
                if (!free_slots.empty()) { slot[va] = free_slots.back(); free_slots.pop_back(); }
                else slot[va] = nslots++;
            }
            free_slots.push_back(slot[vb]);
        }
    }
    const size_t smem = sizeof(double) * ((size_t)ND * ND + (size_t)ND * ST + (size_t)nslots * ND * ST);
    int max_smem = 0;
    DTRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (smem > (size_t)max_smem) { g_dense_err = "tree needs too many live messages; renumber nodes depth-first"; return NXB_ERR_BAD_ARG; }
    int *dparent = nullptr, *dslot = nullptr, *dlast = nullptr;
    double *dP = nullptr, *dprior = nullptr, *dll = nullptr;
    unsigned long long* dmask = nullptr;
    DTRY(ws_take(device, (void**)&dparent, 4 * n_nodes));
    DTRY(ws_take(device, (void**)&dslot, 4 * n_nodes));
    DTRY(ws_take(device, (void**)&dlast, 4 * std::max(E, 1)));
    DTRY(ws_take(device, (void**)&dP, sizeof(double) * (size_t)E * n * n));
    DTRY(ws_take(device, (void**)&dprior, sizeof(double) * n));
    DTRY(ws_take(device, (void**)&dll, sizeof(double) * n_sites));
    DTRY(ws_take(device, (void**)&dmask, 8 * (size_t)n_sites * n_nodes));
    DTRY(cudaMemcpy(dparent, parent, 4 * n_nodes, cudaMemcpyHostToDevice));
    DTRY(cudaMemcpy(dslot, slot.data(), 4 * n_nodes, cudaMemcpyHostToDevice));
    DTRY(cudaMemcpy(dlast, last_child.data(), 4 * std::max(E, 1), cudaMemcpyHostToDevice));
    DTRY(cudaMemcpy(dP, P, sizeof(double) * (size_t)E * n * n, cudaMemcpyHostToDevice));
    DTRY(cudaMemcpy(dprior, prior, sizeof(double) * n, cudaMemcpyHostToDevice));
    DTRY(cudaMemcpy(dmask, masks, 8 * (size_t)n_sites * n_nodes, cudaMemcpyHostToDevice));
    DTRY(cudaFuncSetAttribute(nxbd::pruning_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((n_sites + ST - 1) / ST);
    {
        KernelTimer kt;
        nxbd::pruning_kernel<<<grid, 256, smem>>>(n, n_nodes, nslots, dparent, dslot, dlast, dP, dprior,
                                                  n_sites, dmask, dll);
        kt.stop();
    }
    DTRY(cudaGetLastError());
    DTRY(cudaMemcpy(loglik, dll, sizeof(double) * n_sites, cudaMemcpyDeviceToHost));
    return NXB_OK;
}
