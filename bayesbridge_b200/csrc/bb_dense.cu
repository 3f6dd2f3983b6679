// Dense design matrix (reference: bayesbridge/design_matrix/dense_matrix.py:9-58).  The
// reference materialises [1, X - mean]; here X is kept raw in HBM (row-major, n x p) and the
// intercept / centring algebra is applied implicitly, exactly as in the sparse class, so the
// CG pipeline above it is shared.
#include "bb_internal.cuh"
#include <stdlib.h>

constexpr int DD_THREADS = 256;
constexpr int DT_COLS = 4;          // columns per thread in the transposed product
constexpr int DT_THREADS = 256;

// y_i = sum_j X_ij sv_j + shift ; one warp per row, rows strided over the grid
template <int MODE>
__global__ void __launch_bounds__(DD_THREADS)
k_dense_dot(const double* __restrict__ X, i64 n, i64 p, const double* __restrict__ sv,
            const double* __restrict__ red_shift, int nshift,
            const double* __restrict__ omega, const double* __restrict__ omega_scalar,
            double* __restrict__ out, double* __restrict__ red_w, const int* __restrict__ done_flag) {
    if (done_flag != nullptr && *done_flag) return;
    __shared__ double sm[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = DD_THREADS / 32;
    const double shift = warp_sum_partials(red_shift, nshift);
    double acc_w = 0.0;
    for (i64 i = (i64)blockIdx.x * wpb + warp; i < n; i += (i64)gridDim.x * wpb) {
        const double* row = X + i * p;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        i64 j = lane;
        for (; j + 96 < p; j += 128) {
            double x0 = row[j], x1 = row[j + 32], x2 = row[j + 64], x3 = row[j + 96];
            a0 += x0 * sv[j]; a1 += x1 * sv[j + 32]; a2 += x2 * sv[j + 64]; a3 += x3 * sv[j + 96];
        }
        for (; j < p; j += 32) a0 += row[j] * sv[j];
        double u = warp_sum((a0 + a1) + (a2 + a3)) + shift;
        if (lane == 0) {
            if (MODE == 0) out[i] = u;
            else { double w = (omega ? omega[i] : omega_scalar[0]) * u; out[i] = w; acc_w += w; }
        }
    }
    if (MODE == 1) {
        acc_w = block_sum(acc_w, sm);
        if (threadIdx.x == 0) red_w[blockIdx.x] = acc_w;
    }
}

// part[rb*p + j] = sum_{i in row block rb} f(X_ij) w_i ; SQ: also the squared version (fisher diag)
template <bool SQ>
__global__ void __launch_bounds__(DT_THREADS)
k_dense_tdot(const double* __restrict__ X, i64 n, i64 p, const double* __restrict__ w, int nblk,
             double* __restrict__ part, double* __restrict__ part_sq, const int* __restrict__ done_flag) {
    if (done_flag != nullptr && *done_flag) return;
    const int rb = blockIdx.x;
    const i64 cbase = (i64)blockIdx.y * (DT_THREADS * DT_COLS);
    const i64 r0 = n * rb / nblk, r1 = n * (rb + 1) / nblk;
    double acc[DT_COLS], acc2[DT_COLS];
#pragma unroll
    for (int k = 0; k < DT_COLS; ++k) { acc[k] = 0.0; acc2[k] = 0.0; }
    for (i64 i = r0; i < r1; ++i) {
        const double wi = w[i];
        const double* row = X + i * p;
#pragma unroll
        for (int k = 0; k < DT_COLS; ++k) {
            i64 j = cbase + (i64)k * DT_THREADS + threadIdx.x;
            if (j < p) {
                double x = row[j];
                acc[k] += x * wi;
                if (SQ) acc2[k] += x * x * wi;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < DT_COLS; ++k) {
        i64 j = cbase + (i64)k * DT_THREADS + threadIdx.x;
        if (j < p) {
            part[(i64)rb * p + j] = acc[k];
            if (SQ) part_sq[(i64)rb * p + j] = acc2[k];
        }
    }
}

__global__ void k_colsum_parts(const double* __restrict__ part, int nblk, i64 p, double* __restrict__ out) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < p; j += (i64)gridDim.x * blockDim.x) {
        double t = 0.0;
        for (int r = 0; r < nblk; ++r) t += part[(i64)r * p + j];
        out[j] = t;
    }
}

static int dense_dot_grid(bb_mat* m) {
    i64 g = (m->n + (DD_THREADS / 32) - 1) / (DD_THREADS / 32);
    i64 cap = (i64)m->ctx->sm_count * 6;
    if (cap > RED_MAX) cap = RED_MAX;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int bb_dense_dot(bb_mat* m, int mode, const int* done_flag) {
    bb_ctx* ctx = m->ctx;
    const double* red_shift = m->red + RED_SHIFT * RED_MAX;
    i64 gP = (m->P + 1023) / 1024; if (gP < 1) gP = 1; if (gP > RED_MAX) gP = RED_MAX;
    int nshift = (int)gP;
    int grid = dense_dot_grid(m);
    const double* sv = m->sv + m->add_intercept;
    if (mode == 0) {
        k_dense_dot<0><<<grid, DD_THREADS, 0, ctx->stream>>>(m->Xd, m->n, m->p, sv, red_shift, nshift, nullptr, nullptr,
                                                             m->u_n, nullptr, done_flag);
    } else {
        k_dense_dot<1><<<grid, DD_THREADS, 0, ctx->stream>>>(m->Xd, m->n, m->p, sv, red_shift, nshift,
                                                             m->use_omega_scalar ? nullptr : m->omega, m->omega_scalar_dev, m->w_n, m->red + RED_W * RED_MAX, done_flag);
        m->nred_w = grid;
    }
    BB_LAUNCHED(ctx);
    return BB_OK;
}

int bb_dense_tdot(bb_mat* m, const double* w, const int* done_flag) {
    bb_ctx* ctx = m->ctx;
    if (m->p == 0) return BB_OK;
    dim3 grid(m->dense_nblk, (unsigned)((m->p + DT_THREADS * DT_COLS - 1) / (DT_THREADS * DT_COLS)));
    k_dense_tdot<false><<<grid, DT_THREADS, 0, ctx->stream>>>(m->Xd, m->n, m->p, w, m->dense_nblk, m->dense_part, nullptr, done_flag);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

int bb_dense_fisher_diag(bb_mat* m, const double* weight_dev, double* d2, double* d1) {
    bb_ctx* ctx = m->ctx;
    if (m->p == 0) return BB_OK;
    double* part_sq = nullptr;
    BB_CUDA(cudaMalloc((void**)&part_sq, (size_t)m->dense_nblk * m->p * sizeof(double)));
    dim3 grid(m->dense_nblk, (unsigned)((m->p + DT_THREADS * DT_COLS - 1) / (DT_THREADS * DT_COLS)));
    k_dense_tdot<true><<<grid, DT_THREADS, 0, ctx->stream>>>(m->Xd, m->n, m->p, weight_dev, m->dense_nblk, m->dense_part, part_sq, nullptr);
    BB_LAUNCHED(ctx);
    int g = (int)((m->p + 255) / 256); if (g > 1024) g = 1024;
    k_colsum_parts<<<g, 256, 0, ctx->stream>>>(m->dense_part, m->dense_nblk, m->p, d1);
    BB_LAUNCHED(ctx);
    k_colsum_parts<<<g, 256, 0, ctx->stream>>>(part_sq, m->dense_nblk, m->p, d2);
    BB_LAUNCHED(ctx);
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(part_sq);
    return BB_OK;
}

extern "C" int bb_dense_upload(bb_ctx* ctx, int64_t n, int64_t p, const double* X,
                               const double* column_offset, int add_intercept,
                               int64_t row_offset, int64_t n_global, bb_mat** out) {
    BB_ARG(ctx && out, "ctx/out");
    BB_ARG(n >= 0 && p >= 0, "negative size");
    BB_ARG(n * p == 0 || X != nullptr, "X");
    BB_CUDA(cudaSetDevice(ctx->device));
    bb_mat* m = (bb_mat*)calloc(1, sizeof(bb_mat));
    m->ctx = ctx;
    m->is_sparse = 0;
    m->is_binary = 0;
    m->add_intercept = add_intercept ? 1 : 0;
    m->centered = (column_offset != nullptr);
    m->n = n; m->p = p; m->P = p + m->add_intercept; m->nnz = n * p;
    m->row_offset = row_offset; m->n_global = n_global > 0 ? n_global : n;
    int rc = BB_OK;
    do {
#define CKC(e) { cudaError_t e_ = (e); if (e_ != cudaSuccess) { bb_set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); rc = BB_ERR_CUDA; break; } }
        size_t nb = (size_t)(n * p > 0 ? n * p : 1) * sizeof(double);
        CKC(cudaMalloc((void**)&m->Xd, nb));
        if (n * p > 0) CKC(cudaMemcpyAsync(m->Xd, X, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CKC(cudaMalloc((void**)&m->col_offset, (size_t)(p > 0 ? p : 1) * sizeof(double)));
        CKC(cudaMemsetAsync(m->col_offset, 0, (size_t)(p > 0 ? p : 1) * sizeof(double), ctx->stream));
        if (column_offset && p > 0)
            CKC(cudaMemcpyAsync(m->col_offset, column_offset, (size_t)p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        i64 nblk = ctx->sm_count * 2;
        if (nblk > n) nblk = n > 0 ? n : 1;
        if (nblk > RED_MAX) nblk = RED_MAX;
        m->dense_nblk = (int)nblk;
        CKC(cudaMalloc((void**)&m->dense_part, (size_t)nblk * (p > 0 ? p : 1) * sizeof(double)));
        if ((rc = bb_mat_alloc_work(m)) != BB_OK) break;
        CKC(cudaStreamSynchronize(ctx->stream));
#undef CKC
    } while (0);
    if (rc != BB_OK) { bb_mat_free(m); return rc; }
    *out = m;
    return BB_OK;
}
