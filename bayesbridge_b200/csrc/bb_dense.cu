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

// ------------------------------------------------------------------------------------------
// Streaming kernel: X is read ONCE per operator application (dense_matrix.py:37-58 does two GEMV passes).
// One persistent CTA per SM walks groups of R consecutive rows; a group is one contiguous chunk of R*p doubles, brought
// into shared memory by TMA bulk copies (cp.async.bulk + mbarrier, two stages), so HBM sees long sequential bursts.
// Thread t owns columns t, t+1024, ... (KMAX of them): its entries of the gather vector and its column accumulators
// stay in registers for the whole launch.  Per group:
//   pass 1  u_r = shift + sum_j X_rj sv_j      (block reduction, fixed tree)          -- DS_DOT, DS_DOT_W, DS_FUSED
//           w_r = omega_r u_r                                                         -- DS_DOT_W, DS_FUSED
//   pass 2  acc_j += w_r X_rj  from the same shared-memory copy                        -- DS_TDOT, DS_FUSED
// and at the end part[cta][j] = acc_j (summed over CTAs by the consumer in CTA order).  Algorithmic bytes of the fused
// operator: 8 n p (+ vectors), SURVEY section 8d.
constexpr int DS_THREADS = 256;      // few threads with many columns each: the reductions per row stay cheap, TMA does the loading
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_RMAX = 8;
enum { DS_DOT = 0, DS_DOT_W = 1, DS_TDOT = 2, DS_FUSED = 3 };

struct DenseStreamArgs {
    const double* X; i64 n, p; int R, nstage;
    const double* sv; const double* red_shift; int nshift;
    const double* omega; const double* omega_scalar;
    const double* w_in;
    double* out; double* red_w; double* part;
    const int* done_flag;
};

__device__ __forceinline__ void ds_issue(const DenseStreamArgs& a, i64 g, unsigned dst, unsigned mbar) {
    const i64 i0 = g * a.R;
    i64 rows = a.n - i0; if (rows > a.R) rows = a.R;
    const unsigned bytes = (unsigned)((((size_t)rows * a.p + 1) & ~(size_t)1) * sizeof(double));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    const char* src = reinterpret_cast<const char*>(a.X + i0 * a.p);
    for (unsigned off = 0; off < bytes; off += 32768u) {
        const unsigned sz = min(32768u, bytes - off);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + off), "l"(src + off), "r"(sz), "r"(mbar) : "memory");
    }
}

// RR = rows per group (compile time, so the per-row work of a group is independent instruction streams)
template <int KMAX, int MODE, int RR>
__global__ void __launch_bounds__(DS_THREADS, 1)
k_dense_stream(const DenseStreamArgs a) {
    if (a.done_flag != nullptr && *a.done_flag) return;
    extern __shared__ __align__(128) unsigned char ds_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 p = a.p;
    const size_t stage_elems = ((size_t)RR * p + 1) & ~(size_t)1;
    double* stage0 = reinterpret_cast<double*>(ds_smem);
    double* red = stage0 + (size_t)a.nstage * stage_elems;                    // [RR][DS_WARPS]
    unsigned long long* mbar_p = reinterpret_cast<unsigned long long*>(red + DS_RMAX * DS_WARPS);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage0);
    const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(mbar_p);
    const i64 ngroups = (a.n + RR - 1) / RR;
    if (tid == 0) {
        for (int s = 0; s < a.nstage; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0 + 8u * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < a.nstage; ++s) {
            const i64 g = (i64)blockIdx.x + (i64)s * gridDim.x;
            if (g < ngroups) ds_issue(a, g, sbase + (unsigned)(s * stage_elems * sizeof(double)), mbar0 + 8u * s);
        }
    double svr[KMAX], acc[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const i64 j = tid + (i64)k * DS_THREADS;
        svr[k] = (MODE != DS_TDOT && j < p) ? a.sv[j] : 0.0;
        acc[k] = 0.0;
    }
    double shift = 0.0;
    if (MODE != DS_TDOT) shift = warp_sum_partials(a.red_shift, a.nshift);
    double sw = 0.0;
    unsigned phase_bits = 0;
    int s = 0;
    for (i64 g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const unsigned mb = mbar0 + 8u * s;
        const i64 i0 = g * RR;
        const int rows = (int)((a.n - i0 < RR) ? (a.n - i0) : RR);
        // per-row scalars of this group: loaded before the wait so that their latency hides behind it
        double scal[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            scal[r] = 0.0;
            if (r < rows) {
                if (MODE == DS_TDOT) scal[r] = a.w_in[i0 + r];
                else if (MODE != DS_DOT) scal[r] = a.omega ? a.omega[i0 + r] : a.omega_scalar[0];
            }
        }
        {
            unsigned ok = 0;
            const unsigned ph = (phase_bits >> s) & 1u;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\t"
                             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(mb), "r"(ph) : "memory");
            }
            phase_bits ^= 1u << s;
        }
        const double* xs = stage0 + (size_t)s * stage_elems;
        double wr[RR];
        if (MODE != DS_TDOT) {
            double t[RR];
#pragma unroll
            for (int r = 0; r < RR; ++r) t[r] = 0.0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                const i64 j = tid + (i64)k * DS_THREADS;
                if (j < p) {
#pragma unroll
                    for (int r = 0; r < RR; ++r) t[r] += xs[(size_t)r * p + j] * svr[k];    // rows past the end read stale, finite data
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int r = 0; r < RR; ++r) t[r] += __shfl_xor_sync(0xffffffffu, t[r], o);
            }
            if (lane < RR) {
                double mine = t[0];
#pragma unroll
                for (int r = 1; r < RR; ++r) if (lane == r) mine = t[r];
                red[lane * DS_WARPS + warp] = mine;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                double u = (lane < DS_WARPS) ? red[r * DS_WARPS + lane] : 0.0;
#pragma unroll
                for (int o = DS_WARPS / 2; o > 0; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
                u = __shfl_sync(0xffffffffu, u, 0) + shift;
                wr[r] = 0.0;
                if (r < rows) {
                    if (MODE == DS_DOT) {
                        if (tid == 0) a.out[i0 + r] = u;
                    } else {
                        const double w = scal[r] * u;
                        wr[r] = w;
                        sw += w;
                        if (tid == 0 && a.out != nullptr) a.out[i0 + r] = w;
                    }
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < RR; ++r) wr[r] = scal[r];
        }
        if (MODE == DS_TDOT || MODE == DS_FUSED) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                const i64 j = tid + (i64)k * DS_THREADS;
                if (j < p) {
#pragma unroll
                    for (int r = 0; r < RR; ++r)
                        if (r < rows) acc[k] += wr[r] * xs[(size_t)r * p + j];
                }
            }
        }
        __syncthreads();                          // everybody is done with stage s (and with red[])
        const i64 gn = g + (i64)a.nstage * gridDim.x;
        if (tid == 0 && gn < ngroups) ds_issue(a, gn, sbase + (unsigned)(s * stage_elems * sizeof(double)), mb);
        s = (s + 1 == a.nstage) ? 0 : s + 1;
    }
    if (MODE == DS_TDOT || MODE == DS_FUSED) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const i64 j = tid + (i64)k * DS_THREADS;
            if (j < p) a.part[(i64)blockIdx.x * p + j] = acc[k];
        }
    }
    if ((MODE == DS_DOT_W || MODE == DS_FUSED) && tid == 0) a.red_w[blockIdx.x] = sw;
}

// stage geometry: R rows (even, <= DS_RMAX) per group, nstage buffers; R = 0 when a row pair does not fit
static void ds_geometry(bb_ctx* ctx, i64 p, int* R_out, int* nstage_out, size_t* smem_out) {
    *R_out = 0; *nstage_out = 0; *smem_out = 0;
    if (p < 1 || p > (i64)32 * DS_THREADS) return;
    const size_t fixed = DS_WARPS * DS_RMAX * sizeof(double) + 64;
    for (int nstage = 2; nstage >= 1; --nstage)
        for (int R = DS_RMAX; R >= 2; R /= 2) {
            const size_t stage = (((size_t)R * p + 1) & ~(size_t)1) * sizeof(double);
            const size_t need = nstage * stage + fixed;
            if (need <= ctx->smem_optin) { *R_out = R; *nstage_out = nstage; *smem_out = need; return; }
        }
}

bool bb_dense_stream_ok(bb_mat* m) {
    int R, ns; size_t sm;
    ds_geometry(m->ctx, m->p, &R, &ns, &sm);
    return R > 0 && m->ctx->opt_dense_stream != 0;
}

template <int MODE>
static int ds_launch_mode(bb_mat* m, const DenseStreamArgs& a, int grid, size_t smem) {
    bb_ctx* ctx = m->ctx;
    const int kmax = (int)((m->p + DS_THREADS - 1) / DS_THREADS);
#define DS_LAUNCH(K, RRV)                                                                                             \
    {                                                                                                                 \
        static BBDeviceOnce attr = {{0, 0, 0, 0}};                                                                    \
        if (attr.first(ctx->device)) {                                                                                \
            BB_CUDA(cudaFuncSetAttribute(k_dense_stream<K, MODE, RRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin)); \
        }                                                                                                             \
        k_dense_stream<K, MODE, RRV><<<grid, DS_THREADS, smem, ctx->stream>>>(a);                                    \
    }
#define DS_CASE(K)                                                                                                    \
    {                                                                                                                 \
        if (a.R == 8) DS_LAUNCH(K, 8) else if (a.R == 4) DS_LAUNCH(K, 4) else DS_LAUNCH(K, 2)                          \
    }
    if (kmax <= 4) DS_CASE(4)
    else if (kmax <= 8) DS_CASE(8)
    else if (kmax <= 12) DS_CASE(12)
    else if (kmax <= 16) DS_CASE(16)
    else if (kmax <= 20) DS_CASE(20)
    else if (kmax <= 24) DS_CASE(24)
    else DS_CASE(32)
#undef DS_CASE
#undef DS_LAUNCH
    BB_LAUNCHED(ctx);
    return BB_OK;
}

// mode: DS_DOT (u -> m->u_n), DS_DOT_W (omega u -> m->w_n), DS_TDOT (w given), DS_FUSED
int bb_dense_stream(bb_mat* m, int mode, const double* w, const int* done_flag) {
    bb_ctx* ctx = m->ctx;
    int R, nstage; size_t smem;
    ds_geometry(ctx, m->p, &R, &nstage, &smem);
    if (R == 0) { bb_set_error("dense streaming kernel: p = %lld does not fit", (long long)m->p); return BB_ERR_ARG; }
    DenseStreamArgs a;
    memset(&a, 0, sizeof(a));
    a.X = m->Xd; a.n = m->n; a.p = m->p; a.R = R; a.nstage = nstage;
    a.sv = m->sv + m->add_intercept;
    a.red_shift = m->red + RED_SHIFT * RED_MAX;
    i64 gP = (m->P + 1023) / 1024; if (gP < 1) gP = 1; if (gP > RED_MAX) gP = RED_MAX;
    a.nshift = (int)gP;
    a.omega = m->use_omega_scalar ? nullptr : m->omega;
    a.omega_scalar = m->omega_scalar_dev;
    a.w_in = w;
    a.red_w = m->red + RED_W * RED_MAX;
    a.part = m->dense_part;
    a.done_flag = done_flag;
    const int grid = m->dense_nblk;
    switch (mode) {
    case DS_DOT: a.out = m->u_n; return ds_launch_mode<DS_DOT>(m, a, grid, smem);
    case DS_DOT_W: a.out = m->w_n; m->nred_w = grid; return ds_launch_mode<DS_DOT_W>(m, a, grid, smem);
    case DS_TDOT: return ds_launch_mode<DS_TDOT>(m, a, grid, smem);
    default: a.out = nullptr; m->nred_w = grid; return ds_launch_mode<DS_FUSED>(m, a, grid, smem);
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
    if (m->dense_stream) return bb_dense_stream(m, mode == 0 ? DS_DOT : DS_DOT_W, nullptr, done_flag);
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
    if (m->dense_stream) return bb_dense_stream(m, DS_TDOT, w, done_flag);
    dim3 grid(m->dense_nblk, (unsigned)((m->p + DT_THREADS * DT_COLS - 1) / (DT_THREADS * DT_COLS)));
    k_dense_tdot<false><<<grid, DT_THREADS, 0, ctx->stream>>>(m->Xd, m->n, m->p, w, m->dense_nblk, m->dense_part, nullptr, done_flag);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

// the fused operator of the CG loop: one pass over X
int bb_dense_fused(bb_mat* m, const int* done_flag) { return bb_dense_stream(m, DS_FUSED, nullptr, done_flag); }

int bb_dense_fisher_diag(bb_mat* m, const double* weight_dev, double* d2, double* d1) {
    bb_ctx* ctx = m->ctx;
    if (m->p == 0) return BB_OK;
    double* part_sq = nullptr;       // persistent context scratch: no allocation per call
    BB_TRY(bb_ctx_scratch(ctx, 1, (size_t)m->dense_nblk * m->p * sizeof(double), (void**)&part_sq));
    dim3 grid(m->dense_nblk, (unsigned)((m->p + DT_THREADS * DT_COLS - 1) / (DT_THREADS * DT_COLS)));
    k_dense_tdot<true><<<grid, DT_THREADS, 0, ctx->stream>>>(m->Xd, m->n, m->p, weight_dev, m->dense_nblk, m->dense_part, part_sq, nullptr);
    BB_LAUNCHED(ctx);
    int g = (int)((m->p + 255) / 256); if (g > 1024) g = 1024;
    k_colsum_parts<<<g, 256, 0, ctx->stream>>>(m->dense_part, m->dense_nblk, m->p, d1);
    BB_LAUNCHED(ctx);
    k_colsum_parts<<<g, 256, 0, ctx->stream>>>(part_sq, m->dense_nblk, m->p, d2);
    BB_LAUNCHED(ctx);
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
        size_t nb = (size_t)(n * p > 0 ? n * p : 1) * sizeof(double) + 64;    // the bulk copies round a chunk up to 16 bytes
        CKC(cudaMalloc((void**)&m->Xd, nb));
        CKC(cudaMemsetAsync(m->Xd, 0, nb, ctx->stream));
        if (n * p > 0) CKC(cudaMemcpyAsync(m->Xd, X, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CKC(cudaMalloc((void**)&m->col_offset, (size_t)(p > 0 ? p : 1) * sizeof(double)));
        CKC(cudaMemsetAsync(m->col_offset, 0, (size_t)(p > 0 ? p : 1) * sizeof(double), ctx->stream));
        if (column_offset && p > 0)
            CKC(cudaMemcpyAsync(m->col_offset, column_offset, (size_t)p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        i64 nblk = ctx->sm_count * 2;
        if (nblk > n) nblk = n > 0 ? n : 1;
        if (nblk > RED_MAX) nblk = RED_MAX;
        m->dense_stream = bb_dense_stream_ok(m) ? 1 : 0;
        if (m->dense_stream) {       // one partial row per persistent CTA of the streaming kernel
            int R, ns; size_t sm;
            ds_geometry(ctx, p, &R, &ns, &sm);
            const i64 ngroups = (n + R - 1) / R;
            nblk = ctx->sm_count < ngroups ? ctx->sm_count : (ngroups > 0 ? ngroups : 1);
        }
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
