// The comparator of BASELINE config 2 ("cg vs cholesky"): full Fisher information X'WX and the direct Gaussian draw.
//
// Reference: design_matrix/dense_matrix.py:54-58 and sparse_matrix.py:131-162 (compute_fisher_info, diag_only=False);
// reg_coef_sampler/direct_gaussian_sampler.py:4-44 (generate_gaussian_with_weight, compute_precond_post_prec):
//     d = pps^2 + diag(X'WX) ; J = d^-1/2 ; Prec = J X'WX J + diag((J pps)^2) ; U'U = Prec (Cholesky, upper)
//     mean = (U'U)^-1 (J z) ; sample = J (mean + U^-1 g),  g ~ N(0, I)
// X'WX is this library's own fp64 tensor-core kernel (mma.sync m8n8k4 f64 -- tcgen05 has no fp64 kind): 128 x 128
// output tiles of the lower triangle, X streamed through shared memory 16 rows at a time, the weight applied to one
// operand on the way in; the intercept / centring terms are added algebraically afterwards from X'w and sum(w), so X
// is used raw, exactly as in the products.  The factorisation itself (P^3/3 flops, 3 % of the work at config 2) is
// cuSOLVER's potrf, loaded with dlopen -- plain LAPACK, not the path this library is about; the three triangular
// solves around it are small kernels here.  A sparse design is densified on the device first (the Cholesky sampler is
// only sensible for p up to ~1e4 anyway).
#include "bb_internal.cuh"
#include <dlfcn.h>
#include <stdlib.h>

constexpr int FT = 128;          // output tile
constexpr int FK = 16;           // rows of X per stage
constexpr int FS = FT + 4;       // padded shared-memory row stride: the m8n8k4 fragment loads hit 16 distinct banks
constexpr int F_THREADS = 512;   // 16 warps, 4 x 4, each 32 x 32 of the tile

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// G0[ja][jb] = sum_i w_i X[i][ja] X[i][jb] for the tile pair (ti >= tj) of this block; mirrored into the upper triangle
__global__ void __launch_bounds__(F_THREADS)
k_fisher_syrk(const double* __restrict__ X, i64 n, i64 p, const double* __restrict__ w, double* __restrict__ G0) {
    extern __shared__ __align__(16) double fs_smem[];
    double* sA = fs_smem;                          // [2][FK][FS]
    double* sB = fs_smem + 2 * FK * FS;
    // tile pair from the linear block index: b = ti (ti + 1) / 2 + tj
    int ti = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((i64)(ti + 1) * (ti + 2) / 2 <= (i64)blockIdx.x) ++ti;
    while ((i64)ti * (ti + 1) / 2 > (i64)blockIdx.x) --ti;
    const int tj = (int)((i64)blockIdx.x - (i64)ti * (ti + 1) / 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const i64 ca = (i64)ti * FT, cb = (i64)tj * FT;
    const int col = tid & (FT - 1), row0 = tid >> 7;      // this thread stages column `col` of rows row0, row0+4, ..
    const bool ina = (ca + col) < p, inb = (cb + col) < p;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    double ra[4], rb[4];
    auto fetch = [&](i64 i0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const i64 i = i0 + row0 + 4 * k;
            const bool ok = i < n;
            ra[k] = (ok && ina) ? X[i * p + ca + col] : 0.0;
            rb[k] = (ok && inb) ? X[i * p + cb + col] * w[i] : 0.0;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sA[(buf * FK + row0 + 4 * k) * FS + col] = ra[k];
            sB[(buf * FK + row0 + 4 * k) * FS + col] = rb[k];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (i64 i0 = 0; i0 < n; i0 += FK) {
        const bool more = (i0 + FK) < n;
        if (more) fetch(i0 + FK);                  // global loads in flight while this stage is multiplied
        const double* a_base = sA + buf * FK * FS + (lane & 3) * FS + wm * 32 + (lane >> 2);
        const double* b_base = sB + buf * FK * FS + (lane & 3) * FS + wn * 32 + (lane >> 2);
#pragma unroll
        for (int k4 = 0; k4 < FK / 4; ++k4) {
            double af[4], bf[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { af[q] = a_base[k4 * 4 * FS + q * 8]; bf[q] = b_base[k4 * 4 * FS + q * 8]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const i64 ja = ca + wm * 32 + a * 8 + (lane >> 2);
                const i64 jb = cb + wn * 32 + b * 8 + (lane & 3) * 2 + e;
                if (ja < p && jb < p) {
                    G0[ja * p + jb] = acc[a][b][e];
                    if (ti != tj) G0[jb * p + ja] = acc[a][b][e];
                }
            }
}

// sparse -> dense image of the local rows (row-major n x p)
__global__ void k_densify(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                          i64 n, i64 p, double* __restrict__ Xd) {
    const i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    for (int k = ptr[i] + lane; k < ptr[i + 1]; k += 32) atomicAdd(&Xd[i * p + idx[k]], val ? val[k] : 1.0);   // duplicates add up
}

// G (P x P) from G0 = X'WX, t = [sum w; X'w], c: the intercept / centring algebra of sparse_matrix.py:144-160
//   MODE 0: out = G ; MODE 1: out = J G J + diag((J pps)^2) with J = (pps^2 + diag G)^-1/2 (also written to Jout)
__device__ __forceinline__ double fisher_entry(const double* __restrict__ G0, const double* __restrict__ t,
                                               const double* __restrict__ c, i64 p, int icpt, int centered, i64 a, i64 b) {
    const double sw = t[0];
    if (a < icpt && b < icpt) return sw;
    if (a < icpt || b < icpt) {
        const i64 j = (a < icpt ? b : a) - icpt;
        return t[1 + j] - sw * c[j];
    }
    const i64 j = a - icpt, k = b - icpt;
    double g = G0[j * p + k];
    if (centered) g += -c[j] * t[1 + k] - t[1 + j] * c[k] + sw * c[j] * c[k];
    return g;
}

__global__ void k_fisher_scale_vec(const double* __restrict__ G0, const double* __restrict__ t, const double* __restrict__ c,
                                   i64 p, int icpt, int centered, const double* __restrict__ pps, double* __restrict__ J) {
    const i64 P = p + icpt;
    for (i64 a = (i64)blockIdx.x * blockDim.x + threadIdx.x; a < P; a += (i64)gridDim.x * blockDim.x) {
        const double d = pps[a] * pps[a] + fisher_entry(G0, t, c, p, icpt, centered, a, a);
        J[a] = 1.0 / sqrt(d);
    }
}

template <int MODE>
__global__ void k_fisher_assemble(const double* __restrict__ G0, const double* __restrict__ t, const double* __restrict__ c,
                                  i64 p, int icpt, int centered, const double* __restrict__ pps, const double* __restrict__ J,
                                  double* __restrict__ out) {
    const i64 P = p + icpt;
    const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (b >= P) return;
    double g = fisher_entry(G0, t, c, p, icpt, centered, a, b);
    if (MODE == 1) {
        g = J[a] * g * J[b];
        if (a == b) { const double s = J[a] * pps[a]; g += s * s; }
    }
    out[a * P + b] = g;
}

// Triangular solves with the factor U (upper, column-major as cuSOLVER leaves it: U(i,j) at a[i + j P], i <= j), one CTA.
//   TRANS = 0: U x = b (back substitution) ; TRANS = 1: U' x = b (forward substitution).  In place on x.
template <int TRANS>
__global__ void __launch_bounds__(1024) k_trsv_upper(const double* __restrict__ U, i64 P, double* __restrict__ x) {
    __shared__ double xb[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 nblk = (P + 31) / 32;
    for (i64 bb = 0; bb < nblk; ++bb) {
        const i64 blk = TRANS ? bb : nblk - 1 - bb;
        const i64 j0 = blk * 32;
        const int bw = (int)((P - j0 < 32) ? (P - j0) : 32);
        if (warp == 0) {
            // 32 x 32 diagonal block by one warp: lane l owns unknown j0 + l
            double v = (lane < bw) ? x[j0 + lane] : 0.0;
            if (TRANS) {
                for (int k = 0; k < bw; ++k) {
                    const double xk = __shfl_sync(0xffffffffu, v, k) / U[(j0 + k) + (j0 + k) * P];
                    if (lane == k) v = xk;
                    if (lane > k && lane < bw) v -= U[(j0 + k) + (j0 + lane) * P] * xk;      // U'(l,k) = U(k,l)
                }
            } else {
                for (int k = bw - 1; k >= 0; --k) {
                    const double xk = __shfl_sync(0xffffffffu, v, k) / U[(j0 + k) + (j0 + k) * P];
                    if (lane == k) v = xk;
                    if (lane < k) v -= U[(j0 + lane) + (j0 + k) * P] * xk;
                }
            }
            if (lane < bw) { x[j0 + lane] = v; xb[lane] = v; }
        }
        __syncthreads();
        // update the remaining unknowns with the block just solved
        if (TRANS) {
            for (i64 i = j0 + 32 + tid; i < P; i += 1024) {
                double s = 0.0;
                for (int k = 0; k < bw; ++k) s += U[(j0 + k) + i * P] * xb[k];               // column i of U: contiguous
                x[i] -= s;
            }
        } else {
            for (i64 i = tid; i < j0; i += 1024) {
                double s = 0.0;
                for (int k = 0; k < bw; ++k) s += U[i + (j0 + k) * P] * xb[k];               // coalesced over i
                x[i] -= s;
            }
        }
        __syncthreads();
    }
}

__global__ void k_fill_const(double* __restrict__ out, i64 n, const double* __restrict__ value) {
    const double v = value[0];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) out[i] = v;
}
__global__ void k_vec_mul(const double* __restrict__ a, const double* __restrict__ b, i64 n, double* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) out[i] = a[i] * b[i];
}
__global__ void k_chol_finish(const double* __restrict__ J, const double* __restrict__ mean, const double* __restrict__ x,
                              i64 n, double* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        out[i] = J[i] * (mean[i] + x[i]);
}

// ---- cuSOLVER through dlopen (potrf only) ----------------------------------------------------------
typedef int (*cs_create_t)(void**);
typedef int (*cs_set_stream_t)(void*, cudaStream_t);
typedef int (*cs_potrf_bs_t)(void*, int, int, double*, int, int*);
typedef int (*cs_potrf_t)(void*, int, int, double*, int, double*, int, int*);
struct CusolverApi { void* lib; void* handle; cs_potrf_bs_t bufsize; cs_potrf_t potrf; };
static CusolverApi g_cs = {nullptr, nullptr, nullptr, nullptr};

static int cusolver_load(bb_ctx* ctx) {
    if (g_cs.handle) return BB_OK;
    const char* cands[] = {getenv("BB_CUSOLVER"), "libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so", "libcusolver.so.12"};
    void* lib = nullptr;
    for (const char* c : cands) { if (c && c[0]) { lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL); if (lib) break; } }
    if (!lib) { bb_set_error("cannot dlopen cuSOLVER (set BB_CUSOLVER): %s", dlerror()); return BB_ERR_STATE; }
    cs_create_t create = (cs_create_t)dlsym(lib, "cusolverDnCreate");
    cs_set_stream_t set_stream = (cs_set_stream_t)dlsym(lib, "cusolverDnSetStream");
    g_cs.bufsize = (cs_potrf_bs_t)dlsym(lib, "cusolverDnDpotrf_bufferSize");
    g_cs.potrf = (cs_potrf_t)dlsym(lib, "cusolverDnDpotrf");
    if (!create || !set_stream || !g_cs.bufsize || !g_cs.potrf) { bb_set_error("cuSOLVER symbols not found"); return BB_ERR_STATE; }
    void* h = nullptr;
    if (create(&h) != 0) { bb_set_error("cusolverDnCreate failed"); return BB_ERR_STATE; }
    if (set_stream(h, ctx->stream) != 0) { bb_set_error("cusolverDnSetStream failed"); return BB_ERR_STATE; }
    g_cs.lib = lib; g_cs.handle = h;
    return BB_OK;
}

// ---- host side -------------------------------------------------------------------------------------
struct FisherWork { double* Xdense; bool own_x; double* G0; };

// G0 = X'WX of the local rows (device, p x p), summed over the row shards; weight on the device (n doubles)
static int fisher_g0(bb_mat* m, const double* w_dev, FisherWork* fw) {
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const i64 n = m->n, p = m->p;
    fw->Xdense = m->Xd; fw->own_x = false; fw->G0 = nullptr;
    if (m->is_sparse) {
        if ((double)n * (double)p * 8.0 > 16e9) { bb_set_error("full Fisher information of a sparse design: the dense image (%lld x %lld) is too large", (long long)n, (long long)p); return BB_ERR_ARG; }
        BB_CUDA(cudaMalloc((void**)&fw->Xdense, (size_t)(n * p > 0 ? n * p : 1) * sizeof(double)));
        fw->own_x = true;
        BB_CUDA(cudaMemsetAsync(fw->Xdense, 0, (size_t)(n * p > 0 ? n * p : 1) * sizeof(double), st));
        if (n > 0 && m->nnz > 0) {
            k_densify<<<(int)((n * 32 + 255) / 256), 256, 0, st>>>(m->csr_ptr, m->csr_idx, m->csr_val, n, p, fw->Xdense);
            BB_LAUNCHED(ctx);
        }
    }
    BB_CUDA(cudaMalloc((void**)&fw->G0, (size_t)(p * p > 0 ? p * p : 1) * sizeof(double)));
    if (p > 0) {
        static BBDeviceOnce attr = {{0, 0, 0, 0}};
        const size_t smem = (size_t)4 * FK * FS * sizeof(double);
        if (attr.first(ctx->device)) BB_CUDA(cudaFuncSetAttribute(k_fisher_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const i64 nt = (p + FT - 1) / FT;
        k_fisher_syrk<<<(unsigned)(nt * (nt + 1) / 2), F_THREADS, smem, st>>>(fw->Xdense, n, p, w_dev, fw->G0);
        BB_LAUNCHED(ctx);
        BB_TRY(bb_allreduce_dev(ctx, fw->G0, p * p));
    }
    return BB_OK;
}
static void fisher_free(FisherWork* fw) {
    if (fw->own_x && fw->Xdense) cudaFree(fw->Xdense);
    if (fw->G0) cudaFree(fw->G0);
    fw->Xdense = nullptr; fw->G0 = nullptr;
}

// compute_fisher_info(weight, diag_only=False): out is P x P, row-major (symmetric)
extern "C" int bb_fisher_full(bb_mat* m, const double* weight, double* out, double* device_ms) {
    BB_ARG(m && weight && out, "mat/weight/out");
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    const i64 P = m->P;
    BB_CUDA(cudaMemcpyAsync(m->eps_n, weight, (size_t)m->n * sizeof(double), cudaMemcpyHostToDevice, st));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (device_ms) { BB_CUDA(cudaEventCreate(&e0)); BB_CUDA(cudaEventCreate(&e1)); BB_CUDA(cudaEventRecord(e0, st)); }
    FisherWork fw;
    int rc = fisher_g0(m, m->eps_n, &fw);
    if (rc != BB_OK) { fisher_free(&fw); return rc; }
    if (device_ms) BB_CUDA(cudaEventRecord(e1, st));
    BB_TRY(bb_op_tdot(m, m->eps_n));                 // traw = [sum w; X'w], all-reduced
    double* G = nullptr;
    BB_CUDA(cudaMalloc((void**)&G, (size_t)(P * P > 0 ? P * P : 1) * sizeof(double)));
    if (P > 0) {
        dim3 grid((unsigned)((P + 255) / 256), (unsigned)P);
        k_fisher_assemble<0><<<grid, 256, 0, st>>>(fw.G0, m->traw, m->col_offset, m->p, m->add_intercept, m->centered, nullptr, nullptr, G);
        BB_LAUNCHED(ctx);
    }
    BB_CUDA(cudaMemcpyAsync(out, G, (size_t)P * P * sizeof(double), cudaMemcpyDeviceToHost, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    if (device_ms) {
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *device_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(G);
    fisher_free(&fw);
    return BB_OK;
}

// generate_gaussian_with_weight: coef ~ N(Sigma z, Sigma), Sigma^-1 = X' diag(omega) X + diag(pps)^2, with the Gaussian
// vector g supplied by the caller (the reference draws it from numpy's global stream, direct_gaussian_sampler.py:26-29).
//   omega: host pointer, or NULL to use the resident precisions.  stats[0] = ms of X'WX, stats[1] = ms of the factorisation.
extern "C" int bb_cholesky_sample(bb_mat* m, const double* omega, const double* prior_prec_sqrt, const double* z,
                                  const double* gaussian_vec, double* coef_out, double* stats) {
    BB_ARG(m && prior_prec_sqrt && z && gaussian_vec && coef_out, "null pointer");
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BB_ARG(m->P < ((i64)1 << 15) * 2, "P too large for the dense Cholesky sampler");
    BB_TRY(cusolver_load(ctx));
    BBTimer timer_(ctx);
    const i64 P = m->P;
    const size_t Pb = (size_t)P * sizeof(double);
    if (omega) { BB_CUDA(cudaMemcpyAsync(m->omega, omega, (size_t)m->n * sizeof(double), cudaMemcpyHostToDevice, st)); m->use_omega_scalar = 0; }
    BB_CUDA(cudaMemcpyAsync(m->pps, prior_prec_sqrt, Pb, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(m->z, z, Pb, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(m->eps_P, gaussian_vec, Pb, cudaMemcpyHostToDevice, st));
    // the weights as a vector (linear model: omega = scalar * 1)
    const double* w_dev = m->omega;
    if (m->use_omega_scalar) {
        BB_CUDA(cudaMemcpyAsync(m->omega_scalar_dev, &m->omega_scalar, sizeof(double), cudaMemcpyHostToDevice, st));
        k_fill_const<<<256, 256, 0, st>>>(m->eps_n, m->n, m->omega_scalar_dev);
        BB_LAUNCHED(ctx);
        w_dev = m->eps_n;
    }
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    for (auto& e : ev) BB_CUDA(cudaEventCreate(&e));
    BB_CUDA(cudaEventRecord(ev[0], st));
    FisherWork fw;
    int rc = fisher_g0(m, w_dev, &fw);
    if (rc != BB_OK) { fisher_free(&fw); return rc; }
    BB_CUDA(cudaEventRecord(ev[1], st));
    BB_TRY(bb_op_tdot(m, w_dev));                    // traw = [sum w; X'w]
    double *Prec = nullptr, *work = nullptr; int* info_dev = nullptr;
    BB_CUDA(cudaMalloc((void**)&Prec, (size_t)(P * P > 0 ? P * P : 1) * sizeof(double)));
    BB_CUDA(cudaMalloc((void**)&info_dev, sizeof(int)));
    double* J = m->D;                                // P-vectors of the CG work space serve as scratch
    double* rhs = m->b;
    double* xg = m->q;
    const int gP = (int)((P + 255) / 256 > 0 ? (P + 255) / 256 : 1);
    k_fisher_scale_vec<<<gP, 256, 0, st>>>(fw.G0, m->traw, m->col_offset, m->p, m->add_intercept, m->centered, m->pps, J);
    BB_LAUNCHED(ctx);
    {
        dim3 grid((unsigned)((P + 255) / 256), (unsigned)P);
        k_fisher_assemble<1><<<grid, 256, 0, st>>>(fw.G0, m->traw, m->col_offset, m->p, m->add_intercept, m->centered, m->pps, J, Prec);
        BB_LAUNCHED(ctx);
    }
    int lwork = 0;
    if (g_cs.bufsize(g_cs.handle, /*CUBLAS_FILL_MODE_UPPER*/ 1, (int)P, Prec, (int)P, &lwork) != 0) { bb_set_error("cusolverDnDpotrf_bufferSize failed"); rc = BB_ERR_STATE; }
    if (rc == BB_OK) {
        cudaError_t e = cudaMalloc((void**)&work, (size_t)(lwork > 0 ? lwork : 1) * sizeof(double));
        if (e != cudaSuccess) { bb_set_error("cholesky workspace: %s", cudaGetErrorString(e)); rc = BB_ERR_CUDA; }
    }
    if (rc == BB_OK && g_cs.potrf(g_cs.handle, 1, (int)P, Prec, (int)P, work, lwork, info_dev) != 0) { bb_set_error("cusolverDnDpotrf failed"); rc = BB_ERR_STATE; }
    ctx->launches += 1;
    int info_host = 0;
    if (rc == BB_OK) {
        BB_CUDA(cudaEventRecord(ev[2], st));
        k_vec_mul<<<gP, 256, 0, st>>>(J, m->z, P, rhs);                 // J z
        k_trsv_upper<1><<<1, 1024, 0, st>>>(Prec, P, rhs);              // U' y = J z
        k_trsv_upper<0><<<1, 1024, 0, st>>>(Prec, P, rhs);              // U mean = y
        BB_CUDA(cudaMemcpyAsync(xg, m->eps_P, Pb, cudaMemcpyDeviceToDevice, st));
        k_trsv_upper<0><<<1, 1024, 0, st>>>(Prec, P, xg);               // U x = g
        k_chol_finish<<<gP, 256, 0, st>>>(J, rhs, xg, P, m->out_P);
        ctx->launches += 5;
        BB_CUDA(cudaMemcpyAsync(coef_out, m->out_P, Pb, cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaMemcpyAsync(&info_host, info_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    timer_.end();
    cudaError_t es = cudaStreamSynchronize(st);
    timer_.commit();
    if (stats && rc == BB_OK) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]);
        stats[0] = a; stats[1] = b;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (work) cudaFree(work);
    cudaFree(Prec); cudaFree(info_dev);
    fisher_free(&fw);
    if (rc != BB_OK) return rc;
    if (es != cudaSuccess) { bb_set_error("cholesky sampler: %s", cudaGetErrorString(es)); return BB_ERR_CUDA; }
    if (info_host != 0) { bb_set_error("potrf: the posterior precision is not positive definite (info = %d)", info_host); return BB_ERR_ARG; }
    return BB_OK;
}
