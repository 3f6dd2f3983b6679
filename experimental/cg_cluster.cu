// PROTOTYPE + micro-benchmark: the three P-vector kernels of one CG iteration (k_cg_q, k_cg_update, k_cg_dir in
// bb_cg.cu) fused into ONE kernel that runs as a single thread-block cluster, with the two dot products reduced
// through distributed shared memory and cluster barriers instead of kernel boundaries.
//
// STATUS: compiles for sm_100a; HAS NOT RUN ON A GPU YET.  Not built into libbbgpu.so.
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/cgc experimental/cg_cluster.cu
//     /tmp/cgc [P = 100001] [iterations = 300]
//
// Why: at the per-rank size of the 8-GPU run a CG iteration spends ~40 us in seven small kernels and the gaps between
// them, as much as in the two SpMVs (profiles/r01_launches.md).  q = D.p + s.t, alpha = rho / p.q, x += alpha p,
// r -= alpha q, rho' = r.r, p = r + (rho'/rho) p, sv = s.p are three kernels today because p.q and r.r are grid-wide
// reductions.  A cluster of 16 CTAs (non-portable size, one GPC) covers P = 100k with 6 250 entries per CTA; p, q and r
// of the chunk stay in shared memory between the phases (150 KB), every CTA writes its partial sum into every CTA's
// shared memory (DSMEM), and after a cluster barrier (~0.2 us) all CTAs add the 16 partials in rank order -- same
// bits everywhere, deterministic.  Rounding of the element-wise updates is the library's (__dmul_rn/__dadd_rn, numpy's
// operation order); only the grouping of the two reductions differs from the three-kernel form.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

typedef long long i64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

struct Scalars { double rho[2]; double atol; double rnorm; int iter, done, maxiter, pad; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum(double v, double* sm33) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm33[w] = v;
    __syncthreads();
    if (w == 0) { double t = lane < nw ? sm33[lane] : 0.0; t = warp_sum(t); if (lane == 0) sm33[32] = t; }
    __syncthreads();
    return sm33[32];
}
__device__ __forceinline__ double partial_sum(const double* buf, int count) {     // all 32 lanes of a warp
    double t = 0.0;
    for (int i = threadIdx.x & 31; i < count; i += 32) t += buf[i];
    return warp_sum(t);
}
__device__ __forceinline__ double tdot_entry(const double* traw, const double* c, i64 j, int icpt) {
    return j < icpt ? traw[0] : __dsub_rn(traw[1 + (j - icpt)], __dmul_rn(traw[0], c[j - icpt]));
}

// ---- the three-kernel form (as in bb_cg.cu, without the exchange and the graph plumbing) ----------------------------
__global__ void k_q(Scalars* st, const double* traw, const double* c, int icpt, i64 P, const double* pvec, const double* s,
                    const double* D, double* q, double* red_pq) {
    if (st->done) return;
    __shared__ double sm[33];
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        const double pj = pvec[j], qj = __dadd_rn(__dmul_rn(D[j], pj), __dmul_rn(s[j], tdot_entry(traw, c, j, icpt)));
        q[j] = qj; acc += pj * qj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) { red_pq[blockIdx.x] = acc; if (blockIdx.x == 0) st->iter = st->iter + 1; }
}
__global__ void k_update(const Scalars* st, const double* red_pq, int nred, i64 P, double* x, double* r, const double* pvec,
                         const double* q, double* red_rr) {
    if (st->done) return;
    __shared__ double sm[33];
    const int it = st->iter - 1;
    const double alpha = st->rho[it & 1] / partial_sum(red_pq, nred);
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        x[j] = __dadd_rn(x[j], __dmul_rn(alpha, pvec[j]));
        const double rj = __dsub_rn(r[j], __dmul_rn(alpha, q[j]));
        r[j] = rj; acc += rj * rj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_rr[blockIdx.x] = acc;
}
__global__ void k_dir(Scalars* st, const double* red_rr, int nred, i64 P, int icpt, const double* r, double* pvec,
                      const double* s, const double* c, double* sv, double* red_shift) {
    if (st->done) return;
    __shared__ double sm[33];
    const int it = st->iter;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    const double rho = partial_sum(red_rr, nred), rn = sqrt(rho);
    if (it >= st->maxiter) { if (lead) { st->done = 2; st->rnorm = rn; } return; }
    if (rn < st->atol) { if (lead) { st->done = 1; st->rnorm = rn; } return; }
    const double beta = it > 0 ? rho / st->rho[(it + 1) & 1] : 0.0;
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        const double pj = it > 0 ? __dadd_rn(__dmul_rn(beta, pvec[j]), r[j]) : r[j];
        pvec[j] = pj;
        const double xs = __dmul_rn(s[j], pj);
        sv[j] = xs; acc += j < icpt ? xs : -c[j - icpt] * xs;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_shift[blockIdx.x] = acc;
    if (lead) { st->rho[it & 1] = rho; st->rnorm = rn; }
}

// ---- fused: one cluster ---------------------------------------------------------------------------------------
constexpr int FUSED_THREADS = 1024;
// every CTA deposits `v` into slot [rank] of `red` in every CTA of the cluster; after the barrier all CTAs add the
// partials in rank order
__device__ __forceinline__ double cluster_sum(cg::cluster_group& cluster, double v, double* red /* shared, >= cluster size */) {
    const unsigned n = cluster.num_blocks(), me = cluster.block_rank();
    if (threadIdx.x < n) cluster.map_shared_rank(red, threadIdx.x)[me] = v;
    cluster.sync();
    double t = 0.0;
    for (unsigned i = 0; i < n; ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(FUSED_THREADS, 1)
k_cg_fused(Scalars* st, const double* __restrict__ traw, const double* __restrict__ c, int icpt, i64 P, int chunk,
           double* __restrict__ pvec, const double* __restrict__ s, const double* __restrict__ D, double* __restrict__ x,
           double* __restrict__ r, double* __restrict__ sv, double* __restrict__ red_shift) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double smem[];
    double* sp = smem;                   // p, q, r of this CTA's chunk
    double* sq = smem + chunk;
    double* sr = smem + 2 * chunk;
    __shared__ double sm[33];
    __shared__ double red_a[16], red_b[16];
    // NB: every CTA of the cluster must reach every cluster barrier, so `done` is folded into the work instead of an
    // early return: a finished solve runs the barriers with empty loops.
    const bool active = st->done == 0;
    const int me = (int)cluster.block_rank();
    const i64 lo = (i64)me * chunk, hi = active ? min(P, lo + chunk) : lo;
    const int it0 = st->iter;            // iterations completed before this call
    const double rho = st->rho[it0 & 1];
    // phase 1: q = D.p + s.t, p.q
    double acc = 0.0;
    for (i64 j = lo + threadIdx.x; j < hi; j += FUSED_THREADS) {
        const double pj = pvec[j], qj = __dadd_rn(__dmul_rn(D[j], pj), __dmul_rn(s[j], tdot_entry(traw, c, j, icpt)));
        sp[j - lo] = pj; sq[j - lo] = qj; acc += pj * qj;
    }
    const double pq = cluster_sum(cluster, block_sum(acc, sm), red_a);
    // phase 2: x += alpha p, r -= alpha q, r.r
    const double alpha = rho / pq;
    acc = 0.0;
    for (i64 j = lo + threadIdx.x; j < hi; j += FUSED_THREADS) {
        x[j] = __dadd_rn(x[j], __dmul_rn(alpha, sp[j - lo]));
        const double rj = __dsub_rn(r[j], __dmul_rn(alpha, sq[j - lo]));
        r[j] = rj; sr[j - lo] = rj; acc += rj * rj;
    }
    const double rho_new = cluster_sum(cluster, block_sum(acc, sm), red_b);
    // phase 3 (head of the next iteration): convergence test, p = r + beta p, sv = s.p, shift partials
    const int it = it0 + 1;
    const double rn = sqrt(rho_new);
    const bool lead = me == 0 && threadIdx.x == 0;
    bool stop = !active;
    if (active && it >= st->maxiter) { if (lead) { st->done = 2; st->rnorm = rn; st->iter = it; } stop = true; }
    else if (active && rn < st->atol) { if (lead) { st->done = 1; st->rnorm = rn; st->iter = it; } stop = true; }
    if (!stop) {
        const double beta = rho_new / rho;
        acc = 0.0;
        for (i64 j = lo + threadIdx.x; j < hi; j += FUSED_THREADS) {
            const double pj = __dadd_rn(__dmul_rn(beta, sp[j - lo]), sr[j - lo]);
            pvec[j] = pj;
            const double xs = __dmul_rn(s[j], pj);
            sv[j] = xs; acc += j < icpt ? xs : -c[j - icpt] * xs;
        }
        acc = block_sum(acc, sm);
        if (threadIdx.x == 0) red_shift[me] = acc;
        if (lead) { st->rho[it & 1] = rho_new; st->rnorm = rn; st->iter = it; }
    }
    cluster.sync();                      // no CTA exits while a peer may still write into its shared memory
}

int main(int argc, char** argv) {
    const i64 P = argc > 1 ? atoll(argv[1]) : 100001;
    const int iters = argc > 2 ? atoi(argv[2]) : 300;
    const int icpt = 1;
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    std::vector<double> h_t((size_t)P + 1), h_c((size_t)P), h_p((size_t)P), h_s((size_t)P), h_D((size_t)P), h_x((size_t)P), h_r((size_t)P);
    for (i64 j = 0; j < P; ++j) {
        h_t[(size_t)j] = std::sin(0.3 * j) * 1e-3; h_c[(size_t)j] = 0.01 * std::cos(0.7 * j); h_p[(size_t)j] = std::sin(1.1 * j + 0.2);
        h_s[(size_t)j] = 0.5 + 0.4 * std::cos(0.13 * j); h_D[(size_t)j] = 1.0 + 0.3 * std::sin(0.05 * j);
        h_x[(size_t)j] = std::cos(0.9 * j); h_r[(size_t)j] = 0.7 * std::sin(1.1 * j + 0.25);
    }
    h_t[(size_t)P] = 1e-3;
    auto up = [&](const std::vector<double>& h) { double* d; CK(cudaMalloc((void**)&d, h.size() * sizeof(double))); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice)); return d; };
    double *traw = up(h_t), *c = up(h_c), *s = up(h_s), *D = up(h_D);
    double *pA = up(h_p), *xA = up(h_x), *rA = up(h_r), *pB = up(h_p), *xB = up(h_x), *rB = up(h_r);
    double *q, *svA, *svB, *red;
    CK(cudaMalloc((void**)&q, P * sizeof(double))); CK(cudaMalloc((void**)&svA, P * sizeof(double))); CK(cudaMalloc((void**)&svB, P * sizeof(double)));
    CK(cudaMalloc((void**)&red, 4 * 128 * sizeof(double)));
    Scalars h_st = {{0.0, 0.0}, 0.0, 0.0, 0, 0, 1 << 30, 0};
    for (i64 j = 0; j < P; ++j) h_st.rho[0] += h_r[(size_t)j] * h_r[(size_t)j];      // rho of "iteration 0"
    Scalars *stA, *stB;
    CK(cudaMalloc((void**)&stA, sizeof(Scalars))); CK(cudaMalloc((void**)&stB, sizeof(Scalars)));
    CK(cudaMemcpy(stA, &h_st, sizeof(Scalars), cudaMemcpyHostToDevice)); CK(cudaMemcpy(stB, &h_st, sizeof(Scalars), cudaMemcpyHostToDevice));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const int nblk = (int)std::min<i64>(98, (P + 1023) / 1024);
    auto three = [&]() {
        k_q<<<nblk, 256, 0, st>>>(stA, traw, c, icpt, P, pA, s, D, q, red);
        k_update<<<nblk, 256, 0, st>>>(stA, red, nblk, P, xA, rA, pA, q, red + 128);
        k_dir<<<nblk, 256, 0, st>>>(stA, red + 128, nblk, P, icpt, rA, pA, s, c, svA, red + 256);
    };
    // fused launch: cluster of `csize` CTAs
    int csize = argc > 3 ? atoi(argv[3]) : 16;
    const int chunk = (int)(((P + csize - 1) / csize + 31) & ~31);
    const size_t smem = (size_t)3 * chunk * sizeof(double);
    if (smem + 2048 > prop.sharedMemPerBlockOptin) { printf("chunk of %d entries does not fit (cluster size %d)\n", chunk, csize); return 0; }
    CK(cudaFuncSetAttribute(k_cg_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8) CK(cudaFuncSetAttribute(k_cg_fused, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize); cfg.blockDim = dim3(FUSED_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    auto fused = [&]() { CK(cudaLaunchKernelEx(&cfg, k_cg_fused, stB, (const double*)traw, (const double*)c, icpt, P, chunk, pB, (const double*)s, (const double*)D, xB, rB, svB, red + 384)); };
    // one iteration each, compare
    three(); fused();
    CK(cudaStreamSynchronize(st));
    auto down = [&](const double* d) { std::vector<double> h((size_t)P); CK(cudaMemcpy(h.data(), d, P * sizeof(double), cudaMemcpyDeviceToHost)); return h; };
    double worst = 0.0;
    const double* pairs[4][2] = {{xA, xB}, {rA, rB}, {pA, pB}, {svA, svB}};
    for (auto& pr : pairs) {
        std::vector<double> a = down(pr[0]), b = down(pr[1]);
        double num = 0.0, den = 0.0;
        for (i64 j = 0; j < P; ++j) { num = std::max(num, std::fabs(a[(size_t)j] - b[(size_t)j])); den = std::max(den, std::fabs(a[(size_t)j])); }
        worst = std::max(worst, num / std::max(den, 1e-300));
    }
    Scalars a, b; CK(cudaMemcpy(&a, stA, sizeof(a), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&b, stB, sizeof(b), cudaMemcpyDeviceToHost));
    const bool ok = worst < 1e-12 && a.iter == b.iter && std::fabs(a.rho[1] - b.rho[1]) <= 1e-12 * std::fabs(a.rho[1]);
    printf("fused vs three kernels after one iteration: max rel diff %.2e, iter %d/%d, rho %.17g / %.17g -> %s\n", worst, a.iter, b.iter, a.rho[1], b.rho[1], ok ? "PASS" : "FAIL");
    // timing: a graph of `iters` iterations of each form (values are irrelevant for the timing)
    for (int mode = 0; mode < 2; ++mode) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        for (int i = 0; i < iters; ++i) { if (mode == 0) three(); else fused(); }
        CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(e0, st)); CK(cudaGraphLaunch(ge, st)); CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%s: %.2f us per CG iteration (P = %lld, %d iterations in one graph)\n", mode == 0 ? "three kernels (98 x 256)" : "one cluster kernel", ms * 1e3 / iters, P, iters);
        CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
    }
    return ok ? 0 : 1;
}
