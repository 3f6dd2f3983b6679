// Mode search of the chain initialisation, entirely on the device.
//
// Reference: reg_coef_sampler/reg_coef_sampler.py:281-358 (search_mode / define_function_for_optim): the conditional
// posterior mode of the coefficients is found with scipy's L-BFGS-B (no bounds; maxcor = 200, gtol = 1e-6/sqrt(P),
// maxiter = 250) in prior-preconditioned coordinates theta = coef / scale:
//     F(theta) = -loglik(scale . theta) + 1/2 sum_j prior_prec_j theta_j^2 ,   grad F = -scale . grad loglik + prior_prec . theta
// With the likelihood already evaluated on the device (bb_loglik_and_gradient) the remaining cost of that search is
// scipy's own limited-memory bookkeeping on the host: 1.4 s of the 1.6 s initialisation of the 1M x 100k problem
// (profiles/r02_chain_init_profile.log).  This file is the same algorithm without the host: an unbounded L-BFGS
// (two-loop recursion over up to `maxcor` correction pairs, the same curvature safeguard and stopping rules as L-BFGS-B:
// max-norm of the gradient <= gtol, relative decrease <= ftol, maxiter) with a strong-Wolfe line search (c1 = 1e-3,
// c2 = 0.9 as dcsrch is called by L-BFGS-B; first trial step 1/||d|| at the first iteration, 1 afterwards).  The search
// direction is ONE kernel per iteration: a small co-resident grid walks the two loops with a grid barrier per
// correction pair, each thread keeping its slice of the vector.  Being a different implementation of the same method it
// does not reproduce scipy's iterates, only its optimum (to the tolerances above); the chain initialisation is not on
// the parity path (the Gibbs draws that follow are random), and `init_optimizer='scipy'` keeps the host optimiser.
#include "bb_internal.cuh"
#include <stdlib.h>
#include <vector>

constexpr int LB_THREADS = 512;
constexpr int LB_MAX_CTAS = 64;
constexpr int LB_MAX_COR = 256;

struct LbfgsWork {
    i64 P; int maxcor;
    double *theta, *g, *d, *theta_new, *g_new, *scale, *pprec;     // [P]
    double *S, *Y;                                                 // [maxcor][P] ring of correction pairs
    double *rho;                                                   // [maxcor]
    double *red;                                                   // [LB_MAX_CTAS * 4] partial sums
    double *scal;                                                  // [8] device scalars
    double *scal_host;                                             // pinned
    unsigned long long* bar;
};

__device__ __forceinline__ unsigned long long lb_ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lb_grid_barrier(unsigned long long* ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1ull);
        while (lb_ld_acquire(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ double lb_sum_partials(const double* buf, int count) {
    const int lane = threadIdx.x & 31;
    double t = 0.0;
    for (int i = lane; i < count; i += 32) t += __ldcg(buf + i);
    return warp_sum(t);
}

// d = -H g by the two-loop recursion (Nocedal & Wright alg. 7.4); the pairs are stored oldest..newest at ring positions
// (head - k + i) mod maxcor.  Each step is "dot over the whole vector -> axpy", so there is one grid barrier per pair and
// loop; the axpy of one step and the partial dot of the next are fused in one pass over the thread's elements.
__global__ void __launch_bounds__(LB_THREADS)
k_lbfgs_direction(i64 P, int k, int head, int maxcor, double gamma, const double* __restrict__ g,
                  const double* __restrict__ S, const double* __restrict__ Y, const double* __restrict__ rho,
                  double* __restrict__ q /* work + output d */, double* __restrict__ red, unsigned long long* bar,
                  unsigned long long bar_base) {
    __shared__ double sm[33];
    __shared__ double alpha[LB_MAX_COR];
    const int G = gridDim.x, tid = threadIdx.x;
    const i64 gtid = (i64)blockIdx.x * LB_THREADS + tid, gthreads = (i64)G * LB_THREADS;
    unsigned long long target = bar_base;
    // q = g ; partial of s_{k-1} . q
    double acc = 0.0;
    {
        const double* s_new = (k > 0) ? S + (i64)((head - 1 + maxcor) % maxcor) * P : nullptr;
        for (i64 j = gtid; j < P; j += gthreads) {
            const double v = g[j];
            q[j] = v;
            if (s_new) acc += s_new[j] * v;
        }
    }
    for (int t = 0; t < k; ++t) {                       // newest -> oldest
        const int pos = (head - 1 - t + 2 * maxcor) % maxcor;
        acc = block_sum(acc, sm);
        if (tid == 0) red[(t & 1) * LB_MAX_CTAS + blockIdx.x] = acc;
        target += (unsigned long long)G;
        lb_grid_barrier(bar, target);
        const double a = rho[pos] * lb_sum_partials(red + (t & 1) * LB_MAX_CTAS, G);
        if (tid == 0) alpha[t] = a;
        const double* y = Y + (i64)pos * P;
        const double* s_next = (t + 1 < k) ? S + (i64)((pos - 1 + maxcor) % maxcor) * P : nullptr;
        acc = 0.0;
        for (i64 j = gtid; j < P; j += gthreads) {
            const double v = q[j] - a * y[j];
            q[j] = v;
            if (s_next) acc += s_next[j] * v;
        }
    }
    __syncthreads();
    // r = gamma q ; partial of y_oldest . r
    {
        const double* y_old = (k > 0) ? Y + (i64)((head - k + 2 * maxcor) % maxcor) * P : nullptr;
        acc = 0.0;
        for (i64 j = gtid; j < P; j += gthreads) {
            const double v = gamma * q[j];
            q[j] = v;
            if (y_old) acc += y_old[j] * v;
        }
    }
    for (int t = k - 1; t >= 0; --t) {                  // oldest -> newest (t indexes alpha: t = k-1 is the oldest pair)
        const int pos = (head - 1 - t + 2 * maxcor) % maxcor;
        acc = block_sum(acc, sm);
        if (tid == 0) red[(2 + (t & 1)) * LB_MAX_CTAS + blockIdx.x] = acc;
        target += (unsigned long long)G;
        lb_grid_barrier(bar, target);
        const double b = rho[pos] * lb_sum_partials(red + (2 + (t & 1)) * LB_MAX_CTAS, G);
        const double coef = alpha[t] - b;
        const double* s = S + (i64)pos * P;
        const double* y_next = (t > 0) ? Y + (i64)((pos + 1) % maxcor) * P : nullptr;
        acc = 0.0;
        for (i64 j = gtid; j < P; j += gthreads) {
            const double v = q[j] + coef * s[j];
            q[j] = v;
            if (y_next) acc += y_next[j] * v;
        }
    }
    for (i64 j = gtid; j < P; j += gthreads) q[j] = -q[j];
}

// v = scale . (theta + step d)   (the coefficient vector the likelihood is evaluated at); also theta_new
__global__ void k_lbfgs_point(i64 P, const double* __restrict__ theta, const double* __restrict__ d, double step,
                              const double* __restrict__ scale, double* __restrict__ theta_new, double* __restrict__ coef) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        const double t = d ? theta[j] + step * d[j] : theta[j];
        theta_new[j] = t;
        coef[j] = scale[j] * t;
    }
}

// g = -scale . grad_loglik + prior_prec . theta ; partial sums: [0] 1/2 sum prior_prec theta^2, [1] g . d, [2] g . g ; max |g|
__global__ void __launch_bounds__(256)
k_lbfgs_grad(i64 P, const double* __restrict__ theta, const double* __restrict__ grad_ll, const double* __restrict__ scale,
             const double* __restrict__ pprec, const double* __restrict__ d, double* __restrict__ g, double* __restrict__ red /*[4][grid]*/) {
    __shared__ double sm[33];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        const double th = theta[j];
        const double gj = -scale[j] * grad_ll[j] + pprec[j] * th;
        g[j] = gj;
        a0 += 0.5 * pprec[j] * th * th;
        if (d) a1 += gj * d[j];
        a2 += gj * gj;
        a3 = fmax(a3, fabs(gj));
    }
    a0 = block_sum(a0, sm); a1 = block_sum(a1, sm); a2 = block_sum(a2, sm);
    // block max
    for (int o = 16; o > 0; o >>= 1) a3 = fmax(a3, __shfl_xor_sync(0xffffffffu, a3, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a3;
    __syncthreads();
    if (threadIdx.x == 0) {
        double mx = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, sm[w]);
        red[0 * gridDim.x + blockIdx.x] = a0; red[1 * gridDim.x + blockIdx.x] = a1;
        red[2 * gridDim.x + blockIdx.x] = a2; red[3 * gridDim.x + blockIdx.x] = mx;
    }
}
// scal[0] = F = -ll + quad ; scal[1] = g.d ; scal[2] = ||g||_2^2 ; scal[3] = ||g||_inf      (ll in ll_dev[0], already all-reduced)
__global__ void k_lbfgs_scalars(const double* __restrict__ red, int nred, const double* __restrict__ ll_dev, int is_linear,
                                double prec, double* __restrict__ scal) {
    const double q = warp_sum_partials(red, nred), gd = warp_sum_partials(red + nred, nred), gg = warp_sum_partials(red + 2 * nred, nred);
    double mx = 0.0;
    for (int i = threadIdx.x & 31; i < nred; i += 32) mx = fmax(mx, red[3 * nred + i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) {
        const double ll = is_linear ? -0.5 * prec * ll_dev[0] : ll_dev[0];
        scal[0] = -ll + q; scal[1] = gd; scal[2] = gg; scal[3] = mx;
    }
}
// correction pair: s = theta_new - theta, y = g_new - g into ring slot; partial s.y, y.y
__global__ void __launch_bounds__(256)
k_lbfgs_pair(i64 P, const double* __restrict__ theta, const double* __restrict__ theta_new, const double* __restrict__ g,
             const double* __restrict__ g_new, double* __restrict__ s_out, double* __restrict__ y_out, double* __restrict__ red) {
    __shared__ double sm[33];
    double sy = 0.0, yy = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        const double s = theta_new[j] - theta[j], y = g_new[j] - g[j];
        s_out[j] = s; y_out[j] = y;
        sy += s * y; yy += y * y;
    }
    sy = block_sum(sy, sm); yy = block_sum(yy, sm);
    if (threadIdx.x == 0) { red[blockIdx.x] = sy; red[gridDim.x + blockIdx.x] = yy; }
}
__global__ void k_lbfgs_pair_scalars(const double* __restrict__ red, int nred, double* __restrict__ scal) {
    const double sy = warp_sum_partials(red, nred), yy = warp_sum_partials(red + nred, nred);
    if (threadIdx.x == 0) { scal[4] = sy; scal[5] = yy; }
}
__global__ void k_lbfgs_set(double* p, double v) { p[0] = v; }

// defined in bb_rand.cu
int bb_loglik_resid_dev(bb_mat* m, const double* coef_dev, double obs_prec, double* ll_dev, double* grad_dev);

static void lb_free(LbfgsWork* w) {
    void* ptrs[] = {w->theta, w->g, w->d, w->theta_new, w->g_new, w->scale, w->pprec, w->S, w->Y, w->rho, w->red, w->scal, w->bar};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (w->scal_host) cudaFreeHost(w->scal_host);
}

static int lb_grid(bb_ctx* ctx, i64 P) {
    i64 g = (P + 2047) / 2048;
    if (g > LB_MAX_CTAS) g = LB_MAX_CTAS;
    if (g > ctx->sm_count) g = ctx->sm_count;
    if (g < 1) g = 1;
    return (int)g;
}

// status: 0 converged (gradient), 1 converged (relative decrease), 2 maxiter reached, 3 line search failed
extern "C" int bb_mode_search(bb_mat* m, const double* coef0, const double* scale, const double* prior_prec, double obs_prec,
                              int maxiter, double gtol, double ftol, int maxcor, double* coef_out,
                              int* n_iter_out, int* n_eval_out, int* status_out) {
    BB_ARG(m && coef0 && scale && prior_prec && coef_out, "null pointer");
    BB_ARG(maxiter >= 0 && maxcor >= 1 && maxcor <= LB_MAX_COR, "maxiter / maxcor (1..256)");
    if (!m->has_outcome) { bb_set_error("bb_mode_search needs bb_set_outcome"); return BB_ERR_STATE; }
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    const i64 P = m->P;
    if (maxcor > maxiter + 1) maxcor = maxiter + 1;
    LbfgsWork w;
    memset(&w, 0, sizeof(w));
    w.P = P; w.maxcor = maxcor;
    int rc = BB_OK;
#define LBALLOC(ptr, count) if (rc == BB_OK && cudaMalloc((void**)&(ptr), (size_t)(count) * sizeof(*(ptr))) != cudaSuccess) { bb_set_error("bb_mode_search: out of device memory"); rc = BB_ERR_CUDA; }
    LBALLOC(w.theta, P + 1); LBALLOC(w.g, P + 1); LBALLOC(w.d, P + 1); LBALLOC(w.theta_new, P + 1); LBALLOC(w.g_new, P + 1);
    LBALLOC(w.scale, P + 1); LBALLOC(w.pprec, P + 1);
    LBALLOC(w.S, (i64)maxcor * P + 1); LBALLOC(w.Y, (i64)maxcor * P + 1); LBALLOC(w.rho, maxcor);
    LBALLOC(w.red, 4 * 1024); LBALLOC(w.scal, 8); LBALLOC(w.bar, 1);
#undef LBALLOC
    if (rc == BB_OK && cudaMallocHost((void**)&w.scal_host, 8 * sizeof(double)) != cudaSuccess) { bb_set_error("bb_mode_search: pinned alloc"); rc = BB_ERR_CUDA; }
    if (rc != BB_OK) { lb_free(&w); return rc; }
    const size_t Pb = (size_t)P * sizeof(double);
    const int gv = (int)((P + 1023) / 1024 > 0 ? ((P + 1023) / 1024 < 1024 ? (P + 1023) / 1024 : 1024) : 1);   // grid of the vector kernels
    const int G = lb_grid(ctx, P);
    unsigned long long bar_base = 0;
    int n_eval = 0, n_iter = 0, status = 2;
    std::vector<double> rho_host((size_t)maxcor, 0.0);

    auto fail = [&](int code) { lb_free(&w); return code; };
#define LB_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { bb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); return fail(BB_ERR_CUDA); } } while (0)
#define LB_TRY(expr) do { int rc_ = (expr); if (rc_ != BB_OK) return fail(rc_); } while (0)
    LB_CUDA(cudaMemsetAsync(w.bar, 0, sizeof(unsigned long long), st));
    LB_CUDA(cudaMemcpyAsync(w.scale, scale, Pb, cudaMemcpyHostToDevice, st));
    LB_CUDA(cudaMemcpyAsync(w.pprec, prior_prec, Pb, cudaMemcpyHostToDevice, st));
    // theta = coef0 / scale (P doubles, on the host side of the data)
    {
        std::vector<double> th((size_t)P);
        for (i64 j = 0; j < P; ++j) th[(size_t)j] = coef0[j] / scale[j];
        LB_CUDA(cudaMemcpyAsync(w.theta, th.data(), Pb, cudaMemcpyHostToDevice, st));
        LB_CUDA(cudaStreamSynchronize(st));
    }
    // F, g at (theta + step d); results in theta_new / g_new and scal_host[0..3]
    auto evaluate = [&](const double* d_dev, double step) -> int {
        k_lbfgs_point<<<gv, 256, 0, st>>>(P, w.theta, d_dev, step, w.scale, w.theta_new, m->v_P);
        ctx->launches++;
        BB_TRY(bb_loglik_resid_dev(m, m->v_P, obs_prec, w.scal + 6, m->t_P));
        k_lbfgs_grad<<<gv, 256, 0, st>>>(P, w.theta_new, m->t_P, w.scale, w.pprec, d_dev, w.g_new, w.red);
        k_lbfgs_scalars<<<1, 32, 0, st>>>(w.red, gv, w.scal + 6, m->is_linear, obs_prec, w.scal);
        ctx->launches += 2;
        BB_CUDA(cudaMemcpyAsync(w.scal_host, w.scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        ++n_eval;
        return BB_OK;
    };
    LB_TRY(evaluate(nullptr, 0.0));
    double f = w.scal_host[0], ginf = w.scal_host[3], gnorm2 = w.scal_host[2];
    LB_CUDA(cudaMemcpyAsync(w.g, w.g_new, Pb, cudaMemcpyDeviceToDevice, st));
    int k = 0, head = 0;
    double gamma = 1.0;
    if (!(ginf > gtol)) status = 0;
    const double epsmch = 2.220446049250313e-16;
    while (status == 2 && n_iter < maxiter) {
        // search direction
        k_lbfgs_direction<<<G, LB_THREADS, 0, st>>>(P, k, head, maxcor, gamma, w.g, w.S, w.Y, w.rho, w.d, w.red, w.bar, bar_base);
        ctx->launches++;
        bar_base += (unsigned long long)G * (unsigned long long)(2 * k);
        // line search along d (strong Wolfe, c1 = 1e-3, c2 = 0.9): bracket, then bisect / interpolate
        // g.d at the current point: from a zero-step evaluation of the stored g? cheaper: one tiny reduction
        k_lbfgs_grad<<<gv, 256, 0, st>>>(P, w.theta, m->t_P, w.scale, w.pprec, w.d, w.g_new, w.red);   // recomputes g from the stored grad_ll (t_P still holds it)
        k_lbfgs_scalars<<<1, 32, 0, st>>>(w.red, gv, w.scal + 6, m->is_linear, obs_prec, w.scal);
        ctx->launches += 2;
        LB_CUDA(cudaMemcpyAsync(w.scal_host, w.scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
        LB_CUDA(cudaStreamSynchronize(st));
        const double gd0 = w.scal_host[1];
        if (!(gd0 < 0.0)) {                                  // not a descent direction: drop the history and retry with steepest descent
            if (k == 0) { status = 3; break; }
            k = 0; gamma = 1.0;
            continue;
        }
        const double c1 = 1e-3, c2 = 0.9;
        double step = (n_iter == 0) ? fmin(1.0, 1.0 / sqrt(gnorm2)) : 1.0;
        double lo = 0.0, hi = 0.0, f_lo = f;
        bool have_hi = false, ok = false;
        double f_new = f, gd_new = gd0;
        for (int ls = 0; ls < 25; ++ls) {
            LB_TRY(evaluate(w.d, step));
            f_new = w.scal_host[0]; gd_new = w.scal_host[1];
            if (!(f_new <= f + c1 * step * gd0) || (ls > 0 && !have_hi && f_new >= f_lo && lo > 0.0)) {
                hi = step; have_hi = true;
            } else if (fabs(gd_new) <= -c2 * gd0) {
                ok = true; break;
            } else if (gd_new >= 0.0) {
                hi = step; have_hi = true;
            } else {
                lo = step; f_lo = f_new;
            }
            step = have_hi ? 0.5 * (lo + hi) : 2.0 * step;
            if (have_hi && (hi - lo) <= 1e-14 * hi) break;
        }
        if (!ok) {
            // accept a point that at least satisfies sufficient decrease; otherwise give up
            if (!(f_new <= f + c1 * step * gd0)) { status = 3; break; }
        }
        // accept: correction pair, then move
        const int pos = head;
        k_lbfgs_pair<<<gv, 256, 0, st>>>(P, w.theta, w.theta_new, w.g, w.g_new, w.S + (i64)pos * P, w.Y + (i64)pos * P, w.red);
        k_lbfgs_pair_scalars<<<1, 32, 0, st>>>(w.red, gv, w.scal);
        ctx->launches += 2;
        LB_CUDA(cudaMemcpyAsync(w.scal_host + 4, w.scal + 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        LB_CUDA(cudaMemcpyAsync(w.theta, w.theta_new, Pb, cudaMemcpyDeviceToDevice, st));
        LB_CUDA(cudaMemcpyAsync(w.g, w.g_new, Pb, cudaMemcpyDeviceToDevice, st));
        LB_CUDA(cudaStreamSynchronize(st));
        const double sy = w.scal_host[4], yy = w.scal_host[5];
        if (sy > epsmch * yy && yy > 0.0) {                  // L-BFGS-B's curvature safeguard (lbfgsb.f: dr <= epsmch*ddum -> skip)
            k_lbfgs_set<<<1, 1, 0, st>>>(w.rho + pos, 1.0 / sy);
            ctx->launches++;
            gamma = sy / yy;
            head = (head + 1) % maxcor;
            if (k < maxcor) ++k;
        }
        ++n_iter;
        const double f_old = f;
        f = f_new; ginf = w.scal_host[3]; gnorm2 = w.scal_host[2];
        if (!(ginf > gtol)) { status = 0; break; }
        if ((f_old - f) <= ftol * fmax(fmax(fabs(f_old), fabs(f)), 1.0)) { status = 1; break; }
    }
    // coef = scale . theta
    k_lbfgs_point<<<gv, 256, 0, st>>>(P, w.theta, nullptr, 0.0, w.scale, w.theta_new, m->v_P);
    ctx->launches++;
    LB_CUDA(cudaMemcpyAsync(coef_out, m->v_P, Pb, cudaMemcpyDeviceToHost, st));
    timer_.end();
    LB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
#undef LB_CUDA
#undef LB_TRY
    lb_free(&w);
    if (n_iter_out) *n_iter_out = n_iter;
    if (n_eval_out) *n_eval_out = n_eval;
    if (status_out) *status_out = status;
    return BB_OK;
}
