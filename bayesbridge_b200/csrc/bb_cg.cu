// The prior-preconditioned conjugate-gradient sampler (reference: reg_coef_sampler/cg_sampler.py:20-151)
// and the CG loop it delegates to scipy.sparse.linalg.cg (scipy 1.18.1, _isolve/iterative.py:cg).
// The recurrences below follow scipy's update order exactly:
//     r = b - A x0
//     loop: if ||r|| < atol: stop ; rho = r.r ; p = r + (rho/rho_prev) p ; q = A p ;
//           alpha = rho / (p.q) ; x += alpha p ; r -= alpha q
// with A(x) = D.x + s.X'(omega.(X(s.x))), D = (s.prior_prec_sqrt)^2  (cg_sampler.py:105-112).
// All vector state stays on the device; the host only reads back {iter, done} between chunks.
#include "bb_internal.cuh"

__device__ __forceinline__ double tdot_entry(const double* __restrict__ traw, const double* __restrict__ c, i64 j, int icpt) {
    const double sw = traw[0];
    if (j < icpt) return sw;
    // sparse_matrix.py:126-128   result = X.T.dot(v); result -= sum(v) * column_offset
    return __dsub_rn(traw[1 + (j - icpt)], __dmul_rn(sw, c[j - icpt]));
}

// u_n = sqrt(omega) .* eps1  (cg_sampler.py:66), eps1 injected or Philox
__global__ void k_rhs_noise(const double* __restrict__ omega, double omega_scalar, const double* __restrict__ eps1,
                            int philox, uint64_t seed, uint64_t offset, i64 row_offset, i64 n, double* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double e;
        if (philox) { RandStream rs; rs.init(seed, offset, (uint64_t)(row_offset + i), STREAM_EPS1); e = rs.normal(); }
        else e = eps1[i];
        double om = omega ? omega[i] : omega_scalar;
        out[i] = sqrt(om) * e;
    }
}

// y_gaussian weighting for z = X'(omega .* y): logit -> kappa ; linear -> omega * y
__global__ void k_z_weight(const double* __restrict__ n_trial, const double* __restrict__ n_success, int is_linear,
                           const double* __restrict__ omega, double omega_scalar, i64 n, double* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        if (is_linear) out[i] = (omega ? omega[i] : omega_scalar) * n_success[i];
        else out[i] = n_success[i] - 0.5 * n_trial[i];
    }
}

// b = s.(z + v), v = Tdot(sqrt(omega) eps1) + pps.eps2 ; D = (s.pps)^2 ; x = x0 / s ; partial b.b
__global__ void k_cg_init(const double* __restrict__ traw, const double* __restrict__ c, int icpt, i64 P,
                          const double* __restrict__ z, const double* __restrict__ pps, const double* __restrict__ s,
                          const double* __restrict__ x0, const double* __restrict__ eps2,
                          int philox, uint64_t seed, uint64_t offset,
                          double* __restrict__ b, double* __restrict__ D, double* __restrict__ x, double* __restrict__ red_bb) {
    __shared__ double sm[33];
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        double e2;
        if (philox) { RandStream rs; rs.init(seed, offset, (uint64_t)j, STREAM_EPS2); e2 = rs.normal(); }
        else e2 = eps2[j];
        double t = tdot_entry(traw, c, j, icpt);
        double v = __dadd_rn(t, __dmul_rn(pps[j], e2));
        double bj = __dmul_rn(s[j], __dadd_rn(z[j], v));
        b[j] = bj;
        double sp = __dmul_rn(s[j], pps[j]);
        D[j] = __dmul_rn(sp, sp);
        x[j] = x0[j] / s[j];
        acc += bj * bj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_bb[blockIdx.x] = acc;
}

__global__ void k_cg_scalars(CgScalars* st, const double* __restrict__ red_bb, int nred, double atol, int maxiter) {
    double bb = warp_sum_partials(red_bb, nred);
    if (threadIdx.x == 0) {
        double bn = sqrt(bb);
        st->bnorm = bn;
        // cg_sampler.py:75 rtol = atol/||b|| ; scipy: atol_eff = max(0, rtol*||b||)
        st->atol_eff = (bn > 0.0) ? (atol / bn) * bn : atol;
        st->iter = 0;
        st->done = (bn == 0.0) ? 3 : 0;
        st->maxiter = maxiter;
        st->rho[0] = 0.0; st->rho[1] = 0.0;
        st->rnorm = 0.0;
    }
}

// q = D.p + s.Tdot ; partial p.q ; bumps the iteration counter when count_iter.
// With a peer-memory exchange (pub != nullptr) the all-reduce of [sum w; X'w] is done HERE: the block waits for
// every rank's publication and each thread sums the ranks' partials in rank order (compute + collective fused).
__global__ void k_cg_q(CgScalars* st, const double* __restrict__ traw, const double* __restrict__ c, int icpt, i64 P,
                       const double* __restrict__ pvec, const double* __restrict__ s, const double* __restrict__ D,
                       double* __restrict__ q, double* __restrict__ red_pq, int count_iter,
                       const P2PView* __restrict__ pub_ptr) {
    if (st->done) return;
    __shared__ double sm[33];
    __shared__ int p2p_ok;
    P2PView pub;
    double sw = 0.0;
    if (pub_ptr != nullptr) {
        pub = *pub_ptr;
        if (!p2p_wait_all(pub, &p2p_ok)) return;
        sw = p2p_sum(pub, 0);
    }
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        double t;
        if (pub_ptr != nullptr) {
            t = (j < icpt) ? sw : __dsub_rn(p2p_sum(pub, 1 + (j - icpt)), __dmul_rn(sw, c[j - icpt]));
        } else {
            t = tdot_entry(traw, c, j, icpt);
        }
        double pj = pvec[j];
        double qj = __dadd_rn(__dmul_rn(D[j], pj), __dmul_rn(s[j], t));
        q[j] = qj;
        acc += pj * qj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) {
        red_pq[blockIdx.x] = acc;
        if (blockIdx.x == 0 && count_iter) st->iter = st->iter + 1;
    }
}

// r = b - q ; partial r.r      (initial residual)
__global__ void k_cg_resid(const CgScalars* st, const double* __restrict__ b, const double* __restrict__ q, i64 P,
                           double* __restrict__ r, double* __restrict__ red_rr) {
    if (st->done) return;
    __shared__ double sm[33];
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        double rj = __dsub_rn(b[j], q[j]);
        r[j] = rj;
        acc += rj * rj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_rr[blockIdx.x] = acc;
}

// convergence test, rho, search direction, and the scaled gather vector of the next product
__global__ void k_cg_dir(CgScalars* st, const double* __restrict__ red_rr, int nred, i64 P, int icpt,
                         const double* __restrict__ r, double* __restrict__ pvec, const double* __restrict__ s,
                         const double* __restrict__ c, double* __restrict__ sv, double* __restrict__ red_shift) {
    if (st->done) return;
    __shared__ double sm[33];
    const int it = st->iter;
    const bool lead = (blockIdx.x == 0 && threadIdx.x == 0);
    const double rho = warp_sum_partials(red_rr, nred);
    const double rn = sqrt(rho);
    // scipy: `for iteration in range(maxiter): if norm(r) < atol: return x, 0 ...` -- after maxiter
    // updates the loop ends without another convergence test.
    if (it >= st->maxiter) { if (lead) { st->done = 2; st->rnorm = rn; } return; }
    if (rn < st->atol_eff) { if (lead) { st->done = 1; st->rnorm = rn; } return; }
    const double beta = (it > 0) ? rho / st->rho[(it + 1) & 1] : 0.0;
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        double pj = (it > 0) ? __dadd_rn(__dmul_rn(beta, pvec[j]), r[j]) : r[j];
        pvec[j] = pj;
        double x = __dmul_rn(s[j], pj);
        sv[j] = x;
        acc += (j < icpt) ? x : -c[j - icpt] * x;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_shift[blockIdx.x] = acc;
    if (lead) { st->rho[it & 1] = rho; st->rnorm = rn; }
}

// alpha = rho/(p.q) ; x += alpha p ; r -= alpha q ; partial r.r
__global__ void k_cg_update(const CgScalars* st, const double* __restrict__ red_pq, int nred, i64 P,
                            double* __restrict__ x, double* __restrict__ r, const double* __restrict__ pvec,
                            const double* __restrict__ q, double* __restrict__ red_rr) {
    if (st->done) return;
    __shared__ double sm[33];
    const int it = st->iter - 1;            // k_cg_q already counted this iteration
    const double pq = warp_sum_partials(red_pq, nred);
    const double alpha = st->rho[it & 1] / pq;
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        x[j] = __dadd_rn(x[j], __dmul_rn(alpha, pvec[j]));
        double rj = __dsub_rn(r[j], __dmul_rn(alpha, q[j]));
        r[j] = rj;
        acc += rj * rj;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_rr[blockIdx.x] = acc;
}

__global__ void k_cg_final(const CgScalars* st, const double* __restrict__ s, const double* __restrict__ x, i64 P,
                           double* __restrict__ coef) {
    const bool zero_rhs = (st->done == 3);   // scipy: `if bnrm2 == 0: return b, 0`
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x)
        coef[j] = zero_rhs ? 0.0 : __dmul_rn(s[j], x[j]);
}

static int P_grid(i64 P) {
    i64 g = (P + 1023) / 1024; if (g < 1) g = 1; if (g > RED_MAX) g = RED_MAX; return (int)g;
}
static int N_grid(i64 n) { return P_grid(n); }

// one application of the operator to the vector whose scaled image is already in m->sv
// (k_prepare / k_cg_dir wrote sv and the shift partials): q = D.v + s.X'(omega.(X sv))
static int apply_operator(bb_mat* m, const double* vP, int count_iter) {
    bb_ctx* ctx = m->ctx;
    const int* done = &m->cg->done;
    BB_TRY(bb_op_dot_flag(m, 1, done));
    P2PView view;
    const bool p2p = (ctx->nranks > 1) && bb_p2p_view(ctx, m->p + 1, &view);
    BB_TRY(bb_op_tdot_flag(m, m->w_n, true, done, /*fuse_reduce_into_consumer=*/p2p));
    // red_pq holds one partial per block and k_cg_update sums P_grid(P) of them: keep that grid (<= RED_MAX);
    k_cg_q<<<P_grid(m->P), 256, 0, ctx->stream>>>(m->cg, m->traw, m->col_offset, m->add_intercept, m->P, vP,
                                                  m->s, m->D, m->q, m->red + RED_PQ * RED_MAX, count_iter,
                                                  p2p ? m->p2p_view_dev : nullptr);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

static int cg_iteration(bb_mat* m) {
    bb_ctx* ctx = m->ctx;
    const int gP = P_grid(m->P);
    k_cg_dir<<<gP, 256, 0, ctx->stream>>>(m->cg, m->red + RED_RR * RED_MAX, gP, m->P, m->add_intercept, m->r, m->pvec, m->s,
                                          m->col_offset, m->sv, m->red + RED_SHIFT * RED_MAX);
    BB_LAUNCHED(ctx);
    BB_TRY(apply_operator(m, m->pvec, 1));
    k_cg_update<<<gP, 256, 0, ctx->stream>>>(m->cg, m->red + RED_PQ * RED_MAX, gP, m->P, m->x, m->r, m->pvec, m->q,
                                             m->red + RED_RR * RED_MAX);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

// compute z = X'(omega.y_gaussian) into m->z (device)
static int compute_z(bb_mat* m) {
    bb_ctx* ctx = m->ctx;
    if (!m->has_outcome) { bb_set_error("z == NULL requires bb_set_outcome"); return BB_ERR_STATE; }
    if (!m->is_linear && m->zk_valid) {
        BB_CUDA(cudaMemcpyAsync(m->z, m->zk, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        return BB_OK;
    }
    k_z_weight<<<N_grid(m->n), 256, 0, ctx->stream>>>(m->n_trial, m->n_success, m->is_linear,
                                                      m->use_omega_scalar ? nullptr : m->omega, m->omega_scalar, m->n, m->u_n);
    BB_LAUNCHED(ctx);
    BB_TRY(bb_op_tdot(m, m->u_n));
    BB_TRY(bb_op_tdot_finish(m, m->z));
    if (!m->is_linear) {
        BB_CUDA(cudaMemcpyAsync(m->zk, m->z, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        m->zk_valid = 1;
    }
    return BB_OK;
}

extern "C" int bb_cg_sample(bb_mat* m, const double* omega, const double* prior_prec_sqrt,
                            const double* z, const double* x0, const double* precond_scale,
                            double atol, int maxiter, int noise_mode,
                            const double* eps1, const double* eps2, uint64_t seed, uint64_t offset,
                            double* coef_out, int* n_iter, int* info, double* stats) {
    BB_ARG(m && prior_prec_sqrt && x0 && precond_scale && coef_out, "null pointer");
    BB_ARG(noise_mode == BB_NOISE_PHILOX || (eps1 && eps2), "BB_NOISE_INJECT needs eps1 and eps2");
    BB_ARG(maxiter >= 0, "maxiter");
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    const size_t Pb = (size_t)m->P * sizeof(double), nb = (size_t)m->n * sizeof(double);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (stats) { BB_CUDA(cudaEventCreate(&ev0)); BB_CUDA(cudaEventCreate(&ev1)); BB_CUDA(cudaEventRecord(ev0, st)); }

    if (omega) { BB_CUDA(cudaMemcpyAsync(m->omega, omega, nb, cudaMemcpyHostToDevice, st)); m->use_omega_scalar = 0; }
    BB_CUDA(cudaMemcpyAsync(m->pps, prior_prec_sqrt, Pb, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(m->x0, x0, Pb, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(m->s, precond_scale, Pb, cudaMemcpyHostToDevice, st));
    if (z) BB_CUDA(cudaMemcpyAsync(m->z, z, Pb, cudaMemcpyHostToDevice, st));
    else BB_TRY(compute_z(m));
    const int philox = (noise_mode == BB_NOISE_PHILOX);
    if (!philox) {
        BB_CUDA(cudaMemcpyAsync(m->eps_n, eps1, nb, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(m->eps_P, eps2, Pb, cudaMemcpyHostToDevice, st));
    }
    const double* om = m->use_omega_scalar ? nullptr : m->omega;
    // the captured iteration graph bakes pointer arguments: rebuild it when omega switches between
    // the vector and the scalar representation; the scalar VALUE travels through device memory
    if (m->cg_graph && m->cg_graph_scalar_mode != m->use_omega_scalar) {
        cudaGraphExecDestroy(m->cg_graph);
        m->cg_graph = nullptr;
    }
    m->cg_graph_scalar_mode = m->use_omega_scalar;
    BB_CUDA(cudaMemcpyAsync(m->omega_scalar_dev, &m->omega_scalar, sizeof(double), cudaMemcpyHostToDevice, st));
    // right-hand side
    k_rhs_noise<<<N_grid(m->n), 256, 0, st>>>(om, m->omega_scalar, m->eps_n, philox, seed, offset, m->row_offset, m->n, m->u_n);
    BB_LAUNCHED(ctx);
    BB_TRY(bb_op_tdot(m, m->u_n));
    const int gP = P_grid(m->P);
    k_cg_init<<<gP, 256, 0, st>>>(m->traw, m->col_offset, m->add_intercept, m->P, m->z, m->pps, m->s, m->x0, m->eps_P,
                                  philox, seed, offset, m->b, m->D, m->x, m->red + RED_BB * RED_MAX);
    BB_LAUNCHED(ctx);
    k_cg_scalars<<<1, 32, 0, st>>>(m->cg, m->red + RED_BB * RED_MAX, gP, atol, maxiter);
    BB_LAUNCHED(ctx);
    // initial residual r = b - A x  (exactly b when x0 == 0)
    BB_TRY(bb_op_prepare_flag(m, m->x, m->s, &m->cg->done));
    BB_TRY(apply_operator(m, m->x, 0));
    k_cg_resid<<<gP, 256, 0, st>>>(m->cg, m->b, m->q, m->P, m->r, m->red + RED_RR * RED_MAX);
    BB_LAUNCHED(ctx);

    // iterations, enqueued in chunks; the device skips work once done != 0.  One CUDA graph holds
    // CG_GRAPH_ITERS iterations (fewer graph launches; surplus iterations are early-exit kernels).
    const int CG_GRAPH_ITERS = 4;
    int total = 0;
    int first = (ctx->opt_cg_chunk > 0) ? (int)ctx->opt_cg_chunk : (m->last_n_iter > 0 ? m->last_n_iter + 1 : 8);
    int chunk = first;
    const bool use_graph = ctx->opt_use_graph != 0;
    for (;;) {
        // +1: the convergence test of iteration k runs at the start of launch k
        int todo = chunk;
        if (total + todo > maxiter + 1) todo = maxiter + 1 - total;
        if (todo < 1) todo = 1;
        if (use_graph) {
            if (!m->cg_graph) {
                cudaGraph_t g = nullptr;
                const i64 l0 = ctx->launches;
                BB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                int rc = BB_OK;
                for (int k = 0; k < CG_GRAPH_ITERS && rc == BB_OK; ++k) rc = cg_iteration(m);
                cudaError_t e = cudaStreamEndCapture(st, &g);
                m->cg_graph_launches = (int)(ctx->launches - l0);
                ctx->launches = l0;
                if (rc != BB_OK) { if (g) cudaGraphDestroy(g); return rc; }
                if (e != cudaSuccess) { bb_set_error("graph capture: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
                e = cudaGraphInstantiate(&m->cg_graph, g, 0);
                cudaGraphDestroy(g);
                if (e != cudaSuccess) { bb_set_error("graph instantiate: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
            }
            const int ngraphs = (todo + CG_GRAPH_ITERS - 1) / CG_GRAPH_ITERS;
            for (int k = 0; k < ngraphs; ++k) BB_CUDA(cudaGraphLaunch(m->cg_graph, st));
            ctx->launches += (i64)ngraphs * m->cg_graph_launches;
            todo = ngraphs * CG_GRAPH_ITERS;
        } else {
            for (int k = 0; k < todo; ++k) BB_TRY(cg_iteration(m));
        }
        total += todo;
        BB_CUDA(cudaMemcpyAsync(m->cg_host, m->cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        if (m->cg_host->done != 0) break;
        if (total >= maxiter + 1) break;   // cannot happen: launch maxiter+1 sets done=2
        chunk = (ctx->opt_cg_chunk > 0) ? (int)ctx->opt_cg_chunk : CG_GRAPH_ITERS;
    }
    k_cg_final<<<gP, 256, 0, st>>>(m->cg, m->s, m->x, m->P, m->out_P);
    BB_LAUNCHED(ctx);
    BB_CUDA(cudaMemcpyAsync(coef_out, m->out_P, Pb, cudaMemcpyDeviceToHost, st));
    if (stats) BB_CUDA(cudaEventRecord(ev1, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    const int done = m->cg_host->done;
    m->last_n_iter = m->cg_host->iter;
    if (n_iter) *n_iter = m->cg_host->iter;
    if (info) *info = (done == 1 || done == 3) ? 0 : maxiter;
    if (stats) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        stats[0] = m->cg_host->bnorm; stats[1] = m->cg_host->rnorm; stats[2] = (double)ms;
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    }
    return BB_OK;
}
