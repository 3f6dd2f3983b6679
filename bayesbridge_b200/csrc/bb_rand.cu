// Device random variates: Polya-Gamma (reference: bayesbridge/random/polya_gamma/polya_gamma.pyx:40-216,
// log_ndtr from scipy_ndtr.c:367-396), exponentially tilted stable (tilted_stable.pyx:65-332) and the
// Gaussian streams of the CG right-hand side.  One thread per variate; each variate owns a
// counter-based Philox4x32-10 stream keyed by (seed, call offset, GLOBAL element index, stream id),
// so results do not depend on how rows are sharded across GPUs and a chain can be resumed from
// (seed, offset) alone.
#include "bb_internal.cuh"

#define BB_PI 3.14159265358979323846264338327950288

// ---- log of the standard normal CDF, three regimes as in scipy_ndtr.c:367-396 ----------------
__device__ double dev_log_ndtr(double a) {
    if (a > 6.0) return -normcdf(-a);           // log(1 - eps) ~ -eps
    if (a > -20.0) return log(normcdf(a));
    double log_lhs = -0.5 * a * a - log(-a) - 0.5 * log(2.0 * BB_PI);
    double last_total = 0.0, rhs = 1.0, numerator = 1.0, denom_factor = 1.0, denom_cons = 1.0 / (a * a);
    double sign = 1.0;
    long i = 0;
    while (fabs(last_total - rhs) > 2.2204460492503131e-16) {
        i += 1;
        last_total = rhs;
        sign = -sign;
        denom_factor *= denom_cons;
        numerator *= (double)(2 * i - 1);
        rhs += sign * numerator * denom_factor;
    }
    return log_lhs + log(rhs);
}

// ---- Polya-Gamma ------------------------------------------------------------------------------
#define PG_T 0.63661977236758134308   /* 2/pi: where the two series representations meet */
#define PG_MAX_TERMS 100

// a_n(x): terms of the alternating series, Polson-Scott-Windle (2013) eq. (12)-(13)
__device__ __forceinline__ double pg_series_term(int n, double x) {
    double nh = (double)n + 0.5;
    double lr = log(BB_PI * nh);
    if (x <= PG_T) lr += -1.5 * log(0.5 * x * BB_PI) - 2.0 * nh * nh / x;
    else lr += -0.5 * x * BB_PI * BB_PI * nh * nh;
    return exp(lr);
}

__device__ double pg_prob_right(double z, double K) {
    double lm_expo = -log(K) - K * PG_T + log(0.25 * BB_PI);
    double st = sqrt(PG_T);
    double lm_ig1 = -z + dev_log_ndtr((PG_T * z - 1.0) / st);
    double lm_ig2 = z + dev_log_ndtr(-(PG_T * z + 1.0) / st);
    double ratio = exp(lm_ig1 - lm_expo) + exp(lm_ig2 - lm_expo);
    return 1.0 / (1.0 + ratio);
}

// inverse Gaussian(mean 1/z, shape 1) restricted to (0, t)
__device__ double pg_trunc_invgauss(RandStream& rs, double z) {
    double X;
    double mean = 1.0 / z;
    if (mean > PG_T) {
        // 1/chi^2_1 restricted to (0, t) as proposal; accept with exp(-z^2 X / 2)
        for (;;) {
            double E;
            for (;;) {   // chi^2_1 left-truncated at 1/t: shifted Exp(2) proposal
                E = 0.5 * BB_PI - 2.0 * log(1.0 - rs.uniform());
                if (rs.uniform() <= sqrt(0.5 * BB_PI / E)) break;
            }
            X = 1.0 / E;
            if (log(rs.uniform()) < -0.5 * X * z * z) break;
        }
    } else {
        for (;;) {   // Michael-Schucany-Haas, repeated until it lands below t
            double N = rs.normal();
            double V = N * N;
            X = mean + 0.5 * mean * (mean * V - sqrt(4.0 * mean * V + mean * mean * V * V));
            if (rs.uniform() > mean / (mean + X)) X = mean * mean / X;
            if (X < PG_T) break;
        }
    }
    return X;
}

// J*(1, z): Devroye's alternating-series rejection sampler
__device__ double pg_tilted_jacobi(RandStream& rs, double z) {
    const double K = 0.5 * z * z + 0.125 * BB_PI * BB_PI;
    const double p_right = pg_prob_right(z, K);
    for (;;) {
        double X;
        if (rs.uniform() < p_right) X = PG_T - log(1.0 - rs.uniform()) / K;
        else X = pg_trunc_invgauss(rs, z);
        const double a0 = pg_series_term(0, X);
        const double U = rs.uniform() * a0;
        double partial = a0;
        int n = 1;
        double sign = -1.0;
        for (;;) {
            partial += sign * pg_series_term(n, X);
            n += 1;
            if (sign < 0.0) { if (U <= partial) return X; }
            else {
                if (U > partial) break;                 // rejected
                if (n >= PG_MAX_TERMS) return X;        // lower bound taken as the target
            }
            sign = -sign;
        }
    }
}

__device__ double pg_draw(RandStream& rs, int shape, double tilt) {
    const double z = 0.5 * fabs(tilt);
    double acc = 0.0;
    for (int j = 0; j < shape; ++j) acc += 0.25 * pg_tilted_jacobi(rs, z);
    return acc;
}

__global__ void k_pg_sample(i64 n, const int* __restrict__ shape, const double* __restrict__ shape_d,
                            const double* __restrict__ tilt, uint64_t seed, uint64_t offset, i64 index_offset,
                            double* __restrict__ out, int* __restrict__ bad_shape) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RandStream rs;
    rs.init(seed, offset, (uint64_t)(index_offset + i), STREAM_PG);
    int b = shape ? shape[i] : (int)shape_d[i];
    if (b < 0) { if (bad_shape) atomicOr(bad_shape, 1); out[i] = 0.0; return; }     // polya_gamma.pyx:55-61: validated here, not by a host loop
    out[i] = pg_draw(rs, b, tilt[i]);
}

// fused: omega_i ~ PG(n_trial_i, eta_i) and the logistic log-likelihood terms
__global__ void k_pg_loglik(i64 n, const double* __restrict__ n_trial, const double* __restrict__ n_success,
                            const double* __restrict__ eta, uint64_t seed, uint64_t offset, i64 index_offset,
                            double* __restrict__ omega, double* __restrict__ red_ll) {
    __shared__ double sm[33];
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    double ll = 0.0;
    if (i < n) {
        RandStream rs;
        rs.init(seed, offset, (uint64_t)(index_offset + i), STREAM_PG);
        double e = eta[i], nt = n_trial[i];
        omega[i] = pg_draw(rs, (int)nt, e);
        // logistic_model.py:52-55   n_success*eta - n_trial*logaddexp(0, eta)
        double lae = (e > 0.0) ? e + log1p(exp(-e)) : log1p(exp(e));
        ll = n_success[i] * e - nt * lae;
    }
    ll = block_sum(ll, sm);
    if (threadIdx.x == 0) red_ll[blockIdx.x] = ll;
}

// ---- exponentially tilted stable ---------------------------------------------------------------
__device__ __forceinline__ double ts_sinc(double x) {
    if (fabs(x) < 0.01) { double x2 = x * x; return 1.0 - x2 / 6.0 * (1.0 - x2 / 20.0); }
    return sin(x) / x;
}
__device__ double ts_zolotarev(double x, double al) {
    double v = pow((1.0 - al) * ts_sinc((1.0 - al) * x), 1.0 - al) * pow(al * ts_sinc(al * x), al) / ts_sinc(x);
    return pow(v, 1.0 / (1.0 - al));
}
__device__ double ts_zolotarev_pdf_exp(double x, double al) {
    double den = pow(ts_sinc(al * x), al) * pow(ts_sinc((1.0 - al) * x), 1.0 - al);
    return ts_sinc(x) / den;
}

__device__ double ts_divide_conquer(RandStream& rs, double al, double tilt) {
    double tp = floor(pow(tilt, al));
    long m = (tp >= 1.0) ? (long)tp : 1;
    double c = pow(1.0 / (double)m, 1.0 / al);
    double X = 0.0;
    for (long i = 0; i < m; ++i) {
        double S;
        for (;;) {
            double u1 = rs.uniform(), u2 = rs.uniform();
            S = c * pow(-ts_zolotarev(BB_PI * u1, al) / log(u2), (1.0 - al) / al);
            if (rs.uniform() < exp(-tilt * S)) break;
        }
        X += S;
    }
    return X;
}

__device__ double ts_double_rejection(RandStream& rs, double al, double tilt) {
    const double b = pow(tilt, al);
    const double gam = b * al * (1.0 - al);
    const double sg = sqrt(gam);
    const double c2 = 2.0 + sqrt(0.5 * BB_PI);
    const double xi = (1.0 + sqrt(2.0 * gam) * c2) / BB_PI;
    const double psi = sqrt(gam / BB_PI) * c2 * exp(-gam * BB_PI * BB_PI / 8.0);
    const double w1 = sqrt(0.5 * BB_PI / gam) * xi, w2 = 2.0 * sqrt(BB_PI) * psi, w3 = xi * BB_PI;
    const double odds = (1.0 - al) / al;
    for (;;) {
        // auxiliary variable U (and V, z)
        double U, V, z;
        for (;;) {
            double Vs = rs.uniform();
            if (gam >= 1.0) {
                if (Vs < w1 / (w1 + w2)) U = fabs(rs.normal()) / sg;
                else { double W = rs.uniform(); U = BB_PI * (1.0 - W * W); }
            } else {
                double W = rs.uniform();
                if (Vs < w3 / (w2 + w3)) U = BB_PI * W;
                else U = BB_PI * (1.0 - W * W);
            }
            if (U > BB_PI) continue;
            double zeta = sqrt(ts_zolotarev_pdf_exp(U, al));
            z = 1.0 / (1.0 - pow(1.0 + al * zeta / sg, -1.0 / al));
            double inv = BB_PI * exp(-b * (1.0 - 1.0 / (zeta * zeta))) / ((1.0 + sqrt(0.5 * BB_PI)) * sg / zeta + z);
            double d = 0.0;
            if (U >= 0.0 && gam >= 1.0) d += xi * exp(-gam * U * U / 2.0);
            if (U > 0.0 && U < BB_PI) d += psi / sqrt(BB_PI - U);
            if (U >= 0.0 && U <= BB_PI && gam < 1.0) d += xi;
            inv *= d;
            double accept_prob = 1.0 / inv;
            if (accept_prob > 0.0) {
                V = rs.uniform() / accept_prob;
                if (U < BB_PI && V <= 1.0) break;
            }
        }
        // reference variable X given U
        double a = ts_zolotarev(U, al);
        double left = pow(odds / a, al) * b;
        double delta = sqrt(left * al / a);
        double right = left + delta;
        double m_left = delta * sqrt(0.5 * BB_PI), m_mid = delta, m_right = z / a;
        double total = m_left + m_mid + m_right;
        double Vr = rs.uniform();
        double N = 0.0, E = 0.0, X;
        if (Vr < m_left / total) { N = rs.normal(); X = left - delta * fabs(N); }
        else if (Vr < (m_left + m_mid) / total) X = left + delta * rs.uniform();
        else { E = -log(rs.uniform()); X = right + E * m_right; }
        double lacc;
        if (X < 0.0) lacc = -INFINITY;
        else {
            lacc = -(a * (X - left) + exp(log(b) / al - odds * log(left)) * (pow(left / X, odds) - 1.0));
            if (X < left) lacc += N * N / 2.0;
            else if (X > right) lacc += E;
        }
        if (lacc > log(V)) return pow(X, -odds);
    }
}

__global__ void k_tilted_stable(i64 n, double al, const double* __restrict__ tilt, uint64_t seed, uint64_t offset,
                                i64 index_offset, double* __restrict__ out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RandStream rs;
    rs.init(seed, offset, (uint64_t)(index_offset + i), STREAM_TS);
    double t = tilt[i];
    // tilted_stable.pyx:103-108: divide-and-conquer is cheaper while tilt^alpha < 2
    out[i] = (pow(t, al) < 2.0) ? ts_divide_conquer(rs, al, t) : ts_double_rejection(rs, al, t);
}

__global__ void k_philox_normal(i64 n, int stream, uint64_t seed, uint64_t offset, i64 index_offset, double* __restrict__ out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RandStream rs;
    rs.init(seed, offset, (uint64_t)(index_offset + i), (uint32_t)stream);
    out[i] = rs.normal();
}

__global__ void k_rss(i64 n, const double* __restrict__ y, const double* __restrict__ eta, double* __restrict__ red) {
    __shared__ double sm[33];
    double acc = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double d = y[i] - eta[i];
        acc += d * d;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red[blockIdx.x] = acc;
}

__global__ void k_finish_scalar(const double* __restrict__ red, int nred, double* __restrict__ out) {
    double s = warp_sum_partials(red, nred);
    if (threadIdx.x == 0) out[0] = s;
}

// ---- host entry points -------------------------------------------------------------------------

extern "C" int bb_pg_sample(bb_ctx* ctx, int64_t n, const int32_t* shape, const double* tilt,
                            uint64_t seed, uint64_t offset, int64_t index_offset, double* out) {
    BB_ARG(ctx && n >= 0, "ctx/n");
    if (n == 0) return BB_OK;
    BB_ARG(shape && tilt && out, "null pointer");
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    cudaStream_t st = ctx->stream;
    int* d_shape = nullptr; double *d_tilt = nullptr, *d_out = nullptr;
    BB_TRY(bb_ctx_scratch(ctx, 0, (size_t)n * sizeof(int), (void**)&d_shape));
    BB_TRY(bb_ctx_scratch(ctx, 1, (size_t)n * sizeof(double), (void**)&d_tilt));
    BB_TRY(bb_ctx_scratch(ctx, 2, (size_t)n * sizeof(double), (void**)&d_out));
    int* d_bad = nullptr;
    BB_TRY(bb_ctx_scratch(ctx, 3, 64, (void**)&d_bad));
    int rc = BB_OK, bad = 0;
    cudaError_t e;
    e = cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_shape, shape, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tilt, tilt, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        k_pg_sample<<<(int)((n + 127) / 128), 128, 0, st>>>(n, d_shape, nullptr, d_tilt, seed, offset, index_offset, d_out, d_bad);
        ctx->launches++;
        e = cudaPeekAtLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    timer_.end();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    timer_.commit();
    if (e != cudaSuccess) { bb_set_error("bb_pg_sample: %s", cudaGetErrorString(e)); rc = BB_ERR_CUDA; }
    else if (bad) { bb_set_error("bb_pg_sample: invalid argument: shape must be non-negative"); rc = BB_ERR_ARG; }
    return rc;
}

extern "C" int bb_tilted_stable_sample(bb_ctx* ctx, int64_t n, double char_exp, const double* tilt,
                                       uint64_t seed, uint64_t offset, int64_t index_offset, double* out) {
    BB_ARG(ctx && n >= 0, "ctx/n");
    if (n == 0) return BB_OK;
    BB_ARG(tilt && out, "null pointer");
    BB_ARG(char_exp > 0.0 && char_exp < 1.0, "characteristic exponent must be in (0,1)");
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    cudaStream_t st = ctx->stream;
    double *d_tilt = nullptr, *d_out = nullptr;
    BB_TRY(bb_ctx_scratch(ctx, 1, (size_t)n * sizeof(double), (void**)&d_tilt));
    BB_TRY(bb_ctx_scratch(ctx, 2, (size_t)n * sizeof(double), (void**)&d_out));
    int rc = BB_OK;
    cudaError_t e = cudaMemcpyAsync(d_tilt, tilt, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        k_tilted_stable<<<(int)((n + 127) / 128), 128, 0, st>>>(n, char_exp, d_tilt, seed, offset, index_offset, d_out);
        ctx->launches++;
        e = cudaPeekAtLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st);
    timer_.end();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    timer_.commit();
    if (e != cudaSuccess) { bb_set_error("bb_tilted_stable_sample: %s", cudaGetErrorString(e)); rc = BB_ERR_CUDA; }
    return rc;
}

extern "C" int bb_philox_normal(bb_ctx* ctx, int64_t n, int stream, uint64_t seed, uint64_t offset,
                                int64_t index_offset, double* out) {
    BB_ARG(ctx && n >= 0 && (n == 0 || out), "ctx/n/out");
    if (n == 0) return BB_OK;
    BB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    double* d_out = nullptr;
    BB_TRY(bb_ctx_scratch(ctx, 2, (size_t)n * sizeof(double), (void**)&d_out));
    k_philox_normal<<<(int)((n + 255) / 256), 256, 0, st>>>(n, stream, seed, offset, index_offset, d_out);
    ctx->launches++;
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { bb_set_error("bb_philox_normal: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
    return BB_OK;
}

// ---- resident observation-side vectors ---------------------------------------------------------
extern "C" int bb_set_outcome(bb_mat* m, const double* n_trial, const double* n_success) {
    BB_ARG(m && n_success, "mat/n_success");
    bb_ctx* ctx = m->ctx;
    BB_CUDA(cudaSetDevice(ctx->device));
    size_t nb = (size_t)m->n * sizeof(double);
    BB_CUDA(cudaMemcpyAsync(m->n_success, n_success, nb, cudaMemcpyHostToDevice, ctx->stream));
    if (n_trial) BB_CUDA(cudaMemcpyAsync(m->n_trial, n_trial, nb, cudaMemcpyHostToDevice, ctx->stream));
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    m->is_linear = (n_trial == nullptr);
    m->has_outcome = 1;
    m->zk_valid = 0;
    return BB_OK;
}

extern "C" int bb_set_obs_prec(bb_mat* m, const double* omega) {
    BB_ARG(m && omega, "mat/omega");
    BB_CUDA(cudaSetDevice(m->ctx->device));
    BB_CUDA(cudaMemcpyAsync(m->omega, omega, (size_t)m->n * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    m->use_omega_scalar = 0;
    return BB_OK;
}

extern "C" int bb_set_obs_prec_scalar(bb_mat* m, double omega) {
    BB_ARG(m != nullptr, "mat");
    m->omega_scalar = omega;
    m->use_omega_scalar = 1;
    return BB_OK;
}

extern "C" int bb_get_obs_prec(bb_mat* m, double* out) {
    BB_ARG(m && out, "mat/out");
    BB_CUDA(cudaSetDevice(m->ctx->device));
    if (m->use_omega_scalar) { for (i64 i = 0; i < m->n; ++i) out[i] = m->omega_scalar; return BB_OK; }
    BB_CUDA(cudaMemcpyAsync(out, m->omega, (size_t)m->n * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return BB_OK;
}

extern "C" int bb_get_linear_predictor(bb_mat* m, double* out) {
    BB_ARG(m && out, "mat/out");
    BB_CUDA(cudaSetDevice(m->ctx->device));
    BB_CUDA(cudaMemcpyAsync(out, m->eta, (size_t)m->n * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return BB_OK;
}

static int N_grid(i64 n) { i64 g = (n + 1023) / 1024; if (g < 1) g = 1; if (g > RED_MAX) g = RED_MAX; return (int)g; }

// eta = X coef into m->eta (device)
static int linear_predictor(bb_mat* m, const double* coef) {
    bb_ctx* ctx = m->ctx;
    // coef == NULL: the coefficients of the last CG draw, still resident in m->out_P
    if (coef) BB_CUDA(cudaMemcpyAsync(m->v_P, coef, (size_t)m->P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    else BB_CUDA(cudaMemcpyAsync(m->v_P, m->out_P, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    BB_TRY(bb_op_prepare(m, m->v_P, nullptr));
    BB_TRY(bb_op_dot(m, 0));
    BB_CUDA(cudaMemcpyAsync(m->eta, m->u_n, (size_t)m->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return BB_OK;
}

extern "C" int bb_pg_from_coef(bb_mat* m, const double* coef, uint64_t seed, uint64_t offset,
                               double* omega_out, double* loglik) {
    BB_ARG(m != nullptr, "mat");      // coef == NULL: use the coefficients of the last CG draw (resident)
    if (!m->has_outcome || m->is_linear) { bb_set_error("bb_pg_from_coef needs a logit outcome (bb_set_outcome)"); return BB_ERR_STATE; }
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    BB_TRY(linear_predictor(m, coef));
    const int TB = 128;
    i64 nblk = (m->n + TB - 1) / TB;
    if (nblk < 1) nblk = 1;
    // partial log-likelihoods: one per block; collapse in passes of RED_MAX
    double* red_ll = nullptr;
    BB_TRY(bb_ctx_scratch(ctx, 3, (size_t)nblk * sizeof(double), (void**)&red_ll));
    k_pg_loglik<<<(int)nblk, TB, 0, st>>>(m->n, m->n_trial, m->n_success, m->eta, seed, offset, m->row_offset, m->omega, red_ll);
    ctx->launches++;
    m->use_omega_scalar = 0;
    // simple and deterministic: one warp sums all block partials in fixed order
    k_finish_scalar<<<1, 32, 0, st>>>(red_ll, (int)nblk, m->traw);
    ctx->launches++;
    BB_TRY(bb_allreduce_dev(ctx, m->traw, 1));
    double ll = 0.0;
    BB_CUDA(cudaMemcpyAsync(&ll, m->traw, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (omega_out) BB_CUDA(cudaMemcpyAsync(omega_out, m->omega, (size_t)m->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    timer_.end();
    cudaError_t e = cudaStreamSynchronize(st);
    timer_.commit();
    if (e != cudaSuccess) { bb_set_error("bb_pg_from_coef: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
    if (loglik) *loglik = ll;
    return BB_OK;
}

extern "C" int bb_linear_rss(bb_mat* m, const double* coef, double* rss) {
    BB_ARG(m && rss, "mat/rss");          // coef == NULL: resident coefficients of the last CG draw
    if (!m->has_outcome) { bb_set_error("bb_linear_rss needs bb_set_outcome"); return BB_ERR_STATE; }
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    BB_TRY(linear_predictor(m, coef));
    int g = N_grid(m->n);
    k_rss<<<g, 256, 0, st>>>(m->n, m->n_success, m->eta, m->red + RED_LL * RED_MAX);
    ctx->launches++;
    k_finish_scalar<<<1, 32, 0, st>>>(m->red + RED_LL * RED_MAX, g, m->traw);
    ctx->launches++;
    BB_TRY(bb_allreduce_dev(ctx, m->traw, 1));
    BB_CUDA(cudaMemcpyAsync(rss, m->traw, sizeof(double), cudaMemcpyDeviceToHost, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    return BB_OK;
}


// ---- batched chains (bb_batch.cu): omega[i][c] ~ PG(n_trial_i, eta[i][c]) on chain c's own stream, log-likelihood per chain ----
constexpr int PGB_BC = 16;
__global__ void __launch_bounds__(256)
k_pg_loglik_batched(i64 n, const double* __restrict__ n_trial, const double* __restrict__ n_success, const double* __restrict__ eta_b,
                    int C, const uint64_t* __restrict__ seeds, const uint64_t* __restrict__ offsets, i64 row_offset,
                    double* __restrict__ omega_b, double* __restrict__ red /*[grid][16]*/) {
    __shared__ double sm[8][PGB_BC];
    const i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
    const i64 i = idx >> 4;
    const int c = (int)(idx & 15), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double ll = 0.0;
    if (i < n) {
        double om = 0.0;
        if (c < C) {
            RandStream rs;
            rs.init(seeds[c], offsets[c], (uint64_t)(row_offset + i), STREAM_PG);
            const double e = eta_b[idx], nt = n_trial[i];
            om = pg_draw(rs, (int)nt, e);
            const double lae = (e > 0.0) ? e + log1p(exp(-e)) : log1p(exp(e));
            ll = n_success[i] * e - nt * lae;
        }
        omega_b[idx] = om;
    }
    ll += __shfl_xor_sync(0xffffffffu, ll, 16);            // the two rows of a warp
    if (lane < 16) sm[warp][lane] = ll;
    __syncthreads();
    if (threadIdx.x < PGB_BC) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
        red[(i64)blockIdx.x * PGB_BC + threadIdx.x] = t;
    }
}
__global__ void k_pg_ll_finish_batched(const double* __restrict__ red, i64 nblk, double* __restrict__ ll_b) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int c = warp; c < PGB_BC; c += nw) {
        double t = 0.0;
        for (i64 k = lane; k < nblk; k += 32) t += red[k * PGB_BC + c];
        t = warp_sum(t);
        if (lane == 0) ll_b[c] = t;
    }
}
int bb_batch_pg_launch(bb_mat* m, const double* eta_b, int C, const uint64_t* seeds_dev, const uint64_t* offsets_dev,
                       double* omega_b, double* red_scratch, double* ll_b) {
    bb_ctx* ctx = m->ctx;
    const i64 nblk = (m->n * PGB_BC + 255) / 256;
    if (nblk > 0) {
        k_pg_loglik_batched<<<(unsigned)nblk, 256, 0, ctx->stream>>>(m->n, m->n_trial, m->n_success, eta_b, C, seeds_dev, offsets_dev,
                                                                    m->row_offset, omega_b, red_scratch);
        BB_LAUNCHED(ctx);
    }
    k_pg_ll_finish_batched<<<1, 512, 0, ctx->stream>>>(red_scratch, nblk, ll_b);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

// ---- log-likelihood and its gradient with everything n-length resident (chain initialisation) ----
// logit  (logistic_model.py:49-55): ll = sum n_success eta - n_trial log(1 + e^eta) ; grad = X'(n_success - n_trial sigmoid(eta))
// linear (linear_model.py:13-24):   ll = -prec/2 sum (y - eta)^2 (the n/2 log prec term is the host's) ; grad = prec X'(y - eta)
__global__ void k_loglik_resid(i64 n, const double* __restrict__ n_trial, const double* __restrict__ y_or_success, int is_linear,
                               double prec, const double* __restrict__ eta, double* __restrict__ w, double* __restrict__ red) {
    __shared__ double sm[33];
    double acc = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double e = eta[i];
        if (is_linear) {
            const double r = y_or_success[i] - e;
            acc += r * r;
            w[i] = prec * r;
        } else {
            const double nt = n_trial[i], ns = y_or_success[i];
            acc += ns * e - nt * (fmax(e, 0.0) + log1p(exp(-fabs(e))));       // np.logaddexp(0, eta)
            w[i] = ns - nt / (1.0 + exp(-e));
        }
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red[blockIdx.x] = acc;
}

// device-resident form used by the mode search (bb_lbfgs.cu): coef in m->v_P; ll_dev[0] = sum of the likelihood terms
// (logit) or of the squared residuals (linear), summed over the shards; grad_dev = X' residual (P entries)
int bb_loglik_resid_dev(bb_mat* m, const double* coef_dev, double obs_prec, double* ll_dev, double* grad_dev) {
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_TRY(bb_op_prepare(m, coef_dev, nullptr));
    BB_TRY(bb_op_dot(m, 0));
    const int g = N_grid(m->n);
    k_loglik_resid<<<g, 256, 0, st>>>(m->n, m->n_trial, m->n_success, m->is_linear, obs_prec, m->u_n, m->w_n, m->red + RED_LL * RED_MAX);
    ctx->launches++;
    k_finish_scalar<<<1, 32, 0, st>>>(m->red + RED_LL * RED_MAX, g, ll_dev);
    ctx->launches++;
    BB_TRY(bb_allreduce_dev(ctx, ll_dev, 1));
    BB_TRY(bb_op_tdot(m, m->w_n));
    BB_TRY(bb_op_tdot_finish(m, grad_dev));
    return BB_OK;
}

extern "C" int bb_loglik_and_gradient(bb_mat* m, const double* coef, double obs_prec, int loglik_only,
                                      double* loglik, double* grad) {
    BB_ARG(m && coef && loglik && (loglik_only || grad), "mat/coef/loglik/grad");
    if (!m->has_outcome) { bb_set_error("bb_loglik_and_gradient needs bb_set_outcome"); return BB_ERR_STATE; }
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    BB_TRY(linear_predictor(m, coef));
    const int g = N_grid(m->n);
    k_loglik_resid<<<g, 256, 0, st>>>(m->n, m->n_trial, m->n_success, m->is_linear, obs_prec, m->eta, m->w_n, m->red + RED_LL * RED_MAX);
    ctx->launches++;
    double* sc = m->red + RED_MISC * RED_MAX;                      // one double of scratch for the scalar
    k_finish_scalar<<<1, 32, 0, st>>>(m->red + RED_LL * RED_MAX, g, sc);
    ctx->launches++;
    BB_TRY(bb_allreduce_dev(ctx, sc, 1));
    double ll = 0.0;
    BB_CUDA(cudaMemcpyAsync(&ll, sc, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!loglik_only) {
        BB_TRY(bb_op_tdot(m, m->w_n));
        BB_TRY(bb_op_tdot_finish(m, m->t_P));
        BB_CUDA(cudaMemcpyAsync(grad, m->t_P, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    *loglik = m->is_linear ? -0.5 * obs_prec * ll : ll;
    return BB_OK;
}

// ---- local scales on the device (bayesbridge.py:458-478 with the state of bb_state_*) --------------------------
// lambda_j = sqrt(0.5 / TS(alpha/2, (beta_j / tau)^2)); counts[0] = #(tilt <= 0), [1] = #(lambda == 0), [2] = #(lambda == inf)
// Rank r of a row-sharded job draws only the scales [lo, hi) (the streams are keyed by the global coefficient index, so
// the values do not depend on who draws them); the pieces are then summed over ranks against zeros.
__global__ void k_local_scale(i64 lo, i64 hi, int k, double char_exp, double gscale, const double* __restrict__ coef,
                              uint64_t seed, uint64_t offset, double* __restrict__ lscale, int* __restrict__ counts) {
    i64 i = lo + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    double r = coef[k + i] / gscale;
    double tilt = __dmul_rn(r, r);
    if (!(tilt > 0.0)) { atomicAdd(&counts[0], 1); return; }
    RandStream rs;
    rs.init(seed, offset, (uint64_t)i, STREAM_TS);
    double ts = (pow(tilt, char_exp) < 2.0) ? ts_divide_conquer(rs, char_exp, tilt) : ts_double_rejection(rs, char_exp, tilt);
    double l = sqrt(0.5 / ts);
    if (l == 0.0) atomicAdd(&counts[1], 1);
    else if (isinf(l)) atomicAdd(&counts[2], 1);
    lscale[i] = l;
}
__global__ void k_counts_to_double(const int* __restrict__ c, double* __restrict__ d) { if (threadIdx.x < 3) d[threadIdx.x] = (double)c[threadIdx.x]; }
__global__ void k_double_to_counts(const double* __restrict__ d, int* __restrict__ c) { if (threadIdx.x < 3) c[threadIdx.x] = (int)(d[threadIdx.x] + 0.5); }
// the reference's repair: zeros -> 10e-16 if any zero, else infinities -> 2/tau
__global__ void k_local_scale_fix(i64 nshrunk, int mode, double gscale, double* __restrict__ lscale) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nshrunk) return;
    double l = lscale[i];
    if (mode == 1 && l == 0.0) lscale[i] = 10e-16;
    if (mode == 2 && isinf(l)) lscale[i] = 2.0 / gscale;
}

extern "C" int bb_local_scale_resident(bb_mat* m, double gscale, double char_exp, uint64_t seed, uint64_t offset,
                                       int* counts_out /*[3]*/, double* lscale_out) {
    BB_ARG(m && counts_out && gscale > 0.0 && char_exp > 0.0 && char_exp < 1.0, "mat/counts/gscale/char_exp");
    if (!m->st_ready) { bb_set_error("device state not set"); return BB_ERR_STATE; }
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    const i64 ns = m->P - m->st_k;
    int* d_counts = nullptr;
    BB_TRY(bb_ctx_scratch(ctx, 0, 4 * sizeof(int), (void**)&d_counts));
    BB_CUDA(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int), st));
    // this rank's share of the draws (the tilted-stable sampler is fp64-ALU bound: ~0.7 ms for 1e5 draws on one GPU)
    const i64 lo = ns * ctx->rank / ctx->nranks, hi = ns * (ctx->rank + 1) / ctx->nranks;
    if (ctx->nranks > 1) {
        if (lo > 0) BB_CUDA(cudaMemsetAsync(m->st_lscale, 0, (size_t)lo * sizeof(double), st));
        if (hi < ns) BB_CUDA(cudaMemsetAsync(m->st_lscale + hi, 0, (size_t)(ns - hi) * sizeof(double), st));
    }
    if (hi > lo) {
        k_local_scale<<<(int)((hi - lo + 127) / 128), 128, 0, st>>>(lo, hi, m->st_k, char_exp, gscale, m->out_P, seed, offset,
                                                                   m->st_lscale, d_counts);
        ctx->launches++;
    }
    if (ctx->nranks > 1) {
        k_counts_to_double<<<1, 32, 0, st>>>(d_counts, m->st_lscale + ns);
        ctx->launches++;
        BB_TRY(bb_allreduce_dev(ctx, m->st_lscale, ns + 3));
        k_double_to_counts<<<1, 32, 0, st>>>(m->st_lscale + ns, d_counts);
        ctx->launches++;
    }
    BB_CUDA(cudaMemcpyAsync(counts_out, d_counts, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    if (counts_out[0] == 0 && (counts_out[1] > 0 || counts_out[2] > 0) && ns > 0) {
        k_local_scale_fix<<<(int)((ns + 127) / 128), 128, 0, st>>>(ns, counts_out[1] > 0 ? 1 : 2, gscale, m->st_lscale);
        ctx->launches++;
    }
    if (lscale_out) {
        BB_CUDA(cudaMemcpyAsync(lscale_out, m->st_lscale, (size_t)ns * sizeof(double), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
    }
    return BB_OK;
}
