// Batched multi-chain coefficient update on a dense design (BASELINE config 5: 16 independent chains on one X).
//
// The reference has one chain per process (bayesbridge.py:109-277, random/random.py:12-18); C chains on the same design
// share the expensive operand X, so their CG iterations are run in lock-step and the two products of an operator
// application become skinny GEMMs:
//     U = X V (n x C),   W = Omega o U (each chain its own omega),   T = X' W (p x C)
// on the fp64 tensor cores (mma.sync m8n8k4 f64; tcgen05 has no fp64 kind).  X is read once per product for all chains
// (16 n p bytes per operator application instead of 16 n p per CHAIN).  A one-pass form (8 n p) would need ~52 TFLOP/s
// of fp64 MMA to stay HBM-bound at C = 16 against the 37 TFLOP/s measured here, i.e. it would be
// tensor-bound at the same speed, so the simpler two-pass form is used.
//
// Layouts: n-side batched vectors are [n][BC] (chain fastest: one 128-byte line per observation), the gather operand of
// X V is [p][BC]; P-side vectors are chain-major [C][Ps] (every chain's CG state contiguous).  BC = 16 is the padded
// batch width; chains c >= C carry zeros.  Every chain has its own CgScalars; a chain that has converged stops updating
// while the others continue, exactly as C separate calls of the single-chain sampler would.
//
// Row sharding: as in the single-chain path the local T (and sum w) of every chain is summed over ranks once per
// iteration (C (p+1) doubles through bb_allreduce_dev).
#include "bb_internal.cuh"
#include <stdlib.h>

constexpr int BC = 16;

__device__ __forceinline__ void bdmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct BatchWork {
    int C;
    i64 Ps;                      // stride of the chain-major P-vectors
    double *omega_b, *w_b, *u_b; // [n][BC]
    double *stage;               // [BC][max(n, Ps)] chain-major staging for host <-> device transposes
    double *part_b; int nblk;    // [nblk][p][BC]
    double *red_w_b; int ndot;   // [ndot][BC] per-CTA sums of W from the X V kernel
    double *traw_b;              // [BC][p + 1]
    double *x, *r, *pv, *q, *b, *s, *D, *pps, *z, *x0, *eps2, *out;   // [BC][Ps]
    double *sv_b, *shift_b;      // [p][BC], [BC]
    CgScalars *cg, *cg_host;     // [BC]
    int* done_count;
    uint64_t *seeds, *offsets;   // [BC] device
    double* ll_b;                // [BC]
    int last_n_iter;
};

// ------------------------------------------------------------------------------------------
// U = X V (+ shift), W = omega o U.   CTA: 128 rows, K chunks of 32 columns through shared memory, 8 warps x (16 rows x 16 chains)
constexpr int BD_ROWS = 128, BD_KC = 32, BD_XS = BD_KC + 4, BD_VS = BC + 4;

template <int MODE>      // 0: out = U ; 1: out = omega o U and per-CTA column sums of it
__global__ void __launch_bounds__(256)
k_batch_dot(const double* __restrict__ X, i64 n, i64 p, const double* __restrict__ sv_b, const double* __restrict__ shift_b,
            const double* __restrict__ omega_b, double* __restrict__ out, double* __restrict__ red_w_b,
            const int* __restrict__ done_count, int C) {
    if (done_count != nullptr && *done_count >= C) return;
    extern __shared__ __align__(16) double bd_smem[];
    double* sX = bd_smem;                                  // [2][BD_ROWS][BD_XS]
    double* sV = bd_smem + 2 * BD_ROWS * BD_XS;            // [2][BD_KC][BD_VS]
    double* sW = sV + 2 * BD_KC * BD_VS;                   // [8][BC]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 row0 = (i64)blockIdx.x * BD_ROWS;
    const int nch = (int)((p + BD_KC - 1) / BD_KC);
    double acc[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    double rx[16], rv[2];
    auto fetch = [&](int ch) {
        const i64 j = (i64)ch * BD_KC + lane;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const i64 i = row0 + warp + 8 * k;
            rx[k] = (i < n && j < p) ? X[i * p + j] : 0.0;
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid * 2 + e;
            const i64 jj = (i64)ch * BD_KC + (idx >> 4);
            rv[e] = (jj < p) ? sv_b[jj * BC + (idx & 15)] : 0.0;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int k = 0; k < 16; ++k) sX[(buf * BD_ROWS + warp + 8 * k) * BD_XS + lane] = rx[k];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid * 2 + e;
            sV[(buf * BD_KC + (idx >> 4)) * BD_VS + (idx & 15)] = rv[e];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int ch = 0; ch < nch; ++ch) {
        const bool more = ch + 1 < nch;
        if (more) fetch(ch + 1);
        const double* xa = sX + (buf * BD_ROWS + warp * 16 + (lane >> 2)) * BD_XS + (lane & 3);
        const double* vb = sV + (buf * BD_KC + (lane & 3)) * BD_VS + (lane >> 2);
#pragma unroll
        for (int k4 = 0; k4 < BD_KC / 4; ++k4) {
            const double a0 = xa[k4 * 4], a1 = xa[8 * BD_XS + k4 * 4];
            const double b0 = vb[k4 * 4 * BD_VS], b1 = vb[k4 * 4 * BD_VS + 8];
            bdmma(acc[0][0][0], acc[0][0][1], a0, b0);
            bdmma(acc[0][1][0], acc[0][1][1], a0, b1);
            bdmma(acc[1][0][0], acc[1][0][1], a1, b0);
            bdmma(acc[1][1][0], acc[1][1][1], a1, b1);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    // epilogue: element (row = warp*16 + mi*8 + lane/4, chain = ni*8 + (lane%4)*2 + e)
    double colsum[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
        const i64 i = row0 + warp * 16 + mi * 8 + (lane >> 2);
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            const int c = ni * 8 + (lane & 3) * 2;
            double u0 = acc[mi][ni][0] + shift_b[c], u1 = acc[mi][ni][1] + shift_b[c + 1];
            if (i < n) {
                if (MODE == 1) {
                    const double2 om = *reinterpret_cast<const double2*>(omega_b + i * BC + c);
                    u0 *= om.x; u1 *= om.y;
                    colsum[ni][0] += u0; colsum[ni][1] += u1;
                }
                *reinterpret_cast<double2*>(out + i * BC + c) = make_double2(u0, u1);
            }
        }
    }
    if (MODE == 1) {
        // sum over the 16 rows of the warp (lane bits 2..4), then over the 8 warps in warp order
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double v = colsum[ni][e];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if ((lane >> 2) == 0) sW[warp * BC + ni * 8 + (lane & 3) * 2 + e] = v;
            }
        __syncthreads();
        if (tid < BC) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += sW[w * BC + tid];
            red_w_b[(i64)blockIdx.x * BC + tid] = t;
        }
    }
}

// T = X' W for a block of 128 columns and one block of rows: part_b[blk][j][c]
constexpr int BT_COLS = 128, BT_KR = 32, BT_XS = BT_COLS + 4, BT_WS = BC + 4;

__global__ void __launch_bounds__(256)
k_batch_tdot(const double* __restrict__ X, i64 n, i64 p, const double* __restrict__ W, int nblk,
             double* __restrict__ part_b, const int* __restrict__ done_count, int C) {
    if (done_count != nullptr && *done_count >= C) return;
    extern __shared__ __align__(16) double bt_smem[];
    double* sX = bt_smem;                                  // [2][BT_KR][BT_XS]
    double* sW = bt_smem + 2 * BT_KR * BT_XS;              // [2][BT_KR][BT_WS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 cb = (i64)blockIdx.x * BT_COLS;
    const int by = blockIdx.y;
    const i64 r_lo = n * by / nblk, r_hi = n * (by + 1) / nblk;
    const int nch = (int)((r_hi - r_lo + BT_KR - 1) / BT_KR);
    double acc[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    double rx[16], rw[2];
    const int col = tid & 127, rsub = tid >> 7;
    auto fetch = [&](int ch) {
        const i64 j = cb + col;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const i64 i = r_lo + (i64)ch * BT_KR + rsub + 2 * k;
            rx[k] = (i < r_hi && j < p) ? X[i * p + j] : 0.0;
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid * 2 + e;
            const i64 i = r_lo + (i64)ch * BT_KR + (idx >> 4);
            rw[e] = (i < r_hi) ? W[i * BC + (idx & 15)] : 0.0;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int k = 0; k < 16; ++k) sX[(buf * BT_KR + rsub + 2 * k) * BT_XS + col] = rx[k];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid * 2 + e;
            sW[(buf * BT_KR + (idx >> 4)) * BT_WS + (idx & 15)] = rw[e];
        }
    };
    if (nch > 0) { fetch(0); stash(0); }
    __syncthreads();
    int buf = 0;
    for (int ch = 0; ch < nch; ++ch) {
        const bool more = ch + 1 < nch;
        if (more) fetch(ch + 1);
        const double* xa = sX + (buf * BT_KR + (lane & 3)) * BT_XS + warp * 16 + (lane >> 2);
        const double* wb = sW + (buf * BT_KR + (lane & 3)) * BT_WS + (lane >> 2);
#pragma unroll
        for (int k4 = 0; k4 < BT_KR / 4; ++k4) {
            const double a0 = xa[k4 * 4 * BT_XS], a1 = xa[k4 * 4 * BT_XS + 8];
            const double b0 = wb[k4 * 4 * BT_WS], b1 = wb[k4 * 4 * BT_WS + 8];
            bdmma(acc[0][0][0], acc[0][0][1], a0, b0);
            bdmma(acc[0][1][0], acc[0][1][1], a0, b1);
            bdmma(acc[1][0][0], acc[1][0][1], a1, b0);
            bdmma(acc[1][1][0], acc[1][1][1], a1, b1);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
        const i64 j = cb + warp * 16 + mi * 8 + (lane >> 2);
        if (j < p) {
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int c = ni * 8 + (lane & 3) * 2;
                *reinterpret_cast<double2*>(part_b + ((i64)by * p + j) * BC + c) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
        }
    }
}

// traw_b[c][0] = sum of the W column sums ; traw_b[c][1 + j] = sum over row blocks of part_b[blk][j][c]
__global__ void k_batch_collect(const double* __restrict__ part_b, int nblk, i64 p, const double* __restrict__ red_w_b, int ndot,
                                double* __restrict__ traw_b, const int* __restrict__ done_count, int C) {
    if (done_count != nullptr && *done_count >= C) return;
    const i64 total = p * BC;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
        const i64 j = idx >> 4;
        const int c = (int)(idx & 15);
        double t0 = 0.0, t1 = 0.0;
        int b = 0;
        for (; b + 1 < nblk; b += 2) { t0 += part_b[((i64)b * p + j) * BC + c]; t1 += part_b[((i64)(b + 1) * p + j) * BC + c]; }
        if (b < nblk) t0 += part_b[((i64)b * p + j) * BC + c];
        traw_b[(i64)c * (p + 1) + 1 + j] = t0 + t1;
    }
    if (blockIdx.x == 0 && red_w_b != nullptr) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int c = warp; c < BC; c += nw) {
            double t = 0.0;
            for (int k = lane; k < ndot; k += 32) t += red_w_b[(i64)k * BC + c];
            t = warp_sum(t);
            if (lane == 0) traw_b[(i64)c * (p + 1)] = t;
        }
    }
}
// column sums of an [n][BC] array -> red_w_b[grid][BC] (used when W does not come out of k_batch_dot<1>)
__global__ void __launch_bounds__(256) k_batch_colsum(const double* __restrict__ W, i64 n, double* __restrict__ red_w_b) {
    __shared__ double sm[16][BC];
    const int c = threadIdx.x & 15, r = threadIdx.x >> 4;
    double t = 0.0;
    for (i64 i = (i64)blockIdx.x * 16 + r; i < n; i += (i64)gridDim.x * 16) t += W[i * BC + c];
    sm[r][c] = t;
    __syncthreads();
    if (threadIdx.x < BC) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += sm[k][threadIdx.x];
        red_w_b[(i64)blockIdx.x * BC + threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------
// chain-major <-> chain-fastest transposes (host arrays are [C][len])
__global__ void k_batch_to_inner(const double* __restrict__ src, int C, i64 len, i64 src_stride, double* __restrict__ dst) {
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < len * BC; idx += (i64)gridDim.x * blockDim.x) {
        const i64 i = idx >> 4;
        const int c = (int)(idx & 15);
        dst[idx] = (c < C) ? src[(i64)c * src_stride + i] : 0.0;
    }
}
__global__ void k_batch_to_outer(const double* __restrict__ src, int C, i64 len, i64 dst_stride, double* __restrict__ dst) {
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < len * C; idx += (i64)gridDim.x * blockDim.x) {
        const int c = (int)(idx / len);
        const i64 i = idx - (i64)c * len;
        dst[(i64)c * dst_stride + i] = src[i * BC + c];
    }
}

// sqrt(omega) * eps1 per chain (cg_sampler.py:66); eps injected ([n][BC]) or Philox keyed by (seed_c, offset_c, global row)
__global__ void k_batch_rhs(i64 n, const double* __restrict__ omega_b, const double* __restrict__ eps_b, int philox,
                            const uint64_t* __restrict__ seeds, const uint64_t* __restrict__ offsets, i64 row_offset, int C,
                            double* __restrict__ out) {
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n * BC; idx += (i64)gridDim.x * blockDim.x) {
        const i64 i = idx >> 4;
        const int c = (int)(idx & 15);
        double v = 0.0;
        if (c < C) {
            double e;
            if (philox) { RandStream rs; rs.init(seeds[c], offsets[c], (uint64_t)(row_offset + i), STREAM_EPS1); e = rs.normal(); }
            else e = eps_b[idx];
            v = sqrt(omega_b[idx]) * e;
        }
        out[idx] = v;
    }
}

// gather operand of X V and the shift of every chain from chain-major P-vectors: sv = (scale? scale.v : v)
__global__ void __launch_bounds__(512)
k_batch_prepare(const double* __restrict__ v, const double* __restrict__ scale, i64 Ps, i64 P, int icpt,
                const double* __restrict__ coff, double* __restrict__ sv_b, double* __restrict__ shift_b) {
    __shared__ double sm[33];
    const int c = blockIdx.x;
    double acc = 0.0;
    for (i64 j = threadIdx.x; j < P; j += blockDim.x) {
        const double x = scale ? __dmul_rn(scale[(i64)c * Ps + j], v[(i64)c * Ps + j]) : v[(i64)c * Ps + j];
        if (j >= icpt) sv_b[(j - icpt) * BC + c] = x;
        acc += (j < icpt) ? x : -coff[j - icpt] * x;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) shift_b[c] = acc;
}

// ------------------------------------------------------------------------------------------
// The P-side of every chain: one CTA per chain, all reductions block-local (P is a few thousand here)
enum { BP_INIT = 0, BP_RESID = 1, BP_ITER = 2, BP_FINAL = 3 };
struct BatchPArgs {
    int mode, C, icpt, philox, maxiter;
    i64 P, p, Ps;
    double atol;
    const double* coff; const double* traw_b;
    double *x, *r, *pv, *q, *b, *s, *D, *pps, *z, *x0, *eps2, *out;
    double *sv_b, *shift_b;
    CgScalars* cg; int* done_count;
    const uint64_t* seeds; const uint64_t* offsets;
};

__global__ void __launch_bounds__(512) k_batch_pside(const BatchPArgs a) {
    __shared__ double sm[33];
    const int c = blockIdx.x, tid = threadIdx.x;
    CgScalars* st = a.cg + c;
    const i64 o = (i64)c * a.Ps;
    const double* t = a.traw_b + (i64)c * (a.p + 1);
    const double sw = t[0];
    auto tent = [&](i64 j) { return (j < a.icpt) ? sw : __dsub_rn(t[1 + (j - a.icpt)], __dmul_rn(sw, a.coff[j - a.icpt])); };
    // writes sv = s.p (gather operand, chain-fastest layout) and the shift; returns nothing
    auto emit_dir = [&](bool first, double beta) {
        double acc = 0.0;
        for (i64 j = tid; j < a.P; j += blockDim.x) {
            const double pj = first ? a.r[o + j] : __dadd_rn(__dmul_rn(beta, a.pv[o + j]), a.r[o + j]);
            a.pv[o + j] = pj;
            const double xs = __dmul_rn(a.s[o + j], pj);
            if (j >= a.icpt) a.sv_b[(j - a.icpt) * BC + c] = xs;
            acc += (j < a.icpt) ? xs : -a.coff[j - a.icpt] * xs;
        }
        acc = block_sum(acc, sm);
        if (tid == 0) a.shift_b[c] = acc;
    };
    if (a.mode == BP_INIT) {
        // b = s.(z + X' sqrt(omega) eps1 + pps.eps2) ; D = (s.pps)^2 ; x = x0 / s ; gather operand of the initial residual
        double bb = 0.0, sh = 0.0;
        for (i64 j = tid; j < a.P; j += blockDim.x) {
            double e2;
            if (a.philox) { RandStream rs; rs.init(a.seeds[c], a.offsets[c], (uint64_t)j, STREAM_EPS2); e2 = rs.normal(); }
            else e2 = a.eps2[o + j];
            const double v = __dadd_rn(tent(j), __dmul_rn(a.pps[o + j], e2));
            const double bj = __dmul_rn(a.s[o + j], __dadd_rn(a.z[o + j], v));
            a.b[o + j] = bj;
            const double sp = __dmul_rn(a.s[o + j], a.pps[o + j]);
            a.D[o + j] = __dmul_rn(sp, sp);
            const double xj = a.x0[o + j] / a.s[o + j];
            a.x[o + j] = xj;
            const double xs = __dmul_rn(a.s[o + j], xj);
            if (j >= a.icpt) a.sv_b[(j - a.icpt) * BC + c] = xs;
            sh += (j < a.icpt) ? xs : -a.coff[j - a.icpt] * xs;
            bb += bj * bj;
        }
        bb = block_sum(bb, sm);
        sh = block_sum(sh, sm);
        if (tid == 0) {
            const double bn = sqrt(bb);
            st->bnorm = bn;
            st->atol_eff = (bn > 0.0) ? (a.atol / bn) * bn : a.atol;
            st->iter = 0; st->maxiter = a.maxiter; st->rho[0] = 0.0; st->rho[1] = 0.0; st->rnorm = 0.0;
            st->done = (bn == 0.0) ? 3 : 0;
            if (bn == 0.0) atomicAdd(a.done_count, 1);
            a.shift_b[c] = sh;
        }
        return;
    }
    if (a.mode == BP_FINAL) {
        const bool zero_rhs = (st->done == 3);
        for (i64 j = tid; j < a.P; j += blockDim.x) a.out[o + j] = zero_rhs ? 0.0 : __dmul_rn(a.s[o + j], a.x[o + j]);
        return;
    }
    if (st->done) return;
    if (a.mode == BP_RESID) {
        // r = b - A x ; then the stop test and the direction of iteration 0
        double rr = 0.0;
        for (i64 j = tid; j < a.P; j += blockDim.x) {
            const double qj = __dadd_rn(__dmul_rn(a.D[o + j], a.x[o + j]), __dmul_rn(a.s[o + j], tent(j)));
            const double rj = __dsub_rn(a.b[o + j], qj);
            a.r[o + j] = rj;
            rr += rj * rj;
        }
        rr = block_sum(rr, sm);
        const double rn = sqrt(rr);
        int done = 0;
        if (0 >= st->maxiter) done = 2; else if (rn < st->atol_eff) done = 1;
        if (!done) emit_dir(true, 0.0);
        if (tid == 0) {
            st->rnorm = rn;
            if (done) { st->done = done; atomicAdd(a.done_count, 1); } else st->rho[0] = rr;
        }
        return;
    }
    // BP_ITER: q = D.p + s.t ; alpha ; x, r ; stop test ; direction
    const int it0 = st->iter;
    const double rho_old = st->rho[it0 & 1];
    double pq = 0.0;
    for (i64 j = tid; j < a.P; j += blockDim.x) {
        const double pj = a.pv[o + j];
        const double qj = __dadd_rn(__dmul_rn(a.D[o + j], pj), __dmul_rn(a.s[o + j], tent(j)));
        a.q[o + j] = qj;
        pq += pj * qj;
    }
    pq = block_sum(pq, sm);
    const double alpha = rho_old / pq;
    double rr = 0.0;
    for (i64 j = tid; j < a.P; j += blockDim.x) {
        a.x[o + j] = __dadd_rn(a.x[o + j], __dmul_rn(alpha, a.pv[o + j]));
        const double rj = __dsub_rn(a.r[o + j], __dmul_rn(alpha, a.q[o + j]));
        a.r[o + j] = rj;
        rr += rj * rj;
    }
    rr = block_sum(rr, sm);
    const int it = it0 + 1;
    const double rn = sqrt(rr);
    int done = 0;
    if (it >= st->maxiter) done = 2; else if (rn < st->atol_eff) done = 1;
    if (!done) emit_dir(false, rr / rho_old);
    if (tid == 0) {
        st->iter = it; st->rnorm = rn;
        if (done) { st->done = done; atomicAdd(a.done_count, 1); } else st->rho[it & 1] = rr;
    }
}

// ------------------------------------------------------------------------------------------
// host side
static BatchWork* bw_of(bb_mat* m) { return (BatchWork*)m->batch; }

static int bgrid(i64 n) { i64 g = (n + 255) / 256; if (g < 1) g = 1; if (g > 4096) g = 4096; return (int)g; }

extern "C" int bb_batch_free(bb_mat* m) {
    if (!m || !m->batch) return BB_OK;
    BatchWork* w = bw_of(m);
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    void* ptrs[] = {w->omega_b, w->w_b, w->u_b, w->stage, w->part_b, w->red_w_b, w->traw_b, w->x, w->r, w->pv, w->q, w->b, w->s, w->D,
                    w->pps, w->z, w->x0, w->eps2, w->out, w->sv_b, w->shift_b, w->cg, w->done_count, w->seeds, w->offsets, w->ll_b};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (w->cg_host) cudaFreeHost(w->cg_host);
    free(w);
    m->batch = nullptr;
    return BB_OK;
}

extern "C" int bb_batch_init(bb_mat* m, int n_chains) {
    BB_ARG(m != nullptr, "mat");
    BB_ARG(n_chains >= 1 && n_chains <= BC, "1 <= n_chains <= 16");
    BB_ARG(!m->is_sparse, "the batched multi-chain path is implemented for dense designs");
    bb_ctx* ctx = m->ctx;
    BB_CUDA(cudaSetDevice(ctx->device));
    if (m->batch) BB_TRY(bb_batch_free(m));
    BatchWork* w = (BatchWork*)calloc(1, sizeof(BatchWork));
    m->batch = w;
    w->C = n_chains;
    w->Ps = (m->P + 8) & ~(i64)7;
    const i64 n = m->n > 0 ? m->n : 1, p = m->p > 0 ? m->p : 1;
    w->nblk = (int)((ctx->sm_count * 4 + (p + BT_COLS - 1) / BT_COLS - 1) / ((p + BT_COLS - 1) / BT_COLS));   // ~4 CTAs per SM in the X'W kernel
    if (w->nblk > n / 32 + 1) w->nblk = (int)(n / 32 + 1);
    if (w->nblk < 1) w->nblk = 1;
    w->ndot = (int)((n + BD_ROWS - 1) / BD_ROWS);
    cudaStream_t st = ctx->stream;
#define BALLOC(ptr, count)                                                                                  \
    do {                                                                                                    \
        BB_CUDA(cudaMalloc((void**)&(ptr), (size_t)(count) * sizeof(*(ptr))));                              \
        BB_CUDA(cudaMemsetAsync((ptr), 0, (size_t)(count) * sizeof(*(ptr)), st));                           \
    } while (0)
    BALLOC(w->omega_b, n * BC); BALLOC(w->w_b, n * BC); BALLOC(w->u_b, n * BC);
    BALLOC(w->stage, (i64)BC * (n > w->Ps ? n : w->Ps));
    BALLOC(w->part_b, (i64)w->nblk * p * BC);
    BALLOC(w->red_w_b, (i64)(w->ndot > 4096 ? w->ndot : 4096) * BC);
    BALLOC(w->traw_b, (i64)BC * (p + 1));
    double** pv[] = {&w->x, &w->r, &w->pv, &w->q, &w->b, &w->s, &w->D, &w->pps, &w->z, &w->x0, &w->eps2, &w->out};
    for (auto pp : pv) BALLOC(*pp, (i64)BC * w->Ps);
    BALLOC(w->sv_b, p * BC); BALLOC(w->shift_b, BC);
    BALLOC(w->cg, BC); BALLOC(w->done_count, 1); BALLOC(w->seeds, BC); BALLOC(w->offsets, BC); BALLOC(w->ll_b, BC);
#undef BALLOC
    BB_CUDA(cudaMallocHost((void**)&w->cg_host, BC * sizeof(CgScalars)));
    static BBDeviceOnce attr = {{0, 0, 0, 0}};
    if (attr.first(ctx->device)) {
        BB_CUDA(cudaFuncSetAttribute(k_batch_dot<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        BB_CUDA(cudaFuncSetAttribute(k_batch_dot<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        BB_CUDA(cudaFuncSetAttribute(k_batch_tdot, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    BB_CUDA(cudaStreamSynchronize(st));
    return BB_OK;
}

static size_t dot_smem() { return (size_t)(2 * BD_ROWS * BD_XS + 2 * BD_KC * BD_VS + 8 * BC) * sizeof(double); }
static size_t tdot_smem() { return (size_t)(2 * BT_KR * BT_XS + 2 * BT_KR * BT_WS) * sizeof(double); }

// U (mode 0, into u_b) or W = omega o U (mode 1, into w_b + red_w_b) from sv_b / shift_b
static int batch_dot(bb_mat* m, int mode, const int* done) {
    BatchWork* w = bw_of(m);
    bb_ctx* ctx = m->ctx;
    if (m->n == 0) return BB_OK;
    if (mode == 0) k_batch_dot<0><<<w->ndot, 256, dot_smem(), ctx->stream>>>(m->Xd, m->n, m->p, w->sv_b, w->shift_b, nullptr, w->u_b, nullptr, done, w->C);
    else k_batch_dot<1><<<w->ndot, 256, dot_smem(), ctx->stream>>>(m->Xd, m->n, m->p, w->sv_b, w->shift_b, w->omega_b, w->w_b, w->red_w_b, done, w->C);
    BB_LAUNCHED(ctx);
    return BB_OK;
}
// traw_b = [sum W; X' W] per chain, summed over the row shards.  have_sums: red_w_b holds w->ndot partial sums from k_batch_dot<1>
static int batch_tdot(bb_mat* m, const double* W, bool have_sums, const int* done) {
    BatchWork* w = bw_of(m);
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    int nsum = w->ndot;
    if (!have_sums) {
        nsum = 1024;
        k_batch_colsum<<<nsum, 256, 0, st>>>(W, m->n, w->red_w_b);
        BB_LAUNCHED(ctx);
    }
    if (m->p > 0) {
        dim3 grid((unsigned)((m->p + BT_COLS - 1) / BT_COLS), (unsigned)w->nblk);
        k_batch_tdot<<<grid, 256, tdot_smem(), st>>>(m->Xd, m->n, m->p, W, w->nblk, w->part_b, done, w->C);
        BB_LAUNCHED(ctx);
    }
    k_batch_collect<<<bgrid(m->p * BC), 256, 0, st>>>(w->part_b, w->nblk, m->p, w->red_w_b, nsum, w->traw_b, done, w->C);
    BB_LAUNCHED(ctx);
    BB_TRY(bb_allreduce_dev(ctx, w->traw_b, (i64)BC * (m->p + 1)));
    return BB_OK;
}

static int upload_chain_major(bb_mat* m, const double* host, i64 len, double* dst /*[BC][Ps]*/) {
    BatchWork* w = bw_of(m);
    for (int c = 0; c < w->C; ++c)
        BB_CUDA(cudaMemcpyAsync(dst + (i64)c * w->Ps, host + (i64)c * len, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
    return BB_OK;
}
// host [C][n] -> device [n][BC]
static int upload_n_side(bb_mat* m, const double* host, double* dst) {
    BatchWork* w = bw_of(m);
    cudaStream_t st = m->ctx->stream;
    BB_CUDA(cudaMemcpyAsync(w->stage, host, (size_t)w->C * m->n * sizeof(double), cudaMemcpyHostToDevice, st));
    k_batch_to_inner<<<bgrid(m->n * BC), 256, 0, st>>>(w->stage, w->C, m->n, m->n, dst);
    BB_LAUNCHED(m->ctx);
    return BB_OK;
}
static int download_n_side(bb_mat* m, const double* src, double* host) {
    BatchWork* w = bw_of(m);
    cudaStream_t st = m->ctx->stream;
    k_batch_to_outer<<<bgrid(m->n * w->C), 256, 0, st>>>(src, w->C, m->n, m->n, w->stage);
    BB_LAUNCHED(m->ctx);
    BB_CUDA(cudaMemcpyAsync(host, w->stage, (size_t)w->C * m->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    return BB_OK;
}

extern "C" int bb_batch_set_obs_prec(bb_mat* m, const double* omega) {
    BB_ARG(m && m->batch && omega, "mat (bb_batch_init first) / omega");
    BB_CUDA(cudaSetDevice(m->ctx->device));
    BB_TRY(upload_n_side(m, omega, bw_of(m)->omega_b));
    BB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return BB_OK;
}
extern "C" int bb_batch_get_obs_prec(bb_mat* m, double* omega_out) {
    BB_ARG(m && m->batch && omega_out, "mat (bb_batch_init first) / out");
    BB_CUDA(cudaSetDevice(m->ctx->device));
    BB_TRY(download_n_side(m, bw_of(m)->omega_b, omega_out));
    BB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return BB_OK;
}

// out[c][:] = X v_c   (v: [C][P], out: [C][n])
extern "C" int bb_dot_batched(bb_mat* m, const double* v, double* out) {
    BB_ARG(m && m->batch && v && out, "mat (bb_batch_init first) / v / out");
    BatchWork* w = bw_of(m);
    bb_ctx* ctx = m->ctx;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    BB_TRY(upload_chain_major(m, v, m->P, w->x0));
    k_batch_prepare<<<w->C, 512, 0, ctx->stream>>>(w->x0, nullptr, w->Ps, m->P, m->add_intercept, m->col_offset, w->sv_b, w->shift_b);
    BB_LAUNCHED(ctx);
    BB_TRY(batch_dot(m, 0, nullptr));
    BB_TRY(download_n_side(m, w->u_b, out));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    timer_.commit();
    return BB_OK;
}

__global__ void k_batch_tfinish(const double* __restrict__ traw_b, i64 p, int icpt, const double* __restrict__ coff, i64 Ps, int C,
                                double* __restrict__ out) {
    const int c = blockIdx.x;
    const double* t = traw_b + (i64)c * (p + 1);
    const double sw = t[0];
    for (i64 j = threadIdx.x; j < p + icpt; j += blockDim.x)
        out[(i64)c * Ps + j] = (j < icpt) ? sw : t[1 + (j - icpt)] - sw * coff[j - icpt];
}

// out[c][:] = X' w_c   (w: [C][n], out: [C][P])
extern "C" int bb_tdot_batched(bb_mat* m, const double* wv, double* out) {
    BB_ARG(m && m->batch && wv && out, "mat (bb_batch_init first) / w / out");
    BatchWork* w = bw_of(m);
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    BB_TRY(upload_n_side(m, wv, w->u_b));
    BB_TRY(batch_tdot(m, w->u_b, false, nullptr));
    k_batch_tfinish<<<w->C, 512, 0, st>>>(w->traw_b, m->p, m->add_intercept, m->col_offset, w->Ps, w->C, w->out);
    BB_LAUNCHED(ctx);
    for (int c = 0; c < w->C; ++c)
        BB_CUDA(cudaMemcpyAsync(out + (i64)c * m->P, w->out + (i64)c * w->Ps, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToHost, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    return BB_OK;
}

static BatchPArgs pargs(bb_mat* m, int mode, double atol, int maxiter, int philox) {
    BatchWork* w = bw_of(m);
    BatchPArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = mode; a.C = w->C; a.icpt = m->add_intercept; a.philox = philox; a.maxiter = maxiter;
    a.P = m->P; a.p = m->p; a.Ps = w->Ps; a.atol = atol;
    a.coff = m->col_offset; a.traw_b = w->traw_b;
    a.x = w->x; a.r = w->r; a.pv = w->pv; a.q = w->q; a.b = w->b; a.s = w->s; a.D = w->D; a.pps = w->pps; a.z = w->z; a.x0 = w->x0;
    a.eps2 = w->eps2; a.out = w->out; a.sv_b = w->sv_b; a.shift_b = w->shift_b; a.cg = w->cg; a.done_count = w->done_count;
    a.seeds = w->seeds; a.offsets = w->offsets;
    return a;
}

// The batched form of bb_cg_sample: C coefficient draws on one design, CG iterations in lock-step.
//   omega: [C][n] host, or NULL for the resident batched precisions (bb_batch_set_obs_prec / bb_pg_from_coef_batched)
//   prior_prec_sqrt, z, x0, precond_scale, eps2, coef_out: [C][P] ; eps1: [C][n] ; seeds, offsets, n_iter, info: [C]
extern "C" int bb_cg_sample_batched(bb_mat* m, const double* omega, const double* prior_prec_sqrt, const double* z,
                                    const double* x0, const double* precond_scale, double atol, int maxiter, int noise_mode,
                                    const double* eps1, const double* eps2, const uint64_t* seeds, const uint64_t* offsets,
                                    double* coef_out, int* n_iter, int* info) {
    BB_ARG(m && m->batch, "bb_batch_init first");
    BB_ARG(prior_prec_sqrt && z && x0 && precond_scale && coef_out, "null pointer");
    BB_ARG(noise_mode == BB_NOISE_PHILOX ? (seeds && offsets) : (eps1 && eps2), "noise arguments");
    BB_ARG(maxiter >= 0, "maxiter");
    BatchWork* w = bw_of(m);
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    const int C = w->C, philox = (noise_mode == BB_NOISE_PHILOX);
    if (omega) BB_TRY(upload_n_side(m, omega, w->omega_b));
    BB_TRY(upload_chain_major(m, prior_prec_sqrt, m->P, w->pps));
    BB_TRY(upload_chain_major(m, z, m->P, w->z));
    BB_TRY(upload_chain_major(m, x0, m->P, w->x0));
    BB_TRY(upload_chain_major(m, precond_scale, m->P, w->s));
    if (philox) {
        BB_CUDA(cudaMemcpyAsync(w->seeds, seeds, C * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemcpyAsync(w->offsets, offsets, C * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    } else {
        BB_TRY(upload_n_side(m, eps1, w->u_b));
        BB_TRY(upload_chain_major(m, eps2, m->P, w->eps2));
    }
    BB_CUDA(cudaMemsetAsync(w->done_count, 0, sizeof(int), st));
    // right-hand side: X'(sqrt(omega) eps1) for every chain
    k_batch_rhs<<<bgrid(m->n * BC), 256, 0, st>>>(m->n, w->omega_b, w->u_b, philox, w->seeds, w->offsets, m->row_offset, C, w->w_b);
    BB_LAUNCHED(ctx);
    BB_TRY(batch_tdot(m, w->w_b, false, nullptr));
    {
        BatchPArgs a = pargs(m, BP_INIT, atol, maxiter, philox);
        k_batch_pside<<<C, 512, 0, st>>>(a);
        BB_LAUNCHED(ctx);
    }
    // initial residual
    BB_TRY(batch_dot(m, 1, w->done_count));
    BB_TRY(batch_tdot(m, w->w_b, true, w->done_count));
    {
        BatchPArgs a = pargs(m, BP_RESID, atol, maxiter, philox);
        k_batch_pside<<<C, 512, 0, st>>>(a);
        BB_LAUNCHED(ctx);
    }
    int total = 0;
    int chunk = (ctx->opt_cg_chunk > 0) ? (int)ctx->opt_cg_chunk : (w->last_n_iter > 0 ? w->last_n_iter + 1 : 8);
    const BatchPArgs it_args = pargs(m, BP_ITER, atol, maxiter, philox);
    for (;;) {
        int todo = chunk;
        if (total + todo > maxiter) todo = maxiter - total;
        if (todo < 1) todo = 1;
        for (int k = 0; k < todo; ++k) {
            BB_TRY(batch_dot(m, 1, w->done_count));
            BB_TRY(batch_tdot(m, w->w_b, true, w->done_count));
            k_batch_pside<<<C, 512, 0, st>>>(it_args);
            BB_LAUNCHED(ctx);
        }
        total += todo;
        BB_CUDA(cudaMemcpyAsync(w->cg_host, w->cg, C * sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        bool all = true;
        for (int c = 0; c < C; ++c) all = all && (w->cg_host[c].done != 0);
        if (all || total >= maxiter) break;
        chunk = (ctx->opt_cg_chunk > 0) ? (int)ctx->opt_cg_chunk : 4;
    }
    {
        BatchPArgs a = pargs(m, BP_FINAL, atol, maxiter, philox);
        k_batch_pside<<<C, 512, 0, st>>>(a);
        BB_LAUNCHED(ctx);
    }
    for (int c = 0; c < C; ++c)
        BB_CUDA(cudaMemcpyAsync(coef_out + (i64)c * m->P, w->out + (i64)c * w->Ps, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(w->cg_host, w->cg, C * sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    int mx = 0;
    for (int c = 0; c < C; ++c) {
        const int done = w->cg_host[c].done;
        if (n_iter) n_iter[c] = w->cg_host[c].iter;
        if (info) info[c] = (done == 1 || done == 3) ? 0 : maxiter;
        if (w->cg_host[c].iter > mx) mx = w->cg_host[c].iter;
    }
    w->last_n_iter = mx;
    return BB_OK;
}

// defined in bb_rand.cu: omega_b[i][c] ~ PG(n_trial_i, eta_b[i][c]) and the per-chain log-likelihoods
int bb_batch_pg_launch(bb_mat* m, const double* eta_b, int C, const uint64_t* seeds_dev, const uint64_t* offsets_dev,
                       double* omega_b, double* red_scratch, double* ll_b);

// omega_c | beta_c for every chain, batched: eta = X beta (tensor-core product), PG draws on per-chain Philox streams keyed
// by the global row index, logistic log-likelihood per chain.  coef: [C][P] host, or NULL for the last batched CG draw.
extern "C" int bb_pg_from_coef_batched(bb_mat* m, const double* coef, const uint64_t* seeds, const uint64_t* offsets, double* loglik) {
    BB_ARG(m && m->batch && seeds && offsets, "mat (bb_batch_init first) / seeds / offsets");
    if (!m->has_outcome || m->is_linear) { bb_set_error("bb_pg_from_coef_batched needs a logit outcome (bb_set_outcome)"); return BB_ERR_STATE; }
    BatchWork* w = bw_of(m);
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    const double* src = w->out;
    if (coef) { BB_TRY(upload_chain_major(m, coef, m->P, w->x0)); src = w->x0; }
    BB_CUDA(cudaMemcpyAsync(w->seeds, seeds, w->C * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(w->offsets, offsets, w->C * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    k_batch_prepare<<<w->C, 512, 0, st>>>(src, nullptr, w->Ps, m->P, m->add_intercept, m->col_offset, w->sv_b, w->shift_b);
    BB_LAUNCHED(ctx);
    BB_TRY(batch_dot(m, 0, nullptr));
    BB_TRY(bb_batch_pg_launch(m, w->u_b, w->C, w->seeds, w->offsets, w->omega_b, w->stage, w->ll_b));
    BB_TRY(bb_allreduce_dev(ctx, w->ll_b, BC));
    double ll[BC];
    BB_CUDA(cudaMemcpyAsync(ll, w->ll_b, BC * sizeof(double), cudaMemcpyDeviceToHost, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    if (loglik) for (int c = 0; c < w->C; ++c) loglik[c] = ll[c];
    return BB_OK;
}


// timing hook of bb_time_kernel("batch_op"): W = Omega o (X V) and [sum W; X' W] for all chains (one operator application)
int bb_batch_time_op(bb_mat* m) {
    if (!m->batch) { bb_set_error("batch_op: bb_batch_init first"); return BB_ERR_STATE; }
    BB_TRY(batch_dot(m, 1, nullptr));
    BB_TRY(batch_tdot(m, bw_of(m)->w_b, true, nullptr));
    return BB_OK;
}
