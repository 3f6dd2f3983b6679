// Internal declarations shared by the translation units of libbbgpu.so.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <math.h>
#include "../../include/bbgpu.h"

typedef long long i64;

void bb_set_error(const char* fmt, ...);

#define BB_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            bb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return BB_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define BB_TRY(expr)                                                                       \
    do {                                                                                   \
        int rc_ = (expr);                                                                  \
        if (rc_ != BB_OK) return rc_;                                                      \
    } while (0)

#define BB_ARG(cond, msg)                                                                  \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            bb_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg);          \
            return BB_ERR_ARG;                                                             \
        }                                                                                  \
    } while (0)

// Function attributes (opt-in shared memory) are per device: a once-flag per device, not per process, so that a process
// holding contexts on several devices sets them on each.
struct BBDeviceOnce {
    unsigned long long mask[4];
    bool first(int device) {
        const int w = (device >> 6) & 3, b = device & 63;
        if ((mask[w] >> b) & 1ull) return false;
        mask[w] |= 1ull << b;
        return true;
    }
};

// counts a kernel launch and checks the launch error
#define BB_LAUNCHED(ctx)                                                                   \
    do {                                                                                   \
        (ctx)->launches++;                                                                 \
        cudaError_t e_ = cudaPeekAtLastError();                                            \
        if (e_ != cudaSuccess) {                                                           \
            bb_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return BB_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

// peer-memory exchange state (see bb_p2p.cu)
struct P2PState {
    unsigned long long seq;        // publications completed by this rank
    unsigned int blocks_done;      // publish: block counter
    unsigned int error;            // set when a wait timed out
    unsigned long long seq2;       // two-shot exchanges (fused CG iteration, bb_pside.cu) completed by this rank
};
struct P2PView {
    double* const* peer_base;      // [nranks] mapped exchange buffers (device array)
    P2PState* st;
    i64 cap;
    int nranks, rank;
    int variant;                   // protocol variant bits (option "p2p_variant"), see bb_internal.cuh
};

// ------------------------------------------------------------------------------------------
struct bb_ctx {
    int device;
    cudaStream_t stream;
    int sm_count;
    size_t smem_optin;     // max dynamic shared memory per block (opt-in)
    i64 launches;
    // options
    i64 opt_spmv_stage;    // 1: stage the gather vector in shared memory; 0: gather through L2
    i64 opt_slab_width;    // max doubles of the gather vector staged per CTA (0 = auto)
    i64 opt_bank_permute;  // 1: reorder nnz inside (segment x tile) pieces so that staged gathers avoid bank conflicts
    i64 opt_spmv_bulk;     // 1 (default): stage the window with cp.async.bulk + mbarrier; 0: cooperative copy loop
    i64 opt_sell_lmax;     // sliced format: fragment length cap (0 = automatic)
    i64 opt_spmv_variant;  // 1 (default): sliced lane-per-fragment kernel with 16-bit in-slab indices; 0: tile + segmented-scan kernel;
                           // 2: sub-warp-per-segment kernel on the canonical CSR / CSC image (gathers through L2)
    i64 opt_rowwise_max_nnz;  // with variant 1: matrices with at most this many nnz use variant 2 instead
    i64 opt_cg_chunk;      // CG iterations enqueued between host checks (0 = adaptive)
    i64 opt_use_graph;     // capture the CG iteration chunk into a CUDA graph
    i64 opt_cg_fused;      // 1 (default): the P-side of a CG iteration (+ its all-reduce) is one kernel (bb_pside.cu)
    i64 opt_pside_ctas;    // grid of that kernel (0 = automatic)
    i64 opt_pside_ll;      // 1 (default): the exchange inside that kernel sends flag-in-data lines (no flag round trips, no grid barriers
                           // around them) when the exchange buffer has room for 2 (p + 1) doubles per region; 0: data + flags
    i64 opt_pside_barrier; // 1 (default): its grid barriers arrive with a release reduction and poll at once; 0: fence + atomicAdd + fence
    i64 opt_pside_fold_ovf;     // overflow fragments folded inside the fused kernel: -1 automatic (few of them), 0 never, 1 always
    i64 opt_pside_collect_max;  // slab partials per column the fused kernel sums itself (0 = default 8); above: k_tdot_collect
    i64 opt_dense_stream;  // 1 (default): dense products through the one-pass TMA streaming kernel when a row pair fits in shared memory
    i64 opt_sell_slice_cost;   // sliced format, work partition: per-slice overhead in row equivalents (0 = default 3)
    i64 opt_sell_lpt;          // 1 (default): the warp strips of a section are filled longest-slice-first (balanced to within one short
                               // slice) and the slices laid out strip by strip; 0: contiguous cuts of the sorted slice order
    i64 opt_sell_partition;    // 0 (default): slab-aligned CTA ranges when that shortens the estimated critical path; 1: always
                               // equal-cost ranges (a CTA may stage two windows); 2: always slab-aligned
    i64 opt_pdl;           // 1 (default): the kernels of a fused CG iteration are launched with programmatic dependent launch
                           // (their launch and data-independent prologue overlap the tail of the previous kernel)
    i64 opt_uniform_carveout;  // 1 (default 0: measured, no gain): every kernel of the CG iteration asks for the maximum shared-memory carve-out, so
                               // that the SMs are never reconfigured (and drained) between the SpMV and the small kernels
    // communicator (NCCL via dlopen)
    void* nccl_handle;
    void* nccl_comm;
    int nranks, rank;
    int comm_local;          // communicator created by bb_comm_init_local: no NCCL, every exchange over peer memory
    // one-shot all-reduce over NVLink peer memory (bb_p2p.cu)
    void* p2p;
    int p2p_ready;
    i64 opt_allreduce_p2p;   // 1: use it when attached (default), 0: always NCCL
    i64 opt_p2p_variant;     // bit0: one system fence + relaxed flag stores; bit1: parallel flag polls;
                             // bit2: batched peer loads; bit3: flag stores issued by nranks threads
    // L2 flush scratch for bb_time_kernel
    void* flush_buf;
    size_t flush_bytes;
    // generic host pinned staging
    double* pinned;
    size_t pinned_bytes;
    // persistent device scratch (grown on demand) so that per-iteration entry points never cudaMalloc/cudaFree
    void* scratch[4];
    size_t scratch_bytes[4];
    // device-time accounting of the public entry points (CUDA events on `stream`)
    cudaEvent_t tev0, tev1;
    double dev_ms;
    int timer_depth;
};

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// A kernel launched through bb_launch(..., pdl = true) may start while the previous kernel of the stream is still
// running.  Rules every such kernel follows: (1) pdl_trigger() first (lets ITS successor be scheduled early);
// (2) before pdl_wait() it touches only data no kernel ever writes during a solve (the matrix formats) and its own
// shared memory; (3) pdl_wait() -- which returns once the previous kernel has completed and its writes are visible --
// precedes every other access, including the `done` flag, and is executed on every path (so completion of a kernel
// implies completion of all its predecessors).  Launched without the attribute both instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t bb_launch(bb_ctx* ctx, bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                    Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    if (pdl && ctx->opt_pdl != 0) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// every kernel of the CG iteration asks for the same (maximum) shared-memory carve-out: see opt_uniform_carveout
template <typename K>
static inline cudaError_t bb_prefer_max_smem(bb_ctx* ctx, K kernel) {
    if (ctx->opt_uniform_carveout == 0) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}
#endif

int bb_ctx_pinned(bb_ctx* ctx, size_t bytes, double** out);
int bb_ctx_scratch(bb_ctx* ctx, int slot, size_t bytes, void** out);
// bracket the device work of one entry point: begin() before the first enqueue, end() after the
// last enqueue and BEFORE the final stream sync, commit() after that sync.
struct BBTimer {
    bb_ctx* c; bool active;
    explicit BBTimer(bb_ctx* ctx) : c(ctx), active(false) {
        if (c->timer_depth++ == 0) { active = (cudaEventRecord(c->tev0, c->stream) == cudaSuccess); }
    }
    void end() { if (active) cudaEventRecord(c->tev1, c->stream); }
    void commit() {
        if (active) { float ms = 0.f; if (cudaEventElapsedTime(&ms, c->tev0, c->tev1) == cudaSuccess) c->dev_ms += ms; active = false; }
    }
    ~BBTimer() { c->timer_depth--; }
};
int bb_allreduce_dev(bb_ctx* ctx, double* dbuf, i64 count);   // in place, on ctx->stream
bool bb_p2p_allreduce(bb_ctx* c, double* dbuf, i64 count, const int* done_flag, int* rc_out);
int bb_p2p_free(bb_ctx* c);
i64 bb_p2p_capacity(bb_ctx* c);                               // doubles per exchange, 0 when not attached
bool bb_p2p_view(bb_ctx* c, i64 count, P2PView* out);          // true when the exchange can carry `count` doubles
bool bb_p2p_view2(bb_ctx* c, i64 count, P2PView* out);         // same for the two-shot exchange of the fused CG iteration
int bb_p2p_reduce_into(bb_ctx* c, double* dst, i64 count);      // wait + rank-ordered sum of the last publication

// ------------------------------------------------------------------------------------------
// Slab format: the nnz of a compressed (CSR or CSC) matrix regrouped so that every contiguous
// run of nnz gathers from one <=W-wide window of the input vector (a "slab"), which one CTA
// stages in shared memory.  Virtual segment v = slab * n_seg + seg.
struct TileMeta { int start, end, vlo, vhi; };   // nnz range; owns virtual segments [vlo, vhi)

struct SlabFmt {
    int   nslab;
    i64   n_seg;        // rows (dot format) or columns (Tdot format)
    i64   n_gather;     // length of the gathered vector
    int   W;            // slab width (gather indices [slab*W, slab*W+W))
    i64   nnz;
    int*  ptr;          // [nslab*n_seg + 1] nnz offsets of the virtual segments
    int*  idx;          // [nnz] gather index (global)
    double* val;        // [nnz] or NULL (pattern-only)
    bool  owns_arrays;  // false when aliasing the canonical CSR/CSC (nslab == 1)
    bool  has_val;      // sliced format: the matrix carries values (val itself is released after the build)
    bool  staged;       // the kernel stages the slab of the gather vector in shared memory
    int   ntiles;
    TileMeta* tiles;    // [ntiles]
    int*  head_seg;     // [ntiles] virtual segment continued from the previous tile, or -1
    int2* tmeta;        // [ntiles] {first owned virtual segment, number owned}
    int*  slab_tile0;   // [nslab+1] first tile of each slab
    int*  slab_nnz0;    // [nslab+1] first nnz of each slab (padded to a multiple of 4)
    int*  slab_nnz1;    // [nslab+1] one past the last nnz of each slab
    double* part;       // [nslab*n_seg (+ overflow slots + 1)] per-slab partial sums (output of the kernel)
    double* head_part;  // [ntiles]
    // ---- sliced format (spmv_variant 1, bb_sell.cu): fragments of <= SELL_LMAX nnz, one lane each ----
    int   variant;            // 0: tile + segmented-scan kernel (k_seg_spmv); 1: sliced lane-per-fragment kernel (k_sell_spmv);
                              // 2: sub-warp-per-segment kernel on the canonical arrays (k_csr_rowwise)
    int   nslices;            // slices of 32 fragments
    int   sl_lmax;            // fragment length cap chosen at build time (<= 256)
    unsigned* sl_off;         // [nslices+1] first row of every slice (sorted slice order) in the row stream; build-time only
    unsigned* sl_slot;        // [nslices*32] output slot of every lane's fragment (index into part)
    unsigned* sl_pairs;       // [32 * sl_off[nslices]] two 16-bit in-slab gather indices per word, blocked per lane
    double*   sl_vals;        // [64 * sl_off[nslices]] values in the same order, or NULL (pattern-only)
    int*  sl_slab_slice0;     // [nslab+1] first slice of each slab
    int*  sl_cta_slice0;      // [sl_ncta+1] first slice of each CTA's contiguous range (balanced by cost)
    int   sl_partition;       // 1: equal-cost CTA ranges (a CTA may stage two windows), 2: slab-aligned ranges (bb_sell.cu)
    int   sl_ncta, sl_nsec;   // sl_cta_slice0 packs [cta_sec0 (sl_ncta+1) | sec_slab (sl_nsec) | sec_wstart (sl_nsec*33)]
    int   n_ovf_pieces;       // virtual segments longer than SELL_LMAX (their extra fragments land in overflow slots)
    int*  ovf_piece;          // [n_ovf_pieces] virtual segment id
    int*  ovf_first;          // [n_ovf_pieces+1] first overflow slot (relative to part + nslab*n_seg)
    i64   n_ovf;              // overflow slots
};

struct CgScalars {
    double rho[2];
    double atol_eff;
    double bnorm;
    double rnorm;
    int    iter;      // completed CG iterations
    int    done;      // 0 running, 1 converged, 2 maxiter reached, 3 zero right-hand side
    int    maxiter;
    int    pad;
    unsigned long long bar_base;   // grid-barrier count at the start of the next fused P-side launch (bb_pside.cu)
};

struct bb_mat {
    bb_ctx* ctx;
    int  is_sparse, is_binary, add_intercept, centered;
    i64  n, p, P, nnz;          // local rows, predictors, P = p + intercept
    i64  row_offset, n_global;
    // canonical storage (bit-exact images of scipy's arrays)
    int *csr_ptr, *csr_idx; double* csr_val;
    int *csc_ptr, *csc_idx; double* csc_val;
    SlabFmt fdot, ftdot;
    double* col_offset;         // [p] (zeros when not centred)
    // dense storage: row-major [n x p], raw (no intercept / centring)
    double* Xd;
    // resident n-vectors
    double *omega, *n_trial, *n_success, *eta, *w_n, *u_n, *eps_n;
    int has_outcome, is_linear;
    double omega_scalar; int use_omega_scalar;   // linear model: omega = scalar * 1_n
    double* omega_scalar_dev;                    // device copy: kernels inside the captured CG graph read it here
    P2PView* p2p_view_dev;                       // device copy of the exchange view (kernel argument by pointer)
    int p2p_view_valid;
    // P-vectors
    double *sv_base;            // allocation behind sv (sv = sv_base + add_intercept, see bb_mat_alloc_work)
    double *v_P, *sv, *traw /*[1+p]*/, *t_P, *x, *r, *pvec, *q, *b, *s, *D, *pps, *z, *x0, *eps_P, *out_P;
    // reduction scratch
    double* red;                // [RED_SLOTS * RED_MAX]
    CgScalars* cg;              // device
    CgScalars* cg_host;         // pinned
    int last_n_iter;
    cudaGraphExec_t cg_graph; int cg_graph_launches; int cg_graph_scalar_mode; int cg_graph_fused;   // kernels per captured CG iteration
    int nred_w;                 // number of valid partials in red[RED_W]
    // dense Tdot partials [dense_nblk x p]
    double* dense_part; int dense_nblk; int dense_stream;    // dense_stream: the one-pass streaming kernel serves this matrix
    // device-resident P-side Gibbs state (bb_state_*): local scales, running summaries of the scaled coefficients
    double *st_lscale, *st_mean, *st_square, *st_prior_sd, *st_sums;
    int st_k, st_ready; double st_slab; long long st_n_avg;
    // cached z = X' kappa for the logit model (kappa = n_success - n_trial/2 is constant)
    double* zk; int zk_valid;
    unsigned long long* ps_bar;  // grid-barrier counter of the fused P-side kernel
    void* batch;                 // batched multi-chain work space (bb_batch.cu), or NULL
};

enum { RED_MAX = 1024, RED_SLOTS = 12 };
enum { RED_SHIFT = 0, RED_W = 1, RED_PQ = 2, RED_RR = 3, RED_BB = 4, RED_MISC = 5, RED_X0 = 6, RED_LL = 7, RED_STATE = 8 /* ..11 */ };

// device-side op pipeline (all on ctx->stream, no sync)
//   sv_shift: computes mat->sv (gather vector, p entries) and the shift partials from a P-vector
int bb_op_prepare(bb_mat* m, const double* vP, const double* scale /*nullable*/);
//   u_n = X sv + shift  (mode 0) ;  w_n = omega .* u_n with sum(w) partials (mode 1)
int bb_op_dot(bb_mat* m, int mode);
//   traw[0] = sum(w_n), traw[1+j] = (X' w_n)_j ; allreduced over shards
int bb_op_tdot(bb_mat* m, const double* w);
//   t_P from traw (intercept + centring)
int bb_op_tdot_finish(bb_mat* m, double* tP);

int bb_op_prepare_flag(bb_mat* m, const double* vP, const double* scale, const int* done_flag);
int bb_op_dot_flag(bb_mat* m, int mode, const int* done_flag);
int bb_op_tdot_flag(bb_mat* m, const double* w, bool have_w_partials, const int* done_flag,
                    bool fuse_reduce_into_consumer = false);
int bb_op_tdot_local(bb_mat* m, const double* w, const int* done_flag);   // product + collect into traw, no exchange
int bb_op_collect_local(bb_mat* m, const int* done_flag);                // the collect alone
int bb_mat_alloc_work(bb_mat* m);

int bb_slab_free(SlabFmt* f);
int bb_launch_spmv(bb_mat* m, SlabFmt* f, const double* gvec, const int* done_flag, bool skip_overflow_add = false);
// sliced format (bb_sell.cu)
i64 bb_sell_max_width(bb_ctx* ctx);
int bb_sell_build(bb_ctx* ctx, SlabFmt* f);          // from f->ptr / f->idx / f->val / f->slab_nnz1 (device arrays of the slab format)
int bb_sell_launch(bb_ctx* ctx, SlabFmt* f, const double* gvec, const int* done_flag, bool skip_overflow_add = false);
// fused P-side of a CG iteration (bb_pside.cu)
bool bb_pside_available(bb_mat* m);
int bb_pside_prepare(bb_mat* m);     // outside graph capture: uploads the exchange view
int bb_pside_enqueue(bb_mat* m);
bool bb_pside_folds_overflow(bb_mat* m);   // true: the fused kernel adds the overflow fragments (few long columns)
bool bb_pside_precollect(bb_mat* m);   // true: the caller launches the overflow fold + k_tdot_collect before the fused kernel
int bb_dense_fused(bb_mat* m, const int* done_flag);   // one pass over X: omega.(X sv + shift) and its X' product
void bb_sell_free(SlabFmt* f);

// ------------------------------------------------------------------------------------------
// device helpers
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum; result valid in every thread. blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double* sm32 /* >= 33 doubles */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm32[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? sm32[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sm32[32] = t;
    }
    __syncthreads();
    return sm32[32];
}

// fixed-order sum of `count` partials by one warp (all 32 lanes must call); result in all lanes
__device__ __forceinline__ double warp_sum_partials(const double* buf, int count) {
    int lane = threadIdx.x & 31;
    double t = 0.0;
    for (int i = lane; i < count; i += 32) t += buf[i];
    return warp_sum(t);
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based ----
struct Philox {
    uint32_t c[4];
    uint32_t k[2];
};

__host__ __device__ __forceinline__ void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint64_t p0 = (uint64_t)M0 * c[0];
    uint64_t p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ __forceinline__ void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// A per-element random stream: element `index` of call `offset` of stream `stream_id` under `seed`.
// Each refill consumes one counter value and yields two 53-bit uniforms in (0,1).
struct RandStream {
    uint32_t ctr[4];
    uint32_t key[2];
    double   spare;
    int      has_spare;
    __host__ __device__ void init(uint64_t seed, uint64_t offset, uint64_t index, uint32_t stream_id) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = (uint32_t)index;
        ctr[1] = (uint32_t)offset;
        ctr[2] = 0u;                                  // draw counter
        ctr[3] = (stream_id << 24) | ((uint32_t)((index >> 32) & 0xFFu) << 16) | (uint32_t)((offset >> 32) & 0xFFFFu);
        has_spare = 0; spare = 0.0;
    }
    __host__ __device__ double uniform() {
        if (has_spare) { has_spare = 0; return spare; }
        uint32_t o[4];
        philox4x32_10(ctr, key, o);
        ctr[2] += 1u;
        uint64_t a = (((uint64_t)o[0] << 32) | o[1]) >> 11;
        uint64_t b = (((uint64_t)o[2] << 32) | o[3]) >> 11;
        spare = ((double)b + 0.5) * (1.0 / 9007199254740992.0);
        has_spare = 1;
        return ((double)a + 0.5) * (1.0 / 9007199254740992.0);
    }
    // Box-Muller; consumes exactly two uniforms, returns one normal (the sine branch is dropped so
    // that the stream position after a normal does not depend on caching).
    __host__ __device__ double normal() {
        double u1 = uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
    }
};

enum { STREAM_EPS1 = 0, STREAM_EPS2 = 1, STREAM_PG = 2, STREAM_TS = 3 };

// ---- peer-memory exchange (bb_p2p.cu): device-side protocol pieces, usable from any kernel ----
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// slot this rank writes its next publication into (call before any thread of the grid can have advanced seq)
__device__ __forceinline__ double* p2p_publish_slot(const P2PView& v) {
    return v.peer_base[v.rank] + 32 + (v.st->seq & 1ull) * v.cap;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// to be called by every block after its last store into the slot: the last block to arrive signals all peers.
// Ordering: writers fence + relaxed counter RMW; the last block's thread 0 fences after its RMW (fence-fence
// synchronisation => it observes every block's stores) and then releases the flags at system scope.
__device__ __forceinline__ void p2p_publish_done(const P2PView& v) {
    __shared__ int p2p_is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(&v.st->blocks_done, 1u);
        p2p_is_last = (prev == gridDim.x * gridDim.y - 1) ? 1 : 0;
        if (p2p_is_last) __threadfence();
    }
    __syncthreads();
    if (!p2p_is_last) return;
    const unsigned long long seq = v.st->seq;
    if (v.variant & 8) {                       // nranks threads, one remote store each
        if ((int)threadIdx.x < v.nranks) {
            __threadfence_system();
            unsigned long long* f = reinterpret_cast<unsigned long long*>(v.peer_base[threadIdx.x]) + v.rank;
            if (v.variant & 1) st_relaxed_sys_u64(f, seq + 1ull); else st_release_sys_u64(f, seq + 1ull);
        }
        __syncthreads();
        if (threadIdx.x == 0) { v.st->blocks_done = 0u; v.st->seq = seq + 1ull; }
    } else if (threadIdx.x == 0) {
        v.st->blocks_done = 0u;
        __threadfence_system();
        for (int q = 0; q < v.nranks; ++q) {
            unsigned long long* f = reinterpret_cast<unsigned long long*>(v.peer_base[q]) + v.rank;
            if (v.variant & 1) st_relaxed_sys_u64(f, seq + 1ull); else st_release_sys_u64(f, seq + 1ull);
        }
        v.st->seq = seq + 1ull;
    }
}
// block-wide wait until every rank has published number st->seq; returns false after a time-out
__device__ __forceinline__ bool p2p_wait_all(const P2PView& v, int* sm_flag) {
    const unsigned long long want = v.st->seq;
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(v.peer_base[v.rank]);
    if (v.variant & 2) {                       // thread q polls rank q's flag
        if (threadIdx.x == 0) *sm_flag = 1;
        __syncthreads();
        if ((int)threadIdx.x < v.nranks) {
            unsigned long long spins = 0;
            while (ld_acquire_sys_u64(flags + threadIdx.x) < want) {
                if (++spins > (1ull << 26)) { *sm_flag = 0; v.st->error = 1u; break; }
                __nanosleep(20);
            }
        }
    } else if (threadIdx.x == 0) {
        int good = 1;
        for (int q = 0; q < v.nranks && good; ++q) {
            unsigned long long spins = 0;
            while (ld_acquire_sys_u64(flags + q) < want) {
                if (++spins > (1ull << 26)) { good = 0; v.st->error = 1u; break; }
                __nanosleep(20);
            }
        }
        *sm_flag = good;
    }
    __syncthreads();
    return *sm_flag != 0;
}
// element i of the published vectors summed over ranks in rank order (identical bits on every rank)
__device__ __forceinline__ double p2p_sum(const P2PView& v, i64 i) {
    const i64 off = 32 + ((v.st->seq - 1ull) & 1ull) * v.cap + i;
    double acc = 0.0;
    if (v.variant & 4) {                       // issue the peer loads together, then add in order
        for (int q0 = 0; q0 < v.nranks; q0 += 8) {
            double x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = (q0 + k < v.nranks) ? __ldcv(v.peer_base[q0 + k] + off) : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (q0 + k < v.nranks) acc += x[k];
        }
    } else {
        for (int q = 0; q < v.nranks; ++q) acc += __ldcv(v.peer_base[q] + off);
    }
    return acc;
}
#endif  // __CUDACC__
