// Sliced SpMV ("spmv_variant 1"): the kernel behind SparseDesignMatrix.dot / Tdot
// (reference: bayesbridge/design_matrix/sparse_matrix.py:90-101 main_dot, :121-129 main_Tdot; the CPU
// path there is scipy's csr_matvec / csc_matvec).
//
// The slab format (bb_sparse.cu) already groups the nnz so that one CTA gathers from a <= W-wide window of the
// input vector staged in shared memory.  Inside a slab a row (dot) or column (Tdot) keeps only ~20-30 nnz, so
// the cost of a product is dominated by how the short per-segment sums are formed.  Here every virtual segment
// (slab, segment) is cut into FRAGMENTS of at most SELL_LMAX nnz and ONE LANE sums one fragment serially:
//
//   * fragments are sorted by length (inside a slab, inside windows of 2^17 segments) and grouped 32 at a time
//     into SLICES; a slice is padded to a common length L (multiple of 4).  After the sort neighbours have
//     (nearly) equal length, so the padding is ~6 %;
//   * gather indices are 16-bit offsets into the slab window (W <= 29 024 < 65 536): 2 B per nnz instead of 4;
//     a padded entry points at a zero kept behind the window;
//   * a slice is stored as L/4 ROWS of 256 B: row k holds, for each of the 32 lanes, its entries 4k..4k+3
//     (two 32-bit words = four 16-bit indices) -> one coalesced 8-byte load per lane and row;
//   * the slices of one (CTA, slab) section are split into 32 contiguous STRIPS of equal cost, one per warp, so a
//     warp streams a contiguous byte range through a register ring (R rows in flight) with no dependent
//     address loads; per entry the lane does: extract index, LDS.64 gather, DADD;
//   * every slice starts with a HEADER row in the same stream: lane l's 8 bytes are {output slot of its fragment,
//     number of data rows of the slice}.  The header therefore arrives through the same register ring as the
//     indices -- no separate, dependent load per slice (with the slot in an array of its own, a warp walking short
//     slices waited a full memory latency per slice: 25 % of all stall samples in the first ncu capture);
//   * at the end of a slice every lane stores its sum to part[slot]; the first fragment of a virtual segment
//     owns slot v = slab*n_seg + seg (the layout k_dot_finish / k_tdot_collect read), further fragments of a
//     long segment own overflow slots that k_ovf_add folds in afterwards (rare);
//   * the window is staged with cp.async.bulk (TMA 1-D bulk copy) + mbarrier when alignment allows.
//
// There is no cross-lane communication, no shared-memory scratch and no tile straddling (no fix-up launch).
// The summation order of a segment is static (fixed by the format), so results are bit-reproducible; it differs
// from scipy's left-to-right order, i.e. results agree with the reference to rounding, not bit for bit.
// Inside a fragment the entries are reordered at build time so that the 16 lanes of a half-warp hit different
// 8-byte banks of the window where possible (option bank_permute).
#include "bb_internal.cuh"
#include <cub/cub.cuh>
#include <vector>
#include <algorithm>
#include <queue>
#include <functional>

constexpr int SELL_LMAX = 256;          // max nnz per fragment (multiple of 4)
constexpr int SELL_WINDOW_BITS = 17;    // fragments are length-sorted inside windows of 2^17 segments (write locality)
constexpr int SELL_THREADS = 1024;
constexpr int SELL_WARPS = SELL_THREADS / 32;
constexpr int SELL_RING_BIN = 8;        // rows in flight per warp, pattern-only
constexpr int SELL_RING_VAL = 2;        // rows in flight per warp, valued
constexpr int SELL_SLICE_COST = 3;      // per-slice overhead in row equivalents (partitioning)
constexpr int SELL_RESTAGE_COST = 400;  // a second window staged by the same CTA, in row equivalents of the CTA's cost: measured
                                        // 2.5-3.8 us against ~0.31 us per 32 rows (one row per warp), profiles/r02_spmv_timeline.md

i64 bb_sell_max_width(bb_ctx* ctx) {
    i64 w = ((i64)ctx->smem_optin - 64) / 8 - 2;
    w &= ~(i64)31;
    if (w > 65504) w = 65504;           // 16-bit in-slab indices; index W is the zero slot
    if (w < 32) w = 32;
    return w;
}

// ------------------------------------------------------------------------------------------
// build kernels
// number of fragments the matrix would have for every candidate fragment length 256 >> c, c = 0..5
__global__ void k_sell_survey(const int* __restrict__ ptr, const int* __restrict__ slab_nnz1, i64 V, i64 n_seg,
                              unsigned long long* __restrict__ count /* [6] */) {
    unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};
    for (i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (i64)gridDim.x * blockDim.x) {
        int slab = (int)(v / n_seg);
        int len = min(ptr[v + 1], slab_nnz1[slab]) - ptr[v];
        if (len <= 0) continue;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] += (unsigned long long)((len + (256 >> c) - 1) / (256 >> c));
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        unsigned long long t = acc[c];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0 && t) atomicAdd(&count[c], t);
    }
}

__global__ void k_sell_count(const int* __restrict__ ptr, const int* __restrict__ slab_nnz1, i64 V, i64 n_seg, int lmax,
                             int* __restrict__ nfr, int* __restrict__ novf) {
    i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    int slab = (int)(v / n_seg);
    int len = min(ptr[v + 1], slab_nnz1[slab]) - ptr[v];
    if (len < 0) len = 0;
    int nf = (len + lmax - 1) / lmax;
    nfr[v] = nf;
    novf[v] = nf > 1 ? nf - 1 : 0;
}

__global__ void k_sell_emit(const int* __restrict__ ptr, const int* __restrict__ slab_nnz1, i64 V, i64 n_seg,
                            const int* __restrict__ nfr, const int* __restrict__ frag_base, const int* __restrict__ ovf_base,
                            int slab_shift, int lmax, int* __restrict__ frag_src, int* __restrict__ frag_len,
                            unsigned* __restrict__ frag_slot, i64* __restrict__ key) {
    i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int nf = nfr[v];
    if (nf == 0) return;
    const i64 slab = v / n_seg, seg = v - slab * n_seg, window = seg >> SELL_WINDOW_BITS;
    const int start = ptr[v];
    const int len = min(ptr[v + 1], slab_nnz1[slab]) - start;
    const int fb = frag_base[v], ob = ovf_base[v];
    for (int f = 0; f < nf; ++f) {
        const int id = fb + f;
        const int l = min(lmax, len - f * lmax);
        frag_src[id] = start + f * lmax;
        frag_len[id] = l;
        frag_slot[id] = (f == 0) ? (unsigned)v : (unsigned)(V + ob + (f - 1));
        key[id] = (slab << slab_shift) | (window << 9) | (i64)(SELL_LMAX - l);
    }
}

__global__ void k_sell_iota(int* a, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int)i;
}

// out[s] = first position q with key[q] >= (s << shift), s in [0, nslab]
__global__ void k_sell_slab_lb(const i64* __restrict__ key, i64 n, int nslab, int shift, int* __restrict__ out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nslab) return;
    const i64 target = (i64)s << shift;
    i64 lo = 0, hi = n;
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        if (key[mid] < target) lo = mid + 1; else hi = mid;
    }
    out[s] = (int)lo;
}

// virtual segments with more than one fragment, in order
__global__ void k_sell_ovf_list(const int* __restrict__ nfr, const int* __restrict__ ovf_base, const int* __restrict__ ovf_rank,
                                i64 V, int* __restrict__ ovf_piece, int* __restrict__ ovf_first) {
    i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    if (nfr[v] > 1) { ovf_piece[ovf_rank[v]] = (int)v; ovf_first[ovf_rank[v]] = ovf_base[v]; }
}
__global__ void k_sell_flag_ovf(const int* __restrict__ nfr, i64 V, int* __restrict__ flag) {
    i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) flag[v] = nfr[v] > 1 ? 1 : 0;
}

__device__ __forceinline__ int sell_slab_of(const int* __restrict__ slab_slice0, int nslab, int slice) {
    // last slab s with slab_slice0[s] <= slice (slabs without slices have equal consecutive entries)
    int lo = 0, hi = nslab;          // answer in [0, nslab-1]
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (slab_slice0[mid] <= slice) lo = mid; else hi = mid; }
    return lo;
}

// rows (of 4 entries per lane) of every slice: one warp per slice
__global__ void k_sell_slice_rows(const int* __restrict__ slab_slice0, const int* __restrict__ slab_frag0, int nslab, int nslices,
                                  const int* __restrict__ sorted_id, const int* __restrict__ frag_len, int* __restrict__ nrows) {
    const int slice = (int)(((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (slice >= nslices) return;
    const int slab = sell_slab_of(slab_slice0, nslab, slice);
    const int q = slab_frag0[slab] + (slice - slab_slice0[slab]) * 32 + lane;
    int len = (q < slab_frag0[slab + 1]) ? frag_len[sorted_id[q]] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) nrows[slice] = 1 + ((len + 3) >> 2);       // header row + data rows
}

// Fill one slice per warp.  Lane l owns fragment l of the slice.  With `permute` the entries of every fragment are
// reordered so that, entry position by entry position, the 16 lanes of a half-warp read different 8-byte banks of the
// window where possible: a lane keeps its entries bucketed by bank (index mod 16) and, in lane order, takes one from
// a bank no earlier lane of its half-warp has taken at this position (rotating preference), else from any bank.
struct SellFillSmem {
    unsigned short el[32][SELL_LMAX];
    unsigned char ord[32][SELL_LMAX];
    unsigned short bptr[32][16];
    unsigned short bend[32][16];
};
constexpr int SELL_FILL_WARPS = 4;

__global__ void __launch_bounds__(SELL_FILL_WARPS * 32)
k_sell_fill(const int* __restrict__ slab_slice0, const int* __restrict__ slab_frag0, int nslab, int nslices,
            const int* __restrict__ sorted_id, const int* __restrict__ frag_src, const int* __restrict__ frag_len,
            const unsigned* __restrict__ frag_slot, const int* __restrict__ idx, const double* __restrict__ val, int W,
            const unsigned* __restrict__ sl_off, const int* __restrict__ sl_nrows, unsigned trash_slot, int permute,
            unsigned* __restrict__ words, double* __restrict__ vals) {
    extern __shared__ __align__(16) unsigned char fill_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slice = blockIdx.x * SELL_FILL_WARPS + warp;
    if (slice >= nslices) return;
    SellFillSmem& S = reinterpret_cast<SellFillSmem*>(fill_smem)[warp];
    const int slab = sell_slab_of(slab_slice0, nslab, slice);
    const int q = slab_frag0[slab] + (slice - slab_slice0[slab]) * 32 + lane;
    const bool valid = q < slab_frag0[slab + 1];
    int src = 0, len = 0;
    unsigned slot = trash_slot;
    if (valid) { const int id = sorted_id[q]; src = frag_src[id]; len = frag_len[id]; slot = frag_slot[id]; }
    const int gbase = slab * W;
    for (int j = 0; j < len; ++j) S.el[lane][j] = (unsigned short)(idx[src + j] - gbase);
    unsigned mask = 0;
    if (permute) {
        for (int b = 0; b < 16; ++b) S.bend[lane][b] = 0;
        for (int j = 0; j < len; ++j) S.bend[lane][S.el[lane][j] & 15] += 1;
        int run = 0;
        for (int b = 0; b < 16; ++b) { int c = S.bend[lane][b]; S.bptr[lane][b] = (unsigned short)run; run += c; }
        for (int j = 0; j < len; ++j) { int b = S.el[lane][j] & 15; S.ord[lane][S.bptr[lane][b]++] = (unsigned char)j; }
        for (int b = 0; b < 16; ++b) S.bend[lane][b] = S.bptr[lane][b];                      // ends
        for (int b = 15; b > 0; --b) S.bptr[lane][b] = S.bend[lane][b - 1];                  // starts
        S.bptr[lane][0] = 0;
        for (int b = 0; b < 16; ++b) if (S.bptr[lane][b] < S.bend[lane][b]) mask |= 1u << b;
    }
    __syncwarp();
    const unsigned r0 = sl_off[slice] + 1u;                 // first data row (the header row precedes it); sl_off = the slice's
    const int nrows = sl_nrows[slice] - 1;                  // place in the row stream, which the work partition decides
    {   // header row: {output slot, data rows of the slice}
        const size_t o32 = ((size_t)(r0 - 1u) * 32 + lane) * 2;
        words[o32] = slot;
        words[o32 + 1] = (unsigned)nrows;
    }
    // permute == 2: per-bank load of this lane's half-warp; lane hl is the book-keeper of bank hl
    const int hl = lane & 15, hbase = lane & 16;
    int load = 0;
    if (permute == 2)
        for (int l = 0; l < 16; ++l) load += (int)S.bend[hbase + l][hl] - (int)S.bptr[hbase + l][hl];
    for (int j = 0; j < 4 * nrows; ++j) {
        int pos = -1;
        unsigned used = 0;
        if (permute == 2) {
            // "Most loaded bank first": a conflict-free position is a matching between lanes and banks, and the number
            // of positions a half-warp needs is bounded below by its fullest bank, so each position serves the banks
            // in order of their remaining load; a bank goes to the lane with the fewest other banks still open.
            bool assigned = (mask == 0u);
            unsigned avail = 0;                 // as book-keeper of bank hl: lanes that could still take an entry from it
            for (int l = 0; l < 16; ++l) {
                const unsigned ml = __shfl_sync(0xffffffffu, mask, l, 16);
                avail |= ((ml >> hl) & 1u) << l;
            }
            for (int pick = 0; pick < 16; ++pick) {
                int key = (!((used >> hl) & 1u) && avail != 0u) ? ((load << 4) | hl) : -1;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o, 16));
                if (!__any_sync(0xffffffffu, key >= 0)) break;
                const int bstar = key & 15;
                int k2 = (key >= 0 && !assigned && ((mask >> bstar) & 1u)) ? ((__popc(mask & ~used) << 5) | hl) : 0x7fffffff;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) k2 = min(k2, __shfl_xor_sync(0xffffffffu, k2, o, 16));
                if (key >= 0) {
                    const int kstar = k2 & 15;
                    if (hl == kstar) {
                        const int pp = S.bptr[lane][bstar];
                        pos = S.ord[lane][pp];
                        S.bptr[lane][bstar] = (unsigned short)(pp + 1);
                        if (pp + 1 == S.bend[lane][bstar]) mask &= ~(1u << bstar);
                        assigned = true;
                    }
                    used |= 1u << bstar;
                    if (hl == bstar) load -= 1;
                    avail &= ~(1u << kstar);
                }
            }
            // lanes left over share a bank with another lane at this position (unavoidable here)
            int forced = -1;
            if (!assigned) {
                const int r = (j + hl) & 15;
                const unsigned mm = ((mask >> r) | (mask << (16 - r))) & 0xffffu;
                forced = (__ffs(mm) - 1 + r) & 15;
                const int pp = S.bptr[lane][forced];
                pos = S.ord[lane][pp];
                S.bptr[lane][forced] = (unsigned short)(pp + 1);
                if (pp + 1 == S.bend[lane][forced]) mask &= ~(1u << forced);
            }
            for (int l = 0; l < 16; ++l) {
                const int fb = __shfl_sync(0xffffffffu, forced, l, 16);
                if (fb == hl) load -= 1;
            }
        } else if (permute) {
            for (int k = 0; k < 16; ++k) {
                int chosen = -1;
                if ((lane & 15) == k && mask != 0u) {
                    const unsigned av = mask & ~used;
                    const unsigned m = av ? av : mask;
                    const int r = (j + k) & 15;
                    const unsigned mm = ((m >> r) | (m << (16 - r))) & 0xffffu;
                    const int b = (__ffs(mm) - 1 + r) & 15;
                    const int pp = S.bptr[lane][b];
                    pos = S.ord[lane][pp];
                    S.bptr[lane][b] = (unsigned short)(pp + 1);
                    if (pp + 1 == S.bend[lane][b]) mask &= ~(1u << b);
                    chosen = b;
                }
                const int cb = __shfl_sync(0xffffffffu, chosen, k, 16);
                if (cb >= 0) used |= 1u << cb;
            }
        } else if (j < len) {
            pos = j;
        }
        // padded entry -> one of the two zeros kept behind the window (banks 0 and 1): the one whose bank is free
        const unsigned e = (pos >= 0) ? (unsigned)S.el[lane][pos] : (unsigned)W + (((used & 3u) == 1u) ? 1u : 0u);
        // row j>>2, lane, entry j&3 : 16-bit entries, 4 per lane and row
        const size_t o16 = ((size_t)(r0 + (unsigned)(j >> 2)) * 32 + lane) * 4 + (j & 3);
        reinterpret_cast<unsigned short*>(words)[o16] = (unsigned short)e;
        if (vals) vals[o16] = (pos >= 0) ? val[src + pos] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// the product kernel
__device__ __forceinline__ double sell_lds(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 sell_ldg_row(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 sell_ldg_val(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <bool BINARY>
struct SellRow {
    uint2 q;
};
template <>
struct SellRow<false> {
    uint2 q;
    double2 va, vb;
};

// row `k` (relative to the warp's running pointers) of the strip
template <bool BINARY>
__device__ __forceinline__ void sell_load_row(SellRow<BINARY>& r, const uint2* __restrict__ rp, const double2* __restrict__ vp, int k) {
    r.q = sell_ldg_row(rp + k * 32);
    if constexpr (!BINARY) {
        r.va = sell_ldg_val(vp + k * 64);
        r.vb = sell_ldg_val(vp + k * 64 + 1);
    }
}

// per-warp state of the strip walk
struct SellWalk {
    double a0, a1;            // the two accumulators of the current fragment
    int rows_left;            // data rows until the end of the current slice (0: the next row is a header)
    unsigned slot;            // output slot of this lane's fragment
};

// One batch of R rows: first issue the loads of the NEXT batch into `nxt` (they are only consumed one batch later,
// so the warp never waits on a load it has just issued), then consume `cur` row by row.
template <bool BINARY, int R>
__device__ __forceinline__ void sell_batch(SellRow<BINARY> (&cur)[R], SellRow<BINARY> (&nxt)[R],
                                           const uint2* __restrict__ rp, const double2* __restrict__ vp,
                                           unsigned r, unsigned r_end, unsigned sbase,
                                           double* __restrict__ part, SellWalk& w) {
#pragma unroll
    for (int k = 0; k < R; ++k)
        if (r + R + k < r_end) sell_load_row<BINARY>(nxt[k], rp, vp, R + k);
#pragma unroll
    for (int k = 0; k < R; ++k) {
        if (r + k < r_end) {
            if (w.rows_left == 0) {             // header row (warp-uniform: rows_left is the same in every lane)
                w.slot = cur[k].q.x;
                w.rows_left = __shfl_sync(0xffffffffu, (int)cur[k].q.y, 0);
            } else {
                const double g0 = sell_lds(sbase + ((cur[k].q.x & 0xffffu) << 3));
                const double g1 = sell_lds(sbase + ((cur[k].q.x >> 16) << 3));
                const double g2 = sell_lds(sbase + ((cur[k].q.y & 0xffffu) << 3));
                const double g3 = sell_lds(sbase + ((cur[k].q.y >> 16) << 3));
                if constexpr (BINARY) {
                    w.a0 += g0; w.a1 += g1; w.a0 += g2; w.a1 += g3;
                } else {
                    w.a0 += cur[k].va.x * g0; w.a1 += cur[k].va.y * g1; w.a0 += cur[k].vb.x * g2; w.a1 += cur[k].vb.y * g3;
                }
                if (--w.rows_left == 0) {
                    part[w.slot] = w.a0 + w.a1;
                    w.a0 = 0.0; w.a1 = 0.0;
                }
            }
        }
    }
}

__device__ __forceinline__ unsigned long long sell_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr int SELL_DBG_WORDS = 8 + SELL_WARPS;      // per CTA: entry, after wait, window staged, end, #sections, spare; per-warp end

// DBG = true is a separate instantiation used only by bb_spmv_timeline (per-CTA / per-warp %globaltimer stamps into `dbg`);
// the production instantiation carries none of that code.
template <bool BINARY, int R, bool DBG = false>
__global__ void __launch_bounds__(SELL_THREADS, 1)
k_sell_spmv(const uint2* __restrict__ rows, const double* __restrict__ vals,
            const int* __restrict__ cta_sec0, const int* __restrict__ sec_slab, const int* __restrict__ sec_wstart,
            const double* __restrict__ gvec, int W, i64 n_gather, int use_bulk,
            double* __restrict__ part, const int* __restrict__ done_flag, unsigned long long* __restrict__ dbg = nullptr)
{
    if constexpr (DBG) { if (threadIdx.x == 0) { dbg[blockIdx.x * SELL_DBG_WORDS + 0] = sell_now(); dbg[blockIdx.x * SELL_DBG_WORDS + 4] = 0ull; } }
    // PDL (bb_internal.cuh): everything up to pdl_wait() reads only the matrix format and writes only shared memory
    pdl_trigger();
    extern __shared__ __align__(128) double sell_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sec_lo = cta_sec0[blockIdx.x], sec_hi = cta_sec0[blockIdx.x + 1];
    if (sec_lo >= sec_hi) { pdl_wait(); return; }
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sell_smem);
    const unsigned mbar = sbase + (unsigned)(W + 2) * 8u;
    if (tid == 0) {
        sell_smem[W] = 0.0;                       // target of padded entries
        sell_smem[W + 1] = 0.0;
        if (use_bulk) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    unsigned phase = 0;
    for (int sec = sec_lo; sec < sec_hi; ++sec) {
        const int slab = sec_slab[sec];
        const i64 gbase = (i64)slab * W;
        const i64 rem = n_gather - gbase;
        const int wlen = rem < (i64)W ? (int)rem : W;
        // this warp's strip (a contiguous row range that starts at a slice header) -- fetched before the barrier so
        // that the loads overlap the staging
        unsigned r = (unsigned)sec_wstart[sec * (SELL_WARPS + 1) + warp];
        const unsigned r_end = (unsigned)sec_wstart[sec * (SELL_WARPS + 1) + warp + 1];
        // strip prologue: issue the first loads of the row stream (matrix format: independent of the previous kernel)
        // before the window is requested, so that they overlap the tail of the previous kernel and the staging
        SellWalk w;
        w.a0 = 0.0; w.a1 = 0.0; w.rows_left = 0; w.slot = 0u;
        SellRow<BINARY> bufA[R], bufB[R];
        const uint2* rp = rows + (size_t)r * 32 + lane;
        const double2* vp = reinterpret_cast<const double2*>(vals);
        if constexpr (!BINARY) vp += ((size_t)r * 32 + lane) * 2;
#pragma unroll
        for (int k = 0; k < R; ++k)
            if (r + k < r_end) sell_load_row<BINARY>(bufA[k], rp, vp, k);
        if (sec == sec_lo) {
            pdl_wait();                           // the gathered vector and the done flag come from the previous kernel
            if (done_flag != nullptr && *done_flag) return;
            if constexpr (DBG) { if (tid == 0) dbg[blockIdx.x * SELL_DBG_WORDS + 1] = sell_now(); }
        }
        __syncthreads();                          // readers of the previous window are done; mbarrier init is visible
        if (use_bulk) {
            if (tid == 0) {
                const unsigned bytes = (unsigned)((wlen + 1) & ~1) * 8u;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
                const char* src = reinterpret_cast<const char*>(gvec + gbase);
                for (unsigned off = 0; off < bytes; off += 32768u) {
                    const unsigned sz = min(32768u, bytes - off);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(sbase + off), "l"(src + off), "r"(sz), "r"(mbar) : "memory");
                }
            }
        } else {
            const double* src = gvec + gbase;
            int i = tid;
            for (; i + 3 * SELL_THREADS < wlen; i += 4 * SELL_THREADS) {
                double a0 = src[i], a1 = src[i + SELL_THREADS], a2 = src[i + 2 * SELL_THREADS], a3 = src[i + 3 * SELL_THREADS];
                sell_smem[i] = a0; sell_smem[i + SELL_THREADS] = a1; sell_smem[i + 2 * SELL_THREADS] = a2; sell_smem[i + 3 * SELL_THREADS] = a3;
            }
            for (; i < wlen; i += SELL_THREADS) sell_smem[i] = src[i];
        }
        if (use_bulk) {
            unsigned ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\t"
                             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(mbar), "r"(phase) : "memory");
            }
            phase ^= 1u;
        } else {
            __syncthreads();
        }
        if constexpr (DBG) { if (tid == 0) { if (sec == sec_lo) dbg[blockIdx.x * SELL_DBG_WORDS + 2] = sell_now(); dbg[blockIdx.x * SELL_DBG_WORDS + 4] += 1ull; } }
        while (r < r_end) {
            sell_batch<BINARY, R>(bufA, bufB, rp, vp, r, r_end, sbase, part, w);
            r += R; rp += R * 32; if constexpr (!BINARY) vp += R * 64;
            if (r >= r_end) break;
            sell_batch<BINARY, R>(bufB, bufA, rp, vp, r, r_end, sbase, part, w);
            r += R; rp += R * 32; if constexpr (!BINARY) vp += R * 64;
        }
        if constexpr (DBG) { if (lane == 0) dbg[blockIdx.x * SELL_DBG_WORDS + 8 + warp] = sell_now(); }
    }
    if constexpr (DBG) {
        __syncthreads();
        if (tid == 0) dbg[blockIdx.x * SELL_DBG_WORDS + 3] = sell_now();
    }
}

// part[v] += overflow fragments of v, in fragment order (one warp per long virtual segment)
__global__ void k_sell_ovf_add(const int* __restrict__ ovf_piece, const int* __restrict__ ovf_first, int n_pieces,
                               i64 V, double* __restrict__ part, const int* __restrict__ done_flag) {
    pdl_trigger();
    pdl_wait();
    if (done_flag != nullptr && *done_flag) return;
    const int i = (int)(((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= n_pieces) return;
    const int a = ovf_first[i], b = ovf_first[i + 1];
    const double t = warp_sum_partials(part + V + a, b - a);
    if ((threadIdx.x & 31) == 0) part[ovf_piece[i]] += t;
}

// ------------------------------------------------------------------------------------------
// host side
void bb_sell_free(SlabFmt* f) {
    void* ptrs[] = {f->sl_off, f->sl_slot, f->sl_pairs, f->sl_vals, f->sl_slab_slice0, f->sl_cta_slice0, f->ovf_piece, f->ovf_first};
    for (void* p : ptrs) if (p) cudaFree(p);
    f->sl_off = nullptr; f->sl_slot = nullptr; f->sl_pairs = nullptr; f->sl_vals = nullptr;
    f->sl_slab_slice0 = nullptr; f->sl_cta_slice0 = nullptr; f->ovf_piece = nullptr; f->ovf_first = nullptr;
}

namespace {
struct DevBuf {           // frees its allocations on scope exit (build temporaries)
    std::vector<void*> v;
    ~DevBuf() { for (void* p : v) if (p) cudaFree(p); }
    template <typename T> cudaError_t alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) v.push_back(p);
        *out = (T*)p;
        return e;
    }
};
}  // namespace

static int sell_exclusive_scan(bb_ctx* ctx, const int* in, int* out, i64 n) {
    size_t tmp_bytes = 0;
    BB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, ctx->stream));
    void* tmp = nullptr;
    BB_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, ctx->stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        bb_set_error("sliced format: scan failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return BB_ERR_CUDA;
    }
    ctx->launches += 2;
    return BB_OK;
}

static int sell_last_plus(bb_ctx* ctx, const int* scan, const int* count, i64 n, i64* total) {
    int a = 0, b = 0;
    if (n > 0) {
        BB_CUDA(cudaMemcpyAsync(&a, scan + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        BB_CUDA(cudaMemcpyAsync(&b, count + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        BB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *total = (i64)a + b;
    return BB_OK;
}

static int bits_for64(i64 maxval) {
    int b = 1;
    while (b < 62 && ((i64)1 << b) <= maxval) ++b;
    return b;
}

// Builds the sliced arrays from the slab format's device arrays (f->ptr, f->idx, f->val, f->slab_nnz1).
int bb_sell_build(bb_ctx* ctx, SlabFmt* f) {
    cudaStream_t st = ctx->stream;
    const i64 n_seg = f->n_seg, V = (i64)f->nslab * n_seg;
    const int nslab = f->nslab, W = f->W;
    const int TB = 256;
    DevBuf tmp;
    f->variant = 1;
    f->sl_partition = 1;
    f->nslices = 0; f->n_ovf = 0; f->n_ovf_pieces = 0;
    if (W > 65504) { bb_set_error("sliced format: slab width %d exceeds the 16-bit index range", W); return BB_ERR_ARG; }
    int *nfr = nullptr, *novf = nullptr, *frag_base = nullptr, *ovf_base = nullptr;
    BB_CUDA(tmp.alloc(&nfr, (size_t)V + 1)); BB_CUDA(tmp.alloc(&novf, (size_t)V + 1));
    BB_CUDA(tmp.alloc(&frag_base, (size_t)V + 1)); BB_CUDA(tmp.alloc(&ovf_base, (size_t)V + 1));
    i64 F = 0, n_ovf = 0, n_ovf_pieces = 0;
    // Fragment length: one lane sums one fragment, so a matrix with few long segments (a small problem with dense
    // columns) would leave most warps idle.  Take the longest cap 256, 128, ... 8 that still yields two slices of 32
    // fragments for every warp of the grid (when even 8 does not, the problem is latency-bound anyway).
    int lmax = SELL_LMAX;
    if (V > 0) {
        unsigned long long* d_cnt = nullptr;
        BB_CUDA(tmp.alloc(&d_cnt, 6));
        BB_CUDA(cudaMemsetAsync(d_cnt, 0, 6 * sizeof(unsigned long long), st));
        k_sell_survey<<<ctx->sm_count * 4, 256, 0, st>>>(f->ptr, f->slab_nnz1, V, n_seg, d_cnt);
        ctx->launches += 1;
        unsigned long long h_cnt[6];
        BB_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        const unsigned long long want = (unsigned long long)ctx->sm_count * SELL_WARPS * 32 * 2;
        int c = 0;
        while (c < 5 && h_cnt[c] < want) ++c;
        lmax = 256 >> c;
        if (ctx->opt_sell_lmax > 0) lmax = (int)std::min<i64>(256, std::max<i64>(4, ctx->opt_sell_lmax & ~(i64)3));
    }
    f->sl_lmax = lmax;
    if (V > 0) {
        const int g = (int)((V + TB - 1) / TB);
        k_sell_count<<<g, TB, 0, st>>>(f->ptr, f->slab_nnz1, V, n_seg, lmax, nfr, novf);
        ctx->launches += 1;
        BB_TRY(sell_exclusive_scan(ctx, nfr, frag_base, V));
        BB_TRY(sell_exclusive_scan(ctx, novf, ovf_base, V));
        BB_TRY(sell_last_plus(ctx, frag_base, nfr, V, &F));
        BB_TRY(sell_last_plus(ctx, ovf_base, novf, V, &n_ovf));
    }
    if (F >= ((i64)1 << 31) - 64 || V + n_ovf + 1 >= ((i64)1 << 32) - 1) { bb_set_error("sliced format: too many fragments"); return BB_ERR_ARG; }
    f->n_ovf = n_ovf;
    // output array: [V] first fragments (the layout the consumers read) + overflow slots + one trash slot
    if (f->part) { cudaFree(f->part); f->part = nullptr; }
    BB_CUDA(cudaMalloc((void**)&f->part, (size_t)(V + n_ovf + 1) * sizeof(double)));
    BB_CUDA(cudaMemsetAsync(f->part, 0, (size_t)(V + n_ovf + 1) * sizeof(double), st));   // empty segments stay 0 for ever
    const unsigned trash_slot = (unsigned)(V + n_ovf);
    // list of long segments
    if (n_ovf > 0) {
        int *flag = nullptr, *rank = nullptr;
        BB_CUDA(tmp.alloc(&flag, (size_t)V)); BB_CUDA(tmp.alloc(&rank, (size_t)V));
        const int g = (int)((V + TB - 1) / TB);
        k_sell_flag_ovf<<<g, TB, 0, st>>>(nfr, V, flag);
        ctx->launches += 1;
        BB_TRY(sell_exclusive_scan(ctx, flag, rank, V));
        BB_TRY(sell_last_plus(ctx, rank, flag, V, &n_ovf_pieces));
        BB_CUDA(cudaMalloc((void**)&f->ovf_piece, (size_t)n_ovf_pieces * sizeof(int)));
        BB_CUDA(cudaMalloc((void**)&f->ovf_first, ((size_t)n_ovf_pieces + 1) * sizeof(int)));
        k_sell_ovf_list<<<g, TB, 0, st>>>(nfr, ovf_base, rank, V, f->ovf_piece, f->ovf_first);
        ctx->launches += 1;
        const int last = (int)n_ovf;
        BB_CUDA(cudaMemcpyAsync(f->ovf_first + n_ovf_pieces, &last, sizeof(int), cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaStreamSynchronize(st));
    }
    f->n_ovf_pieces = (int)n_ovf_pieces;

    // fragments, sorted by (slab, window, length descending)
    std::vector<int> slab_frag0((size_t)nslab + 1, 0), slab_slice0((size_t)nslab + 1, 0);
    int *frag_src = nullptr, *frag_len = nullptr, *ids = nullptr, *sorted_id = nullptr, *d_slab_frag0 = nullptr;
    unsigned* frag_slot = nullptr;
    i64 *key = nullptr, *key_sorted = nullptr;
    BB_CUDA(tmp.alloc(&frag_src, (size_t)F)); BB_CUDA(tmp.alloc(&frag_len, (size_t)F)); BB_CUDA(tmp.alloc(&frag_slot, (size_t)F));
    BB_CUDA(tmp.alloc(&ids, (size_t)F)); BB_CUDA(tmp.alloc(&sorted_id, (size_t)F));
    BB_CUDA(tmp.alloc(&key, (size_t)F)); BB_CUDA(tmp.alloc(&key_sorted, (size_t)F));
    BB_CUDA(tmp.alloc(&d_slab_frag0, (size_t)nslab + 1));
    const int wbits = bits_for64(n_seg > 0 ? ((n_seg - 1) >> SELL_WINDOW_BITS) : 0);
    const int slab_shift = 9 + wbits;
    const int key_bits = slab_shift + bits_for64(nslab > 0 ? nslab - 1 : 0);
    if (F > 0) {
        const int g = (int)((V + TB - 1) / TB);
        k_sell_emit<<<g, TB, 0, st>>>(f->ptr, f->slab_nnz1, V, n_seg, nfr, frag_base, ovf_base, slab_shift, lmax,
                                      frag_src, frag_len, frag_slot, key);
        k_sell_iota<<<(int)((F + TB - 1) / TB), TB, 0, st>>>(ids, F);
        ctx->launches += 2;
        size_t tb = 0;
        BB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key_sorted, ids, sorted_id, (int)F, 0, key_bits, st));
        void* t = nullptr;
        BB_CUDA(cudaMalloc(&t, tb ? tb : 1));
        cudaError_t e = cub::DeviceRadixSort::SortPairs(t, tb, key, key_sorted, ids, sorted_id, (int)F, 0, key_bits, st);
        cudaError_t e2 = cudaStreamSynchronize(st);
        cudaFree(t);
        if (e != cudaSuccess || e2 != cudaSuccess) {
            bb_set_error("sliced format: sort failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
            return BB_ERR_CUDA;
        }
        ctx->launches += 4;
    }
    k_sell_slab_lb<<<(nslab + 1 + TB - 1) / TB, TB, 0, st>>>(key_sorted, F, nslab, slab_shift, d_slab_frag0);
    ctx->launches += 1;
    BB_CUDA(cudaMemcpyAsync(slab_frag0.data(), d_slab_frag0, ((size_t)nslab + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    for (int s = 0; s < nslab; ++s)
        slab_slice0[(size_t)s + 1] = slab_slice0[(size_t)s] + (slab_frag0[(size_t)s + 1] - slab_frag0[(size_t)s] + 31) / 32;
    const int nslices = slab_slice0[(size_t)nslab];
    f->nslices = nslices;
    BB_CUDA(cudaMalloc((void**)&f->sl_slab_slice0, ((size_t)nslab + 1) * sizeof(int)));
    BB_CUDA(cudaMemcpyAsync(f->sl_slab_slice0, slab_slice0.data(), ((size_t)nslab + 1) * sizeof(int), cudaMemcpyHostToDevice, st));

    // rows per slice (device) -> host; `off` = prefix sums in the sorted slice order (the cost scale of the partition below)
    int* nrows = nullptr;
    BB_CUDA(tmp.alloc(&nrows, (size_t)nslices + 1));
    BB_CUDA(cudaMemsetAsync(nrows, 0, ((size_t)nslices + 1) * sizeof(int), st));
    BB_CUDA(cudaMalloc((void**)&f->sl_off, ((size_t)nslices + 1) * sizeof(unsigned)));
    BB_CUDA(cudaMemsetAsync(f->sl_off, 0, ((size_t)nslices + 1) * sizeof(unsigned), st));
    std::vector<int> nr((size_t)nslices + 1, 0);
    std::vector<unsigned> off((size_t)nslices + 1, 0u);
    if (nslices > 0) {
        k_sell_slice_rows<<<(int)(((i64)nslices * 32 + TB - 1) / TB), TB, 0, st>>>(f->sl_slab_slice0, d_slab_frag0, nslab, nslices,
                                                                                  sorted_id, frag_len, nrows);
        ctx->launches += 1;
        BB_CUDA(cudaMemcpyAsync(nr.data(), nrows, (size_t)nslices * sizeof(int), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
    }
    {
        unsigned long long run = 0;
        for (int s2 = 0; s2 < nslices; ++s2) { off[(size_t)s2] = (unsigned)run; run += (unsigned long long)nr[(size_t)s2]; }
        if (run >= (1ull << 31)) { bb_set_error("sliced format: too many rows"); return BB_ERR_ARG; }
        off[(size_t)nslices] = (unsigned)run;
    }
    const size_t total_rows = off[(size_t)nslices];
    BB_CUDA(cudaMalloc((void**)&f->sl_pairs, (total_rows * 64 + 4) * sizeof(unsigned)));
    if (f->val) BB_CUDA(cudaMalloc((void**)&f->sl_vals, (total_rows * 128 + 4) * sizeof(double)));

    // work partition: CTA ranges of equal cost, cut at slab boundaries into sections, every section into 32 warp strips
    const int ncta = std::max(1, std::min(ctx->sm_count, nslices));
    const i64 slice_cost = (ctx->opt_sell_slice_cost > 0) ? ctx->opt_sell_slice_cost : SELL_SLICE_COST;
    auto cost = [&](i64 s) { return (i64)off[(size_t)s] + (slice_cost - 1) * s; };   // off counts the header rows
    auto cut = [&](i64 lo, i64 hi, i64 target) {      // first s in [lo, hi] with cost(s) >= target
        while (lo < hi) { i64 mid = (lo + hi) >> 1; if (cost(mid) < target) lo = mid + 1; else hi = mid; }
        return lo;
    };
    std::vector<int> cta_sec0((size_t)ncta + 1, 0), sec_slab, sec_wstart;
    const i64 total_cost = cost(nslices);
    // CTA b owns slices [cta_lo[b], cta_lo[b + 1]).  Two ways to draw the ranges:
    //  A  equal cost, wherever the cuts fall: a CTA whose range straddles a slab boundary stages two windows one after the
    //     other.  Measured (profiles/r02_spmv_timeline.md): +3.8 us on the N = 8 shard (15.3 us against 11.5 us for the
    //     one-section CTAs), +2.5 us on C4 -- about SELL_RESTAGE_COST rows of streaming, and it sits on the critical path.
    //  B  slab-aligned: slab s gets k_s CTAs in proportion to its cost (largest remainders, k_s >= 1) and is cut into k_s
    //     equal parts; nobody stages twice, but the parts of different slabs differ by the rounding of k_s.
    // The one with the smaller estimated critical path is taken (B needs at least as many CTAs as non-empty slabs).
    std::vector<i64> cta_lo((size_t)ncta + 1, 0);
    for (int b = 0; b <= ncta; ++b) cta_lo[(size_t)b] = (b == ncta) ? nslices : cut(0, nslices, total_cost * b / ncta);
    {
        bool straddles = false;
        int sl = 0;
        for (int b = 0; b < ncta && !straddles; ++b) {
            if (cta_lo[(size_t)b] >= cta_lo[(size_t)b + 1]) continue;
            while (slab_slice0[(size_t)sl + 1] <= cta_lo[(size_t)b]) ++sl;
            straddles = cta_lo[(size_t)b + 1] > slab_slice0[(size_t)sl + 1];
        }
        int n_nonempty = 0;
        for (int s2 = 0; s2 < nslab; ++s2) n_nonempty += (slab_slice0[(size_t)s2 + 1] > slab_slice0[(size_t)s2]) ? 1 : 0;
        const bool try_aligned = (ctx->opt_sell_partition == 0) ? (straddles && n_nonempty <= ncta) : (ctx->opt_sell_partition == 2 && n_nonempty <= ncta);
        if (try_aligned && total_cost > 0) {
            std::vector<int> k((size_t)nslab, 0);
            std::vector<std::pair<double, int>> rem;
            int used = 0;
            for (int s2 = 0; s2 < nslab; ++s2) {
                const i64 cs = cost(slab_slice0[(size_t)s2 + 1]) - cost(slab_slice0[(size_t)s2]);
                if (slab_slice0[(size_t)s2 + 1] == slab_slice0[(size_t)s2]) continue;
                const double share = (double)cs * ncta / (double)total_cost;
                k[(size_t)s2] = std::max(1, (int)share);
                used += k[(size_t)s2];
                rem.push_back({share - (double)k[(size_t)s2], s2});
            }
            std::sort(rem.begin(), rem.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b2) {
                return a.first > b2.first || (a.first == b2.first && a.second < b2.second); });
            for (size_t q = 0; used < ncta && !rem.empty(); q = (q + 1) % rem.size()) { k[(size_t)rem[q].second] += 1; ++used; }
            while (used > ncta) {              // the floor of 1 CTA per slab overshot: take from the slab with the lightest parts
                int best = -1; double lightest = 0.0;
                for (int s2 = 0; s2 < nslab; ++s2) {
                    if (k[(size_t)s2] < 2) continue;
                    const double per = (double)(cost(slab_slice0[(size_t)s2 + 1]) - cost(slab_slice0[(size_t)s2])) / k[(size_t)s2];
                    if (best < 0 || per < lightest) { best = s2; lightest = per; }
                }
                if (best < 0) break;
                k[(size_t)best] -= 1; --used;
            }
            i64 worst_aligned = 0;
            for (int s2 = 0; s2 < nslab; ++s2)
                if (k[(size_t)s2] > 0)
                    worst_aligned = std::max(worst_aligned, (cost(slab_slice0[(size_t)s2 + 1]) - cost(slab_slice0[(size_t)s2]) + k[(size_t)s2] - 1) / k[(size_t)s2]);
            const i64 worst_equal = (total_cost + ncta - 1) / ncta + (i64)SELL_RESTAGE_COST;
            if (used == ncta && (ctx->opt_sell_partition == 2 || worst_aligned < worst_equal)) {
                int b = 0;
                for (int s2 = 0; s2 < nslab; ++s2) {
                    const i64 s_lo = slab_slice0[(size_t)s2], s_hi = slab_slice0[(size_t)s2 + 1];
                    const i64 c_lo = cost(s_lo), c_hi = cost(s_hi);
                    for (int j = 0; j < k[(size_t)s2]; ++j, ++b)
                        cta_lo[(size_t)b] = (j == 0) ? s_lo : cut(s_lo, s_hi, c_lo + (c_hi - c_lo) * j / k[(size_t)s2]);
                }
                cta_lo[(size_t)ncta] = nslices;
                f->sl_partition = 2;
            }
        }
    }
    int slab = 0;
    std::vector<unsigned> dst((size_t)nslices + 1, 0u);       // first row of every slice in the row stream
    std::vector<int> order, wlist[SELL_WARPS];
    unsigned next_row = 0;
    for (int b = 0; b < ncta; ++b) {
        const i64 c0 = cta_lo[(size_t)b], c1 = cta_lo[(size_t)b + 1];
        cta_sec0[(size_t)b] = (int)sec_slab.size();
        i64 cur = c0;
        while (cur < c1) {
            while (slab + 1 <= nslab && slab_slice0[(size_t)slab + 1] <= cur) ++slab;      // slab containing slice `cur`
            const i64 e = std::min<i64>(c1, slab_slice0[(size_t)slab + 1]);
            sec_slab.push_back(slab);
            if (ctx->opt_sell_lpt != 0) {
                // The 32 warp strips of a section: longest-processing-time-first assignment of its slices (the longest
                // slice is up to 65 rows, a strip ~25 rows on the N = 8 shard and ~180 on C4: contiguous cuts at slice
                // granularity left the slowest warp of a CTA 3-9 us behind the median, profiles/r02_spmv_timeline.md).
                // The slices are then LAID OUT strip by strip, so that a warp still streams one contiguous byte range.
                const int cnt = (int)(e - cur);
                order.resize((size_t)cnt);
                for (int q = 0; q < cnt; ++q) order[(size_t)q] = (int)cur + q;
                std::stable_sort(order.begin(), order.end(), [&](int a2, int b2) { return nr[(size_t)a2] > nr[(size_t)b2]; });
                for (int w = 0; w < SELL_WARPS; ++w) wlist[w].clear();
                // min-heap over (load, warp): deterministic
                std::priority_queue<std::pair<i64, int>, std::vector<std::pair<i64, int>>, std::greater<std::pair<i64, int>>> heap;
                for (int w = 0; w < SELL_WARPS; ++w) heap.push({0, w});
                for (int q = 0; q < cnt; ++q) {
                    const std::pair<i64, int> top = heap.top(); heap.pop();
                    const int sl2 = order[(size_t)q];
                    wlist[top.second].push_back(sl2);
                    heap.push({top.first + nr[(size_t)sl2] + (slice_cost - 1), top.second});
                }
                for (int w = 0; w < SELL_WARPS; ++w) {
                    sec_wstart.push_back((int)next_row);
                    for (int sl2 : wlist[w]) { dst[(size_t)sl2] = next_row; next_row += (unsigned)nr[(size_t)sl2]; }
                }
                sec_wstart.push_back((int)next_row);
            } else {
                const i64 ca = cost(cur), cb = cost(e);
                for (int w = 0; w <= SELL_WARPS; ++w) {
                    i64 sw = (w == 0) ? cur : (w == SELL_WARPS) ? e : cut(cur, e, ca + (cb - ca) * w / SELL_WARPS);
                    sec_wstart.push_back((int)off[(size_t)sw]);       // strips are stored as ROW offsets
                }
                for (i64 q = cur; q < e; ++q) dst[(size_t)q] = off[(size_t)q];
                next_row = off[(size_t)e];
            }
            cur = e;
        }
    }
    cta_sec0[(size_t)ncta] = (int)sec_slab.size();
    if ((size_t)next_row != total_rows) { bb_set_error("sliced format: layout covers %u of %zu rows", next_row, total_rows); return BB_ERR_STATE; }
    dst[(size_t)nslices] = (unsigned)total_rows;
    BB_CUDA(cudaMemcpyAsync(f->sl_off, dst.data(), ((size_t)nslices + 1) * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    if (nslices > 0) {
        static BBDeviceOnce fill_attr = {{0, 0, 0, 0}};
        const size_t fill_smem = sizeof(SellFillSmem) * SELL_FILL_WARPS;
        if (fill_attr.first(ctx->device))
            BB_CUDA(cudaFuncSetAttribute(k_sell_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem));
        k_sell_fill<<<(nslices + SELL_FILL_WARPS - 1) / SELL_FILL_WARPS, SELL_FILL_WARPS * 32, fill_smem, st>>>(
            f->sl_slab_slice0, d_slab_frag0, nslab, nslices, sorted_id, frag_src, frag_len, frag_slot, f->idx, f->val, W,
            f->sl_off, nrows, trash_slot, (int)ctx->opt_bank_permute, f->sl_pairs, f->sl_vals);
        ctx->launches += 1;
    }
    f->sl_ncta = ncta;
    const size_t nsec = sec_slab.size();
    // one allocation: [cta_sec0 (ncta+1) | sec_slab (nsec) | sec_wstart (nsec*33)]
    std::vector<int> packed;
    packed.insert(packed.end(), cta_sec0.begin(), cta_sec0.end());
    packed.insert(packed.end(), sec_slab.begin(), sec_slab.end());
    packed.insert(packed.end(), sec_wstart.begin(), sec_wstart.end());
    BB_CUDA(cudaMalloc((void**)&f->sl_cta_slice0, (packed.size() + 1) * sizeof(int)));
    BB_CUDA(cudaMemcpyAsync(f->sl_cta_slice0, packed.data(), packed.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    f->sl_nsec = (int)nsec;
    BB_CUDA(cudaStreamSynchronize(st));
    return BB_OK;
}

static size_t sell_smem_bytes(const SlabFmt* f) { return (size_t)(f->W + 2) * sizeof(double) + 16; }

int bb_sell_launch(bb_ctx* ctx, SlabFmt* f, const double* gvec, const int* done_flag, bool skip_overflow_add) {
    if (f->nslices == 0) return BB_OK;        // no nnz: part stays zero
    const size_t smem = sell_smem_bytes(f);
    if (smem > ctx->smem_optin) { bb_set_error("sliced spmv: shared memory %zu exceeds %zu", smem, ctx->smem_optin); return BB_ERR_ARG; }
    static BBDeviceOnce attr_set = {{0, 0, 0, 0}};
    if (attr_set.first(ctx->device)) {
        const int mx = (int)ctx->smem_optin;
        BB_CUDA(cudaFuncSetAttribute(k_sell_spmv<true, SELL_RING_BIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        BB_CUDA(cudaFuncSetAttribute(k_sell_spmv<false, SELL_RING_VAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        BB_CUDA(bb_prefer_max_smem(ctx, k_sell_spmv<true, SELL_RING_BIN>));
        BB_CUDA(bb_prefer_max_smem(ctx, k_sell_spmv<false, SELL_RING_VAL>));
        BB_CUDA(bb_prefer_max_smem(ctx, k_sell_ovf_add));
    }
    // TMA bulk staging needs a 16-byte aligned source (the window base is a multiple of W, W is a multiple of 32)
    const int use_bulk = (ctx->opt_spmv_bulk != 0 && (reinterpret_cast<uintptr_t>(gvec) & 15u) == 0 && (f->W & 1) == 0) ? 1 : 0;
    const int* cta_sec0 = f->sl_cta_slice0;
    const int* sec_slab = cta_sec0 + (f->sl_ncta + 1);
    const int* sec_wstart = sec_slab + f->sl_nsec;
    const uint2* rows = reinterpret_cast<const uint2*>(f->sl_pairs);
    if (f->sl_vals == nullptr)
        BB_CUDA(bb_launch(ctx, true, k_sell_spmv<true, SELL_RING_BIN>, dim3(f->sl_ncta), dim3(SELL_THREADS), smem,
                          rows, nullptr, cta_sec0, sec_slab, sec_wstart, gvec, f->W, f->n_gather, use_bulk, f->part, done_flag, nullptr));
    else
        BB_CUDA(bb_launch(ctx, true, k_sell_spmv<false, SELL_RING_VAL>, dim3(f->sl_ncta), dim3(SELL_THREADS), smem,
                          rows, f->sl_vals, cta_sec0, sec_slab, sec_wstart, gvec, f->W, f->n_gather, use_bulk, f->part, done_flag, nullptr));
    BB_LAUNCHED(ctx);
    if (f->n_ovf_pieces > 0 && !skip_overflow_add) {
        BB_CUDA(bb_launch(ctx, true, k_sell_ovf_add, dim3((unsigned)(((i64)f->n_ovf_pieces * 32 + 255) / 256)), dim3(256), 0,
                          f->ovf_piece, f->ovf_first, f->n_ovf_pieces, (i64)f->nslab * f->n_seg, f->part, done_flag));
        BB_LAUNCHED(ctx);
    }
    return BB_OK;
}

// ---- profiling aid: per-CTA time line of one SpMV launch (debug instantiation of the kernel) ------------------------
// out[cta * (8 + 32) + k]: k = 0 kernel entry, 1 after the dependency wait, 2 first window staged, 3 CTA end, 4 number of
// sections, 8.. = end of each warp's last strip; all in ns of %globaltimer.  which: 0 = dot format, 1 = Tdot format.
extern "C" int bb_spmv_timeline(bb_mat* m, int which, int flush_l2, uint64_t* out, int64_t capacity, int* ncta_out) {
    BB_ARG(m && out && ncta_out, "mat/out/ncta_out");
    BB_ARG(m->is_sparse, "sparse matrix needed");
    bb_ctx* ctx = m->ctx;
    BB_CUDA(cudaSetDevice(ctx->device));
    SlabFmt* f = which ? &m->ftdot : &m->fdot;
    BB_ARG(f->variant == 1 && f->nslices > 0 && f->sl_vals == nullptr, "pattern-only sliced format needed");
    const int ncta = f->sl_ncta;
    BB_ARG(capacity >= (int64_t)ncta * SELL_DBG_WORDS, "capacity");
    const size_t smem = sell_smem_bytes(f);
    static BBDeviceOnce attr_set = {{0, 0, 0, 0}};
    if (attr_set.first(ctx->device))
        BB_CUDA(cudaFuncSetAttribute(k_sell_spmv<true, SELL_RING_BIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    unsigned long long* dbg = nullptr;
    BB_TRY(bb_ctx_scratch(ctx, 3, (size_t)ncta * SELL_DBG_WORDS * sizeof(unsigned long long), (void**)&dbg));
    const double* gvec = which ? m->eps_n : (m->sv + m->add_intercept);
    const int use_bulk = (ctx->opt_spmv_bulk != 0 && (reinterpret_cast<uintptr_t>(gvec) & 15u) == 0 && (f->W & 1) == 0) ? 1 : 0;
    const int* cta_sec0 = f->sl_cta_slice0;
    const int* sec_slab = cta_sec0 + (f->sl_ncta + 1);
    const int* sec_wstart = sec_slab + f->sl_nsec;
    for (int rep = 0; rep < 3; ++rep) {       // the last repetition is the one reported
        if (flush_l2) {
            if (!ctx->flush_buf) { BB_CUDA(cudaMalloc(&ctx->flush_buf, (size_t)512 << 20)); ctx->flush_bytes = (size_t)512 << 20; }
            BB_CUDA(cudaMemsetAsync(ctx->flush_buf, rep, ctx->flush_bytes, ctx->stream));
        }
        BB_CUDA(cudaMemsetAsync(dbg, 0, (size_t)ncta * SELL_DBG_WORDS * sizeof(unsigned long long), ctx->stream));
        k_sell_spmv<true, SELL_RING_BIN, true><<<ncta, SELL_THREADS, smem, ctx->stream>>>(
            reinterpret_cast<const uint2*>(f->sl_pairs), nullptr, cta_sec0, sec_slab, sec_wstart, gvec, f->W, f->n_gather, use_bulk,
            f->part, nullptr, dbg);
        BB_LAUNCHED(ctx);
    }
    BB_CUDA(cudaMemcpyAsync(out, dbg, (size_t)ncta * SELL_DBG_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    *ncta_out = ncta;
    return BB_OK;
}
