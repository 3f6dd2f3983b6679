// Sparse design matrix: CSR upload, bit-exact CSC construction, slab formats and the SpMV kernels.
//
// Reference seams: bayesbridge/design_matrix/sparse_matrix.py:21-49 (construction; X.tocsr()),
// :68-101 (dot), :103-129 (Tdot), :164-177 (fisher diag).  scipy's csr_tocsc (called by the
// reference's `X.T.dot`) is a stable counting sort by column; the device CSC below is a stable
// LSD radix sort by column of the CSR nnz sequence, which yields the same arrays bit for bit.
#include "bb_internal.cuh"
#include <cub/cub.cuh>
#include <vector>
#include <algorithm>
#include <stdlib.h>

constexpr int SPMV_THREADS = 1024;                      // one CTA per SM, 32 independent warps
constexpr int SPMV_WARPS = SPMV_THREADS / 32;
constexpr int SPMV_ITEMS = 8;                           // nnz per lane per tile
constexpr int SPMV_TILE = 32 * SPMV_ITEMS;              // nnz per (warp) tile

// ------------------------------------------------------------------------------------------
// setup kernels (run once per matrix)
__global__ void k_iota(int* a, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int)i;
}

// seg_of[k] = segment containing nnz k  (largest s with ptr[s] <= k)
__global__ void k_expand_ptr(const int* __restrict__ ptr, i64 n_seg, i64 nnz, int* __restrict__ seg_of) {
    i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    i64 lo = 0, hi = n_seg;   // find first s in [0, n_seg] with ptr[s] > k, answer s-1
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        if (ptr[mid] <= (int)k) lo = mid + 1; else hi = mid;
    }
    seg_of[k] = (int)(lo - 1);
}

__global__ void k_gather_i(const int* __restrict__ src, const int* __restrict__ perm, i64 n, int* __restrict__ dst) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}
__global__ void k_gather_d(const double* __restrict__ src, const int* __restrict__ perm, i64 n, double* __restrict__ dst) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}
__global__ void k_slab_key(const int* __restrict__ idx, i64 n, int W, int* __restrict__ key) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = idx[i] / W;
}
__global__ void k_vkey(const int* __restrict__ slab_sorted, const int* __restrict__ seg_of,
                       const int* __restrict__ perm, i64 n, i64 n_seg, i64* __restrict__ vkey) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vkey[i] = (i64)slab_sorted[i] * n_seg + seg_of[perm[i]];
}
// out[v] = first position m with key[m] >= v, for v in [0, nv]
template <typename K>
__global__ void k_lower_bound(const K* __restrict__ key, i64 n, i64 nv, int* __restrict__ out) {
    i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nv) return;
    i64 lo = 0, hi = n;
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        if ((i64)key[mid] < v) lo = mid + 1; else hi = mid;
    }
    out[v] = (int)lo;
}

// slab-major copy with every slab start padded to a multiple of 4 nnz (128-bit loads)
__global__ void k_scatter_padded(const int* __restrict__ cidx, const double* __restrict__ cval,
                                 const int* __restrict__ perm, const int* __restrict__ slab_sorted,
                                 const int* __restrict__ delta, i64 n, int* __restrict__ oidx, double* __restrict__ oval) {
    i64 m = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    i64 dst = m + delta[slab_sorted[m]];
    int src = perm[m];
    oidx[dst] = cidx[src];
    if (cval) oval[dst] = cval[src];
}
// ptr[v] (position in the unpadded sorted sequence) -> position in the padded arrays
__global__ void k_shift_ptr(int* __restrict__ ptr, i64 V, i64 n_seg, const int* __restrict__ delta) {
    i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > V) return;
    i64 slab = (n_seg > 0) ? (v / n_seg) : 0;             // the sentinel v == V indexes delta[nslab] (= last slab's shift)
    ptr[v] += delta[slab];
}

// per-tile ownership: slab/last flags are packed in vlo/vhi on input
__global__ void k_tile_meta(const int* __restrict__ ptr, i64 V, i64 n_seg, int ntiles,
                            TileMeta* __restrict__ tiles, int* __restrict__ head_seg) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    TileMeta m = tiles[t];
    int slab = m.vlo, is_last = m.vhi;
    i64 vbeg = (i64)slab * n_seg, vend = vbeg + n_seg;
    // vlo = first v in [vbeg, vend] with ptr[v] >= start
    i64 lo = vbeg, hi = vend;
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        if (ptr[mid] < m.start) lo = mid + 1; else hi = mid;
    }
    i64 vlo = lo;
    i64 vhi = vend;
    if (!is_last) {
        lo = vlo; hi = vend;
        while (lo < hi) {
            i64 mid = (lo + hi) >> 1;
            if (ptr[mid] < m.end) lo = mid + 1; else hi = mid;
        }
        vhi = lo;
    }
    m.vlo = (int)vlo;
    m.vhi = (int)vhi;
    tiles[t] = m;
    head_seg[t] = (ptr[vlo] > m.start) ? (int)(vlo - 1) : -1;
}

__global__ void k_compact_meta(const TileMeta* __restrict__ tiles, int ntiles, int2* __restrict__ tmeta) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntiles) tmeta[t] = make_int2(tiles[t].vlo, tiles[t].vhi - tiles[t].vlo);
}

// Bank-aware ordering of the nnz (build time, one thread per tile).  The SpMV kernel gathers the staged vector
// with one 8-byte shared-memory load per nnz: load j of lane l reads element 8l + j of the tile, and the 16
// lanes of a half-warp are served together, one wavefront per distinct address that falls into the same
// 8-byte bank (index mod 16).  With the nnz in canonical order the gather indices of those 16 lanes are
// random, so a load needs ~3 wavefronts per half-warp instead of 1, and the gathers are what saturates the
// shared-memory pipe.  A sum does not depend (beyond rounding) on the order of its terms, so the nnz of a
// piece (one segment inside one tile) may be stored in any order: for every position, in order, pick among
// the piece's remaining nnz one whose bank is least used so far by the (load j, half-warp) group of that
// position.  Pieces, cut points and tile boundaries do not move; the canonical CSR/CSC images are separate
// arrays and stay bit-exact.
__global__ void __launch_bounds__(64)
k_bank_permute(const TileMeta* __restrict__ tiles, int ntiles, const int* __restrict__ ptr, i64 V,
               const int* __restrict__ idx_in, const double* __restrict__ val_in,
               int* __restrict__ idx_out, double* __restrict__ val_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int start = tiles[t].start, len = tiles[t].end - tiles[t].start;
    if (len <= 1) return;
    int e[SPMV_TILE];
    unsigned char src[SPMV_TILE];
    unsigned char occ[2 * SPMV_ITEMS * 16];
    for (int k = 0; k < len; ++k) { e[k] = idx_in[start + k]; src[k] = (unsigned char)k; }
    for (int k = 0; k < 2 * SPMV_ITEMS * 16; ++k) occ[k] = 0;
    // virtual segment that contains nnz `start`: the largest v < V with ptr[v] <= start
    i64 v = 0;
    {
        i64 lo = 0, hi = V;
        while (hi - lo > 1) { i64 mid = (lo + hi) >> 1; if (ptr[mid] <= start) lo = mid; else hi = mid; }
        v = lo;
    }
    int a = 0;
    while (a < len) {
        while (v + 1 < V && ptr[v + 1] <= start + a) ++v;
        int b = min(ptr[v + 1] - start, len);
        if (b <= a) b = len;                                  // defensive: never loop on an empty piece
        for (int k = a; k < b; ++k) {
            unsigned char* og = occ + ((((k & (SPMV_ITEMS - 1)) << 1) | (k >> 7)) << 4);
            int best = k, bo = og[e[k] & 15];
            for (int m = k + 1; bo > 0 && m < b; ++m) {
                int o = og[e[m] & 15];
                if (o < bo) { bo = o; best = m; }
            }
            int te = e[k]; e[k] = e[best]; e[best] = te;
            unsigned char ts = src[k]; src[k] = src[best]; src[best] = ts;
            og[e[k] & 15] = (unsigned char)(bo + 1);
        }
        a = b;
    }
    for (int k = 0; k < len; ++k) idx_out[start + k] = e[k];
    if (val_in != nullptr)
        for (int k = 0; k < len; ++k) val_out[start + k] = val_in[start + src[k]];
}

// ------------------------------------------------------------------------------------------
// The SpMV kernel.  The tile sequence (all slabs concatenated) is cut into gridDim.x equal contiguous
// ranges, one per CTA (grid = #SMs, one CTA per SM): perfectly balanced in nnz.  A CTA walks its
// range slab by slab: it stages the slab of the gather vector in shared memory (block barrier), then
// its 32 warps run independently over the slab's tiles: coalesced loads of (idx, val) into registers,
// software-prefetched one tile ahead (the tile metadata two ahead, so no load in the steady state
// waits on another), products into a per-warp shared buffer, and a segmented reduction in which
// groups of G = 32/2^ceil(log2(pieces)) lanes each sum one piece.  The pieces of a tile are delimited
// by its cut points c_k = min(ptr[vlo+k], end) - start: piece 0 is the head (a segment continued from
// the previous tile -> head_part), piece k>0 is owned segment vlo+k-1 (-> part).  The reduction shape
// depends only on static metadata, so results are bit-reproducible.
// Loads of one tile in BLOCKED layout: lane l owns the 8 consecutive nnz [start+8l, start+8l+8), fetched
// with 128-bit loads (tile starts are multiples of 4 nnz from a 256 B-aligned base).  A partial tile (the
// last of a slab) uses clamped scalar loads: elements past `end` duplicate the last valid nnz, and their
// products land beyond every piece, where nothing reads them.  Indices and values are fetched separately so
// that the index prefetch can be issued early and the value prefetch once the product registers are dead.
__device__ __forceinline__ void tile_fetch_idx(int (&ri)[SPMV_ITEMS], int start, int end,
                                               const int* __restrict__ idx, int lane) {
    if (end - start == SPMV_TILE) {
        const int4* ip = reinterpret_cast<const int4*>(idx + start) + lane * 2;
        int4 a = ip[0], b = ip[1];
        ri[0] = a.x; ri[1] = a.y; ri[2] = a.z; ri[3] = a.w;
        ri[4] = b.x; ri[5] = b.y; ri[6] = b.z; ri[7] = b.w;
    } else if (end > start) {
#pragma unroll
        for (int j = 0; j < SPMV_ITEMS; ++j) ri[j] = idx[min(start + lane * SPMV_ITEMS + j, end - 1)];
    }
}
__device__ __forceinline__ void tile_fetch_val(double (&rv)[SPMV_ITEMS], int start, int end,
                                               const double* __restrict__ val, int lane) {
    if (end - start == SPMV_TILE) {
        const double2* vp = reinterpret_cast<const double2*>(val + start) + lane * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) { double2 v = vp[q]; rv[2 * q] = v.x; rv[2 * q + 1] = v.y; }
    } else if (end > start) {
#pragma unroll
        for (int j = 0; j < SPMV_ITEMS; ++j) rv[j] = val[min(start + lane * SPMV_ITEMS + j, end - 1)];
    }
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// The SpMV kernel.  Per 256-nnz tile a warp computes the products, marks the
// starts of the tile's pieces in a 256-bit head-flag map (cut points c_k = min(ptr[vlo+k], end) - start),
// runs a segmented inclusive scan (8 serial steps per lane + 5 shuffle steps across lanes), parks the
// lane-local prefix values and the per-lane carries in shared memory, and lane k picks the total of piece k
// at position c_k - 1 (adding the carry of that lane when no piece starts in the lane at or before it).
template <bool BINARY, bool STAGE>
__global__ void __launch_bounds__(SPMV_THREADS, 1)
k_seg_spmv(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
           const int2* __restrict__ tmeta, const int* __restrict__ slab_tile0, const int* __restrict__ slab_nnz0,
           const int* __restrict__ slab_nnz1, int nslab, int ntiles,
           const double* __restrict__ gvec, int W, i64 n_gather, int wstage,
           double* __restrict__ part, double* __restrict__ head_part, const int* __restrict__ done_flag)
{
    if (done_flag != nullptr && *done_flag) return;
    extern __shared__ double smem[];
    double* sv = smem;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    double* wprod = smem + wstage + warp * SPMV_TILE;                       // per-warp prefix buffer
    double* wcarry = smem + wstage + SPMV_WARPS * SPMV_TILE + warp * 32;     // per-lane carries
    unsigned* fl = reinterpret_cast<unsigned*>(smem + wstage + SPMV_WARPS * (SPMV_TILE + 32)) + warp * 8;   // head flags
    const int t_lo = (int)((i64)ntiles * blockIdx.x / gridDim.x);
    const int t_hi = (int)((i64)ntiles * (blockIdx.x + 1) / gridDim.x);
    if (t_lo >= t_hi) return;
    int slab;
    {
        int lo = 0, hi = nslab;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (slab_tile0[mid] <= t_lo) lo = mid; else hi = mid; }
        slab = lo;
    }
    int cur = t_lo;
    while (cur < t_hi) {
        const int s_t0 = slab_tile0[slab];
        const int sec_end = min(t_hi, slab_tile0[slab + 1]);
        const int nnz0 = slab_nnz0[slab], nnz1 = slab_nnz1[slab];
        const i64 gbase = (i64)slab * W;
        if (STAGE) {
            __syncthreads();                               // previous section's readers are done
            i64 rem = n_gather - gbase;
            const int wlen = rem < (i64)W ? (int)rem : W;
            const double* src = gvec + gbase;
            int i = tid;
            for (; i + 3 * SPMV_THREADS < wlen; i += 4 * SPMV_THREADS) {
                double a0 = src[i], a1 = src[i + SPMV_THREADS], a2 = src[i + 2 * SPMV_THREADS], a3 = src[i + 3 * SPMV_THREADS];
                sv[i] = a0; sv[i + SPMV_THREADS] = a1; sv[i + 2 * SPMV_THREADS] = a2; sv[i + 3 * SPMV_THREADS] = a3;
            }
            for (; i < wlen; i += SPMV_THREADS) sv[i] = src[i];
        }
        // byte address such that (sbase + 8*global_index) addresses the staged entry
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(sv) - (unsigned)((unsigned)gbase << 3);
        int t = cur + warp;
        int ri[SPMV_ITEMS];
        double rv[SPMV_ITEMS];
        int2 m_cur = make_int2(0, 0), m_next = make_int2(0, 0);
        int cut = 0;
        if (t < sec_end) {
            const int st0 = nnz0 + (t - s_t0) * SPMV_TILE;
            tile_fetch_idx(ri, st0, min(st0 + SPMV_TILE, nnz1), idx, lane);
            if (!BINARY) tile_fetch_val(rv, st0, min(st0 + SPMV_TILE, nnz1), val, lane);
            m_cur = tmeta[t];
            if (t + SPMV_WARPS < sec_end) m_next = tmeta[t + SPMV_WARPS];
            cut = (lane <= m_cur.y) ? ptr[m_cur.x + lane] : 0x7fffffff;
        }
        if (STAGE) __syncthreads();

        while (t < sec_end) {
            const int start = nnz0 + (t - s_t0) * SPMV_TILE;
            const int end = min(start + SPMV_TILE, nnz1);
            const int len = end - start;
            const int vlo = m_cur.x, nown = m_cur.y;
            const int c = min(cut, end) - start;            // cut point k = lane (valid for lane <= nown)
            const int tn = t + SPMV_WARPS;
            const int st1 = nnz0 + (tn - s_t0) * SPMV_TILE;
            const int en1 = min(st1 + SPMV_TILE, nnz1);
            double p[SPMV_ITEMS];
            if (len > 0) {
                // products (blocked: p[j] is element 8*lane + j)
#pragma unroll
                for (int j = 0; j < SPMV_ITEMS; ++j) {
                    double g = STAGE ? lds_f64(sbase + ((unsigned)ri[j] << 3)) : __ldg(gvec + ri[j]);
                    p[j] = BINARY ? g : rv[j] * g;
                }
                // early prefetch (indices, cut points, metadata) of the next tile: `ri` is dead now
                if (tn < sec_end) {
                    tile_fetch_idx(ri, st1, en1, idx, lane);
                    m_cur = m_next;
                    cut = (lane <= m_cur.y) ? ptr[m_cur.x + lane] : 0x7fffffff;
                    if (tn + SPMV_WARPS < sec_end) m_next = tmeta[tn + SPMV_WARPS];
                }
                // head flags: piece k+1 starts at c_k
                __syncwarp();
                if (lane < 8) fl[lane] = 0u;
                __syncwarp();
                if (lane < nown && c < len) atomicOr(&fl[c >> 5], 1u << (c & 31));
                if (nown > 32) {
                    for (int k = 32 + lane; k < nown; k += 32) {
                        int cc = min(ptr[vlo + k], end) - start;
                        if (cc < len) atomicOr(&fl[cc >> 5], 1u << (cc & 31));
                    }
                }
                __syncwarp();
                const unsigned f = (fl[lane >> 2] >> ((lane & 3) * 8)) & 0xffu;
                // segmented inclusive scan: serial inside the lane ...
                double run = 0.0;
#pragma unroll
                for (int j = 0; j < SPMV_ITEMS; ++j) {
                    if ((f >> j) & 1u) run = 0.0;
                    run += p[j];
                    wprod[j * 32 + lane] = run;           // lane-local prefix of element 8*lane + j
                }
                // late prefetch (values): the product registers are dead now
                if (!BINARY && tn < sec_end) tile_fetch_val(rv, st1, en1, val, lane);
                // ... and across lanes (segments start at lanes that contain a head)
                // (lane l takes the partial sum of lane l-d unless a head lies in lanes (l-d, l]: one ballot
                //  replaces a shuffled flag per step)
                double x = run;
                const unsigned hm = __ballot_sync(0xffffffffu, f != 0u);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    double y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= d && ((hm >> (lane - d + 1)) & ((1u << d) - 1u)) == 0u) x += y;
                }
                double carry = __shfl_up_sync(0xffffffffu, x, 1);
                if (lane == 0) carry = 0.0;
                wcarry[lane] = carry;                       // sum of the open piece over the previous lanes
                __syncwarp();
                // piece k = [c_{k-1}, c_k): total = prefix at e = c_k - 1 (+ carry of e's lane if the piece began before it)
                {
                    int prev = __shfl_up_sync(0xffffffffu, c, 1);
                    if (lane == 0) prev = 0;
                    if (lane <= nown) {
                        double tot = 0.0;
                        if (c > prev) {
                            int e = c - 1, le = e >> 3, je = e & 7;
                            unsigned fe = (fl[le >> 2] >> ((le & 3) * 8)) & 0xffu;
                            tot = wprod[je * 32 + le];
                            if ((fe & ((2u << je) - 1u)) == 0u) tot += wcarry[le];
                        }
                        if (lane == 0) head_part[t] = tot; else part[vlo + lane - 1] = tot;
                    }
                }
                if (nown >= 32) {                           // rare: more than 32 pieces in a tile
                    for (int k = 32 + lane; k <= nown; k += 32) {
                        int ck = min(ptr[vlo + k], end) - start;
                        int pk = min(ptr[vlo + k - 1], end) - start;
                        double tot = 0.0;
                        if (ck > pk) {
                            int e = ck - 1, le = e >> 3, je = e & 7;
                            unsigned fe = (fl[le >> 2] >> ((le & 3) * 8)) & 0xffu;
                            tot = wprod[je * 32 + le];
                            if ((fe & ((2u << je) - 1u)) == 0u) tot += wcarry[le];
                        }
                        part[vlo + k - 1] = tot;
                    }
                }
            } else {
                // empty tile (a slab without nnz): every owned segment is empty
                if (lane == 0) head_part[t] = 0.0;
                for (int k = 1 + lane; k <= nown; k += 32) part[vlo + k - 1] = 0.0;
                if (tn < sec_end) {
                    tile_fetch_idx(ri, st1, en1, idx, lane);
                    if (!BINARY) tile_fetch_val(rv, st1, en1, val, lane);
                    m_cur = m_next;
                    cut = (lane <= m_cur.y) ? ptr[m_cur.x + lane] : 0x7fffffff;
                    if (tn + SPMV_WARPS < sec_end) m_next = tmeta[tn + SPMV_WARPS];
                }
            }
            t = tn;
        }
        cur = sec_end;
        slab += 1;
    }
}

// adds the head partials (tails of segments that straddle tiles) in tile order
__global__ void k_fixup(const int* __restrict__ head_seg, int ntiles, const double* __restrict__ head_part,
                        double* __restrict__ part, const int* __restrict__ done_flag) {
    if (done_flag != nullptr && *done_flag) return;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    int v = head_seg[t];
    if (v < 0) return;
    if (t > 0 && head_seg[t - 1] == v) return;   // not the first continuation tile
    double acc = 0.0;
    for (int tt = t; tt < ntiles && head_seg[tt] == v; ++tt) acc += head_part[tt];
    part[v] += acc;
}

// ------------------------------------------------------------------------------------------
// element-wise kernels around the SpMV
// sv[j] = (scale? scale[j]:1) * v[j]; shift partial = sum_{j<icpt} sv[j] - sum_j c[j] sv[icpt+j]
__global__ void k_prepare(const double* __restrict__ v, const double* __restrict__ scale, i64 P, int icpt,
                          const double* __restrict__ c, double* __restrict__ sv, double* __restrict__ red_shift,
                          const int* __restrict__ done_flag) {
    if (done_flag != nullptr && *done_flag) return;
    __shared__ double sm[33];
    double acc = 0.0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (i64)gridDim.x * blockDim.x) {
        double x = scale ? scale[j] * v[j] : v[j];
        sv[j] = x;
        acc += (j < icpt) ? x : -c[j - icpt] * x;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red_shift[blockIdx.x] = acc;
}

// u_i = shift + sum_s part[s*n+i];  MODE 0: out=u ; MODE 1: out = omega*u, partial sums of out
template <int MODE>
__global__ void k_dot_finish(const double* __restrict__ part, int nslab, i64 n,
                             const double* __restrict__ red_shift, int nshift,
                             const double* __restrict__ omega, const double* __restrict__ omega_scalar,
                             double* __restrict__ out, double* __restrict__ u_out, double* __restrict__ red_w,
                             const int* __restrict__ done_flag) {
    pdl_trigger();
    pdl_wait();
    if (done_flag != nullptr && *done_flag) return;
    __shared__ double sm[33];
    const double shift = warp_sum_partials(red_shift, nshift);
    double acc = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double u = 0.0;
        for (int s = 0; s < nslab; ++s) u += part[(i64)s * n + i];
        u += shift;
        if (MODE == 0) {
            out[i] = u;
        } else {
            double w = (omega ? omega[i] : omega_scalar[0]) * u;
            out[i] = w;
            if (u_out) u_out[i] = u;
            acc += w;
        }
    }
    if (MODE == 1) {
        acc = block_sum(acc, sm);
        if (threadIdx.x == 0) red_w[blockIdx.x] = acc;
    }
}

__global__ void k_sum_partials(const double* __restrict__ w, i64 n, double* __restrict__ red) {
    __shared__ double sm[33];
    double acc = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) acc += w[i];
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) red[blockIdx.x] = acc;
}

// traw[0] = sum of the w partials; traw[1+j] = sum_r part[r*p+j]
// Block = 32 columns x 8 slab groups: each thread sums slabs r = gy, gy+8, ... (independent loads in flight),
// then the 8 partial sums of a column are added in fixed order through shared memory.
// With a peer-memory exchange attached (pub != nullptr) the result goes straight into this rank's exchange
// slot and the last block signals every peer: the local reduction and the publication are one kernel.
__global__ void __launch_bounds__(256)
k_tdot_collect(const double* __restrict__ part, int nslab, i64 p,
               const double* __restrict__ red_w, int nred, double* __restrict__ traw,
               const int* __restrict__ done_flag, const P2PView* __restrict__ pub_ptr) {
    pdl_trigger();
    pdl_wait();
    if (done_flag != nullptr && *done_flag) return;
    __shared__ double sm[8][33];
    P2PView pub;
    if (pub_ptr != nullptr) { pub = *pub_ptr; traw = p2p_publish_slot(pub); }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        double s = warp_sum_partials(red_w, nred);
        if (threadIdx.x == 0) traw[0] = s;
    }
    const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
    for (i64 j0 = (i64)blockIdx.x * 32; j0 < p; j0 += (i64)gridDim.x * 32) {
        const i64 j = j0 + cx;
        double t0 = 0.0, t1 = 0.0;
        if (j < p) {
            int r = gy;
            for (; r + 8 < nslab; r += 16) { t0 += part[(i64)r * p + j]; t1 += part[(i64)(r + 8) * p + j]; }
            if (r < nslab) t0 += part[(i64)r * p + j];
        }
        sm[gy][cx] = t0 + t1;
        __syncthreads();
        if (gy == 0 && j < p) {
            double t = sm[0][cx];
#pragma unroll
            for (int g = 1; g < 8; ++g) t += sm[g][cx];
            traw[1 + j] = t;
        }
        __syncthreads();
    }
    if (pub_ptr != nullptr) p2p_publish_done(pub);
}

// t_P from the (allreduced) traw: t[0] = sum w (intercept); t[icpt+j] = traw[1+j] - sum_w * c[j]
__global__ void k_tdot_finish(const double* __restrict__ traw, i64 p, int icpt, const double* __restrict__ c,
                              double* __restrict__ tP) {
    const double sw = traw[0];
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < p + icpt; j += (i64)gridDim.x * blockDim.x) {
        if (j < icpt) tP[j] = sw;
        else tP[j] = traw[1 + (j - icpt)] - sw * c[j - icpt];
    }
}

__global__ void k_mul(const double* __restrict__ a, const double* __restrict__ b, i64 n, double* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) out[i] = a[i] * b[i];
}

// fisher diag: sum_i weight_i X_ij^2 (values squared on the fly): reuse the Tdot format with a
// squared-value pass; implemented as a column-parallel gather over the canonical CSC.
__global__ void k_fisher_diag_csc(const int* __restrict__ cptr, const int* __restrict__ cidx, const double* __restrict__ cval,
                                  i64 p, const double* __restrict__ weight, double* __restrict__ d2, double* __restrict__ d1) {
    // one warp per column
    i64 j = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (j >= p) return;
    double a2 = 0.0, a1 = 0.0;
    for (int k = cptr[j] + lane; k < cptr[j + 1]; k += 32) {
        double x = cval ? cval[k] : 1.0;
        double w = weight[cidx[k]];
        a2 += w * x * x;
        a1 += w * x;
    }
    a2 = warp_sum(a2);
    a1 = warp_sum(a1);
    if (lane == 0) { d2[j] = a2; d1[j] = a1; }
}

// ------------------------------------------------------------------------------------------
static inline int grid_for(i64 n, int threads, int cap) {
    i64 g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

template <typename KeyT>
static int sort_pairs(bb_ctx* ctx, const KeyT* kin, KeyT* kout, const int* vin, int* vout, i64 n, int end_bit) {
    size_t tmp_bytes = 0;
    BB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin, kout, vin, vout, (int)n, 0, end_bit, ctx->stream));
    void* tmp = nullptr;
    BB_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, (int)n, 0, end_bit, ctx->stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        bb_set_error("radix sort failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return BB_ERR_CUDA;
    }
    ctx->launches += 4;
    return BB_OK;
}

static int bits_for(i64 maxval) {   // number of bits needed to represent values in [0, maxval]
    int b = 1;
    while (b < 63 && ((i64)1 << b) <= maxval) ++b;
    return b;
}

int bb_slab_free(SlabFmt* f) {
    bb_sell_free(f);
    if (f->owns_arrays) {
        if (f->ptr) cudaFree(f->ptr);
        if (f->idx) cudaFree(f->idx);
        if (f->val) cudaFree(f->val);
    }
    if (f->tiles) cudaFree(f->tiles);
    if (f->head_seg) cudaFree(f->head_seg);
    if (f->tmeta) cudaFree(f->tmeta);
    if (f->slab_tile0) cudaFree(f->slab_tile0);
    if (f->slab_nnz0) cudaFree(f->slab_nnz0);
    if (f->slab_nnz1) cudaFree(f->slab_nnz1);
    if (f->part) cudaFree(f->part);
    if (f->head_part) cudaFree(f->head_part);
    memset(f, 0, sizeof(*f));
    return BB_OK;
}

static i64 max_stage_width(bb_ctx* ctx) {
    i64 avail = (i64)ctx->smem_optin - (i64)(SPMV_WARPS * (SPMV_TILE + 32)) * 8 - SPMV_WARPS * 8 * 4 - 64;
    i64 w = avail / 8;
    w &= ~(i64)31;
    if (w < 32) w = 32;
    return w;
}

// Build a slab format from a canonical compressed matrix (ptr[n_seg+1], idx, val).
static int build_slab_format(bb_ctx* ctx, const int* cptr, const int* cidx, const double* cval,
                             i64 n_seg, i64 n_gather, i64 nnz, bool staged, SlabFmt* f) {
    memset(f, 0, sizeof(*f));
    f->staged = staged;
    cudaStream_t st = ctx->stream;
    // sub-warp-per-segment kernel on the canonical image: forced by spmv_variant = 2, or picked for small, L2-resident
    // problems (option rowwise_max_nnz: largest nnz it is used for; 0 = never)
    if (ctx->opt_spmv_variant == 2 || (ctx->opt_spmv_variant == 1 && nnz <= ctx->opt_rowwise_max_nnz)) {
        f->variant = 2;
        f->nslab = 1; f->n_seg = n_seg; f->n_gather = n_gather; f->W = (int)(n_gather < ((i64)1 << 30) ? n_gather : ((i64)1 << 30)); f->nnz = nnz;
        f->ptr = (int*)cptr; f->idx = (int*)cidx; f->val = (double*)cval;
        f->owns_arrays = false;
        BB_CUDA(cudaMalloc((void**)&f->part, (size_t)(n_seg > 0 ? n_seg : 1) * sizeof(double)));
        BB_CUDA(cudaMemsetAsync(f->part, 0, (size_t)(n_seg > 0 ? n_seg : 1) * sizeof(double), st));
        return BB_OK;
    }
    // the sliced kernel (bb_sell.cu) always stages its window; the tile kernel also runs unstaged (one slab, gathers through L2)
    const bool sliced = staged && ctx->opt_spmv_variant == 1;
    i64 wmax = sliced ? bb_sell_max_width(ctx) : max_stage_width(ctx);
    if (!staged) wmax = ((i64)1 << 40);     // gathers go through L2: no slab limit
    if (ctx->opt_slab_width > 0) {
        i64 w = (ctx->opt_slab_width + 31) & ~(i64)31;
        if (w < wmax) wmax = w;
    }
    i64 ng = n_gather > 0 ? n_gather : 1;
    i64 nslab = (ng + wmax - 1) / wmax;
    i64 W = ((ng + nslab - 1) / nslab + 31) & ~(i64)31;
    if (W > wmax) W = wmax;
    nslab = (ng + W - 1) / W;
    i64 V = nslab * n_seg;
    if (V + 1 >= ((i64)1 << 31)) { bb_set_error("slab format too large (nslab*n_seg = %lld)", (long long)V); return BB_ERR_ARG; }
    f->nslab = (int)nslab;
    f->n_seg = n_seg;
    f->n_gather = n_gather;
    f->W = (int)W;
    f->nnz = nnz;
    const int TB = 256;
    std::vector<int> padded_start((size_t)nslab + 1, 0), padded_end((size_t)nslab + 1, 0);
    i64 padded_total = nnz;
    if (nslab == 1) {
        f->ptr = (int*)cptr; f->idx = (int*)cidx; f->val = (double*)cval;
        f->owns_arrays = false;
        padded_start[0] = 0; padded_end[0] = (int)nnz;
    } else {
        f->owns_arrays = true;
        BB_CUDA(cudaMalloc((void**)&f->ptr, (size_t)(V + 1) * sizeof(int)));
        int *seg_of = nullptr, *key = nullptr, *key_sorted = nullptr, *iota = nullptr, *perm = nullptr;
        i64* vkey = nullptr;
        size_t nb = (size_t)(nnz > 0 ? nnz : 1);
        BB_CUDA(cudaMalloc((void**)&seg_of, nb * sizeof(int)));
        BB_CUDA(cudaMalloc((void**)&key, nb * sizeof(int)));
        BB_CUDA(cudaMalloc((void**)&key_sorted, nb * sizeof(int)));
        BB_CUDA(cudaMalloc((void**)&iota, nb * sizeof(int)));
        BB_CUDA(cudaMalloc((void**)&perm, nb * sizeof(int)));
        BB_CUDA(cudaMalloc((void**)&vkey, nb * sizeof(i64)));
        int rc = BB_OK;
        int* d_delta = nullptr;           // per-slab shift that pads every slab start to a multiple of 4 nnz
        int* d_old = nullptr;
        BB_CUDA(cudaMalloc((void**)&d_delta, ((size_t)nslab + 1) * sizeof(int)));
        BB_CUDA(cudaMalloc((void**)&d_old, ((size_t)nslab + 1) * sizeof(int)));
        if (nnz > 0) {
            int g = (int)((nnz + TB - 1) / TB);
            k_expand_ptr<<<g, TB, 0, st>>>(cptr, n_seg, nnz, seg_of);
            k_slab_key<<<g, TB, 0, st>>>(cidx, nnz, (int)W, key);
            k_iota<<<g, TB, 0, st>>>(iota, nnz);
            ctx->launches += 3;
            rc = sort_pairs<int>(ctx, key, key_sorted, iota, perm, nnz, bits_for(nslab - 1));
        }
        std::vector<int> old_start((size_t)nslab + 1, 0), delta((size_t)nslab + 1, 0);
        if (rc == BB_OK) {
            // first position of every slab in the sorted sequence
            k_lower_bound<int><<<(int)((nslab + 1 + TB - 1) / TB), TB, 0, st>>>(key_sorted, nnz, nslab, d_old);
            ctx->launches += 1;
            BB_CUDA(cudaMemcpyAsync(old_start.data(), d_old, ((size_t)nslab + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
            BB_CUDA(cudaStreamSynchronize(st));
            i64 pos = 0;
            for (i64 sl = 0; sl < nslab; ++sl) {
                pos = (pos + 3) & ~(i64)3;
                padded_start[(size_t)sl] = (int)pos;
                delta[(size_t)sl] = (int)(pos - old_start[(size_t)sl]);
                pos += old_start[(size_t)sl + 1] - old_start[(size_t)sl];
                padded_end[(size_t)sl] = (int)pos;
            }
            delta[(size_t)nslab] = delta[(size_t)nslab - 1];
            padded_total = pos;
            if (padded_total + 8 >= ((i64)1 << 31)) { bb_set_error("slab format exceeds int32 nnz range"); rc = BB_ERR_ARG; }
        }
        if (rc == BB_OK) {
            BB_CUDA(cudaMemcpyAsync(d_delta, delta.data(), ((size_t)nslab + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
            BB_CUDA(cudaMalloc((void**)&f->idx, (size_t)(padded_total + 8) * sizeof(int)));
            BB_CUDA(cudaMemsetAsync(f->idx, 0, (size_t)(padded_total + 8) * sizeof(int), st));
            if (cval) {
                BB_CUDA(cudaMalloc((void**)&f->val, (size_t)(padded_total + 8) * sizeof(double)));
                BB_CUDA(cudaMemsetAsync(f->val, 0, (size_t)(padded_total + 8) * sizeof(double), st));
            }
            if (nnz > 0) {
                int g = (int)((nnz + TB - 1) / TB);
                k_scatter_padded<<<g, TB, 0, st>>>(cidx, cval, perm, key_sorted, d_delta, nnz, f->idx, f->val);
                k_vkey<<<g, TB, 0, st>>>(key_sorted, seg_of, perm, nnz, n_seg, vkey);
                ctx->launches += 2;
            }
            int g = (int)((V + 1 + TB - 1) / TB);
            k_lower_bound<i64><<<g, TB, 0, st>>>(vkey, nnz, V, f->ptr);
            k_shift_ptr<<<g, TB, 0, st>>>(f->ptr, V, n_seg, d_delta);
            ctx->launches += 2;
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { bb_set_error("slab build: %s", cudaGetErrorString(e)); rc = BB_ERR_CUDA; }
        }
        cudaFree(d_delta); cudaFree(d_old);
        cudaFree(seg_of); cudaFree(key); cudaFree(key_sorted); cudaFree(iota); cudaFree(perm); cudaFree(vkey);
        BB_TRY(rc);
    }
    if (sliced) {
        BB_CUDA(cudaMalloc((void**)&f->slab_nnz1, ((size_t)nslab + 1) * sizeof(int)));
        BB_CUDA(cudaMemcpyAsync(f->slab_nnz1, padded_end.data(), ((size_t)nslab + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        BB_TRY(bb_sell_build(ctx, f));
        // the slab-ordered copies were only the source of the sliced arrays
        if (f->owns_arrays) {
            cudaFree(f->ptr); cudaFree(f->idx); if (f->val) cudaFree(f->val);
        }
        f->ptr = nullptr; f->idx = nullptr;
        f->has_val = (f->val != nullptr);
        f->val = nullptr;
        f->owns_arrays = false;
        return BB_OK;
    }
    // slab nnz ranges -> tiles (host), ownership (device)
    std::vector<int>& slab_off = padded_start;
    std::vector<TileMeta> tiles;
    std::vector<int> slab_tile0((size_t)nslab + 1);
    for (i64 s = 0; s < nslab; ++s) {
        slab_tile0[(size_t)s] = (int)tiles.size();
        int a = padded_start[(size_t)s], b = padded_end[(size_t)s];
        int nt = (b - a + SPMV_TILE - 1) / SPMV_TILE;
        if (nt < 1) nt = 1;
        for (int t = 0; t < nt; ++t) {
            TileMeta m;
            m.start = a + t * SPMV_TILE;
            m.end = std::min(b, m.start + SPMV_TILE);
            if (m.end < m.start) m.end = m.start;
            m.vlo = (int)s;               // packed inputs for k_tile_meta
            m.vhi = (t == nt - 1) ? 1 : 0;
            tiles.push_back(m);
        }
    }
    slab_tile0[(size_t)nslab] = (int)tiles.size();
    f->ntiles = (int)tiles.size();
    BB_CUDA(cudaMalloc((void**)&f->tiles, tiles.size() * sizeof(TileMeta)));
    BB_CUDA(cudaMalloc((void**)&f->head_seg, tiles.size() * sizeof(int)));
    BB_CUDA(cudaMemcpyAsync(f->tiles, tiles.data(), tiles.size() * sizeof(TileMeta), cudaMemcpyHostToDevice, st));
    k_tile_meta<<<(f->ntiles + TB - 1) / TB, TB, 0, st>>>(f->ptr, V, n_seg, f->ntiles, f->tiles, f->head_seg);
    ctx->launches += 1;
    // compact per-tile metadata {vlo, nown} and the per-slab tile / nnz offsets read by the kernel
    BB_CUDA(cudaMalloc((void**)&f->tmeta, tiles.size() * sizeof(int2)));
    k_compact_meta<<<(f->ntiles + TB - 1) / TB, TB, 0, st>>>(f->tiles, f->ntiles, f->tmeta);
    ctx->launches += 1;
    if (f->owns_arrays && staged && ctx->opt_bank_permute != 0 && nnz > 0) {
        // reorder the nnz inside every (segment x tile) piece for conflict-free staged gathers
        static_assert(SPMV_TILE == 256 && SPMV_ITEMS == 8, "k_bank_permute assumes 32 lanes x 8 nnz per tile");
        int* idx_tmp = nullptr;
        double* val_tmp = nullptr;
        size_t cnt = (size_t)(padded_total + 8);
        BB_CUDA(cudaMalloc((void**)&idx_tmp, cnt * sizeof(int)));
        BB_CUDA(cudaMemcpyAsync(idx_tmp, f->idx, cnt * sizeof(int), cudaMemcpyDeviceToDevice, st));
        if (f->val) {
            cudaError_t e = cudaMalloc((void**)&val_tmp, cnt * sizeof(double));
            if (e != cudaSuccess) { cudaFree(idx_tmp); bb_set_error("slab build: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
            BB_CUDA(cudaMemcpyAsync(val_tmp, f->val, cnt * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        k_bank_permute<<<(f->ntiles + 63) / 64, 64, 0, st>>>(f->tiles, f->ntiles, f->ptr, V, idx_tmp, val_tmp, f->idx, f->val);
        ctx->launches += 1;
        cudaError_t e = cudaStreamSynchronize(st);
        cudaFree(idx_tmp);
        if (val_tmp) cudaFree(val_tmp);
        if (e != cudaSuccess) { bb_set_error("bank permute: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
    }
    BB_CUDA(cudaMalloc((void**)&f->slab_tile0, ((size_t)nslab + 1) * sizeof(int)));
    BB_CUDA(cudaMalloc((void**)&f->slab_nnz0, ((size_t)nslab + 1) * sizeof(int)));
    BB_CUDA(cudaMemcpyAsync(f->slab_tile0, slab_tile0.data(), ((size_t)nslab + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMalloc((void**)&f->slab_nnz1, ((size_t)nslab + 1) * sizeof(int)));
    BB_CUDA(cudaMemcpyAsync(f->slab_nnz0, slab_off.data(), ((size_t)nslab + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(f->slab_nnz1, padded_end.data(), ((size_t)nslab + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMalloc((void**)&f->part, (size_t)(V > 0 ? V : 1) * sizeof(double)));
    BB_CUDA(cudaMalloc((void**)&f->head_part, (size_t)f->ntiles * sizeof(double)));
    BB_CUDA(cudaMemsetAsync(f->part, 0, (size_t)(V > 0 ? V : 1) * sizeof(double), st));
    BB_CUDA(cudaMemsetAsync(f->head_part, 0, (size_t)f->ntiles * sizeof(double), st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BB_OK;
}

// ---- sub-warp-per-segment variant (spmv_variant 2) -----------------------------------------------------------------
// For matrices whose index arrays and gather vector live in the 126 MB L2 (BASELINE configs 1 and 3) the persistent
// staged kernels pay more for staging a window per CTA than for the product itself.  Here TPR lanes (2 ... 32, chosen
// from the mean segment length) walk one row (dot, CSR) or column (Tdot, CSC) of the CANONICAL image, gathering through
// L1/L2; the lanes' partial sums are added in a fixed shuffle tree.  One slab, no staging, no build cost.
template <int TPR, bool BINARY>
__global__ void __launch_bounds__(256)
k_csr_rowwise(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val, i64 n_seg,
              const double* __restrict__ gvec, double* __restrict__ part, const int* __restrict__ done_flag) {
    pdl_trigger();
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 seg = t / TPR;
    const int sub = (int)(t % TPR);
    int k0 = 0, k1 = 0;
    if (seg < n_seg) { k0 = ptr[seg]; k1 = ptr[seg + 1]; }      // matrix format: readable before the previous kernel is done
    pdl_wait();
    if (done_flag != nullptr && *done_flag) return;
    double acc = 0.0;
    if (seg < n_seg) {
        for (int k = k0 + sub; k < k1; k += TPR) {
            const double g = __ldg(gvec + idx[k]);
            acc += BINARY ? g : val[k] * g;
        }
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, TPR);
    if (sub == 0 && seg < n_seg) part[seg] = acc;
}

template <bool BINARY>
static cudaError_t launch_rowwise(bb_ctx* ctx, const SlabFmt* f, const double* gvec, const int* done_flag) {
    cudaError_t e = cudaSuccess;
    const double mean_len = f->n_seg > 0 ? (double)f->nnz / (double)f->n_seg : 0.0;
    int tpr = 2;
    while (tpr < 32 && tpr * 4 < mean_len) tpr *= 2;            // ~4-8 nnz per lane
    const i64 threads = f->n_seg * tpr;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    switch (tpr) {
    case 2: e = bb_launch(ctx, true, k_csr_rowwise<2, BINARY>, dim3(grid), dim3(256), 0, f->ptr, f->idx, f->val, f->n_seg, gvec, f->part, done_flag); break;
    case 4: e = bb_launch(ctx, true, k_csr_rowwise<4, BINARY>, dim3(grid), dim3(256), 0, f->ptr, f->idx, f->val, f->n_seg, gvec, f->part, done_flag); break;
    case 8: e = bb_launch(ctx, true, k_csr_rowwise<8, BINARY>, dim3(grid), dim3(256), 0, f->ptr, f->idx, f->val, f->n_seg, gvec, f->part, done_flag); break;
    case 16: e = bb_launch(ctx, true, k_csr_rowwise<16, BINARY>, dim3(grid), dim3(256), 0, f->ptr, f->idx, f->val, f->n_seg, gvec, f->part, done_flag); break;
    default: e = bb_launch(ctx, true, k_csr_rowwise<32, BINARY>, dim3(grid), dim3(256), 0, f->ptr, f->idx, f->val, f->n_seg, gvec, f->part, done_flag); break;
    }
    return e;
}

// launch the SpMV + fix-up for one format; gvec has f->n_gather entries
int bb_launch_spmv(bb_mat* m, SlabFmt* f, const double* gvec, const int* done_flag, bool skip_overflow_add) {
    bb_ctx* ctx = m->ctx;
    if (f->variant == 1) return bb_sell_launch(ctx, f, gvec, done_flag, skip_overflow_add);
    if (f->variant == 2) {
        if (f->n_seg == 0) return BB_OK;
        if (f->val == nullptr) BB_CUDA(launch_rowwise<true>(ctx, f, gvec, done_flag));
        else BB_CUDA(launch_rowwise<false>(ctx, f, gvec, done_flag));
        BB_LAUNCHED(ctx);
        return BB_OK;
    }
    // a format built without staging in mind may have slabs wider than shared memory
    const bool stage = f->staged && (i64)f->W <= max_stage_width(ctx);
    int wstage = stage ? f->W : 0;
    size_t smem = (size_t)(wstage + SPMV_WARPS * (SPMV_TILE + 32)) * sizeof(double) + SPMV_WARPS * 8 * sizeof(unsigned);
    static BBDeviceOnce attr_set = {{0, 0, 0, 0}};
    if (attr_set.first(ctx->device)) {
        int mx = (int)ctx->smem_optin;
        BB_CUDA(cudaFuncSetAttribute(k_seg_spmv<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        BB_CUDA(cudaFuncSetAttribute(k_seg_spmv<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        BB_CUDA(cudaFuncSetAttribute(k_seg_spmv<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        BB_CUDA(cudaFuncSetAttribute(k_seg_spmv<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    }
    if (smem > ctx->smem_optin) { bb_set_error("spmv shared memory %zu exceeds %zu", smem, ctx->smem_optin); return BB_ERR_ARG; }
    const bool binary = (f->val == nullptr);
    int ncta = ctx->sm_count < f->ntiles ? ctx->sm_count : f->ntiles;
    dim3 grid(ncta), block(SPMV_THREADS);
#define SPMV_ARGS f->ptr, f->idx, f->val, f->tmeta, f->slab_tile0, f->slab_nnz0, f->slab_nnz1, f->nslab, f->ntiles, gvec, f->W, f->n_gather, wstage, f->part, f->head_part, done_flag
    if (binary && stage) k_seg_spmv<true, true><<<grid, block, smem, ctx->stream>>>(SPMV_ARGS);
    else if (binary && !stage) k_seg_spmv<true, false><<<grid, block, smem, ctx->stream>>>(SPMV_ARGS);
    else if (!binary && stage) k_seg_spmv<false, true><<<grid, block, smem, ctx->stream>>>(SPMV_ARGS);
    else k_seg_spmv<false, false><<<grid, block, smem, ctx->stream>>>(SPMV_ARGS);
#undef SPMV_ARGS
    BB_LAUNCHED(ctx);
    k_fixup<<<(f->ntiles + 255) / 256, 256, 0, ctx->stream>>>(f->head_seg, f->ntiles, f->head_part, f->part, done_flag);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

// ---- op pipeline (sparse part; dense twin in bb_dense.cu) ----------------------------------
int bb_dense_dot(bb_mat* m, int mode, const int* done_flag);
int bb_dense_tdot(bb_mat* m, const double* w, const int* done_flag);

static int P_grid(i64 P) { return grid_for(P, 1024, RED_MAX); }   // 256 threads, ~4 elements each
static int N_grid(i64 n) { return grid_for(n, 1024, RED_MAX); }

int bb_op_prepare_flag(bb_mat* m, const double* vP, const double* scale, const int* done_flag) {
    bb_ctx* ctx = m->ctx;
    k_prepare<<<P_grid(m->P), 256, 0, ctx->stream>>>(vP, scale, m->P, m->add_intercept, m->col_offset, m->sv,
                                                     m->red + RED_SHIFT * RED_MAX, done_flag);
    BB_LAUNCHED(ctx);
    return BB_OK;
}
int bb_op_prepare(bb_mat* m, const double* vP, const double* scale) { return bb_op_prepare_flag(m, vP, scale, nullptr); }

// the small kernels between the SpMV launches ask for the same shared-memory carve-out as the SpMV (opt_uniform_carveout)
static int iteration_kernel_attrs(bb_ctx* ctx) {
    static BBDeviceOnce once = {{0, 0, 0, 0}};
    if (!once.first(ctx->device)) return BB_OK;
    BB_CUDA(bb_prefer_max_smem(ctx, k_dot_finish<0>));
    BB_CUDA(bb_prefer_max_smem(ctx, k_dot_finish<1>));
    BB_CUDA(bb_prefer_max_smem(ctx, k_tdot_collect));
    BB_CUDA(bb_prefer_max_smem(ctx, k_prepare));
    return BB_OK;
}

int bb_op_dot_flag(bb_mat* m, int mode, const int* done_flag) {
    bb_ctx* ctx = m->ctx;
    if (!m->is_sparse) return bb_dense_dot(m, mode, done_flag);
    BB_TRY(iteration_kernel_attrs(ctx));
    BB_TRY(bb_launch_spmv(m, &m->fdot, m->sv + m->add_intercept, done_flag));
    const double* red_shift = m->red + RED_SHIFT * RED_MAX;
    int nshift = P_grid(m->P);
    if (mode == 0) {
        BB_CUDA(bb_launch(ctx, true, k_dot_finish<0>, dim3(N_grid(m->n)), dim3(256), 0, m->fdot.part, m->fdot.nslab, m->n,
                          red_shift, nshift, nullptr, nullptr, m->u_n, nullptr, nullptr, done_flag));
    } else {
        BB_CUDA(bb_launch(ctx, true, k_dot_finish<1>, dim3(N_grid(m->n)), dim3(256), 0, m->fdot.part, m->fdot.nslab, m->n,
                          red_shift, nshift, m->use_omega_scalar ? nullptr : m->omega, m->omega_scalar_dev, m->w_n, nullptr,
                          m->red + RED_W * RED_MAX, done_flag));
        m->nred_w = N_grid(m->n);
    }
    BB_LAUNCHED(ctx);
    return BB_OK;
}
int bb_op_dot(bb_mat* m, int mode) { return bb_op_dot_flag(m, mode, nullptr); }

// traw = [sum w; X' w], allreduced.  have_w_partials: red[RED_W] already holds N_grid(n) partial sums of w
int bb_op_tdot_flag(bb_mat* m, const double* w, bool have_w_partials, const int* done_flag,
                    bool fuse_reduce_into_consumer) {
    bb_ctx* ctx = m->ctx;
    if (!have_w_partials) {
        k_sum_partials<<<N_grid(m->n), 256, 0, ctx->stream>>>(w, m->n, m->red + RED_W * RED_MAX);
        BB_LAUNCHED(ctx);
        m->nred_w = N_grid(m->n);
    }
    // exchange: peer-memory one-shot (publication fused into the collect kernel) or NCCL
    P2PView view;
    const bool p2p = (ctx->nranks > 1) && bb_p2p_view(ctx, m->p + 1, &view);
    const P2PView* pub = nullptr;
    if (p2p) {
        if (m->p2p_view_valid != 1 + view.variant) {      // first use is never inside a graph capture (the RHS product precedes the loop)
            BB_CUDA(cudaStreamSynchronize(ctx->stream));
            BB_CUDA(cudaMemcpy(m->p2p_view_dev, &view, sizeof(P2PView), cudaMemcpyHostToDevice));
            m->p2p_view_valid = 1 + view.variant;
        }
        pub = m->p2p_view_dev;
    }
    if (!m->is_sparse) {
        BB_TRY(bb_dense_tdot(m, w, done_flag));
        k_tdot_collect<<<grid_for(m->p, 32, 4096), 256, 0, ctx->stream>>>(
            m->dense_part, m->dense_nblk, m->p, m->red + RED_W * RED_MAX, m->nred_w, m->traw, done_flag, pub);
        BB_LAUNCHED(ctx);
    } else {
        BB_TRY(bb_launch_spmv(m, &m->ftdot, w, done_flag));
        k_tdot_collect<<<grid_for(m->p, 32, 4096), 256, 0, ctx->stream>>>(
            m->ftdot.part, m->ftdot.nslab, m->p, m->red + RED_W * RED_MAX, m->nred_w, m->traw, done_flag, pub);
        BB_LAUNCHED(ctx);
    }
    if (p2p) {
        if (!fuse_reduce_into_consumer) BB_TRY(bb_p2p_reduce_into(ctx, m->traw, m->p + 1));
    } else {
        BB_TRY(bb_allreduce_dev(ctx, m->traw, m->p + 1));
    }
    return BB_OK;
}
int bb_op_tdot(bb_mat* m, const double* w) { return bb_op_tdot_flag(m, w, false, nullptr, false); }

// the local part only: traw = [sum w; X' w] of this shard (w partials already in red[RED_W]); the fused P-side kernel
// exchanges it
int bb_op_tdot_local(bb_mat* m, const double* w, const int* done_flag) {
    if (!m->is_sparse) BB_TRY(bb_dense_tdot(m, w, done_flag));
    else BB_TRY(bb_launch_spmv(m, &m->ftdot, w, done_flag));
    return bb_op_collect_local(m, done_flag);
}
// traw = [sum of the w partials; column sums of the slab / row-block partials]
int bb_op_collect_local(bb_mat* m, const int* done_flag) {
    bb_ctx* ctx = m->ctx;
    BB_CUDA(bb_launch(ctx, true, k_tdot_collect, dim3(grid_for(m->p, 32, 4096)), dim3(256), 0,
                      m->is_sparse ? m->ftdot.part : m->dense_part, m->is_sparse ? m->ftdot.nslab : m->dense_nblk, m->p,
                      m->red + RED_W * RED_MAX, m->nred_w, m->traw, done_flag, nullptr));
    BB_LAUNCHED(ctx);
    return BB_OK;
}

int bb_op_tdot_finish(bb_mat* m, double* tP) {
    bb_ctx* ctx = m->ctx;
    k_tdot_finish<<<P_grid(m->P), 256, 0, ctx->stream>>>(m->traw, m->p, m->add_intercept, m->col_offset, tP);
    BB_LAUNCHED(ctx);
    return BB_OK;
}

// ---- matrix lifetime -------------------------------------------------------------------------
// NB: ctx->stream is non-blocking, i.e. NOT ordered with the legacy default stream: every memset
// must go to ctx->stream or it races with the async copies issued right after the allocation.
static int alloc_d(cudaStream_t st, double** p, i64 n) {
    BB_CUDA(cudaMalloc((void**)p, (size_t)(n > 0 ? n : 1) * sizeof(double)));
    BB_CUDA(cudaMemsetAsync(*p, 0, (size_t)(n > 0 ? n : 1) * sizeof(double), st));
    return BB_OK;
}

int bb_mat_alloc_work(bb_mat* m) {
    // n-vectors carry two spare doubles: the bulk window copy of the sliced SpMV rounds an odd window up to 16 bytes
    const i64 npad = m->n + 2;
    BB_TRY(alloc_d(m->ctx->stream, &m->omega, npad)); BB_TRY(alloc_d(m->ctx->stream, &m->n_trial, npad)); BB_TRY(alloc_d(m->ctx->stream, &m->n_success, npad));
    BB_TRY(alloc_d(m->ctx->stream, &m->eta, npad)); BB_TRY(alloc_d(m->ctx->stream, &m->w_n, npad)); BB_TRY(alloc_d(m->ctx->stream, &m->u_n, npad));
    BB_TRY(alloc_d(m->ctx->stream, &m->eps_n, npad));
    // sv + add_intercept is the vector `dot` gathers from: keep THAT address 16-byte aligned
    BB_TRY(alloc_d(m->ctx->stream, &m->sv_base, m->P + 4));
    m->sv = m->sv_base + (m->add_intercept ? 1 : 0);
    double** pv[] = {&m->v_P, &m->t_P, &m->x, &m->r, &m->pvec, &m->q, &m->b, &m->s, &m->D, &m->pps,
                     &m->z, &m->x0, &m->eps_P, &m->out_P};
    for (auto pp : pv) BB_TRY(alloc_d(m->ctx->stream, pp, m->P + 1));
    BB_TRY(alloc_d(m->ctx->stream, &m->traw, m->p + 1));
    BB_TRY(alloc_d(m->ctx->stream, &m->zk, m->P + 1));
    BB_TRY(alloc_d(m->ctx->stream, &m->omega_scalar_dev, 1));
    BB_CUDA(cudaMalloc((void**)&m->p2p_view_dev, sizeof(P2PView)));
    BB_TRY(alloc_d(m->ctx->stream, &m->red, (i64)RED_SLOTS * RED_MAX));
    BB_CUDA(cudaMalloc((void**)&m->cg, sizeof(CgScalars)));
    BB_CUDA(cudaMemsetAsync(m->cg, 0, sizeof(CgScalars), m->ctx->stream));
    BB_CUDA(cudaMallocHost((void**)&m->cg_host, sizeof(CgScalars)));
    BB_CUDA(cudaMalloc((void**)&m->ps_bar, sizeof(unsigned long long)));
    BB_CUDA(cudaMemsetAsync(m->ps_bar, 0, sizeof(unsigned long long), m->ctx->stream));
    m->omega_scalar = 1.0;
    m->use_omega_scalar = 0;
    m->last_n_iter = 0;
    return BB_OK;
}

extern "C" int bb_batch_free(bb_mat* m);

extern "C" int bb_mat_free(bb_mat* m) {
    if (!m) return BB_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    bb_batch_free(m);
    if (m->cg_graph) cudaGraphExecDestroy(m->cg_graph);
    bb_slab_free(&m->fdot);
    bb_slab_free(&m->ftdot);
    void* ptrs[] = {m->csr_ptr, m->csr_idx, m->csr_val, m->csc_ptr, m->csc_idx, m->csc_val, m->col_offset, m->Xd,
                    m->omega, m->n_trial, m->n_success, m->eta, m->w_n, m->u_n, m->eps_n, m->dense_part, m->zk, m->omega_scalar_dev, m->p2p_view_dev,
                    m->st_lscale, m->st_mean, m->st_square, m->st_prior_sd, m->st_sums,
                    m->v_P, m->sv_base, m->traw, m->t_P, m->x, m->r, m->pvec, m->q, m->b, m->s, m->D, m->pps, m->z, m->x0,
                    m->eps_P, m->out_P, m->red, m->cg, m->ps_bar};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (m->cg_host) cudaFreeHost(m->cg_host);
    free(m);
    return BB_OK;
}

extern "C" int bb_csr_upload(bb_ctx* ctx, int64_t n, int64_t p, int64_t nnz,
                             const int32_t* indptr, const int32_t* indices, const double* data,
                             const double* column_offset, int add_intercept,
                             int64_t row_offset, int64_t n_global, bb_mat** out) {
    BB_ARG(ctx && out && indptr, "ctx/out/indptr");
    BB_ARG(n >= 0 && p >= 0 && nnz >= 0, "negative size");
    BB_ARG(nnz == 0 || indices != nullptr, "indices");
    BB_ARG(n < ((i64)1 << 31) - 1 && p < ((i64)1 << 31) - 1 && nnz < ((i64)1 << 31) - 1, "int32 index range exceeded");
    BB_ARG(indptr[0] == 0 && indptr[n] == nnz, "indptr[0] != 0 or indptr[n] != nnz");
    BB_CUDA(cudaSetDevice(ctx->device));
    bb_mat* m = (bb_mat*)calloc(1, sizeof(bb_mat));
    m->ctx = ctx;
    m->is_sparse = 1;
    m->is_binary = (data == nullptr);
    m->add_intercept = add_intercept ? 1 : 0;
    m->centered = (column_offset != nullptr);
    m->n = n; m->p = p; m->P = p + m->add_intercept; m->nnz = nnz;
    m->row_offset = row_offset; m->n_global = n_global > 0 ? n_global : n;
    cudaStream_t st = ctx->stream;
    size_t nb = (size_t)(nnz > 0 ? nnz : 1);
    int rc = BB_OK;
    do {
#define CK(e) if ((rc = (e)) != BB_OK) break
#define CKC(e) { cudaError_t e_ = (e); if (e_ != cudaSuccess) { bb_set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); rc = BB_ERR_CUDA; break; } }
        CKC(cudaMalloc((void**)&m->csr_ptr, (size_t)(n + 1) * sizeof(int)));
        CKC(cudaMalloc((void**)&m->csr_idx, nb * sizeof(int)));
        CKC(cudaMemcpyAsync(m->csr_ptr, indptr, (size_t)(n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        if (nnz > 0) CKC(cudaMemcpyAsync(m->csr_idx, indices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, st));
        if (data) {
            CKC(cudaMalloc((void**)&m->csr_val, nb * sizeof(double)));
            if (nnz > 0) CKC(cudaMemcpyAsync(m->csr_val, data, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        CK(alloc_d(st, &m->col_offset, p));
        if (column_offset && p > 0)
            CKC(cudaMemcpyAsync(m->col_offset, column_offset, (size_t)p * sizeof(double), cudaMemcpyHostToDevice, st));
        // canonical CSC: stable sort of the CSR nnz sequence by column
        CKC(cudaMalloc((void**)&m->csc_ptr, (size_t)(p + 1) * sizeof(int)));
        CKC(cudaMalloc((void**)&m->csc_idx, nb * sizeof(int)));
        if (data) CKC(cudaMalloc((void**)&m->csc_val, nb * sizeof(double)));
        int *row_of = nullptr, *iota = nullptr, *perm = nullptr, *col_sorted = nullptr;
        CKC(cudaMalloc((void**)&row_of, nb * sizeof(int)));
        CKC(cudaMalloc((void**)&iota, nb * sizeof(int)));
        CKC(cudaMalloc((void**)&perm, nb * sizeof(int)));
        CKC(cudaMalloc((void**)&col_sorted, nb * sizeof(int)));
        const int TB = 256;
        if (nnz > 0) {
            int g = (int)((nnz + TB - 1) / TB);
            k_expand_ptr<<<g, TB, 0, st>>>(m->csr_ptr, n, nnz, row_of);
            k_iota<<<g, TB, 0, st>>>(iota, nnz);
            ctx->launches += 2;
            rc = sort_pairs<int>(ctx, m->csr_idx, col_sorted, iota, perm, nnz, bits_for(p > 0 ? p - 1 : 0));
            if (rc == BB_OK) {
                k_gather_i<<<g, TB, 0, st>>>(row_of, perm, nnz, m->csc_idx);
                if (data) k_gather_d<<<g, TB, 0, st>>>(m->csr_val, perm, nnz, m->csc_val);
                ctx->launches += 2;
            }
        }
        if (rc == BB_OK) {
            k_lower_bound<int><<<(int)((p + 1 + TB - 1) / TB), TB, 0, st>>>(col_sorted, nnz, p, m->csc_ptr);
            ctx->launches += 1;
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { bb_set_error("csc build: %s", cudaGetErrorString(e)); rc = BB_ERR_CUDA; }
        }
        cudaFree(row_of); cudaFree(iota); cudaFree(perm); cudaFree(col_sorted);
        if (rc != BB_OK) break;
        // spmv_stage: 0 = gather through L2 (one slab), 1 = stage both products' vectors in shared memory,
        //             2 = stage only dot's (p-vector), 3 = stage only Tdot's (n-vector)
        const i64 sopt = ctx->opt_spmv_stage;
        CK(build_slab_format(ctx, m->csr_ptr, m->csr_idx, m->csr_val, n, p, nnz, sopt == 1 || sopt == 2, &m->fdot));
        CK(build_slab_format(ctx, m->csc_ptr, m->csc_idx, m->csc_val, p, n, nnz, sopt == 1 || sopt == 3, &m->ftdot));
        CK(bb_mat_alloc_work(m));
        CKC(cudaStreamSynchronize(st));
#undef CK
#undef CKC
    } while (0);
    if (rc != BB_OK) { bb_mat_free(m); return rc; }
    *out = m;
    return BB_OK;
}

extern "C" int bb_mat_info(bb_mat* m, int64_t* n_local, int64_t* P, int64_t* nnz, int* is_sparse, int* is_binary) {
    BB_ARG(m != nullptr, "mat");
    if (n_local) *n_local = m->n;
    if (P) *P = m->P;
    if (nnz) *nnz = m->nnz;
    if (is_sparse) *is_sparse = m->is_sparse;
    if (is_binary) *is_binary = m->is_binary;
    return BB_OK;
}

static int export_compressed(bb_mat* m, const int* dptr, i64 nptr, const int* didx, const double* dval,
                             int32_t* indptr, int32_t* indices, double* data) {
    BB_ARG(m->is_sparse, "not a sparse matrix");
    cudaStream_t st = m->ctx->stream;
    BB_CUDA(cudaSetDevice(m->ctx->device));
    if (indptr) BB_CUDA(cudaMemcpyAsync(indptr, dptr, (size_t)nptr * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (indices && m->nnz > 0) BB_CUDA(cudaMemcpyAsync(indices, didx, (size_t)m->nnz * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (data && m->nnz > 0) {
        if (dval) BB_CUDA(cudaMemcpyAsync(data, dval, (size_t)m->nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
        else for (i64 k = 0; k < m->nnz; ++k) data[k] = 1.0;
    }
    BB_CUDA(cudaStreamSynchronize(st));
    return BB_OK;
}

extern "C" int bb_mat_export_csr(bb_mat* m, int32_t* indptr, int32_t* indices, double* data) {
    BB_ARG(m != nullptr, "mat");
    return export_compressed(m, m->csr_ptr, m->n + 1, m->csr_idx, m->csr_val, indptr, indices, data);
}
extern "C" int bb_mat_export_csc(bb_mat* m, int32_t* indptr, int32_t* indices, double* data) {
    BB_ARG(m != nullptr, "mat");
    return export_compressed(m, m->csc_ptr, m->p + 1, m->csc_idx, m->csc_val, indptr, indices, data);
}

// ---- host-buffer products (seam 1) -------------------------------------------------------------
extern "C" int bb_dot(bb_mat* m, const double* v, double* out) {
    BB_ARG(m && v && out, "mat/v/out");
    bb_ctx* ctx = m->ctx;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    BB_CUDA(cudaMemcpyAsync(m->v_P, v, (size_t)m->P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    BB_TRY(bb_op_prepare(m, m->v_P, nullptr));
    BB_TRY(bb_op_dot(m, 0));
    BB_CUDA(cudaMemcpyAsync(out, m->u_n, (size_t)m->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    timer_.commit();
    return BB_OK;
}

extern "C" int bb_tdot(bb_mat* m, const double* w, double* out) {
    BB_ARG(m && w && out, "mat/w/out");
    bb_ctx* ctx = m->ctx;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    BB_CUDA(cudaMemcpyAsync(m->eps_n, w, (size_t)m->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    BB_TRY(bb_op_tdot(m, m->eps_n));
    BB_TRY(bb_op_tdot_finish(m, m->t_P));
    BB_CUDA(cudaMemcpyAsync(out, m->t_P, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    timer_.commit();
    return BB_OK;
}

// diag(X' W X) with the intercept / centring algebra of sparse_matrix.py:164-177:
//   d_j = sum_i w_i x_ij^2 - 2 c_j sum_i w_i x_ij + (sum w) c_j^2 ; intercept entry = sum w
__global__ void k_fisher_finish(const double* __restrict__ d2, const double* __restrict__ d1, const double* __restrict__ sw_p,
                                const double* __restrict__ c, i64 p, int icpt, int centered, double* __restrict__ out) {
    const double sw = sw_p[0];
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < p + icpt; j += (i64)gridDim.x * blockDim.x) {
        if (j < icpt) { out[j] = sw; continue; }
        i64 k = j - icpt;
        double d = d2[1 + k];
        if (centered) { d -= 2.0 * c[k] * d1[1 + k]; d += sw * c[k] * c[k]; }
        out[j] = d;
    }
}

int bb_dense_fisher_diag(bb_mat* m, const double* weight_dev, double* d2, double* d1);

extern "C" int bb_fisher_diag(bb_mat* m, const double* weight, double* out) {
    BB_ARG(m && weight && out, "mat/weight/out");
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    BB_CUDA(cudaMemcpyAsync(m->eps_n, weight, (size_t)m->n * sizeof(double), cudaMemcpyHostToDevice, st));
    // d2 -> q[1..p], d1 -> b[1..p], sum w -> traw[0]; all three allreduced
    double *d2 = m->q, *d1 = m->b;
    if (m->is_sparse) {
        if (m->p > 0) {
            i64 threads = m->p * 32;
            k_fisher_diag_csc<<<(int)((threads + 255) / 256), 256, 0, st>>>(m->csc_ptr, m->csc_idx, m->csc_val, m->p,
                                                                            m->eps_n, d2 + 1, d1 + 1);
            BB_LAUNCHED(ctx);
        }
    } else {
        BB_TRY(bb_dense_fisher_diag(m, m->eps_n, d2 + 1, d1 + 1));
    }
    k_sum_partials<<<N_grid(m->n), 256, 0, st>>>(m->eps_n, m->n, m->red + RED_MISC * RED_MAX);
    BB_LAUNCHED(ctx);
    k_tdot_collect<<<1, 256, 0, st>>>(nullptr, 0, 0, m->red + RED_MISC * RED_MAX, N_grid(m->n), m->traw, nullptr, nullptr);
    BB_LAUNCHED(ctx);
    BB_CUDA(cudaMemcpyAsync(d2, m->traw, sizeof(double), cudaMemcpyDeviceToDevice, st));
    BB_TRY(bb_allreduce_dev(ctx, d2, m->p + 1));
    BB_TRY(bb_allreduce_dev(ctx, d1 + 1, m->p));
    k_fisher_finish<<<P_grid(m->P), 256, 0, st>>>(d2, d1, d2, m->col_offset, m->p, m->add_intercept, m->centered, m->out_P);
    BB_LAUNCHED(ctx);
    BB_CUDA(cudaMemcpyAsync(out, m->out_P, (size_t)m->P * sizeof(double), cudaMemcpyDeviceToHost, st));
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    return BB_OK;
}


// ---- column moments from the resident CSC image (construction: abstract_matrix.py:93-107 remove_intercept_indicator) ----
__global__ void k_fill_ones(double* __restrict__ a, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) a[i] = 1.0;
}

// sum_out[j] = sum_i x_ij, sumsq_out[j] = sum_i x_ij^2 over the LOCAL rows (the caller adds the shards)
extern "C" int bb_column_moments(bb_mat* m, double* sum_out, double* sumsq_out) {
    BB_ARG(m && sum_out && sumsq_out, "mat/sum/sumsq");
    BB_ARG(m->is_sparse, "bb_column_moments: sparse designs only");
    bb_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BB_CUDA(cudaSetDevice(ctx->device));
    BBTimer timer_(ctx);
    if (m->p > 0) {
        k_fill_ones<<<N_grid(m->n), 256, 0, st>>>(m->eps_n, m->n);
        BB_LAUNCHED(ctx);
        const i64 threads = m->p * 32;
        k_fisher_diag_csc<<<(int)((threads + 255) / 256), 256, 0, st>>>(m->csc_ptr, m->csc_idx, m->csc_val, m->p, m->eps_n, m->q, m->b);
        BB_LAUNCHED(ctx);
        BB_CUDA(cudaMemcpyAsync(sumsq_out, m->q, (size_t)m->p * sizeof(double), cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaMemcpyAsync(sum_out, m->b, (size_t)m->p * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    timer_.end();
    BB_CUDA(cudaStreamSynchronize(st));
    timer_.commit();
    return BB_OK;
}

// (re)sets the centring offsets of a resident design: offset == NULL -> not centred
extern "C" int bb_set_column_offset(bb_mat* m, const double* offset) {
    BB_ARG(m != nullptr, "mat");
    bb_ctx* ctx = m->ctx;
    BB_CUDA(cudaSetDevice(ctx->device));
    if (offset && m->p > 0) BB_CUDA(cudaMemcpyAsync(m->col_offset, offset, (size_t)m->p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    else BB_CUDA(cudaMemsetAsync(m->col_offset, 0, (size_t)(m->p > 0 ? m->p : 1) * sizeof(double), ctx->stream));
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    m->centered = offset ? 1 : 0;
    m->zk_valid = 0;
    return BB_OK;
}
