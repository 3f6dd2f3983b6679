// Device-side construction of the v7 tile metadata from the arrays the current slab format already has
// (ptr[V+1] = padded nnz offsets of the virtual segments, per-slab nnz ranges and first tiles).
// Every step is a __host__ __device__ function of one index so that the same code runs in a CUDA kernel and, for
// validation without a GPU, in a host loop (spmv_v7.cu --host cross-checks it against the straightforward host builder).
//
//   step 1  v7_nonempty(v)            -> flag[v]                      (then cidx = exclusive scan of flag, CUB)
//   step 2  v7_head_pos(v)            -> hpos[cidx[v]] = ptr[v]       for non-empty v
//   step 3  v7_segment_slot(v)        -> cslot[v]                     slot of segment v in the compact output, -1 if empty
//   step 4  v7_tile_meta(t)           -> tile_out[t] = {first slot, n_heads}, chead_slot[t]
//   step 5  v7_lane_meta(t, lane)     -> lane_meta[32 t + lane] = flags | heads_in_lower_lanes << 16
#pragma once
#include <cuda_runtime.h>

#ifndef V7_HD
#define V7_HD __host__ __device__ __forceinline__
#endif

struct V7Geometry {
    const int* ptr;          // [V + 1]
    long long V, n_seg;
    int nslab, ntiles, tile;           // tile = nnz per tile (512)
    const int* slab_tile0;   // [nslab + 1]
    const int* slab_nnz0;    // [nslab + 1]
    const int* slab_nnz1;    // [nslab + 1]
};

V7_HD int v7_slab_of_segment(const V7Geometry& g, long long v) { return g.n_seg > 0 ? (int)(v / g.n_seg) : 0; }

V7_HD int v7_segment_end(const V7Geometry& g, long long v) {
    const int s = v7_slab_of_segment(g, v);
    const int e = g.ptr[v + 1], cap = g.slab_nnz1[s];
    return e < cap ? e : cap;                      // the last segment of a slab stops before the padding
}

V7_HD int v7_nonempty(const V7Geometry& g, long long v) { return v7_segment_end(g, v) > g.ptr[v] ? 1 : 0; }

V7_HD int v7_tile_of(const V7Geometry& g, int slab, int pos) { return g.slab_tile0[slab] + (pos - g.slab_nnz0[slab]) / g.tile; }

// first v in [0, V] with ptr[v] >= pos
V7_HD long long v7_lower_bound(const V7Geometry& g, int pos) {
    long long lo = 0, hi = g.V;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (g.ptr[mid] < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}

V7_HD int v7_segment_slot(const V7Geometry& g, const int* cidx, long long v) {
    if (!v7_nonempty(g, v)) return -1;
    return cidx[v] + v7_tile_of(g, v7_slab_of_segment(g, v), g.ptr[v]) + 1;
}

V7_HD int v7_slab_of_tile(const V7Geometry& g, int t) {
    int lo = 0, hi = g.nslab;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (g.slab_tile0[mid] <= t) lo = mid; else hi = mid; }
    return lo;
}

V7_HD void v7_tile_range(const V7Geometry& g, int t, int* slab, int* start, int* end) {
    const int s = v7_slab_of_tile(g, t);
    const int a = g.slab_nnz0[s] + (t - g.slab_tile0[s]) * g.tile;
    int b = a + g.tile;
    if (b > g.slab_nnz1[s]) b = g.slab_nnz1[s];
    if (b < a) b = a;
    *slab = s; *start = a; *end = b;
}

// cidx has V + 1 entries (cidx[V] = number of non-empty segments); hpos[c] = ptr of the c-th non-empty segment
V7_HD void v7_tile_meta(const V7Geometry& g, const int* cidx, const int* hpos, int t, int2* tile_out, int* chead_slot) {
    int s, start, end;
    v7_tile_range(g, t, &s, &start, &end);
    const long long va = v7_lower_bound(g, start), vb = v7_lower_bound(g, end);
    const int ca = cidx[va], cb = end > start ? cidx[vb] : cidx[va];
    *tile_out = make_int2(ca + t, cb - ca);
    int ch = -1;
    if (end > start) {
        const bool head_at_start = (cb > ca) && hpos[ca] == start;
        if (!head_at_start && ca > 0) {
            const int hp = hpos[ca - 1];
            if (hp >= g.slab_nnz0[s]) ch = (ca - 1) + v7_tile_of(g, s, hp) + 1;       // same slab by construction
        }
    }
    *chead_slot = ch;
}

V7_HD unsigned v7_lane_meta(const V7Geometry& g, const int* cidx, int t, int lane, int items) {
    int s, start, end;
    v7_tile_range(g, t, &s, &start, &end);
    if (end <= start) return 0u;
    int a = start + lane * items, b = a + items;
    if (a > end) a = end;
    if (b > end) b = end;
    const long long v0 = v7_lower_bound(g, start), va = v7_lower_bound(g, a);
    unsigned flags = 0u;
    for (long long v = va; v < g.V && g.ptr[v] < b; ++v)
        if (v7_nonempty(g, v)) flags |= 1u << (g.ptr[v] - (start + lane * items));
    return flags | ((unsigned)(cidx[va] - cidx[v0]) << 16);
}

#ifdef __CUDACC__
__global__ void k_v7_nonempty(V7Geometry g, int* flag) {
    long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < g.V) flag[v] = v7_nonempty(g, v);
    else if (v == g.V) flag[v] = 0;
}
__global__ void k_v7_segments(V7Geometry g, const int* cidx, int* hpos, int* cslot) {
    long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.V) return;
    const int slot = v7_segment_slot(g, cidx, v);
    cslot[v] = slot;
    if (slot >= 0) hpos[cidx[v]] = g.ptr[v];
}
__global__ void k_v7_tiles(V7Geometry g, const int* cidx, const int* hpos, int2* tile_out, int* chead_slot) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < g.ntiles) v7_tile_meta(g, cidx, hpos, t, tile_out + t, chead_slot + t);
}
__global__ void k_v7_lanes(V7Geometry g, const int* cidx, unsigned* lane_meta, int items) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)g.ntiles * 32) lane_meta[i] = v7_lane_meta(g, cidx, (int)(i >> 5), (int)(i & 31), items);
}
#endif
