// Device code of the v7 SpMV prototype (see README.md and emulate_spmv_v7.py in this directory): tile bodies, the
// kernel and the fix-up.  Included by the standalone self-test spmv_v7.cu; kept in a header so that the library can
// include exactly the code the self-test validates.  STATUS: compiles for sm_100a, not yet run on a GPU.
#pragma once
#include <cuda_runtime.h>

typedef long long v7_i64;
#ifndef V7_CARRY_RED
#define V7_CARRY_RED 1
#endif
#ifndef V7_EARLY_PREFETCH          // 1: next tile's indices are fetched right after the gathers (costs a few spills), 0: after the tile
#define V7_EARLY_PREFETCH 1
#endif
constexpr int V7_THREADS = 1024, V7_WARPS = V7_THREADS / 32, V7_ITEMS = 16, V7_TILE = 32 * V7_ITEMS;


__device__ __forceinline__ double v7_lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void v7_fetch_idx(int (&ri)[V7_ITEMS], const int* __restrict__ idx, int start, int end, int lane) {
    if (end - start == V7_TILE) {
        const int4* ip = reinterpret_cast<const int4*>(idx + start) + lane * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int4 a = ip[q];
            ri[4 * q] = a.x; ri[4 * q + 1] = a.y; ri[4 * q + 2] = a.z; ri[4 * q + 3] = a.w;
        }
    } else if (end > start) {
#pragma unroll
        for (int j = 0; j < V7_ITEMS; ++j) ri[j] = idx[min(start + lane * V7_ITEMS + j, end - 1)];
    }
}

// One tile.  `out_tile` = out + tile_out[t]; slot 0 = piece continued from the previous tile.
template <bool BINARY, bool PARTIAL>
__device__ __forceinline__ void v7_tile_body(const int (&ri)[V7_ITEMS], const double* __restrict__ val, int start, int len,
                                          unsigned meta, int nheads, int lane, unsigned sbase, double* __restrict__ out_tile) {
    const unsigned f = meta & 0xffffu;
    unsigned o = meta >> 16;                      // heads in the lower lanes = ordinal of the piece open at lane start
    double g[V7_ITEMS];
#pragma unroll
    for (int j = 0; j < V7_ITEMS; ++j) g[j] = v7_lds_f64(sbase + ((unsigned)ri[j] << 3));
    if (!BINARY) {
#pragma unroll
        for (int j = 0; j < V7_ITEMS; ++j) {
            const int q = lane * V7_ITEMS + j;
            g[j] *= val[start + (PARTIAL ? min(q, len - 1) : q)];
        }
    }
    // Branch-free serial pass.  At a head the running sum is the total of the piece that ends there and is stored to
    // the piece's slot right away (predicated store, running 32-bit slot offset).  For the FIRST head of a lane that
    // value still lacks the carry of the lower lanes: it is added after the scan, either by a fire-and-forget
    // red.global.add.f64 onto the stored value (V7_CARRY_RED: one store + one add per slot, so the result is the
    // exactly rounded sum of the two, deterministic) or by keeping the lane's leading sum in a register.
    double run = 0.0;
#if !V7_CARRY_RED
    double first_run = 0.0;
    const int jf = __ffs((int)f) - 1;             // position of the lane's first head (-1: none)
#endif
    double* op = out_tile + o;                    // slot of the piece open at the current position
#pragma unroll
    for (int j = 0; j < V7_ITEMS; ++j) {
        double gj = g[j];
        if (PARTIAL) gj = (lane * V7_ITEMS + j < len) ? gj : 0.0;
#if !V7_CARRY_RED
        first_run = (j == jf) ? run : first_run;
#endif
        // head at j:  *op++ = run; run = gj      else:  run += gj      (one predicate, no branch)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t.reg .f64 s;\n\t"
                     "and.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\t"
                     "@p st.global.f64 [%1], %0;\n\t@p add.u64 %1, %1, 8;\n\t"
                     "add.rn.f64 s, %0, %4;\n\tselp.f64 %0, %4, s, p;\n\t}"
                     : "+d"(run), "+l"(op) : "r"(f), "r"(1u << j), "d"(gj) : "memory");
    }
    // segmented inclusive scan of the lane tails across the warp
    const unsigned hm = __ballot_sync(0xffffffffu, f != 0u);
    double x = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d && ((hm >> (lane - d + 1)) & ((1u << d) - 1u)) == 0u) x += y;
    }
    double carry = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane == 0) carry = 0.0;
#if V7_CARRY_RED
    if (f != 0u) asm volatile("red.global.add.f64 [%0], %1;" : : "l"(out_tile + (meta >> 16)), "d"(carry) : "memory");
#else
    if (f != 0u) out_tile[meta >> 16] = first_run + carry;            // the piece that ends at the lane's first head
#endif
    if (lane == 31) out_tile[nheads] = (f != 0u) ? run : run + carry; // the piece that reaches the end of the tile
}

// Fast path for tiles with at most V7_SLOTS - 1 heads: the piece sums are parked in a small per-warp shared buffer
// (32-bit addresses: one predicated STS + one predicated add per element instead of a 64-bit pointer chain), the lane
// that owns a piece's first head adds the carry to its own entry, and the warp then writes the tile's slots
// to global memory with ONE coalesced store instead of ~26 scattered 8-byte ones.
constexpr int V7_SLOTS = 64;
static_assert(V7_SLOTS == 64, "the copy-out below handles exactly two slots per lane");
template <bool BINARY, bool PARTIAL>
__device__ __forceinline__ void v7_tile_body_staged(int (&ri)[V7_ITEMS], const double* __restrict__ val, int start, int len,
                                                 unsigned meta, int nheads, int lane, unsigned sbase, unsigned wbuf,
                                                 double* __restrict__ out_tile,
                                                 const int* __restrict__ idx, int next_start, int next_end) {
    const unsigned f = meta & 0xffffu;
    double run = 0.0;
    const unsigned first = wbuf + ((meta >> 16) << 3);   // shared address of the slot open at the start of the lane
    unsigned sp = first;
    // pattern-only: all 16 gathers in flight, then the serial pass; valued: two chunks of 8 so that gathered entries
    // and values (4 x 128-bit loads per chunk, issued before the gathers) fit the 64-register budget
    constexpr int CH = BINARY ? V7_ITEMS : 8;
#pragma unroll
    for (int c = 0; c < V7_ITEMS; c += CH) {
        double g[CH];
        if (!BINARY) {
            const double2* vp = reinterpret_cast<const double2*>(val + start + lane * V7_ITEMS + c);
#pragma unroll
            for (int q = 0; q < CH / 2; ++q) { const double2 v = vp[q]; g[2 * q] = v.x; g[2 * q + 1] = v.y; }
#pragma unroll
            for (int j = 0; j < CH; ++j) g[j] *= v7_lds_f64(sbase + ((unsigned)ri[c + j] << 3));
        } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) g[j] = v7_lds_f64(sbase + ((unsigned)ri[c + j] << 3));
        }
        // early prefetch: the index registers are dead once the last gather has been issued (next_end <= next_start: none)
        if (c + CH == V7_ITEMS) v7_fetch_idx(ri, idx, next_start, next_end, lane);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            // head at c + j:  *sp++ = run; run = g      else:  run += g      (one predicate, no branch)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t.reg .f64 s;\n\t"
                         "and.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\t"
                         "@p st.shared.f64 [%1], %0;\n\t@p add.u32 %1, %1, 8;\n\t"
                         "add.rn.f64 s, %0, %4;\n\tselp.f64 %0, %4, s, p;\n\t}"
                         : "+d"(run), "+r"(sp) : "r"(f), "r"(1u << (c + j)), "d"(g[j]) : "memory");
        }
    }
    const unsigned hm = __ballot_sync(0xffffffffu, f != 0u);
    double x = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d && ((hm >> (lane - d + 1)) & ((1u << d) - 1u)) == 0u) x += y;
    }
    double carry = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane == 0) carry = 0.0;
    if (f != 0u) {                                        // this lane parked the piece that ends at its first head
        double lead;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(lead) : "r"(first) : "memory");
        lead += carry;
        asm volatile("st.shared.f64 [%0], %1;" : : "r"(first), "d"(lead) : "memory");
    }
    if (lane == 31) {                                     // the piece that reaches the end of the tile
        const double last = (f != 0u) ? run : run + carry;
        asm volatile("st.shared.f64 [%0], %1;" : : "r"(wbuf + ((unsigned)nheads << 3)), "d"(last) : "memory");
    }
    __syncwarp();
    // nheads < V7_SLOTS = 64: at most two coalesced stores per lane
    if (lane <= nheads) out_tile[lane] = v7_lds_f64(wbuf + ((unsigned)lane << 3));
    if (lane + 32 <= nheads) out_tile[lane + 32] = v7_lds_f64(wbuf + ((unsigned)(lane + 32) << 3));
    __syncwarp();                                         // the buffer is reused by the next tile
}

template <bool BINARY>
__global__ void __launch_bounds__(V7_THREADS, 1)
k_seg_spmv_v7(const int* __restrict__ idx, const double* __restrict__ val, const unsigned* __restrict__ lane_meta,
              const int2* __restrict__ tile_out /* {first slot, n_heads} */, const int* __restrict__ slab_tile0,
              const int* __restrict__ slab_nnz0, const int* __restrict__ slab_nnz1, int nslab, int ntiles,
              const double* __restrict__ gvec, int W, v7_i64 n_gather, double* __restrict__ out,
              const int* __restrict__ done_flag) {
    if (done_flag != nullptr && *done_flag) return;
    extern __shared__ double sv[];                        // [W staged entries][V7_WARPS x V7_SLOTS piece sums]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned wbuf = (unsigned)__cvta_generic_to_shared(sv + W + warp * V7_SLOTS);
    const int t_lo = (int)((v7_i64)ntiles * blockIdx.x / gridDim.x);
    const int t_hi = (int)((v7_i64)ntiles * (blockIdx.x + 1) / gridDim.x);
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
        const v7_i64 gbase = (v7_i64)slab * W;
        __syncthreads();                                   // the previous section's readers are done
        {
            const v7_i64 rem = n_gather - gbase;
            const int wlen = rem < (v7_i64)W ? (int)rem : W;
            const double* src = gvec + gbase;
            int i = tid;
            for (; i + 3 * V7_THREADS < wlen; i += 4 * V7_THREADS) {
                double a0 = src[i], a1 = src[i + V7_THREADS], a2 = src[i + 2 * V7_THREADS], a3 = src[i + 3 * V7_THREADS];
                sv[i] = a0; sv[i + V7_THREADS] = a1; sv[i + 2 * V7_THREADS] = a2; sv[i + 3 * V7_THREADS] = a3;
            }
            for (; i < wlen; i += V7_THREADS) sv[i] = src[i];
        }
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(sv) - (unsigned)((unsigned)gbase << 3);
        int t = cur + warp;
        int ri[V7_ITEMS];
        unsigned meta = 0u;
        int2 tm = make_int2(0, 0);
        if (t < sec_end) {
            const int st0 = nnz0 + (t - s_t0) * V7_TILE;
            v7_fetch_idx(ri, idx, st0, min(st0 + V7_TILE, nnz1), lane);
            meta = lane_meta[(v7_i64)t * 32 + lane];
            tm = tile_out[t];
        }
        __syncthreads();
        while (t < sec_end) {
            const int start = nnz0 + (t - s_t0) * V7_TILE;
            const int len = min(start + V7_TILE, nnz1) - start;
            double* out_tile = out + tm.x;
            const int tn = t + V7_WARPS;
            const int st1 = nnz0 + (tn - s_t0) * V7_TILE;
            const int en1 = tn < sec_end ? min(st1 + V7_TILE, nnz1) : st1;       // empty range: nothing to prefetch
            unsigned meta_next = 0u;
            int2 tm_next = make_int2(0, 0);
            if (tn < sec_end) { meta_next = lane_meta[(v7_i64)tn * 32 + lane]; tm_next = tile_out[tn]; }
            if (len == V7_TILE && tm.y < V7_SLOTS) {
#if V7_EARLY_PREFETCH
                v7_tile_body_staged<BINARY, false>(ri, val, start, len, meta, tm.y, lane, sbase, wbuf, out_tile, idx, st1, en1);
#else
                v7_tile_body_staged<BINARY, false>(ri, val, start, len, meta, tm.y, lane, sbase, wbuf, out_tile, idx, st1, st1);
                v7_fetch_idx(ri, idx, st1, en1, lane);
#endif
            } else {
                if (len == V7_TILE) v7_tile_body<BINARY, false>(ri, val, start, len, meta, tm.y, lane, sbase, out_tile);
                else if (len > 0)   v7_tile_body<BINARY, true>(ri, val, start, len, meta, tm.y, lane, sbase, out_tile);
                else if (lane == 0) out_tile[0] = 0.0;     // a slab without nnz: its single empty tile
                v7_fetch_idx(ri, idx, st1, en1, lane);
            }
            t = tn; meta = meta_next; tm = tm_next;
        }
        cur = sec_end;
        slab += 1;
    }
}

// adds the pieces continued from previous tiles to the slot of the segment they belong to, in tile order
__global__ void k_fixup_v7(const int* __restrict__ chead_slot, const int2* __restrict__ tile_out, int ntiles,
                           double* __restrict__ out, const int* __restrict__ done_flag) {
    if (done_flag != nullptr && *done_flag) return;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int s = chead_slot[t];
    if (s < 0) return;
    if (t > 0 && chead_slot[t - 1] == s) return;
    double acc = 0.0;
    for (int tt = t; tt < ntiles && chead_slot[tt] == s; ++tt) acc += out[tile_out[tt].x];
    out[s] += acc;
}

