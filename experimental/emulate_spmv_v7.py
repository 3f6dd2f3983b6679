"""Executable specification (numpy, CPU) of the NEXT SpMV kernel design ("v7"): 512-nnz tiles, 16 nnz per lane, head
flags precomputed at build time, piece sums emitted straight from registers (no shared-memory prefix buffer).

STATUS: design prototype.  Nothing here is built into libbbgpu.so or used by the product path; the CUDA translation
(experimental/spmv_v7.cu) has not run on a GPU yet.  The emulator mirrors the warp algorithm lane by lane (arrays of
32 lanes, shuffles as array shifts) so that the index logic -- flags, piece ordinals, empty segments, partial tiles,
segments that straddle tiles and slabs -- can be checked against a plain segmented sum on the CPU
(tests/test_spmv_v7_emulator.py).  Motivation and instruction budget: profiles/r01_spmv_history.md.

Format (per product, built once):
  idx/val     slab-major nnz stream (as today); every slab start padded to a multiple of 4 nnz
  ptr[V+1]    nnz offsets of the virtual segments v = slab * n_seg + seg
  tiles       per slab, consecutive runs of T = 512 nnz (the last one of a slab may be partial)
  lane_meta   per tile and lane: bit j (j < 16) set  <=>  position 16*lane + j starts a NON-EMPTY virtual segment
              ("head"); bits 16.. = number of heads in the lower lanes of the tile
  cbase[t]    number of heads in all earlier tiles; n_heads[t]
  cpart       COMPACT output: one entry per non-empty virtual segment, in nnz order (= head order), so the entry of
              the o-th head of tile t is cbase[t] + o - 1 -- no segment ids in the hot loop, and empty virtual
              segments (very common: a rare column has no entry in most row slabs) cost nothing
  cstart[V+1] exclusive prefix count of the non-empty virtual segments: segment v is non-empty iff
              cstart[v+1] > cstart[v], and then its sum is cpart[cstart[v]] (read by the kernels that consume the product)
  chead[t]    compact index of the segment continued from the previous tile (or -1)

Warp algorithm for one tile (lane l owns positions 16l .. 16l+15):
  run = 0; ord = heads in lower lanes
  for j in 0..15:  g = val * x[idx]   (0 past the end of a partial tile)
      if head at j:  the piece with ordinal `ord` ends just before j:
                        first head of the lane -> remember first_run = run (its piece may have started in a lower lane)
                        otherwise              -> emit(ord, run)
                     ord += 1; run = g
      else:          run += g
  segmented inclusive scan of (run, lane has a head) across lanes; carry = value of the lane below (0 for lane 0)
  lane with a head:   emit(heads in lower lanes, first_run + carry)
  lane 31:            emit(n_heads, run + (carry if the lane has no head))          # the piece that reaches the tile end
  emit(0, s) -> head_part[tile] = s ; emit(o > 0, s) -> cpart[cbase[tile] + o - 1] = s
k_fixup: cpart[chead[t]] += head_part[t] for the continuation tiles, in tile order.
consumer: y[seg] = sum over slabs of (cpart[cstart[v]] if cstart[v+1] > cstart[v] else 0), v = slab * n_seg + seg.
"""
import numpy as np

T, LANES, ITEMS = 512, 32, 16


def build_format(indptr, indices, data, n_gather, W):
    """Slab-major copy of a compressed matrix (rows = segments) + the static per-tile metadata."""
    n_seg = len(indptr) - 1
    nslab = max(1, -(-n_gather // W))
    seg_of = np.repeat(np.arange(n_seg), np.diff(indptr))
    slab_of = indices // W
    order = np.lexsort((np.arange(len(indices)), seg_of, slab_of))         # stable: slab, then segment
    V = nslab * n_seg
    counts = np.bincount(slab_of * n_seg + seg_of, minlength=V)
    idx, val, ptr = [], [], np.zeros(V + 1, np.int64)
    slab_range, pos = [], 0
    sorted_idx, sorted_val = indices[order], (None if data is None else data[order])
    per_slab = np.bincount(slab_of, minlength=nslab)
    src = 0
    for s in range(nslab):
        pad = (-pos) % 4
        idx.append(np.zeros(pad, np.int64)); val.append(np.zeros(pad)); pos += pad
        a = pos
        ptr[s * n_seg:(s + 1) * n_seg] = a + np.concatenate(([0], np.cumsum(counts[s * n_seg:(s + 1) * n_seg])[:-1]))
        idx.append(sorted_idx[src:src + per_slab[s]])
        val.append(np.ones(per_slab[s]) if sorted_val is None else sorted_val[src:src + per_slab[s]])
        src += per_slab[s]; pos += per_slab[s]
        slab_range.append((a, pos))
    ptr[V] = pos
    # a segment's end is the next segment's start, except at a slab boundary (padding in between)
    seg_end = ptr[1:].copy()
    for s in range(nslab):
        seg_end[(s + 1) * n_seg - 1] = slab_range[s][1]
    fmt = dict(idx=np.concatenate(idx), val=np.concatenate(val), ptr=ptr, seg_end=seg_end, V=V, n_seg=n_seg,
               nslab=nslab, W=W, slab_range=slab_range)
    nonempty = np.nonzero(seg_end > ptr[:-1])[0]
    head_pos = ptr[nonempty]                                    # ascending
    cstart = np.concatenate(([0], np.cumsum(seg_end > ptr[:-1])))
    tiles, lane_meta, cbase, n_heads, chead = [], [], [], [], []
    for s, (a, b) in enumerate(slab_range):
        nt = max(1, -(-(b - a) // T))
        for k in range(nt):
            start, end = a + k * T, min(a + k * T + T, b)
            end = max(end, start)
            lo, hi = np.searchsorted(head_pos, start), np.searchsorted(head_pos, end)
            segs = nonempty[lo:hi]
            rel = head_pos[lo:hi] - start
            flags = np.zeros(LANES, np.int64)
            np.bitwise_or.at(flags, rel // ITEMS, 1 << (rel % ITEMS))
            heads_per_lane = np.bincount(rel // ITEMS, minlength=LANES)
            below = np.concatenate(([0], np.cumsum(heads_per_lane)[:-1]))
            cbase.append(int(lo))
            n_heads.append(len(segs))
            lane_meta.append(flags | (below << 16))
            # segment continued from the previous tile: the one containing nnz `start`, unless a head sits there
            if end > start and not (len(rel) and rel[0] == 0):
                chead.append(int(lo) - 1)
            else:
                chead.append(-1)
            tiles.append((start, end))
    fmt.update(tiles=tiles, lane_meta=lane_meta, cbase=cbase, n_heads=n_heads, chead=np.array(chead, np.int64),
               cstart=cstart, n_compact=len(nonempty))
    return fmt


def warp_tile(fmt, t, x, cpart, head_part):
    """One tile, lane by lane, exactly as the kernel is meant to do it."""
    start, end = fmt['tiles'][t]
    n = end - start
    cbase, n_heads = fmt['cbase'][t], fmt['n_heads'][t]
    meta = fmt['lane_meta'][t]

    def emit(o, s):
        if o == 0:
            head_part[t] = s
        else:
            cpart[cbase + o - 1] = s

    if n <= 0:
        head_part[t] = 0.0
        return
    run = np.zeros(LANES)
    first_run = np.zeros(LANES)
    has_head = np.zeros(LANES, bool)
    for l in range(LANES):
        f, o = int(meta[l]) & 0xffff, int(meta[l]) >> 16
        has_head[l] = f != 0
        r = 0.0
        for j in range(ITEMS):
            q = ITEMS * l + j
            g = fmt['val'][start + q] * x[fmt['idx'][start + q]] if q < n else 0.0
            if (f >> j) & 1:
                if f & ((1 << j) - 1) == 0:
                    first_run[l] = r
                else:
                    emit(o, r)
                o += 1
                r = g
            else:
                r += g
        run[l] = r
    # segmented inclusive scan across lanes (5 shuffle steps; a ballot of `has_head` gates the adds)
    xs = run.copy()
    d = 1
    while d < LANES:
        y = np.concatenate((np.zeros(d), xs[:-d]))
        for l in range(d, LANES):
            if not has_head[l - d + 1:l + 1].any():
                xs[l] = xs[l] + y[l]
        d *= 2
    carry = np.concatenate(([0.0], xs[:-1]))
    for l in range(LANES):
        if has_head[l]:
            emit(int(meta[l]) >> 16, first_run[l] + carry[l])
    last = LANES - 1
    emit(n_heads, run[last] + (0.0 if has_head[last] else carry[last]))


def spmv(fmt, x):
    """y[seg] = sum over the slabs of part[slab * n_seg + seg], after the fix-up of the straddling segments."""
    cpart = np.full(fmt['n_compact'], np.nan)                  # every entry must be written by exactly one emit
    head_part = np.zeros(len(fmt['tiles']))
    for t in range(len(fmt['tiles'])):
        warp_tile(fmt, t, x, cpart, head_part)
    hs = fmt['chead']
    for t in range(len(hs)):                                   # k_fixup
        if hs[t] >= 0:
            cpart[hs[t]] += head_part[t]
    cs = fmt['cstart']
    dense = np.where(cs[1:] > cs[:-1], cpart[np.minimum(cs[:-1], max(len(cpart) - 1, 0))] if len(cpart) else 0.0, 0.0)
    return dense.reshape(fmt['nslab'], fmt['n_seg']).sum(axis=0)
