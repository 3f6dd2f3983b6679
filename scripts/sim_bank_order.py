"""CPU model of the shared-memory gather of k_seg_spmv and of the build-time nnz ordering (k_bank_permute).

A tile is 256 nnz = 32 lanes x 8 loads; load j of lane l reads element 8l + j; the 16 lanes of a half-warp are served
together and cost one wavefront per distinct entry of the busiest 8-byte bank (index mod 16).  The script draws random
gather indices, cuts the tile into pieces (segments) of a given mean length, applies the same greedy ordering as the
kernel and prints the mean wavefronts per half-warp load before / after (1.0 would be conflict-free).
DESIGN.md 3.1 and profiles/r01_spmv_history.md quote its output for C4-like pieces (~20 nnz): 3.06 -> 2.04, which is
what ncu's source page then showed on the GPU (4.1 wavefronts per LDS.64 = 2 half-warps x 2.04).

usage: python scripts/sim_bank_order.py [tiles per setting]"""
import sys
import numpy as np

GROUP = [((k & 7) << 1) | (k >> 7) for k in range(256)]      # (load j, half-warp) group of tile position k


def wavefronts(idx):
    banks = idx % 16
    total = 0
    for g in range(16):
        pos = [k for k in range(256) if GROUP[k] == g]
        total += np.bincount(banks[pos], minlength=16).max()
    return total / 16.0


def greedy_order(idx, cuts):
    """Same rule as k_bank_permute: walk the positions of a piece in order and take, among the piece's remaining
    nnz, one whose bank is least used so far by the position's group (first such nnz wins ties)."""
    e = idx.copy()
    occ = np.zeros((16, 16), int)
    a = 0
    for b in list(cuts) + [256]:
        for k in range(a, b):
            g = GROUP[k]
            best, bo = k, occ[g, e[k] % 16]
            m = k + 1
            while bo > 0 and m < b:
                o = occ[g, e[m] % 16]
                if o < bo:
                    bo, best = o, m
                m += 1
            e[k], e[best] = e[best], e[k]
            occ[g, e[k] % 16] = bo + 1
        a = b
    return e


def main():
    tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    print('mean piece length | canonical order | bank-aware order   (wavefronts per half-warp load)')
    for mean in (5, 20, 50, 100, 256):
        rng = np.random.default_rng(0)
        w0 = w1 = 0.0
        for _ in range(tiles):
            idx = rng.integers(0, 20864, 256)
            cuts, pos = [], 0
            while True:
                pos += max(1, rng.poisson(mean))
                if pos >= 256:
                    break
                cuts.append(pos)
            w0 += wavefronts(idx)
            w1 += wavefronts(greedy_order(idx, cuts))
        print('%17d | %15.2f | %16.2f' % (mean, w0 / tiles, w1 / tiles))


if __name__ == '__main__':
    main()
