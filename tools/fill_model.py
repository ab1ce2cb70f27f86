#!/usr/bin/env python
"""How much of a tile's pair slots can be in range?  A numpy / scipy model of the tile list's geometry (CPU only).

The sorted order is rebuilt as the library builds it (cells of ~4 atoms ranked along a Hilbert curve, atoms of a cell by index:
csrc/sort.cu, api.cu setupCells / hilbertIndex); consecutive atoms are then cut into i groups of G atoms, and for each group the
j atoms are the LATER atoms that come within the cutoff of the group's bounding box (the list as built every step) or of one of
its atoms (the refined list).  fill = interacting pairs / (G x list entries): the fraction of the pair slots a direct-space
kernel visits that are real work (DESIGN.md sections 7 and 9).  G = 8 is the octet of k_direct / k_directPacked.

    python tools/fill_model.py [dhfr|fep|water] [--sample 400]
"""
import argparse
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def hilbert_index(x, y, z, bits):
    # Skilling's transpose form (api.cu hilbertIndex)
    X = [x, y, z]
    M = 1 << (bits-1)
    Q = M
    while Q > 1:
        P = Q-1
        for i in range(3):
            if X[i] & Q:
                X[0] ^= P
            else:
                t = (X[0] ^ X[i]) & P
                X[0] ^= t
                X[i] ^= t
        Q >>= 1
    for i in range(1, 3):
        X[i] ^= X[i-1]
    t = 0
    Q = M
    while Q > 1:
        if X[2] & Q:
            t ^= Q-1
        Q >>= 1
    for i in range(3):
        X[i] ^= t
    h = 0
    for b in range(bits-1, -1, -1):
        for i in range(3):
            h = (h << 1) | ((X[i] >> b) & 1)
    return h


def sorted_order(pos, box):
    n = len(pos)
    cs = (4.0*np.prod(box)/n)**(1.0/3.0)
    dim = np.maximum(1, np.minimum(256, np.floor(box/cs+0.5).astype(int)))
    bits = 1
    while (1 << bits) < dim.max():
        bits += 1
    frac = pos/box
    frac -= np.floor(frac)
    c = np.minimum(dim-1, (frac*dim).astype(int))
    rank = {}
    keys = np.empty(n, dtype=np.int64)
    for i in range(n):
        cell = (int(c[i, 0]), int(c[i, 1]), int(c[i, 2]))
        if cell not in rank:
            rank[cell] = hilbert_index(cell[0], cell[1], cell[2], bits)
        keys[i] = rank[cell]
    return np.lexsort((np.arange(n), keys))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="dhfr")
    ap.add_argument("--sample", type=int, default=400, help="i groups sampled per group size")
    ap.add_argument("--cutoff", type=float, default=0.9)
    args = ap.parse_args()
    from openmmlab_b200 import systems
    w = systems.WORKLOADS[args.workload]()
    box = np.array([w.box[0][0], w.box[1][1], w.box[2][2]])
    order = sorted_order(w.positions, box)
    pos = w.positions[order] % box
    n = len(pos)
    tree = cKDTree(pos, boxsize=box)
    rc = args.cutoff
    rng = np.random.default_rng(0)
    print("%s: %d atoms, cutoff %.2f nm; fill = interacting pairs / (G x list entries), %d groups sampled" % (w.name, n, rc, args.sample))
    print("%4s %12s %12s %12s %12s" % ("G", "fill (box)", "entries/grp", "fill (atoms)", "entries/grp"))
    for G in (1, 2, 4, 8, 16, 32):
        groups = rng.choice(n//G-1, size=min(args.sample, n//G-1), replace=False)
        pairs = box_entries = atom_entries = 0
        for g in groups:
            lo, hi = g*G, g*G+G
            members = pos[lo:hi]
            # minimum-image coordinates around the first member
            rel = members-members[0]
            rel -= box*np.round(rel/box)
            centre, half = 0.5*(rel.max(0)+rel.min(0)), 0.5*(rel.max(0)-rel.min(0))
            cand = np.array(sorted(set(j for m in members for j in tree.query_ball_point(m, rc+2*np.linalg.norm(half)+1e-9) if j >= hi)), dtype=int)
            if len(cand) == 0:
                continue
            d = pos[cand]-members[0]
            d -= box*np.round(d/box)
            # distance to the group's bounding box / to its atoms
            gap = np.maximum(0.0, np.abs(d-centre)-half)
            in_box = (gap**2).sum(1) < rc*rc
            dist2 = ((d[:, None, :]-rel[None, :, :])**2).sum(2)
            in_range = dist2 < rc*rc
            pairs += in_range.sum()
            box_entries += in_box.sum()
            atom_entries += in_range.any(1).sum()
        print("%4d %12.3f %12.1f %12.3f %12.1f" % (G, pairs/max(G*box_entries, 1), box_entries/len(groups), pairs/max(G*atom_entries, 1), atom_entries/len(groups)))


if __name__ == "__main__":
    main()
