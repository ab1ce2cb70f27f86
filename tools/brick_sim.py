#!/usr/bin/env python
"""CPU model of the PME brick segmentation (pme_brick.cu nextBrickSet): how many bricks does a CTA of T sorted atoms need?
Usage: python tools/brick_sim.py dhfr 128 24 24 28"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openmmlab_b200 as mm
from openmmlab_b200 import systems


def hilbert(x, y, z, bits):
    X = [x.astype(np.uint32).copy(), y.astype(np.uint32).copy(), z.astype(np.uint32).copy()]
    M = np.uint32(1 << (bits-1))
    Q = int(M)
    while Q > 1:
        P = np.uint32(Q-1)
        for i in range(3):
            m = (X[i] & np.uint32(Q)) != 0
            X[0] = np.where(m, X[0] ^ P, X[0])
            t = np.where(m, 0, (X[0] ^ X[i]) & P).astype(np.uint32)
            X[0] ^= t
            X[i] ^= t
        Q >>= 1
    for i in range(1, 3):
        X[i] ^= X[i-1]
    t = np.zeros_like(X[0])
    Q = int(M)
    while Q > 1:
        t = np.where((X[2] & np.uint32(Q)) != 0, t ^ np.uint32(Q-1), t)
        Q >>= 1
    for i in range(3):
        X[i] ^= t
    h = np.zeros(x.shape, dtype=np.uint64)
    for b in range(bits-1, -1, -1):
        for i in range(3):
            h = (h << np.uint64(1)) | ((X[i] >> np.uint32(b)) & np.uint32(1)).astype(np.uint64)
    return h


def main():
    name = sys.argv[1]
    T = int(sys.argv[2])
    B = [int(v) for v in sys.argv[3:6]]
    w = {"dhfr": systems.dhfr_like, "apoa1": systems.apoa1_like, "fep": systems.fep_ligand, "stmv": systems.stmv_like, "water": systems.water_box}[name]()
    pos = np.asarray(w.positions, dtype=np.float64)
    box = np.asarray(w.box, dtype=np.float64).reshape(3, 3)
    L = np.array([box[0, 0], box[1, 1], box[2, 2]])
    N = len(pos)
    cs = (4.0*L.prod()/N)**(1/3)
    dim = np.maximum(1, np.minimum(256, np.floor(L/cs+0.5).astype(int)))
    frac = pos/L
    frac -= np.floor(frac)
    c = np.minimum(dim-1, (frac*dim).astype(int))
    bits = 1
    while (1 << bits) < dim.max():
        bits += 1
    gx, gy, gz = np.meshgrid(np.arange(dim[0]), np.arange(dim[1]), np.arange(dim[2]), indexing="ij")
    keys = hilbert(gx.ravel(), gy.ravel(), gz.ravel(), bits)
    order = np.argsort(keys, kind="stable")
    rank = np.empty(len(keys), dtype=np.int64)
    rank[order] = np.arange(len(keys))
    lin = (c[:, 0]*dim[1]+c[:, 1])*dim[2]+c[:, 2]
    atom_rank = rank[lin]
    sorted_atoms = np.lexsort((np.arange(N), atom_rank))
    f = w.force
    _, nx, ny, nz = f.getPMEParameters()
    n = np.array([nx, ny, nz])
    idx = (frac*n).astype(int) % n
    sub = np.array([f.getParticleSubset(i) for i in range(N)])
    q = np.array([f.getParticleParameters(i)[0] for i in range(N)])
    idx, sub, q = idx[sorted_atoms], sub[sorted_atoms], q[sorted_atoms]
    rounds = []
    sizes = []
    for b0 in range(0, N, T):
        I, S, Q = idx[b0:b0+T], sub[b0:b0+T], q[b0:b0+T]
        pending = Q != 0
        r = 0
        while pending.any():
            first = np.argmax(pending)
            cand = pending & (S == S[first])
            ids = np.nonzero(cand)[0]
            limit = len(ids)
            while True:
                sel = ids[:limit]
                lo, hi = I[sel].min(0), I[sel].max(0)
                org = lo.copy(); org[2] &= ~3
                if ((hi-org+5) <= np.array(B)).all() or limit <= 1:
                    break
                limit = (limit+1)//2
            pending[sel] = False
            sizes.append(len(sel))
            r += 1
        rounds.append(r)
    rounds = np.array(rounds)
    print(name, "N", N, "cells", dim, "grid", n, "T", T, "brick", B)
    print("CTAs", len(rounds), "bricks", rounds.sum(), "mean rounds", rounds.mean(), "max", rounds.max(), "hist", np.bincount(rounds)[:12])
    print("atoms per brick: mean", np.mean(sizes), "median", np.median(sizes))


main()
