#!/usr/bin/env python3
"""Synthetic MatrixMarket inputs for the BASELINE.json configs.

All matrices are "coordinate real general" (so the reference's LU path runs, not LLT)
unless --symmetric is given, use natural lexicographic node numbering, and are emitted
row by row with ascending columns at %.17g.  The right-hand side <base>_b.mtx is
"array real general" with b_i = 1 + 0.25*(i mod 7).

  lap2d NX [NY]     5-point Laplacian, diag 4, off -1, Dirichlet
  lap3d NX [NY NZ]  7-point Laplacian, diag 6, off -1
  nine2d NX [NY]    9-point stencil, diag 8, 8 neighbours -1
  banded N [W] [K]  unsymmetric, diagonally dominant, K (=9) distinct off-diagonals per
                    row at columns i + U{-W..W}\\{0} (W=2048), values U(-1,1),
                    diag = 1 + sum|off|; numpy PCG64 seed 12345
"""
import argparse
import os
import sys

import numpy as np


def _stencil(shape, offsets, diag, off):
    """COO of a constant-coefficient stencil on a grid with Dirichlet boundary."""
    dims = len(shape)
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.int64)
    coords = np.unravel_index(idx, shape)  # C order: last axis fastest
    rows = [idx]
    cols = [idx]
    vals = [np.full(n, float(diag))]
    for o in offsets:
        ok = np.ones(n, dtype=bool)
        nb = []
        for d in range(dims):
            c = coords[d] + o[d]
            ok &= (c >= 0) & (c < shape[d])
            nb.append(c)
        j = np.ravel_multi_index([np.clip(nb[d], 0, shape[d] - 1) for d in range(dims)], shape)
        rows.append(idx[ok])
        cols.append(j[ok])
        vals.append(np.full(int(ok.sum()), float(off)))
    r = np.concatenate(rows)
    c = np.concatenate(cols)
    v = np.concatenate(vals)
    order = np.lexsort((c, r))
    return n, r[order], c[order], v[order]


def lap2d(nx, ny=None):
    ny = ny or nx
    return _stencil((ny, nx), [(-1, 0), (1, 0), (0, -1), (0, 1)], 4.0, -1.0)


def lap3d(nx, ny=None, nz=None):
    ny = ny or nx
    nz = nz or nx
    offs = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
    return _stencil((nz, ny, nx), offs, 6.0, -1.0)


def nine2d(nx, ny=None):
    ny = ny or nx
    offs = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1) if (a, b) != (0, 0)]
    return _stencil((ny, nx), offs, 8.0, -1.0)


def banded(n, w=2048, k=9, seed=12345):
    rng = np.random.Generator(np.random.PCG64(seed))
    rows, cols, vals = [], [], []
    chunk = 1 << 16
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        m = e - s
        i = np.arange(s, e, dtype=np.int64)
        # k distinct non-zero offsets in [-w, w] per row: sample without replacement by
        # drawing k distinct values from 2w candidates via argpartition of random keys
        # would be heavy; rejection on duplicates is cheap because k << 2w.
        off = rng.integers(1, w + 1, size=(m, k)) * rng.choice(np.array([-1, 1]), size=(m, k))
        for _ in range(8):
            so = np.sort(off, axis=1)
            dup = np.zeros((m, k), dtype=bool)
            dup[:, 1:] = so[:, 1:] == so[:, :-1]
            if not dup.any():
                off = so
                break
            repl = rng.integers(1, w + 1, size=int(dup.sum())) * rng.choice(np.array([-1, 1]), size=int(dup.sum()))
            so[dup] = repl
            off = so
        else:
            off = np.sort(off, axis=1)
        j = i[:, None] + off
        ok = (j >= 0) & (j < n)
        v = rng.uniform(-1.0, 1.0, size=(m, k))
        v[v == 0.0] = 0.5
        diag = 1.0 + np.where(ok, np.abs(v), 0.0).sum(axis=1)
        rr = np.concatenate([np.repeat(i, k)[ok.ravel()], i])
        cc = np.concatenate([j.ravel()[ok.ravel()], i])
        vv = np.concatenate([v.ravel()[ok.ravel()], diag])
        rows.append(rr)
        cols.append(cc)
        vals.append(vv)
    r = np.concatenate(rows)
    c = np.concatenate(cols)
    v = np.concatenate(vals)
    # drop accidental duplicate (row, col) pairs left after the rejection rounds
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    keep = np.ones(len(r), dtype=bool)
    keep[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    return n, r[keep], c[keep], v[keep]


def rhs(n):
    return 1.0 + 0.25 * (np.arange(n) % 7)


def write_mtx(path, n, r, c, v, symmetric=False):
    if symmetric:
        keep = r >= c
        r, c, v = r[keep], c[keep], v[keep]
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real %s\n" % ("symmetric" if symmetric else "general"))
        f.write("%d %d %d\n" % (n, n, len(r)))
        step = 1 << 20
        for s in range(0, len(r), step):
            e = min(len(r), s + step)
            lines = ["%d %d %.17g\n" % (a + 1, b + 1, x) for a, b, x in zip(r[s:e].tolist(), c[s:e].tolist(), v[s:e].tolist())]
            f.write("".join(lines))
    b = rhs(n)
    base = path[: path.find(".mtx")]
    with open(base + "_b.mtx", "w") as f:
        f.write("%%MatrixMarket matrix array real general\n")
        f.write("%d 1\n" % n)
        f.write("".join("%.17g\n" % x for x in b.tolist()))


KINDS = {"lap2d": lap2d, "lap3d": lap3d, "nine2d": nine2d, "banded": banded}


def generate(kind, *args):
    return KINDS[kind](*args)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("kind", choices=sorted(KINDS))
    ap.add_argument("dims", type=int, nargs="+")
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--symmetric", action="store_true")
    a = ap.parse_args(argv)
    n, r, c, v = generate(a.kind, *a.dims)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    write_mtx(a.out, n, r, c, v, a.symmetric)
    print("wrote %s n=%d nnz=%d" % (a.out, n, len(r)))


if __name__ == "__main__":
    sys.exit(main())
