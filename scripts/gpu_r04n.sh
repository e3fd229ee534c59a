#!/bin/bash
OUT=gpurun_out/r04n; mkdir -p $OUT
MPS_B200_DUMP_NONCONV=$OUT/stuck timeout 600 python scripts/configs_fullsize.py --which c5 --chi5 512 --fuse-both-upto 0 --budget 300 --out $OUT/configs.jsonl > $OUT/configs.log 2>&1
ls -la $OUT
python - <<PY
import numpy as np, glob
for fn in sorted(glob.glob("$OUT/stuck_*.bin")):
    raw = np.fromfile(fn)
    M, N, Mg, tol, ntol, f2 = int(raw[0]), int(raw[1]), int(raw[2]), raw[3], raw[4], raw[5]
    G = raw[8:8 + 2 * M * N].view(np.complex128).reshape(N, M).T   # column-major M x N
    cn = np.linalg.norm(G, axis=0)
    print(fn, "M N", M, N, "tol", tol, "ntol", ntol, "fro2", f2, "sum cn2", (cn ** 2).sum())
    print(" col norms: max %.3e min %.3e  sorted head %s tail %s" % (cn.max(), cn.min(), np.sort(cn)[::-1][:4], np.sort(cn)[:6]))
    dead = np.sqrt(ntol ** 2 * f2 / N)
    print(" dead threshold on norm %.3e ; columns below: %d ; within 100x above: %d" % (dead, (cn < dead).sum(), ((cn >= dead) & (cn < 100 * dead)).sum()))
    W = G.conj().T @ G
    rel = np.abs(W) / np.maximum(1e-300, np.outer(cn, cn))
    np.fill_diagonal(rel, 0)
    alive = cn > dead
    r2 = rel[np.ix_(alive, alive)]
    idx = np.argwhere(r2 > tol)
    print(" alive %d ; pairs above tol: %d ; above 16 tol: %d ; max rel %.3e" % (alive.sum(), (r2 > tol).sum() // 2, (r2 > 16 * tol).sum() // 2, r2.max()))
    ai = np.where(alive)[0]
    big = np.argwhere(r2 > 16 * tol)[:10]
    for i, j in big:
        if i < j: print("   pair", ai[i], ai[j], "rel %.3e norms %.3e %.3e" % (r2[i, j], cn[ai[i]], cn[ai[j]]))
    print(" nan/inf:", np.isnan(G).sum(), np.isinf(G).sum())
PY
rm -f $OUT/stuck_*.bin
