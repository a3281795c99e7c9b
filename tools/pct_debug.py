import numpy as np, sys
sys.path.insert(0, '.')
from ravest_b200 import _lib
rng = np.random.default_rng(1)
for S, T in [(101, 41), (5, 3), (100, 8)]:
    A = rng.normal(3, 2, size=(S, T))
    for q in ([50.0], [0.0], [100.0], [25.0], [15.85, 50, 84.15]):
        got = _lib.percentile_columns(A, q); ref = np.percentile(A, q, axis=0)
        d = np.abs(got - ref)
        print(S, T, q, "max diff", d.max(), "n diff", (d > 0).sum(), "of", d.size)
        if d.max() > 0:
            i = np.unravel_index(np.argmax(d), d.shape)
            col = np.sort(A[:, i[1]])
            n = S; qq = q[i[0]] / 100
            v = n * qq + (1 + qq * -1) - 1
            print("  worst", i, got[i], ref[i], "virt", v, "neighbours", col[int(np.floor(v))], col[min(int(np.floor(v)) + 1, n - 1)])
