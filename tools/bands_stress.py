"""K6 two-pass path against numpy over odd shapes / distributions / percentile sets: python tools/bands_stress.py"""
import sys, numpy as np
sys.path.insert(0, ".")
from ravest_b200 import _lib
rng = np.random.default_rng(5)
bad = 0
for (S, T) in [(8192, 1), (8193, 8), (10000, 9), (65536, 33), (200000, 3), (12345, 17), (300001, 2), (9000, 64)]:
    for kind in range(6):
        if kind == 0: A = rng.normal(size=(S, T))
        elif kind == 1: A = rng.integers(-3, 4, size=(S, T)).astype(np.float64)
        elif kind == 2: A = np.exp(rng.normal(0, 10, size=(S, T))) * rng.choice([-1.0, 1.0], size=(S, T))
        elif kind == 3: A = np.cumsum(rng.normal(size=(S, T)), axis=0)              # trending with the row index
        elif kind == 4: A = rng.normal(1e9, 1e-3, size=(S, T))                      # tiny spread on a huge offset
        else:
            A = rng.normal(size=(S, T)); A[rng.random((S, T)) < 0.4] = 0.25        # 40 % ties
        for q in ([15.85, 50, 84.15], [0, 100], [0.01, 99.99, 33.3, 66.6, 5, 95, 50, 2.5], 73.0):
            ref = np.percentile(A, q, axis=0)
            got = _lib.percentile_columns(A, q)
            if not np.array_equal(got, ref):
                bad += 1
                print("MISMATCH", S, T, kind, q, np.argwhere(got != ref)[:3])
print("stress done, mismatches:", bad)
