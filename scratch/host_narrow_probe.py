"""How fast can the host narrow int64 ids to int32 (round-2 idea: ship ids at 4 B/row)?  numpy releases the GIL."""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

n = 50_000_000
a = np.random.default_rng(0).integers(0, 1_000_000, n)
out = np.zeros(n, dtype=np.int32)


def run(nt):
    cuts = [n * t // nt for t in range(nt + 1)]

    def w(t):
        s = slice(cuts[t], cuts[t + 1])
        out[s] = a[s]
        return int(a[s].min()), int(a[s].max())

    t0 = time.perf_counter()
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(w, range(nt)))
    return (time.perf_counter() - t0) * 1e3


for nt in (1, 4, 8, 16, 24):
    print(nt, "threads:", round(min(run(nt) for _ in range(3)), 1), "ms for 50M int64 -> int32 + min/max")
