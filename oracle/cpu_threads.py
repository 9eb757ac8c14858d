"""Pick the torch thread count that actually runs the oracle's GEMMs fastest on this host
(containers often expose more logical CPUs than their CPU quota can feed).  Test infrastructure."""
import os
import time

import torch


def tune_threads(max_threads: int = 0) -> int:
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if max_threads:
        avail = min(avail, max_threads)
    a, w = torch.randn(388, 1024), torch.randn(4096, 1024)
    best_n, best_t = 1, float("inf")
    n = 1
    cands = []
    while n < avail:
        cands.append(n); n *= 2
    cands.append(avail)
    for n in cands:
        torch.set_num_threads(n)
        torch.nn.functional.linear(a, w)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.linear(a, w)
        dt = time.perf_counter() - t0
        if dt < best_t * 0.95:
            best_n, best_t = n, dt
    torch.set_num_threads(best_n)
    return best_n
