"""Host thread selection for the CPU oracle.  Test infrastructure.

``physical_threads()`` is the DETERMINISTIC choice used by every timed CPU leg (bench.py ``--impl reference`` and
``cpu_baseline``): one thread per physical core of ONE NUMA node, intersected with the process's affinity mask, capped at
64 (the oracle's batch-of-one GEMMs, M = 388, stop scaling there and cross-node traffic makes them slower).  The previous
chooser timed three small GEMMs and picked 32 or 64 depending on noise, which moved the reported CPU RTFx by 5x between
runs on identical boxes.  ``tune_threads()`` (measure and pick) remains for the pytest session, where only wall time matters.
"""
import os
import time
from typing import Dict, List, Tuple

import torch


def _parse_cpulist(text: str) -> List[int]:
    out: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out


def _topology() -> Tuple[Dict[int, Tuple[int, int]], Dict[int, List[int]]]:
    """cpu -> (package, core id) and node -> cpus, from sysfs; empty dicts when sysfs is not readable."""
    cores: Dict[int, Tuple[int, int]] = {}
    nodes: Dict[int, List[int]] = {}
    base = "/sys/devices/system"
    try:
        for name in os.listdir(f"{base}/cpu"):
            if name.startswith("cpu") and name[3:].isdigit():
                c = int(name[3:])
                try:
                    pkg = int(open(f"{base}/cpu/{name}/topology/physical_package_id").read())
                    cid = int(open(f"{base}/cpu/{name}/topology/core_id").read())
                    cores[c] = (pkg, cid)
                except (OSError, ValueError):
                    pass
        for name in os.listdir(f"{base}/node"):
            if name.startswith("node") and name[4:].isdigit():
                try:
                    nodes[int(name[4:])] = _parse_cpulist(open(f"{base}/node/{name}/cpulist").read())
                except (OSError, ValueError):
                    pass
    except OSError:
        pass
    return cores, nodes


def physical_cpus(cap: int = 64) -> Tuple[List[int], str]:
    """One logical CPU per physical core of the NUMA node that holds most of this process's CPUs; returns (cpus, how)."""
    allowed = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    cores, nodes = _topology()
    pool, how = allowed, "affinity mask"
    if nodes:
        best = max(nodes, key=lambda n: (len(set(nodes[n]) & set(allowed)), -n))
        inter = [c for c in allowed if c in set(nodes[best])]
        if inter:
            pool, how = inter, f"NUMA node {best}"
    seen, cpus = set(), []
    for c in pool:
        key = cores.get(c, (0, c))
        if key not in seen:
            seen.add(key)
            cpus.append(c)
    cpus = cpus[:cap]
    return cpus, f"{len(cpus)} physical cores of {how}" + (f" (capped at {cap})" if len(cpus) == cap else "")


def physical_threads(cap: int = 64, pin: bool = False) -> Tuple[int, str]:
    """Set torch's intra-op threads to the physical-core count (see module docstring); ``pin`` also restricts the process
    to those CPUs (the reference arm, which does nothing else)."""
    cpus, how = physical_cpus(cap)
    torch.set_num_threads(max(len(cpus), 1))
    if pin and hasattr(os, "sched_setaffinity"):
        try:
            os.sched_setaffinity(0, cpus)
            how += ", pinned"
        except OSError:
            pass
    return max(len(cpus), 1), how


def tune_threads(max_threads: int = 0) -> int:
    """Measure-and-pick (powers of two up to the affinity mask): for the test session only, never for a reported number."""
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
