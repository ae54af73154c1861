"""Host-side placement for the host-buffer entry points (ballenv_step_host / ballenv_step_many_host).

The reference has nothing like this (it is a single CPU process); it exists because the end-to-end path of the vector
environment is bounded by the copies between pinned host memory and the GPU, and on a multi-socket box those copies
only reach the PCIe rate when the pinned pages and the issuing thread sit on the NUMA node the GPU hangs off.
``torchrun`` does not bind its workers, so a rank that wants the full rate calls ``bind_to_gpu_numa`` BEFORE it
allocates pinned memory: the pages are then first-touched (and pinned) on the right node.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional


def gpu_cpu_affinity(device_index: int) -> Optional[List[int]]:
    """CPUs NVML reports as local to the GPU (``nvmlDeviceGetCpuAffinity``), or None if NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        return [c for c in cpus if c < ncpu] or None
    except Exception:
        return None


def bind_to_gpu_numa(device_index: int) -> Dict[str, object]:
    """Restrict this process to the CPUs local to GPU ``device_index`` (intersected with the CPUs it may already run
    on).  Returns what was done, for logging: {"bound": bool, "cpus": n, "of": m, "why": ...}."""
    info: Dict[str, object] = {"bound": False, "cpus": None, "of": None, "why": None}
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:   # pragma: no cover  (non-Linux)
        info["why"] = "sched_getaffinity unavailable"
        return info
    info["of"] = len(allowed)
    local = gpu_cpu_affinity(device_index)
    if not local:
        info["why"] = "NVML reports no CPU affinity"
        return info
    want = sorted(set(local) & set(allowed))
    if not want:
        info["why"] = "no overlap between the GPU's CPUs and the allowed set"
        return info
    if len(want) == len(allowed):
        info.update(bound=False, cpus=len(want), why="single NUMA node (or already bound)")
        return info
    try:
        os.sched_setaffinity(0, want)
        info.update(bound=True, cpus=len(want))
    except OSError as e:   # pragma: no cover
        info["why"] = repr(e)
    return info


def pinned_empty(shape, dtype):
    """A pinned (page-locked) CPU tensor; call after bind_to_gpu_numa so that its pages land on the GPU's node."""
    import torch
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def copy_ceiling_gbs(device, nbytes: int, iters: int = 10, direction: str = "d2h") -> float:
    """Bare ``cudaMemcpyAsync`` rate between a device buffer and a pinned host buffer of ``nbytes`` (GB/s, CUDA events
    on the current stream): the ceiling the host-buffer entry points can reach on this box."""
    import torch
    dev = torch.device(device)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = pinned_empty((nbytes,), torch.uint8)
    src, dst = (d, h) if direction == "d2h" else (h, d)
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
