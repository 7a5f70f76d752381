"""Host-side placement helpers for the multi-GPU data-prep path (plumbing, no compute).

With one process per GPU the host buffers of the PCIe-bound host pipeline should live on the NUMA node the
GPU hangs off: page-locked memory is placed by first touch, so binding the process to the GPU-local CPUs
before allocating is enough."""
from __future__ import annotations

import os
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_local_cpus(device: int) -> Optional[List[int]]:
    """CPUs local to CUDA device `device` according to sysfs, or None when it cannot be told."""
    try:
        import torch

        props = torch.cuda.get_device_properties(device)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa(device: int) -> Optional[List[int]]:
    """Restrict this process to the CPUs local to `device` (intersected with its current affinity).
    Returns the CPU list applied, or None if nothing was changed."""
    cpus = gpu_local_cpus(device)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return None
    allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
    if not allowed:
        return None
    os.sched_setaffinity(0, allowed)
    return allowed
