"""Process placement for the host-facing (e2e) paths: bind the calling process to the CPUs of the NUMA node its GPU hangs
off, BEFORE pinned host buffers are allocated (first touch places them on that node), so that the H2D / D2H streams of the
ranks of one box do not all cross the socket interconnect.  Linux sysfs only; silently a no-op where the information is
missing (containers without /sys/bus/pci, single-node hosts)."""
from __future__ import annotations

import os


def _read(path: str) -> str | None:
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s: str) -> list[int]:
    cpus: list[int] = []
    for part in s.split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(device: int) -> int | None:
    """NUMA node of CUDA device `device` (torch ordinal), from its PCI bus id."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None
    v = _read(f"/sys/bus/pci/devices/{bus}/numa_node")
    if v is None or not v.lstrip("-").isdigit() or int(v) < 0:
        return None
    return int(v)


def bind_to_gpu_node(device: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns what was done (for the bench record)."""
    info = {"node": None, "cpus": len(os.sched_getaffinity(0)), "bound": False}
    node = gpu_numa_node(device)
    if node is None:
        return info
    info["node"] = node
    cl = _read(f"/sys/devices/system/node/node{node}/cpulist")
    if not cl:
        return info
    allowed = set(_parse_cpulist(cl)) & os.sched_getaffinity(0)
    if not allowed:
        return info
    try:
        os.sched_setaffinity(0, allowed)
    except OSError:
        return info
    info.update(cpus=len(allowed), bound=True)
    return info
