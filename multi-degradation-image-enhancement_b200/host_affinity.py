"""NUMA placement of a rank's host side (one process per GPU).

`cdan_forward_host` / `cdan_forward_host_u8` stream pinned host buffers over PCIe while the forward runs.  With eight ranks
on one box the copies of all GPUs otherwise come out of whichever NUMA node the processes happened to start on (round 1:
e2e efficiency 0.42 at 8 GPUs).  `bind_to_gpu_node` pins the calling process to the CPUs of the NUMA node the GPU's PCIe
root port hangs off, BEFORE the pinned buffers are allocated, so that first-touch places them in node-local memory.
Everything is read from sysfs; if anything is missing the function reports it and changes nothing.
"""
from __future__ import annotations

import os
from typing import Dict, Optional


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def gpu_cpu_affinity_from_topo(device_index: int):
    """Fallback when sysfs carries no NUMA node for the GPU's PCI function (containers often show -1): the CPU-affinity column
    of `nvidia-smi topo -m`.  Returns (cpu set, numa node or None) or (None, None)."""
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None, None
    for line in out.splitlines():
        fields = [f.strip() for f in re.split(r"\t+|\s{2,}", line.strip()) if f.strip()]
        if not fields or fields[0] != f"GPU{device_index}":
            continue
        for k, f in enumerate(fields[1:], start=1):
            if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", f) and ("-" in f or "," in f):
                node = int(fields[k + 1]) if k + 1 < len(fields) and fields[k + 1].isdigit() else None
                return _parse_cpulist(f), node
    return None, None


def bind_to_gpu_node(device_index: int) -> Dict[str, object]:
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns a small report for the bench JSON line."""
    node = gpu_numa_node(device_index)
    report: Dict[str, object] = {"gpu": device_index, "numa_node": node, "bound": False}
    if node is None:
        cpus, node = gpu_cpu_affinity_from_topo(device_index)
        report["numa_node"] = node
        if cpus:
            try:
                target = cpus & os.sched_getaffinity(0)
                if target and len(target) < len(os.sched_getaffinity(0)):
                    os.sched_setaffinity(0, target)
                    report["bound"] = True
                    report["cpus"] = len(target)
                    report["source"] = "nvidia-smi topo -m"
                else:
                    report["note"] = "GPU affinity covers all allowed CPUs (single NUMA node visible)"
            except Exception as exc:  # pragma: no cover - depends on the host
                report["error"] = str(exc)
        return report
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target:
            os.sched_setaffinity(0, target)
            report["bound"] = True
            report["cpus"] = len(target)
    except Exception as exc:  # pragma: no cover - depends on the host
        report["error"] = str(exc)
    return report
