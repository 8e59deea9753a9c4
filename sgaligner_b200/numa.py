"""NUMA placement of a rank: run on the CPUs of the GPU's NUMA node and allocate the pinned staging pool there.

The host-to-host serving step is PCIe bound; with 8 ranks on one box all pinned buffers landed on NUMA node 0
(round-1 SCALE record: every rank on CPUs 0-31, per-GPU H2D 54 -> 25 GB/s).  ``bind_to_gpu_node`` is called once per
process BEFORE any pinned allocation: Linux places pages on the node of the touching CPU (first touch), so confining
the process to the GPU's node is enough to make ``cudaHostAlloc`` memory local to the GPU's PCIe root.
No external library: sysfs + ``os.sched_setaffinity``.  Everything is best effort -- on a box without NUMA
information nothing changes and the returned dict says why.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional


def _parse_cpulist(txt: str) -> List[int]:
    cpus: List[int] = []
    for part in txt.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_pci_address(device_index: int) -> Optional[str]:
    """``domain:bus:device.0`` of a CUDA device (from torch's device properties)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        return '%04x:%02x:%02x.0' % (int(p.pci_domain_id), int(p.pci_bus_id), int(p.pci_device_id))
    except Exception:   # noqa: BLE001
        return None


def gpu_numa_node(device_index: int) -> Optional[int]:
    addr = gpu_pci_address(device_index)
    if addr is None:
        return None
    try:
        with open(f'/sys/bus/pci/devices/{addr}/numa_node') as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:   # noqa: BLE001
        return None


def node_cpus(node: int) -> List[int]:
    try:
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            return _parse_cpulist(f.read())
    except Exception:   # noqa: BLE001
        return []


def bind_to_gpu_node(device_index: int, local_rank: int = 0, ranks_on_node: int = 1) -> Dict:
    """Confine this process to the CPUs of the NUMA node the GPU hangs off (intersected with the CPUs the process is
    allowed to use).  When several ranks share a node each gets its own slice of the node's CPUs so that their
    collation threads do not fight.  Returns what was done (for the bench line)."""
    info: Dict = {'gpu': device_index, 'pci': gpu_pci_address(device_index), 'numa_node': None, 'bound': False}
    node = gpu_numa_node(device_index)
    if node is None:
        info['why'] = 'no NUMA node in sysfs for this device'
        return info
    info['numa_node'] = node
    allowed = set(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else set()
    cpus = [c for c in node_cpus(node) if not allowed or c in allowed]
    if not cpus:
        info['why'] = 'the node has no CPU this process may use'
        return info
    if ranks_on_node > 1 and len(cpus) >= 2 * ranks_on_node:
        per = len(cpus) // ranks_on_node
        k = local_rank % ranks_on_node
        cpus = cpus[k * per:(k + 1) * per]
    try:
        os.sched_setaffinity(0, cpus)
        info['bound'] = True
        info['cpus'] = '%d-%d (%d)' % (min(cpus), max(cpus), len(cpus))
    except Exception as e:   # noqa: BLE001
        info['why'] = f'sched_setaffinity failed: {e}'
    return info
