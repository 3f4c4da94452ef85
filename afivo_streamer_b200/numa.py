"""Host-side placement for the transfers either side of the solver (afmg_upload / afmg_download): on a multi-socket
GPU node a page-locked buffer that lives on the other socket crosses the inter-socket link on every copy.  Binding the
calling thread to the cores of its GPU's NUMA node BEFORE it allocates (first touch) keeps buffers and copies local.
Pure host logic, no CUDA calls: the PCI address comes from torch's device properties, the rest from sysfs."""
import os
from typing import Optional, Set


def parse_cpulist(text: str) -> Set[int]:
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the format of /sys/devices/system/node/node*/cpulist)"""
    cpus: Set[int] = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def pci_address(device_index: int) -> Optional[str]:
    """'0000:1b:00.0' of a CUDA device (as CUDA numbers them in this process), or None"""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:  # noqa: BLE001  (older torch, no device)
        return None


def gpu_numa_node(device_index: int, sysfs: str = "/sys") -> Optional[int]:
    """NUMA node the GPU hangs off, None when unknown or the machine has a single node"""
    addr = pci_address(device_index)
    if addr is None:
        return None
    try:
        with open(os.path.join(sysfs, "bus/pci/devices", addr, "numa_node")) as f:
            node = int(f.read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node: int, sysfs: str = "/sys") -> Set[int]:
    try:
        with open(os.path.join(sysfs, f"devices/system/node/node{node}/cpulist")) as f:
            return parse_cpulist(f.read())
    except OSError:
        return set()


def bind_to_gpu_node(device_index: int, sysfs: str = "/sys") -> Optional[int]:
    """Restrict the calling thread (and the threads it starts later) to the cores of the GPU's NUMA node that it is
    already allowed on.  Returns the node, or None if nothing was changed (unknown topology, no allowed core there)."""
    node = gpu_numa_node(device_index, sysfs)
    if node is None:
        return None
    allowed = os.sched_getaffinity(0)
    local = allowed & node_cpus(node, sysfs)
    if not local or local == allowed:
        return node if local else None
    try:
        os.sched_setaffinity(0, local)
    except OSError:
        return None
    return node
