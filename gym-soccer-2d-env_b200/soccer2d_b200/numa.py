"""NUMA placement of a rank's host side: the process and its pinned staging buffers next to its GPU.

The host-buffer path (s2d_submit_host / s2d_wait_host) moves 46 bytes per env and launch to the host; with one process
per GPU on a two-socket box the copies of half the GPUs cross the inter-socket link unless the pinned buffers live on
the GPU's own NUMA node.  Everything here is best effort and reports what it did (containers often hide the topology
or restrict the cpuset): nothing fails when the platform says no.

Linux only; uses sysfs, sched_setaffinity and the set_mempolicy / mbind system calls through ctypes (no libnuma).
"""
from __future__ import annotations

import ctypes as C
import mmap
import os

import numpy as np
import torch

_SYS_set_mempolicy, _SYS_mbind = 238, 237  # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2
_libc = C.CDLL(None, use_errno=True)


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(text: str) -> set:
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11}"""
    out = set()
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


def online_nodes() -> list:
    return sorted(parse_cpulist(_read("/sys/devices/system/node/online") or ""))


def gpu_pci_bdf(index: int) -> str | None:
    """'0000:1b:00.0' of CUDA device `index` (torch device properties, else NVML)"""
    try:
        p = torch.cuda.get_device_properties(index)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:  # noqa: BLE001
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByUUID(uuid)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        return bus.lower()[-12:]
    except Exception:  # noqa: BLE001
        return None


def gpu_numa_node(index: int) -> int | None:
    """NUMA node the GPU hangs off (sysfs), None when the platform does not say (-1 in most VMs)"""
    bdf = gpu_pci_bdf(index)
    if not bdf:
        return None
    text = _read(f"/sys/bus/pci/devices/{bdf}/numa_node")
    try:
        node = int(text)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def _nodemask(node: int):
    words = node // 64 + 1
    mask = (C.c_ulong * words)()
    mask[node // 64] = 1 << (node % 64)
    return mask, words * 64 + 1


def bind_process_to_gpu(index: int) -> dict:
    """Pin this process to the CPUs of the GPU's NUMA node (those the cpuset allows) and make that node the preferred
    one for its memory.  Call before allocating pinned buffers.  Returns a report for the bench line."""
    rep = {"gpu": index, "pci": gpu_pci_bdf(index), "node": gpu_numa_node(index), "nodes_online": online_nodes(),
           "cpus_before": len(os.sched_getaffinity(0)), "bound_cpus": False, "mempolicy": False}
    node = rep["node"]
    if node is None:
        rep["why"] = "the platform does not expose the GPU's NUMA node (sysfs numa_node missing or -1)"
        return rep
    cpus = parse_cpulist(_read(f"/sys/devices/system/node/node{node}/cpulist")) & os.sched_getaffinity(0)
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
            rep["bound_cpus"] = True
            rep["cpus_after"] = len(cpus)
        except OSError as e:
            rep["why"] = f"sched_setaffinity: {e}"
    else:
        rep["why"] = f"none of node {node}'s CPUs is in this process's cpuset"
    mask, maxnode = _nodemask(node)
    if _libc.syscall(_SYS_set_mempolicy, MPOL_PREFERRED, C.byref(mask), maxnode) == 0:
        rep["mempolicy"] = True
    else:
        rep["why_mempolicy"] = os.strerror(C.get_errno())
    return rep


class PinnedBlock:
    """`nbytes` of page-locked host memory on NUMA node `node` (mmap + mbind + first touch + cudaHostRegister), as a
    uint8 torch tensor in `.tensor`.  Falls back to torch's pinned allocator (node = wherever the driver puts it)."""

    def __init__(self, nbytes: int, node: int | None = None):
        self.node, self.registered, self._map = None, False, None
        nbytes = int(nbytes)
        if node is not None:
            try:
                size = (nbytes + mmap.PAGESIZE - 1) // mmap.PAGESIZE * mmap.PAGESIZE
                self._map = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
                arr = np.frombuffer(self._map, dtype=np.uint8)
                mask, maxnode = _nodemask(node)
                if _libc.syscall(_SYS_mbind, C.c_void_p(arr.ctypes.data), C.c_ulong(size), MPOL_BIND, C.byref(mask),
                                 C.c_ulong(maxnode), 0) != 0:
                    raise OSError(C.get_errno(), "mbind")
                arr[::mmap.PAGESIZE] = 0  # first touch: the pages materialise on `node`
                rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, size, 0)
                if int(rc) != 0:
                    raise OSError(int(rc), "cudaHostRegister")
                self._ptr, self.registered, self.node = arr.ctypes.data, True, node
                self.tensor = torch.from_numpy(arr)[:nbytes]
                return
            except Exception:  # noqa: BLE001 - fall back to the plain pinned allocator
                self._map = None
        self.tensor = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self._ptr)
            self.registered = False
