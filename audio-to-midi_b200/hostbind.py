"""Host-side placement for the host-fed paths (a2m_submit_host_ex, TrainEngine.train_pipelined).

One process drives one GPU.  Its page-locked staging buffers should live on the NUMA node the GPU hangs off: a pinned buffer on
the other socket crosses the inter-socket link on every H2D / D2H copy, and with eight ranks feeding ~20 GB/s each that link,
not PCIe, is what saturates (round 1: end-to-end efficiency 0.62 at 8 GPUs with every rank's buffers on node 0).

bind_to_gpu_numa(device) does two things, each best-effort and silent on failure (containers often forbid one or the other):
  * sched_setaffinity to the CPUs of the GPU's NUMA node -- only if the process is allowed to run on some of them;
  * set_mempolicy(MPOL_PREFERRED, node) so that pages touched / page-locked from now on come from that node even when the
    CPUs of that node are not available to the container.
Call it before the first pinned allocation (model.pinned_empty, torch .pin_memory()).
"""
from __future__ import annotations

import ctypes
import os

_MPOL_PREFERRED = 1
_SYS_set_mempolicy = 238   # x86_64


def gpu_numa_node(device: int):
    """NUMA node of CUDA device `device` from sysfs (PCI bus id via torch), or None when unknown / single-node."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def _node_cpus(node: int):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            text = f.read().strip()
    except OSError:
        return set()
    cpus = set()
    for part in text.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa(device: int) -> dict:
    """Returns what was done: {"node": n or None, "cpus": k bound CPUs or 0, "mempolicy": bool}."""
    done = {"node": None, "cpus": 0, "mempolicy": False}
    node = gpu_numa_node(device)
    if node is None:
        return done
    done["node"] = node
    try:
        allowed = os.sched_getaffinity(0)
        want = _node_cpus(node) & allowed
        if want and want != allowed:
            os.sched_setaffinity(0, want)
            done["cpus"] = len(want)
        elif want:
            done["cpus"] = len(want)
    except OSError:
        pass
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        nbits = 1024
        mask = (ctypes.c_ulong * (nbits // (8 * ctypes.sizeof(ctypes.c_ulong))))()
        mask[node // (8 * ctypes.sizeof(ctypes.c_ulong))] |= 1 << (node % (8 * ctypes.sizeof(ctypes.c_ulong)))
        rc = libc.syscall(_SYS_set_mempolicy, _MPOL_PREFERRED, ctypes.byref(mask), nbits + 1)
        done["mempolicy"] = rc == 0
    except Exception:
        pass
    return done
