"""NUMA placement of the HOST side of the host-buffer path (gemb200_step_host / gemb200_reset_host).

On an 8-GPU HGX box the GPUs hang off two CPU sockets (GPU 0-3 -> NUMA node 0, GPU 4-7 -> node 1).  The caller owns the host
buffers of `step_host`; when they sit on the other socket every D2H/H2D byte crosses the inter-socket fabric, which is what limited
the 8-GPU end-to-end rate in round 1 (8 x 72 MB per step).  Linux places the pages of a pinned allocation on the node the allocating
thread runs on, so a process that drives GPU d should run on d's node BEFORE it allocates its pinned buffers:

    hostmem.bind_to_device_numa_node(local_rank)      # CPU affinity + preferred memory node of this process
    buf = hostmem.pinned_empty((n, 14), torch.float32, local_rank)

Both are no-ops (returning None / plain pinned memory) on single-node machines or when sysfs does not expose the topology.
"""
import ctypes
import os

_bound = {}


def device_numa_node(device):
    """NUMA node of CUDA device ordinal `device` (respecting CUDA_VISIBLE_DEVICES through the PCI bus id), or None."""
    pci = None
    try:
        import torch

        p = torch.cuda.get_device_properties(int(device))
        if hasattr(p, "pci_domain_id"):
            pci = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        pci = None
    if pci is None:
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(device)
            if vis:
                ent = [v.strip() for v in vis.split(",") if v.strip()]
                if idx < len(ent) and ent[idx].isdigit():
                    idx = int(ent[idx])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            b = pynvml.nvmlDeviceGetPciInfo(h).busId
            b = b.decode() if isinstance(b, bytes) else b
            pci = b.lower()[-12:]  # nvml prints an 8-digit domain; sysfs uses 4
        except Exception:
            return None
    try:
        with open(f"/sys/bus/pci/devices/{pci}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def _node_cpus(node):
    with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
        txt = f.read().strip()
    cpus = set()
    for part in txt.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_device_numa_node(device):
    """Pin this process (all threads created afterwards) to the CPUs of the GPU's NUMA node and prefer that node for new pages.
    Returns the node, or None when there is nothing to do."""
    node = device_numa_node(device)
    if node is None:
        return None
    try:
        cpus = _node_cpus(node) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        return None
    try:  # set_mempolicy(MPOL_PREFERRED = 1, nodemask, maxnode): x86-64 syscall 238, aarch64 237
        libc = ctypes.CDLL(None, use_errno=True)
        nr = 238 if os.uname().machine == "x86_64" else 237
        mask = ctypes.c_ulong(1 << node)
        libc.syscall(ctypes.c_long(nr), ctypes.c_int(1), ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(ctypes.c_ulong)))
    except Exception:
        pass
    _bound[int(device)] = node
    return node


def pinned_empty(shape, dtype, device=0):
    """Page-locked host tensor for the host-buffer path; allocate it AFTER bind_to_device_numa_node(device) so that its pages are local
    to the GPU's socket (first touch happens here: the tensor is zero-filled once)."""
    import torch

    t = torch.empty(shape, dtype=dtype, pin_memory=True)
    t.zero_()
    return t
