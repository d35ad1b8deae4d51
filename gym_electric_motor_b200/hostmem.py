"""NUMA placement of the HOST side of the host-buffer path (gemb200_step_host / gemb200_reset_host).

On an 8-GPU HGX box the GPUs hang off two CPU sockets (GPU 0-3 -> NUMA node 0, GPU 4-7 -> node 1).  The caller owns the host
buffers of `step_host`; when they sit on the other socket every D2H/H2D byte crosses the inter-socket fabric, which is what limited
the 8-GPU end-to-end rate in round 1 (8 x 72 MB per step).  Linux places the pages of a pinned allocation on the node the allocating
thread runs on, so a process that drives GPU d should run on d's node BEFORE it allocates its pinned buffers:

    hostmem.bind_to_device_numa_node(local_rank)      # CPU affinity + preferred memory node of this process
    buf = hostmem.pinned_empty((n, 14), torch.float32, local_rank)

Where sysfs does not expose the GPU's node (containers often read -1) the node is MEASURED (probe_numa_node: device-to-host copy rate into
page-locked memory of every node).  Both are no-ops (returning None / plain pinned memory) on single-node machines.
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


def _set_preferred_node(node):
    """set_mempolicy(MPOL_PREFERRED = 1, nodemask, maxnode) — x86-64 syscall 238, aarch64 237; node None: back to the default policy"""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        nr = 238 if os.uname().machine == "x86_64" else 237
        if node is None:
            libc.syscall(ctypes.c_long(nr), ctypes.c_int(0), ctypes.c_void_p(0), ctypes.c_ulong(0))
        else:
            mask = ctypes.c_ulong(1 << node)
            libc.syscall(ctypes.c_long(nr), ctypes.c_int(1), ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(ctypes.c_ulong)))
    except Exception:
        pass


def probe_numa_node(device, mbytes=48):
    """When sysfs does not say which NUMA node a GPU hangs off (containers often read -1): MEASURE it.  For every node with CPUs this
    process may run on, touch a buffer from a thread bound to that node, page-lock it (cudaHostRegister) and time device-to-host copies
    into it; the node with the highest rate is the GPU's.  Returns (node or None, {node: GB/s}).  ~0.1 s; leaves affinity and memory
    policy as they were."""
    try:
        import numpy as np
        import torch

        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except Exception:
        return None, {}
    if len(nodes) < 2:
        return None, {}
    import time

    saved = os.sched_getaffinity(0)
    rates = {}
    try:
        rt = torch.cuda.cudart()
        nbytes = int(mbytes) << 20
        src = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{int(device)}")
        for nd in nodes:
            try:
                cpus = _node_cpus(nd) & saved
            except Exception:
                continue
            if not cpus:
                continue
            os.sched_setaffinity(0, cpus)
            _set_preferred_node(nd)
            arr = np.empty(nbytes, dtype=np.uint8)
            arr[:] = 0  # first touch: the pages land on node nd
            if int(rt.cudaHostRegister(arr.ctypes.data, nbytes, 0)) != 0:
                continue
            try:
                dst = torch.from_numpy(arr)
                for _ in range(2):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize(int(device))
                t0 = time.perf_counter()
                for _ in range(4):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize(int(device))
                rates[nd] = 4 * nbytes / (time.perf_counter() - t0) / 1e9
            finally:
                rt.cudaHostUnregister(arr.ctypes.data)
    except Exception:
        pass
    finally:
        os.sched_setaffinity(0, saved)
        _set_preferred_node(None)
    if len(rates) < 2:
        return None, rates
    best = max(rates, key=rates.get)
    worst = min(rates.values())
    return (best if rates[best] > 1.08 * worst else None), rates  # within 8 %: no preference to speak of


placement = {}  # device -> {"node": ..., "how": "sysfs" | "probed" | None, "d2h_GBps_by_node": {...}}: what bind_to_device_numa_node found


def bind_to_device_numa_node(device, probe=True):
    """Pin this process (all threads created afterwards) to the CPUs of the GPU's NUMA node and prefer that node for new pages.
    The node comes from sysfs, or — when sysfs does not tell and `probe` — from a measurement (probe_numa_node).
    Returns the node, or None when there is nothing to do."""
    node, how, rates = device_numa_node(device), "sysfs", {}
    if node is None and probe:
        node, rates = probe_numa_node(device)
        how = "probed"
    placement[int(device)] = {"node": node, "how": how if node is not None else None, "d2h_GBps_by_node": {str(k): round(v, 1) for k, v in rates.items()}}
    if node is None:
        return None
    try:
        cpus = _node_cpus(node) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        return None
    _set_preferred_node(node)
    _bound[int(device)] = node
    return node


def pcie_link(device):
    """{"gen": current PCIe generation, "width": lanes} of the GPU's link (NVML), or {} — context for the host-buffer rates"""
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
        return {"gen": int(pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h)), "width": int(pynvml.nvmlDeviceGetCurrPcieLinkWidth(h)),
                "max_gen": int(pynvml.nvmlDeviceGetMaxPcieLinkGeneration(h)), "max_width": int(pynvml.nvmlDeviceGetMaxPcieLinkWidth(h))}
    except Exception:
        return {}


def pinned_empty(shape, dtype, device=0):
    """Page-locked host tensor for the host-buffer path; allocate it AFTER bind_to_device_numa_node(device) so that its pages are local
    to the GPU's socket (first touch happens here: the tensor is zero-filled once)."""
    import torch

    t = torch.empty(shape, dtype=dtype, pin_memory=True)
    t.zero_()
    return t
