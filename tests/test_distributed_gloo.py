"""N>1 host logic on CPU: world_size-2 gloo run of the sharding helpers, plus world-size independence of the env
streams (checked with the CPU oracle, which shares the device's RNG keying by global env index)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from helpers import ROOT, config_from_meta, load_golden
from gym_electric_motor_b200 import _cabi as K
from gym_electric_motor_b200.distributed import shard_envs


def test_shard_partition_is_exact():
    for total in (1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            parts = [shard_envs(total, r, world) for r in range(world)]
            assert sum(c for c, _ in parts) == total
            off = 0
            for c, o in parts:
                assert o == off
                off += c
            assert max(c for c, _ in parts) - min(c for c, _ in parts) <= 1


def test_results_do_not_depend_on_world_size(oracle_lib):
    g = load_golden("pmsm_cc_rk4")
    total, steps = 64, 30
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, size=(steps, total, 3))

    def run(count, offset):
        cfg = config_from_meta(g["meta"], n_envs=count, reset_ode=g["reset_ode"], solver="rk4", ref_kind=K.REF_WIENER,
                               autoreset=K.AUTORESET_SAME_STEP, seed=21)
        cfg.env_index_offset = offset
        sim = oracle_lib.Oracle(cfg)
        out = [np.concatenate(sim.reset(), axis=1)]
        for k in range(steps):
            o, r, w, t = sim.step(acts[k, offset : offset + count])
            out.append(np.concatenate([o, r, w[:, None], t[:, None]], axis=1))
        return out

    whole = run(total, 0)
    parts = [run(*shard_envs(total, r, 2)) for r in range(2)]
    assert np.array_equal(np.concatenate([parts[0][0], parts[1][0]]), whole[0])
    for k in range(1, steps + 1):
        assert np.array_equal(np.concatenate([parts[0][k], parts[1][k]]), whole[k])


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from gym_electric_motor_b200.distributed import shard_envs, rank_world, global_stats, all_gather_batch, PackedStepOutputs
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
    rank, world = rank_world()
    assert (rank, world) == (int(os.environ["RANK"]), 2)
    count, offset = shard_envs(10, rank, world)
    reward = torch.arange(offset, offset + count, dtype=torch.float32)
    term = (reward %% 3 == 0)
    mean, n_term = global_stats(reward, term)
    assert abs(mean - 4.5) < 1e-12 and n_term == 4, (mean, n_term)
    (g,) = all_gather_batch(reward.reshape(-1, 1))
    assert g.flatten().tolist() == [float(i) for i in range(10)]
    # the single packed all-gather of (obs, ref, reward, terminated): sections 16-byte aligned, rank-major result
    n_loc, n_state, n_ref = 5, 14, 2
    out = PackedStepOutputs(n_loc, n_state, n_ref, torch.float32, torch.device("cpu"))
    obs, ref, rew, trm = out.local_views()
    assert all(t.data_ptr() %% 16 == 0 for t in (obs, ref, rew, trm)) and obs.shape == (5, 14) and trm.dtype == torch.uint8
    obs.copy_(torch.arange(n_loc * n_state, dtype=torch.float32).reshape(n_loc, n_state) + 1000 * rank)
    ref.fill_(rank + 0.5); rew.copy_(torch.arange(n_loc, dtype=torch.float32) - rank); trm.fill_(rank)
    g_obs, g_ref, g_rew, g_trm = out.gather()
    assert g_obs.shape == (2, 5, 14) and g_trm.shape == (2, 5)
    for r in range(2):
        assert torch.equal(g_obs[r], torch.arange(n_loc * n_state, dtype=torch.float32).reshape(n_loc, n_state) + 1000 * r)
        assert (g_ref[r] == r + 0.5).all() and torch.equal(g_rew[r], torch.arange(n_loc, dtype=torch.float32) - r) and (g_trm[r] == r).all()
    dist.destroy_process_group()
    print("ok", rank)
""") % ROOT


def test_two_rank_gloo_collectives(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port, MASTER_ADDR="127.0.0.1"),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_host_placement_helpers_degrade_quietly(monkeypatch):
    """hostmem: on a machine with one NUMA node (or without sysfs / NVML / CUDA) there is nothing to bind and nothing to probe — the helpers
    return None / {} and leave the process's CPU affinity alone; with a fake two-node sysfs answer the binding narrows the affinity to that
    node's CPUs (the e2e arm of bench.py allocates its pinned buffers after this call)."""
    import os

    from gym_electric_motor_b200 import hostmem

    before = os.sched_getaffinity(0)
    nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()] if os.path.isdir("/sys/devices/system/node") else []
    if len(nodes) < 2:
        assert hostmem.probe_numa_node(0) == (None, {})
    assert isinstance(hostmem.pcie_link(0), dict)
    monkeypatch.setattr(hostmem, "device_numa_node", lambda d: None)
    monkeypatch.setattr(hostmem, "probe_numa_node", lambda d: (None, {}))
    assert hostmem.bind_to_device_numa_node(0) is None and hostmem.placement[0]["node"] is None
    assert os.sched_getaffinity(0) == before
    some = sorted(before)[: max(1, len(before) // 2)]
    monkeypatch.setattr(hostmem, "device_numa_node", lambda d: 0)
    monkeypatch.setattr(hostmem, "_node_cpus", lambda n: set(some))
    try:
        assert hostmem.bind_to_device_numa_node(0) == 0
        assert os.sched_getaffinity(0) == set(some) and hostmem.placement[0] == {"node": 0, "how": "sysfs", "d2h_GBps_by_node": {}}
    finally:
        os.sched_setaffinity(0, before)
        hostmem._set_preferred_node(None)
