"""2..8 GPUs (torchrun): the fused step + gather (distributed.PeerGather: peer stores over NVLink) against the NCCL all-gather of the same
step (distributed.OverlappedGather) — gathered bytes must be identical, then both are timed.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/peer_gather_check.py [--envs 1048576] [--steps 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from gym_electric_motor_b200.distributed import OverlappedGather, PeerGather  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, K = a.envs, a.steps
e1, e2 = bench.make_env("pmsm", n, lr, rank), bench.make_env("pmsm", n, lr, rank)
e1.reset()
e2.reset()
gen = torch.Generator(device=dev).manual_seed(7 + rank)
pool = [torch.rand((n, 3), generator=gen, device=dev) * 2 - 1 for _ in range(8)]
og, pg = OverlappedGather(e1.sim, torch.float32), PeerGather(e2.sim, torch.float32)
ok = True
for k in range(6):  # correctness: the same step through both paths, gathered buffers byte for byte
    b1 = og.step(pool[k % 8])
    og.finish()
    b2 = pg.step(pool[k % 8])
    pg.finish()
    torch.cuda.synchronize()
    ref = og.bufs[b1].gathered if world > 1 else og.bufs[b1].local
    got = pg.buf[b2 * world * pg.nbytes:(b2 + 1) * world * pg.nbytes]
    same = bool(torch.equal(ref, got))
    ok = ok and same
pg.check()
t = torch.tensor([0 if ok else 1], device=dev)
if world > 1:
    dist.all_reduce(t)
all_ok = int(t.item()) == 0


def timed(fn, fin):
    for k in range(3):
        fn(pool[k % 8])
    fin()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        fn(pool[k % 8])
    fin()
    e1_.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1_)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / K


ms_nccl = timed(og.step, og.finish)
ms_peer = timed(pg.step, pg.finish)
pg.check()
if rank == 0:
    print(json.dumps({"world": world, "envs_per_gpu": n, "bytes_per_rank_per_step": pg.nbytes, "identical": all_ok, "ms_per_step_nccl_overlapped": ms_nccl,
                      "ms_per_step_peer_store": ms_peer, "env_steps_per_s_nccl": n * world / (ms_nccl * 1e-3), "env_steps_per_s_peer": n * world / (ms_peer * 1e-3),
                      "ingress_GBps_peer": (world - 1) * pg.nbytes / (ms_peer * 1e-3) / 1e9}))
og.release()
pg.release()
if world > 1:
    dist.destroy_process_group()
