"""2-GPU tests of the sharded env (needs >= 2 CUDA devices; skipped on a 1-GPU box): the global batch returned by
every rank -- gathered over peer memory by the step kernel's epilogue (fused) or by one NCCL all-gather (fallback) --
equals, bit for bit, the rows of ONE single-GPU env of the same global batch and seeds."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, task, mode, steps, global_batch, q):
    try:
        sys.path.insert(0, ROOT)
        import contextlib
        import io
        import torch.distributed as dist
        from pybullet_multigoal_gym_b200.sharded import ShardedKukaEnv
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        kw = dict(num_block=3, check_actions=False, max_episode_steps=3)
        if mode == "auto":
            kw.update(device_sampling=True, auto_reset=True)
        with contextlib.redirect_stdout(io.StringIO()):
            env = ShardedKukaEnv(task, global_batch, seed=11, device=rank, fused=(mode != "nccl"), **kw)
        assert env.fused == (mode != "nccl")
        gen = torch.Generator(device="cuda")
        gen.manual_seed(5)  # the same global action tape on every rank
        rows = []
        for t in range(steps):
            a = torch.rand((global_batch, env.env.action_dim), device="cuda", generator=gen) * 2 - 1
            obs, reward, done, ok = env.step_gathered(env.local_actions(a).contiguous())
            rows.append((obs.cpu().numpy().copy(), reward.cpu().numpy().copy(), done.cpu().numpy().copy(), ok.cpu().numpy().copy()))
        host = env.step_host(env.local_actions(a).cpu().numpy()) if mode != "nccl" else None
        if host is not None:
            rows.append(tuple(np.array(x) for x in host))
        launches = env.env.launch_count
        dist.barrier()
        q.put((rank, rows, launches))
        dist.destroy_process_group()
    except Exception as e:  # surface the failure instead of a queue timeout
        import traceback
        q.put((rank, "ERROR: %s\n%s" % (e, traceback.format_exc()), 0))


@pytest.mark.parametrize("task,mode", [("reach", "fused"), ("pick_and_place", "fused"), ("block_stack", "fused"),
                                       ("reach", "auto"), ("push", "auto"), ("reach", "nccl")])
def test_sharded_rows_equal_single_gpu_rows(task, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import contextlib
    import io
    import torch.multiprocessing as mp
    import pybullet_multigoal_gym_b200 as pmg
    world, steps, Bg = 2, 7, 48
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, task, mode, steps, Bg, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert not isinstance(r[1], str), r[1]
    # the single-GPU twin
    kw = dict(num_block=3, check_actions=False, max_episode_steps=3)
    if mode == "auto":
        kw.update(device_sampling=True, auto_reset=True)
    with contextlib.redirect_stdout(io.StringIO()):
        env = pmg.make_env(task=task, batch=Bg, seed=11, **kw)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    want = []
    nsteps = steps + (1 if mode != "nccl" else 0)
    for t in range(nsteps):
        if t < steps:
            a = torch.rand((Bg, env.action_dim), device="cuda", generator=gen) * 2 - 1
        obs, reward, done, info = env.step(a)   # the extra host-path step repeats the last action
        packed = torch.cat([obs[k] for k in ("observation", "policy_state", "achieved_goal", "desired_goal")], dim=1)
        want.append((packed.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), info["goal_achieved"].cpu().numpy()))
    for rank, rows, launches in results:
        assert len(rows) == nsteps
        for t in range(nsteps):
            for got, ref in zip(rows[t], want[t]):
                assert np.array_equal(got, ref), (task, mode, rank, t)
        if mode == "fused":
            assert launches == 2 + nsteps     # ctor state init + ctor reset + ONE kernel per step: the gather is in it
