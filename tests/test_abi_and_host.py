"""CPU tests: the C-ABI library loads and exports every symbol include/pmg.h declares (no compute
calls without a GPU), and the host-side logic (spaces, seeding, make_env validation, sharding)."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from pybullet_multigoal_gym_b200 import build
    build.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    from pybullet_multigoal_gym_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pmg.h")).read()
    declared = sorted(set(re.findall(r"\b(pmg_[a-z_]+)\s*\(", header)))
    assert declared, "no declarations found in include/pmg.h"
    import ctypes
    L = ctypes.CDLL(_lib.SO_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, "libpmg.so lacks: %s" % missing
    assert sorted(_lib.SYMBOLS) == declared
    assert _lib.load().pmg_abi_version() == _lib.ABI_VERSION == 6
    # the library is sm_100a SASS produced by our own sources (kept in-tree, not in site-packages)
    assert os.path.dirname(_lib.SO_PATH) == os.path.join(ROOT, "pybullet_multigoal_gym_b200")


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu():
    _ensure_built()
    import ctypes as C
    from pybullet_multigoal_gym_b200 import _lib
    L = _lib.load()
    h = C.c_void_p()
    assert L.pmg_create(None, C.byref(h)) == -1
    cfg = _lib.PmgConfig(7, 4, 8, 1, 0.05, 50, 0, 0, 0, 0, 0, 0)
    assert L.pmg_create(C.byref(cfg), C.byref(h)) == -1 and b"task" in L.pmg_last_error()
    cfg = _lib.PmgConfig(3, 6, 8, 1, 0.05, 50, 0, 0, 0, 0, 0, 0)
    assert L.pmg_create(C.byref(cfg), C.byref(h)) == -1 and b"5 blocks" in L.pmg_last_error()
    cfg = _lib.PmgConfig(4, 6, 8, 1, 0.05, 50, 0, 0, 0, 0, 0, 0)
    assert L.pmg_create(C.byref(cfg), C.byref(h)) == -1 and b"5 blocks" in L.pmg_last_error()
    cfg = _lib.PmgConfig(4, 3, 8, 1, 0.05, 50, 0, 1, 0, 0, 0, 0)  # grip-informed goals are a block_stack option
    assert L.pmg_create(C.byref(cfg), C.byref(h)) == -1 and b"grip_informed_goal" in L.pmg_last_error()
    cfg = _lib.PmgConfig(0, 0, 0, 1, 0.05, 50, 0, 0, 0, 0, 0, 0)
    assert L.pmg_create(C.byref(cfg), C.byref(h)) == -1
    with pytest.raises(ValueError):
        _lib.check(-1)


def test_make_env_validation_matches_reference():
    import pybullet_multigoal_gym_b200 as pmg
    with pytest.raises(ValueError):          # __init__.py:55
        pmg.make_env(task="juggle")
    with pytest.raises(AssertionError):      # __init__.py:18
        pmg.make_env(task="reach", gripper="claw")
    with pytest.raises(AssertionError):      # __init__.py:108
        pmg.make_env(task="block_stack", num_block=6)
    with pytest.raises(AssertionError):      # kuka_multi_step_envs.py:158
        pmg.make_env(task="block_rearrange", num_block=3, grip_informed_goal=True)
    with pytest.raises(AssertionError):      # kuka_multi_step_envs.py:159
        pmg.make_env(task="block_rearrange", num_block=3, task_decomposition=True)
    for kw in (dict(task="insertion"), dict(task="chest_push"), dict(task="reach", gripper="robotiq85"),
               dict(task="push", image_observation=True), dict(task="chest_pick_and_place", use_curriculum=True)):
        with pytest.raises(NotImplementedError):
            pmg.make_env(**kw)


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    import pybullet_multigoal_gym_b200 as pmg
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            pmg.make_env(task="reach", batch=4)
    pkg = os.path.join(ROOT, "pybullet_multigoal_gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pmg_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_seeding_matches_gym_and_numpy():
    from pybullet_multigoal_gym_b200 import seeding
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pybullet_shim"))
    try:
        from gym.utils import seeding as gym_seeding  # the shim's restatement of gym 0.17.3
    finally:
        sys.path.pop(0)
    for seed in (0, 1, 7, 123456789, 2 ** 63 + 5):
        rs, _ = gym_seeding.np_random(seed)
        mine = np.random.RandomState()
        mine.seed(seeding.seed_key(seed))
        assert np.array_equal(rs.uniform(size=5), mine.uniform(size=5))
    with pytest.raises(ValueError):
        seeding.seed_key(-1)
    # golden: RandomState seeded like gym for seed 0 starts with these draws (numpy legacy stream)
    rs = np.random.RandomState()
    rs.seed(seeding.seed_key(0))
    np.testing.assert_allclose(rs.uniform(0, 1, 3), [0.05436006, 0.96539094, 0.63269095], atol=1e-8)


def test_spaces():
    from pybullet_multigoal_gym_b200 import spaces
    b = spaces.Box(-np.ones([3]), np.ones([3]))
    assert b.shape == (3,) and b.contains(np.zeros(3)) and not b.contains(np.zeros(4)) and not b.contains(np.array([0, 0, 1.5]))
    assert b.contains(b.sample())
    d = spaces.Dict(dict(state=spaces.Box(-np.inf, np.inf, shape=(8192, 3), dtype="float32")))
    assert d["state"].shape == (8192, 3)


def test_flat_layout_roundtrip():
    from pybullet_multigoal_gym_b200.sharded import flat_layout, shard_range, split_flat
    B, W, world = 5, 12, 3
    lay = flat_layout(B, W)
    assert lay["bytes"] % 16 == 0 and lay["reward"] == 4 * B * W
    g = torch.zeros((world, lay["bytes"]), dtype=torch.uint8)
    for r in range(world):
        g[r, lay["obs"]:lay["reward"]] = (torch.arange(B * W, dtype=torch.float32) + 100 * r).view(torch.uint8)
        g[r, lay["reward"]:lay["done"]] = torch.full((B,), -float(r)).view(torch.uint8)
        g[r, lay["done"]:lay["success"]] = r % 2
        g[r, lay["success"]:lay["success"] + B] = 1
    obs, reward, done, ok = split_flat(g, B, W)
    assert obs.shape == (world * B, W) and obs[B, 0] == 100 and obs[2 * B + 1, 3] == 215
    assert reward.tolist() == [0.0] * B + [-1.0] * B + [-2.0] * B
    assert done.tolist() == [False] * B + [True] * B + [False] * B and ok.all()
    assert shard_range(8192, 3, 8) == (3072, 4096)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from pybullet_multigoal_gym_b200.sharded import flat_layout, gather_flat, split_flat
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    B, W = 4, 6
    lay = flat_layout(B, W)
    flat = torch.zeros((lay["bytes"],), dtype=torch.uint8)
    flat[lay["obs"]:lay["reward"]] = (torch.arange(B * W, dtype=torch.float32) + 1000 * rank).view(torch.uint8)
    flat[lay["reward"]:lay["done"]] = torch.full((B,), -1.0 * rank).view(torch.uint8)
    flat[lay["done"]:lay["success"]] = rank
    g = gather_flat(flat)
    obs, reward, done, ok = split_flat(g, B, W)
    q.put((rank, obs[:, 0].tolist(), reward.tolist(), done.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_single_all_gather_of_the_flat_step_output_gloo_world2():
    """N > 1 host logic on CPU: gloo, world_size 2, the one collective of a sharded step."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, col0, reward, done in results:
        assert col0 == [0.0, 6.0, 12.0, 18.0, 1000.0, 1006.0, 1012.0, 1018.0]
        assert reward == [0.0] * 4 + [-1.0] * 4
        assert done == [False] * 4 + [True] * 4


class _FakeGatherLib:
    """Stands in for libpmg.so's gather entry points: rank `bad` cannot create its gather buffer."""
    def __init__(self, rank, bad):
        self.rank, self.bad, self.connected = rank, bad, False

    def pmg_gather_create(self, h, rank, world, out):
        return -2 if self.rank == self.bad else 0

    def pmg_gather_connect(self, h, blob):
        self.connected = True
        return 0

    def pmg_gather_layout(self, h, lay):
        for i, v in enumerate((1024, 0, 256, 512, 768, 0)):
            lay[i] = v
        return 0


def _connect_worker(rank, world, port, bad, required, q):
    import types
    import warnings
    import torch.distributed as dist
    from pybullet_multigoal_gym_b200 import _lib, sharded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    _lib.check = lambda rc: None if rc == 0 else (_ for _ in ()).throw(_lib.PmgError("pmg error %d: no peer access" % rc))
    fake = _FakeGatherLib(rank, bad)
    e = object.__new__(sharded.ShardedKukaEnv)
    e.group, e.rank, e.world, e.fused, e._fused_required = None, rank, world, True, required
    e.global_batch, e.local_batch = 8, 4
    e.env = types.SimpleNamespace(_L=fake, _h=None, device=torch.device("cpu"), row_width=6, action_dim=3)
    outcome = "ok"
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            # pinned host mirrors need a CUDA runtime: the test only follows the agreement, so stop before them
            torch.Tensor.pin_memory = lambda self, *a, **k: self
            e._connect_peers()
        except RuntimeError as err:
            outcome = "raised: %s" % err
    q.put((rank, e.fused, fake.connected, outcome, [str(x.message) for x in w]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bad,required", [(-1, False), (1, False), (1, True)])
def test_peer_gather_is_agreed_by_all_ranks_gloo_world2(bad, required):
    """ShardedKukaEnv._connect_peers: when any rank cannot set up the peer-memory gather, EVERY rank falls back to the
    NCCL all-gather (fused=None) or every rank raises (fused=True) -- never a mix, never a hang."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() + 7 * (bad + 2) + int(required)) % 2000
    procs = [ctx.Process(target=_connect_worker, args=(r, 2, port, bad, required, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, fused, connected, outcome, warns in results:
        if bad < 0:
            assert fused and connected and outcome == "ok" and not warns
        elif required:
            assert outcome.startswith("raised: fused gather unavailable")
        else:
            assert not fused and outcome == "ok" and any("falling back to one NCCL all-gather" in m for m in warns)
            assert not connected   # nobody maps a peer buffer unless everybody can


def test_step_demonstrator_matches_reference_golden():
    """envs.StepDemonstrator against a trace of the reference's utils/demonstrator.py (tools/gen_demonstrator_golden.py)."""
    import json
    from pybullet_multigoal_gym_b200.envs import StepDemonstrator
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "step_demonstrator.json")))
    for run in g["runs"]:
        d = StepDemonstrator(g["demonstrations"], stick_with_final_goal=run["stick_with_final_goal"])
        for rec in run["trace"]:
            if rec[0] == "next":
                assert [d.get_next_goal(), bool(d.final), d.current_goal, d.demon_ind] == rec[1:]
            elif rec[0] == "manual_reset":
                d.manual_reset(rec[1])
                assert [bool(d.final), d.current_goal, d.demon_ind, d.current_final_goal] == rec[2:]
            else:
                d.reset_with_the_last_sub_goal_index(rec[1])
                assert [bool(d.final), d.current_goal, d.demon_ind, d.current_final_goal] == rec[2:]
