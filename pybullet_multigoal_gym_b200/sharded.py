"""Batch sharding over several GPUs: one process per GPU (torch.distributed), envs partitioned in
contiguous ranges, no communication during physics, and exactly ONE all-gather per step for the
returned observation batch (north_star).

The step kernel writes observation rows, rewards and flags straight into one flat byte buffer
[obs f32 B*W | reward f32 B | done u8 B | success u8 B]; that buffer is the all-gather operand, so
the collective needs no packing kernels.  Works with the NCCL backend (CUDA tensors) and, for the
CPU tests of the host logic, with gloo through `gather_flat`.
"""
import torch
import torch.distributed as dist


def flat_layout(batch, width):
    """Byte offsets of the sections of one rank's flat step-output buffer."""
    obs = 0
    reward = obs + 4 * batch * width
    done = reward + 4 * batch
    success = done + batch
    total = success + batch
    total_padded = (total + 15) // 16 * 16
    return {"obs": obs, "reward": reward, "done": done, "success": success, "bytes": total_padded}


def split_flat(gathered, batch, width):
    """gathered: [world, bytes] uint8 -> global (obs [world*batch, width] f32, reward, done, success)."""
    lay = flat_layout(batch, width)
    world = gathered.shape[0]
    obs = gathered[:, lay["obs"]:lay["reward"]].contiguous().view(torch.float32).reshape(world * batch, width)
    reward = gathered[:, lay["reward"]:lay["done"]].contiguous().view(torch.float32).reshape(world * batch)
    done = gathered[:, lay["done"]:lay["success"]].reshape(world * batch).bool()
    success = gathered[:, lay["success"]:lay["success"] + batch].reshape(world * batch).bool()
    return obs, reward, done, success


def gather_flat(local_flat, group=None):
    """One all-gather of every rank's flat buffer -> [world, bytes]."""
    world = dist.get_world_size(group)
    out = torch.empty((world, local_flat.numel()), dtype=torch.uint8, device=local_flat.device)
    dist.all_gather_into_tensor(out.view(-1), local_flat, group=group)
    return out


def shard_range(global_batch, rank, world):
    """Contiguous env range [lo, hi) owned by `rank`."""
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


class ShardedKukaEnv:
    """Rank-local slice of a global batch; step() returns the gathered global observation batch."""

    def __init__(self, task, global_batch, seed=0, group=None, **kw):
        from .envs import KukaBulletMGEnv
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lo, hi = shard_range(global_batch, self.rank, self.world)
        self.lo, self.hi, self.global_batch = lo, hi, global_batch
        device = kw.pop("device", torch.cuda.current_device())
        # env i of the global batch is seeded with seed + i, whatever the sharding
        self.env = KukaBulletMGEnv(task, batch=hi - lo, device=device, seed=seed + lo, **kw)
        self.local_batch = hi - lo
        self.lay = flat_layout(self.local_batch, self.env.row_width)
        self._flat = torch.empty((self.lay["bytes"],), dtype=torch.uint8, device=self.env.device)
        f = self._flat
        self._obs = f[self.lay["obs"]:self.lay["reward"]].view(torch.float32).view(self.local_batch, self.env.row_width)
        self._reward = f[self.lay["reward"]:self.lay["done"]].view(torch.float32)
        self._done = f[self.lay["done"]:self.lay["success"]]
        self._success = f[self.lay["success"]:self.lay["success"] + self.local_batch]

    def local_actions(self, global_actions):
        return global_actions[self.lo:self.hi]

    def step_gathered(self, local_action):
        """local_action: [local_batch, A] CUDA tensor.  Returns (packed obs [global, W], reward, done, success)."""
        self.env.step_packed(local_action, self._obs, self._reward, self._done, self._success)
        g = gather_flat(self._flat, self.group)
        return split_flat(g, self.local_batch, self.env.row_width)

    def step(self, local_action):
        obs, reward, done, ok = self.step_gathered(local_action)
        return self.env._split(obs), reward, done, {"goal_achieved": ok, "is_success": ok, "TimeLimit.truncated": done}

    def reset(self):
        local = self.env.reset(device_output=True)
        packed = torch.cat([local[k] for k in ("observation", "policy_state", "achieved_goal", "desired_goal")], dim=1).contiguous()
        out = torch.empty((self.world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(out.view(-1), packed.view(-1), group=self.group)
        return self.env._split(out.reshape(self.global_batch, self.env.row_width))

    def close(self):
        self.env.close()
