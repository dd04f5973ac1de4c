"""Batch sharding over several GPUs: one process per GPU (torch.distributed), envs partitioned in
contiguous ranges, no communication during physics, and exactly ONE gather per step for the
returned observation batch (north_star).

Default on GPUs (`fused=True`): the gather is part of the step itself -- the kernel's epilogue stores
each environment's row, reward and flags into every rank's gather buffer over peer-mapped memory
(NVLink / NVSwitch), the launch's last arrival publishes a sequence flag and waits for the peers'
(include/pmg.h, pmg_step_gather): no collective launch, no packing or splitting copies, no per-step
allocation; the returned tensors are views of the gather buffer (valid until the step after the next).
torch.distributed only carries the one-off exchange of the IPC handles.

Fallback (`fused=False`, and the gloo CPU tests of the host logic): the step kernel writes into one
flat byte buffer [obs f32 B*W | reward f32 B | done u8 B | success u8 B] which is the operand of a
single `all_gather_into_tensor`.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class _DeviceMemory:
    """A raw device allocation owned by libpmg.so, exposed to torch through the CUDA array interface."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


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

    def __init__(self, task, global_batch, seed=0, group=None, fused=None, **kw):
        from .envs import KukaBulletMGEnv
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lo, hi = shard_range(global_batch, self.rank, self.world)
        self.lo, self.hi, self.global_batch = lo, hi, global_batch
        device = kw.pop("device", torch.cuda.current_device())
        # env i of the global batch is seeded with seed + i, whatever the sharding (host MT19937 streams: seed + lo + i;
        # device Philox streams: (seed, env_index_base + i))
        if kw.get("device_sampling") or kw.get("auto_reset"):
            self.env = KukaBulletMGEnv(task, batch=hi - lo, device=device, seed=seed, env_index_base=lo, **kw)
        else:
            self.env = KukaBulletMGEnv(task, batch=hi - lo, device=device, seed=seed + lo, **kw)
        self.local_batch = hi - lo
        self._fused_required = fused is True  # None: use it when the box allows (NCCL group of <= 8 ranks with peer access)
        if fused is None:
            fused = dist.get_backend(group) == "nccl" and self.world <= 8
        self.fused = bool(fused)
        if self.fused:
            self._connect_peers()
        self.lay = flat_layout(self.local_batch, self.env.row_width)
        self._flat = torch.empty((self.lay["bytes"],), dtype=torch.uint8, device=self.env.device)
        f = self._flat
        self._obs = f[self.lay["obs"]:self.lay["reward"]].view(torch.float32).view(self.local_batch, self.env.row_width)
        self._reward = f[self.lay["reward"]:self.lay["done"]].view(torch.float32)
        self._done = f[self.lay["done"]:self.lay["success"]]
        self._success = f[self.lay["success"]:self.lay["success"] + self.local_batch]

    def _connect_peers(self):
        """pmg_gather_create on every rank, one exchange of the 64-byte IPC handles, pmg_gather_connect; then torch
        views of the two parity copies of the gather buffer, built once."""
        L, h = self.env._L, self.env._h
        mine = (C.c_char * 64)()
        err = None
        try:
            _lib.check(L.pmg_gather_create(h, self.rank, self.world, mine))
        except (RuntimeError, ValueError) as e:
            err = e
        handles = [None] * self.world
        dist.all_gather_object(handles, None if err else bytes(mine.raw), group=self.group)
        if err is None and all(x is not None for x in handles):
            try:
                _lib.check(L.pmg_gather_connect(h, C.c_char_p(b"".join(handles))))
            except (RuntimeError, ValueError) as e:
                err = e
        # one reduction doubles as the barrier (every rank has mapped every buffer before the first step publishes
        # into them) and as the agreement on the path: a box without peer access between some pair of GPUs (no
        # NVLink / IPC disabled in the container) puts ALL ranks on the NCCL all-gather, never a mix
        ok = torch.tensor([0 if (err is not None or any(x is None for x in handles)) else 1], dtype=torch.int32, device=self.env.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            if self._fused_required:
                raise RuntimeError("fused gather unavailable on this box: %s" % (err or "a peer rank could not map the gather buffers"))
            import warnings
            warnings.warn("peer-memory gather unavailable (%s): falling back to one NCCL all-gather per step" % (err or "peer rank failed"))
            self.fused = False
            return
        lay = (C.c_int64 * 6)()
        _lib.check(L.pmg_gather_layout(h, lay))
        self._parity_bytes, _, off_r, off_d, off_s, _ = [int(v) for v in lay]
        self._offsets = (0, off_r, off_d, off_s)
        self._views = {}
        self._ptrs = (C.c_void_p * 4)()
        Bg, W = self.global_batch, self.env.row_width
        # pinned host mirrors for the host-buffer path (step with numpy actions)
        self._h_action = torch.empty((self.local_batch, self.env.action_dim), dtype=torch.float32).pin_memory()
        self._d_action = torch.empty((self.local_batch, self.env.action_dim), dtype=torch.float32, device=self.env.device)
        self._h_out = torch.empty((self._parity_bytes,), dtype=torch.uint8).pin_memory()
        self._gather_bytes = off_s + Bg

    def _view(self, base):
        v = self._views.get(base)
        if v is None:
            Bg, W = self.global_batch, self.env.row_width
            _, off_r, off_d, off_s = self._offsets
            raw = torch.as_tensor(_DeviceMemory(base, self._parity_bytes), device=self.env.device)
            v = (raw, raw[:4 * Bg * W].view(torch.float32).view(Bg, W), raw[off_r:off_r + 4 * Bg].view(torch.float32),
                 raw[off_d:off_d + Bg].view(torch.bool), raw[off_s:off_s + Bg].view(torch.bool))
            self._views[base] = v
        return v

    def local_actions(self, global_actions):
        return global_actions[self.lo:self.hi]

    def step_gathered(self, local_action):
        """local_action: [local_batch, A] float32 CUDA tensor.  Returns (packed obs [global, W], reward, done, success)
        of the GLOBAL batch; with the fused gather these are views of the gather buffer."""
        if self.fused:
            with torch.cuda.device(self.env.device):
                _lib.check(self.env._L.pmg_step_gather(self.env._h, C.c_void_p(local_action.data_ptr()), self._ptrs, self.env._stream()))
            return self._view(self._ptrs[0])[1:]
        self.env.step_packed(local_action, self._obs, self._reward, self._done, self._success)
        g = gather_flat(self._flat, self.group)
        return split_flat(g, self.local_batch, self.env.row_width)

    def step_host(self, local_action_np):
        """Host-buffer path of the sharded env: numpy actions of the LOCAL shard in, numpy arrays of the GLOBAL batch
        out (H2D from pinned memory, step + gather, D2H of the gathered batch, one synchronise)."""
        if not self.fused:
            a = torch.from_numpy(np.ascontiguousarray(local_action_np, dtype=np.float32)).to(self.env.device)
            obs, reward, done, ok = self.step_gathered(a)
            return obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), ok.cpu().numpy()
        self._h_action.numpy()[...] = local_action_np
        self._d_action.copy_(self._h_action, non_blocking=True)
        self.step_gathered(self._d_action)
        raw = self._view(self._ptrs[0])[0]
        n = self._gather_bytes
        self._h_out[:n].copy_(raw[:n], non_blocking=True)
        torch.cuda.current_stream(self.env.device).synchronize()
        Bg, W = self.global_batch, self.env.row_width
        _, off_r, off_d, off_s = self._offsets
        host = self._h_out.numpy()
        return (host[:4 * Bg * W].view(np.float32).reshape(Bg, W), host[off_r:off_r + 4 * Bg].view(np.float32),
                host[off_d:off_d + Bg].view(np.bool_), host[off_s:off_s + Bg].view(np.bool_))

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
