"""Builds tests/emu/libpmg_coop_emu.so: the lane-cooperative device code compiled for the host (g++).

TEST INFRASTRUCTURE (see pmg_coop_emu.cpp); the product never loads it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "pybullet_multigoal_gym_b200", "csrc")
SO = os.path.join(HERE, "libpmg_coop_emu.so")
SRC = os.path.join(HERE, "pmg_coop_emu.cpp")
DEPS = [SRC] + [os.path.join(CSRC, f) for f in ("pmg_coop.cuh", "pmg_sim.cuh", "pmg_physics.cuh", "pmg_emu_shim.h", "pmg_spawn.cuh")] + \
       [os.path.join(HERE, "..", "..", "include", "pmg_model_constants.h")]


def build(force=False):
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
    return SO


if __name__ == "__main__":
    print(build(force=True))
