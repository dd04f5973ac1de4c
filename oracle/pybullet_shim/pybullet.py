"""`pybullet` stand-in (TEST INFRASTRUCTURE): the 20 entry points the reference's in-scope step path
calls, answered by the CPU oracle in oracle/libpmg_oracle.so.

The authoritative simulation state is kept here in numpy (joint state, motor settings, free-body
poses) and synchronised with an oracle environment around stepSimulation(); the contact caches live
in the oracle and survive resets exactly as Bullet's persistent manifolds do.
"""
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle import pmg_oracle as _O  # noqa: E402

GUI, DIRECT = 1, 2
POSITION_CONTROL, VELOCITY_CONTROL, TORQUE_CONTROL = 2, 0, 1
COV_ENABLE_GUI = 1
URDF_USE_SELF_COLLISION = 8
JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_FIXED = 0, 1, 4
_OUTER_DT = 0.04  # setPhysicsEngineParameter(fixedTimeStep=...) scales the motor max impulse

_next_client = [0]


def _urdf_joints(path):
    """Joint list in Bullet's link order: depth-first from the root link, children in file order."""
    root = ET.parse(path).getroot()
    joints = root.findall("joint")
    children = {}
    child_links = set()
    for j in joints:
        children.setdefault(j.find("parent").get("link"), []).append(j)
        child_links.add(j.find("child").get("link"))
    roots = [l.get("name") for l in root.findall("link") if l.get("name") not in child_links]
    out = []

    def visit(link):
        for j in children.get(link, []):
            out.append(j)
            visit(j.find("child").get("link"))
    visit(roots[0])
    return out


class _World(object):
    def __init__(self):
        self.client_id = _next_client[0]
        _next_client[0] += 1
        self.bodies = []          # dicts: kind in {'kuka','static','block','marker'}
        self.q = np.zeros(9)
        self.qd = np.zeros(9)
        self.mot_target = np.zeros(9)
        self.mot_maximp = np.zeros(9)
        self.blocks = []          # body ids of the dynamic blocks, in load order
        self.slide = False        # long_table.urdf + cylinder_bulk.urdf: the Slide scene (one puck instead of a cube)
        self.oracle = None
        self.scratch = _O.OracleEnv("reach")

    # ---- oracle synchronisation ---------------------------------------------------------
    def _pack(self, nb, goal_dim):
        s = np.concatenate([self.q, self.qd, np.zeros(3), np.zeros(7), self.mot_target, self.mot_maximp])
        for bid in self.blocks:
            b = self.bodies[bid]
            s = np.concatenate([s, b["pos"], b["orn"], b["lin"], b["ang"]])
        # dummy goal rows (z levels keep the oracle's block-stack bookkeeping well defined)
        goal = np.zeros(goal_dim)
        goal[2::3] = 0.175 + 0.03 * np.arange(goal_dim // 3)
        return np.concatenate([s, goal, [0.0]])

    def _ensure_oracle(self):
        nb = len(self.blocks)
        if self.oracle is None or self.oracle.nb != nb:
            if self.slide:
                assert nb == 1
                self.oracle = _O.OracleEnv("slide")
            else:
                self.oracle = _O.OracleEnv("block_stack", num_block=nb) if nb else _O.OracleEnv("reach")
        return self.oracle

    def step(self):
        o = self._ensure_oracle()
        o.poke_state(self._pack(o.nb, o.dims[3]))
        o.step_simulation()
        s = o.get_state()
        self.q, self.qd = s[0:9].copy(), s[9:18].copy()
        for k, bid in enumerate(self.blocks):
            b = self.bodies[bid]
            base = 46 + 13 * k
            b["pos"], b["orn"] = s[base:base + 3].copy(), s[base + 3:base + 7].copy()
            b["lin"], b["ang"] = s[base + 7:base + 10].copy(), s[base + 10:base + 13].copy()

    def link_state(self, which):
        s = np.concatenate([self.q, self.qd, np.zeros(3), np.zeros(7), np.zeros(9), np.zeros(9), np.zeros(3), [0.0]])
        self.scratch.set_state(s)
        return self.scratch.link_state(which)


_ARM = ["iiwa_joint_%d" % i for i in range(1, 8)]
_FINGERS = ["iiwa_gripper_finger1_joint", "iiwa_gripper_finger2_joint"]
_LINK_QUERY = {"iiwa_gripper_tip": 0, "iiwa_gripper_base_link": 1, "iiwa_gripper_finger1_finger_tab_link": 2,
               "iiwa_gripper_finger2_finger_tab_link": 3, "iiwa_gripper_finger1": 4, "iiwa_gripper_finger2": 5}


def _dof_of(world, body, joint_index):
    name = world.bodies[body]["joints"][joint_index].get("name")
    if name in _ARM:
        return _ARM.index(name)
    if name in _FINGERS:
        return 7 + _FINGERS.index(name)
    raise KeyError("joint %s has no degree of freedom" % name)


# ---- world configuration (base_env.py:203-220) -------------------------------------------------
def setGravity(world, x, y, z):
    assert (x, y) == (0, 0) and abs(z + 9.81) < 1e-12


def setDefaultContactERP(world, erp):
    assert erp == 0.9


def setPhysicsEngineParameter(world, fixedTimeStep=None, numSolverIterations=None, numSubSteps=None, **kw):
    assert abs(fixedTimeStep - 0.04) < 1e-12 and numSolverIterations == 5 and numSubSteps == 20


def setRealTimeSimulation(world, flag):
    assert not flag


def configureDebugVisualizer(world, *a, **k):
    pass


def resetDebugVisualizerCamera(world, *a, **k):
    pass


def computeViewMatrix(world, **k):
    return [0.0] * 16


def computeProjectionMatrixFOV(world, **k):
    return [0.0] * 16


def enableJointForceTorqueSensor(world, **k):
    pass


def disconnect(world):
    pass


# ---- model loading -----------------------------------------------------------------------------
def loadURDF(world, path, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), useFixedBase=False, flags=0):
    name = os.path.basename(path)
    body = {"path": path, "pos": np.array(basePosition, dtype=float), "orn": np.array(baseOrientation, dtype=float),
            "lin": np.zeros(3), "ang": np.zeros(3)}
    if name == "iiwa14_parallel_jaw.urdf":
        body["kind"] = "kuka"
        body["joints"] = _urdf_joints(path)
    elif name in ("table.urdf",):
        body["kind"] = "static"
        assert np.allclose(basePosition, [-0.52, 0.0, 0.08])
    elif name == "long_table.urdf":  # Slide (kuka_single_step_base_env.py:53-56)
        body["kind"] = "static"
        assert np.allclose(basePosition, [-0.70, 0.0, 0.08])
        world.slide = True
    elif name == "cylinder_bulk.urdf":
        body["kind"] = "block"
        assert world.slide, "the puck only exists on the long table"
    elif name.startswith("block"):
        body["kind"] = "block"
    elif name.startswith("target"):
        body["kind"] = "marker"
    else:
        raise NotImplementedError("shim: %s is outside the in-scope scenes" % name)
    world.bodies.append(body)
    bid = len(world.bodies) - 1
    if body["kind"] == "block":
        world.blocks.append(bid)
    return bid


def getNumJoints(world, body):
    return len(world.bodies[body].get("joints", []))


def getJointInfo(world, body, j):
    jt = world.bodies[body]["joints"][j]
    typ = {"revolute": JOINT_REVOLUTE, "prismatic": JOINT_PRISMATIC, "fixed": JOINT_FIXED}[jt.get("type")]
    lim = jt.find("limit")
    lo = float(lim.get("lower")) if lim is not None else 0.0
    hi = float(lim.get("upper")) if lim is not None else -1.0
    vel = float(lim.get("velocity")) if lim is not None else 0.0
    eff = float(lim.get("effort")) if lim is not None else 0.0
    return (j, jt.get("name").encode("utf8"), typ, -1, -1, 0, 0.0, 0.0, lo, hi, eff, vel,
            jt.find("child").get("link").encode("utf8"), (0.0, 0.0, 1.0), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), j - 1)


# ---- joint state / motors ------------------------------------------------------------------------
def resetJointState(world, body, jointIndex, targetValue, targetVelocity=0.0):
    d = _dof_of(world, body, jointIndex)
    world.q[d], world.qd[d] = targetValue, targetVelocity


def getJointState(world, body, jointIndex):
    d = _dof_of(world, body, jointIndex)
    return (float(world.q[d]), float(world.qd[d]), (0.0,) * 6, 0.0)


def _set_motor(world, body, jointIndex, mode, target, target_vel, force, kp, kd):
    assert mode == POSITION_CONTROL and target_vel == 0
    d = _dof_of(world, body, jointIndex)
    if force != 0:
        assert abs(kp - 0.03) < 1e-12 and abs(kd - 1.0) < 1e-12, "the oracle hard-codes the reference's gains"
    world.mot_target[d] = target
    world.mot_maximp[d] = force * _OUTER_DT


def setJointMotorControl2(world, bodyIndex, jointIndex, controlMode, targetPosition=0.0, targetVelocity=0.0,
                          positionGain=0.1, velocityGain=1.0, force=0.0, **kw):
    _set_motor(world, bodyIndex, jointIndex, controlMode, targetPosition, targetVelocity, force, positionGain, velocityGain)


def setJointMotorControlArray(world, bodyUniqueId, jointIndices, controlMode, targetPositions, targetVelocities,
                              forces, positionGains, velocityGains):
    for k, j in enumerate(jointIndices):
        _set_motor(world, bodyUniqueId, j, controlMode, float(targetPositions[k]), float(targetVelocities[k]),
                   float(forces[k]), float(positionGains[k]), float(velocityGains[k]))


def calculateInverseKinematics(world, bodyUniqueId, endEffectorLinkIndex, targetPosition, targetOrientation,
                               lowerLimits=None, upperLimits=None, jointRanges=None, restPoses=None,
                               maxNumIterations=20, residualThreshold=1e-4):
    joints = world.bodies[bodyUniqueId]["joints"]
    assert joints[endEffectorLinkIndex].get("name") == "iiwa_gripper_tip_joint"
    # 7-element null-space lists on a 9-DoF body: pybullet ignores them
    assert len(lowerLimits) == 7 and len(restPoses) == 7
    return tuple(_O.ik(world.q, np.asarray(targetPosition, dtype=float), np.asarray(targetOrientation, dtype=float),
                       maxNumIterations, residualThreshold))


def stepSimulation(world):
    world.step()


# ---- state queries -------------------------------------------------------------------------------
def getLinkState(world, body, link, computeLinkVelocity=0):
    child = world.bodies[body]["joints"][link].find("child").get("link")
    which = _LINK_QUERY.get(child)
    if which is None:
        # links the step path never reads (BodyPart.__init__ records their initial pose only)
        pos, orn, lin, ang = (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)
    else:
        s = world.link_state(which)
        pos, orn, lin, ang = tuple(s[0:3]), tuple(s[3:7]), tuple(s[7:10]), tuple(s[10:13])
    if computeLinkVelocity:
        return (pos, orn, (0, 0, 0), (0, 0, 0, 1), pos, orn, lin, ang)
    return (pos, orn, (0, 0, 0), (0, 0, 0, 1), pos, orn)


def getEulerFromQuaternion(world, q):
    x, y, z, w = q
    roll = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    pitch = np.arcsin(np.clip(2 * (w * y - z * x), -1, 1))
    yaw = np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    return (roll, pitch, yaw)


def getBasePositionAndOrientation(world, body):
    b = world.bodies[body]
    return tuple(b["pos"]), tuple(b["orn"])


def getBaseVelocity(world, body):
    b = world.bodies[body]
    return tuple(b["lin"]), tuple(b["ang"])


def resetBasePositionAndOrientation(world, body, pos, orn):
    b = world.bodies[body]
    b["pos"], b["orn"] = np.array(pos, dtype=float), np.array(orn, dtype=float)
    b["lin"], b["ang"] = np.zeros(3), np.zeros(3)
