#!/usr/bin/env python3
"""Bake the reference's model constants into include/pmg_model_constants.h.

Runs ONLY in the build container (it reads /root/reference); the generated
header is committed, so nothing on the GPU box needs the reference tree.

What is extracted (all from the reference's assets, nothing from its code):
  * assets/robots/kuka/iiwa14_parallel_jaw.urdf : kinematic tree, joint frames,
    axes, limits, joint damping, link masses, inertial origins, finger boxes,
    lateral friction, inertia_scaling.
  * assets/robots/kuka/meshes/iiwa14/collision/link_[1-7].stl : AABBs, because
    PyBullet's loadURDF (no URDF_USE_INERTIA_FROM_FILE flag, robot_bases.py:68-77)
    replaces the URDF <inertia> by collisionShape->calculateLocalInertia(mass)
    ([BULLET-MEMORY]).
  * assets/objects/table.urdf, block.urdf : box sizes, masses, friction.

Inertia rule restated:
  * single primitive child at identity inertial frame -> the primitive's own
    inertia (box: m/12*(ly^2+lz^2,...), cylinder-Z: m/12*(3r^2+h^2) x2, m r^2/2);
  * otherwise (mesh links; inertial origin != 0) -> box inertia of the compound
    AABB expressed in the inertial frame, AABB grown by the convex-hull margin
    (0.001) plus the compound margin (0.001) on every side;
  * multiplied by <inertia_scaling> when the link has one.
"""
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference/pybullet_multigoal_gym/assets"
URDF = os.path.join(REF, "robots/kuka/iiwa14_parallel_jaw.urdf")
OUT = os.path.join(os.path.dirname(__file__), "..", "include", "pmg_model_constants.h")

URDF_MARGIN = 0.001  # gUrdfDefaultCollisionMargin [BULLET-MEMORY]


def floats(s):
    return [float(x) for x in s.replace(",", " ").split()]


def rpy_to_mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def stl_aabb(path):
    with open(path, "rb") as f:
        data = f.read()
    n = struct.unpack_from("<I", data, 80)[0]
    assert len(data) == 84 + 50 * n, "not a binary STL: %s" % path
    rec = np.frombuffer(data, dtype=np.uint8, offset=84).reshape(n, 50)
    v = rec[:, 12:48].copy().view("<f4").reshape(n * 3, 3).astype(np.float64)
    return v.min(0), v.max(0)


def box_inertia(mass, full):
    lx, ly, lz = full
    return mass / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])


def link_info(link, urdf_dir):
    inertial = link.find("inertial")
    mass = float(inertial.find("mass").get("value")) if inertial is not None else 0.0
    com = np.zeros(3)
    if inertial is not None and inertial.find("origin") is not None:
        com = np.array(floats(inertial.find("origin").get("xyz", "0 0 0")))
        assert np.allclose(floats(inertial.find("origin").get("rpy", "0 0 0")), 0.0)
    contact = link.find("contact")
    friction, iscale = 0.5, 1.0  # Bullet default lateral friction 0.5
    if contact is not None:
        if contact.find("lateral_friction") is not None:
            friction = float(contact.find("lateral_friction").get("value"))
        if contact.find("inertia_scaling") is not None:
            iscale = float(contact.find("inertia_scaling").get("value"))
    col = link.find("collision")
    shape, dims = None, None
    inertia = np.zeros(3)
    if col is not None:
        corg = col.find("origin")
        cxyz = np.array(floats(corg.get("xyz", "0 0 0"))) if corg is not None else np.zeros(3)
        geom = col.find("geometry")[0]
        identity = np.allclose(com, 0.0) and np.allclose(cxyz, 0.0)
        if geom.tag == "box":
            shape, dims = "box", np.array(floats(geom.get("size")))
            lo, hi = cxyz - dims / 2, cxyz + dims / 2
        elif geom.tag == "cylinder":
            r, L = float(geom.get("radius")), float(geom.get("length"))
            shape, dims = "cylinder", np.array([r, L])
            lo, hi = cxyz - np.array([r, r, L / 2]), cxyz + np.array([r, r, L / 2])
        elif geom.tag == "mesh":
            shape = "mesh"
            lo, hi = stl_aabb(os.path.join(urdf_dir, geom.get("filename")))
            lo, hi = lo + cxyz - URDF_MARGIN, hi + cxyz + URDF_MARGIN  # hull margin
            dims = np.concatenate([lo, hi])
            identity = False
        if mass > 0:
            if identity and shape == "box":
                inertia = box_inertia(mass, dims)
            elif identity and shape == "cylinder":
                r, L = dims
                t1 = mass / 12.0 * L * L + mass / 4.0 * r * r
                inertia = np.array([t1, t1, mass / 2.0 * r * r])
            else:
                full = (hi - lo) + 2 * URDF_MARGIN  # compound margin
                inertia = box_inertia(mass, full)
            inertia = inertia * iscale
    return dict(mass=mass, com=com, inertia=inertia, friction=friction, shape=shape, dims=dims)


def simple_box_urdf(path):
    root = ET.parse(path).getroot()
    link = root.find("link")
    info = link_info(link, os.path.dirname(path))
    assert info["shape"] == "box"
    return info


def simple_cylinder_urdf(path):
    root = ET.parse(path).getroot()
    info = link_info(root.find("link"), os.path.dirname(path))
    assert info["shape"] == "cylinder"
    return info


def fmt(a):
    return ", ".join("%.17g" % float(x) for x in np.asarray(a).ravel())


def main():
    root = ET.parse(URDF).getroot()
    urdf_dir = os.path.dirname(URDF)
    links = {l.get("name"): l for l in root.findall("link")}
    joints = {j.get("name"): j for j in root.findall("joint")}

    # dynamic bodies in Bullet link order (depth-first, file order): arm 1..7,
    # gripper base (fixed to link_7), finger1, finger2. Massless fixed links
    # (tip, hand cam, mocap, tabs) carry no dynamics and are kept as offsets.
    chain = ["iiwa_joint_1", "iiwa_joint_2", "iiwa_joint_3", "iiwa_joint_4", "iiwa_joint_5",
             "iiwa_joint_6", "iiwa_joint_7", "iiwa_gripper_base_joint",
             "iiwa_gripper_finger1_joint", "iiwa_gripper_finger2_joint"]
    body_of_link = {"iiwa_link_0": -1}
    rows = []
    for bi, jn in enumerate(chain):
        j = joints[jn]
        parent, child = j.find("parent").get("link"), j.find("child").get("link")
        org = j.find("origin")
        xyz = np.array(floats(org.get("xyz", "0 0 0")))
        R = rpy_to_mat(floats(org.get("rpy", "0 0 0")))
        jt = {"revolute": 0, "prismatic": 1, "fixed": 2}[j.get("type")]
        axis = np.array(floats(j.find("axis").get("xyz"))) if j.find("axis") is not None else np.zeros(3)
        lim = j.find("limit")
        lo = float(lim.get("lower")) if lim is not None else 0.0
        hi = float(lim.get("upper")) if lim is not None else 0.0
        dyn = j.find("dynamics")
        damp = float(dyn.get("damping", "0")) if dyn is not None else 0.0
        info = link_info(links[child], urdf_dir)
        body_of_link[child] = bi
        rows.append(dict(name=jn, parent=body_of_link[parent], xyz=xyz, R=R, type=jt, axis=axis,
                         lo=lo, hi=hi, damp=damp, **info))
    # base link 0 pose: plane_iiwa_joint is identity; link_0 is the fixed base.
    assert np.allclose(floats(joints["plane_iiwa_joint"].find("origin").get("xyz")), 0)
    tip = joints["iiwa_gripper_tip_joint"]
    assert tip.find("parent").get("link") == "iiwa_link_7"
    tip_xyz = floats(tip.find("origin").get("xyz"))
    tab1 = floats(joints["iiwa_gripper_finger1_finger_tab_joint"].find("origin").get("xyz"))
    tab2 = floats(joints["iiwa_gripper_finger2_finger_tab_joint"].find("origin").get("xyz"))
    plane = link_info(links["plane"], urdf_dir)

    table = simple_box_urdf(os.path.join(REF, "objects/table.urdf"))
    block = simple_box_urdf(os.path.join(REF, "objects/block.urdf"))
    for c in ["blue", "green", "purple", "red", "yellow"]:
        b2 = simple_box_urdf(os.path.join(REF, "objects/block_%s.urdf" % c))
        assert np.allclose(b2["dims"], block["dims"]) and b2["mass"] == block["mass"] and \
            b2["friction"] == block["friction"] and np.allclose(b2["inertia"], block["inertia"]), c

    nb = len(rows)
    dof = 0
    dof_index = []
    for r in rows:
        if r["type"] != 2:
            dof_index.append(dof)
            dof += 1
        else:
            dof_index.append(-1)
    o = []
    w = o.append
    w("/* GENERATED by tools/gen_model.py from the reference's assets -- do not edit.")
    w(" * Source: pybullet_multigoal_gym/assets/robots/kuka/iiwa14_parallel_jaw.urdf:37-523,")
    w(" *         assets/robots/kuka/meshes/iiwa14/collision/link_[1-7].stl (AABBs only),")
    w(" *         assets/objects/table.urdf:8-30, assets/objects/block.urdf:8-34 (= block_<colour>.urdf).")
    w(" * Inertia rule: SURVEY.md A.3 (PyBullet recomputes inertia from the collision shape).")
    w(" * Data table only: shared by oracle/ (double) and csrc/ (float). */")
    w("#ifndef PMG_MODEL_CONSTANTS_H")
    w("#define PMG_MODEL_CONSTANTS_H")
    w("")
    w("#define PMG_NBODY %d   /* dynamic robot bodies: link_1..7, gripper_base, finger1, finger2 */" % nb)
    w("#define PMG_NDOF  %d   /* j1..j7, finger1, finger2 */" % dof)
    w("#define PMG_BODY_LINK7   6")
    w("#define PMG_BODY_GBASE   7")
    w("#define PMG_BODY_FINGER1 8")
    w("#define PMG_BODY_FINGER2 9")
    w("/* joint type: 0 revolute, 1 prismatic, 2 fixed */")
    w("#define PMG_BODY_PARENT   { %s }" % ", ".join(str(r["parent"]) for r in rows))
    w("#define PMG_BODY_JTYPE    { %s }" % ", ".join(str(r["type"]) for r in rows))
    w("#define PMG_BODY_DOF      { %s }" % ", ".join(str(d) for d in dof_index))
    w("#define PMG_BODY_JXYZ     { %s }" % ", ".join("{%s}" % fmt(r["xyz"]) for r in rows))
    w("/* fixed rotation parent->child at q=0, row-major (URDF rpy: Rz*Ry*Rx) */")
    w("#define PMG_BODY_JROT     { %s }" % ", \\\n                            ".join("{%s}" % fmt(r["R"]) for r in rows))
    w("#define PMG_BODY_AXIS     { %s }" % ", ".join("{%s}" % fmt(r["axis"]) for r in rows))
    w("#define PMG_BODY_MASS     { %s }" % fmt([r["mass"] for r in rows]))
    w("#define PMG_BODY_COM      { %s }" % ", ".join("{%s}" % fmt(r["com"]) for r in rows))
    w("/* principal inertia about the COM, link axes (inertial rpy is 0 for every link) */")
    w("#define PMG_BODY_INERTIA  { %s }" % ", \\\n                            ".join("{%s}" % fmt(r["inertia"]) for r in rows))
    mov = [r for r in rows if r["type"] != 2]
    w("#define PMG_DOF_LOWER     { %s }" % fmt([r["lo"] for r in mov]))
    w("#define PMG_DOF_UPPER     { %s }" % fmt([r["hi"] for r in mov]))
    w("#define PMG_DOF_DAMPING   { %s }" % fmt([r["damp"] for r in mov]))
    w("#define PMG_DOF_BODY      { %s }" % ", ".join(str(i) for i, r in enumerate(rows) if r["type"] != 2))
    w("/* IK end effector: iiwa_gripper_tip_joint (fixed), child of link_7 */")
    w("#define PMG_TIP_OFFSET    { %s }" % fmt(tip_xyz))
    w("#define PMG_TAB1_OFFSET   { %s }  /* finger1 -> tab1 */" % fmt(tab1))
    w("#define PMG_TAB2_OFFSET   { %s }  /* finger2 -> tab2 */" % fmt(tab2))
    f1 = rows[8]
    assert f1["shape"] == "box" and np.allclose(rows[9]["dims"], f1["dims"])
    w("#define PMG_FINGER_HALF   { %s }" % fmt(f1["dims"] / 2))
    w("#define PMG_FINGER_FRICTION %s" % fmt([f1["friction"]]))
    gb = rows[7]
    assert gb["shape"] == "cylinder"
    w("/* gripper base: a cylinder about its z axis centred on its link frame (iiwa14_parallel_jaw.urdf:399-416, no <contact> tag) */")
    w("#define PMG_GBASE_RADIUS   %s" % fmt([gb["dims"][0]]))
    w("#define PMG_GBASE_HALF_LEN %s" % fmt([gb["dims"][1] / 2]))
    w("#define PMG_GBASE_FRICTION %s" % fmt([gb["friction"]]))
    w("/* static boxes: table (kuka_single_step_base_env.py:49) and the robot's own 'plane' base link */")
    w("#define PMG_TABLE_CENTER  { -0.52, 0.0, 0.08 }")
    w("#define PMG_TABLE_HALF    { %s }" % fmt(table["dims"] / 2))
    w("#define PMG_TABLE_FRICTION %s" % fmt([table["friction"]]))
    w("#define PMG_FLOOR_CENTER  { 0.0, 0.0, 0.0 }")
    w("#define PMG_FLOOR_HALF    { %s }" % fmt(plane["dims"] / 2))
    w("#define PMG_FLOOR_FRICTION %s" % fmt([plane["friction"]]))
    w("#define PMG_BLOCK_HALF    %s" % fmt([block["dims"][0] / 2]))
    assert np.allclose(block["dims"], block["dims"][0])
    w("#define PMG_BLOCK_MASS    %s" % fmt([block["mass"]]))
    w("#define PMG_BLOCK_INERTIA %s" % fmt([block["inertia"][0]]))
    w("#define PMG_BLOCK_FRICTION %s" % fmt([block["friction"]]))
    # Slide scene (kuka_single_step_base_env.py:53-56,89-93): long table at x = -0.70, puck = cylinder_bulk.urdf (axis z)
    long_table = simple_box_urdf(os.path.join(REF, "objects/long_table.urdf"))
    puck = simple_cylinder_urdf(os.path.join(REF, "objects/cylinder_bulk.urdf"))
    w("/* Slide: assets/objects/long_table.urdf:10,20 at (-0.70, 0, 0.08); assets/objects/cylinder_bulk.urdf:10-31 */")
    w("#define PMG_LONG_TABLE_CENTER  { -0.70, 0.0, 0.08 }")
    w("#define PMG_LONG_TABLE_HALF    { %s }" % fmt(long_table["dims"] / 2))
    w("#define PMG_LONG_TABLE_FRICTION %s" % fmt([long_table["friction"]]))
    w("#define PMG_PUCK_RADIUS   %s" % fmt([puck["dims"][0]]))
    w("#define PMG_PUCK_HALF_LEN %s" % fmt([puck["dims"][1] / 2]))
    w("#define PMG_PUCK_MASS     %s" % fmt([puck["mass"]]))
    w("#define PMG_PUCK_INERTIA  { %s }  /* cylinder about z, x inertia_scaling */" % fmt(puck["inertia"]))
    w("#define PMG_PUCK_FRICTION %s" % fmt([puck["friction"]]))
    w("#define PMG_PUCK_SPAWN_Z  0.17  /* kuka_single_step_base_env.py:56 */")
    # Order in which Bullet visits the robot's 18 non-contact constraints (ids 0..8 = joint-limit
    # constraint of dof i, created while the URDF loads; 9..17 = joint motor of dof i-9, created
    # after the load).  btMultiBodyDynamicsWorld copies them in creation order and quick-sorts by
    # island id every step; all 18 share one island, and Bullet's (unstable) quickSort applied to
    # equal keys yields a fixed permutation, reproduced here by running that partition scheme.
    def qs(a, lo, hi):
        i, j = lo, hi
        while True:  # all keys equal: neither inner scan advances
            if i <= j:
                a[i], a[j] = a[j], a[i]
                i += 1
                j -= 1
            if not i <= j:
                break
        if lo < j:
            qs(a, lo, j)
        if i < hi:
            qs(a, i, hi)
    order = list(range(2 * dof))
    qs(order, 0, 2 * dof - 1)
    w("/* visit order of the 18 non-contact constraints: id<9 joint limit of dof id, id>=9 motor of dof id-9 */")
    w("#define PMG_NONCONTACT_ORDER { %s }" % ", ".join(str(i) for i in order))
    w("")
    w("#endif /* PMG_MODEL_CONSTANTS_H */")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write("\n".join(o) + "\n")
    print("wrote", os.path.abspath(OUT))
    for r in rows:
        print("%-30s m=%.4f com=%s I=%s" % (r["name"], r["mass"], r["com"], r["inertia"]))


if __name__ == "__main__":
    sys.exit(main())
