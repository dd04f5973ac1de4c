#!/usr/bin/env python3
"""Can arm link 7 -- the one robot geometry next to the gripper base that is NOT a collision shape in the oracle and the
kernels -- touch a block before the gripper base does?  (missing item 5 / next step 7: "model it, or prove
with a soak over C5 rollouts that penetration never occurs (report max overlap)".)

Two parts, both on the CPU:

1. Geometry.  Link 7's collision shape is the convex hull of meshes/iiwa14/collision/link_7.stl
   (iiwa14_parallel_jaw.urdf:300-306); the gripper base is a cylinder r 0.05 x 0.04 fixed 0.055 up link 7's z axis
   (:397-410).  The hull's (r, z) profile is read from the mesh when /root/reference is present (build container), else the
   profile recorded below is used.  It tells where link 7 is wider than the base.
2. Soak.  Oracle rollouts of the multi-block scenes (random policy, and scripted sweeps of the gripper through
   pre-built 3-, 4- and 5-block stacks at every height the workspace allows); after every env.step the signed distance
   of every block to link 7's hull of revolution and to the gripper base cylinder is evaluated (block surface sampled
   on a 3 mm grid, exact point-to-solid distances).  Reported: the deepest overlap of a block with link 7, and the
   deepest overlap *while the base is clear of that block* (only that could change a trajectory: once the base touches,
   it is the contact that stops the arm).

Writes nothing; prints the report (kept as profiles/r02_link7_clearance.txt).  Test infrastructure: uses oracle/.
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pmg_oracle as O  # noqa: E402

STL = "/root/reference/pybullet_multigoal_gym/assets/robots/kuka/meshes/iiwa14/collision/link_7.stl"
# convex (r, z) profile of the hull in link 7's frame, as read from the mesh in the build container (fallback)
PROFILE_FALLBACK = [(0.0, -0.0099), (0.0250, -0.0099), (0.0516, -0.0050), (0.0519, 0.0011), (0.0513, 0.0093), (0.0492, 0.0190),
                    (0.0466, 0.0330), (0.0414, 0.0390), (0.0315, 0.0450), (0.0, 0.0450)]
BASE_R, BASE_Z0, BASE_Z1 = 0.05, 0.035, 0.075   # gripper base cylinder in link 7's frame (joint origin 0.055, length 0.04)
TIP_Z = 0.12                                     # iiwa_gripper_tip_joint, urdf:311-315
BLOCK_HALF = 0.015


def hull_profile():
    """Upper convex hull of the mesh vertices in the (r, z) half plane -> closed convex polygon, counter-clockwise."""
    if not os.path.exists(STL):
        return np.array(PROFILE_FALLBACK), "recorded profile (mesh not on this machine)"
    raw = open(STL, "rb").read()
    n = struct.unpack("<I", raw[80:84])[0]
    tri = np.frombuffer(raw[84:84 + 50 * n], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    v = tri["v"].reshape(-1, 3).astype(np.float64)
    pts = np.stack([np.hypot(v[:, 0], v[:, 1]), v[:, 2]], axis=1)
    pts = np.concatenate([pts, [[0.0, pts[:, 1].min()], [0.0, pts[:, 1].max()]]])
    from scipy.spatial import ConvexHull
    h = ConvexHull(pts)
    return pts[h.vertices], "convex hull of %d mesh vertices (link_7.stl)" % len(v)


def signed_dist_convex(poly, p):
    """Signed distance of points p [n, 2] to a convex polygon (counter-clockwise vertices), negative inside."""
    a, b = poly, np.roll(poly, -1, axis=0)
    e = b - a
    L = np.linalg.norm(e, axis=1)
    keep = L > 1e-12
    a, b, e, L = a[keep], b[keep], e[keep], L[keep]
    nrm = np.stack([e[:, 1], -e[:, 0]], axis=1) / L[:, None]          # outward normals of a ccw polygon
    d_plane = np.einsum("nkj,kj->nk", p[:, None, :] - a[None], nrm)   # [n, edges]
    inside = d_plane.max(axis=1)
    t = np.clip(np.einsum("nkj,kj->nk", p[:, None, :] - a[None], e) / (L * L), 0, 1)
    closest = a[None] + t[..., None] * e[None]
    d_out = np.linalg.norm(p[:, None, :] - closest, axis=2).min(axis=1)
    return np.where(inside <= 0, inside, d_out)


def quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def block_surface():
    g = np.linspace(-BLOCK_HALF, BLOCK_HALF, 11)   # 3 mm grid
    pts = []
    for ax in range(3):
        for s in (-BLOCK_HALF, BLOCK_HALF):
            u, v = np.meshgrid(g, g)
            p = np.zeros((u.size, 3))
            p[:, ax] = s
            p[:, (ax + 1) % 3] = u.ravel()
            p[:, (ax + 2) % 3] = v.ravel()
            pts.append(p)
    return np.unique(np.concatenate(pts), axis=0)


SURF = block_surface()


def distances(env, poly, base_poly):
    """-> per block (signed distance to link 7's hull, signed distance to the gripper base)."""
    st = env.get_state()
    tip = env.link_state(0)
    R7 = quat_to_mat(tip[3:7])
    o7 = tip[:3] - R7 @ np.array([0, 0, TIP_Z])          # link 7's origin
    out = []
    for b in range(env.nb):
        pos, quat = st[46 + 13 * b:49 + 13 * b], st[49 + 13 * b:53 + 13 * b]
        pw = pos + SURF @ quat_to_mat(quat).T
        loc = (pw - o7) @ R7                               # in link 7's frame
        rz = np.stack([np.hypot(loc[:, 0], loc[:, 1]), loc[:, 2]], axis=1)
        out.append((signed_dist_convex(poly, rz).min(), signed_dist_convex(base_poly, rz).min()))
    return out


def soak(poly, base_poly, nb, episodes, mode, rng, report):
    env = O.OracleEnv("block_stack", num_block=nb, seed=int(rng.integers(1 << 30)))
    for ep in range(episodes):
        env.reset()
        if mode != "random":
            # a pre-built stack of `mode` blocks somewhere on the table, the other blocks where they spawned
            st = env.get_state()
            x, y = rng.uniform(-0.60, -0.44), rng.uniform(-0.12, 0.12)
            for k in range(mode):
                st[46 + 13 * k:49 + 13 * k] = [x, y, 0.175 + 0.03 * k]
                st[49 + 13 * k:53 + 13 * k] = [0, 0, 0, 1]
            env.set_state(st)
            side = rng.uniform(0, 2 * np.pi)
            start = np.array([x + 0.11 * np.cos(side), y + 0.11 * np.sin(side), rng.uniform(0.175, 0.175 + 0.03 * mode + 0.02)])
        for t in range(50):
            if mode == "random":
                a = rng.uniform(-1, 1, 4)
            else:
                tip = env.link_state(0)[:3]
                # go to the start point beside the stack at the chosen height, then straight through the stack
                tgt = start if t < 14 else np.array([x - 0.11 * np.cos(side), y - 0.11 * np.sin(side), start[2]])
                a = np.r_[np.clip((tgt - tip) / 0.01, -1, 1), -1.0 if rng.random() < 0.5 else 1.0]
            env.step(a)
            for b, (d7, db) in enumerate(distances(env, poly, base_poly)):
                report["steps"] += 1
                report["min_d7"] = min(report["min_d7"], d7)
                report["min_dbase"] = min(report["min_dbase"], db)
                if d7 < 0:
                    report["link7_overlap_steps"] += 1
                    if db > 0:   # link 7 inside a block the base does not touch
                        report["link7_first_steps"] += 1
                        report["worst_link7_first"] = min(report["worst_link7_first"], d7)


def main():
    poly, src = hull_profile()
    base_poly = np.array([(0.0, BASE_Z0), (BASE_R, BASE_Z0), (BASE_R, BASE_Z1), (0.0, BASE_Z1)])
    print("link 7 hull profile: %s" % src)
    zs = np.linspace(poly[:, 1].min(), poly[:, 1].max(), 2001)
    # radius of the convex profile at height z = the largest r with (r, z) inside
    rr = np.array([max((a[0] + (b[0] - a[0]) * (z - a[1]) / (b[1] - a[1]) for a, b in zip(poly, np.roll(poly, -1, axis=0))
                        if (a[1] - z) * (b[1] - z) <= 0 and a[1] != b[1]), default=0.0) for z in zs])
    for z in np.arange(-0.0075, 0.0451, 0.0075):
        r = rr[np.argmin(np.abs(zs - z))]
        print("  z %+.4f (%.4f above the tip)  r %.4f%s" % (z, TIP_Z - z, r, "   <- wider than the gripper base (0.05)" if r > BASE_R else ""))
    over = zs[rr > BASE_R]
    print("link 7 is wider than the base for z in [%.4f, %.4f] of its frame = %.3f..%.3f above the tip, by at most %.2f mm;"
          % (over.min(), over.max(), TIP_Z - over.max(), TIP_Z - over.min(), 1e3 * (rr.max() - BASE_R)))
    print("next to the base (z = %.3f) its radius is %.4f < 0.05: whatever meets the gripper from below or from the side at the"
          % (BASE_Z0, rr[np.argmin(np.abs(zs - BASE_Z0))]))
    print("base's height meets the base first.  With the tip at the workspace floor (z = 0.175) the wide band is at z >= %.3f:"
          % (0.175 + TIP_Z - over.max()))
    print("above a 4-block stack (top 0.280), within a 5-block stack (top 0.310) only.")
    rng = np.random.default_rng(7)
    for nb, mode, episodes in ((4, "random", 40), (5, "random", 40), (4, 3, 30), (4, 4, 40), (5, 5, 60)):
        rep = dict(steps=0, min_d7=1.0, min_dbase=1.0, link7_overlap_steps=0, link7_first_steps=0, worst_link7_first=0.0)
        soak(poly, base_poly, nb, episodes, mode, rng, rep)
        print("num_block %d, %-22s %5d block-steps: closest block to link 7 %+.4f m, to the base %+.4f m; link 7 overlapping a block: %d "
              "block-steps, of which with the base clear of that block: %d (deepest %.2f mm)"
              % (nb, "random policy" if mode == "random" else "sweeps through a %d-stack" % mode, rep["steps"], rep["min_d7"],
                 rep["min_dbase"], rep["link7_overlap_steps"], rep["link7_first_steps"], -1e3 * rep["worst_link7_first"]))


if __name__ == "__main__":
    main()
