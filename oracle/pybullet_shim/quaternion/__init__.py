"""numpy-quaternion stand-in: the in-scope step path never calls it (kuka.py:215-220 is the
rotation-control branch)."""


def _unavailable(*a, **k):
    raise NotImplementedError("numpy-quaternion is not available; rotation control is out of scope")


from_euler_angles = as_float_array = _unavailable
