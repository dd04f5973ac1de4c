"""pybullet_utils.bullet_client stand-in: BulletClient forwards attribute access to the shim module
with its own world, like the real one forwards to pybullet with physicsClientId bound."""
import pybullet


class BulletClient(object):
    def __init__(self, connection_mode=None):
        self._world = pybullet._World()
        self._client = self._world.client_id

    def __getattr__(self, name):
        attr = getattr(pybullet, name)
        if callable(attr):
            world = self._world

            def bound(*a, **k):
                return attr(world, *a, **k)
            return bound
        return attr
