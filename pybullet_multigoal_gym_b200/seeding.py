"""gym 0.17.3 `utils/seeding.py` restated (the reference seeds with it, base_env.py:120-122):
np_random(seed) hashes the integer seed with sha512 and feeds the first 8 bytes, as 32-bit
words, to numpy's legacy RandomState.seed == MT19937 init_by_array."""
import hashlib
import struct


def seed_key(seed):
    if not (isinstance(seed, int) and seed >= 0):
        raise ValueError("Seed must be a non-negative integer or omitted, not %r" % (seed,))
    seed = seed % 2 ** 64
    digest = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    digest += b"\0" * (4 - len(digest) % 4)  # gym pads even when already aligned
    words = struct.unpack("%dI" % (len(digest) // 4), digest)
    big = sum(2 ** (32 * i) * v for i, v in enumerate(words))
    if big == 0:
        return [0]
    key = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        key.append(mod)
    return key
