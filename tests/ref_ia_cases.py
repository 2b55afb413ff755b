"""Seeded input-assembly cases for tests/test_reference_ia.py and tests/golden/make_ref_golden.py.
A case = (header, index buffer bytes); header = u32 {indexed, first, count, vertexOffset, indexStride, topology, bindingOffset, bytes}:
vkCmdDraw with a first vertex, vkCmdDrawIndexed with 8 / 16 / 32-bit indices, a first index, a binding offset, positive and negative
vertex offsets (the reference adds them as uint32: Draw.cpp:703), every topology the path draws, ragged and empty counts."""
import numpy as np


def cases():
    rng = np.random.default_rng(675760)
    out = []
    for topology in (0, 1, 2, 3, 4, 5):
        for count in (0, 1, 2, 3, 7, 64):
            out.append((np.array([0, int(rng.integers(0, 1 << 20)), count, 0, 0, topology, 0, 0], dtype=np.uint32), b""))
    out.append((np.array([0, 0xFFFFFFF0, 40, 0, 0, 3, 0, 0], dtype=np.uint32), b""))  # first vertex + i wraps
    for stride, dtype in ((1, np.uint8), (2, np.uint16), (4, np.uint32)):
        for topology in (0, 1, 2, 3, 4, 5):
            for count in (0, 3, 5, 96):
                first = int(rng.integers(0, 9))
                binding_offset = int(rng.integers(0, 5)) * stride
                hi = np.iinfo(dtype).max  # the maximum itself is the restart index: never drawn without primitive restart, excluded here too
                idx = rng.integers(0, min(hi, 1 << 24), size=first + count + 4).astype(dtype)
                data = bytes(binding_offset) + idx.tobytes()
                offset = int(rng.choice([0, 1, 17, 1000, -1, -3, -100000])) & 0xFFFFFFFF
                out.append((np.array([1, first, count, offset, stride, topology, binding_offset, len(data)], dtype=np.uint32), data))
    return out


def payload(cs):
    parts = [np.array([len(cs)], dtype=np.uint32).tobytes()]
    for h, data in cs:
        parts += [h.tobytes(), data]
    return b"".join(parts)


def parse(raw, n):
    words = np.frombuffer(raw, dtype="<u4")
    out, off = [], 0
    for _ in range(n):
        k = int(words[off]); off += 1
        out.append(words[off:off + 2 * k].reshape(k, 2).copy()); off += 2 * k
    assert off == len(words)
    return out
