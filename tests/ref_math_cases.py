"""Seeded inputs for the reference pin of the shader runtime math (tests/test_reference_math.py, tests/golden/make_ref_golden.py):
one operation per case, raw 32-bit lanes. header = (kind, op, type, n, 0); kind 0 = GLSL.std.450 instruction `op` on n lanes of
type 0 float / 1 signed / 2 unsigned, 1 = dot(n), 2 = mat(n x n) * scalar, 3 = vec4 * mat4, 4 = mat(n x n) * vec(n), 5 = mat4 * mat4."""
import struct

import numpy as np

FLOAT_GLSL = [4, 13, 14, 26, 37, 40, 43, 46, 79, 80, 81]
SINT_GLSL = [5, 7, 39, 42, 45]
UINT_GLSL = [38, 41, 44]
LIBM = {13, 14, 26}  # sin / cos / pow go through libm on both sides: same library in one process image, not a GPU parity claim


def cases():
    rng = np.random.RandomState(31337)
    hdr, A, B, C = [], [], [], []

    def floats(kind_of):
        v = rng.uniform(-4.0, 4.0, size=16).astype(np.float32)
        if kind_of == 1:
            v = (v * rng.choice([1e-3, 1.0, 1e4], size=16)).astype(np.float32)
        elif kind_of == 2:
            sel = rng.rand(16) < 0.3
            v[sel] = rng.choice(np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1.0, -1.0], dtype=np.float32), size=int(sel.sum()))
        elif kind_of == 3:  # ties between the operands (min / max / clamp edges)
            pass
        return v

    def add(kind, op, ty, n, a, b, c):
        hdr.append((kind, op, ty, n, 0)); A.append(a.view(np.uint32).copy()); B.append(b.view(np.uint32).copy()); C.append(c.view(np.uint32).copy())

    for rep in range(40):
        for op in FLOAT_GLSL:
            for n in (1, 2, 3, 4):
                a, b, c = floats(rep % 3), floats(rep % 3), floats(rep % 3)
                if rep % 4 == 3:
                    b[:2] = a[:2]; c[1:3] = a[1:3]
                if op in (43, 81) and rep % 3 != 2:  # clamp: min <= max as the API requires (std::clamp is undefined otherwise)
                    lo, hi = np.minimum(b, c), np.maximum(b, c); b, c = lo, hi
                if op == 43 and rep % 3 == 2:
                    continue  # std::clamp with NaN bounds is undefined behaviour in the reference itself
                if op == 26:
                    a = np.abs(a) + np.float32(0.01)
                add(0, op, 0, n, a, b, c)
        for op in SINT_GLSL:
            for n in (1, 2, 3, 4):
                a, b, c = (rng.randint(-2**31, 2**31, size=16, dtype=np.int64).astype(np.int32) for _ in range(3))
                if rep % 2:
                    a = rng.randint(-5, 6, size=16).astype(np.int32); b = rng.randint(-5, 6, size=16).astype(np.int32); c = rng.randint(-5, 6, size=16).astype(np.int32)
                if op == 5:
                    a[a == -2**31] = 7  # abs(INT_MIN) overflows in the reference (std::abs): undefined
                if op == 45:
                    lo, hi = np.minimum(b, c), np.maximum(b, c); b, c = lo, hi
                add(0, op, 1, n, a, b, c)
        for op in UINT_GLSL:
            for n in (1, 2, 3, 4):
                a, b, c = (rng.randint(0, 2**32, size=16, dtype=np.int64).astype(np.uint32) for _ in range(3))
                if op == 44:
                    lo, hi = np.minimum(b, c), np.maximum(b, c); b, c = lo, hi
                add(0, op, 2, n, a, b, c)
        for op in (69, 71):
            for n in (2, 3, 4):
                a, b = floats(rep % 2), floats(rep % 2)
                add(0, op, 0, n, a, b, floats(0))
        for n in (2, 3, 4):
            add(1, 0, 0, n, floats(rep % 3), floats(rep % 3), floats(0))
            add(2, 0, 0, n, floats(rep % 2), floats(rep % 2), floats(0))
        add(3, 0, 0, 4, floats(rep % 2), floats(rep % 2), floats(0))
        for n in (3, 4):
            add(4, 0, 0, n, floats(rep % 2), floats(rep % 2), floats(0))
        add(5, 0, 0, 4, floats(rep % 2), floats(rep % 2), floats(0))
    return np.array(hdr, dtype=np.uint32), np.array(A, dtype=np.uint32), np.array(B, dtype=np.uint32), np.array(C, dtype=np.uint32)


def file_bytes(hdr, A, B, C):
    rows = np.concatenate([hdr, A, B, C], axis=1).astype("<u4")
    return struct.pack("<I", len(hdr)) + rows.tobytes()


def canonical(bits):
    b = np.array(bits, dtype=np.uint32, copy=True)
    b[(b & 0x7FFFFFFF) > 0x7F800000] = 0x7FC00000
    return b
