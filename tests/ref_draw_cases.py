"""Seeded inputs for the reference pin of the rasteriser / interpolator / blend core (tests/test_reference_draw.py,
tests/golden/make_ref_golden.py): vertex-stage output records + state for oracle/_ref/draw_check `raster` and for the
oracle's cpvk_oracle_raster_records, and blend states + operands for `blend` / cpvk_oracle_apply_blend.

A raster case is one draw: the records are the packed `{vec4 position, f32 pointSize, f32 clip[1], outputs...}` structs the
vertex wrapper stores (PipelineData.h:4-9, PipelineCompiler.cpp:532-547); the four fragment inputs exercise every branch
of GetFragmentInput / SetDatum (Draw.cpp:816-954): a perspective vec4, a noperspective vec2, a flat uint (the provoking
vertex) and a perspective scalar. Only outputs are stored in tests/golden/; the inputs are regenerated from the seeds here."""
import struct

import numpy as np

POINT_LIST, LINE_LIST, LINE_STRIP, TRIANGLE_LIST, TRIANGLE_STRIP, TRIANGLE_FAN = range(6)
CCW, CW = 0, 1
CULL_NONE, CULL_FRONT, CULL_BACK, CULL_BOTH = 0, 1, 2, 3
PERSPECTIVE, LINEAR, FLAT = 0, 1, 2
R32_UINT, R32_SFLOAT, R32G32_SFLOAT, R32G32B32A32_SFLOAT = 98, 100, 103, 109

STRIDE = 56
INPUTS = np.array([[24, R32G32B32A32_SFLOAT, PERSPECTIVE, 16],
                   [40, R32G32_SFLOAT, LINEAR, 8],
                   [48, R32_UINT, FLAT, 4],
                   [52, R32_SFLOAT, PERSPECTIVE, 4]], dtype=np.uint32)
WORDS = 8 + 4 + 2 + 1 + 1


class RasterCase:
    def __init__(self, name, positions, width, height, topology=TRIANGLE_LIST, front_face=CCW, cull=CULL_NONE, origin_upper=1,
                 dynamic_viewport=1, min_depth=0.0, max_depth=1.0, line_width=1.0, point_size=None, seed=0, full=False):
        self.name, self.width, self.height = name, float(width), float(height)
        self.topology, self.front_face, self.cull = topology, front_face, cull
        self.origin_upper, self.dynamic_viewport = origin_upper, dynamic_viewport
        self.min_depth, self.max_depth, self.line_width = float(min_depth), float(max_depth), float(line_width)
        self.full = full  # keep the whole fragment stream in the golden file (else: count + SHA-256)
        positions = np.asarray(positions, dtype=np.float32).reshape(-1, 4)
        n = len(positions)
        rng = np.random.RandomState(1000 + seed)
        rec = np.zeros((n, STRIDE // 4), dtype=np.uint32)
        rec[:, 0:4] = positions.view(np.uint32)
        ps = np.full(n, 1.0, dtype=np.float32) if point_size is None else np.asarray(point_size, dtype=np.float32)
        rec[:, 4] = ps.view(np.uint32)
        rec[:, 5] = np.float32(0.0).view(np.uint32)
        rec[:, 6:10] = (rng.uniform(-2.0, 2.0, size=(n, 4)) * rng.choice([1.0, 1e-3, 300.0], size=(n, 4))).astype(np.float32).view(np.uint32)
        rec[:, 10:12] = rng.uniform(0.0, 1.0, size=(n, 2)).astype(np.float32).view(np.uint32)
        rec[:, 12] = (np.arange(n, dtype=np.uint64) * 2654435761 % (1 << 32)).astype(np.uint32)
        rec[:, 13] = rng.uniform(-1.0, 1.0, size=n).astype(np.float32).view(np.uint32)
        self.records = np.ascontiguousarray(rec)
        self.vertex_count = n

    def serialise(self):
        head = struct.pack("<5f8I", self.width, self.height, self.min_depth, self.max_depth, self.line_width, self.topology,
                           self.front_face, self.cull, self.origin_upper, self.dynamic_viewport, self.vertex_count, STRIDE, len(INPUTS))
        return head + INPUTS.tobytes() + self.records.tobytes()


def _ndc(rng, n, lo=-1.2, hi=1.2, w=None, z=None):
    p = np.zeros((n, 4), dtype=np.float32)
    p[:, 0:2] = rng.uniform(lo, hi, size=(n, 2))
    p[:, 2] = rng.uniform(0.0, 1.0, size=n) if z is None else z
    p[:, 3] = 1.0 if w is None else w
    if w is not None:  # clip-space positions: the reference divides by w
        p[:, 0:3] *= p[:, 3:4]
    return p.astype(np.float32)


def _small_triangles(rng, n, size, w=None):
    c = rng.uniform(-1.1, 1.1, size=(n, 1, 2))
    p = np.zeros((n, 3, 4), dtype=np.float32)
    p[:, :, 0:2] = c + rng.uniform(-size, size, size=(n, 3, 2))
    p[:, :, 2] = rng.uniform(0.0, 1.0, size=(n, 3))
    p[:, :, 3] = 1.0
    p = p.reshape(-1, 4)
    if w is not None:
        p[:, 3] = w(len(p))
        p[:, 0:3] *= p[:, 3:4]
    return p.astype(np.float32)


def raster_cases():
    cases = []
    rng = np.random.RandomState(20261018)
    # 1. random small triangles: both windings x every cull mode; the first pair keeps its whole stream
    k = 0
    for ff in (CCW, CW):
        for cull in (CULL_NONE, CULL_FRONT, CULL_BACK, CULL_BOTH):
            cases.append(RasterCase("random ff=%d cull=%d" % (ff, cull), _small_triangles(rng, 260, 0.22), 64, 48, front_face=ff, cull=cull,
                                    seed=k, full=(cull == CULL_NONE and ff == CCW)))
            k += 1
    # 2. perspective: w in [0.4, 5], some negative w (behind the eye: the reference does not clip)
    cases.append(RasterCase("perspective", _small_triangles(rng, 300, 0.3, w=lambda n: rng.uniform(0.4, 5.0, size=n)), 80, 60, seed=20, full=True))
    cases.append(RasterCase("negative w", _small_triangles(rng, 120, 0.3, w=lambda n: rng.choice([-2.0, -0.5, 1.0, 3.0], size=n)), 48, 48, seed=21))
    # 3. vertices snapped to pixel centres and pixel corners of the viewport: edges run exactly through sample points
    W, H = 40, 30
    for snap, nm in ((0.5, "centres"), (0.0, "corners")):
        ix = rng.randint(-2, W + 3, size=(200 * 3)); iy = rng.randint(-2, H + 3, size=(200 * 3))
        p = np.zeros((600, 4), dtype=np.float32)
        p[:, 0] = ((ix + snap) / W * 2 - 1).astype(np.float32); p[:, 1] = ((iy + snap) / H * 2 - 1).astype(np.float32)
        p[:, 2] = rng.uniform(0, 1, 600); p[:, 3] = 1
        cases.append(RasterCase("snapped " + nm, p, W, H, seed=30 + int(snap * 2), full=(snap == 0.5)))
    # 4. a shared-edge mesh (grid of quads, two triangles each): pixels on shared edges are emitted by both triangles
    gx, gy = 9, 7
    xs = np.linspace(-1, 1, gx + 1, dtype=np.float32); ys = np.linspace(-1, 1, gy + 1, dtype=np.float32)
    tris = []
    for j in range(gy):
        for i in range(gx):
            a, b, c, d = (xs[i], ys[j]), (xs[i + 1], ys[j]), (xs[i + 1], ys[j + 1]), (xs[i], ys[j + 1])
            tris += [a, b, c, a, c, d]
    p = np.zeros((len(tris), 4), dtype=np.float32); p[:, 0:2] = np.array(tris, dtype=np.float32); p[:, 2] = 0.5; p[:, 3] = 1
    cases.append(RasterCase("shared edges", p, 36, 28, seed=40, full=True))
    cases.append(RasterCase("shared edges cw cull back", p, 36, 28, front_face=CW, cull=CULL_BACK, seed=41))
    # 5. zero-area triangles: collinear and repeated vertices (area == 0: front facing, weights 0/0)
    t = rng.uniform(-1, 1, size=(60, 1)).astype(np.float32)
    a = _ndc(rng, 60, -0.9, 0.9); b = _ndc(rng, 60, -0.9, 0.9)
    mid = (a + (b - a) * t).astype(np.float32); mid[:, 3] = 1
    deg = np.stack([a, b, mid], axis=1).reshape(-1, 4)
    rep = np.stack([a, a, b], axis=1).reshape(-1, 4)
    axis = np.zeros((30, 3, 4), dtype=np.float32)  # exactly collinear: horizontal / vertical through pixel centres
    axis[:, :, 3] = 1; axis[:, :, 2] = 0.25
    yv = ((rng.randint(0, 24, size=30) + 0.5) / 24 * 2 - 1).astype(np.float32)
    axis[:, 0, 0], axis[:, 1, 0], axis[:, 2, 0] = -0.75, 0.5, 0.1
    axis[:, :, 1] = yv[:, None]
    cases.append(RasterCase("zero area", np.concatenate([deg, rep, axis.reshape(-1, 4)]), 32, 24, seed=50, full=True))
    # 6. non-finite and enormous positions, w == 0: the bounding box casts go through cvttss2si (INT_MIN), the edge functions
    #    see inf / NaN (NaN weights pass the `< 0` tests)
    special = np.array([np.inf, -np.inf, np.nan, 1e30, -1e30, 3e9, -3e9, 1e-30, 0.0, -0.0, 0.5, -0.5, 1.0, -1.0], dtype=np.float32)
    p = _small_triangles(rng, 150, 0.5)
    sel = rng.rand(len(p), 4) < 0.12
    p[sel] = rng.choice(special, size=int(sel.sum()))
    p[rng.rand(len(p)) < 0.05, 3] = 0.0
    cases.append(RasterCase("non-finite", p, 24, 20, seed=60, full=True))
    # 7. strips and fans, odd viewport sizes, a depth range, lower-left origin, static viewport
    cases.append(RasterCase("strip", _ndc(rng, 90, -1.0, 1.0), 37, 29, topology=TRIANGLE_STRIP, min_depth=0.125, max_depth=0.75, seed=70, full=True))
    cases.append(RasterCase("fan", _ndc(rng, 40, -1.0, 1.0), 29, 37, topology=TRIANGLE_FAN, origin_upper=0, dynamic_viewport=0, seed=71))
    cases.append(RasterCase("fan cw", _ndc(rng, 40, -1.0, 1.0), 31, 17, topology=TRIANGLE_FAN, front_face=CW, cull=CULL_FRONT, seed=72))
    cases.append(RasterCase("strip perspective", _ndc(rng, 70, -1.0, 1.0, w=rng.uniform(0.5, 3.0, size=70)), 33, 31, topology=TRIANGLE_STRIP, cull=CULL_BACK, seed=73))
    cases.append(RasterCase("ragged counts", _ndc(rng, 5, -1.0, 1.0), 16, 16, seed=74, full=True))      # 5 vertices: one triangle
    cases.append(RasterCase("two vertices", _ndc(rng, 2, -1.0, 1.0), 16, 16, topology=TRIANGLE_STRIP, seed=75, full=True))
    cases.append(RasterCase("no vertices", np.zeros((0, 4)), 16, 16, seed=76, full=True))
    # 8. large triangles (every pixel of the viewport, several times) and a fractional viewport size
    cases.append(RasterCase("large", _ndc(rng, 24, -3.0, 3.0), 96, 64, seed=80))
    cases.append(RasterCase("fractional viewport", _small_triangles(rng, 120, 0.4), 50.5, 37.25, seed=81))
    # 9. bulk: 4 000 random triangles on 160x120
    cases.append(RasterCase("bulk", _small_triangles(rng, 4000, 0.12), 160, 120, cull=CULL_BACK, seed=90))
    # 10. lines (every line tests every pixel of the viewport): widths, strips, degenerate segments, perspective
    cases.append(RasterCase("lines", _ndc(rng, 80, -1.1, 1.1), 40, 30, topology=LINE_LIST, line_width=1.0, seed=100, full=True))
    cases.append(RasterCase("wide lines", _ndc(rng, 60, -1.1, 1.1), 40, 30, topology=LINE_LIST, line_width=3.5, seed=101))
    cases.append(RasterCase("line strip perspective", _ndc(rng, 40, -1.0, 1.0, w=rng.uniform(0.5, 3.0, size=40)), 36, 28, topology=LINE_STRIP,
                            line_width=2.0, min_depth=0.25, max_depth=1.0, origin_upper=0, seed=102, full=True))
    p = _ndc(rng, 24, -0.8, 0.8)
    p[1] = p[0]; p[3, 0:2] = p[2, 0:2]; p[5, 0] = np.inf; p[7, 1] = np.nan; p[9, 3] = 0.0  # zero-length, same xy, non-finite
    cases.append(RasterCase("degenerate lines", p, 20, 16, topology=LINE_LIST, line_width=1.0, seed=103, full=True))
    cases.append(RasterCase("one vertex line", _ndc(rng, 1, -1.0, 1.0), 8, 8, topology=LINE_STRIP, seed=104, full=True))
    # 11. points: sizes from 0 to 9.5, negative, off-screen centres
    n = 160
    sizes = rng.choice([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.25, 7.0, 9.5, -3.0], size=n).astype(np.float32)
    cases.append(RasterCase("points", _ndc(rng, n, -1.2, 1.2), 48, 36, topology=POINT_LIST, point_size=sizes, seed=110, full=True))
    cases.append(RasterCase("points perspective", _ndc(rng, n, -1.0, 1.0, w=rng.uniform(0.5, 2.0, size=n)), 31, 23, topology=POINT_LIST,
                            point_size=rng.uniform(0.25, 6.0, size=n), origin_upper=0, min_depth=0.5, max_depth=0.25, seed=111))
    return cases


def raster_file(cases):
    return struct.pack("<I", len(cases)) + b"".join(c.serialise() for c in cases)


def parse_raster_output(data, ncases):
    """-> list of (nFragments, words) arrays of shape (nFragments, wordsPerFragment)"""
    out, at = [], 0
    for _ in range(ncases):
        n, words = struct.unpack_from("<II", data, at)
        at += 8
        a = np.frombuffer(data, dtype="<u4", count=n * words, offset=at).reshape(n, words)
        at += n * words * 4
        out.append(a)
    assert at == len(data)
    return out


FLOAT_COLUMNS = [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15]  # every word of a fragment record that is a float (14 = the flat uint)


def canonical(stream):
    """NaN results compare as NaN, not by sign / payload: an invalid operation on x86 yields the negative default NaN
    (0xFFC00000) or propagates an operand's payload depending on which operand the compiler put first, and the GPU's
    canonical NaN is 0x7FFFFFFF — no two builds agree on those bits, and nothing downstream distinguishes them."""
    s = np.array(stream, dtype=np.uint32, copy=True).reshape(-1, WORDS)
    for k in FLOAT_COLUMNS:
        col = s[:, k]
        col[(col & 0x7FFFFFFF) > 0x7F800000] = 0x7FC00000
    return s


def canonical_floats(bits):
    b = np.array(bits, dtype=np.uint32, copy=True)
    b[(b & 0x7FFFFFFF) > 0x7F800000] = 0x7FC00000
    return b


# ---- blend ----
FACTORS = list(range(15))  # ZERO .. SRC_ALPHA_SATURATE (SRC1_* abort in the reference)
OPS = list(range(5))       # ADD, SUBTRACT, REVERSE_SUBTRACT, MIN, MAX


def blend_cases():
    """-> (states uint32 [n, 8], operands float32 [n, 12]): every colour factor pair x every op with alpha = colour; every
    alpha factor pair x every alpha op under a different colour op; every (colour op, alpha op) pair; random mixtures.
    Operands rotate through in-range colours, HDR / negative values and non-finite ones."""
    rng = np.random.RandomState(4242)
    states = []
    for op in OPS:
        for s in FACTORS:
            for d in FACTORS:
                states.append([1, s, d, op, s, d, op, 0xF])
    for aop in OPS:
        for s in FACTORS:
            for d in FACTORS:
                cop = (aop + 1 + (s + d) % 4) % 5
                states.append([1, int(rng.randint(15)), int(rng.randint(15)), cop, s, d, aop, 0xF])
    for cop in OPS:
        for aop in OPS:
            for _ in range(12):
                states.append([1] + [int(v) for v in rng.randint(15, size=2)] + [cop] + [int(v) for v in rng.randint(15, size=2)] + [aop, 0xF])
    for _ in range(2500):
        f = rng.randint(15, size=4); o = rng.randint(5, size=2)
        states.append([1, int(f[0]), int(f[1]), int(o[0]), int(f[2]), int(f[3]), int(o[1]), 0xF])
    states = np.array(states, dtype=np.uint32)
    n = len(states)
    operands = rng.uniform(0.0, 1.0, size=(n, 12)).astype(np.float32)
    kind = np.arange(n) % 4
    hdr = kind == 1
    operands[hdr] = (rng.uniform(-3.0, 3.0, size=(int(hdr.sum()), 12)) * rng.choice([1.0, 1e-4, 1e4], size=(int(hdr.sum()), 12))).astype(np.float32)
    special = np.array([np.inf, -np.inf, np.nan, 0.0, -0.0, 1.0, 0.5, 65504.0, 1e-45, 3.4e38], dtype=np.float32)
    sp = kind == 2
    vals = operands[sp]
    mask = rng.rand(*vals.shape) < 0.35
    vals[mask] = rng.choice(special, size=int(mask.sum()))
    operands[sp] = vals
    eq = kind == 3  # ties: source == destination in some lanes (MIN / MAX, x - x)
    vals = operands[eq]
    tie = rng.rand(len(vals), 4) < 0.5
    vals[:, 4:8][tie] = vals[:, 0:4][tie]
    operands[eq] = vals
    return states, np.ascontiguousarray(operands)


def blend_file(states, operands):
    rows = np.concatenate([states.view(np.uint32), operands.view(np.uint32)], axis=1).astype("<u4")
    return struct.pack("<I", len(states)) + rows.tobytes()
