"""Generates tests/golden/ref_*.{txt,npz}: outputs of the REFERENCE's own code for the pieces of the path that execute
in this image — CPVulkanBase/Formats.cpp (format table, image layout, GetImagePixelOffset), CPVulkanBase/FloatFormat.h
(half <-> float, the codec behind R16G16B16A16_SFLOAT) and CPVulkan/ImageSampler.cpp (the texture sampler) and the rasteriser / interpolator /
blend functions of CPVulkan/CommandBuffer.Draw.cpp — compiled in
place into oracle/_ref/formats_check, oracle/_ref/sampler_check and oracle/_ref/draw_check (oracle/Makefile, oracle/ref_*_check.cpp). Run in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_ref_golden.py

tests/test_reference_formats.py checks the oracle against these files everywhere, and — where oracle/_ref exists — that
the files still equal what the reference binary prints."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CHECK = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "formats_check")


def run(mode, stdin=None):
    return subprocess.run([CHECK, mode], input=stdin, stdout=subprocess.PIPE, text=True, check=True).stdout


def tohalf_inputs():
    """float32 bit patterns around every half rounding boundary: for each pair of adjacent non-negative half codes the
    exact midpoint (a tie), one ulp either side, and the end points; both signs; plus specials."""
    h = np.arange(0, 0x7C00, dtype=np.uint32)
    lo = np.array([np.float32(np.array([c], dtype=np.uint16).view(np.float16)[0]) for c in (0, 1)], dtype=np.float32)  # warm-up for numpy
    f = h.astype(np.uint16).view(np.float16).astype(np.float32)
    nxt = (h + 1).astype(np.uint16).view(np.float16).astype(np.float32)  # 0x7C00 -> inf
    fb = f.view(np.uint32).astype(np.uint64)
    nb = np.where(np.isinf(nxt), np.float32(65536.0).view(np.uint32), nxt.view(np.uint32)).astype(np.uint64)
    mid = ((fb + nb) // 2).astype(np.uint32)  # same-binade neighbours: the bit average is the arithmetic midpoint
    mid = np.where(f == 0, (np.float32(2.0 ** -25)).view(np.uint32), mid)
    sel = (h % 3 == 0) | (h < 1100) | (h > 0x7B00)
    pats = np.concatenate([fb[sel].astype(np.uint32), mid[sel], mid[sel] - 1, mid[sel] + 1])
    specials = np.array([0x00000000, 0x80000000, 0x7F800000, 0xFF800000, 0x7FC00000, 0x7F800001, 0x7FA12345, 0xFFC00001, 0x7FFFFFFF,
                         0x00000001, 0x007FFFFF, 0x00800000, 0x33000000, 0x33000001, 0x32FFFFFF, 0x33800000, 0x477FE000, 0x477FEFFF,
                         0x477FF000, 0x477FF001, 0x47800000, 0x7F7FFFFF, 0x38800000, 0x387FFFFF, 0x387FE000, 0x387FF000], dtype=np.uint32)
    pats = np.concatenate([pats, pats | 0x80000000, specials])
    return np.unique(pats)


SAMPLER_CHECK = os.path.join(os.path.dirname(CHECK), "sampler_check")


def sampler_inputs():
    """A 3-level RGBA32F mip chain (8x4, 4x2, 2x1) of awkward floats, sampler configurations covering every address mode,
    both filters on the magnification and minification paths, both mipmap modes, fractional LODs, three border colours and
    mixed U/V modes, and coordinates that include negatives, exact 0 / 1, texel centres and edges."""
    rng = np.random.RandomState(20261017)
    levels = []
    for w, h in ((8, 4), (4, 2), (2, 1)):
        t = (rng.uniform(-4.0, 4.0, size=(h, w, 4)) * rng.choice([1.0, 1e-3, 257.0], size=(h, w, 4))).astype(np.float32)
        levels.append(t)
    NEAREST, LINEAR = 0, 1
    cfg = []
    for mode in range(5):
        for filt in (NEAREST, LINEAR):
            for border in ((0, 2, 4) if mode == 3 else (0,)):
                cfg.append((filt, filt, 0, mode, mode, border, 0.0))            # lod <= 0: magnification filter, level 0
    for u_mode, v_mode in ((0, 2), (1, 3), (4, 0), (3, 1)):
        cfg.append((LINEAR, NEAREST, 0, u_mode, v_mode, 2, 0.0))
    for lod in (0.25, 0.5, 1.0, 1.5, 2.0, 3.7):
        for mip in (0, 1):
            for filt in (NEAREST, LINEAR):
                for mode in (0, 1, 2):
                    cfg.append((1 - filt, filt, mip, mode, mode, 0, lod))       # lod > 0: minification filter, mip selection
    cfg = np.array(cfg, dtype=[("mag", "<u4"), ("min", "<u4"), ("mipmap", "<u4"), ("au", "<u4"), ("av", "<u4"), ("border", "<u4"), ("lod", "<f4")])
    us = np.array([-1.25, -0.5 / 8, 0.0, 0.5 / 8, 0.37, 1 - 1e-7, 1.0, 2.3], dtype=np.float32)
    grid = np.array([[u, v] for u in us for v in us], dtype=np.float32)
    coords = np.concatenate([grid, rng.uniform(-2.0, 3.0, size=(64, 2)).astype(np.float32)])
    return levels, cfg, coords


def sampler_golden():
    import tempfile
    levels, cfg, coords = sampler_inputs()
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        f.write(np.array([len(levels), len(cfg), len(coords)], dtype="<u4").tobytes())
        for t in levels:
            f.write(np.array([t.shape[1], t.shape[0]], dtype="<u4").tobytes())
            f.write(t.tobytes())
        f.write(cfg.tobytes())
        f.write(coords.tobytes())
        f.flush()
        out = subprocess.run([SAMPLER_CHECK, f.name], stdout=subprocess.PIPE, text=True, check=True).stdout
    bits = np.array([[int(x) for x in l.split()] for l in out.splitlines()], dtype=np.uint32).reshape(len(cfg), len(coords), 4)
    np.savez_compressed(os.path.join(HERE, "ref_sampler.npz"), level0=levels[0], level1=levels[1], level2=levels[2], configs=cfg, coords=coords, result_bits=bits)
    print("sampler: %d configurations x %d coordinates" % (len(cfg), len(coords)))


def sampler3d_golden():
    """The 3-D SampleImage overload vkCmdBlitImage goes through, on a 6x5x1 image (what a 2-D blit source is) and a 4x3x2 one."""
    import tempfile
    rng = np.random.RandomState(77)
    out = {}
    for tag, (w, h, d) in (("a", (6, 5, 1)), ("b", (4, 3, 2))):
        tex = rng.uniform(-3.0, 3.0, size=(d, h, w, 4)).astype(np.float32)
        us = np.array([-0.3, 0.0, 0.5 / w, 0.37, 0.5, 1 - 1e-7, 1.0, 1.4], dtype=np.float32)
        coords = np.array([[u, v, z] for u in us for v in us[1:6] for z in (0.0, 0.25, 0.5, 0.75, 1.0)], dtype=np.float32)
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            f.write(np.array([w, h, d, len(coords)], dtype="<u4").tobytes()); f.write(tex.tobytes()); f.write(coords.tobytes()); f.flush()
            lines = subprocess.run([SAMPLER_CHECK, "3d", f.name], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
        bits = np.array([[int(x) for x in l.split()] for l in lines], dtype=np.uint32).reshape(2, len(coords), 4)
        out["tex_" + tag], out["coords_" + tag], out["bits_" + tag] = tex, coords, bits
    np.savez_compressed(os.path.join(HERE, "ref_sampler3d.npz"), **out)
    print("sampler 3d: %d + %d coordinates x 2 filters" % (len(out["coords_a"]), len(out["coords_b"])))


DRAW_CHECK = os.path.join(os.path.dirname(CHECK), "draw_check")


def run_draw_check(mode, payload):
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(src, "wb") as f:
            f.write(payload)
        subprocess.run([DRAW_CHECK, mode, src, dst], check=True)
        with open(dst, "rb") as f:
            return f.read()


def draw_golden():
    """The reference's CalculatePrimitives / ProcessTriangles / ProcessLines / ProcessPoints / GetFragmentInput / SetDatum /
    DrawPixel (fragment streams) and ApplyBlend, run by oracle/_ref/draw_check on the seeded cases of tests/ref_draw_cases.py.
    Cases flagged `full` keep every fragment record; the others keep the fragment count and the SHA-256 of the stream. NaN bit patterns are canonicalised first
    (rc.canonical: sign and payload of a NaN differ between any two builds)."""
    import hashlib
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    import ref_draw_cases as rc
    cases = rc.raster_cases()
    streams = [rc.canonical(t) for t in rc.parse_raster_output(run_draw_check("raster", rc.raster_file(cases)), len(cases))]
    out = {"names": np.array([c.name for c in cases]), "counts": np.array([len(s) for s in streams], dtype=np.uint64),
           "sha256": np.array([hashlib.sha256(np.ascontiguousarray(s).tobytes()).hexdigest() for s in streams])}
    for i, (c, s) in enumerate(zip(cases, streams)):
        assert s.shape[1] == rc.WORDS or len(s) == 0
        if c.full:
            out["stream_%d" % i] = s
    states, operands = rc.blend_cases()
    res = np.frombuffer(run_draw_check("blend", rc.blend_file(states, operands)), dtype="<u4").reshape(len(states), 4)
    out["blend_bits"] = rc.canonical_floats(res)
    np.savez_compressed(os.path.join(HERE, "ref_draw.npz"), **out)
    total = int(sum(len(s) for s in streams))
    full = int(sum(len(s) for c, s in zip(cases, streams) if c.full))
    print("draw: %d raster cases, %d fragments (%d stored in full), %d blend cases" % (len(cases), total, full, len(states)))
    for c, s in zip(cases, streams):
        print("   %-28s %8d fragments%s" % (c.name, len(s), "  (full)" if c.full else ""))


MATH_CHECK = os.path.join(os.path.dirname(CHECK), "math_check")


def math_golden():
    """The reference's shader runtime math (LLVMRuntime/SpirvFunctions.cpp through its name table, the GLSL.std.450 templates of
    CPVulkan/GlslFunctions.cpp:19-321), run by oracle/_ref/math_check on the seeded cases of tests/ref_math_cases.py."""
    import sys
    import tempfile
    sys.path.insert(0, os.path.dirname(HERE))
    import ref_math_cases as mc
    hdr, A, B, C = mc.cases()
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(src, "wb") as f:
            f.write(mc.file_bytes(hdr, A, B, C))
        subprocess.run([MATH_CHECK, src, dst], check=True)
        out = np.fromfile(dst, dtype="<u4").reshape(len(hdr), 16)
    # float results only are canonicalised for NaN (integer lanes are compared as they are)
    isf = hdr[:, 2] == 0
    out[isf] = mc.canonical(out[isf])
    np.savez_compressed(os.path.join(HERE, "ref_math.npz"), result_bits=out)
    print("math: %d cases" % len(hdr))


def image_golden():
    """The shader-side image functions (GlslFunctions.cpp:324-737) run by oracle/_ref/image_check on tests/ref_image_cases.py."""
    sys.path.insert(0, os.path.dirname(HERE))
    import ref_image_cases
    cs = ref_image_cases.cases()
    check = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "image_check")
    with tempfile.TemporaryDirectory() as tmp:
        inp, outp = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        open(inp, "wb").write(ref_image_cases.payload(cs))
        subprocess.run([check, inp, outp], check=True)
        raw = np.fromfile(outp, dtype=np.uint32)
    out, off = {"count": np.array(len(cs))}, 0
    for i, (hdr, f3, data, coords) in enumerate(cs):
        out["result_%d" % i] = raw[off:off + 4 * len(coords)].reshape(-1, 4).copy()
        off += 4 * len(coords)
    assert off == len(raw)
    np.savez_compressed(os.path.join(HERE, "ref_image.npz"), **out)
    print("image: %d cases, %d results" % (len(cs), off // 4))


def ia_golden():
    """Input assembly: the (rawId, vertexId) pairs ProcessInputAssembler[Indexed] (draw_check ia) produce for tests/ref_ia_cases.py."""
    sys.path.insert(0, os.path.dirname(HERE))
    import ref_ia_cases
    cs = ref_ia_cases.cases()
    pairs = ref_ia_cases.parse(run_draw_check("ia", ref_ia_cases.payload(cs)), len(cs))
    out = {"count": np.array(len(cs))}
    for i, p in enumerate(pairs):
        out["pairs_%d" % i] = p
    np.savez_compressed(os.path.join(HERE, "ref_ia.npz"), **out)
    print("ia: %d cases, %d vertices" % (len(cs), sum(len(p) for p in pairs)))


def blit_golden():
    """vkCmdBlitImage: the destinations oracle/_ref/blit_check (BlitImageCommand::Process compiled in place) produces for the
    seeded cases of tests/ref_blit_cases.py."""
    sys.path.insert(0, os.path.dirname(HERE))
    import ref_blit_cases
    cs = ref_blit_cases.cases()
    check = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "blit_check")
    with tempfile.TemporaryDirectory() as tmp:
        inp, outp = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        open(inp, "wb").write(ref_blit_cases.payload(cs))
        subprocess.run([check, inp, outp], check=True)
        raw = np.fromfile(outp, dtype=np.uint32)
    out, off = {"count": np.array(len(cs))}, 0
    for i, (h, src, dst) in enumerate(cs):
        out["dst_%d" % i] = raw[off:off + dst.size].copy()
        off += dst.size
    assert off == len(raw)
    np.savez_compressed(os.path.join(HERE, "ref_blit.npz"), **out)
    print("blit: %d cases, %d destination texels" % (len(cs), off // 4))


def main():
    image_golden()
    ia_golden()
    blit_golden()
    draw_golden()
    math_golden()
    sampler_golden()
    sampler3d_golden()
    open(os.path.join(HERE, "ref_formats.txt"), "w").write(run("formats"))
    open(os.path.join(HERE, "ref_layout.txt"), "w").write(run("layout"))
    half = np.array([int(l.split()[2]) for l in run("half").splitlines()], dtype=np.uint32)
    assert half.shape == (65536,)
    pats = tohalf_inputs()
    out = run("tohalf", "\n".join(str(int(p)) for p in pats) + "\n").splitlines()
    codes = np.array([int(l.split()[2]) for l in out], dtype=np.uint16)
    assert codes.shape == pats.shape
    np.savez_compressed(os.path.join(HERE, "ref_half.npz"), half_to_float_bits=half, float_bits=pats, half_codes=codes)
    print("formats %d lines, half table %d, float->half vectors %d" % (130, len(half), len(pats)))


if __name__ == "__main__":
    main()
