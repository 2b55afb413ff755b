"""GPU parity through the OUTER boundary: manifest -> dlopen -> vk_icdGetInstanceProcAddr -> the Vulkan entry points
(vkCreateGraphicsPipelines / vkCmdBind* / vkCmdDraw* / vkQueueSubmit / mapped host-coherent memory), driven by the
loader-harness, byte-compared with the CPU oracle on the same inputs (configs C1, C2 and reduced C3 / C4)."""
import numpy as np
import pytest

from cpvulkan_b200 import scenes

pytestmark = pytest.mark.gpu


def check(scene, tmp_path, frames=1):
    oc, od, _ = scenes.run_oracle(scene)
    gc, gd, info = scenes.run_icd(scene, str(tmp_path), frames=frames)
    assert np.array_equal(oc, gc), "colour read back through the ICD differs from the oracle"
    if od is not None:
        assert np.array_equal(od, gd), "depth read back through the ICD differs from the oracle"
    return info


def test_icd_draw_cube(built, tmp_path):
    info = check(scenes.draw_cube(), tmp_path)
    assert "B200" in info["device"]


@pytest.mark.parametrize("filt", [scenes.NEAREST, scenes.LINEAR])
def test_icd_draw_textured_cube(built, tmp_path, filt):
    check(scenes.draw_textured_cube(filt=filt), tmp_path)


def test_icd_draw_indexed_mesh(built, tmp_path):
    check(scenes.mesh_indexed(width=640, height=360, nx=160, ny=90), tmp_path)


def test_icd_blended_overdraw(built, tmp_path):
    check(scenes.overdraw_quads(width=96, height=64, quads=12, tex_size=32), tmp_path)


def test_icd_frame_loop_rewrites_and_reads_mapped_memory(built, tmp_path):
    # several frames: vertex data rewritten through a mapping each frame, result read through a persistent mapping
    info = check(scenes.draw_cube(200, 120), tmp_path, frames=5)
    assert info["frames"] == 5


def test_icd_texel_buffer_sample(built, tmp_path):
    # Samples/texel_buffer: vkCreateBufferView + UNIFORM_TEXEL_BUFFER descriptor, texelFetch in the vertex shader
    check(scenes.texel_buffer(200, 120), tmp_path)


@pytest.mark.parametrize("w,h,fmt,filt", [(150, 90, 37, 0), (300, 200, 97, 1), (64, 64, 44, 1)])
def test_icd_blit_and_copy_after_the_render_pass(built, tmp_path, w, h, fmt, filt):
    """Samples/copy_blit_image through the ICD: vkCmdBlitImage (scaling + format conversion) of the rendered image,
    vkCmdCopyImage of the result, vkCmdCopyImageToBuffer readback — against the oracle's blit of the oracle's frame."""
    import ctypes as C
    from cpvulkan_b200 import capi
    scene = scenes.draw_cube(200, 120)
    oc, _, _ = scenes.run_oracle(scene)
    gc, _, info = scenes.run_icd(scene, str(tmp_path), blit=(w, h, fmt, filt))
    assert np.array_equal(oc, gc)
    lib = capi.load_oracle()
    texel = {37: 4, 44: 4, 97: 8}[fmt]
    src = np.ascontiguousarray(oc)
    dst = np.zeros(w * h * texel, dtype=np.uint8)
    b = capi.Blit(capi.Attachment(src.ctypes.data, 200, 120, 200 * 4, scene.color.format), capi.Attachment(dst.ctypes.data, w, h, w * texel, fmt),
                  0, 0, 200, 120, 0, 0, w, h, filt)
    assert lib.cpvk_oracle_blit(C.byref(b)) == 0
    assert np.array_equal(info["blit"], dst)


@pytest.mark.parametrize("flags", [("--indirect",), ("--indirect-count",), ("--secondary",), ("--update-buffers",), ("--indirect", "--secondary", "--update-buffers")],
                         ids=lambda f: "+".join(x.strip("-") for x in f))
@pytest.mark.parametrize("indexed", [False, True])
def test_icd_other_ways_to_issue_the_same_draw(built, tmp_path, flags, indexed):
    """vkCmdDraw[Indexed]Indirect, vkCmdExecuteCommands of a secondary command buffer, and uniform data delivered by
    vkCmdFillBuffer + vkCmdUpdateBuffer must produce the frame of the plain draw."""
    scene = scenes.mesh_indexed(width=160, height=90, nx=40, ny=22) if indexed else scenes.draw_cube(160, 90)
    oc, od, _ = scenes.run_oracle(scene)
    gc, gd, _ = scenes.run_icd(scene, str(tmp_path), flags=flags)
    assert np.array_equal(oc, gc) and np.array_equal(od, gd)


def test_icd_clear_attachments_rectangle(built, tmp_path):
    """vkCmdClearAttachments after the draw: the rectangle of colour and depth takes the clear values, the rest is kept."""
    import ctypes as C
    from cpvulkan_b200 import capi
    scene = scenes.draw_cube(160, 90)
    oc, od, _ = scenes.run_oracle(scene)
    gc, gd, _ = scenes.run_icd(scene, str(tmp_path), flags=("--clear-rect", 20, 10, 70, 45))
    lib = capi.load_oracle()
    oc, od = np.ascontiguousarray(oc), np.ascontiguousarray(od)
    cv = capi.ClearValue()
    for i, v in enumerate((0.5, 0.25, 0.75, 1.0)):
        cv.float32[i] = v
    sub = capi.Attachment(oc.ctypes.data + 10 * scene.color.pitch + 20 * 4, 70, 45, scene.color.pitch, scene.color.format)
    assert lib.cpvk_oracle_clear(C.byref(sub), C.byref(cv), 0) == 0
    dv = capi.ClearValue(); dv.depthStencil.depth, dv.depthStencil.stencil = 0.5, 0
    dsub = capi.Attachment(od.ctypes.data + 10 * scene.depth.pitch + 20 * 2, 70, 45, scene.depth.pitch, scene.depth.format)
    assert lib.cpvk_oracle_clear(C.byref(dsub), C.byref(dv), 1) == 0
    assert np.array_equal(oc, gc) and np.array_equal(od, gd)


@pytest.mark.parametrize("topology", [scenes.POINT_LIST, scenes.LINE_LIST, scenes.LINE_STRIP])
def test_icd_points_and_lines(built, tmp_path, topology):
    check(scenes.random_points_lines(topology=topology, count=50, seed=9, line_width=2.5), tmp_path)


@pytest.mark.parametrize("filt", [scenes.NEAREST, scenes.LINEAR])
@pytest.mark.parametrize("immutable", [False, True])
def test_icd_separate_image_sampler(built, tmp_path, filt, immutable):
    # Samples/separate_image_sampler: SAMPLED_IMAGE + SAMPLER descriptors combined by OpSampledImage in the shader
    # (ImageCombine, GlslFunctions.cpp:812-820); with immutable=True the sampler comes from the set layout instead of a write
    check(scenes.separate_image_sampler(200, 160, filt, immutable), tmp_path)


@pytest.mark.parametrize("filt", [scenes.NEAREST, scenes.LINEAR])
def test_icd_immutable_sampler(built, tmp_path, filt):
    # Samples/immutable_sampler: pImmutableSamplers in the layout, image_info.sampler = 0 in the write (DescriptorSet.cpp:38-48, :79-101)
    check(scenes.immutable_sampler(200, 160, filt), tmp_path)


def test_icd_input_attachment(built, tmp_path):
    # Samples/input_attachment: INPUT_ATTACHMENT descriptor read with subpassLoad (OpImageRead -> @Image.Read = ImageFetch at the
    # coordinate as written, GlslFunctions.cpp:739-743)
    check(scenes.input_attachment(200, 160), tmp_path)


@pytest.mark.parametrize("iterations,threshold,scale", [(4, 0.9, None), (9, 100.0, 1.25)])
def test_icd_push_constants_and_specialization(built, tmp_path, iterations, threshold, scale):
    # Samples/push_constants + Samples/spirv_specialization: vkCmdPushConstants (CommandBuffer.cpp:552-556) and
    # VkSpecializationInfo on the fragment stage, feeding the loop / function-call / OpKill shader of the front-end breadth test
    import struct
    sc = scenes.random_triangles(width=64, height=48, tris=120, seed=81)
    sc.fs = "complex.frag"
    sc.push_constants = struct.pack("<4fif", 0.3, 0.1, 0.2, 0.05, iterations, threshold)
    if scale is not None:
        sc.spec_constants["fragment"] = [(3, struct.unpack("<I", struct.pack("<f", scale))[0])]
    check(sc, tmp_path)


def test_icd_dynamic_uniform_offset(built, tmp_path):
    # Samples/dynamic_uniform: UNIFORM_BUFFER_DYNAMIC whose offset arrives with vkCmdBindDescriptorSets (Binding.cpp:58-80); the
    # matrix the draw must use sits 256 bytes into the buffer, behind a decoy
    sc = scenes.draw_cube(200, 160)
    mvp = sc.buffers["ubo"]
    decoy = np.zeros(256, dtype=np.uint8)
    sc.buffers["ubo"] = np.concatenate([decoy, mvp, np.zeros(192, dtype=np.uint8)])
    sc.uniform_dynamic["ubo"] = (256, 64)
    info = check(sc, tmp_path)
    oc, _, _ = scenes.run_oracle(scenes.draw_cube(200, 160))
    gc, _, _ = scenes.run_icd(sc, str(tmp_path))
    assert np.array_equal(oc, gc), "the dynamic offset must select the same matrix the plain cube uses"


def test_icd_multiple_descriptor_sets(built, tmp_path):
    """Samples/multiple_sets: the uniform buffer in set 0, the sampler in set 1 — two set layouts in the pipeline layout, one
    vkCmdBindDescriptorSets per set with firstSet = its number (Binding.cpp:58-80, LoadUniforms Draw.cpp:356-408)."""
    check(scenes.multiple_sets(filt=scenes.LINEAR), tmp_path)


def test_icd_uniform_arrays_in_a_second_set(built, tmp_path):
    """A second uniform buffer in set 1 with std140 arrays (vec4[4] indexed dynamically, float[3] at ArrayStride 16:
    SPIRVCompiler.cpp:104-184)."""
    check(scenes.ubo_arrays(), tmp_path)
