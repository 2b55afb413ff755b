"""Workload builders for the BASELINE.json configs and for parity fuzzing.

A Scene is a backend-neutral description: pipeline state, named byte buffers, render-target images, one draw.
`materialize()` turns it into the C-ABI PODs (cpvk_cuda.h) given an allocator that places bytes somewhere and
returns an address — host memory for the CPU oracle, HBM for libcpvk_cuda.so. Inputs follow the reference's
samples (Samples/15-draw_cube, Samples/draw_textured_cube, Samples/utils/util_init.cpp) in layout and state;
geometry and matrices are generated here, not copied.
"""
import ctypes as C
import math
import os

import numpy as np

from . import capi, spvasm

SHADER_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shaders")

# VkFormat
R8G8B8A8_UNORM, B8G8R8A8_UNORM, R16G16B16A16_SFLOAT = 37, 44, 97
R32_SFLOAT, R32G32_SFLOAT, R32G32B32_SFLOAT, R32G32B32A32_SFLOAT = 100, 103, 106, 109
D16_UNORM, D32_SFLOAT, D24_UNORM_S8_UINT = 124, 126, 129
# VkPrimitiveTopology / VkCullMode / VkFrontFace / VkCompareOp
POINT_LIST, LINE_LIST, LINE_STRIP = 0, 1, 2
TRIANGLE_LIST, TRIANGLE_STRIP, TRIANGLE_FAN = 3, 4, 5
CULL_NONE, CULL_FRONT, CULL_BACK = 0, 1, 2
FRONT_CCW, FRONT_CW = 0, 1
NEVER, LESS, EQUAL, LESS_OR_EQUAL, GREATER, NOT_EQUAL, GREATER_OR_EQUAL, ALWAYS = range(8)
# VkFilter / VkSamplerAddressMode
NEAREST, LINEAR = 0, 1
REPEAT, MIRRORED_REPEAT, CLAMP_TO_EDGE, CLAMP_TO_BORDER, MIRROR_CLAMP_TO_EDGE = range(5)
# VkBlendFactor / VkBlendOp
BF_ZERO, BF_ONE, BF_SRC_COLOR, BF_ONE_MINUS_SRC_COLOR, BF_DST_COLOR, BF_ONE_MINUS_DST_COLOR, BF_SRC_ALPHA, \
    BF_ONE_MINUS_SRC_ALPHA, BF_DST_ALPHA, BF_ONE_MINUS_DST_ALPHA = range(10)
BO_ADD, BO_SUBTRACT, BO_REVERSE_SUBTRACT, BO_MIN, BO_MAX = range(5)

TEXEL_SIZE = {R8G8B8A8_UNORM: 4, B8G8R8A8_UNORM: 4, R16G16B16A16_SFLOAT: 8, D16_UNORM: 2, D32_SFLOAT: 4, 41: 4, 95: 8, 107: 16, 42: 4, 96: 8, 108: 16,  # ..._UINT / ..._SINT: R8G8B8A8, R16G16B16A16, R32G32B32A32
              D24_UNORM_S8_UINT: 4, R32_SFLOAT: 4, R32G32B32A32_SFLOAT: 16,
              125: 4, 127: 1, 128: 3, 130: 8}  # X8_D24_UNORM_PACK32, S8_UINT, D16_UNORM_S8_UINT, D32_SFLOAT_S8_UINT

_shader_cache = {}


def shader(name):
    """SPIR-V words (numpy uint32) of cpvulkan_b200/shaders/<name>.spvasm."""
    if name not in _shader_cache:
        with open(os.path.join(SHADER_DIR, name + ".spvasm")) as f:
            _shader_cache[name] = np.array(spvasm.assemble(f.read()), dtype=np.uint32)
    return _shader_cache[name]


class Image:
    def __init__(self, fmt, width, height, data=None, clear=None):
        self.format, self.width, self.height = fmt, width, height
        self.pitch = TEXEL_SIZE[fmt] * width  # Stride = texel * width (CPVulkanBase/Formats.cpp:455-483)
        self.data = data      # optional initial bytes (np.uint8, pitch*height)
        self.clear = clear    # ('color', (r,g,b,a)) or ('depth', (d, s)) applied before the draw

    @property
    def nbytes(self):
        return getattr(self, "chain_bytes", self.pitch * self.height)  # chain_bytes: a whole mip chain, levels back to back


class Texture:
    """sampler_binding: None = combined image sampler; an int = texture2D at `binding` + sampler object at `sampler_binding`
    (OpSampledImage in the shader, Samples/separate_image_sampler). immutable: through the ICD the sampler is given in the
    set layout's pImmutableSamplers and the descriptor write carries none (Samples/immutable_sampler)."""

    def __init__(self, binding, image, filt=NEAREST, address=CLAMP_TO_EDGE, set_=0, sampler_binding=None, immutable=False, input_attachment=False):
        self.set, self.binding, self.image, self.filter, self.address = set_, binding, image, filt, address
        self.sampler_binding, self.immutable = sampler_binding, immutable
        self.input_attachment = input_attachment  # bound as VK_DESCRIPTOR_TYPE_INPUT_ATTACHMENT (no sampler), read with subpassLoad


class Scene:
    def __init__(self, name):
        self.name = name
        self.vs = self.fs = None                 # shader names
        self.bindings = []                       # (binding, stride, inputRate)
        self.attributes = []                     # (location, binding, format, offset)
        self.topology = TRIANGLE_LIST
        self.cull, self.front_face = CULL_NONE, FRONT_CCW
        self.depth_test = self.depth_write = False
        self.depth_op = LESS_OR_EQUAL
        self.blend = None                        # dict(src, dst, op, srcA, dstA, opA) or None
        self.write_mask = 0xF
        self.buffers = {}                        # name -> np.uint8 array
        self.vertex_buffers = {}                 # binding -> buffer name
        self.index_buffer, self.index_stride = None, 0
        self.uniforms = []                       # (set, binding, buffer name)
        self.textures = []                       # Texture
        self.texel_buffers = []                  # (set, binding, buffer name, format)
        self.color = None                        # Image
        self.depth = None                        # Image or None
        self.viewport = None                     # (x, y, w, h, minDepth, maxDepth)
        self.count, self.instances, self.first, self.vertex_offset, self.first_instance = 0, 1, 0, 0, 0
        self.push_constants = b""
        self.spec_constants = {"vertex": [], "fragment": []}   # per stage: (constantId, 32-bit value as u32)
        self.uniform_dynamic = {}                # uniform buffer name -> (dynamic offset, range): bound as UNIFORM_BUFFER_DYNAMIC
        self.line_width = 1.0
        self.mutate = None                       # optional callable(Materialized): last-minute edits of the PODs, applied on every backend


class Materialized:
    """ctypes PODs + the keep-alive objects behind their pointers."""

    def __init__(self):
        self.desc = capi.PipelineDesc()
        self.state = capi.DrawState()
        self.addr = {}      # resource name -> address
        self.keep = []
        self.color_attachment = None
        self.depth_attachment = None


def materialize(scene, alloc):
    """alloc(name, nbytes, init_bytes_or_None) -> address (int)."""
    m = Materialized()
    d, s = m.desc, m.state
    vsw, fsw = shader(scene.vs), (shader(scene.fs) if scene.fs else None)
    m.keep += [vsw, fsw]
    d.vertex.spirv = vsw.ctypes.data_as(C.POINTER(C.c_uint32))
    d.vertex.wordCount = len(vsw)
    d.vertex.entryPoint = b"main"
    if fsw is not None:
        d.fragment.spirv = fsw.ctypes.data_as(C.POINTER(C.c_uint32))
        d.fragment.wordCount = len(fsw)
        d.fragment.entryPoint = b"main"
    d.bindingCount = len(scene.bindings)
    for i, (b, stride, rate) in enumerate(scene.bindings):
        d.bindings[i] = capi.VertexBinding(b, stride, rate)
    d.attributeCount = len(scene.attributes)
    for i, a in enumerate(scene.attributes):
        d.attributes[i] = capi.VertexAttribute(*a)
    d.topology = scene.topology
    d.polygonMode, d.cullMode, d.frontFace, d.lineWidth = 0, scene.cull, scene.front_face, scene.line_width
    d.rasterizationSamples = 1
    d.depthTestEnable, d.depthWriteEnable, d.depthCompareOp = int(scene.depth_test), int(scene.depth_write), scene.depth_op
    d.minDepthBounds, d.maxDepthBounds = 0.0, 1.0
    d.colorAttachmentCount = 1
    d.colorFormats[0] = scene.color.format
    bl = d.blend[0]
    bl.colorWriteMask = scene.write_mask
    if scene.blend:
        bl.blendEnable = 1
        bl.srcColorBlendFactor, bl.dstColorBlendFactor, bl.colorBlendOp = scene.blend["src"], scene.blend["dst"], scene.blend["op"]
        bl.srcAlphaBlendFactor = scene.blend.get("srcA", scene.blend["src"])
        bl.dstAlphaBlendFactor = scene.blend.get("dstA", scene.blend["dst"])
        bl.alphaBlendOp = scene.blend.get("opA", scene.blend["op"])
    d.depthStencilFormat = scene.depth.format if scene.depth else 0
    d.dynamicViewport = 1

    for name, data in scene.buffers.items():
        m.addr[name] = alloc(name, data.nbytes, data)
    vp = scene.viewport or (0.0, 0.0, float(scene.color.width), float(scene.color.height), 0.0, 1.0)
    s.viewport = capi.Viewport(*vp)
    for b, name in scene.vertex_buffers.items():
        s.vertexBuffers[b] = m.addr[name]
    if scene.index_buffer:
        s.indexBuffer, s.indexStride = m.addr[scene.index_buffer], scene.index_stride
    s.count, s.instanceCount, s.first = scene.count, scene.instances, scene.first
    s.vertexOffset, s.firstInstance = scene.vertex_offset, scene.first_instance
    nd = 0
    for set_, binding, name in scene.uniforms:
        ds = s.descriptors[nd]
        ds.set, ds.binding, ds.type = set_, binding, capi.DESC_BUFFER
        ds.address, ds.range = m.addr[name], scene.buffers[name].nbytes
        if name in scene.uniform_dynamic:  # LoadUniforms adds the dynamic offset to the descriptor's own (Draw.cpp:379-396)
            off, rng = scene.uniform_dynamic[name]
            ds.address, ds.range = m.addr[name] + off, rng
        nd += 1
    for stage, target in (("vertex", m.desc.vertex), ("fragment", m.desc.fragment)):
        target.specCount = len(scene.spec_constants[stage])
        for i, (cid, value) in enumerate(scene.spec_constants[stage]):
            target.spec[i].constantId, target.spec[i].value = cid, value
    for t in scene.textures:
        ds = s.descriptors[nd]
        img = t.image
        addr = alloc("tex%d" % t.binding, img.nbytes, img.data)
        m.addr["tex%d" % t.binding] = addr
        ds.set, ds.binding, ds.type = t.set, t.binding, capi.DESC_IMAGE
        ds.format, ds.dimensions, ds.levelCount = img.format, 2, 1
        ds.levels[0] = capi.MipLevel(addr, img.width, img.height, 1, 0)
        if t.sampler_binding is not None:  # the image descriptor carries no sampler; a CPVK_DESC_SAMPLER record follows
            nd += 1
            ds = s.descriptors[nd]
            ds.set, ds.binding, ds.type = t.set, t.sampler_binding, capi.DESC_SAMPLER
        sm = ds.sampler
        sm.magFilter = sm.minFilter = t.filter
        sm.addressModeU = sm.addressModeV = sm.addressModeW = t.address
        sm.minLod, sm.maxLod = 0.0, 0.0
        if t.sampler_binding is not None or t.immutable:
            sm.borderColor = 4  # the harness' vkCreateSampler state (FLOAT_OPAQUE_WHITE); never sampled with CLAMP_TO_BORDER here
        nd += 1
    for set_, binding, name, fmt in scene.texel_buffers:
        ds = s.descriptors[nd]
        ds.set, ds.binding, ds.type = set_, binding, capi.DESC_TEXEL_BUFFER
        ds.address, ds.range, ds.format = m.addr[name], scene.buffers[name].nbytes, fmt
        nd += 1
    s.descriptorCount = nd
    pc = scene.push_constants
    s.pushConstantSize = len(pc)
    for i, byte in enumerate(pc):
        s.pushConstants[i] = byte

    caddr = alloc("color", scene.color.nbytes, scene.color.data)
    m.addr["color"] = caddr
    m.color_attachment = capi.Attachment(caddr, scene.color.width, scene.color.height, scene.color.pitch, scene.color.format)
    s.color[0] = m.color_attachment
    if scene.depth:
        ddata = scene.depth.data
        if ddata is None and scene.depth.format == 130:  # D32_SFLOAT_S8_UINT has 3 bytes per texel nothing ever writes: start them at 0 on every backend
            ddata = np.zeros(scene.depth.nbytes, dtype=np.uint8)
        daddr = alloc("depth", scene.depth.nbytes, ddata)
        m.addr["depth"] = daddr
        m.depth_attachment = capi.Attachment(daddr, scene.depth.width, scene.depth.height, scene.depth.pitch, scene.depth.format)
        s.depthStencil = m.depth_attachment
    if scene.mutate:
        scene.mutate(m)
    return m


def clear_value(image):
    cv = capi.ClearValue()
    kind, val = image.clear
    if kind == "depth":
        cv.depthStencil.depth, cv.depthStencil.stencil = val
    elif kind == "color_uint":  # VkClearColorValue.uint32, what an integer attachment reads (ImageSampler.cpp:735-760)
        for i in range(4):
            cv.uint32[i] = val[i]
    else:
        for i in range(4):
            cv.float32[i] = val[i]
    return cv, (1 if kind == "depth" else 0)


# ---------------------------------------------------------------------------------------------------
# matrices (float32, column-major like the sample's glm matrices; Samples/utils/util_init.cpp:1050-1068)

def _perspective(fovy, aspect, near, far):
    t = np.float32(math.tan(fovy / 2.0))
    m = np.zeros((4, 4), dtype=np.float32)  # m[col][row]
    m[0][0] = np.float32(1.0) / (np.float32(aspect) * t)
    m[1][1] = np.float32(1.0) / t
    m[2][2] = -np.float32(far + near) / np.float32(far - near)
    m[2][3] = -1.0
    m[3][2] = -np.float32(2.0 * far * near) / np.float32(far - near)
    return m


def _look_at(eye, center, up):
    eye, center, up = (np.array(v, dtype=np.float32) for v in (eye, center, up))
    f = center - eye
    f = f / np.float32(np.linalg.norm(f))
    s = np.cross(f, up)
    s = s / np.float32(np.linalg.norm(s))
    u = np.cross(s, f)
    m = np.identity(4, dtype=np.float32)
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0], m[3][1], m[3][2] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m.astype(np.float32)


def _mul(a, b):  # column-major product a*b with m[col][row] storage
    return (b.astype(np.float32) @ a.astype(np.float32)).astype(np.float32)


def cube_mvp(width, height):
    fov = math.radians(45.0)
    if width > height:
        fov *= float(height) / float(width)
    proj = _perspective(fov, float(width) / float(height), 0.1, 100.0)
    view = _look_at((-5, 3, -10), (0, 0, 0), (0, -1, 0))
    clip = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 0.5, 0], [0, 0, 0.5, 1]], dtype=np.float32)
    return _mul(_mul(clip, proj), view)


_FACES = [  # (axis, sign): the face x/y/z = +-1; colour per face as in the sample's solid-colour cube
    ((2, +1), (1, 0, 0)), ((2, -1), (0, 1, 0)), ((0, -1), (0, 0, 1)),
    ((0, +1), (1, 1, 0)), ((1, +1), (1, 0, 1)), ((1, -1), (0, 1, 1)),
]


def _cube_faces():
    """36 positions (two triangles per face, clockwise seen from outside in a right-handed frame) + face uv."""
    pos, uv, face = [], [], []
    for fi, ((axis, sign), _) in enumerate(_FACES):
        a1, a2 = (axis + 1) % 3, (axis + 2) % 3
        corners = []
        for (u, v) in ((0, 0), (1, 0), (1, 1), (0, 1)):
            p = [0.0, 0.0, 0.0]
            p[axis] = float(sign)
            p[a1] = 2.0 * u - 1.0
            p[a2] = (2.0 * v - 1.0) * sign  # flip so every face has the same handedness seen from outside
            corners.append((p, (float(u), float(v))))
        for k in (0, 2, 1, 0, 3, 2):
            pos.append(corners[k][0] + [1.0])
            uv.append(corners[k][1])
            face.append(fi)
    return np.array(pos, dtype=np.float32), np.array(uv, dtype=np.float32), face


def _render_targets(scene, width, height, color_fmt, depth_fmt, clear_color, clear_depth=1.0):
    scene.color = Image(color_fmt, width, height, clear=("color", clear_color))
    if depth_fmt:
        scene.depth = Image(depth_fmt, width, height, clear=("depth", (clear_depth, 0)))


def draw_cube(width=500, height=500):
    """C1 = Samples/15-draw_cube: 36-vertex coloured cube, BGRA8 + D16, cull BACK, front CLOCKWISE, LESS_OR_EQUAL."""
    s = Scene("draw_cube")
    s.vs, s.fs = "cube.vert", "cube.frag"
    pos, _, face = _cube_faces()
    col = np.array([list(_FACES[f][1]) + [1.0] for f in face], dtype=np.float32)
    vb = np.concatenate([pos, col], axis=1).astype(np.float32)  # Vertex{vec4 pos, vec4 rgba}, stride 32
    s.buffers["vb"] = vb.view(np.uint8).reshape(-1)
    s.buffers["ubo"] = cube_mvp(width, height).reshape(-1).view(np.uint8)
    s.bindings = [(0, 32, 0)]
    s.attributes = [(0, 0, R32G32B32A32_SFLOAT, 0), (1, 0, R32G32B32A32_SFLOAT, 16)]
    s.vertex_buffers = {0: "vb"}
    s.uniforms = [(0, 0, "ubo")]
    s.cull, s.front_face = CULL_BACK, FRONT_CW
    s.depth_test = s.depth_write = True
    s.count = 36
    _render_targets(s, width, height, B8G8R8A8_UNORM, D16_UNORM, (0.2, 0.2, 0.2, 0.2))
    return s


def checker_texture(size=256, seed=7):
    """Stand-in for Samples/data/lunarg.ppm (256x256 RGBA8, alpha forced to 255 by read_ppm)."""
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, size=(size, size, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:size, 0:size]
    img[((xx // 16 + yy // 16) % 2) == 0, :3] //= 3
    img[..., 3] = 255
    return img


def draw_textured_cube(width=500, height=500, filt=NEAREST):
    """C2 = Samples/draw_textured_cube: VertexUV{vec4 pos, vec2 uv} stride 24, 256^2 RGBA8 texture at binding 1."""
    s = Scene("draw_textured_cube_%s" % ("linear" if filt == LINEAR else "nearest"))
    s.vs, s.fs = "texcube.vert", "texcube.frag"
    pos, uv, _ = _cube_faces()
    vb = np.concatenate([pos, uv], axis=1).astype(np.float32)
    s.buffers["vb"] = vb.view(np.uint8).reshape(-1)
    s.buffers["ubo"] = cube_mvp(width, height).reshape(-1).view(np.uint8)
    s.bindings = [(0, 24, 0)]
    s.attributes = [(0, 0, R32G32B32A32_SFLOAT, 0), (1, 0, R32G32_SFLOAT, 16)]
    s.vertex_buffers = {0: "vb"}
    s.uniforms = [(0, 0, "ubo")]
    tex = checker_texture()
    s.textures = [Texture(1, Image(R8G8B8A8_UNORM, 256, 256, data=tex.reshape(-1)), filt, CLAMP_TO_EDGE)]
    s.cull, s.front_face = CULL_BACK, FRONT_CW
    s.depth_test = s.depth_write = True
    s.count = 36
    _render_targets(s, width, height, B8G8R8A8_UNORM, D16_UNORM, (0.2, 0.2, 0.2, 0.2))
    return s


def multiple_sets(width=500, height=500, filt=LINEAR):
    """Samples/multiple_sets: the textured cube with its uniform buffer in descriptor set 0 and its sampler in set 1, binding 0
    (two set layouts, one vkCmdBindDescriptorSets per set through the ICD)."""
    s = draw_textured_cube(width, height, filt)
    s.name = "multiple_sets_%dx%d" % (width, height)
    s.fs = "multisets.frag"
    t = s.textures[0]
    s.textures = [Texture(0, t.image, filt, CLAMP_TO_EDGE, set_=1)]
    return s


def ubo_arrays(width=500, height=500):
    """The coloured cube with a second uniform buffer, in descriptor set 1, that holds std140 arrays: vec4 tint[4] (indexed
    dynamically) and float scale[3] whose ArrayStride (16) is four times its element size."""
    s = draw_cube(width, height)
    s.name = "ubo_arrays_%dx%d" % (width, height)
    s.vs = "uboarray.vert"
    rng = np.random.RandomState(99)
    params = np.zeros(28, dtype=np.float32)                    # 4 x vec4, then 3 floats at a stride of 4 floats
    params[0:16] = rng.uniform(0.2, 1.0, 16)
    params[16:28] = rng.uniform(-9.0, 9.0, 12)                 # the padding lanes hold garbage a wrong stride would pick up
    params[16], params[20], params[24] = 0.03, 0.07, 0.11      # scale[0..2]
    s.buffers["params"] = params.view(np.uint8).reshape(-1)
    s.uniforms = [(0, 0, "ubo"), (1, 0, "params")]
    return s


def separate_image_sampler(width=500, height=500, filt=LINEAR, immutable=False):
    """Samples/separate_image_sampler: the textured cube with `texture2D tex` at binding 1 and `sampler samp` at binding 2,
    combined in the shader by OpSampledImage (sampler2D(tex, samp)); the fragment shader also darkens a 1 % border."""
    s = draw_textured_cube(width, height, filt)
    s.name = "separate_image_sampler_%d%s" % (filt, "_immutable" if immutable else "")
    s.fs = "sepsampler.frag"
    t = s.textures[0]
    s.textures = [Texture(1, t.image, filt, CLAMP_TO_EDGE, sampler_binding=2, immutable=immutable)]
    return s


def input_attachment(width=500, height=500):
    """Samples/input_attachment: the fragment shader returns subpassLoad() of an image bound as an input attachment. The
    reference reads the coordinate glslang writes — ivec2(0, 0) — so every fragment gets texel (0, 0); the cube geometry of
    C2 stands in for the sample's full-screen triangle so that coverage and depth are exercised too."""
    s = draw_textured_cube(width, height, NEAREST)
    s.name = "input_attachment"
    s.fs = "subpass.frag"
    t = s.textures[0]
    s.textures = [Texture(1, t.image, NEAREST, CLAMP_TO_EDGE, input_attachment=True)]
    return s


def immutable_sampler(width=500, height=500, filt=LINEAR):
    """Samples/immutable_sampler: draw_textured_cube with the sampler baked into the descriptor set layout."""
    s = draw_textured_cube(width, height, filt)
    s.name = "immutable_sampler_%d" % filt
    t = s.textures[0]
    s.textures = [Texture(1, t.image, filt, CLAMP_TO_EDGE, immutable=True)]
    return s


def _grid_mesh(nx, ny, z_of, seed):
    """(nx x ny)-quad grid over NDC [-1,1]^2: vertices {vec4 pos, vec4 rgba}, u32 indices, CCW in framebuffer space."""
    ix, iy = np.meshgrid(np.arange(nx + 1, dtype=np.float32), np.arange(ny + 1, dtype=np.float32))
    x = (np.float32(-1.0) + np.float32(2.0) * ix / np.float32(nx)).astype(np.float32)
    y = (np.float32(-1.0) + np.float32(2.0) * iy / np.float32(ny)).astype(np.float32)
    z = z_of(ix, iy).astype(np.float32)
    pos = np.stack([x, y, z, np.ones_like(x)], axis=-1).reshape(-1, 4)
    rng = np.random.RandomState(seed)
    col = rng.random_sample((pos.shape[0], 4)).astype(np.float32)
    vb = np.concatenate([pos, col], axis=1).astype(np.float32)
    qx, qy = np.meshgrid(np.arange(nx, dtype=np.uint32), np.arange(ny, dtype=np.uint32))
    v00 = (qy * (nx + 1) + qx).reshape(-1)
    v10, v01, v11 = v00 + 1, v00 + (nx + 1), v00 + (nx + 2)
    # framebuffer y grows downwards with NDC y: (v00, v01, v11) and (v00, v11, v10) have positive edge-function area
    idx = np.stack([v00, v01, v11, v00, v11, v10], axis=-1).reshape(-1).astype(np.uint32)
    return vb, idx


def mesh_indexed(width=3840, height=2160, nx=1000, ny=500, layers=1, seed=1234):
    """C3: synthetic indexed mesh, RGBA8 + D32, opaque, LESS_OR_EQUAL. layers=1 -> M1 (2*nx*ny triangles);
    layers=4 with nx=500, ny=250 -> M4 (four stacked grids drawn far to near)."""
    s = Scene("mesh_%dx%dx%d_%dx%d" % (nx, ny, layers, width, height))
    s.vs, s.fs = "cube.vert", "cube.frag"
    vbs, ibs, base = [], [], 0
    for layer in range(layers):
        if layers == 1:
            z_of = lambda ix, iy: np.float32(0.25) + np.float32(0.5) * (ix + iy) / np.float32(nx + ny)
        else:
            zl = np.float32(0.8 - 0.2 * layer)
            z_of = lambda ix, iy, zl=zl: np.full_like(ix, zl)
        vb, idx = _grid_mesh(nx, ny, z_of, seed + layer)
        vbs.append(vb)
        ibs.append(idx + np.uint32(base))
        base += vb.shape[0]
    s.buffers["vb"] = np.concatenate(vbs).view(np.uint8).reshape(-1)
    s.buffers["ib"] = np.concatenate(ibs).view(np.uint8).reshape(-1)
    s.buffers["ubo"] = np.identity(4, dtype=np.float32).reshape(-1).view(np.uint8)
    s.bindings = [(0, 32, 0)]
    s.attributes = [(0, 0, R32G32B32A32_SFLOAT, 0), (1, 0, R32G32B32A32_SFLOAT, 16)]
    s.vertex_buffers = {0: "vb"}
    s.index_buffer, s.index_stride = "ib", 4
    s.uniforms = [(0, 0, "ubo")]
    s.cull, s.front_face = CULL_NONE, FRONT_CCW
    s.depth_test = s.depth_write = True
    s.count = 6 * nx * ny * layers
    _render_targets(s, width, height, R8G8B8A8_UNORM, D32_SFLOAT, (0.0, 0.0, 0.0, 1.0))
    return s


def overdraw_quads(width=7680, height=4320, quads=2000, tex_size=1024, seed=42, blend=True,
                   color_fmt=R16G16B16A16_SFLOAT):
    """C4: alpha-blended full-screen textured quads, RGBA16F, no depth; texture RGBA8 random, LINEAR, REPEAT, uv x4."""
    s = Scene("overdraw_%dq_%dx%d" % (quads, width, height))
    s.vs, s.fs = "texcube.vert", "texcube.frag"
    corner = np.array([[-1, -1, 0, 1, 0, 0], [-1, 1, 0, 1, 0, 4], [1, 1, 0, 1, 4, 4], [1, -1, 0, 1, 4, 0]], dtype=np.float32)
    vb = np.tile(corner, (quads, 1)).astype(np.float32)
    base = (np.arange(quads, dtype=np.uint32) * 4)[:, None]
    idx = (base + np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)[None, :]).reshape(-1).astype(np.uint32)
    s.buffers["vb"] = vb.view(np.uint8).reshape(-1)
    s.buffers["ib"] = idx.view(np.uint8).reshape(-1)
    s.buffers["ubo"] = np.identity(4, dtype=np.float32).reshape(-1).view(np.uint8)
    s.bindings = [(0, 24, 0)]
    s.attributes = [(0, 0, R32G32B32A32_SFLOAT, 0), (1, 0, R32G32_SFLOAT, 16)]
    s.vertex_buffers = {0: "vb"}
    s.index_buffer, s.index_stride = "ib", 4
    s.uniforms = [(0, 0, "ubo")]
    rng = np.random.RandomState(seed)
    tex = rng.randint(0, 256, size=(tex_size, tex_size, 4), dtype=np.uint8)
    s.textures = [Texture(1, Image(R8G8B8A8_UNORM, tex_size, tex_size, data=tex.reshape(-1)), LINEAR, REPEAT)]
    s.cull, s.front_face = CULL_NONE, FRONT_CCW
    if blend:
        s.blend = dict(src=BF_SRC_ALPHA, dst=BF_ONE_MINUS_SRC_ALPHA, op=BO_ADD)
    s.count = 6 * quads
    _render_targets(s, width, height, color_fmt, None, (0.0, 0.0, 0.0, 0.0))
    return s


def sampler_matrix(width=96, height=64, tex_fmt=R8G8B8A8_UNORM, address=(REPEAT, REPEAT), mag=LINEAR, min_=LINEAR, mipmap=0,
                   min_lod=0.0, border=0, tex_size=(8, 4), levels=3, seed=3, bias=0.0, max_lod=1000.0, swizzle=None):
    """One opaque full-screen quad whose uv runs from -1.5 to 2.5 over a small mip-mapped texture, written to an RGBA32F
    target so every sampled float is visible. The sampler state (address modes per axis, mag / min filter, mipmap mode,
    border colour) and the mip chain (levels back to back, Formats.cpp:455-483) are patched into the descriptor on every
    backend; min_lod > 0 forces the minification path without touching the shader: lambda = clamp(0 + bias, minLod, maxLod)
    (GlslFunctions.cpp:598-654)."""
    s = overdraw_quads(width, height, quads=1, tex_size=8, blend=False, color_fmt=R32G32B32A32_SFLOAT)
    s.name = "sampler_matrix"
    vb = s.buffers["vb"].view(np.float32).reshape(-1, 6).copy()
    vb[:, 4:6] = vb[:, 4:6] - 1.5  # 0 / 4 -> -1.5 / 2.5
    s.buffers["vb"] = vb.view(np.uint8).reshape(-1)
    rng = np.random.RandomState(seed)
    texel = TEXEL_SIZE[tex_fmt]
    w, h = tex_size
    chain, dims = [], []
    for _ in range(levels):
        if tex_fmt == R32G32B32A32_SFLOAT:
            data = rng.uniform(-2.0, 2.0, size=(h, w, 4)).astype(np.float32).view(np.uint8).reshape(-1)
        else:
            data = rng.randint(0, 256, size=h * w * texel, dtype=np.uint8)
        chain.append(data); dims.append((w, h))
        w, h = max(w // 2, 1), max(h // 2, 1)
    img = Image(tex_fmt, tex_size[0], tex_size[1], data=np.concatenate(chain))
    img.chain_bytes = sum(len(c) for c in chain)
    s.textures = [Texture(1, img, mag, address[0])]

    def patch(m):
        for i in range(m.state.descriptorCount):
            d = m.state.descriptors[i]
            if d.type != capi.DESC_IMAGE:
                continue
            base, off = d.levels[0].address, 0
            d.levelCount = levels
            for l, (lw, lh) in enumerate(dims):
                d.levels[l] = capi.MipLevel(base + off, lw, lh, 1, 0)
                off += lw * lh * texel
            sm = d.sampler
            sm.magFilter, sm.minFilter, sm.mipmapMode = mag, min_, mipmap
            sm.addressModeU, sm.addressModeV = address
            sm.borderColor = border
            sm.minLod, sm.maxLod, sm.mipLodBias = min_lod, max_lod, bias
            if swizzle is not None:  # VkComponentSwizzle per channel of the view (GlslFunctions.cpp:539-555, :636-651)
                for k, c in enumerate(swizzle):
                    d.swizzle[k] = c
    s.mutate = patch
    return s


def texel_buffer(width=500, height=500, texels=(1.0, 0.0, 1.0), triangles=1):
    """Samples/texel_buffer (BASELINE C5, third item): a uniform texel buffer view of three R32_SFLOAT texels fetched in
    the vertex shader, one triangle from a private array indexed by gl_VertexIndex % 3, no vertex buffers, no depth
    (texel_buffer.cpp:35-62, :71, :286). `triangles` > 1 repeats the same triangle (vertex index modulo 3)."""
    s = Scene("texel_buffer_%dx%d" % (width, height))
    s.vs, s.fs = "texelbuf.vert", "cube.frag"
    s.buffers["texels"] = np.array(texels, dtype=np.float32).view(np.uint8).reshape(-1)
    s.texel_buffers = [(0, 0, "texels", R32_SFLOAT)]
    s.topology = TRIANGLE_LIST
    s.count = 3 * triangles
    s.cull, s.front_face = CULL_NONE, FRONT_CW
    _render_targets(s, width, height, B8G8R8A8_UNORM, None, (0.2, 0.2, 0.2, 0.2))
    return s


def random_points_lines(width=96, height=64, count=60, seed=1, topology=POINT_LIST, line_width=3.0, depth_fmt=D32_SFLOAT,
                        color_fmt=R8G8B8A8_UNORM, perspective=True, fs="cube.frag"):
    """Points / line lists / line strips: random positions (optionally varying w), per-vertex colour
    and point size (1..9 px)."""
    s = Scene("prims_t%d_%d_%dx%d_s%d" % (topology, count, width, height, seed))
    s.vs, s.fs = "points.vert", fs
    rng = np.random.RandomState(seed)
    n = count if topology != LINE_LIST else count * 2
    xy = rng.uniform(-1.05, 1.05, size=(n, 2)).astype(np.float32)
    z = rng.uniform(0.0, 1.0, size=(n, 1)).astype(np.float32)
    w = rng.uniform(0.5, 3.0, size=(n, 1)).astype(np.float32) if perspective else np.ones((n, 1), dtype=np.float32)
    pos = np.concatenate([xy * w, z * w, w], axis=1).astype(np.float32)
    col = rng.random_sample((n, 4)).astype(np.float32)
    size = rng.uniform(1.0, 9.0, size=(n, 1)).astype(np.float32)
    vb = np.concatenate([pos, col, size], axis=1).astype(np.float32)  # 36 bytes per vertex
    s.buffers["vb"] = vb.view(np.uint8).reshape(-1)
    s.bindings = [(0, 36, 0)]
    s.attributes = [(0, 0, R32G32B32A32_SFLOAT, 0), (1, 0, R32G32B32A32_SFLOAT, 16), (2, 0, R32_SFLOAT, 32)]
    s.vertex_buffers = {0: "vb"}
    s.topology, s.count, s.line_width = topology, n, line_width
    s.depth_test = s.depth_write = depth_fmt is not None
    _render_targets(s, width, height, color_fmt, depth_fmt, (0.1, 0.2, 0.3, 1.0))
    return s


def random_triangles(width=256, height=192, tris=200, seed=1, depth_fmt=D32_SFLOAT, color_fmt=R8G8B8A8_UNORM,
                     cull=CULL_NONE, front_face=FRONT_CCW, perspective=True, depth_op=LESS_OR_EQUAL,
                     topology=TRIANGLE_LIST, indexed=None, snap=False):
    """Parity fuzz: random, overlapping, mixed-winding triangles (optionally with varying w and vertices snapped to
    pixel centres so edge-on-centre and shared-edge cases occur)."""
    s = Scene("random_%d_%dx%d_s%d" % (tris, width, height, seed))
    s.vs, s.fs = "cube.vert", "cube.frag"
    rng = np.random.RandomState(seed)
    n = tris * 3 if topology == TRIANGLE_LIST else tris + 2
    xy = rng.uniform(-1.1, 1.1, size=(n, 2)).astype(np.float32)
    if snap:  # put vertices exactly on pixel centres: ((x/W + 0.5/W) * 2 - 1)
        px = rng.randint(0, width, size=n).astype(np.float32)
        py = rng.randint(0, height, size=n).astype(np.float32)
        W, H = np.float32(width), np.float32(height)
        xy[:, 0] = (px / W + (np.float32(1.0) / W) * np.float32(0.5)) * np.float32(2) - np.float32(1)
        xy[:, 1] = (py / H + (np.float32(1.0) / H) * np.float32(0.5)) * np.float32(2) - np.float32(1)
    z = rng.uniform(0.0, 1.0, size=(n, 1)).astype(np.float32)
    w = rng.uniform(0.5, 3.0, size=(n, 1)).astype(np.float32) if perspective else np.ones((n, 1), dtype=np.float32)
    pos = np.concatenate([xy * w, z * w, w], axis=1).astype(np.float32)
    col = rng.random_sample((n, 4)).astype(np.float32)
    vb = np.concatenate([pos, col], axis=1).astype(np.float32)
    s.buffers["vb"] = vb.view(np.uint8).reshape(-1)
    s.buffers["ubo"] = np.identity(4, dtype=np.float32).reshape(-1).view(np.uint8)
    s.bindings = [(0, 32, 0)]
    s.attributes = [(0, 0, R32G32B32A32_SFLOAT, 0), (1, 0, R32G32B32A32_SFLOAT, 16)]
    s.vertex_buffers = {0: "vb"}
    s.uniforms = [(0, 0, "ubo")]
    s.topology = topology
    s.count = n
    if indexed:
        dtype = {1: np.uint8, 2: np.uint16, 4: np.uint32}[indexed]
        perm = rng.permutation(n).astype(dtype) if n <= np.iinfo(dtype).max else np.arange(n, dtype=dtype)
        s.buffers["ib"] = perm.view(np.uint8).reshape(-1)
        s.index_buffer, s.index_stride = "ib", indexed
    s.cull, s.front_face = cull, front_face
    s.depth_test = s.depth_write = depth_fmt is not None
    s.depth_op = depth_op
    _render_targets(s, width, height, color_fmt, depth_fmt, (0.1, 0.2, 0.3, 1.0))
    return s


def large_triangles(width=200, height=136, tris=60, seed=1, scale="mixed", blend=True):
    """Large triangles (bbox far beyond a 16x8 warp region) whose edges cut through regions at every angle: slivers, screen-
    sized triangles with one vertex far outside, both windings, and — with scale="extreme" — vertices at 1e6 .. 1e30, +-inf
    and NaN, where the raster kernel's whole-region rejection must step aside (oracle semantics: NaN edge values pass the
    `w < 0` test, Draw.cpp:879-903). Blended so that every covered fragment changes the result."""
    # extreme vertices give NaN colours, whose payload bits differ between x86 and the GPU: an UNORM target packs them to 0
    s = random_triangles(width=width, height=height, tris=tris, seed=seed, depth_fmt=None,
                         color_fmt=R8G8B8A8_UNORM if scale == "extreme" else R32G32B32A32_SFLOAT)
    s.name = "large_triangles_%s" % scale
    rng = np.random.RandomState(seed + 100)
    n = 3 * tris
    xy = rng.uniform(-1.6, 1.6, size=(n, 2)).astype(np.float32)
    k = np.arange(tris)
    far = rng.choice([1.0, 1.0, 3.0, 40.0, 1e3], size=tris).astype(np.float32)          # push one vertex of some triangles far out
    xy[3 * k] *= far[:, None]
    sliver = k % 5 == 0
    xy[3 * k[sliver] + 1] = xy[3 * k[sliver]] * np.float32(-1.0) + rng.uniform(-0.02, 0.02, size=(sliver.sum(), 2)).astype(np.float32)
    if scale == "extreme":
        big = rng.choice([1e6, 1e12, 1e20, 1e30, np.inf, -np.inf, np.nan], size=tris)
        pick = rng.randint(0, 3, size=tris)
        axis = rng.randint(0, 2, size=tris)
        sel = k % 2 == 0
        xy[3 * k[sel] + pick[sel], axis[sel]] = big[sel].astype(np.float32)
    z = rng.uniform(0.0, 1.0, size=(n, 1)).astype(np.float32)
    pos = np.concatenate([xy, z, np.ones((n, 1), dtype=np.float32)], axis=1).astype(np.float32)
    col = rng.uniform(0.05, 0.9, size=(n, 4)).astype(np.float32)
    s.buffers["vb"] = np.concatenate([pos, col], axis=1).astype(np.float32).view(np.uint8).reshape(-1)
    if blend:
        s.blend = dict(src=BF_SRC_ALPHA, dst=BF_ONE_MINUS_SRC_ALPHA, op=BO_ADD)
    return s


# ---------------------------------------------------------------------------------------------------
# export for the Vulkan loader-harness (cpvulkan_b200/icd/cpvk_harness.cpp)

def export_scene(scene, directory):
    """Write `scene` as scene.txt + raw blobs so the harness can replay it through the Vulkan API of the ICD."""
    os.makedirs(directory, exist_ok=True)
    lines = []
    for tag, name in (("vs", scene.vs), ("fs", scene.fs)):
        fn = name + ".spv"
        shader(name).tofile(os.path.join(directory, fn))
        lines.append("%s %s" % (tag, fn))
    for b in scene.bindings:
        lines.append("binding %d %d %d" % b)
    for a in scene.attributes:
        lines.append("attribute %d %d %d %d" % a)
    lines += ["topology %d" % scene.topology, "cull %d" % scene.cull, "front %d" % scene.front_face,
              "depth_test %d" % int(scene.depth_test), "depth_write %d" % int(scene.depth_write), "depth_op %d" % scene.depth_op,
              "write_mask %d" % scene.write_mask, "line_width %r" % float(np.float32(scene.line_width))]
    if scene.blend:
        bl = scene.blend
        lines.append("blend %d %d %d %d %d %d" % (bl["src"], bl["dst"], bl["op"], bl.get("srcA", bl["src"]), bl.get("dstA", bl["dst"]), bl.get("opA", bl["op"])))
    if scene.push_constants:
        lines.append("push_constants %d %s" % (len(scene.push_constants), bytes(scene.push_constants).hex()))
    for stage, tag in (("vertex", 0), ("fragment", 1)):
        for cid, value in scene.spec_constants[stage]:
            lines.append("spec %d %d %d" % (tag, cid, value))
    for name, data in scene.buffers.items():
        fn = "buf_%s.bin" % name
        np.ascontiguousarray(data).tofile(os.path.join(directory, fn))
        lines.append("buffer %s %s %d" % (name, fn, data.nbytes))
    for b, name in scene.vertex_buffers.items():
        lines.append("vertex_buffer %d %s" % (b, name))
    if scene.index_buffer:
        lines.append("index_buffer %s %d" % (scene.index_buffer, scene.index_stride))
    for set_, binding, name in scene.uniforms:
        dyn = scene.uniform_dynamic.get(name)
        lines.append("uniform %d %d %s" % (set_, binding, name) + (" %d %d" % dyn if dyn else ""))
    for set_, binding, name, fmt in scene.texel_buffers:
        lines.append("texel_buffer %d %d %s %d" % (set_, binding, name, fmt))
    for t in scene.textures:
        fn = "tex_%d.bin" % t.binding
        np.ascontiguousarray(t.image.data).tofile(os.path.join(directory, fn))
        lines.append("texture %d %d %d %d %d %d %d %s %d %d %d" % (t.set, t.binding, t.image.format, t.image.width, t.image.height, t.filter, t.address, fn,
                                                                   -1 if t.sampler_binding is None else t.sampler_binding, int(t.immutable), int(t.input_attachment)))
    cc = scene.color.clear[1] if scene.color.clear else (0, 0, 0, 0)
    lines.append("color %d %d %d %r %r %r %r" % ((scene.color.format, scene.color.width, scene.color.height) + tuple(float(np.float32(c)) for c in cc)))
    if scene.depth:
        d, s = scene.depth.clear[1] if scene.depth.clear else (1.0, 0)
        lines.append("depth %d %r %d" % (scene.depth.format, float(d), int(s)))
    vp = scene.viewport or (0.0, 0.0, float(scene.color.width), float(scene.color.height), 0.0, 1.0)
    lines.append("viewport " + " ".join(repr(float(v)) for v in vp))
    lines.append("draw %d %d %d %d %d" % (scene.count, scene.instances, scene.first, scene.vertex_offset, scene.first_instance))
    with open(os.path.join(directory, "scene.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


def run_icd(scene, workdir, frames=1, env=None, blit=None, flags=()):
    """Render `scene` through the Vulkan ICD (manifest -> vk_icd* -> vkCmdDraw* -> vkQueueSubmit) with the
    loader-harness. Returns (color bytes, depth bytes or None, harness JSON dict)."""
    import json
    import subprocess
    icd_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "icd", "build")
    scene_dir, out_dir = os.path.join(workdir, "scene"), os.path.join(workdir, "out")
    export_scene(scene, scene_dir)
    os.makedirs(out_dir, exist_ok=True)
    e = dict(os.environ)
    e["VK_ICD_FILENAMES"] = os.path.join(icd_dir, "CPVulkan_b200.json")
    if env:
        e.update(env)
    extra = ["--blit"] + [str(v) for v in blit] if blit else []  # (width, height, format, filter): see cpvk_harness.cpp
    extra += [str(f) for f in flags]                              # --indirect, --secondary, --update-buffers, --clear-rect X Y W H
    out = subprocess.run([os.path.join(icd_dir, "cpvk_harness"), scene_dir, out_dir, "--frames", str(frames)] + extra, env=e, check=True,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    info = json.loads(out.stdout.strip().splitlines()[-1])
    color = np.fromfile(os.path.join(out_dir, "color.bin"), dtype=np.uint8)
    depth = np.fromfile(os.path.join(out_dir, "depth.bin"), dtype=np.uint8) if scene.depth else None
    if blit:
        info["blit"] = np.fromfile(os.path.join(out_dir, "blit.bin"), dtype=np.uint8)
    return color, depth, info


# ---------------------------------------------------------------------------------------------------
# host-memory backend used with the CPU oracle (tests / bench cpu_baseline only)

class HostMemory:
    def __init__(self):
        self.arrays = {}

    def alloc(self, name, nbytes, init):
        arr = np.zeros(max(nbytes, 1), dtype=np.uint8)
        if init is not None:
            arr[:nbytes] = np.frombuffer(np.ascontiguousarray(init).tobytes(), dtype=np.uint8)[:nbytes]
        self.arrays[name] = arr
        return arr.ctypes.data


def run_oracle(scene, window=None):
    """Render `scene` with the CPU oracle. Returns (color bytes, depth bytes or None, DrawStats)."""
    lib = capi.load_oracle()
    mem = HostMemory()
    m = materialize(scene, mem.alloc)
    for img, att in ((scene.color, m.color_attachment), (scene.depth, m.depth_attachment)):
        if img is not None and img.clear is not None:
            cv, is_ds = clear_value(img)
            rc = lib.cpvk_oracle_clear(C.byref(att), C.byref(cv), is_ds)
            assert rc == 0, lib.cpvk_oracle_last_error()
    stats = capi.DrawStats()
    if window is None:
        rc = lib.cpvk_oracle_draw(C.byref(m.desc), C.byref(m.state), C.byref(stats))
    else:
        rc = lib.cpvk_oracle_draw_window(C.byref(m.desc), C.byref(m.state), *window, C.byref(stats))
    if rc != 0:
        raise RuntimeError("oracle draw failed: %s" % lib.cpvk_oracle_last_error().decode())
    color = mem.arrays["color"][:scene.color.nbytes].copy()
    depth = mem.arrays["depth"][:scene.depth.nbytes].copy() if scene.depth else None
    return color, depth, stats
