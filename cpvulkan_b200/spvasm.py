"""Minimal SPIR-V assembler (text -> words).

The reference's samples compile GLSL at run time with glslang (Samples/utils/util.cpp:498-550), which is
absent from this image, and no spirv-as exists either. This tool lets the harness and
the tests keep their shaders as readable SPIR-V assembly next to the GLSL they restate; the assembled
words are committed under tests/golden/ so nothing is assembled at run time on the GPU box.

Syntax: one instruction per line, `%res = OpName operands...` or `OpName operands...`; `;` starts a comment.
Operands: %ids, integers (dec/hex), floats (only where a float literal is expected, i.e. OpConstant /
OpSpecConstant of a float type), "strings", and enumerant names.
"""
import re
import struct

MAGIC = 0x07230203
VERSION_1_0 = 0x00010000

# name -> (opcode, has_result_type, has_result_id)
OPS = {
    "Nop": (0, 0, 0), "Undef": (1, 1, 1), "Source": (3, 0, 0), "SourceExtension": (4, 0, 0), "Name": (5, 0, 0),
    "MemberName": (6, 0, 0), "ExtInstImport": (11, 0, 1), "ExtInst": (12, 1, 1), "MemoryModel": (14, 0, 0),
    "EntryPoint": (15, 0, 0), "ExecutionMode": (16, 0, 0), "Capability": (17, 0, 0), "TypeVoid": (19, 0, 1),
    "TypeBool": (20, 0, 1), "TypeInt": (21, 0, 1), "TypeFloat": (22, 0, 1), "TypeVector": (23, 0, 1),
    "TypeMatrix": (24, 0, 1), "TypeImage": (25, 0, 1), "TypeSampler": (26, 0, 1), "TypeSampledImage": (27, 0, 1),
    "TypeArray": (28, 0, 1), "TypeRuntimeArray": (29, 0, 1), "TypeStruct": (30, 0, 1), "TypePointer": (32, 0, 1),
    "TypeFunction": (33, 0, 1), "ConstantTrue": (41, 1, 1), "ConstantFalse": (42, 1, 1), "Constant": (43, 1, 1),
    "ConstantComposite": (44, 1, 1), "ConstantNull": (46, 1, 1), "SpecConstantTrue": (48, 1, 1),
    "SpecConstantFalse": (49, 1, 1), "SpecConstant": (50, 1, 1), "SpecConstantComposite": (51, 1, 1),
    "Function": (54, 1, 1), "FunctionParameter": (55, 1, 1), "FunctionEnd": (56, 0, 0), "FunctionCall": (57, 1, 1),
    "Variable": (59, 1, 1), "Load": (61, 1, 1), "Store": (62, 0, 0), "AccessChain": (65, 1, 1),
    "InBoundsAccessChain": (66, 1, 1), "Decorate": (71, 0, 0), "MemberDecorate": (72, 0, 0),
    "VectorExtractDynamic": (77, 1, 1), "VectorInsertDynamic": (78, 1, 1), "VectorShuffle": (79, 1, 1),
    "CompositeConstruct": (80, 1, 1), "CompositeExtract": (81, 1, 1), "CompositeInsert": (82, 1, 1),
    "CopyObject": (83, 1, 1), "Transpose": (84, 1, 1), "SampledImage": (86, 1, 1),
    "ImageSampleImplicitLod": (87, 1, 1), "ImageSampleExplicitLod": (88, 1, 1), "ImageFetch": (95, 1, 1),
    "ImageRead": (98, 1, 1),
    "Image": (100, 1, 1), "ImageQuerySizeLod": (103, 1, 1), "ImageQuerySize": (104, 1, 1),
    "ConvertFToU": (109, 1, 1), "ConvertFToS": (110, 1, 1), "ConvertSToF": (111, 1, 1), "ConvertUToF": (112, 1, 1),
    "UConvert": (113, 1, 1), "SConvert": (114, 1, 1), "FConvert": (115, 1, 1), "Bitcast": (124, 1, 1),
    "SNegate": (126, 1, 1), "FNegate": (127, 1, 1), "IAdd": (128, 1, 1), "FAdd": (129, 1, 1), "ISub": (130, 1, 1),
    "FSub": (131, 1, 1), "IMul": (132, 1, 1), "FMul": (133, 1, 1), "UDiv": (134, 1, 1), "SDiv": (135, 1, 1),
    "FDiv": (136, 1, 1), "UMod": (137, 1, 1), "SRem": (138, 1, 1), "SMod": (139, 1, 1), "FRem": (140, 1, 1),
    "FMod": (141, 1, 1), "VectorTimesScalar": (142, 1, 1), "MatrixTimesScalar": (143, 1, 1),
    "VectorTimesMatrix": (144, 1, 1), "MatrixTimesVector": (145, 1, 1), "MatrixTimesMatrix": (146, 1, 1),
    "OuterProduct": (147, 1, 1), "Dot": (148, 1, 1), "Any": (154, 1, 1), "All": (155, 1, 1), "IsNan": (156, 1, 1),
    "IsInf": (157, 1, 1), "LogicalEqual": (164, 1, 1), "LogicalNotEqual": (165, 1, 1), "LogicalOr": (166, 1, 1),
    "LogicalAnd": (167, 1, 1), "LogicalNot": (168, 1, 1), "Select": (169, 1, 1), "IEqual": (170, 1, 1),
    "INotEqual": (171, 1, 1), "UGreaterThan": (172, 1, 1), "SGreaterThan": (173, 1, 1),
    "UGreaterThanEqual": (174, 1, 1), "SGreaterThanEqual": (175, 1, 1), "ULessThan": (176, 1, 1),
    "SLessThan": (177, 1, 1), "ULessThanEqual": (178, 1, 1), "SLessThanEqual": (179, 1, 1),
    "FOrdEqual": (180, 1, 1), "FUnordEqual": (181, 1, 1), "FOrdNotEqual": (182, 1, 1), "FUnordNotEqual": (183, 1, 1),
    "FOrdLessThan": (184, 1, 1), "FUnordLessThan": (185, 1, 1), "FOrdGreaterThan": (186, 1, 1),
    "FUnordGreaterThan": (187, 1, 1), "FOrdLessThanEqual": (188, 1, 1), "FUnordLessThanEqual": (189, 1, 1),
    "FOrdGreaterThanEqual": (190, 1, 1), "FUnordGreaterThanEqual": (191, 1, 1), "ShiftRightLogical": (194, 1, 1),
    "ShiftRightArithmetic": (195, 1, 1), "ShiftLeftLogical": (196, 1, 1), "BitwiseOr": (197, 1, 1),
    "BitwiseXor": (198, 1, 1), "BitwiseAnd": (199, 1, 1), "Not": (200, 1, 1), "Phi": (245, 1, 1),
    "LoopMerge": (246, 0, 0), "SelectionMerge": (247, 0, 0), "Label": (248, 0, 1), "Branch": (249, 0, 0),
    "BranchConditional": (250, 0, 0), "Switch": (251, 0, 0), "Kill": (252, 0, 0), "Return": (253, 0, 0),
    "ReturnValue": (254, 0, 0), "Unreachable": (255, 0, 0),
}

STORAGE_CLASS = {"UniformConstant": 0, "Input": 1, "Uniform": 2, "Output": 3, "Workgroup": 4, "CrossWorkgroup": 5,
                 "Private": 6, "Function": 7, "Generic": 8, "PushConstant": 9, "AtomicCounter": 10, "Image": 11,
                 "StorageBuffer": 12}
DECORATION = {"RelaxedPrecision": 0, "SpecId": 1, "Block": 2, "BufferBlock": 3, "RowMajor": 4, "ColMajor": 5,
              "ArrayStride": 6, "MatrixStride": 7, "GLSLShared": 8, "GLSLPacked": 9, "BuiltIn": 11,
              "NoPerspective": 13, "Flat": 14, "Centroid": 16, "Invariant": 18, "NonWritable": 24,
              "NonReadable": 25, "Location": 30, "Component": 31, "Index": 32, "Binding": 33, "DescriptorSet": 34,
              "Offset": 35, "InputAttachmentIndex": 43}
BUILTIN = {"Position": 0, "PointSize": 1, "ClipDistance": 3, "CullDistance": 4, "VertexId": 5, "InstanceId": 6,
           "PrimitiveId": 7, "FragCoord": 15, "PointCoord": 16, "FrontFacing": 17, "FragDepth": 22,
           "VertexIndex": 42, "InstanceIndex": 43}
MISC = {
    # ExecutionModel
    "Vertex": 0, "Fragment": 4, "GLCompute": 5,
    # AddressingModel / MemoryModel
    "Logical": 0, "Simple": 0, "GLSL450": 1,
    # ExecutionMode
    "OriginUpperLeft": 7, "OriginLowerLeft": 8, "EarlyFragmentTests": 9, "DepthReplacing": 12,
    # Capability
    "Matrix": 0, "Shader": 1, "InputAttachment": 40, "SampledBuffer": 46, "ImageBuffer": 47,
    # Dim
    "1D": 0, "2D": 1, "3D": 2, "Cube": 3, "Rect": 4, "Buffer": 5, "SubpassData": 6,
    # ImageFormat
    "Unknown": 0, "Rgba32f": 1, "Rgba8": 4, "R32f": 3,
    # ImageOperands
    "Bias": 1, "Lod": 2, "Grad": 4, "ConstOffset": 8,
    # Function / selection / loop control
    "None": 0, "Inline": 1, "DontInline": 2, "Flatten": 1, "DontFlatten": 2, "Unroll": 1, "DontUnroll": 2,
    # Source language
    "GLSL": 2, "ESSL": 1,
}
GLSL_STD_450 = {"Round": 1, "RoundEven": 2, "Trunc": 3, "FAbs": 4, "SAbs": 5, "FSign": 6, "SSign": 7, "Floor": 8,
                "Ceil": 9, "Fract": 10, "Radians": 11, "Degrees": 12, "Sin": 13, "Cos": 14, "Tan": 15, "Pow": 26,
                "Exp": 27, "Log": 28, "Exp2": 29, "Log2": 30, "Sqrt": 31, "InverseSqrt": 32, "FMin": 37, "UMin": 38,
                "SMin": 39, "FMax": 40, "UMax": 41, "SMax": 42, "FClamp": 43, "UClamp": 44, "SClamp": 45,
                "FMix": 46, "Step": 48, "SmoothStep": 49, "Fma": 50, "Length": 66, "Distance": 67, "Cross": 68,
                "Normalize": 69, "FaceForward": 70, "Reflect": 71, "Refract": 72, "FindILsb": 73, "FindSMsb": 74,
                "FindUMsb": 75, "NMin": 79, "NMax": 80, "NClamp": 81}

_TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|[^\s]+')


def _string_words(s):
    raw = s.encode("utf-8") + b"\0"
    raw += b"\0" * ((4 - len(raw) % 4) % 4)
    return list(struct.unpack("<%dI" % (len(raw) // 4), raw))


class AssemblyError(Exception):
    pass


def assemble(text):
    """Assemble SPIR-V text; returns a list of 32-bit words (header included)."""
    ids = {}

    def get_id(tok):
        name = tok[1:]
        if name not in ids:
            ids[name] = len(ids) + 1
        return ids[name]

    # First pass: tokenise and learn which ids are float / int types (for OpConstant literal encoding).
    lines = []
    for lineno, raw in enumerate(text.splitlines(), 1):
        line = raw.split(";", 1)[0].strip()
        if not line:
            continue
        toks = _TOKEN.findall(line)
        res = None
        if len(toks) >= 3 and toks[1] == "=":
            res = toks[0]
            toks = toks[2:]
        if not toks[0].startswith("Op") or toks[0][2:] not in OPS:
            raise AssemblyError("line %d: unknown instruction %r" % (lineno, toks[0]))
        lines.append((lineno, res, toks[0][2:], toks[1:]))

    float_types = set()
    for _, res, op, args in lines:
        if op == "TypeFloat":
            float_types.add(res)

    words = []
    for lineno, res, op, args in lines:
        opcode, has_type, has_res = OPS[op]
        if bool(has_res) != (res is not None):
            raise AssemblyError("line %d: Op%s result id mismatch" % (lineno, op))
        ops = []
        argi = 0
        if has_type:
            ops.append(get_id(args[0]))
            argi = 1
        if has_res:
            ops.append(get_id(res))
        rest = args[argi:]
        float_literal = op in ("Constant", "SpecConstant") and args[0] in float_types
        for k, tok in enumerate(rest):
            if tok.startswith("%"):
                ops.append(get_id(tok))
            elif tok.startswith('"'):
                ops.extend(_string_words(bytes(tok[1:-1], "utf-8").decode("unicode_escape")))
            elif re.fullmatch(r"-?(0x[0-9a-fA-F]+|\d+)", tok) and not float_literal:
                ops.append(int(tok, 0) & 0xFFFFFFFF)
            elif float_literal and re.fullmatch(r"[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?|inf|nan)", tok):
                ops.append(struct.unpack("<I", struct.pack("<f", float(tok)))[0])
            elif float_literal and tok.startswith("bits:"):
                ops.append(int(tok[5:], 0) & 0xFFFFFFFF)
            else:
                table = None
                if op in ("Variable", "TypePointer") and k == 0:
                    table = STORAGE_CLASS
                elif op == "Decorate" and k == 1:
                    table = DECORATION
                elif op == "MemberDecorate" and k == 2:
                    table = DECORATION
                elif op in ("Decorate", "MemberDecorate") and rest[k - 1] == "BuiltIn":
                    table = BUILTIN
                elif op == "ExtInst" and k == 1:
                    table = GLSL_STD_450
                else:
                    table = MISC
                if "|" in tok:
                    val = 0
                    for part in tok.split("|"):
                        val |= table[part]
                    ops.append(val)
                elif tok in table:
                    ops.append(table[tok])
                else:
                    raise AssemblyError("line %d: cannot encode operand %r of Op%s" % (lineno, tok, op))
        words.append(((len(ops) + 1) << 16) | opcode)
        words.extend(ops)

    header = [MAGIC, VERSION_1_0, 0x00B20001, len(ids) + 1, 0]
    return header + words


def assemble_bytes(text):
    w = assemble(text)
    return struct.pack("<%dI" % len(w), *w)
