"""Compile one scene's pipeline on the CPU (no GPU needed) and print SASS statistics of its stage kernels:
instruction count, registers/stack from the cubin, and an opcode histogram (MUFU.RCP + FCHK pairs = IEEE divides).

    python tools/sass_stats.py c3|c1|c2|c4 [--dump out.sass]
"""
import collections
import ctypes as C
import re
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, ".")
from cpvulkan_b200 import capi, scenes  # noqa: E402

SCENES = {
    "c1": lambda: scenes.draw_cube(64, 64),
    "c2": lambda: scenes.draw_textured_cube(64, 64, scenes.LINEAR),
    "c3": lambda: scenes.mesh_indexed(64, 64, 4, 4),
    "c4": lambda: scenes.overdraw_quads(64, 64, 2, 16),
}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    lib = capi.load_cuda()
    m = scenes.materialize(SCENES[which](), scenes.HostMemory().alloc)
    p = C.c_void_p()
    t0 = time.time()
    rc = lib.cpvk_cuda_pipeline_compile_only(C.byref(m.desc), C.byref(p))
    assert rc == 0, lib.cpvk_cuda_last_error()
    print("compile %.1f s" % (time.time() - t0))
    n = C.c_size_t()
    cubin = lib.cpvk_cuda_pipeline_cubin(p, C.byref(n))
    with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
        f.write(C.string_at(cubin, n.value)); f.flush()
        res = subprocess.run(["cuobjdump", "-res-usage", f.name], capture_output=True, text=True).stdout
        for line in res.splitlines():
            if "Function" in line or "REG" in line:
                print(line.strip())
        sass = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True).stdout
    if "--dump" in sys.argv:
        open(sys.argv[sys.argv.index("--dump") + 1], "w").write(sass)
    cur, hist = None, {}
    for line in sass.splitlines():
        mm = re.search(r"Function : (\S+)", line)
        if mm:
            cur = mm.group(1); hist[cur] = collections.Counter(); continue
        mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if mm and cur:
            hist[cur][mm.group(1).split(".")[0]] += 1
    for k, h in hist.items():
        print(k, sum(h.values()), "instructions;", ", ".join("%s %d" % kv for kv in h.most_common(14)))


if __name__ == "__main__":
    main()
