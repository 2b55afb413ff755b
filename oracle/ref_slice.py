"""ref_slice.py — build-time recipe that lifts LINE RANGES of a reference source file into a scratch include file.

Why: CPVulkan/CommandBuffer.Draw.cpp cannot be compiled as a whole translation unit here (it pulls in the whole ICD:
LLVM-8, the Vulkan SDK, SPIRV classes, xcb ...), but the functions that decide every covered pixel and every
interpolated value are self-contained `static` functions and templates. oracle/Makefile calls this script to cut exactly
those functions out of the file WHERE IT LIES under /root/reference, into a temporary directory outside the repository;
oracle/ref_draw_check.cpp then #includes the result, so the machine code of oracle/_ref/draw_check IS the reference's
own code for them. Nothing of the reference is written into the repository tree (only the linked binary lands in the
git-ignored oracle/_ref/).

Every range carries an anchor: a string its first line must contain, and one its last line must contain. A reference
checkout whose line numbers moved fails the build loudly instead of silently slicing something else.
`#line` directives keep compiler diagnostics pointing at the reference file.

usage: ref_slice.py <source> <output> <first>-<last>@<anchor on first line>@<anchor on last line> ...
TEST INFRASTRUCTURE ONLY."""
import sys


def main(argv):
    if len(argv) < 4:
        sys.stderr.write(__doc__)
        return 2
    source, output, specs = argv[1], argv[2], argv[3:]
    with open(source, encoding="utf-8", errors="replace") as f:
        lines = f.read().split("\n")
    out = ["// generated at build time by oracle/ref_slice.py from %s — scratch file, never committed\n" % source]
    for spec in specs:
        rng, first_anchor, last_anchor = spec.split("@", 2)
        first, last = (int(v) for v in rng.split("-"))
        body = lines[first - 1:last]
        if not body or first_anchor not in body[0] or last_anchor not in body[-1]:
            sys.stderr.write("ref_slice: %s:%d-%d does not start with %r / end with %r — the reference moved; refusing to slice\n"
                             % (source, first, last, first_anchor, last_anchor))
            return 1
        out.append('#line %d "%s"\n' % (first, source))
        out.append("\n".join(body) + "\n")
    with open(output, "w", encoding="utf-8") as f:
        f.write("".join(out))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
