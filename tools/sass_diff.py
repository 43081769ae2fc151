#!/usr/bin/env python
"""Compare the SASS of two builds of libclipdlm.so kernel by kernel (cuobjdump -sass dumps or .so files).

Used when a change adds NEW template instantiations behind a default-off switch and there is no GPU at hand to re-measure: every
kernel of the default path must come out byte-identical (same instructions, same registers), so the measured numbers still hold.

  python tools/sass_diff.py before.{so,txt} after.{so,txt}      exit 1 if a kernel present in both differs
"""
import re
import subprocess
import sys


def dump(path: str) -> str:
    if path.endswith(".txt"):
        return open(path).read()
    return subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout


def functions(text: str) -> dict:
    out, name, body = {}, None, []
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name is not None:
                out[name] = body
            name, body = m.group(1), []
        elif name is not None:
            # instruction lines: "        /*0010*/   IMAD.MOV.U32 R1, RZ, RZ, c[0x0][0x28] ;   /* 0x... */" - keep the text, drop the encoding comment
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
            if m:
                body.append(re.sub(r"\s+", " ", m.group(1)).strip())
    if name is not None:
        out[name] = body
    return out


def canonical(body: list) -> list:
    """Rename registers / predicates by order of first appearance: two bodies that differ only by a permutation of register
    names (same instructions, same operands' data flow) come out equal."""
    names = {}

    def ren(m):
        tok = m.group(0)
        kind = re.match(r"[A-Z]+", tok).group(0)
        if tok not in names:
            names[tok] = f"{kind}#{sum(1 for k in names if re.match(r'[A-Z]+', k).group(0) == kind)}"
        return names[tok]

    return [re.sub(r"\b(?:UR|UP|R|P)\d+\b", ren, line) for line in body]


def main():
    a, b = functions(dump(sys.argv[1])), functions(dump(sys.argv[2]))
    same = [k for k in a if k in b and a[k] == b[k]]
    diff = [k for k in a if k in b and a[k] != b[k]]
    renamed = [k for k in diff if canonical(a[k]) == canonical(b[k])]
    diff = [k for k in diff if k not in renamed]
    # ptxas is not deterministic: two builds of the SAME source differ in a few kernels by the order / register names of a handful
    # of independent instructions (seen in the MMA-descriptor arithmetic of the GEMM issuer warp). Those have the same multiset of
    # instructions once register names are masked.
    mask = lambda body: sorted(re.sub(r"\b(?:UR|UP|R|P)\d+\b", "r", line) for line in body)
    permuted = [k for k in diff if mask(a[k]) == mask(b[k])]
    diff = [k for k in diff if k not in permuted]
    print(f"{len(a)} kernels before, {len(b)} after: {len(same)} identical, {len(renamed)} identical up to register renaming, "
          f"{len(permuted)} same instruction multiset (ptxas scheduling noise), {len(diff)} changed, {len(set(b) - set(a))} added, {len(set(a) - set(b))} removed")
    for k in sorted(renamed):
        print(f"  renamed  {k}  ({len(a[k])} instructions)")
    for k in sorted(permuted):
        print(f"  permuted {k}  ({len(a[k])} instructions)")
    for k in sorted(set(b) - set(a)):
        print(f"  added   {k}  ({len(b[k])} instructions)")
    for k in sorted(set(a) - set(b)):
        print(f"  removed {k}")
    for k in sorted(diff):
        print(f"  CHANGED {k}: {len(a[k])} -> {len(b[k])} instructions")
    sys.exit(1 if diff else 0)


if __name__ == "__main__":
    main()
