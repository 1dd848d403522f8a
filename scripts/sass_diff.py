#!/usr/bin/env python
"""Which kernels of an object file changed?  Compares the instruction streams (cuobjdump -sass,
encodings and addresses stripped) of two builds kernel by kernel, keyed on demangled names.

    cuobjdump -sass sigma_b200/lib/obj/kernels_spmv.o > /tmp/before.sass      # before the edit
    ... edit, make ...
    python scripts/sass_diff.py /tmp/before.sass sigma_b200/lib/obj/kernels_spmv.o

Used in round 1 to show, without a GPU, that the opt-in paths (new template parameters, new
kernels) leave every product kernel's code byte for byte as it was measured."""
import hashlib
import re
import subprocess
import sys


def load(path):
    text = open(path).read() if path.endswith(".sass") else subprocess.run(
        ["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, cur, buf = {}, None, []
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                out[cur] = buf
            cur, buf = m.group(1), []
            continue
        mm = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if mm and cur:
            buf.append(re.sub(r"\s+", " ", mm.group(1)))
    if cur:
        out[cur] = buf
    names = list(out)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().split("\n") if names else []
    res = {}
    for n, d in zip(names, dem):
        d = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", d)
        res[d] = (hashlib.md5("\n".join(out[n]).encode()).hexdigest(), len(out[n]))
    return res


def main():
    a, b = load(sys.argv[1]), load(sys.argv[2])
    changed = [k for k in a if k in b and a[k] != b[k]]
    gone = [k for k in a if k not in b]
    new = [k for k in b if k not in a]
    print(f"{len(a)} kernels before, {len(b)} after: {len(a) - len(changed) - len(gone)} identical, "
          f"{len(changed)} changed, {len(gone)} gone (or renamed), {len(new)} new")
    for k in changed:
        print("  changed:", k, a[k][1], "->", b[k][1], "instructions")
    for k in gone:
        print("  gone   :", k)
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
