#!/usr/bin/env python3
"""Where do the instructions of a kernel come from?  Static SASS instruction count per source line (needs -lineinfo).

    python tools/sass_lines.py minigraph_b200/libmgb200.so k_gwfa [top_n]

Extracts the cubin (cuobjdump -xelf), disassembles it with line information (nvdisasm -g -c) and attributes every
instruction of the kernel to the last `//## File "...", line N` marker in front of it.  Prints the per-file totals and the
heaviest lines: with a 32 KB L1.5 instruction cache (2048 instructions) what is inlined where decides whether the warps of
an SM can share their fetches.  No GPU needed."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    so, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        cubins = sorted((os.path.getsize(os.path.join(tmp, f)), f) for f in os.listdir(tmp) if f.endswith(".cubin"))
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubins[-1][1])], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    per_line, per_file, per_fn = collections.Counter(), collections.Counter(), collections.Counter()
    cur, inside, total, fn = ("?", 0), False, 0, "(kernel body)"
    for line in txt.split("\n"):
        if line.startswith(".text."):
            fn = "(kernel body)"
            inside = re.search(r"\.text\._Z\d+%s\d" % re.escape(kern), line) is not None or line.startswith(".text.%s:" % kern)
            continue
        if not inside:
            continue
        m = re.search(r"\.type\s+\$[^$]+\$(\S+),@function", line)  # a device function kept out of line (MG_NOINLINE)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            per_fn[fn] += 1
            total += 1
            if fn == "(kernel body)":
                per_line[cur] += 1
                per_file[cur[0]] += 1
    print("%s: %d instructions (%.0f KB)" % (kern, total, total * 16 / 1024))
    for f, n in per_fn.most_common():
        print("  %6d  %s" % (n, f))
    total = per_fn["(kernel body)"]
    print("kernel body by file:")
    for f, n in per_file.most_common():
        print("  %-28s %6d  %4.1f%%" % (f, n, 100.0 * n / max(1, total)))
    print("heaviest lines:")
    for (f, l), n in per_line.most_common(top):
        print("  %-28s %5d  %5d" % (f, l, n))


if __name__ == "__main__":
    main()
