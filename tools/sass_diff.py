#!/usr/bin/env python3
"""Which kernels of libmgb200.so did a change touch?  Compares the SASS of two builds kernel by kernel.

    python tools/sass_diff.py /tmp/old.so minigraph_b200/libmgb200.so

A kernel listed as identical has the same instructions in the same order; for the others the instruction counts are
given (branch targets are part of the text, so an inserted instruction also changes the branches that jump across it).
Used to check that a switch that is off by default leaves the default kernels alone, and to see at a glance how much
a source change moved (needs cuobjdump; no GPU)."""
import difflib
import re
import subprocess
import sys


def kernels(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    out, cur = {}, None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m:
            out[cur].append(re.sub(r"\s+", " ", m.group(1)))
    return out


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    for k in sorted(set(a) | set(b)):
        if k not in a:
            print("%-45s only in the second (%d instr)" % (k, len(b[k])))
        elif k not in b:
            print("%-45s only in the first (%d instr)" % (k, len(a[k])))
        elif a[k] == b[k]:
            print("%-45s identical (%d instr)" % (k, len(a[k])))
        else:
            sm = difflib.SequenceMatcher(None, a[k], b[k], autojunk=False)
            same = sum(x.size for x in sm.get_matching_blocks())
            print("%-45s differs: %d -> %d instr, %d lines in common" % (k, len(a[k]), len(b[k]), same))


if __name__ == "__main__":
    main()
