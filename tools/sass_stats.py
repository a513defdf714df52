#!/usr/bin/env python
"""Per-kernel SASS instruction mix of a built library (no GPU needed).

    python tools/sass_stats.py rein48_b200/libr48.so [kernel-substring] [--dump]
"""
import collections
import re
import subprocess
import sys


def main():
    so = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    dump = "--dump" in sys.argv
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    name = None
    kernels = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and name:
            kernels[name].append((m.group(1), m.group(2)))
    for name, ins in kernels.items():
        if pat not in name:
            continue
        mix = collections.Counter()
        for _, text in ins:
            t = text.split()
            op = t[1] if t[0].startswith("@") else t[0]
            mix[op.split(".")[0]] += 1
        print("== %s: %d instructions" % (name, len(ins)))
        print("   " + "  ".join("%s:%d" % kv for kv in mix.most_common()))
        if dump:
            for addr, text in ins:
                print("   %s  %s" % (addr, text))


if __name__ == "__main__":
    main()
