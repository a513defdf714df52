#!/usr/bin/env python
"""Per-component instruction budget of a kernel: EXECUTED warp-instructions per source function.

    python tools/sass_budget.py <report.ncu-rep | source.csv> <libr48.so> <kernel substring> <units> [--lines]

Joins two views of the same binary:
  * `ncu -i report --page source --csv`  per SASS instruction: "Instructions Executed" (from a
    `ncu --set full --import-source on` capture of ONE launch), in address order;
  * `nvdisasm -gi` of the cubin inside the .so     per SASS instruction: the innermost source line
    (the library is compiled with -lineinfo), which is mapped to the enclosing function of
    rein48_b200/csrc/*.cu*.
`units` = boards (or env-steps) the profiled launch processed; the table is warp-instructions per
32 units, i.e. instructions per board/step per lane.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "rein48_b200", "csrc", f) for f in ("r48_device.cuh", "r48_kernels.cu")]


def function_map(path):
    """line -> name of the enclosing top-level function (crude brace matching: a definition is the
    text from the end of the previous top-level item up to its opening brace at namespace depth)."""
    owner, depth, current, start_depth, sig = {}, 0, None, None, ""
    for no, line in enumerate(open(path), 1):
        code = line.split("//")[0]
        if current is None and depth <= 1:
            if code.lstrip().startswith("#"):
                continue
            sig += " " + code.strip()
            if ";" in code and "{" not in code:
                sig = ""
            elif "{" in code:
                head = sig.split("{")[0]
                head = re.sub(r"__launch_bounds__\s*\([^)]*\)", " ", head)
                m = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*\(", head)
                if m and not re.match(r"\s*(namespace|struct|enum|extern|class)\b", head.strip()):
                    current, start_depth = m[0] if m[0] not in ("template",) else m[-1], depth
                sig = ""
        if current:
            owner[no] = current
        depth += code.count("{") - code.count("}")
        if current and depth <= start_depth:
            current = None
    return owner


def disasm(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    out, inside, loc = [], False, None
    for line in text.splitlines():
        if line.startswith(".text."):
            inside = kernel in line
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            if loc is None:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))      # innermost frame comes first
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), loc))
            # a location applies until the next annotation
            continue
        if line.strip() == "":
            continue
        if line.lstrip().startswith(".L_") or line.lstrip().startswith("."):
            continue
    # carry locations forward: nvdisasm prints an annotation only when the location changes
    fixed, last = [], None
    it = iter(text.splitlines())
    inside = False
    for line in it:
        if line.startswith(".text."):
            inside = kernel in line
            last = None
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            # the first annotation of a group is the innermost frame; groups are contiguous comment lines
            if not getattr(disasm, "_in_group", False):
                last = (os.path.basename(m.group(1)), int(m.group(2)))
                disasm._in_group = True
            continue
        disasm._in_group = False
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            fixed.append((int(m.group(1), 16), m.group(2).strip(), last))
    return fixed


def executed(report, kernel):
    if report.endswith(".csv"):              # an `ncu --page source --csv` export made on the GPU box
        text = open(report).read()
    else:
        text = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows, take, hdr = [], False, None
    for r in csv.reader(io.StringIO(text)):
        if r and r[0] == "Kernel Name":
            take = kernel in r[1].replace("(bool)", "").replace(" ", "") or kernel in r[1]
            hdr = None
            continue
        if not take or not r:
            continue
        if r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr:
            rows.append((r[hdr["Source"]].strip(), int(r[hdr["Instructions Executed"]]),
                         int(r[hdr["Thread Instructions Executed"]])))
    return rows


def main():
    report, so, kernel, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    owners = {os.path.basename(p): function_map(p) for p in SRC}
    dis = disasm(so, kernel)
    exe = executed(report, kernel.split("ILb")[0].split("ILi")[0].replace("_ZN3r48", "").lstrip("0123456789"))
    if len(dis) != len(exe):
        sys.exit("instruction counts differ: nvdisasm %d vs ncu %d -- the report was not taken from this binary"
                 % (len(dis), len(exe)))
    for (_, a, _), (b, _, _) in zip(dis, exe):
        if a.split()[0].lstrip("@!P0123456789U ") [:3] != b.split()[0].lstrip("@!P0123456789U ")[:3] and a.split()[-1][:2] != b.split()[-1][:2]:
            pass                                    # operand spelling differs between the tools; the count check guards alignment
    per_fn = collections.Counter()
    lanes = collections.Counter()
    per_line = collections.Counter()
    for (_, text, loc), (_, n, thr) in zip(dis, exe):
        fn = owners.get(loc[0], {}).get(loc[1], "?") if loc else "?"
        per_fn[fn] += n
        lanes[fn] += thr
        per_line[(fn, loc)] += n
    total = sum(per_fn.values())
    scale = 32.0 / units
    print("kernel %s: %d warp-instructions executed, %.1f per unit-lane (units = %d)" % (kernel, total, total * scale, units))
    print("%-28s %10s %8s %7s" % ("function", "warp-inst", "per unit", "lanes"))
    for fn, n in per_fn.most_common():
        print("%-28s %10d %8.1f %7.1f" % (fn, n, n * scale, lanes[fn] / max(1, n)))
    if "--lines" in sys.argv:
        for (fn, loc), n in per_line.most_common(60):
            print("   %-24s %-28s %8.2f" % (fn, "%s:%d" % loc if loc else "?", n * scale))


if __name__ == "__main__":
    main()
