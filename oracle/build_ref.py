#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Build the UNMODIFIED reference into oracle/_ref/ (test / baseline infrastructure only).

The reference's hot path is Python (game/GameClient.py, control/rand.py, main.py:play), so its
"build" is byte-compilation: every file is compiled from where it lies under $R48_REFERENCE
(default /root/reference) into a marshalled code object under oracle/_ref/ -- a built artefact
like a .so (git-ignored, travels to the GPU box with the snapshot); no reference source text
enters the repository.  The files are named *.r48c rather than *.pyc because tree snapshots
commonly drop *.pyc as interpreter litter; oracle/refarm.py loads them.

    python oracle/build_ref.py            # or: make -C oracle _ref

bench.py (`--impl reference`, `cpu_baseline`) imports the result through oracle/refarm.py and
reports kind "reference"; when oracle/_ref/ is absent it falls back to oracle/pyport.py (kind
"port") and says so.
"""
import marshal
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("main.py", "game/GameClient.py", "control/rand.py", "control/hand.py")


def build(reference=None, quiet=False):
    reference = reference or os.environ.get("R48_REFERENCE", "/root/reference")
    if not all(os.path.exists(os.path.join(reference, f)) for f in FILES):
        if not quiet:
            print("build_ref: no reference checkout at %s; oracle/_ref not (re)built" % reference)
        return None
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    import warnings
    for f in FILES:
        dst = os.path.join(DEST, f[:-3] + ".r48c")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(os.path.join(reference, f), "rb") as fh:
            text = fh.read()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)        # main.py's ASCII-art banner has stray backslashes
            code = compile(text, "<reference>/" + f, "exec", dont_inherit=True)
        with open(dst, "wb") as fh:
            fh.write(b"R48C%d.%d\n" % sys.version_info[:2])        # marshal is specific to the minor version
            marshal.dump(code, fh)
    with open(os.path.join(DEST, "BUILD_INFO"), "w") as fh:
        fh.write("byte-compiled from %s by oracle/build_ref.py with python %s\nfiles: %s\n"
                 % (reference, sys.version.split()[0], " ".join(FILES)))
    if not quiet:
        print("build_ref: %d files -> %s" % (len(FILES), DEST))
    return DEST


if __name__ == "__main__":
    sys.exit(0 if build(sys.argv[1] if len(sys.argv) > 1 else None) or True else 1)
