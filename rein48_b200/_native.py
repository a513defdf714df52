# -*- coding: utf-8 -*-
"""ctypes binding of libr48.so (the C ABI declared in include/r48.h).

There is deliberately NO fallback: if the library is missing or a call fails, this raises.
The product path never touches oracle/ or any CPU implementation of the environment.
"""
import ctypes as C
import os
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libr48.so")
SRC = os.path.join(_PKG, "csrc", "r48_kernels.cu")
DEPS = [SRC, os.path.join(_PKG, "csrc", "r48_device.cuh"), os.path.join(_ROOT, "include", "r48.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "186",
]

OK, ERR_NULL, ERR_ALIGN, ERR_ARG, ERR_CUDA, ERR_ACTION = 0, -1, -2, -3, -4, -5
STATS_WORDS = 4120
ROLLOUT_WORKSPACE_BYTES = 256

# every symbol include/r48.h declares (tests check the .so exports exactly these)
SYMBOLS = (
    "r48_version", "r48_build_id", "r48_last_error", "r48_init", "r48_debug_tables_host", "r48_reset", "r48_step",
    "r48_step_injected", "r48_step_injected_view", "r48_env_step", "r48_env_step_ring", "r48_spawn_injected", "r48_spawn", "r48_blank_counts",
    "r48_afterstates", "r48_rollout", "r48_rollout_policy", "r48_rollout_trajectories", "r48_episode_stats",
    "r48_episode_records", "r48_scores", "r48_decode_f32", "r48_decode_i32", "r48_encode_i32",
    "r48_ring_clear", "r48_ring_append", "r48_ring_sample", "r48_debug_copy22",
    "r48_step_host", "r48_afterstates_host", "r48_rollout_host", "r48_rollout_host_ex", "r48_shutdown",
)
VERSION = 201


class Ring(C.Structure):
    """struct r48_ring of include/r48.h (a host struct of device pointers); `slots` points at
    `capacity` 32-byte r48_transition records."""
    _fields_ = [("slots", C.c_void_p), ("cursor", C.c_void_p), ("capacity", C.c_uint64)]


class GameView(C.Structure):
    """struct r48_game_view of include/r48.h (r48_step_injected_view writes one per board)."""
    _fields_ = [("cells", C.c_int32 * 16), ("reward", C.c_int32), ("done", C.c_uint8), ("valid", C.c_uint8),
                ("blanks", C.c_uint8 * 4), ("reserved", C.c_uint8 * 2)]


ACTION_NONE = 255


class R48Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libr48 error %d: %s" % (code, message))
        self.code = code


def source_id():
    """16 hex digits over the contents of every file libr48.so is compiled from."""
    import hashlib
    h = hashlib.sha256()
    for path in DEPS:
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def built_id(path=LIB_PATH):
    """The source id a built library carries (r48_build_id), read from the file without loading
    it; None if the file is missing or carries none."""
    import re
    try:
        with open(path, "rb") as f:
            m = re.search(rb"R48_BUILD_ID=([0-9a-f]{16})", f.read())
    except OSError:
        return None
    return m.group(1).decode() if m else None


def build(force=False, verbose=False):
    """Compile libr48.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  A no-op when
    the library on disk was built from the current sources (compared by content, not by mtime:
    copies of the tree do not keep timestamps).  The new file is moved into place atomically, so
    several ranks calling this at once cannot load a half-written library."""
    want = source_id()
    if not force and built_id() == want:
        return LIB_PATH
    nvcc = os.environ.get("NVCC") or "nvcc"
    if not any(os.access(os.path.join(d, nvcc), os.X_OK) for d in os.environ.get("PATH", "").split(os.pathsep)):
        cand = "/usr/local/cuda/bin/nvcc"
        if os.path.exists(cand):
            nvcc = cand
    tmp = "%s.%d.tmp" % (LIB_PATH, os.getpid())
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        '-DR48_BUILD_ID="%s"' % want, "-o", tmp, SRC]
    try:
        subprocess.check_call(cmd)
        os.replace(tmp, LIB_PATH)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return LIB_PATH


_lib = None
_lock = threading.Lock()

_u64p = C.c_void_p   # all buffers cross the ABI as raw addresses (torch data_ptr / numpy ctypes.data)


def lib():
    """Load libr48.so.  It is (re)built first when it is absent or was built from other sources;
    a stale library that cannot be rebuilt is an error, never silently loaded."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("R48_LIBRARY")       # A/B tooling only: a build variant of the same sources
        if not path:
            path = LIB_PATH
            try:
                build()
            except (OSError, subprocess.CalledProcessError) as e:
                raise RuntimeError("libr48.so is missing or was built from other sources, and rebuilding it "
                                   "failed: %s" % e)
        L = C.CDLL(path)
        L.r48_version.restype = C.c_int
        if L.r48_version() != VERSION:
            raise RuntimeError("libr48.so reports ABI version %d, this package binds version %d: rebuild it "
                               "(python -c 'import rein48_b200; rein48_b200.build(force=True)')"
                               % (L.r48_version(), VERSION))
        vp, i64, u64, u32, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int
        L.r48_version.argtypes = []
        L.r48_version.restype = i32
        L.r48_last_error.argtypes = []
        L.r48_last_error.restype = C.c_char_p
        L.r48_build_id.argtypes = []
        L.r48_init.argtypes = [i32]
        L.r48_debug_tables_host.argtypes = [vp, vp, i32]
        L.r48_reset.argtypes = [vp, i64, u64, u64, vp]
        L.r48_step.argtypes = [vp, vp, vp, vp, vp, i64, u64, u64, u32, i32, vp, vp]
        L.r48_step_injected.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i32, vp, vp]
        L.r48_env_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, u64, u64, u64, i32, i32, vp, vp]
        L.r48_env_step_ring.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, u64, u64, u64, i32, i32, vp,
                                        C.POINTER(Ring), vp]
        L.r48_episode_records.argtypes = [vp, vp, vp, i64, vp]
        L.r48_ring_clear.argtypes = [C.POINTER(Ring), vp]
        L.r48_ring_append.argtypes = [C.POINTER(Ring), vp, vp, vp, vp, vp, i64, i32, vp]
        L.r48_ring_sample.argtypes = [C.POINTER(Ring), i64, u64, u64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
        L.r48_debug_copy22.argtypes = [vp, vp, vp, vp, vp, i64, vp]
        L.r48_rollout_host_ex.argtypes = [i64, u64, u64, i32, vp, vp, vp, vp, i32]
        L.r48_spawn_injected.argtypes = [vp, vp, vp, i64, vp]
        L.r48_step_injected_view.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, vp]
        L.r48_spawn.argtypes = [vp, i64, u64, u64, u32, vp]
        L.r48_blank_counts.argtypes = [vp, vp, i64, vp]
        L.r48_afterstates.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp]
        L.r48_rollout.argtypes = [i64, u64, u64, vp, vp, vp, vp, vp]
        L.r48_rollout_policy.argtypes = [i64, u64, u64, i32, vp, vp, vp, vp, vp]
        L.r48_rollout_trajectories.argtypes = [i64, u64, u64, i32, vp, vp, vp, vp, vp, vp]
        L.r48_episode_stats.argtypes = [vp, vp, i64, vp, vp]
        L.r48_scores.argtypes = [vp, vp, vp, i64, vp]
        L.r48_decode_f32.argtypes = [vp, vp, i64, i32, vp]
        L.r48_decode_i32.argtypes = [vp, vp, i64, vp]
        L.r48_encode_i32.argtypes = [vp, vp, i64, vp, vp]
        L.r48_step_host.argtypes = [vp, vp, vp, vp, vp, i64, u64, u64, u32, i32, i32]
        L.r48_afterstates_host.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32]
        L.r48_rollout_host.argtypes = [i64, u64, u64, vp, vp, vp, i32]
        L.r48_shutdown.argtypes = []
        for name in SYMBOLS:
            if name not in ("r48_last_error", "r48_build_id"):
                getattr(L, name).restype = i32
        L.r48_last_error.restype = C.c_char_p
        L.r48_build_id.restype = C.c_char_p
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().r48_last_error()
        raise R48Error(rc, msg.decode() if msg else "")
    return rc
