#!/usr/bin/env python
"""A/B timing of build variants of libr48.so (no package import: raw ctypes on each file).

    python tools/ab_kernels.py build            # here: compile tools/ab/libr48_<variant>.so (nvcc cross-compiles)
    python tools/ab_kernels.py run [variants]   # on the GPU box: time rollout / step / env / afterstates per variant

Variants are sets of -D switches (VARIANTS below).  Every variant is first checked against the
default build (same final boards / lengths / step outputs) so that a fast wrong kernel cannot win.
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "rein48_b200", "csrc", "r48_kernels.cu")
OUT = os.path.join(ROOT, "tools", "ab")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-shared", "-diag-suppress", "186"]
VARIANTS = {
    "base": [],
    "unroll": ["-DR48_STEP_UNROLL=1", "-DR48_STEP_GATE_EARLY=1"],          # step_kernel: two trips per loop body
    "unroll_pf0": ["-DR48_STEP_UNROLL=1", "-DR48_STEP_GATE_EARLY=1", "-DR48_STEP_PREFETCH=0"],
    "unroll_pf2": ["-DR48_STEP_UNROLL=1", "-DR48_STEP_GATE_EARLY=1", "-DR48_STEP_PREFETCH=2"],
    "gate_early": ["-DR48_STEP_GATE_EARLY=1"],
    "ring64": ["-DR48_RING_L2_HINT=64"],
    "ringnc": ["-DR48_RING_L2_HINT=1"],
    "guarded": ["-DR48_UNGUARDED_TICKS=0u"],                             # rollout: always the range-guarded lookup body (must give the same outputs)
    "nopred": ["-DR48_PRED_TRANSPOSE=0"],                                # transposes behind a branch instead of predicated
    "glb": ["-DR48_STEP_TABLE_GLOBAL=1"],                                # step_kernel: LR table from global memory / L1, no staging
    "nopipe": ["-DR48_AFTER_PIPE=0"],
    "noapf": ["-DR48_AFTER_PREFETCH=0"],
    "fma": ["-DR48_FMA_INDEX=1"],
    "swz": ["-DR48_SWIZZLE=1"],
    "fmaswz": ["-DR48_FMA_INDEX=1", "-DR48_SWIZZLE=1"],
}


def build(only=None):
    os.makedirs(OUT, exist_ok=True)
    for name, defs in VARIANTS.items():
        if only and name not in only:
            continue
        so = os.path.join(OUT, "libr48_%s.so" % name)
        subprocess.check_call(["nvcc"] + FLAGS + defs + ["-o", so, SRC])
        print("built", so)


def load(name):
    L = C.CDLL(os.path.join(OUT, "libr48_%s.so" % name))
    vp, i64, u64, u32, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int
    L.r48_last_error.restype = C.c_char_p
    L.r48_init.argtypes = [i32]
    L.r48_reset.argtypes = [vp, i64, u64, u64, vp]
    L.r48_step.argtypes = [vp, vp, vp, vp, vp, i64, u64, u64, u32, i32, vp, vp]
    L.r48_rollout.argtypes = [i64, u64, u64, vp, vp, vp, vp, vp]
    L.r48_rollout_policy.argtypes = [i64, u64, u64, i32, vp, vp, vp, vp, vp]
    L.r48_afterstates.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp]
    L.r48_env_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, u64, u64, u64, i32, i32, vp, vp]
    L.r48_debug_copy22.argtypes = [vp, vp, vp, vp, vp, i64, vp]
    return L


def main():
    if sys.argv[1] == "build":
        return build(sys.argv[2:])
    import torch
    names = sys.argv[2:] or list(VARIANTS)
    torch.cuda.set_device(0)
    dev = "cuda"
    st = lambda: torch.cuda.current_stream().cuda_stream

    def ck(L, rc):
        if rc:
            raise RuntimeError(L.r48_last_error().decode())

    def graph_time(fn, launches, replays=10):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(launches):
                fn(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(launches):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (replays * launches) * 1e3      # us per launch

    n_roll = 1 << 24
    fb = torch.empty(n_roll, dtype=torch.int64, device=dev)
    ln = torch.empty(n_roll, dtype=torch.int32, device=dev)
    ws = torch.zeros(32, dtype=torch.int64, device=dev)
    n, sets = 1 << 20, 8
    nb = n * sets
    boards = torch.empty(nb, dtype=torch.int64, device=dev)
    out = torch.empty(nb, dtype=torch.int64, device=dev)
    rw = torch.empty(nb, dtype=torch.int32, device=dev)
    dn = torch.empty(nb, dtype=torch.uint8, device=dev)
    acts = torch.randint(0, 4, (nb,), device=dev, dtype=torch.uint8, generator=torch.Generator(device=dev).manual_seed(1))
    ref = {}
    results = {}
    for name in names:
        L = load(name)
        ck(L, L.r48_init(0))
        # mid-game boards: reset + 64 random steps (identical for every correct variant)
        ck(L, L.r48_reset(boards.data_ptr(), nb, 2048, 0, st()))
        for t in range(64):
            a = torch.randint(0, 4, (nb,), device=dev, dtype=torch.uint8, generator=torch.Generator(device=dev).manual_seed(100 + t))
            ck(L, L.r48_step(boards.data_ptr(), a.data_ptr(), boards.data_ptr(), None, None, nb, 2048, 0, t, 0, None, st()))
        ck(L, L.r48_step(boards.data_ptr(), acts.data_ptr(), out.data_ptr(), rw.data_ptr(), dn.data_ptr(), nb, 2048, 0, 64, 0, None, st()))
        ck(L, L.r48_rollout(1 << 20, 7, 5, fb.data_ptr(), ln.data_ptr(), None, ws.data_ptr(), st()))
        torch.cuda.synchronize()
        sig = (int(out.sum().item()), int(dn.sum().item()), int(fb[:1 << 20].sum().item()), int(ln[:1 << 20].sum().item()))
        if not ref:
            ref["sig"] = sig
        ok = sig == ref["sig"]
        r = {"matches_first_variant": ok}
        # fused rollout, 2^24 episodes
        for i in range(2):
            ck(L, L.r48_rollout(n_roll, 2048 + i, 0, fb.data_ptr(), ln.data_ptr(), None, ws.data_ptr(), st()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 3
        for i in range(reps):
            ck(L, L.r48_rollout(n_roll, 2048 + i, 0, fb.data_ptr(), ln.data_ptr(), None, ws.data_ptr(), st()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        steps = int(ln.to(torch.int64).sum().item())
        r["rollout_ms_2^24"] = ms
        r["rollout_env_steps_per_s"] = steps / (ms * 1e-3)
        # greedy, 2^22
        for i in range(2):
            ck(L, L.r48_rollout_policy(1 << 22, 2048 + i, 0, 1, fb.data_ptr(), ln.data_ptr(), None, ws.data_ptr(), st()))
        torch.cuda.synchronize()
        e0.record()
        ck(L, L.r48_rollout_policy(1 << 22, 2048, 0, 1, fb.data_ptr(), ln.data_ptr(), None, ws.data_ptr(), st()))
        e1.record()
        torch.cuda.synchronize()
        r["greedy_env_steps_per_s"] = int(ln[:1 << 22].to(torch.int64).sum().item()) / (e0.elapsed_time(e1) * 1e-3)
        # step, 2^20 boards per launch over 8 rotating windows; and one 2^23-board launch
        pb, pa, po, pr, pd = (t.data_ptr() for t in (boards, acts, out, rw, dn))

        def step_1m(i):
            o = (i % sets) * n
            ck(L, L.r48_step(pb + 8 * o, pa + o, po + 8 * o, pr + 4 * o, pd + o, n, 2048, o, 64, 0, None, st()))

        def step_8m(i):
            ck(L, L.r48_step(pb, pa, po, pr, pd, nb, 2048, 0, 64, 0, None, st()))

        def copy_1m(i):
            o = (i % sets) * n
            ck(L, L.r48_debug_copy22(pb + 8 * o, pa + o, po + 8 * o, pr + 4 * o, pd + o, n, st()))

        def copy_8m(i):
            ck(L, L.r48_debug_copy22(pb, pa, po, pr, pd, nb, st()))

        r["step_1M_us"] = graph_time(step_1m, 64)
        r["step_8M_us"] = graph_time(step_8m, 8)
        r["copy22_1M_us"] = graph_time(copy_1m, 64)
        r["copy22_8M_us"] = graph_time(copy_8m, 8)
        r["step_1M_frac_hbm"] = 22 * n / (r["step_1M_us"] * 1e-6) / 6459.9e9
        r["step_8M_frac_hbm"] = 22 * nb / (r["step_8M_us"] * 1e-6) / 6459.9e9
        for lg in (16, 18, 22):
            m = 1 << lg
            wins = max(1, nb // m)

            def step_m(i, m=m, wins=wins):
                o = (i % wins) * m
                ck(L, L.r48_step(pb + 8 * o, pa + o, po + 8 * o, pr + 4 * o, pd + o, m, 2048, o, 64, 0, None, st()))
            r["step_2^%d_us" % lg] = graph_time(step_m, max(8, min(64, wins)))
        # afterstates 2^23
        aout = torch.empty((4, nb), dtype=torch.int64, device=dev)
        arw = torch.empty((4, nb), dtype=torch.int32, device=dev)

        def after(i):
            ck(L, L.r48_afterstates(pb, aout.data_ptr(), arw.data_ptr(), pd, pd, nb, 0, st()))
        r["afterstates_8M_us"] = graph_time(after, 8)
        r["afterstates_frac_hbm"] = 58 * nb / (r["afterstates_8M_us"] * 1e-6) / 6459.9e9
        del aout, arw
        # env_step 2^20 with obs
        es = torch.zeros(n, dtype=torch.int32, device=dev)
        ee = torch.zeros(n, dtype=torch.int32, device=dev)
        eb = boards[:n].clone()
        obs = torch.empty((n, 4, 4), dtype=torch.float32, device=dev)
        fbz = torch.empty(n, dtype=torch.int64, device=dev)

        def env(i):
            ck(L, L.r48_env_step(eb.data_ptr(), pa, es.data_ptr(), ee.data_ptr(), pr, pd, obs.data_ptr(), 0, fbz.data_ptr(),
                                 n, 2048, 0, n, 0, 1, None, st()))
        r["env_step_1M_us"] = graph_time(env, 32)
        r["env_step_frac_hbm"] = 102 * n / (r["env_step_1M_us"] * 1e-6) / 6459.9e9
        results[name] = r
        print(name, json.dumps(r), flush=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "ab_results.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
