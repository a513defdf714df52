#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- headline benchmark of the B200 2048 environment / random-rollout path.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 (N = 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W        (N > 1, one rank per GPU)
    python bench.py --impl reference [...]                               (CPU arm: the unmodified reference)

Metric (BASELINE.json): 2048 env-steps/sec of fused random-policy rollouts.  One "step" of the
benchmark = one pass of the hot path over one batch: every rank plays `--boards-per-gpu`
episodes (default 2^26 = config 5's per-GPU shard; 2^29 in total at 8 GPUs) from reset to game
over in the fused kernel, reduces them to the statistics vector and, for N > 1, all-reduces
that vector over NCCL.  value = env-steps of all ranks / device time (max over ranks).

The JSON line also carries:
  e2e            the same metric through the host-buffer C-ABI entry point (r48_rollout_host_ex)
                 with one packed record per episode (score, length) + the statistics vector copied
                 to pinned host memory inside the timed region; `e2e_modes` adds the 12-byte
                 full-board records and the statistics-only form
  stats_digest   sha256 of the all-reduced statistics of ONE fixed job (2^22 episodes, seed 2048,
                 ids 0..2^22-1) sharded over the N ranks; rank 0 also plays the whole job alone and
                 asserts the two vectors are equal -- the same digest must appear at N = 1, 2, 4, 8
  roofline       fused rollout kernel vs the integer-issue bound (north_star's bound for it)
  roofline_hbm / kernels   single-step kernel at 1M boards (config 2) with the same-harness
                 22-byte copy ceiling, afterstates at 8M boards (config 4), env_step, the
                 transition ring -- vs the measured HBM copy peak
  cpu_baseline   the unmodified reference (oracle/_ref, byte-compiled by oracle/build_ref.py)
                 timed on this box's host cores; falls back to the Python port if that is absent
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
SEED = 2048


# ---------------------------------------------------------------------- CPU arms (oracle/ is allowed here only)

def cpu_reference(episodes, cores):
    """The reference's own CPU path over all host cores: the unmodified reference from oracle/_ref
    when it was built (kind "reference"), else oracle/pyport.py (kind "port")."""
    from oracle import refarm
    return refarm.timed_rollouts(episodes, cores)            # (steps, seconds, kind)


REF_WHAT = {
    "reference": "the unmodified reference (game/GameClient.py + control/rand.py + main.play), byte-compiled into "
                 "oracle/_ref by oracle/build_ref.py: random.seed(s); Game(); play(game, 'rand')",
    "port": "oracle/pyport.py: Python restatement with the reference's data structures (oracle/_ref was not built)",
}


def cpu_c_port(episodes, threads):
    from oracle import oracle as orc
    orc.lib()
    t0 = time.perf_counter()
    _, ln = orc.rollout(episodes, SEED, 0, threads=threads)
    dt = time.perf_counter() - t0
    return int(ln.sum()), dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path, all host cores, same
    metric/unit/config as our arm.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    cores = host_cores()
    per_step = args.ref_episodes_per_core * cores
    kind = "port"
    for _ in range(args.warmup):
        cpu_reference(max(cores, per_step // 8), cores)
    total_steps, total_dt = 0, 0.0
    for _ in range(args.steps):
        s, dt, kind = cpu_reference(per_step, cores)
        total_steps += s
        total_dt += dt
    value = total_steps / total_dt
    sample = "%d seeded episodes per step (seeds 0..%d), %d steps, multiprocessing.Pool(%d)" % (
        per_step, per_step - 1, args.steps, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "python int",
        "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "what": REF_WHAT[kind]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "episodes_per_sec": per_step * args.steps / total_dt,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n_gpus):
    return {
        "workload": "fused random-policy rollouts to game over (BASELINE configs 3 and 5): %d episodes per GPU, "
                    "%d in total; config 3's 2^24 on one GPU is reported under rollout_config3"
                    % (args.boards_per_gpu, args.boards_per_gpu * n_gpus),
        "boards_per_gpu": args.boards_per_gpu, "episodes_total": args.boards_per_gpu * n_gpus,
        "seed": SEED, "policy": "uniform random (control/rand.py)", "parallelism": "shard%d" % n_gpus,
        "l2": "no inputs; per-episode outputs (12 B x boards_per_gpu) exceed L2; step/afterstates legs rotate "
              "buffer sets larger than L2",
    }


# ---------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; only samples that fall inside the
    timed region are kept."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, ln in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.06):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": max(mx),
                "power_w_max": max(pw), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------- our arm

def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def rollout_issue_profile():
    """Warp-instructions the rollout kernel issues per 32 env-steps, measured by ncu
    (profiles/rollout_issue.json: smsp__inst_executed.sum of launches at several sizes)."""
    path = os.path.join(ROOT, "profiles", "rollout_issue.json")
    if os.path.exists(path):
        return json.load(open(path))
    return None


def ncu_traffic(profile, index=0):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from a committed ncu summary
    (profiles/*.json, one `ncu --set full` capture per kernel); None if the file is missing."""
    path = os.path.join(ROOT, "profiles", profile)
    try:
        d = json.load(open(path))[index]
    except (OSError, IndexError, ValueError, KeyError):
        return None
    total = 0.0
    for key, val in d.items():
        if key.startswith("dram__bytes_read.sum") or key.startswith("dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(key.split("[")[-1].rstrip("]"), None)
            if scale is None:
                return None
            total += float(val) * scale
    return total


def time_launches(torch, fn, iters):
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for i in range(iters):
        fn(i)
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / iters       # ms per launch


def time_graph(torch, fn, launches, replays=10):
    """Capture `launches` calls of fn(i) into one CUDA graph (so that the Python/ctypes launch
    cost, ~10 us per call, is out of the way) and time graph replays with CUDA events on the
    replay stream.  Returns ms per launch."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(launches):                 # warm-up outside capture
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(launches):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(replays):
        g.replay()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / (replays * launches)


def hbm_roofline(alg_bytes, ms, hbm_peak, traffic=None, **extra):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
         "algorithmic_bytes_per_launch": alg_bytes, "traffic": traffic}
    r.update(extra)
    return r


def bench_step_kernel(torch, r48, hbm_peak):
    """Config 2: 1M boards, one batched step per call, host-supplied actions already in HBM.
    Eight rotating buffer sets (8 x 22 MB = 176 MB > 126 MB L2) so every launch reads HBM."""
    n, sets = 1 << 20, 8
    L = r48._native.lib()
    # mid-game boards: 64 random steps from reset
    env = r48.BatchedGame(n * sets, seed=SEED)
    for _ in range(64):
        env.step(torch.randint(0, 4, (n * sets,), device="cuda", dtype=torch.uint8))
    boards_in = env.boards.clone()
    g = torch.Generator(device="cuda").manual_seed(1)
    actions = torch.randint(0, 4, (n * sets,), device="cuda", dtype=torch.uint8, generator=g)
    out = torch.empty_like(boards_in)
    reward = torch.empty(n * sets, dtype=torch.int32, device="cuda")
    done = torch.empty(n * sets, dtype=torch.uint8, device="cuda")
    pb, pa, po, pr, pd = (t.data_ptr() for t in (boards_in, actions, out, reward, done))

    def cur():
        return torch.cuda.current_stream().cuda_stream

    def launch(i):
        o = (i % sets) * n
        r48._native.check(L.r48_step(pb + 8 * o, pa + o, po + 8 * o, pr + 4 * o, pd + o, n, SEED, o, 64, 0, None, cur()))

    def copy(i):
        o = (i % sets) * n
        r48._native.check(L.r48_debug_copy22(pb + 8 * o, pa + o, po + 8 * o, pr + 4 * o, pd + o, n, cur()))
    ms = time_graph(torch, launch, 64)
    ms_copy = time_graph(torch, copy, 64)
    alg_bytes = 22 * n
    res = {"workload": "config 2: 2^20 boards, one step per call, actions in HBM, 8 rotating buffer sets (176 MB > L2), "
                       "64 launches captured in a CUDA graph, timed with CUDA events over 10 replays",
           "us_per_launch": ms * 1e3, "board_steps_per_sec": n / (ms * 1e-3),
           "roofline": hbm_roofline(alg_bytes, ms, hbm_peak, ncu_traffic("r02_step_ncu.json", 0),
                                    traffic_note="ncu, one 2^20-board launch: most of the 13.6 MB of outputs is still in "
                                                 "L2 when the kernel ends"),
           # what 100 % means for a launch of this shape: the same 22 B per board moved by a kernel that
           # computes nothing, in the same 64-launch graph
           "copy_ceiling": {"kernel": "r48::copy22_kernel", "us_per_launch": ms_copy * 1e3,
                            "GBps": alg_bytes / (ms_copy * 1e-3) / 1e9,
                            "frac_of_hbm_peak": alg_bytes / (ms_copy * 1e-3) / 1e9 / hbm_peak,
                            "step_vs_copy": ms_copy / ms}}
    # the same kernel on one 8M-board launch (176 MB): launch latency amortised
    nbig = n * sets
    def launch_big(i):
        r48._native.check(L.r48_step(pb, pa, po, pr, pd, nbig, SEED, 0, 64, 0, None, cur()))
    def copy_big(i):
        r48._native.check(L.r48_debug_copy22(pb, pa, po, pr, pd, nbig, cur()))
    ms_big = time_graph(torch, launch_big, 8)
    ms_copy_big = time_graph(torch, copy_big, 8)
    gbs_big = 22 * nbig / (ms_big * 1e-3) / 1e9
    res["at_8M_boards"] = {"us_per_launch": ms_big * 1e3, "GBps": gbs_big, "frac": gbs_big / hbm_peak,
                           "board_steps_per_sec": nbig / (ms_big * 1e-3),
                           "copy_ceiling_us": ms_copy_big * 1e3,
                           "copy_ceiling_frac_of_hbm_peak": 22 * nbig / (ms_copy_big * 1e-3) / 1e9 / hbm_peak}
    # size sweep (1M boards is launch/staging-latency sized); windows rotate through the 8M buffer
    sweep = {}
    for lg in (16, 18, 19, 20, 21, 22, 23):
        m = 1 << lg
        wins = max(1, nbig // m)
        def launch_m(i, m=m, wins=wins):
            o = (i % wins) * m
            r48._native.check(L.r48_step(pb + 8 * o, pa + o, po + 8 * o, pr + 4 * o, pd + o, m, SEED, o, 64, 0, None, cur()))
        t = time_graph(torch, launch_m, max(8, min(64, wins)))
        sweep["2^%d" % lg] = {"us": t * 1e3, "GBps": 22 * m / (t * 1e-3) / 1e9}
    res["size_sweep"] = sweep
    # the zero-copy Python call a DQN loop makes (BatchedGame.step, device tensors): launch path included
    env1 = r48.BatchedGame(n, seed=SEED)
    env1.boards.copy_(boards_in[:n])
    a1 = actions[:n].contiguous()
    for _ in range(20):
        env1.step(a1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        env1.step(a1)
    torch.cuda.synchronize()
    res["python_call_us"] = (time.perf_counter() - t0) / 200 * 1e6
    # end to end through the host-buffer entry point (pinned host memory, copies inside)
    h_in = boards_in[:n].cpu().pin_memory()
    h_act = actions[:n].cpu().pin_memory()
    h_out = torch.empty(n, dtype=torch.int64).pin_memory()
    h_rw = torch.empty(n, dtype=torch.int32).pin_memory()
    h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    def host_call():
        r48._native.check(L.r48_step_host(h_in.data_ptr(), h_act.data_ptr(), h_out.data_ptr(), h_rw.data_ptr(),
                                          h_dn.data_ptr(), n, SEED, 0, 64, 0, torch.cuda.current_device()))
    for _ in range(3):
        host_call()
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        host_call()
    dt = (time.perf_counter() - t0) / reps
    res["e2e"] = {"value": n / dt, "unit": "board-steps/s", "h2d_bytes_per_step": 9 * n, "d2h_bytes_per_step": 13 * n,
                  "ms_per_call": dt * 1e3}
    return res


def bench_env_step_kernel(torch, r48, hbm_peak):
    """policy-in-the-loop step -- per-env counters, auto-reset, float32 readout fused
    into the epilogue.  2^20 envs, 8 rotating env sets (8 x 102 MB); actions already in HBM.  Second
    figure: the same call with the transition-ring append fused in (8(f).4)."""
    n, sets = 1 << 20, 8
    envs = [r48.BatchedGame(n, seed=SEED + s, board_base=s * n, id_stride=sets * n) for s in range(sets)]
    acts = [torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8) for _ in range(sets)]
    for e, a in zip(envs, acts):
        for _ in range(48):
            e.env_step(a)

    def launch(i):
        envs[i % sets].env_step(acts[i % sets])
    ms = time_graph(torch, launch, 32)
    alg_bytes = 102 * n          # board 8+8, action 1, steps 4+4, episodes 4+4, reward 4, done 1, obs 64
    res = {"workload": "8(f).1: 2^20 envs, Game.step with per-env counters + auto-reset + fused float32 [n,4,4] "
                       "readout, 8 rotating env sets",
           "us_per_launch": ms * 1e3, "board_steps_per_sec": n / (ms * 1e-3),
           "roofline": hbm_roofline(alg_bytes, ms, hbm_peak, ncu_traffic("r02_env_ncu.json", 0))}
    ring = r48.ReplayRing(1 << 24)                 # 16M slots x 22 B = 369 MB (> L2)

    def launch_ring(i):
        envs[i % sets].env_step(acts[i % sets], ring=ring)
    ms_r = time_graph(torch, launch_ring, 32)
    res["with_ring_append"] = {"us_per_launch": ms_r * 1e3,
                               "roofline": hbm_roofline(124 * n, ms_r, hbm_peak, None,
                                                        note="+22 B per env written to the ring (s, a, r, s', done)")}
    return res


def bench_ring(torch, r48, hbm_peak):
    """the transition ring on its own.  append: 2^20 transitions per call into a ring of
    2^24 slots (read 22 B, write one 32-byte record per transition); sample: 2^20 distinct slots
    gathered per call, one random 32-byte record each, written out as separate batch arrays."""
    n, cap = 1 << 20, 1 << 24
    ring = r48.ReplayRing(cap, seed=SEED)
    src = [torch.randint(0, 1 << 62, (n,), device="cuda", dtype=torch.int64) for _ in range(2)]
    a = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8)
    rw = torch.zeros(n, dtype=torch.int32, device="cuda")
    for _ in range(17):
        ring.store(src[0], a, rw, src[1], a)

    def append(i):
        ring.store(src[i & 1], a, rw, src[1 - (i & 1)], a)
    ms_a = time_graph(torch, append, 32)
    L = r48._native.lib()
    out = ring.sample(n)                           # allocates nothing we keep; buffers for the timed loop:
    bufs = {k: torch.empty_like(v) for k, v in out.items()}
    st = lambda: torch.cuda.current_stream().cuda_stream

    def sample(i):
        r48._native.check(L.r48_ring_sample(ring._ref(), n, SEED, i, 0, bufs["index"].data_ptr(), bufs["state"].data_ptr(),
                                            bufs["action"].data_ptr(), bufs["reward"].data_ptr(), bufs["next_state"].data_ptr(),
                                            bufs["done"].data_ptr(), None, None, 0, st()))
    ms_s = time_graph(torch, sample, 16)
    return {"workload": "8(f).4: ring of 2^24 transitions (one 32-byte record each, 537 MB), 2^20 per call",
            "append": {"us_per_launch": ms_a * 1e3, "transitions_per_sec": n / (ms_a * 1e-3),
                       "roofline": hbm_roofline(44 * n, ms_a, hbm_peak, ncu_traffic("r02_ring_append_ncu.json", 0),
                                                note="algorithmic: 22 B read + 22 B of payload written per transition; "
                                                     "the record is padded to one 32-byte sector (54 B moved)")},
            "sample": {"us_per_launch": ms_s * 1e3, "transitions_per_sec": n / (ms_s * 1e-3),
                       "roofline": hbm_roofline(52 * n, ms_s, hbm_peak, ncu_traffic("r02_ring_sample_ncu.json", 0),
                                                note="algorithmic: 22 B of payload gathered + 30 B written (incl. the "
                                                     "int64 slot index) per sample; a random record is one 32-byte "
                                                     "sector, which DRAM serves as a 64-byte access")}}


def bench_afterstates_kernel(torch, r48, hbm_peak):
    """Config 4: 8M boards, all 4 moves + game-over per board; 464 MB per launch (> L2)."""
    n = 1 << 23
    L = r48._native.lib()
    env = r48.BatchedGame(n, seed=SEED)
    for _ in range(64):
        env.step(torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8))
    boards = env.boards
    out = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    reward = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    valid = torch.empty(n, dtype=torch.uint8, device="cuda")
    done = torch.empty(n, dtype=torch.uint8, device="cuda")
    ptrs = tuple(t.data_ptr() for t in (boards, out, reward, valid, done))

    def launch(i):
        r48._native.check(L.r48_afterstates(*ptrs, n, 0, torch.cuda.current_stream().cuda_stream))
    ms = time_graph(torch, launch, 8)
    alg_bytes = 58 * n
    return {"workload": "config 4: 2^23 boards x 4 afterstates + valid mask + done, 464 MB per launch (> L2)",
            "us_per_launch": ms * 1e3, "boards_per_sec": n / (ms * 1e-3),
            "roofline": hbm_roofline(alg_bytes, ms, hbm_peak, ncu_traffic("r02_afterstates_ncu.json", 0))}


def bench_game_adapter(r48):
    """Config 1 through the new backend: the drop-in `Game` (a batch of ONE board, every transition on
    the GPU, the reference's draws injected) playing seeded episodes.  Launch- and sync-bound by
    construction: this is the compatibility surface, not the data path."""
    import random
    steps, t0 = 0, time.perf_counter()
    for s in range(3):
        random.seed(s)
        g = r48.Game()
        over = False
        while not over:
            _, _, over = g.step(r48.Rand.random_action(g.state_matrix))
            steps += 1
    dt = time.perf_counter() - t0
    return {"workload": "config 1 through rein48_b200.Game (rng='python'): seeds 0..2, played to game over",
            "env_steps": steps, "env_steps_per_sec": steps / dt, "us_per_step": dt / steps * 1e6,
            "note": "one launch (r48_step_injected_view: move, spawn, readout and one move of lookahead, through "
                    "pinned buffers) + one stream synchronize per step; the reference's own Python Game.step is "
                    "~27 us (37 k steps/s per core)"}


def stats_digest(torch, dist, r48, world, rank, dev):
    """ONE fixed job -- 2^22 episodes, seed 2048, ids 0..2^22-1 -- sharded over the ranks and all-reduced;
    rank 0 replays the whole job alone and requires the same vector.  Returns (sha256 hex, equal)."""
    import hashlib
    n_total = 1 << 22
    res = r48.sharded_rollouts(n_total, seed=SEED, device=dev)
    reduced = res.stats.clone()
    equal = True
    if rank == 0:
        whole = r48.random_rollouts(n_total, seed=SEED, device=dev).stats
        equal = bool((whole == reduced).all().item())
    digest = hashlib.sha256(reduced.cpu().numpy().tobytes()).hexdigest()
    return digest, equal


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch N>1 with torch.distributed.run" % (args.gpus, world))

    # CPU baselines first (rank 0, N = 1 only), before CUDA is initialised in this process
    cpu_baseline = cpu_c = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        episodes = args.cpu_episodes_per_core * cores
        steps, dt, kind = cpu_reference(episodes, cores)
        cpu_baseline = {"value": steps / dt, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": "%d seeded episodes (seeds 0..%d) of the same workload, multiprocessing.Pool(%d), "
                                  "%.1f s" % (episodes, episodes - 1, cores, dt),
                        "what": REF_WHAT[kind], "episodes_per_sec": episodes / dt}
        c_eps = 60000 * cores
        steps, dt = cpu_c_port(c_eps, cores)
        cpu_c = {"value": steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
                 "sample": "%d episodes, %d pthreads, %.1f s" % (c_eps, cores, dt),
                 "what": "oracle/r48_oracle.c (C restatement, Philox draws)"}

    import torch
    import torch.distributed as dist
    import rein48_b200 as r48

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r48._native.check(r48._native.lib().r48_init(local_rank))

    # N > 1 parity, visible in the line: the same digest at every N
    digest, digest_ok = stats_digest(torch, dist, r48, world, rank, dev)
    if not digest_ok:
        raise SystemExit("the statistics of the job sharded over %d ranks differ from the single-GPU run" % world)

    n = args.boards_per_gpu
    base = rank * n
    L = r48._native.lib()
    buf = r48.RolloutBuffers(n, dev)
    stream = torch.cuda.current_stream().cuda_stream
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    total_stats = torch.zeros(r48.STATS_WORDS, dtype=torch.int64, device=dev)

    def one_step(i, events=None):
        seed = SEED + i
        buf.stats.zero_()
        if events:
            events[0].record()
        r48._native.check(L.r48_rollout(n, seed, base, buf.final_boards.data_ptr(), buf.lengths.data_ptr(), None,
                                        buf.workspace.data_ptr(), stream))
        if events:
            events[1].record()
        r48._native.check(L.r48_episode_stats(buf.final_boards.data_ptr(), buf.lengths.data_ptr(), n,
                                              buf.stats.data_ptr(), stream))
        if events:
            events[2].record()
        r48.allreduce_stats(buf.stats)
        if events:
            events[3].record()
        total_stats.add_(buf.stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                       # nvidia-smi needs ~0.3 s to start: begin before the warm-up
    for i in range(args.warmup):
        one_step(1000 + i)
    total_stats.zero_()
    barrier()
    wall0 = time.time()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        one_step(i, ev[i])
    t_end.record()
    barrier()
    clocks = sampler.stop(wall0, time.time()) if rank == 0 else None
    elapsed_ms = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    k_ms = torch.tensor([sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps,
                         sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps,
                         sum(e[2].elapsed_time(e[3]) for e in ev) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(k_ms, op=dist.ReduceOp.MAX)
    elapsed_s = float(elapsed_ms.item()) * 1e-3
    st = r48.EpisodeStats(total_stats)          # already all-reduced: whole-job counts over K steps
    env_steps = st.steps
    episodes = st.episodes
    value = env_steps / elapsed_s
    buf = None
    torch.cuda.empty_cache()

    # ---- e2e: the public host-buffer API; results land in pinned host memory inside the timed region
    def e2e_leg(make_out, records, d2h_per_episode):
        out = make_out()
        r48.random_rollouts_host(n, seed=SEED + 2000, device=local_rank, board_base=base, out=out, records=records)
        reps = max(1, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        steps_done = 0
        for i in range(reps):
            r48.random_rollouts_host(n, seed=SEED + i, device=local_rank, board_base=base, out=out, records=records)
            steps_done += int(out.stats[r48.stats.SUM_LEN])
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        cnt = torch.tensor([steps_done], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        return {"value": float(cnt.item()) / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": (d2h_per_episode * n + 8 * r48.STATS_WORDS) * world, "steps": reps,
                "ms_per_step": float(dt.item()) / reps * 1e3}

    pin = lambda *shape_dtype: torch.empty(shape_dtype[0], dtype=shape_dtype[1]).pin_memory()
    e2e = e2e_leg(lambda: r48.HostRecords(pin(n, torch.int32), pin(r48.STATS_WORDS, torch.int64)), True, 4)
    e2e.update({
        "api": "rein48_b200.random_rollouts_host(records=True) -> r48_rollout_host_ex (C ABI, host buffers)",
        "payload": "one uint32 per episode (score / 2 << 13 | min(length, 8191): what main.py:48 prints per game, plus "
                   "the episode length) + the 33 KB statistics vector",
        "note": "the path has no per-step host input (episodes are generated from (seed, board id)); the timed region "
                "includes launches, kernels, the D2H copies and the final synchronise"})
    e2e_full = e2e_leg(lambda: r48.RolloutResult(pin(n, torch.int64), pin(n, torch.int32), pin(r48.STATS_WORDS, torch.int64)),
                       False, 12)
    e2e_full["payload"] = "final board (8 B) + length (4 B) per episode + statistics: round 1's e2e payload"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak, peak_src = measured_peaks()
    rollout_ms, stats_ms, reduce_ms = (float(x) for x in k_ms.tolist())
    steps_per_launch = env_steps / args.steps / world          # per GPU
    per_gpu_rate = steps_per_launch / (rollout_ms * 1e-3)
    prof = rollout_issue_profile()
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    roofline = {"bound": "issue", "kernel": "r48::rollout_kernel", "achieved": per_gpu_rate, "unit": "env-steps/s per GPU",
                "peak": None, "frac": None, "traffic": None,
                "ms_per_launch": rollout_ms, "share_of_step": rollout_ms / (elapsed_s * 1e3 / args.steps),
                "algorithmic_bytes_per_launch": 12 * n,
                "hbm_GBps_of_outputs": 12 * n / (rollout_ms * 1e-3) / 1e9,
                "why_not_hbm": "12 B per EPISODE (0.08 B per env-step): HBM traffic is ~0; north_star names integer "
                               "issue slots as this kernel's bound"}
    if prof:
        w = float(prof["warp_instr_per_32_steps"])
        peak = 148 * 4 * sm_mhz * 1e6 * 32 / w
        roofline.update({"peak": peak, "frac": per_gpu_rate / peak,
                         "warp_instr_per_32_steps": w, "instr_source": prof.get("source"),
                         "instr_by_launch_size": prof.get("by_launch_size"),
                         "ipc_per_sm_implied": per_gpu_rate * w / 32 / (148 * sm_mhz * 1e6),
                         "peak_formula": "148 SMs x 4 schedulers x sm_mhz (median under load) x 32 lanes / "
                                         "warp-instructions per 32 env-steps (ncu); frac = issue-slot utilisation",
                         "issue_note": prof.get("issue_note"), "history": prof.get("history"),
                         "traffic": prof.get("dram_bytes_per_launch"),
                         "traffic_note": prof.get("traffic_note")})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_s * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks, "e2e": e2e, "e2e_modes": {"records_4B": e2e["value"], "full_boards_12B": e2e_full},
        "gpu_launches": 2 * args.steps,
        "stats_digest": digest, "stats_digest_job": "2^22 episodes, seed %d, ids 0..2^22-1, sharded over %d rank(s); equal "
                                                    "to rank 0's single-GPU replay: %s" % (SEED, world, digest_ok),
        "episodes_per_sec": episodes / elapsed_s, "env_steps": env_steps, "episodes": episodes,
        "mean_episode_length": st.mean_length, "mean_score": st.mean_score, "max_tile": st.max_tile,
        "kernel_ms": {"rollout": rollout_ms, "episode_stats": stats_ms, "allreduce": reduce_ms},
        "roofline": roofline,
        "hbm_peak": {"GBps": hbm_peak, "source": peak_src},
    }
    if world == 1:
        # statistics-only e2e (33 KB per step to the host)
        sto = pin(r48.STATS_WORDS, torch.int64)
        L.r48_rollout_host_ex(n, SEED, base, 0, None, None, None, sto.data_ptr(), local_rank)
        t0 = time.perf_counter()
        r48._native.check(L.r48_rollout_host_ex(n, SEED + 1, base, 0, None, None, None, sto.data_ptr(), local_rank))
        dt = time.perf_counter() - t0
        line["e2e_modes"]["stats_only"] = {"value": int(sto[r48.stats.SUM_LEN]) / dt, "unit": UNIT,
                                           "d2h_bytes_per_step": 8 * r48.STATS_WORDS, "ms_per_step": dt * 1e3}
    if world == 1 and not args.quick:
        torch.cuda.empty_cache()
        # config 3 exactly: 2^24 episodes on one GPU
        n3 = 1 << 24
        b3 = r48.RolloutBuffers(n3, dev)
        r48.random_rollouts(n3, seed=SEED, buffers=b3)
        ms3 = time_launches(torch, lambda i: r48.random_rollouts(n3, seed=SEED + i, buffers=b3), 3)
        st3 = r48.EpisodeStats(b3.stats)
        line["rollout_config3"] = {"workload": "config 3: 2^24 episodes, 1 GPU (rollout + statistics kernels)",
                                   "ms_per_step": ms3, "env_steps_per_sec": st3.steps / (ms3 * 1e-3),
                                   "episodes_per_sec": n3 / (ms3 * 1e-3)}
        # the same fused kernel with the 1-ply greedy (most blanks) policy
        r48.random_rollouts(n3, seed=SEED, buffers=b3, policy="greedy_blanks")
        msg = time_launches(torch, lambda i: r48.random_rollouts(n3, seed=SEED + i, buffers=b3,
                                                                  policy="greedy_blanks"), 2)
        stg = r48.EpisodeStats(b3.stats)
        line["rollout_greedy_blanks"] = {"workload": "2^24 episodes, 1-ply greedy on blank count (4 afterstates per step)",
                                         "ms_per_step": msg, "env_steps_per_sec": stg.steps / (msg * 1e-3),
                                         "mean_episode_length": stg.mean_length, "mean_score": stg.mean_score,
                                         "max_tile": stg.max_tile}
        b3 = None
        torch.cuda.empty_cache()
        # data generation: every (state, action) of 2^22 random episodes (two passes: play, then replay + write)
        nt = 1 << 22
        r48.rollout_trajectories(nt, seed=SEED)                 # warm the allocator
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        tr = r48.rollout_trajectories(nt, seed=SEED + 1)
        e1.record()
        torch.cuda.synchronize()
        mst = e0.elapsed_time(e1)
        line["trajectories"] = {"workload": "2^22 random episodes, all transitions kept (9 B per step)",
                                "transitions": tr.transitions, "ms": mst,
                                "transitions_per_sec": tr.transitions / (mst * 1e-3),
                                "GBps_written": 9 * tr.boards.numel() / (mst * 1e-3) / 1e9}
        tr = None
        torch.cuda.empty_cache()
        ks = bench_step_kernel(torch, r48, hbm_peak)
        torch.cuda.empty_cache()
        ka = bench_afterstates_kernel(torch, r48, hbm_peak)
        torch.cuda.empty_cache()
        ke = bench_env_step_kernel(torch, r48, hbm_peak)
        torch.cuda.empty_cache()
        kr = bench_ring(torch, r48, hbm_peak)
        line["kernels"] = {"step_1M": ks, "afterstates_8M": ka, "env_step_1M": ke, "ring": kr}
        line["roofline_hbm"] = dict(ks["roofline"], kernel="r48::step_kernel", peak_source=peak_src)
        line["game_adapter"] = bench_game_adapter(r48)
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
        line["cpu_baseline_c"] = cpu_c
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--boards-per-gpu", type=int, default=1 << 26)
    ap.add_argument("--quick", action="store_true", help="skip the config 2/3/4 side legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-episodes-per-core", type=int, default=5000)
    ap.add_argument("--ref-episodes-per-core", type=int, default=1500)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("note: warm-up raised to 3 (timing rule)", file=sys.stderr)
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
