// r48_kernels.cu -- kernels and C ABI of libr48.so (see include/r48.h).
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
//             -Xcompiler -fPIC -shared -o libr48.so r48_kernels.cu
//
// Kernel inventory (DESIGN.md has the roofline of each):
//   build_tables_kernel   row tables (LR: LEFT|RIGHT per row; L16 + merges), once per device
//   reset_kernel          Game.reset                       GameClient.py:33-38
//   step_kernel           Game.step                        GameClient.py:40-51
//   env_step_kernel       step + auto-reset + readout      a3c.py:187-243, ddpg.py:12-70 (worker loops)
//   afterstates_kernel    4 x Game.update_matrix + over    GameClient.py:129-254, 65-94
//   rollout_kernel        main.play(control="rand")        main.py:36-42 + rand.py:9-11
//                         <policy, record>: random | greedy-blanks; play | replay-and-record
//   stats/scores/records/decode/encode  readout             main.py:48, a3c.py:195,205
//   ring_append/ring_sample  Replay.store / Replay.sample   algorithm/ddpg/replay.py:8-47
//   game_view_kernel      Game.step for a host-side player main.py:36-42 over GameClient.py:40-51
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <type_traits>
#include <new>

#include "../../include/r48.h"
#include "r48_device.cuh"

namespace r48 {

// step_kernel A/B switches (tools/ab_kernels.py).  The two-trip loop body executes
// 4 % fewer instructions on early-game boards but issues them more slowly (IPC 2.43 against 2.64) and
// ties at tick 64; the one-trip body is 4 % faster on 2^20 boards, so it is the default.
#ifndef R48_STEP_UNROLL
#define R48_STEP_UNROLL 0          // 1: two trips per loop body, immediate-offset addressing
#endif
#ifndef R48_STEP_PREFETCH
#define R48_STEP_PREFETCH 1        // (two-trip body) L2 prefetch distance in loop bodies, 0 = none
#endif
#ifndef R48_STEP_GATE_EARLY
#define R48_STEP_GATE_EARLY 0      // 1: wait for the first half of the table before the loop instead of per trip
#endif
#ifndef R48_STEP_TABLE_GLOBAL
#define R48_STEP_TABLE_GLOBAL 0    // 1: step_kernel (reward_mode 0) reads the LR table from global memory, no staging
#endif
#ifndef R48_RING_L2_HINT
#define R48_RING_L2_HINT 64          // ring_get: 64 = ld.global.L2::64B (half the DRAM bytes of the default, same time), 1 = also .nc.L1::no_allocate, 0 = plain
#endif
#ifndef R48_UNGUARDED_TICKS
#define R48_UNGUARDED_TICKS 4000u    // rollout: ticks below this cannot hold a 16384 tile (0 = always take the guarded body: test builds)
#endif
#ifndef R48_AFTER_PREFETCH
#define R48_AFTER_PREFETCH 1
#endif

constexpr int kThreads = 1024;                    // one CTA per SM (the table fills its shared memory)
constexpr uint32_t kLeftBytes = 65536 * 2;        // reward-mode tables: LEFT rows (u16) ...
constexpr uint32_t kMergeBytes = 65536;           // ... + merged exponents (u8)
constexpr uint32_t kLrBytes = kLrRows * 4;        // reward-free table: LEFT | RIGHT << 16 per row
constexpr uint32_t kFull = 0xFFFFFFFFu;

// ------------------------------------------------------------------ row tables

// Plain serial restatement of one row sliding LEFT (compress, merge once per pair from
// the left, compress): runs 65536 times per device, never on the hot path.
__device__ uint32_t slow_row_left(uint32_t r, uint32_t &merged)
{
    uint32_t c[4], n = 0, out = 0, o = 0, nm = 0;
    merged = 0;
    for (int t = 0; t < 4; t++) {
        uint32_t e = (r >> (4 * t)) & 15u;
        if (e) c[n++] = e;
    }
    for (uint32_t t = 0; t < n;) {
        if (t + 1 < n && c[t] == c[t + 1]) {
            uint32_t e = c[t] + 1;
            out |= (e > 15u ? 15u : e) << (4 * o++);        // 32768+32768 saturates
            merged |= c[t] << (4 * nm++);
            t += 2;
        } else {
            out |= c[t] << (4 * o++);
            t += 1;
        }
    }
    return out;
}

__device__ uint32_t reverse_row(uint32_t r)
{
    return ((r & 0xFu) << 12) | ((r & 0xF0u) << 4) | ((r >> 4) & 0xF0u) | ((r >> 12) & 0xFu);
}

__global__ void build_tables_kernel(uint16_t *left, uint8_t *merges, uint32_t *lr)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < 65536u) {
        uint32_t m;
        const uint32_t l = slow_row_left(r, m);
        left[r] = (uint16_t)l;
        merges[r] = (uint8_t)m;
        // RIGHT on r = mirror image of LEFT on the mirrored row
        if (r < kLrRows) lr[lr_slot(r)] = l | (reverse_row(slow_row_left(reverse_row(r), m)) << 16);
    }
}

// Stage the tables into shared memory with the bulk-copy engine; returns at once, the
// caller waits on `bar` (parity 0) before its first lookup.
struct Tables {
    const uint16_t *left;        // 65536 x u16
    const uint8_t *merges;       // 65536 x u8
    const uint32_t *lr;          // kLrRows x u32
    PipeConsts pc;
};

// REWARD: [left u16 x 65536][merges u8 x 65536] (192 KB); otherwise [lr u32 x kLrRows] (224 KB).
// Two barriers: bar[0] completes when the first 128 KB are in (the whole LEFT table, or the LR
// rows below kLrSplit), bar[1] when the rest is.
template <bool REWARD>
__device__ __forceinline__ void stage_tables(uint8_t *smem, const Tables &g, uint64_t *bar)
{
    constexpr uint32_t kFirst = REWARD ? kLeftBytes : kLrSplit * 4u;
    constexpr uint32_t kTotal = REWARD ? kLeftBytes + kMergeBytes : kLrBytes;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar[0], kFirst);
        mbar_expect_tx(&bar[1], kTotal - kFirst);
        if (REWARD) {
#pragma unroll
            for (uint32_t off = 0; off < kLeftBytes; off += 32768u)
                bulk_g2s(smem + off, (const uint8_t *)g.left + off, 32768u, &bar[0]);
#pragma unroll
            for (uint32_t off = 0; off < kMergeBytes; off += 32768u)
                bulk_g2s(smem + kLeftBytes + off, g.merges + off, 32768u, &bar[1]);
        } else {
#pragma unroll
            for (uint32_t off = 0; off < kLrBytes; off += 32768u)
                bulk_g2s(smem + off, (const uint8_t *)g.lr + off, 32768u, off < kFirst ? &bar[0] : &bar[1]);
        }
    }
    // The tables never change after r48_init, so staging them does not depend on the previous
    // kernel in the stream; everything after this line may.
    pdl_launch_dependents();
    pdl_wait();
}

// which parts of the table this thread has seen arrive
template <bool REWARD>
struct TableGate {
    uint64_t *bar;
    bool first, second;
    // the reward-mode tables are read through plain pointers: their waits keep the memory clobber
    __device__ __forceinline__ void need_first() { if (!first) { mbar_wait<REWARD>(&bar[0], 0); first = true; } }
    __device__ __forceinline__ void need_second() { if (!second) { mbar_wait<REWARD>(&bar[1], 0); second = true; } }
    __device__ __forceinline__ void need_all() { need_first(); need_second(); }     // also: never exit with a copy in flight
};

template <bool REWARD>
constexpr uint32_t table_bytes() { return REWARD ? kLeftBytes + kMergeBytes : kLrBytes; }

// ------------------------------------------------------------------ reset

struct ResetParams {
    uint64_t *boards;
    int64_t n;
    uint64_t board_base;
    PhiloxKeys keys;
};

// Game.reset: the tick-0 draw on the empty board.  The word's own action field names the axis
// whose order counts the blanks, exactly as the fused rollout does at tick 0.
__device__ __forceinline__ void reset_board(uint32_t &lo, uint32_t &hi, uint32_t a)
{
    lo = 0u; hi = 0u;
    const Blanks b = count_blanks(lo, hi);
    place_tile(lo, hi, b, __umulhi(a << 2, b.n), spawn_exp(a));
    if (is_vertical(a >> 30)) transpose(lo, hi);
}

__global__ void __launch_bounds__(256) reset_kernel(ResetParams p)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t a = draw_word(p.board_base + (uint64_t)i, 0u, p.keys);
        uint32_t lo, hi;
        reset_board(lo, hi, a);
        p.boards[i] = ((uint64_t)hi << 32) | lo;
    }
}

__global__ void __launch_bounds__(256) spawn_injected_kernel(uint64_t *boards, const uint8_t *spawn_k,
                                                             const uint8_t *spawn_exp, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t b = boards[i];
        uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
        const Blanks bl = count_blanks(lo, hi);
        place_tile_checked(lo, hi, bl, spawn_k[i], spawn_exp[i] & 15u);
        boards[i] = ((uint64_t)hi << 32) | lo;
    }
}

// Game.random_fill_grid alone, GPU draws, blanks counted row-major (no move, hence no axis)
__global__ void __launch_bounds__(256) spawn_kernel(ResetParams p, uint32_t tick)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t a = draw_word(p.board_base + (uint64_t)i, tick, p.keys);
        const uint64_t b = p.boards[i];
        uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
        const Blanks bl = count_blanks(lo, hi);
        place_tile(lo, hi, bl, __umulhi(a << 2, bl.n), spawn_exp(a));
        p.boards[i] = ((uint64_t)hi << 32) | lo;
    }
}

__global__ void __launch_bounds__(256) blank_counts_kernel(const uint64_t *__restrict__ boards,
                                                           uint8_t *counts, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t b = boards[i];
        counts[i] = (uint8_t)count_blanks((uint32_t)b, (uint32_t)(b >> 32)).n;
    }
}

// ------------------------------------------------------------------ one call per move for a host-side player

// Game.step with injected draws for a player whose policy AND random draws live on the host (the
// reference's own main.play loop over the `Game` adapter): one launch does the move, the spawn, the
// readout (state_matrix as tile values, reward, has_game_over) and looks one move ahead -- for each
// of the four actions, whether it changes the new board and how many blanks the moved board has,
// which is exactly what the host needs to make the NEXT step's draws in the reference's order
// (GameClient.py:121 draws randint(0, n_blank - 1) only if the move changed the board).  Every
// pointer may be device memory or pinned host memory (the adapter passes pinned buffers, so a
// move costs one launch and one stream synchronize, no copies).  Uses the 16-bit tables straight
// from global memory: a handful of boards does not pay for staging 224 KB.
__global__ void __launch_bounds__(128) game_view_kernel(uint64_t *boards, const uint8_t *action,
                                                        const uint8_t *spawn_k, const uint8_t *spawn_exp,
                                                        int64_t n, int reward_mode, r48_game_view *views,
                                                        int32_t *status, Tables g)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t b = boards[i];
    uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
    const uint32_t act = action[i];
    uint32_t rw = 0;
    bool changed = true;                                   // R48_ACTION_NONE: spawn only (Game.reset)
    if (act < 4u) {
        const uint32_t olo = lo, ohi = hi;
        if (is_vertical(act)) transpose(lo, hi);
        rows_l16<true>(lo, hi, is_toward_high(act), g.left, g.merges, rw);
        if (is_vertical(act)) transpose(lo, hi);
        changed = ((lo ^ olo) | (hi ^ ohi)) != 0u;
    } else if (act != R48_ACTION_NONE) {
        changed = false;                                   // GameClient.py:254: the caller raises
        if (status) atomicOr(status, 1);
    }
    if (changed) place_tile_checked(lo, hi, count_blanks(lo, hi), spawn_k[i], spawn_exp[i] & 15u);
    boards[i] = ((uint64_t)hi << 32) | lo;
    r48_game_view v;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t el = (lo >> (4 * t)) & 15u, eh = (hi >> (4 * t)) & 15u;
        v.cells[t] = el ? (int32_t)(1u << el) : 0;
        v.cells[8 + t] = eh ? (int32_t)(1u << eh) : 0;
    }
    v.reward = reward_mode != 0 ? (int32_t)rw : 0;
    v.done = game_over(lo, hi) ? 1 : 0;
    v.valid = 0;
#pragma unroll
    for (uint32_t a = 0; a < 4; a++) {
        uint32_t l = lo, h = hi, unused;
        if (is_vertical(a)) transpose(l, h);
        const uint32_t tl = l, th = h;
        rows_l16<false>(l, h, is_toward_high(a), g.left, g.merges, unused);
        if ((l ^ tl) | (h ^ th)) v.valid |= (uint8_t)(1u << a);
        v.blanks[a] = (uint8_t)count_blanks(l, h).n;       // the same in either orientation
    }
    v.reserved[0] = v.reserved[1] = 0;
    views[i] = v;
}

// ------------------------------------------------------------------ step

struct StepParams {
    const uint64_t *in;
    const uint8_t *action;
    const uint8_t *spawn_k;      // injected mode only
    const uint8_t *spawn_exp;
    uint64_t *out;
    int32_t *reward;             // may be NULL
    uint8_t *done;               // may be NULL
    int32_t *status;             // may be NULL
    uint32_t n;                  // boards of this launch (the host splits at 2^30 boards ...
    uint32_t id_lo;              // ... and at multiples of 2^32 ids: id.hi is a launch constant)
    PhiloxLaunch pl;
    PhiloxKeys keys;
    Tables tables;
};

// One board through Game.step.  `aw` = the tick's Philox word (or the injected k), `vw` = the
// injected exponent.  Returns whether the board is FULL after the spawn; the caller evaluates
// has_game_over for those (rare) boards -- it is symmetric under transposition, so it does not
// matter that the test runs after the board has been put straight again.  An action byte > 3
// moves nothing useful here; the callers detect it and pass the board through (illegal_action).
//
// The spawn counts blanks in the order of the move's axis, so for UP/DOWN it happens on the
// transposed board, between the two transposes (nothing extra to compute); injected draws index
// the reference's row-major blank list (GameClient.py:109-114), so there the board is transposed
// back first.
// The action byte at bit `SHIFT` of `packed` (two actions travel as one 16-bit load), decoded with
// tests on the packed word itself instead of extracting the byte first: bit 1 clear = UP/DOWN
// (for the legal codes 0..3; anything else is undone by illegal_action), bit 0 = toward the high end.
struct Move {
    uint32_t vertical;              // nonzero for UP / DOWN
    bool toward_high;
    uint32_t pk;                    // PRMT selector of the table halves (pack_selector)
};

template <uint32_t SHIFT>
__device__ __forceinline__ Move decode_move(uint32_t packed)
{
    Move m;
    m.vertical = ~packed & (2u << SHIFT);
    m.toward_high = (packed & (1u << SHIFT)) != 0u;
    m.pk = SHIFT == 0u ? 0x5410u + 0x2222u * (packed & 1u) : (m.toward_high ? 0x7632u : 0x5410u);
    return m;
}

template <bool REWARD, bool INJECT, typename Table>
__device__ __forceinline__ bool step_one(uint32_t &lo, uint32_t &hi, const Move mv, uint32_t aw,
                                         uint32_t vw, const uint8_t *smem, const Table lr, const PipeConsts &pc,
                                         TableGate<REWARD> &gate, int32_t &reward)
{
    const uint32_t vertical = mv.vertical;
    transpose_where(vertical, lo, hi);
    const uint32_t olo = lo, ohi = hi;
    uint32_t rw = 0;
    if (REWARD) rows_l16<true>(lo, hi, mv.toward_high, (const uint16_t *)smem, smem + kLeftBytes, rw);
    else rows_lr(lo, hi, mv.pk, lr, pc, [&] { gate.need_second(); });
    const bool changed = ((lo ^ olo) | (hi ^ ohi)) != 0u;
    bool full;
    if (INJECT) {
        transpose_where(vertical, lo, hi);
        const Blanks b = count_blanks(lo, hi);
        place_tile_checked(lo, hi, b, aw, changed ? (vw & 15u) : 0u);
        full = (any_zero_nibble(lo) | any_zero_nibble(hi)) == 0u;
    } else {
        const Blanks b = count_blanks(lo, hi);
        spawn_tile(lo, hi, b, aw, changed);
        // full after the spawn <=> the moved board had no blank, or exactly one that was filled
        full = b.n == (changed ? 1u : 0u);
        transpose_where(vertical, lo, hi);
    }
    reward = REWARD ? (int32_t)rw : 0;
    return full;
}

// GameClient.py:254 raises ValueError for an action it does not know; a batch cannot raise per
// board, so the board is passed through with reward 0 and the caller's status word is flagged
__device__ __forceinline__ void illegal_action(uint32_t &lo, uint32_t &hi, uint64_t input, int32_t &reward,
                                               uint32_t &done)
{
    lo = (uint32_t)input; hi = (uint32_t)(input >> 32);
    reward = 0;
    done = game_over(lo, hi) ? 1u : 0u;
}

// Work split: every CTA owns one contiguous slice of the batch (equal slices, so all SMs finish
// together even when a launch is only a few trips long -- 2^20 boards are 7 boards per thread);
// a thread takes one unit per trip and loads the unit of its NEXT trip into registers before it
// computes the current one, so that the loads have a whole trip to land.  VEC: a unit
// is a PAIR of boards moved with 128-bit loads/stores (needs 16-byte aligned in/out, 8-byte
// reward, 2-byte action/done); otherwise a unit is one board.  The loop is deliberately bare: the
// kernel runs at the SM's integer issue ceiling, so every bookkeeping instruction per trip is a
// fraction of a percent of its time (profiles/r02_step_budget.txt).
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <uint32_t OFF = 0u>
__device__ __forceinline__ void prefetch_l2_at(uint64_t global_addr)
{
    asm volatile("prefetch.global.L2 [%0+%1];" ::"l"(global_addr), "n"(OFF));
}

// WORD = tick & 3 (see philox_launch_word); injected-draw kernels ignore it
template <bool REWARD, bool INJECT, bool VEC, int WORD>
__global__ void __launch_bounds__(kThreads, 1) step_kernel(StepParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[2];
    // the launch constants of the loop below, pinned in registers (see PinnedWords)
    constexpr int kPinned = 2 * kPhiloxRounds + 4 + 10 + 2;
    __shared__ uint32_t pin_slots[kPinned];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < kPhiloxRounds; r++) {
            sts_u32_at(pin_slots + r, p.keys.k0[r]);
            sts_u32_at(pin_slots + kPhiloxRounds + r, p.keys.k1[r]);
        }
        sts_u32_at(pin_slots + 2 * kPhiloxRounds + 0, p.pl.r1_c1k);
        sts_u32_at(pin_slots + 2 * kPhiloxRounds + 1, p.pl.r1_h0k);
        sts_u32_at(pin_slots + 2 * kPhiloxRounds + 2, p.pl.r2_c3k);
        sts_u32_at(pin_slots + 2 * kPhiloxRounds + 3, p.id_lo);
        sts_u64_at(pin_slots + 2 * kPhiloxRounds + 4, (uint64_t)__cvta_generic_to_global(p.in));
        sts_u64_at(pin_slots + 2 * kPhiloxRounds + 6, (uint64_t)__cvta_generic_to_global(p.action));
        sts_u64_at(pin_slots + 2 * kPhiloxRounds + 8, (uint64_t)__cvta_generic_to_global(p.out));
        sts_u64_at(pin_slots + 2 * kPhiloxRounds + 10, (uint64_t)__cvta_generic_to_global(p.reward));
        sts_u64_at(pin_slots + 2 * kPhiloxRounds + 12, (uint64_t)__cvta_generic_to_global(p.done));
        sts_u32_at(pin_slots + 2 * kPhiloxRounds + 14, p.reward != nullptr ? 1u : 0u);
        sts_u32_at(pin_slots + 2 * kPhiloxRounds + 15, p.done != nullptr ? 1u : 0u);
    }
#if R48_STEP_TABLE_GLOBAL
    // A/B variant: no staging, the LR table is read where it lies through the read-only L1 path
    constexpr bool kStaged = REWARD;
#else
    constexpr bool kStaged = true;
#endif
    if (kStaged) stage_tables<REWARD>(smem, p.tables, bar);            // (its CTA barrier publishes pin_slots)
    else { __syncthreads(); pdl_launch_dependents(); pdl_wait(); }
#if R48_STEP_TABLE_GLOBAL
    const TableInGlobal lr{(uint64_t)__cvta_generic_to_global(p.tables.lr)};
#else
    const uint32_t lr = smem_u32_pinned(smem);
#endif
    TableGate<REWARD> gate{bar, !kStaged, !kStaged};
    PinnedWords<kPinned> pw;
    pw.fetch(pin_slots);

    uint32_t bad = 0;
    const uint32_t units = VEC ? (p.n >> 1) : p.n;
    const uint32_t per = ((units + gridDim.x - 1u) / gridDim.x + 31u) & ~31u;      // whole warps per slice
    const uint32_t begin = min(units, blockIdx.x * per), end = min(units, begin + per);

    PhiloxKeys keys;
    PhiloxLaunch pl = p.pl;
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; r++) { keys.k0[r] = pw.w[r]; keys.k1[r] = pw.w[kPhiloxRounds + r]; }
    pl.r1_c1k = pw.w[2 * kPhiloxRounds + 0]; pl.r1_h0k = pw.w[2 * kPhiloxRounds + 1]; pl.r2_c3k = pw.w[2 * kPhiloxRounds + 2];
    const uint32_t id_lo = pw.w[2 * kPhiloxRounds + 3];
    const uint64_t *const in = p.in;
    const uint8_t *const action = p.action;
    uint64_t *const out = p.out;
    int32_t *const reward = p.reward;
    uint8_t *const done = p.done;
    const uint64_t g_in = pw.u64(2 * kPhiloxRounds + 4), g_action = pw.u64(2 * kPhiloxRounds + 6),
                   g_out = pw.u64(2 * kPhiloxRounds + 8), g_reward = pw.u64(2 * kPhiloxRounds + 10),
                   g_done = pw.u64(2 * kPhiloxRounds + 12);
    const uint32_t has_reward = pw.w[2 * kPhiloxRounds + 14], has_done = pw.w[2 * kPhiloxRounds + 15];

    auto draw = [&](uint32_t board) -> uint32_t {          // the tick's word for board index `board`
        return philox_launch_word<WORD>(id_lo + board, pl, keys);
    };
    auto first_use = [&] {                                 // before the first lookup of a thread
        if (REWARD) gate.need_all(); else gate.need_first();
    };

    // base + index * stride as ONE instruction (IMAD.WIDE); the compiler otherwise spells 64-bit
    // address arithmetic against a base held in registers as an add-with-carry pair
    auto at = [](uint64_t base, uint32_t index, uint32_t stride) -> uint64_t {
        uint64_t a;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(index), "r"(stride), "l"(base));
        return a;
    };
    if (VEC) {
        // One unit (a pair of boards) per thread per trip.  The loop body is TWO trips, A = unit u
        // and B = unit u + kThreads, each with its own registers for the boards and actions in
        // flight: the loads of the next A are issued when A's registers die (right after A is
        // computed), so they have all of B's compute (~3000 cycles) to land, and no register is
        // ever copied from a "next" to a "current" set.  Every address of the body is one of five
        // bases plus a compile-time offset that lives in the instruction (R48_STEP_UNROLL=0 builds
        // the one-trip body with its six moves and ten address instructions per trip for A/B).
        // (a_out/a_reward/a_done: the body's running addresses in the two-trip variant; the one-trip
        // loop passes zeros and the addresses are made from u where they are used)
        auto unit = [&](auto off, const ulonglong2 b, const uint32_t a16, const uint32_t u, uint64_t a_out,
                        uint64_t a_reward, uint64_t a_done) {
            constexpr uint32_t OFF = decltype(off)::value;          // units past the body's base addresses
            uint32_t k0, k1, v0 = 0u, v1 = 0u;
            if (INJECT) {
                const uchar2 kk = ((const uchar2 *)p.spawn_k)[u], vv = ((const uchar2 *)p.spawn_exp)[u];
                k0 = kk.x; k1 = kk.y; v0 = vv.x; v1 = vv.y;
            } else {
                k0 = draw(2u * u); k1 = draw(2u * u + 1u);
            }
            if (!R48_STEP_GATE_EARLY) first_use();
            uint32_t lo0 = (uint32_t)b.x, hi0 = (uint32_t)(b.x >> 32), lo1 = (uint32_t)b.y, hi1 = (uint32_t)(b.y >> 32);
            int32_t r0, r1;
            const bool full0 = step_one<REWARD, INJECT>(lo0, hi0, decode_move<0>(a16), k0, v0, smem, lr, p.tables.pc, gate, r0);
            const bool full1 = step_one<REWARD, INJECT>(lo1, hi1, decode_move<8>(a16), k1, v1, smem, lr, p.tables.pc, gate, r1);
            // Game.has_game_over only where the board is full.  On mid-game boards ~3.5 % are, i.e. most
            // warps have one in some lane and the test's ~25 instructions issue on most trips: one pass
            // tests, in every lane that has a full board, ITS full board (the first of the pair if that
            // one is full, else the second); a lane with both full takes a second pass (1 lane in 800).
            // The test costs the same issue slots with 1 or 32 lanes active.
            uint32_t d0 = 0u, d1 = 0u;
            if (full0 | full1) {
                const uint32_t tl = full0 ? lo0 : lo1, th = full0 ? hi0 : hi1;
                const uint32_t dead = no_equal_neighbours(tl, th) ? 1u : 0u;
                if (full0) d0 = dead; else d1 = dead;
                if (full0 & full1) d1 = no_equal_neighbours(lo1, hi1) ? 1u : 0u;
            }
            if (__builtin_expect((a16 & 0xFCFCu) != 0u, 0)) {            // rare: an action byte > 3
                bad = 1u;
                // (the input pair is read again here rather than kept in four registers for every trip;
                // nothing has been stored yet, so this is right for in-place calls too)
                const ulonglong2 again = ldg_u64x2<0>(at(g_in, u, 16u));
                if ((a16 & 0x00FCu) != 0u) illegal_action(lo0, hi0, again.x, r0, d0);
                if ((a16 & 0xFC00u) != 0u) illegal_action(lo1, hi1, again.y, r1, d1);
            }
            if (!R48_STEP_UNROLL) { a_out = at(g_out, u, 16u); a_reward = at(g_reward, u, 8u); a_done = at(g_done, u, 2u); }
            stg_u64x2<16u * OFF>(a_out, ((uint64_t)hi0 << 32) | lo0, ((uint64_t)hi1 << 32) | lo1);
            if (has_reward) stg_u32x2<8u * OFF>(a_reward, (uint32_t)r0, (uint32_t)r1);
            if (has_done) stg_u16<2u * OFF>(a_done, d0 | (d1 << 8));
        };
        using Off0 = std::integral_constant<uint32_t, 0u>;
#if R48_STEP_UNROLL
        using OffT = std::integral_constant<uint32_t, (uint32_t)kThreads>;
        constexpr uint32_t T = kThreads;
        uint32_t u = begin + threadIdx.x;
        // five running addresses (this thread's unit u of each array) instead of five bases + u
        uint64_t a_in = at(g_in, u, 16u), a_act = at(g_action, u, 2u), a_out = at(g_out, u, 16u),
                 a_reward = at(g_reward, u, 8u), a_done = at(g_done, u, 2u);
        ulonglong2 ba = make_ulonglong2(0ull, 0ull), bb = ba;
        uint32_t aa = 0u, ab = 0u;
        if (u < end) { ba = ldg_u64x2<0>(a_in); aa = ldg_u16<0>(a_act); }
        if (u + T < end) { bb = ldg_u64x2<16u * T>(a_in); ab = ldg_u16<2u * T>(a_act); }
        // the first half of the table is needed from the first lookup on: waiting for it here, with
        // the first loads in flight, instead of testing a "seen it" flag on every trip
        if (R48_STEP_GATE_EARLY) first_use();
        // With the loop this short the kernel is bound by the latency of its own loads (one unit in
        // flight per thread, ~1 us under load: Little's law gives 3.6 TB/s).  A deeper register
        // pipeline does not fit in 64 registers, so the units of the body after next are pulled into
        // L2 by one lane per 128-byte line; the register loads one body ahead then find them there.
        const uint32_t pf_boards = (threadIdx.x & 7u) == 0u ? end : 0u, pf_actions = (threadIdx.x & 31u) == 0u ? end : 0u;
        constexpr uint32_t PF = 2u * T * R48_STEP_PREFETCH;           // units ahead of this body's A
        for (; u < end; u += 2u * T) {
            if (R48_STEP_PREFETCH) {
                if (u + PF < pf_boards) prefetch_l2_at<16u * PF>(a_in);
                if (u + PF + T < pf_boards) prefetch_l2_at<16u * (PF + T)>(a_in);
                if (u + PF < pf_actions) prefetch_l2_at<2u * PF>(a_act);
                if (u + PF + T < pf_actions) prefetch_l2_at<2u * (PF + T)>(a_act);
            }
            unit(Off0{}, ba, aa, u, a_out, a_reward, a_done);
            if (u + 2u * T < end) { ba = ldg_u64x2<32u * T>(a_in); aa = ldg_u16<4u * T>(a_act); }
            if (u + T < end) {
                unit(OffT{}, bb, ab, u + T, a_out, a_reward, a_done);
                if (u + 3u * T < end) { bb = ldg_u64x2<48u * T>(a_in); ab = ldg_u16<6u * T>(a_act); }
            }
            a_in += 32u * T; a_act += 4u * T; a_out += 32u * T; a_reward += 16u * T; a_done += 4u * T;
        }
#else
        ulonglong2 nb = make_ulonglong2(0ull, 0ull);
        uint32_t na16 = 0u;
        if (begin + threadIdx.x < end) {
            nb = ldg_u64x2<0>(at(g_in, begin + threadIdx.x, 16u));
            na16 = ldg_u16<0>(at(g_action, begin + threadIdx.x, 2u));
        }
        if (R48_STEP_GATE_EARLY) first_use();
        for (uint32_t u = begin + threadIdx.x; u < end; u += kThreads) {
            const ulonglong2 b = nb;
            const uint32_t a16 = na16;
            if (u + kThreads < end) {
                nb = ldg_u64x2<0>(at(g_in, u + kThreads, 16u));
                na16 = ldg_u16<0>(at(g_action, u + kThreads, 2u));
            }
            unit(Off0{}, b, a16, u, 0ull, 0ull, 0ull);
        }
#endif
    } else {
        for (uint32_t u = begin + threadIdx.x; u < end; u += kThreads) {
            const uint64_t b = in[u];
            const uint32_t act = action[u];
            const uint32_t k = INJECT ? (uint32_t)p.spawn_k[u] : draw(u);
            const uint32_t v = INJECT ? (uint32_t)p.spawn_exp[u] : 0u;
            first_use();
            uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32), d = 0u;
            int32_t r;
            const bool full = step_one<REWARD, INJECT>(lo, hi, decode_move<0>(act), k, v, smem, lr, p.tables.pc, gate, r);
            if (full) d = no_equal_neighbours(lo, hi) ? 1u : 0u;
            if (__builtin_expect(act > 3u, 0)) { bad = 1u; illegal_action(lo, hi, b, r, d); }
            out[u] = ((uint64_t)hi << 32) | lo;
            if (has_reward) reward[u] = r;
            if (has_done) done[u] = (uint8_t)d;
        }
    }
    // the odd last board of a vectorised launch
    if (VEC && (p.n & 1u) && blockIdx.x == gridDim.x - 1u && threadIdx.x == 0u) {
        const uint32_t i = p.n - 1u;
        const uint64_t b = p.in[i];
        const uint32_t act = p.action[i];
        const uint32_t k = INJECT ? (uint32_t)p.spawn_k[i] : draw(i);
        const uint32_t v = INJECT ? (uint32_t)p.spawn_exp[i] : 0u;
        first_use();
        uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32), d = 0u;
        int32_t r;
        const bool full = step_one<REWARD, INJECT>(lo, hi, decode_move<0>(act), k, v, smem, lr, p.tables.pc, gate, r);
        if (full) d = no_equal_neighbours(lo, hi) ? 1u : 0u;
        if (act > 3u) { bad = 1u; illegal_action(lo, hi, b, r, d); }
        p.out[i] = ((uint64_t)hi << 32) | lo;
        if (p.reward) p.reward[i] = r;
        if (p.done) p.done[i] = (uint8_t)d;
    }
    if (bad && p.status) atomicOr(p.status, 1);
    gate.need_all();                         // never leave with the bulk copy in flight
}

// ------------------------------------------------------------------ vectorised env step

// Optional transition ring (see r48_ring_append): slot (cursor[0] + i) % capacity receives the
// transition of env i; the last CTA to finish adds n to cursor[0].  A slot is one 32-byte record
// (r48_transition: one DRAM sector), written and read as two 128-bit words.
struct RingRefs {
    r48_transition *slots;
    uint64_t *cursor;            // [0] transitions appended so far, [1] CTA ticket (zero between launches)
    uint64_t capacity;
};

// (one 256-bit access per record -- LDG/STG.E.ENL2.256, new with sm_100 -- so that every store
// instruction of a warp writes whole sectors: as two 128-bit halves an append ran at 4.0 TB/s)
__device__ __forceinline__ void ring_put(const RingRefs &r, uint64_t slot, uint64_t state, uint32_t action,
                                         int32_t reward, uint64_t next_state, uint32_t done)
{
    const uint64_t word2 = (uint64_t)(uint32_t)reward | ((uint64_t)(action & 0xFFu) << 32) | ((uint64_t)(done & 0xFFu) << 40);
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};"
                 ::"l"(r.slots + slot), "l"(state), "l"(next_state), "l"(word2), "l"(0ull) : "memory");
}

#pragma nv_diag_suppress 550            // the fourth word of the 256-bit load is padding
__device__ __forceinline__ void ring_get(const RingRefs &r, uint64_t slot, uint64_t &state, uint32_t &action,
                                         int32_t &reward, uint64_t &next_state, uint32_t &done)
{
    uint64_t word2, unused;
#if R48_RING_L2_HINT == 64
    asm volatile("ld.global.L2::64B.v4.u64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(state), "=l"(next_state), "=l"(word2), "=l"(unused) : "l"(r.slots + slot));
#elif R48_RING_L2_HINT == 1
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(state), "=l"(next_state), "=l"(word2), "=l"(unused) : "l"(r.slots + slot));
#else
    asm volatile("ld.global.v4.u64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(state), "=l"(next_state), "=l"(word2), "=l"(unused) : "l"(r.slots + slot));
#endif
    reward = (int32_t)(uint32_t)word2;
    action = (uint32_t)(word2 >> 32) & 0xFFu;
    done = (uint32_t)(word2 >> 40) & 0xFFu;
}
#pragma nv_diag_default 550

// first slot of an append of n (<= capacity unless `skip` says otherwise) transitions and the
// number of leading items that a later item of the same append would overwrite anyway
__device__ __forceinline__ uint64_t ring_start(const RingRefs &r, uint64_t n, uint64_t &skip)
{
    skip = n > r.capacity ? n - r.capacity : 0ull;
    return (r.cursor[0] + skip) % r.capacity;
}

__device__ __forceinline__ void ring_finish(const RingRefs &r, uint64_t n)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long t = atomicAdd((unsigned long long *)&r.cursor[1], 1ull);
        if (t == gridDim.x - 1u) {
            r.cursor[0] += n;
            r.cursor[1] = 0ull;
            __threadfence();
        }
    }
}

struct EnvParams {
    uint64_t *boards;
    const uint8_t *action;
    uint32_t *steps;
    uint32_t *episodes;
    int32_t *reward;             // may be NULL
    uint8_t *done;               // may be NULL
    float *obs;                  // may be NULL
    uint64_t *final_boards;      // may be NULL
    int32_t *status;             // may be NULL
    uint32_t n;
    uint64_t board_base;
    uint64_t id_stride;
    int obs_log2;
    int auto_reset;
    PhiloxKeys keys;
    Tables tables;
    RingRefs ring;               // ring.slots == NULL: no ring
};

// Game.step with per-env tick / episode counters, optional auto-reset, the float readout and the
// replay-ring append fused into the epilogue (one pass over the board instead of step + decode +
// append).
template <bool REWARD>
__global__ void __launch_bounds__(kThreads, 1) env_step_kernel(EnvParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[2];
    stage_tables<REWARD>(smem, p.tables, bar);
    const uint32_t lr = smem_u32_pinned(smem);
    TableGate<REWARD> gate{bar, false, false};

    uint32_t bad = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u;
    uint64_t ring_skip = 0, ring_at = 0;
    if (p.ring.slots) ring_at = ring_start(p.ring, p.n, ring_skip);
    // warp-uniform trip count: the readout below shuffles boards between the lanes of a warp
    // software pipeline, as in the step kernel: the inputs of trip k+1 are loaded before trip k is
    // computed (in place is safe: a trip writes only its own env's slots)
    uint64_t nb = 0ull;
    uint32_t na = 0u, nst = 0u, nep = 0u;
    {
        const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (i0 < p.n) { nb = p.boards[i0]; na = p.action[i0]; nst = p.steps[i0]; nep = p.episodes[i0]; }
    }
    for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x - lane; base < p.n; base += stride) {
        const uint32_t i = base + lane;
        const bool act = i < p.n;
        uint32_t lo = 0u, hi = 0u;
        if (act) {
            const uint64_t b = nb;
            const uint32_t a = na;
            uint32_t st = nst, ep = nep;
            if (i + stride < p.n) {
                nb = p.boards[i + stride]; na = p.action[i + stride];
                nst = p.steps[i + stride]; nep = p.episodes[i + stride];
            }
            uint64_t id = p.board_base + i + (uint64_t)ep * p.id_stride;
            uint32_t aw = draw_word(id, st + 1u, p.keys);
            if (REWARD) gate.need_all(); else gate.need_first();
            lo = (uint32_t)b; hi = (uint32_t)(b >> 32);
            uint32_t d = 0u;
            int32_t r;
            const bool full = step_one<REWARD, false>(lo, hi, decode_move<0>(a), aw, 0u, smem, lr, p.tables.pc, gate, r);
            if (full) d = no_equal_neighbours(lo, hi) ? 1u : 0u;
            if (a > 3u) { bad = 1u; illegal_action(lo, hi, b, r, d); }
            st += 1u;
            if (p.ring.slots && i >= ring_skip) {
                uint64_t slot = ring_at + (i - ring_skip);
                if (slot >= p.ring.capacity) slot -= p.ring.capacity;
                // next_state = the board the step produced, not the auto-reset one
                ring_put(p.ring, slot, b, a, r, ((uint64_t)hi << 32) | lo, d);
            }
            if (d && p.auto_reset) {                 // rare: a lane's game ended
                if (p.final_boards) p.final_boards[i] = ((uint64_t)hi << 32) | lo;
                ep += 1u; st = 0u;
                id = p.board_base + i + (uint64_t)ep * p.id_stride;
                aw = draw_word(id, 0u, p.keys);
                reset_board(lo, hi, aw);
            }
            p.boards[i] = ((uint64_t)hi << 32) | lo;
            p.steps[i] = st;
            p.episodes[i] = ep;
            if (p.reward) p.reward[i] = r;
            if (p.done) p.done[i] = (uint8_t)d;
        }
        if (p.obs) {
            // Readout of the warp's 32 boards (2 KB of float32) as four 512-byte contiguous runs:
            // in run k lane L writes row L%4 of board 8k + L/4, fetched from the owning lane by
            // shuffle.  (Each lane writing its own board's 64 bytes half-fills every sector it
            // touches and stalls the store queue -- see afterstates in DESIGN.md.)
            const uint32_t row = lane & 3u;
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t src = 8u * k + (lane >> 2);
                const uint32_t slo = __shfl_sync(kFull, lo, src), shi = __shfl_sync(kFull, hi, src);
                const uint32_t w = ((row & 2u) ? shi : slo) >> (16u * (row & 1u));
                float v[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const uint32_t e = (w >> (4 * t)) & 15u;
                    v[t] = p.obs_log2 ? (float)e : (float)((1u << e) & ~1u);
                }
                const uint32_t tb = base + src;
                if (tb < p.n) ((float4 *)p.obs)[4ull * tb + row] = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    if (bad && p.status) atomicOr(p.status, 1);
    gate.need_all();
    if (p.ring.slots) ring_finish(p.ring, p.n);
}

// ------------------------------------------------------------------ afterstates

struct AfterParams {
    const uint64_t *in;
    uint64_t *out;               // [4][plane] planar: out[a * plane + i]
    int32_t *reward;             // [4][plane] or NULL
    uint8_t *valid;              // [n] or NULL
    uint8_t *done;               // [n] or NULL
    int64_t n;                   // boards of this launch
    uint64_t plane;              // boards of the whole call = distance between action planes
    Tables tables;
};

template <bool REWARD>
__global__ void __launch_bounds__(kThreads, 1) afterstates_kernel(AfterParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[2];
    stage_tables<REWARD>(smem, p.tables, bar);
    const uint32_t lr = smem_u32_pinned(smem);
    TableGate<REWARD> gate{bar, false, false};

    const uint32_t stride = gridDim.x * blockDim.x, n = (uint32_t)p.n;     // host splits batches >= 2^30
#ifndef R48_AFTER_PIPE
#define R48_AFTER_PIPE 1
#endif
    // This kernel is latency-bound, not bandwidth-bound (one 8-byte load per thread per trip at half
    // occupancy): the board of trip k+1 is loaded before trip k is computed, and the one after that
    // is prefetched into L2.
    uint64_t next = 0ull;
    if (R48_AFTER_PIPE && blockIdx.x * blockDim.x + threadIdx.x < n) next = p.in[blockIdx.x * blockDim.x + threadIdx.x];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t b;
        if (R48_AFTER_PIPE) {
            b = next;
            if (i + stride < n) next = p.in[i + stride];
            if (R48_AFTER_PREFETCH && (threadIdx.x & 3u) == 0u && i + 2u * stride < n) prefetch_l2(p.in + i + 2u * stride);
        } else {
            b = p.in[i];
            if (R48_AFTER_PREFETCH && (threadIdx.x & 3u) == 0u && i + stride < n) prefetch_l2(p.in + i + stride);
        }
        gate.need_all();
        const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
        uint64_t res[4];
        uint32_t rw[4] = {0, 0, 0, 0};
        uint32_t mask = 0;
        if (REWARD) {
#pragma unroll
            for (uint32_t a = 0; a < 4; a++) {
                uint32_t l = lo, h = hi;
                if (is_vertical(a)) transpose(l, h);
                rows_l16<true>(l, h, is_toward_high(a), (const uint16_t *)smem, smem + kLeftBytes, rw[a]);
                if (is_vertical(a)) transpose(l, h);
                if ((l ^ lo) | (h ^ hi)) mask |= 1u << a;
                res[a] = ((uint64_t)h << 32) | l;
            }
        } else {
            uint32_t rl[4], rh[4];
            move_all(lo, hi, lr, rl, rh);
#pragma unroll
            for (uint32_t a = 0; a < 4; a++) {
                if ((rl[a] ^ lo) | (rh[a] ^ hi)) mask |= 1u << a;
                res[a] = ((uint64_t)rh[a] << 32) | rl[a];
            }
        }
#pragma unroll
        for (uint32_t a = 0; a < 4; a++) p.out[a * p.plane + i] = res[a];     // 256 contiguous bytes per warp
        if (p.reward) {
#pragma unroll
            for (uint32_t a = 0; a < 4; a++) p.reward[a * p.plane + i] = (int)rw[a];
        }
        if (p.valid) p.valid[i] = (uint8_t)mask;
        // game over <=> no move changes a non-empty board (proof in DESIGN.md)
        if (p.done) p.done[i] = (uint8_t)(mask == 0u && b != 0ull);
    }
    gate.need_all();
}

// ------------------------------------------------------------------ fused random rollout

struct RolloutParams {
    uint64_t *final_boards;
    uint32_t *lengths;
    unsigned int *counter;           // next unassigned block of episodes (32-bit: the host splits launches)
    uint32_t n;
    uint64_t board_base;
    PhiloxKeys keys;
    Tables tables;
    // trajectory replay (RECORD kernels): lengths[] is then an INPUT, final_boards is unused
    const uint64_t *traj_offsets;    // [n] first slot of each episode (exclusive prefix sum of lengths)
    uint64_t *traj_boards;           // board BEFORE step t of the episode, t = 1..length
    uint8_t *traj_actions;           // the action taken at that step
};

// One lane = one episode at a time, board in two registers; when its game ends the lane takes the
// next episode index (legal because every draw is keyed by the episode's GLOBAL id and tick, not
// by the lane that plays it), so warps stay full until the queue is empty.  A warp draws indices
// from a private block of 32 that it refills with one atomic on the global counter: the common
// switch costs no atomic and no memory round trip.  One Philox call feeds four consecutive ticks,
// and a pass of the loop is two calls (eight ticks) between episode-switch checks.
//
// Orientation is lazy.  The board is kept transposed after an UP/DOWN move and straight after a
// LEFT/RIGHT one (the spawn counts blanks in the order of the move's axis, so it runs on the
// board as stored); a tick transposes only when its axis differs from the previous tick's --
// one 10-instruction transpose on half the ticks instead of two on half the ticks.  The final
// board is put straight before it is written.
//
// Game over is detected lazily: has_game_over (GameClient.py:65-94) holds exactly when no
// move changes the board, and on a FULL board a horizontal move fails iff no two
// horizontal neighbours are equal (same for vertical).  So the lane keeps drawing moves and
// remembers which axes it has seen fail since the board last changed; once both have failed and
// the board is full (tested once per pass, not per tick) the board was dead since the last tick
// that changed it, and that tick is the episode length.  This replaces a ~20-instruction neighbour test per tick by a few predicated ops
// at the price of ~3 extra no-op ticks per episode (2 %); outputs are identical.
enum : int { kPolicyRandom = 0, kPolicyGreedyBlanks = 1 };
constexpr uint32_t kEpisodeBlock = 32;      // episode indices a warp takes per atomic

// number of blank cells of a board (selection only; the spawn uses count_blanks)
__device__ __forceinline__ uint32_t blank_count(uint32_t lo, uint32_t hi)
{
    return __popc(zero_nibbles8(lo) | (zero_nibbles8(hi) >> 1));
}

// RECORD: second pass of r48_rollout_trajectories.  Counter-based draws make an episode
// replayable from (seed, id) alone, so the variable-length trajectories are written by playing
// every episode again once the lengths (hence the output offsets) are known.
template <int POLICY, bool RECORD>
__global__ void __launch_bounds__(kThreads, 1) rollout_kernel(RolloutParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[2];
    // the Philox round keys, pinned in registers for the loop (see PinnedWords: whether ptxas keeps
    // launch constants in uniform registers or re-reads them every iteration changes from build to build)
    constexpr int kPinned = 2 * kPhiloxRounds + 7;
    __shared__ uint32_t pin_slots[kPinned];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < kPhiloxRounds; r++) {
            sts_u32_at(pin_slots + r, p.keys.k0[r]);
            sts_u32_at(pin_slots + kPhiloxRounds + r, p.keys.k1[r]);
        }
        sts_u64_at(pin_slots + 2 * kPhiloxRounds + 0, (uint64_t)__cvta_generic_to_global(p.final_boards));
        sts_u64_at(pin_slots + 2 * kPhiloxRounds + 2, (uint64_t)__cvta_generic_to_global(p.lengths));
        sts_u64_at(pin_slots + 2 * kPhiloxRounds + 4, p.board_base);
        sts_u32_at(pin_slots + 2 * kPhiloxRounds + 6, p.n);
    }
    const uint32_t lr = smem_u32_pinned(smem);
    stage_tables<false>(smem, p.tables, bar);             // (its CTA barrier publishes pin_slots)
    PhiloxKeys keys;
    PinnedWords<kPinned> pw;
    pw.fetch(pin_slots);
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; r++) { keys.k0[r] = pw.w[r]; keys.k1[r] = pw.w[kPhiloxRounds + r]; }
    const uint64_t g_final = pw.u64(2 * kPhiloxRounds + 0), g_lengths = pw.u64(2 * kPhiloxRounds + 2);
    const uint64_t board_base = pw.u64(2 * kPhiloxRounds + 4);
    const uint32_t n_episodes = pw.w[2 * kPhiloxRounds + 6];

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lanes_below = (1u << lane) - 1u;
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    // random policy: `failed` collects the axes seen to fail on the current full board (3 = over);
    // greedy policy: `failed` is simply set to 3 when no move changes the board
    // (before its first episode a lane holds an all-ones "board": full, with failed == 3, so that the
    // first pass of the loop sends every lane to the queue)
    uint32_t lo = 0xFFFFFFFFu, hi = 0xFFFFFFFFu, tick = 0, last_change = 0, failed = 3;
    uint32_t axis_word = 0x80000000u;       // sign bit clear <=> the board is stored transposed
    uint32_t ep = kNone;
    PhiloxEpisode pe = {0u, 0u, 0u};        // per-episode part of the Philox call
    uint32_t id_hi = 0;
    uint32_t blk_next = 0, blk_end = 0;     // warp-uniform: the warp's private block of episode indices
    uint32_t rec_len = 0;       // RECORD: length of the episode being replayed ...
    uint64_t rec_off = 0;       // ... and its first output slot (a multiple of 4)
    // RECORD: four steps are gathered in registers and leave as one full 32-byte sector (two
    // 128-bit stores) + one 32-bit store of the four actions.  One store of 8 bytes per lane per
    // tick, every lane in a different sector, made the replay pass LSU-bound (3x the play pass).
    uint32_t tb_lo[4] = {0u, 0u, 0u, 0u}, tb_hi[4] = {0u, 0u, 0u, 0u}, tact = 0u;
    bool live = true;           // the queue may still have work for this lane
    mbar_wait(&bar[0], 0);
    mbar_wait(&bar[1], 0);

    for (;;) {
        // Episode over: both axes have failed since the last change AND the board is full.  The
        // per-tick bookkeeping records failed axes whether or not the board is full (a failure on a
        // board with blanks proves nothing, but then the board cannot be full here either: it has not
        // changed since); testing fullness once per pass is cheaper than once per tick.
        const bool fin = live && failed == 3u &&
                         (POLICY != kPolicyRandom || (any_zero_nibble(lo) | any_zero_nibble(hi)) == 0u);
        if (__any_sync(kFull, fin)) {
            if (!RECORD && fin && ep != kNone) {
                if ((int32_t)axis_word >= 0) transpose(lo, hi);
                stg_u64(g_final + 8ull * ep, ((uint64_t)hi << 32) | lo);
                stg_u32(g_lengths + 4ull * ep, last_change);
            }
            const uint32_t want = __ballot_sync(kFull, fin);
            const uint32_t cnt = __popc(want), rem = blk_end - blk_next;
            uint32_t fresh = 0;
            if (cnt > rem) {                            // warp-uniform: the private block runs out
                if (lane == 0) fresh = atomicAdd(p.counter, kEpisodeBlock);
                fresh = __shfl_sync(kFull, fresh, 0);
            }
            if (fin) {
                const uint32_t rank = __popc(want & lanes_below);
                const uint32_t mine = rank < rem ? blk_next + rank : fresh + (rank - rem);
                if (mine < n_episodes) {
                    ep = mine; lo = 0; hi = 0; tick = 0; failed = 0; axis_word = 0x80000000u;
                    const uint64_t id = board_base + mine;
                    id_hi = (uint32_t)(id >> 32);
                    pe = philox_episode((uint32_t)id, keys);
                    if (RECORD) { rec_len = p.lengths[mine]; rec_off = p.traj_offsets[mine]; }
                } else {                       // queue empty: park on the empty board (failed stays 3)
                    live = false; ep = kNone; rec_len = 0;
                    lo = 0; hi = 0; tick = 2;
                }
            }
            if (cnt > rem) { blk_next = fresh + (cnt - rem); blk_end = fresh + kEpisodeBlock; }
            else blk_next += cnt;
            if (!__any_sync(kFull, live)) break;
        }

#ifndef R48_CALLS_PER_ITER
#define R48_CALLS_PER_ITER 2          // two Philox calls = eight ticks per pass: half as many episode-switch checks
#endif
#pragma unroll
      for (int call = 0; call < R48_CALLS_PER_ITER; call++) {
        uint32_t w[4];
        philox4x32_episode(id_hi, tick >> 2, pe, keys, w);

        // The tiles of a board sum to at most 4 per spawn, so before tick 4000 no tile can be
        // 16384 and every row is inside the LR table: the range test of rows_lr is provably
        // dead.  One warp vote per four ticks selects the unguarded body (always, in
        // practice: random games last a few hundred ticks); the guarded one stays for the rest.
        auto four_ticks = [&](auto guard_tag) {
            constexpr bool kGuard = decltype(guard_tag)::value;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t aw = w[j];
                uint32_t taken = aw >> 30;
                bool changed;
                uint32_t rec_b_lo = 0, rec_b_hi = 0;        // RECORD: the board before the step, straight
                if (POLICY == kPolicyRandom) {
                    transpose_where((aw ^ axis_word) & 0x80000000u, lo, hi);          // the axis changes
                    axis_word = aw;
                    if (RECORD) { rec_b_lo = lo; rec_b_hi = hi; if ((int32_t)aw >= 0) transpose(rec_b_lo, rec_b_hi); }
                    const uint32_t olo = lo, ohi = hi;
                    rows_lr<kGuard>(lo, hi, pack_selector((aw & 0x40000000u) != 0u), lr, p.tables.pc, [] {});
                    changed = ((lo ^ olo) | (hi ^ ohi)) != 0u;
                    // tick 0 is the reset spawn on the empty board (GameClient.py:33-38); an episode
                    // starts at the top of an iteration, so only j == 0 can be it
                    if (j == 0 && call == 0) changed = changed || (tick == 0u);
                } else {
                    // All four afterstates: the rows of the stored board give one axis, the rows of
                    // its transpose the other; a candidate stays in the orientation it was computed
                    // in.  key = blanks * 4 + (3 - rotation offset), -1 if the move changes nothing:
                    // the maximum is the greedy choice with the first-best tie rule.
                    if (j == 0 && call == 0 && tick == 0u) axis_word = aw;      // the reset draw names its axis
                    const bool stored_t = (int32_t)axis_word >= 0;
                    uint32_t yl = lo, yh = hi;
                    transpose(yl, yh);
                    if (RECORD) { rec_b_lo = stored_t ? yl : lo; rec_b_hi = stored_t ? yh : hi; }
                    const uint32_t r[8] = {lo & 0xFFFFu, lo >> 16, hi & 0xFFFFu, hi >> 16,
                                           yl & 0xFFFFu, yl >> 16, yh & 0xFFFFu, yh >> 16};
                    uint32_t o[8];
                    if (kGuard) {
#pragma unroll
                        for (int t = 0; t < 8; t++) {
                            if (r[t] < kLrRows) o[t] = lds_u32(lr + 4u * lr_slot(r[t]));
                            else o[t] = slow_row(r[t], false) | (slow_row(r[t], true) << 16);
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < 8; t++) o[t] = lds_u32(lr + 4u * lr_slot(r[t]));
                    }
                    uint32_t cl[4], ch[4];
                    cl[0] = prmt(o[0], o[1], 0x5410); ch[0] = prmt(o[2], o[3], 0x5410);   // stored rows, toward low
                    cl[1] = prmt(o[0], o[1], 0x7632); ch[1] = prmt(o[2], o[3], 0x7632);   // stored rows, toward high
                    cl[2] = prmt(o[4], o[5], 0x5410); ch[2] = prmt(o[6], o[7], 0x5410);   // other axis, toward low
                    cl[3] = prmt(o[4], o[5], 0x7632); ch[3] = prmt(o[6], o[7], 0x7632);
                    // action labels: stored rows are UP/DOWN (0,1) when stored transposed, else LEFT/RIGHT (2,3)
                    const uint32_t lab0 = stored_t ? 0u : 2u, lab2 = 2u - lab0;
                    const uint32_t rr = aw >> 30;
                    int best = -1;
                    uint32_t bl = lo, bh = hi, bflip = 0u;
#pragma unroll
                    for (uint32_t c = 0; c < 4; c++) {
                        const uint32_t sl = c < 2 ? lo : yl, sh = c < 2 ? hi : yh;
                        const uint32_t label = (c < 2 ? lab0 : lab2) + (c & 1u);
                        const bool valid = ((cl[c] ^ sl) | (ch[c] ^ sh)) != 0u;
                        const int key = valid ? (int)(blank_count(cl[c], ch[c]) * 4u + (3u - ((label - rr) & 3u))) : -1;
                        if (key > best) { best = key; bl = cl[c]; bh = ch[c]; taken = label; bflip = c < 2 ? 0u : 1u; }
                    }
                    const bool reset_tick = j == 0 && call == 0 && tick == 0u;
                    changed = best >= 0 || reset_tick;
                    if (best < 0 && !reset_tick && failed != 3u) { failed = 3u; last_change = tick - 1u; }
                    lo = bl; hi = bh;
                    if (bflip) axis_word ^= 0x80000000u;
                }
                if (RECORD) {
                    // Ticks enter the loop four at a time starting at a multiple of 4, so the slot
                    // s = tick - 1 of position j sits at (j + 3) & 3 of its 4-slot group.
                    const uint32_t s = tick - 1u;
                    const bool rec = s < rec_len;                // steps 1..length of a live episode
                    constexpr int kShift[4] = {3, 0, 1, 2};
                    const int g = kShift[j];                     // static after unrolling
                    if (rec) {
                        tb_lo[g] = rec_b_lo; tb_hi[g] = rec_b_hi;
                        tact = __byte_perm(tact, taken, g == 0 ? 0x3214 : g == 1 ? 0x3240 : g == 2 ? 0x3410 : 0x4210);
                    }
                    const bool flush = rec && (g == 3 || s + 1u == rec_len);
                    if (flush) {
                        const uint64_t gpos = rec_off + (s & ~3u);
                        uint4 *dst = (uint4 *)(p.traj_boards + gpos);
                        dst[0] = make_uint4(tb_lo[0], tb_hi[0], tb_lo[1], tb_hi[1]);
                        dst[1] = make_uint4(tb_lo[2], tb_hi[2], tb_lo[3], tb_hi[3]);
                        *(uint32_t *)(p.traj_actions + gpos) = tact;
                    }
                }
                const Blanks b = count_blanks(lo, hi);
                spawn_tile(lo, hi, b, aw, changed);
                if (POLICY == kPolicyRandom) {
                    // axis bit: 2 for UP/DOWN (aw >> 31 == 0), 1 for LEFT/RIGHT
                    const uint32_t axis = 2u - (aw >> 31);
                    failed = changed ? 0u : (failed | axis);
                    last_change = changed ? tick : last_change;
                }
                tick++;
            }
        };
        // parked lanes (live == false) sit on the empty board
        if (__all_sync(kFull, tick < R48_UNGUARDED_TICKS || !live)) four_ticks(std::false_type{});
        else four_ticks(std::true_type{});
      }
    }
}

// ------------------------------------------------------------------ episode statistics

// (`records`, optional: the packed per-episode word of records_kernel, written in the same pass --
// the host path wants both and the score is computed here anyway)
__global__ void __launch_bounds__(256) stats_kernel(const uint64_t *__restrict__ boards,
                                                    const uint32_t *__restrict__ lengths, int64_t n,
                                                    unsigned long long *stats, uint32_t *records)
{
    __shared__ uint32_t h_max[16], h_len[2048], h_score[2048];
    __shared__ unsigned long long sums[5];
    for (int t = threadIdx.x; t < 2048; t += blockDim.x) { h_len[t] = 0; h_score[t] = 0; }
    if (threadIdx.x < 16) h_max[threadIdx.x] = 0;
    if (threadIdx.x < 5) sums[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long s_len = 0, s_sc = 0, s_sc2 = 0, s_len2 = 0, cnt = 0;
    // The largest tile of a random game is 128 or 256 nine times out of ten: a shared-memory atomic
    // per episode on that 16-bin histogram serialises ~14 lanes of every warp on one address.  Each
    // thread counts in sixteen 8-bit fields of two registers and flushes them every 255 episodes
    // (0.52 -> 0.48 ms per 2^26 episodes; the rest is the score / max-tile arithmetic -- a 256-entry
    // byte table in shared memory for it was slower: its lookups share the pipe with the atomics).
    unsigned long long mx_lo = 0ull, mx_hi = 0ull;
    uint32_t pending = 0;
    auto flush_max = [&] {
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const uint32_t a = (uint32_t)(mx_lo >> (8 * t)) & 255u, b = (uint32_t)(mx_hi >> (8 * t)) & 255u;
            if (a) atomicAdd(&h_max[t], a);
            if (b) atomicAdd(&h_max[8 + t], b);
        }
        mx_lo = 0ull; mx_hi = 0ull; pending = 0;
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t b = boards[i];
        const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
        const uint32_t sc = board_score(lo, hi), mx = board_max_exp(lo, hi), ln = lengths[i];
        const unsigned long long field = 1ull << (8u * (mx & 7u));
        if (mx < 8u) mx_lo += field; else mx_hi += field;
        if (++pending == 255u) flush_max();
        atomicAdd(&h_len[min(ln, 2047u)], 1u);
        atomicAdd(&h_score[min(sc >> 1, 2047u)], 1u);
        if (records) records[i] = ((sc >> 1) << 13) | min(ln, 8191u);
        cnt += 1; s_len += ln; s_sc += sc;
        s_sc2 += (unsigned long long)sc * sc;
        s_len2 += (unsigned long long)ln * ln;
    }
    flush_max();
    atomicAdd(&sums[0], cnt); atomicAdd(&sums[1], s_len); atomicAdd(&sums[2], s_sc);
    atomicAdd(&sums[3], s_sc2); atomicAdd(&sums[4], s_len2);
    __syncthreads();
    if (threadIdx.x < 5 && sums[threadIdx.x]) atomicAdd(&stats[threadIdx.x], sums[threadIdx.x]);
    if (threadIdx.x < 16 && h_max[threadIdx.x])
        atomicAdd(&stats[R48_STATS_HIST_MAXEXP + threadIdx.x], (unsigned long long)h_max[threadIdx.x]);
    for (int t = threadIdx.x; t < 2048; t += blockDim.x) {
        if (h_len[t]) atomicAdd(&stats[R48_STATS_HIST_LEN + t], (unsigned long long)h_len[t]);
        if (h_score[t]) atomicAdd(&stats[R48_STATS_HIST_SCORE + t], (unsigned long long)h_score[t]);
    }
}

__global__ void __launch_bounds__(256) scores_kernel(const uint64_t *__restrict__ boards,
                                                     uint32_t *score, uint8_t *max_exp, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t b = boards[i];
        if (score) score[i] = board_score((uint32_t)b, (uint32_t)(b >> 32));
        if (max_exp) max_exp[i] = (uint8_t)board_max_exp((uint32_t)b, (uint32_t)(b >> 32));
    }
}

// ------------------------------------------------------------------ readout
// One thread per board ROW: reads 2 bytes of the board, writes one 16-byte vector, so a
// warp stores 512 contiguous bytes.

template <typename T, bool LOG2>
__global__ void __launch_bounds__(256) decode_kernel(const uint64_t *__restrict__ boards, T *out,
                                                     int64_t n)
{
    const int64_t rows = n * 4;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t row = ((const uint16_t *)boards)[r];      // little endian: row r&3 of board r>>2
        T v[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint32_t e = (row >> (4 * t)) & 15u;
            v[t] = LOG2 ? (T)e : (T)((1u << e) & ~1u);
        }
        if (sizeof(T) == 4) {
            uint4 q;
            memcpy(&q, v, 16);
            ((uint4 *)out)[r] = q;
        }
    }
}

__global__ void __launch_bounds__(256) encode_kernel(const int32_t *__restrict__ values,
                                                     uint64_t *boards, int64_t n, int32_t *status)
{
    uint32_t bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t b = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int4 v = ((const int4 *)values)[4 * i + q];
            const int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int x = vv[t];
                uint32_t e = 0;
                if (x != 0) {
                    if (x < 2 || (x & (x - 1)) || x > 32768) bad = 1;
                    else e = 31u - __clz(x);
                }
                b |= (uint64_t)e << (4 * (4 * q + t));
            }
        }
        boards[i] = b;
    }
    if (bad && status) atomicOr(status, 2);
}

// ------------------------------------------------------------------ compact episode records
// One uint32 per episode for the host: score / 2 in bits 31..13 (the sum of a board's tiles is even
// and at most 16 * 32768 = 2^19, so 19 bits hold it exactly), min(length, 8191) in bits 12..0.
__global__ void __launch_bounds__(256) records_kernel(const uint64_t *__restrict__ boards,
                                                      const uint32_t *__restrict__ lengths,
                                                      uint32_t *records, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t b = boards[i];
        const uint32_t sc = board_score((uint32_t)b, (uint32_t)(b >> 32));
        records[i] = ((sc >> 1) << 13) | min(lengths[i], 8191u);
    }
}

// ------------------------------------------------------------------ transition ring
// Replay.store / Replay.sample (algorithm/ddpg/replay.py:8-47) for batches of transitions, as one
// device array of `capacity` 32-byte records (state, next_state, reward, action, done): a random
// slot is one DRAM sector.  (Round 2's first layout, five arrays, made a sample gather five sectors
// -- 431 MB of DRAM reads for 23 MB of sampled transitions.)

struct RingAppendParams {
    RingRefs ring;
    const uint64_t *state;
    const uint8_t *action;
    const int32_t *reward;       // may be NULL (stored as 0, the reference's reward)
    const uint64_t *next_state;
    const uint8_t *done;         // may be NULL (stored as 0)
    uint64_t n;
    int drop_when_full;          // Replay.store: a full buffer ignores the transition (replay.py:18-21)
};

__global__ void __launch_bounds__(256) ring_append_kernel(RingAppendParams p)
{
    const uint64_t cap = p.ring.capacity, cur = p.ring.cursor[0];
    uint64_t skip = 0, at = 0, take = p.n;
    if (p.drop_when_full) {
        const uint64_t room = cur < cap ? cap - cur : 0ull;
        take = p.n < room ? p.n : room;                // the first `take` transitions fit
        at = cur;
    } else {
        at = ring_start(p.ring, p.n, skip);
    }
    for (uint64_t i = skip + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < take;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t slot = at + (i - skip);
        if (slot >= cap) slot -= cap;
        ring_put(p.ring, slot, p.state[i], p.action[i], p.reward ? p.reward[i] : 0, p.next_state[i],
                 p.done ? (uint32_t)p.done[i] : 0u);
    }
    ring_finish(p.ring, take);
}

// A keyed pseudo-random permutation of [0, size): six Feistel rounds on 2h >= log2(size) bits,
// cycle-walked back into range.  perm(0), perm(1), ... are distinct by construction, so the
// first B values are a sample WITHOUT replacement (random.sample, replay.py:33).
struct RingPerm {
    uint32_t k0[6], k1[6];
    uint32_t half_bits;
    uint64_t size;
};

__device__ __forceinline__ uint64_t ring_perm(const RingPerm &P, uint64_t x)
{
    const uint32_t h = P.half_bits;
    const uint32_t mask = h >= 32u ? 0xFFFFFFFFu : ((1u << h) - 1u);
    do {
        uint32_t L = (uint32_t)(x >> h) & mask, R = (uint32_t)x & mask;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const uint64_t m = (uint64_t)(R + P.k0[r]) * R48_PHILOX_M0;
            const uint32_t f = ((uint32_t)(m >> 32) ^ (uint32_t)m ^ P.k1[r]) & mask;
            const uint32_t t = L ^ f;
            L = R; R = t;
        }
        x = ((uint64_t)L << h) | R;
    } while (x >= P.size);
    return x;
}

struct RingSampleParams {
    RingRefs ring;
    uint64_t batch;
    uint64_t seed, draw;
    int with_replacement;
    PhiloxKeys keys;
    int64_t *out_index;          // may be NULL
    uint64_t *out_state;
    uint8_t *out_action;
    int32_t *out_reward;
    uint64_t *out_next;
    uint8_t *out_done;
    float *obs_state;            // may be NULL: float32 [batch][4][4]
    float *obs_next;             // may be NULL
    int obs_log2;
};

// round keys of the permutation: words 0,1 of Philox(ctr = (draw.lo, draw.hi, r, 'RING'), key = seed)
__device__ inline void ring_perm_keys(RingPerm &P, uint64_t size, uint64_t draw, const uint32_t (&k0)[kPhiloxRounds],
                                               const uint32_t (&k1)[kPhiloxRounds])
{
    uint32_t bits = 2;
    while (bits < 64u && ((uint64_t)1 << bits) < size) bits++;
    P.half_bits = (bits + 1u) / 2u;
    P.size = size;
    for (int r = 0; r < 6; r++) {
        uint32_t c0 = (uint32_t)draw, c1 = (uint32_t)(draw >> 32), c2 = (uint32_t)r, c3 = 0x52494E47u;
        for (int q = 0; q < kPhiloxRounds; q++) {
            const uint64_t p0 = (uint64_t)R48_PHILOX_M0 * c0, p1 = (uint64_t)R48_PHILOX_M1 * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0[q], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1[q];
            c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        }
        P.k0[r] = c0; P.k1[r] = c1;
    }
}

__device__ __forceinline__ void decode_row4(uint32_t w, int log2_planes, float (&v)[4])
{
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const uint32_t e = (w >> (4 * t)) & 15u;
        v[t] = log2_planes ? (float)e : (float)((1u << e) & ~1u);
    }
}

__global__ void __launch_bounds__(256) ring_sample_kernel(RingSampleParams p)
{
    __shared__ RingPerm sperm;
    const uint64_t cap = p.ring.capacity, cur = p.ring.cursor[0];
    const uint64_t size = cur < cap ? cur : cap;
    if (threadIdx.x == 0) {
        ring_perm_keys(sperm, size > 0 ? size : 1, p.draw, p.keys.k0, p.keys.k1);
    }
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.batch;
         i += (uint64_t)gridDim.x * blockDim.x) {
        int64_t idx = -1;
        if (size > 0) {
            if (p.with_replacement) {
                // element i of draw `draw`: words 0,1 of Philox(ctr = (i.lo, i.hi, draw.lo, 'SMPL' ^ draw.hi))
                uint32_t w[4];
                philox4x32((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)p.draw, 0x534D504Cu ^ (uint32_t)(p.draw >> 32), p.keys, w);
                idx = (int64_t)__umul64hi(((uint64_t)w[0] << 32) | w[1], size);
            } else if (i < size) {
                idx = (int64_t)ring_perm(sperm, i);      // i >= size: the reference returns the whole list
            }
        }
        if (p.out_index) p.out_index[i] = idx;
        if (idx < 0) continue;
        uint64_t s, nx;
        uint32_t act, dn;
        int32_t rw;
        ring_get(p.ring, (uint64_t)idx, s, act, rw, nx, dn);
        p.out_state[i] = s;
        p.out_next[i] = nx;
        p.out_action[i] = (uint8_t)act;
        p.out_reward[i] = rw;
        p.out_done[i] = (uint8_t)dn;
        if (p.obs_state || p.obs_next) {
#pragma unroll
            for (int row = 0; row < 4; row++) {
                float v[4];
                if (p.obs_state) {
                    decode_row4((uint32_t)(s >> (16 * row)) & 0xFFFFu, p.obs_log2, v);
                    ((float4 *)p.obs_state)[4ull * i + row] = make_float4(v[0], v[1], v[2], v[3]);
                }
                if (p.obs_next) {
                    decode_row4((uint32_t)(nx >> (16 * row)) & 0xFFFFu, p.obs_log2, v);
                    ((float4 *)p.obs_next)[4ull * i + row] = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    }
}

__global__ void ring_clear_kernel(uint64_t *cursor) { cursor[0] = 0ull; cursor[1] = 0ull; }

// 22 B per board in, 22 B out, nothing else: the same memory traffic as step_kernel with no work
// in between -- what "100 % of HBM" means for a launch of this shape and size (bench.py).
__global__ void __launch_bounds__(kThreads, 1) copy22_kernel(const uint64_t *__restrict__ in,
                                                             const uint8_t *__restrict__ action,
                                                             uint64_t *out, int32_t *reward, uint8_t *done, uint32_t n)
{
    const uint32_t pairs = n >> 1;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < pairs; j += gridDim.x * blockDim.x) {
        const ulonglong2 b = ((const ulonglong2 *)in)[j];
        const uchar2 a = ((const uchar2 *)action)[j];
        ((ulonglong2 *)out)[j] = b;
        ((int2 *)reward)[j] = make_int2(a.x, a.y);
        ((uchar2 *)done)[j] = a;
    }
}

}  // namespace r48

using namespace r48;

namespace {

// ------------------------------------------------------------------ host side

thread_local char g_err[256] = "";

int fail(int code, const char *msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

int fail_cuda(cudaError_t e, const char *where)
{
    snprintf(g_err, sizeof g_err, "%s: %s", where, cudaGetErrorString(e));
    return R48_ERR_CUDA;
}

#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call);        \
    } while (0)

constexpr int kMaxDevices = 64;

struct DeviceState {
    bool ready = false;
    int sms = 0;
    uint16_t *left = nullptr;
    uint8_t *merges = nullptr;
    uint32_t *lr = nullptr;
    Tables tables() const { return Tables{left, merges, lr, PipeConsts{1u << 16, 1u << 18}}; }
    // host-API arena (guarded by g_host_mu[device])
    uint8_t *arena = nullptr;
    size_t arena_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // D2H of finished chunks overlaps the next chunk's kernel
    cudaStream_t stream2 = nullptr;          // rollout chunks alternate between `stream` and this one (see r48_rollout_host_ex)
    cudaEvent_t chunk_done[2] = {nullptr, nullptr};
    cudaEvent_t stats_done[2] = {nullptr, nullptr};
    cudaEvent_t join = nullptr;
    cudaEvent_t probe[4] = {nullptr, nullptr, nullptr, nullptr};   // around the first chunk's kernel and its copy
    float copy_over_play = 0.0f;             // measured by the previous r48_rollout_host_ex call on this device
};

DeviceState g_dev[kMaxDevices];
std::mutex g_mu;                         // guards g_dev[].ready / table construction
std::mutex g_host_mu[kMaxDevices];       // one *_host call at a time per device (arena, streams, events)

// switches the current device for a scope and always switches back, error paths included
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

using StepKernel = void (*)(StepParams);

// the instantiation for (reward_mode, injected draws, vector loads, tick & 3)
template <bool REWARD, bool INJECT, bool VEC>
StepKernel step_kernel_word(int word)
{
    if (INJECT) return step_kernel<REWARD, INJECT, VEC, 0>;
    switch (word & 3) {
    case 0: return step_kernel<REWARD, INJECT, VEC, 0>;
    case 1: return step_kernel<REWARD, INJECT, VEC, 1>;
    case 2: return step_kernel<REWARD, INJECT, VEC, 2>;
    default: return step_kernel<REWARD, INJECT, VEC, 3>;
    }
}

StepKernel step_kernel_ptr(bool reward, bool inject, bool vec, int word)
{
    if (reward) {
        if (inject) return vec ? step_kernel_word<true, true, true>(word) : step_kernel_word<true, true, false>(word);
        return vec ? step_kernel_word<true, false, true>(word) : step_kernel_word<true, false, false>(word);
    }
    if (inject) return vec ? step_kernel_word<false, true, true>(word) : step_kernel_word<false, true, false>(word);
    return vec ? step_kernel_word<false, false, true>(word) : step_kernel_word<false, false, false>(word);
}

template <typename K>
cudaError_t opt_in_smem(K kernel, uint32_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int init_device_locked(int dev, DeviceState &d)
{
    DeviceGuard g(dev);
    CK(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaMalloc(&d.left, kLeftBytes));
    CK(cudaMalloc(&d.merges, kMergeBytes));
    CK(cudaMalloc(&d.lr, kLrBytes));
    build_tables_kernel<<<256, 256>>>(d.left, d.merges, d.lr);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    for (int reward = 0; reward < 2; reward++)
        for (int inject = 0; inject < 2; inject++)
            for (int vec = 0; vec < 2; vec++)
                for (int word = 0; word < (inject ? 1 : 4); word++)
                    CK(opt_in_smem(step_kernel_ptr(reward != 0, inject != 0, vec != 0, word),
                                   reward ? kLeftBytes + kMergeBytes : kLrBytes));
    CK(opt_in_smem(env_step_kernel<false>, kLrBytes));
    CK(opt_in_smem(env_step_kernel<true>, kLeftBytes + kMergeBytes));
    CK(opt_in_smem(afterstates_kernel<false>, kLrBytes));
    CK(opt_in_smem(afterstates_kernel<true>, kLeftBytes + kMergeBytes));
    CK(opt_in_smem(rollout_kernel<kPolicyRandom, false>, kLrBytes));
    CK(opt_in_smem(rollout_kernel<kPolicyGreedyBlanks, false>, kLrBytes));
    CK(opt_in_smem(rollout_kernel<kPolicyRandom, true>, kLrBytes));
    CK(opt_in_smem(rollout_kernel<kPolicyGreedyBlanks, true>, kLrBytes));
    return R48_OK;
}

int ensure_device(int dev, DeviceState **out)
{
    if (dev < 0 || dev >= kMaxDevices) return fail(R48_ERR_ARG, "device index out of range");
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceState &d = g_dev[dev];
    if (!d.ready) {
        const int rc = init_device_locked(dev, d);      // the guard inside restores the caller's device
        if (rc) return rc;
        d.ready = true;
    }
    *out = &d;
    return R48_OK;
}

int current_device(DeviceState **out)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    return ensure_device(dev, out);
}

PhiloxKeys make_keys(uint64_t seed)
{
    PhiloxKeys k;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < kPhiloxRounds; r++) {
        k.k0[r] = k0;
        k.k1[r] = k1;
        k0 += R48_PHILOX_W0;
        k1 += R48_PHILOX_W1;
    }
    return k;
}

// the launch-constant half of Philox rounds 0..2 for (id.hi, tick): see philox_launch_word
PhiloxLaunch make_philox_launch(const PhiloxKeys &k, uint32_t id_hi, uint32_t tick)
{
    PhiloxLaunch p;
    const uint64_t p1 = (uint64_t)R48_PHILOX_M1 * (tick >> 2);
    p.r0_n0 = (uint32_t)(p1 >> 32) ^ id_hi ^ k.k0[0];
    p.r1_c1k = (uint32_t)p1 ^ k.k0[1];
    const uint64_t q0 = (uint64_t)R48_PHILOX_M0 * p.r0_n0;
    p.r1_h0k = (uint32_t)(q0 >> 32) ^ k.k1[1];
    p.r2_c3k = (uint32_t)q0 ^ k.k1[2];
    p.word = tick & 3u;
    p.last_mul = p.word >= 2u ? R48_PHILOX_M0 : R48_PHILOX_M1;
    p.last_key = p.word >= 2u ? k.k1[kPhiloxRounds - 1] : k.k0[kPhiloxRounds - 1];
    return p;
}

inline bool aligned(const void *p, size_t a) { return ((uintptr_t)p & (a - 1)) == 0; }

// grid for a grid-stride kernel of `threads`-wide blocks: enough blocks for the work,
// capped at `per_sm` blocks per SM
int grid_for(int64_t work_items, int threads, int sms, int per_sm)
{
    int64_t blocks = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int check_n(int64_t n)
{
    if (n < 0) return fail(R48_ERR_ARG, "n < 0");
    if (n > ((int64_t)1 << 40)) return fail(R48_ERR_ARG, "n too large");
    return R48_OK;
}

constexpr int64_t kChunk = (int64_t)1 << 30;     // boards per launch (kernels index with 32 bits)

// Launch with programmatic stream serialisation: the CTAs of this kernel may be scheduled, and
// stage their tables, as SMs of the previous kernel in the stream free up (see stage_tables).
template <typename P>
cudaError_t launch_pdl(void (*kernel)(P), int grid, int block, uint32_t smem, cudaStream_t s, const P &params)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, params);
}

// `whole` describes the full batch (n in whole_n, first id in board_base); launches are cut at
// 2^30 boards and wherever the global id crosses a multiple of 2^32, so that id.hi is a constant
// of every launch (PhiloxLaunch).
template <bool INJECT>
int launch_step(StepParams whole, int64_t whole_n, uint64_t seed, uint64_t board_base, uint32_t tick,
                int reward_mode, const DeviceState &d, cudaStream_t s)
{
    const PhiloxKeys keys = make_keys(seed);
    for (int64_t off = 0; off < whole_n;) {
        const uint64_t id0 = board_base + (uint64_t)off;
        int64_t m = whole_n - off < kChunk ? whole_n - off : kChunk;
        const uint64_t to_wrap = ((uint64_t)1 << 32) - (id0 & 0xFFFFFFFFull);
        if (!INJECT && (uint64_t)m > to_wrap) m = (int64_t)to_wrap;
        StepParams p = whole;
        p.n = (uint32_t)m;
        p.in += off; p.action += off; p.out += off;
        if (p.reward) p.reward += off;
        if (p.done) p.done += off;
        if (INJECT) { p.spawn_k += off; p.spawn_exp += off; }
        p.id_lo = (uint32_t)id0;
        p.keys = keys;
        p.pl = make_philox_launch(keys, (uint32_t)(id0 >> 32), tick);
        const bool vec = aligned(p.in, 16) && aligned(p.out, 16) && aligned(p.action, 2) &&
                         (!p.reward || aligned(p.reward, 8)) && (!p.done || aligned(p.done, 2)) &&
                         (!INJECT || (aligned(p.spawn_k, 2) && aligned(p.spawn_exp, 2)));
        // one CTA per SM as soon as there is a warp's worth of units for each (a CTA costs the same
        // table staging whether it has work for 1 warp or for 32)
        const int64_t units = vec ? (m + 1) / 2 : m;
        const int grid = grid_for(units, 32, d.sms, 1);
        const uint32_t smem = reward_mode ? kLeftBytes + kMergeBytes : (R48_STEP_TABLE_GLOBAL ? 0u : kLrBytes);
        CK(launch_pdl(step_kernel_ptr(reward_mode != 0, INJECT, vec, (int)(tick & 3u)), grid, kThreads, smem, s, p));
        off += m;
    }
    return R48_OK;
}

// the arena only grows; callers hold g_host_mu[device]
int arena_reserve(DeviceState &d, size_t bytes)
{
    if (!d.stream) CK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    if (!d.copy_stream) {
        CK(cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&d.chunk_done[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&d.chunk_done[1], cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&d.stream2, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&d.join, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&d.stats_done[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&d.stats_done[1], cudaEventDisableTiming));
        for (int e = 0; e < 4; e++) CK(cudaEventCreate(&d.probe[e]));
    }
    if (d.arena_bytes >= bytes) return R48_OK;
    if (d.arena) CK(cudaFree(d.arena));
    d.arena = nullptr;
    d.arena_bytes = 0;
    size_t want = bytes + bytes / 8 + 4096;
    CK(cudaMalloc(&d.arena, want));
    d.arena_bytes = want;
    return R48_OK;
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

int ring_refs(const r48_ring *ring, RingRefs *out, const char *who)
{
    if (!ring || !ring->slots || !ring->cursor) return fail(R48_ERR_NULL, who);
    if (ring->capacity == 0 || ring->capacity > ((uint64_t)1 << 40)) return fail(R48_ERR_ARG, "ring capacity must be in 1 .. 2^40");
    if (!aligned(ring->slots, 32) || !aligned(ring->cursor, 8))
        return fail(R48_ERR_ALIGN, "ring slots must be 32-byte aligned, the cursor 8-byte aligned");
    *out = RingRefs{ring->slots, ring->cursor, ring->capacity};
    return R48_OK;
}

}  // namespace

// ==================================================================== C ABI

extern "C" {

int r48_version(void) { return R48_VERSION; }

#ifndef R48_BUILD_ID
#define R48_BUILD_ID "unknown"
#endif
const char *r48_build_id(void) { return "R48_BUILD_ID=" R48_BUILD_ID; }

const char *r48_last_error(void) { return g_err; }

int r48_init(int device)
{
    DeviceState *d;
    return ensure_device(device, &d);
}

int r48_debug_tables_host(uint16_t *left, uint8_t *merges, int device)
{
    if (!left || !merges) return fail(R48_ERR_NULL, "r48_debug_tables_host: NULL output");
    DeviceState *d;
    int rc = ensure_device(device, &d);
    if (rc) return rc;
    DeviceGuard g(device);
    CK(cudaMemcpy(left, d->left, kLeftBytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(merges, d->merges, kMergeBytes, cudaMemcpyDeviceToHost));
    return R48_OK;
}

int r48_reset(uint64_t *boards, int64_t n, uint64_t seed, uint64_t board_base, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!boards) return fail(R48_ERR_NULL, "r48_reset: boards is NULL");
    if (!aligned(boards, 8)) return fail(R48_ERR_ALIGN, "r48_reset: boards not 8-byte aligned");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    ResetParams p{boards, n, board_base, make_keys(seed)};
    reset_kernel<<<grid_for(n, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(p);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_step(const uint64_t *in, const uint8_t *action, uint64_t *out, int32_t *reward,
             uint8_t *done, int64_t n, uint64_t seed, uint64_t board_base, uint32_t step,
             int reward_mode, int32_t *status, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (reward_mode != 0 && reward_mode != 1) return fail(R48_ERR_ARG, "r48_step: reward_mode must be 0 or 1");
    if (n == 0) return R48_OK;
    if (!in || !action || !out) return fail(R48_ERR_NULL, "r48_step: in/action/out is NULL");
    if (!aligned(in, 8) || !aligned(out, 8) || (reward && !aligned(reward, 4)) ||
        (status && !aligned(status, 4)))
        return fail(R48_ERR_ALIGN, "r48_step: misaligned pointer");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    StepParams p{};
    p.in = in; p.action = action; p.out = out; p.reward = reward; p.done = done; p.status = status;
    p.tables = d->tables();
    return launch_step<false>(p, n, seed, board_base, step + 1u, reward_mode, *d, (cudaStream_t)stream);
}

int r48_step_injected(const uint64_t *in, const uint8_t *action, const uint8_t *spawn_k,
                      const uint8_t *spawn_exp, uint64_t *out, int32_t *reward, uint8_t *done,
                      int64_t n, int reward_mode, int32_t *status, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (reward_mode != 0 && reward_mode != 1) return fail(R48_ERR_ARG, "r48_step_injected: reward_mode must be 0 or 1");
    if (n == 0) return R48_OK;
    if (!in || !action || !out || !spawn_k || !spawn_exp)
        return fail(R48_ERR_NULL, "r48_step_injected: NULL pointer");
    if (!aligned(in, 8) || !aligned(out, 8) || (reward && !aligned(reward, 4)) ||
        (status && !aligned(status, 4)))
        return fail(R48_ERR_ALIGN, "r48_step_injected: misaligned pointer");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    StepParams p{};
    p.in = in; p.action = action; p.spawn_k = spawn_k; p.spawn_exp = spawn_exp;
    p.out = out; p.reward = reward; p.done = done; p.status = status;
    p.tables = d->tables();
    return launch_step<true>(p, n, 0ull, 0ull, 0u, reward_mode, *d, (cudaStream_t)stream);
}

int r48_env_step_ring(uint64_t *boards, const uint8_t *action, uint32_t *steps, uint32_t *episodes,
                      int32_t *reward, uint8_t *done, float *obs, int obs_mode, uint64_t *final_boards,
                      int64_t n, uint64_t seed, uint64_t board_base, uint64_t id_stride,
                      int reward_mode, int auto_reset, int32_t *status, const r48_ring *ring, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (reward_mode != 0 && reward_mode != 1) return fail(R48_ERR_ARG, "r48_env_step: reward_mode must be 0 or 1");
    if (obs_mode != 0 && obs_mode != 1) return fail(R48_ERR_ARG, "r48_env_step: obs_mode must be 0 or 1");
    if (n == 0) return R48_OK;
    if (!boards || !action || !steps || !episodes) return fail(R48_ERR_NULL, "r48_env_step: boards/action/steps/episodes is NULL");
    if ((uint64_t)n > id_stride) return fail(R48_ERR_ARG, "r48_env_step: id_stride must be >= n");
    if (!aligned(boards, 8) || !aligned(steps, 4) || !aligned(episodes, 4) || (reward && !aligned(reward, 4)) ||
        (obs && !aligned(obs, 16)) || (final_boards && !aligned(final_boards, 8)) || (status && !aligned(status, 4)))
        return fail(R48_ERR_ALIGN, "r48_env_step: misaligned pointer");
    RingRefs rr{};
    if (ring) {
        if ((rc = ring_refs(ring, &rr, "r48_env_step_ring: NULL ring array"))) return rc;
        if (n > kChunk) return fail(R48_ERR_ARG, "r48_env_step_ring: at most 2^30 envs per call with a ring");
    }
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    for (int64_t off = 0; off < n; off += kChunk) {
        const int64_t m = n - off < kChunk ? n - off : kChunk;
        EnvParams p{boards + off, action + off, steps + off, episodes + off, reward ? reward + off : nullptr,
                    done ? done + off : nullptr, obs ? obs + 16 * off : nullptr,
                    final_boards ? final_boards + off : nullptr, status, (uint32_t)m,
                    board_base + (uint64_t)off, id_stride, obs_mode, auto_reset, make_keys(seed), d->tables(), rr};
        const int grid = grid_for(m, kThreads, d->sms, 1);
        if (reward_mode)
            CK(launch_pdl(env_step_kernel<true>, grid, kThreads, kLeftBytes + kMergeBytes, (cudaStream_t)stream, p));
        else
            CK(launch_pdl(env_step_kernel<false>, grid, kThreads, kLrBytes, (cudaStream_t)stream, p));
    }
    return R48_OK;
}

int r48_env_step(uint64_t *boards, const uint8_t *action, uint32_t *steps, uint32_t *episodes,
                 int32_t *reward, uint8_t *done, float *obs, int obs_mode, uint64_t *final_boards,
                 int64_t n, uint64_t seed, uint64_t board_base, uint64_t id_stride,
                 int reward_mode, int auto_reset, int32_t *status, void *stream)
{
    return r48_env_step_ring(boards, action, steps, episodes, reward, done, obs, obs_mode, final_boards, n, seed,
                             board_base, id_stride, reward_mode, auto_reset, status, nullptr, stream);
}

int r48_spawn_injected(uint64_t *boards, const uint8_t *spawn_k, const uint8_t *spawn_exp,
                       int64_t n, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!boards || !spawn_k || !spawn_exp) return fail(R48_ERR_NULL, "r48_spawn_injected: NULL pointer");
    if (!aligned(boards, 8)) return fail(R48_ERR_ALIGN, "r48_spawn_injected: boards not 8-byte aligned");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    spawn_injected_kernel<<<grid_for(n, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(boards, spawn_k, spawn_exp, n);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_step_injected_view(uint64_t *boards, const uint8_t *action, const uint8_t *spawn_k,
                           const uint8_t *spawn_exp, int64_t n, int reward_mode,
                           struct r48_game_view *views, int32_t *status, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!boards || !action || !spawn_k || !spawn_exp || !views)
        return fail(R48_ERR_NULL, "r48_step_injected_view: NULL pointer");
    if (!aligned(boards, 8) || !aligned(views, 4))
        return fail(R48_ERR_ALIGN, "r48_step_injected_view: boards must be 8-byte, views 4-byte aligned");
    if (reward_mode != 0 && reward_mode != 1) return fail(R48_ERR_ARG, "r48_step_injected_view: reward_mode must be 0 or 1");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    game_view_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        boards, action, spawn_k, spawn_exp, n, reward_mode, views, status, d->tables());
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_spawn(uint64_t *boards, int64_t n, uint64_t seed, uint64_t board_base, uint32_t tick,
              void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!boards) return fail(R48_ERR_NULL, "r48_spawn: boards is NULL");
    if (!aligned(boards, 8)) return fail(R48_ERR_ALIGN, "r48_spawn: boards not 8-byte aligned");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    ResetParams p{boards, n, board_base, make_keys(seed)};
    spawn_kernel<<<grid_for(n, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(p, tick);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_blank_counts(const uint64_t *boards, uint8_t *counts, int64_t n, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!boards || !counts) return fail(R48_ERR_NULL, "r48_blank_counts: NULL pointer");
    if (!aligned(boards, 8)) return fail(R48_ERR_ALIGN, "r48_blank_counts: boards not 8-byte aligned");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    blank_counts_kernel<<<grid_for(n, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(boards, counts, n);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_afterstates(const uint64_t *in, uint64_t *out, int32_t *reward, uint8_t *valid,
                    uint8_t *done, int64_t n, int reward_mode, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (reward_mode != 0 && reward_mode != 1) return fail(R48_ERR_ARG, "r48_afterstates: reward_mode must be 0 or 1");
    if (n == 0) return R48_OK;
    if (!in || !out) return fail(R48_ERR_NULL, "r48_afterstates: in/out is NULL");
    if (!aligned(in, 8) || !aligned(out, 8) || (reward && !aligned(reward, 4)))
        return fail(R48_ERR_ALIGN, "r48_afterstates: misaligned pointer");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    for (int64_t off = 0; off < n; off += kChunk) {
        const int64_t m = n - off < kChunk ? n - off : kChunk;
        AfterParams p{in + off, out + off, reward ? reward + off : nullptr, valid ? valid + off : nullptr,
                      done ? done + off : nullptr, m, (uint64_t)n, d->tables()};
        const int grid = grid_for(m, kThreads, d->sms, 1);
        if (reward_mode)
            CK(launch_pdl(afterstates_kernel<true>, grid, kThreads, kLeftBytes + kMergeBytes, (cudaStream_t)stream, p));
        else
            CK(launch_pdl(afterstates_kernel<false>, grid, kThreads, kLrBytes, (cudaStream_t)stream, p));
    }
    return R48_OK;
}

int r48_episode_stats(const uint64_t *final_boards, const uint32_t *lengths, int64_t n,
                      uint64_t *stats, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!final_boards || !lengths || !stats) return fail(R48_ERR_NULL, "r48_episode_stats: NULL pointer");
    if (!aligned(final_boards, 8) || !aligned(lengths, 4) || !aligned(stats, 8))
        return fail(R48_ERR_ALIGN, "r48_episode_stats: misaligned pointer");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    stats_kernel<<<grid_for(n, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        final_boards, lengths, n, (unsigned long long *)stats, nullptr);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_episode_records(const uint64_t *final_boards, const uint32_t *lengths, uint32_t *records,
                        int64_t n, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!final_boards || !lengths || !records) return fail(R48_ERR_NULL, "r48_episode_records: NULL pointer");
    if (!aligned(final_boards, 8) || !aligned(lengths, 4) || !aligned(records, 4))
        return fail(R48_ERR_ALIGN, "r48_episode_records: misaligned pointer");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    records_kernel<<<grid_for(n, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(final_boards, lengths, records, n);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_rollout_policy(int64_t n, uint64_t seed, uint64_t board_base, int policy,
                       uint64_t *final_boards, uint32_t *lengths, uint64_t *stats, void *workspace,
                       void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (policy != R48_POLICY_RANDOM && policy != R48_POLICY_GREEDY_BLANKS)
        return fail(R48_ERR_ARG, "r48_rollout_policy: unknown policy");
    if (n == 0) return R48_OK;
    if (!final_boards || !lengths || !workspace) return fail(R48_ERR_NULL, "r48_rollout: NULL pointer");
    if (!aligned(final_boards, 8) || !aligned(lengths, 4) || !aligned(workspace, 8) ||
        (stats && !aligned(stats, 8)))
        return fail(R48_ERR_ALIGN, "r48_rollout: misaligned pointer");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    constexpr int64_t kRolloutChunk = (int64_t)1 << 30;      // the kernel's episode queue is 32-bit
    for (int64_t off = 0; off < n; off += kRolloutChunk) {
        const int64_t m = n - off < kRolloutChunk ? n - off : kRolloutChunk;
        CK(cudaMemsetAsync(workspace, 0, R48_ROLLOUT_WORKSPACE_BYTES, s));
        RolloutParams p{final_boards + off, lengths + off, (unsigned int *)workspace, (uint32_t)m,
                        board_base + (uint64_t)off, make_keys(seed), d->tables(), nullptr, nullptr, nullptr};
        const int grid = grid_for(m, kThreads, d->sms, 1);
        if (policy == R48_POLICY_RANDOM)
            CK(launch_pdl(rollout_kernel<kPolicyRandom, false>, grid, kThreads, kLrBytes, s, p));
        else
            CK(launch_pdl(rollout_kernel<kPolicyGreedyBlanks, false>, grid, kThreads, kLrBytes, s, p));
    }
    if (stats) return r48_episode_stats(final_boards, lengths, n, stats, stream);
    return R48_OK;
}

int r48_rollout_trajectories(int64_t n, uint64_t seed, uint64_t board_base, int policy,
                             const uint32_t *lengths, const uint64_t *offsets, uint64_t *traj_boards,
                             uint8_t *traj_actions, void *workspace, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (policy != R48_POLICY_RANDOM && policy != R48_POLICY_GREEDY_BLANKS)
        return fail(R48_ERR_ARG, "r48_rollout_trajectories: unknown policy");
    if (n == 0) return R48_OK;
    if (!lengths || !offsets || !traj_boards || !traj_actions || !workspace)
        return fail(R48_ERR_NULL, "r48_rollout_trajectories: NULL pointer");
    if (!aligned(lengths, 4) || !aligned(offsets, 8) || !aligned(traj_boards, 32) || !aligned(traj_actions, 4) ||
        !aligned(workspace, 8))
        return fail(R48_ERR_ALIGN, "r48_rollout_trajectories: traj_boards needs 32-byte, traj_actions 4-byte alignment");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    constexpr int64_t kRolloutChunk = (int64_t)1 << 30;
    for (int64_t off = 0; off < n; off += kRolloutChunk) {
        const int64_t m = n - off < kRolloutChunk ? n - off : kRolloutChunk;
        CK(cudaMemsetAsync(workspace, 0, R48_ROLLOUT_WORKSPACE_BYTES, s));
        RolloutParams p{nullptr, const_cast<uint32_t *>(lengths) + off, (unsigned int *)workspace, (uint32_t)m,
                        board_base + (uint64_t)off, make_keys(seed), d->tables(), offsets + off, traj_boards,
                        traj_actions};
        const int grid = grid_for(m, kThreads, d->sms, 1);
        if (policy == R48_POLICY_RANDOM)
            CK(launch_pdl(rollout_kernel<kPolicyRandom, true>, grid, kThreads, kLrBytes, s, p));
        else
            CK(launch_pdl(rollout_kernel<kPolicyGreedyBlanks, true>, grid, kThreads, kLrBytes, s, p));
    }
    return R48_OK;
}

int r48_rollout(int64_t n, uint64_t seed, uint64_t board_base, uint64_t *final_boards,
                uint32_t *lengths, uint64_t *stats, void *workspace, void *stream)
{
    return r48_rollout_policy(n, seed, board_base, R48_POLICY_RANDOM, final_boards, lengths, stats,
                              workspace, stream);
}

int r48_scores(const uint64_t *boards, uint32_t *score, uint8_t *max_exp, int64_t n, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!boards) return fail(R48_ERR_NULL, "r48_scores: boards is NULL");
    if (!aligned(boards, 8) || (score && !aligned(score, 4)))
        return fail(R48_ERR_ALIGN, "r48_scores: misaligned pointer");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    scores_kernel<<<grid_for(n, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(boards, score, max_exp, n);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_decode_f32(const uint64_t *boards, float *out, int64_t n, int log2_planes, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!boards || !out) return fail(R48_ERR_NULL, "r48_decode_f32: NULL pointer");
    if (!aligned(boards, 8) || !aligned(out, 16))
        return fail(R48_ERR_ALIGN, "r48_decode_f32: boards needs 8-byte, out 16-byte alignment");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    const int grid = grid_for(n * 4, 256, d->sms, 8);
    if (log2_planes) decode_kernel<float, true><<<grid, 256, 0, (cudaStream_t)stream>>>(boards, out, n);
    else decode_kernel<float, false><<<grid, 256, 0, (cudaStream_t)stream>>>(boards, out, n);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_decode_i32(const uint64_t *boards, int32_t *out, int64_t n, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!boards || !out) return fail(R48_ERR_NULL, "r48_decode_i32: NULL pointer");
    if (!aligned(boards, 8) || !aligned(out, 16))
        return fail(R48_ERR_ALIGN, "r48_decode_i32: boards needs 8-byte, out 16-byte alignment");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    decode_kernel<int32_t, false><<<grid_for(n * 4, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(boards, out, n);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_encode_i32(const int32_t *values, uint64_t *boards, int64_t n, int32_t *status, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!values || !boards) return fail(R48_ERR_NULL, "r48_encode_i32: NULL pointer");
    if (!aligned(values, 16) || !aligned(boards, 8) || (status && !aligned(status, 4)))
        return fail(R48_ERR_ALIGN, "r48_encode_i32: values needs 16-byte, boards 8-byte alignment");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    encode_kernel<<<grid_for(n, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(values, boards, n, status);
    CK(cudaGetLastError());
    return R48_OK;
}

// ---------------------------------------------------------------- transition ring

int r48_ring_clear(const r48_ring *ring, void *stream)
{
    if (!ring || !ring->cursor) return fail(R48_ERR_NULL, "r48_ring_clear: NULL ring");
    if (!aligned(ring->cursor, 8)) return fail(R48_ERR_ALIGN, "r48_ring_clear: cursor not 8-byte aligned");
    ring_clear_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ring->cursor);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_ring_append(const r48_ring *ring, const uint64_t *state, const uint8_t *action,
                    const int32_t *reward, const uint64_t *next_state, const uint8_t *done, int64_t n,
                    int drop_when_full, void *stream)
{
    int rc = check_n(n);
    if (rc) return rc;
    RingRefs rr;
    if ((rc = ring_refs(ring, &rr, "r48_ring_append: NULL ring array"))) return rc;
    if (n == 0) return R48_OK;
    if (!state || !action || !next_state) return fail(R48_ERR_NULL, "r48_ring_append: state/action/next_state is NULL");
    if (!aligned(state, 8) || !aligned(next_state, 8) || (reward && !aligned(reward, 4)))
        return fail(R48_ERR_ALIGN, "r48_ring_append: misaligned pointer");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    RingAppendParams p{rr, state, action, reward, next_state, done, (uint64_t)n, drop_when_full != 0};
    ring_append_kernel<<<grid_for(n, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(p);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_ring_sample(const r48_ring *ring, int64_t batch, uint64_t seed, uint64_t draw, int with_replacement,
                    int64_t *out_index, uint64_t *out_state, uint8_t *out_action, int32_t *out_reward,
                    uint64_t *out_next_state, uint8_t *out_done, float *obs_state, float *obs_next_state,
                    int obs_mode, void *stream)
{
    int rc = check_n(batch);
    if (rc) return rc;
    RingRefs rr;
    if ((rc = ring_refs(ring, &rr, "r48_ring_sample: NULL ring array"))) return rc;
    if (obs_mode != 0 && obs_mode != 1) return fail(R48_ERR_ARG, "r48_ring_sample: obs_mode must be 0 or 1");
    if (batch == 0) return R48_OK;
    if (!out_state || !out_action || !out_reward || !out_next_state || !out_done)
        return fail(R48_ERR_NULL, "r48_ring_sample: NULL output");
    if (!aligned(out_state, 8) || !aligned(out_next_state, 8) || !aligned(out_reward, 4) ||
        (out_index && !aligned(out_index, 8)) || (obs_state && !aligned(obs_state, 16)) ||
        (obs_next_state && !aligned(obs_next_state, 16)))
        return fail(R48_ERR_ALIGN, "r48_ring_sample: misaligned output");
    DeviceState *d;
    if ((rc = current_device(&d))) return rc;
    RingSampleParams p{rr, (uint64_t)batch, seed, draw, with_replacement != 0, make_keys(seed), out_index, out_state,
                       out_action, out_reward, out_next_state, out_done, obs_state, obs_next_state, obs_mode};
    ring_sample_kernel<<<grid_for(batch, 256, d->sms, 8), 256, 0, (cudaStream_t)stream>>>(p);
    CK(cudaGetLastError());
    return R48_OK;
}

int r48_debug_copy22(const uint64_t *in, const uint8_t *action, uint64_t *out, int32_t *reward, uint8_t *done,
                     int64_t n, void *stream)
{
    if (n <= 0 || n > kChunk || (n & 1)) return fail(R48_ERR_ARG, "r48_debug_copy22: n must be even, 2 .. 2^30");
    if (!in || !action || !out || !reward || !done) return fail(R48_ERR_NULL, "r48_debug_copy22: NULL pointer");
    if (!aligned(in, 16) || !aligned(out, 16) || !aligned(reward, 8) || !aligned(action, 2) || !aligned(done, 2))
        return fail(R48_ERR_ALIGN, "r48_debug_copy22: misaligned pointer");
    DeviceState *d;
    int rc;
    if ((rc = current_device(&d))) return rc;
    copy22_kernel<<<grid_for(n / 2, kThreads, d->sms, 2), kThreads, 0, (cudaStream_t)stream>>>(in, action, out, reward, done, (uint32_t)n);
    CK(cudaGetLastError());
    return R48_OK;
}

// ---------------------------------------------------------------- host-buffer entry points
// Each call holds its device's host mutex from the first arena access to the last
// synchronise: calls from several threads on one device serialise, calls on different devices
// run concurrently.

int r48_step_host(const uint64_t *in, const uint8_t *action, uint64_t *out, int32_t *reward,
                  uint8_t *done, int64_t n, uint64_t seed, uint64_t board_base, uint32_t step,
                  int reward_mode, int device)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!in || !action || !out) return fail(R48_ERR_NULL, "r48_step_host: in/action/out is NULL");
    DeviceState *d;
    if ((rc = ensure_device(device, &d))) return rc;
    std::lock_guard<std::mutex> host_lock(g_host_mu[device]);
    DeviceGuard g(device);
    const size_t nb = (size_t)n;
    const size_t o_board = 0, o_act = up256(nb * 8), o_rew = o_act + up256(nb),
                 o_done = o_rew + up256(nb * 4), o_status = o_done + up256(nb), total = o_status + 256;
    if ((rc = arena_reserve(*d, total))) return rc;
    uint64_t *d_board = (uint64_t *)(d->arena + o_board);
    uint8_t *d_act = d->arena + o_act;
    int32_t *d_rew = (int32_t *)(d->arena + o_rew);
    uint8_t *d_done = d->arena + o_done;
    int32_t *d_status = (int32_t *)(d->arena + o_status);
    cudaStream_t s = d->stream, c = d->copy_stream;
    CK(cudaMemsetAsync(d_status, 0, 4, s));
    // 3-stage pipeline over chunks of 2^18 boards: H2D + kernel of chunk i+1 on the compute stream
    // overlap the D2H of chunk i on the copy stream (PCIe is full duplex)
    const int64_t chunk = n > ((int64_t)1 << 19) ? ((int64_t)1 << 18) : n;
    int slot = 0;
    for (int64_t off = 0; off < n; off += chunk, slot ^= 1) {
        const size_t m = (size_t)(n - off < chunk ? n - off : chunk);
        CK(cudaMemcpyAsync(d_board + off, in + off, m * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_act + off, action + off, m, cudaMemcpyHostToDevice, s));
        rc = r48_step(d_board + off, d_act + off, d_board + off, reward ? d_rew + off : nullptr,
                      done ? d_done + off : nullptr, (int64_t)m, seed, board_base + (uint64_t)off, step,
                      reward_mode, d_status, s);
        if (rc) return rc;
        CK(cudaEventRecord(d->chunk_done[slot], s));
        CK(cudaStreamWaitEvent(c, d->chunk_done[slot], 0));
        CK(cudaMemcpyAsync(out + off, d_board + off, m * 8, cudaMemcpyDeviceToHost, c));
        if (reward) CK(cudaMemcpyAsync(reward + off, d_rew + off, m * 4, cudaMemcpyDeviceToHost, c));
        if (done) CK(cudaMemcpyAsync(done + off, d_done + off, m, cudaMemcpyDeviceToHost, c));
    }
    int32_t h_status = 0;
    CK(cudaMemcpyAsync(&h_status, d_status, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(c));
    if (h_status & 1) return fail(R48_ERR_ACTION, "r48_step_host: action byte > 3");
    return R48_OK;
}

int r48_afterstates_host(const uint64_t *in, uint64_t *out, int32_t *reward, uint8_t *valid,
                         uint8_t *done, int64_t n, int reward_mode, int device)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (n == 0) return R48_OK;
    if (!in || !out) return fail(R48_ERR_NULL, "r48_afterstates_host: in/out is NULL");
    DeviceState *d;
    if ((rc = ensure_device(device, &d))) return rc;
    std::lock_guard<std::mutex> host_lock(g_host_mu[device]);
    DeviceGuard g(device);
    const size_t nb = (size_t)n;
    const size_t o_in = 0, o_out = up256(nb * 8), o_rew = o_out + up256(nb * 32),
                 o_valid = o_rew + up256(nb * 16), o_done = o_valid + up256(nb), total = o_done + up256(nb);
    if ((rc = arena_reserve(*d, total))) return rc;
    uint64_t *d_in = (uint64_t *)(d->arena + o_in), *d_out = (uint64_t *)(d->arena + o_out);
    int32_t *d_rew = (int32_t *)(d->arena + o_rew);
    uint8_t *d_valid = d->arena + o_valid, *d_done = d->arena + o_done;
    cudaStream_t s = d->stream;
    CK(cudaMemcpyAsync(d_in, in, nb * 8, cudaMemcpyHostToDevice, s));
    rc = r48_afterstates(d_in, d_out, reward ? d_rew : nullptr, valid ? d_valid : nullptr,
                         done ? d_done : nullptr, n, reward_mode, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, d_out, nb * 32, cudaMemcpyDeviceToHost, s));
    if (reward) CK(cudaMemcpyAsync(reward, d_rew, nb * 16, cudaMemcpyDeviceToHost, s));
    if (valid) CK(cudaMemcpyAsync(valid, d_valid, nb, cudaMemcpyDeviceToHost, s));
    if (done) CK(cudaMemcpyAsync(done, d_done, nb, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return R48_OK;
}

int r48_rollout_host_ex(int64_t n, uint64_t seed, uint64_t board_base, int policy, uint64_t *final_boards,
                        uint32_t *lengths, uint32_t *records, uint64_t *stats, int device)
{
    int rc = check_n(n);
    if (rc) return rc;
    if (policy != R48_POLICY_RANDOM && policy != R48_POLICY_GREEDY_BLANKS)
        return fail(R48_ERR_ARG, "r48_rollout_host_ex: unknown policy");
    if (n == 0) return R48_OK;
    DeviceState *d;
    if ((rc = ensure_device(device, &d))) return rc;
    std::lock_guard<std::mutex> host_lock(g_host_mu[device]);
    DeviceGuard g(device);
    const size_t nb = (size_t)n;
    const size_t o_fb = 0, o_len = up256(nb * 8), o_rec = o_len + up256(nb * 4),
                 o_stats = o_rec + up256(records ? nb * 4 : 0), o_ws = o_stats + up256(R48_STATS_WORDS * 8),
                 total = o_ws + 2 * R48_ROLLOUT_WORKSPACE_BYTES;
    if ((rc = arena_reserve(*d, total))) return rc;
    uint64_t *d_fb = (uint64_t *)(d->arena + o_fb);
    uint32_t *d_len = (uint32_t *)(d->arena + o_len);
    uint32_t *d_rec = (uint32_t *)(d->arena + o_rec);
    uint64_t *d_stats = (uint64_t *)(d->arena + o_stats);
    cudaStream_t cs[2] = {d->stream, d->stream2}, c = d->copy_stream;
    if (stats) CK(cudaMemsetAsync(d_stats, 0, R48_STATS_WORDS * 8, cs[0]));
    CK(cudaEventRecord(d->join, cs[0]));
    CK(cudaStreamWaitEvent(cs[1], d->join, 0));
    // Large batches go in chunks of 2^24 episodes: the per-episode results of chunk i travel to
    // the host on the copy stream while chunk i+1 is being played (67 MB of records per chunk: 1.2 ms
    // of PCIe under 8 ms of play).
    // The copy of the LAST chunk is the only one left exposed, so the chunks shrink toward the end:
    // each half of what is left (..., 2^24, 2^23, ..., 2^20, 2^20) when a chunk's copy takes less than
    // 0.3 of its play -- one GPU on the host: 0.07 ns per record against 0.5 ns of play per
    // episode -- and a quarter of what is left (16M, 12M, 9M, 7M, ...) when it does not: eight ranks
    // sharing the host's memory bandwidth get ~11 GB/s each, 0.36 ns per record, and halved chunks
    // pile their copies up at the end.  The ratio is the one the previous call measured on its first
    // chunk (events around the kernel and around the copy); every extra chunk costs ~0.1 ms.
    // Chunks ALTERNATE between two streams: a rollout launch ends with ~0.1 ms in which its last long
    // games keep a shrinking number of SMs busy; the next chunk's kernel, already queued on the other
    // stream, takes every SM as it is vacated (a kernel behind it on the SAME stream would wait for
    // the last CTA).  Each stream has its own episode-queue workspace; the statistics kernels add
    // into one vector with atomics.
    const int64_t big = (int64_t)1 << 24, small = (int64_t)1 << 20;
    int64_t shrink = d->copy_over_play > 0.3f ? 4 : 2;
    if (const char *forced = getenv("R48_HOST_CHUNK_SHRINK")) {          // tests: take the other schedule
        if (forced[0] == '2' || forced[0] == '4') shrink = forced[0] - '0';
    }
    bool probed = false;
    int slot = 0;
    for (int64_t off = 0, m = 0; off < n; off += m, slot ^= 1) {
        const int64_t left = n - off;
        if (n <= ((int64_t)1 << 23) || !(final_boards || lengths || records))
            m = left;                                          // small batches, or nothing to copy per episode: one launch
        else if (left > shrink * big) m = big;
        else if (left > 2 * small) m = (left / shrink + small - 1) / small * small;   // in 2^20 units
        else m = left;
        cudaStream_t s = cs[slot];
        const bool probing = off == 0 && m < n && (final_boards || lengths || records);
        if (probing) CK(cudaEventRecord(d->probe[0], s));
        rc = r48_rollout_policy(m, seed, board_base + (uint64_t)off, policy, d_fb + off, d_len + off, nullptr,
                                d->arena + o_ws + (size_t)slot * R48_ROLLOUT_WORKSPACE_BYTES, s);
        if (rc) return rc;
        if (probing) { CK(cudaEventRecord(d->probe[1], s)); CK(cudaStreamWaitEvent(c, d->probe[1], 0)); CK(cudaEventRecord(d->probe[2], c)); }
        // what the rollout kernel itself wrote can leave as soon as it has finished; the statistics
        // pass of this chunk gets its SMs only when the NEXT chunk's kernel starts to drain
        if (final_boards || lengths) {
            CK(cudaEventRecord(d->chunk_done[slot], s));
            CK(cudaStreamWaitEvent(c, d->chunk_done[slot], 0));
            if (final_boards)
                CK(cudaMemcpyAsync(final_boards + off, d_fb + off, (size_t)m * 8, cudaMemcpyDeviceToHost, c));
            if (lengths)
                CK(cudaMemcpyAsync(lengths + off, d_len + off, (size_t)m * 4, cudaMemcpyDeviceToHost, c));
            if (probing) CK(cudaEventRecord(d->probe[3], c));
        }
        if (stats) {                                           // statistics and records in one pass over the chunk
            stats_kernel<<<grid_for(m, 256, d->sms, 8), 256, 0, s>>>(d_fb + off, d_len + off, m,
                                                                     (unsigned long long *)d_stats, records ? d_rec + off : nullptr);
            CK(cudaGetLastError());
        } else if (records && (rc = r48_episode_records(d_fb + off, d_len + off, d_rec + off, m, s))) {
            return rc;
        }
        if (records) {
            CK(cudaEventRecord(d->stats_done[slot], s));
            CK(cudaStreamWaitEvent(c, d->stats_done[slot], 0));
            const bool probe_here = probing && !(final_boards || lengths);
            if (probe_here) CK(cudaEventRecord(d->probe[2], c));      // the copy alone, not the wait before it
            CK(cudaMemcpyAsync(records + off, d_rec + off, (size_t)m * 4, cudaMemcpyDeviceToHost, c));
            if (probe_here) CK(cudaEventRecord(d->probe[3], c));
        }
        probed = probed || probing;
    }
    CK(cudaEventRecord(d->join, cs[1]));                       // the statistics are complete when both streams are
    CK(cudaStreamWaitEvent(cs[0], d->join, 0));
    if (stats) CK(cudaMemcpyAsync(stats, d_stats, R48_STATS_WORDS * 8, cudaMemcpyDeviceToHost, cs[0]));
    CK(cudaStreamSynchronize(cs[0]));
    CK(cudaStreamSynchronize(cs[1]));
    CK(cudaStreamSynchronize(c));
    if (probed) {
        float play_ms = 0.0f, copy_ms = 0.0f;
        if (cudaEventElapsedTime(&play_ms, d->probe[0], d->probe[1]) == cudaSuccess &&
            cudaEventElapsedTime(&copy_ms, d->probe[2], d->probe[3]) == cudaSuccess && play_ms > 0.0f)
            d->copy_over_play = copy_ms / play_ms;
    }
    return R48_OK;
}

int r48_rollout_host(int64_t n, uint64_t seed, uint64_t board_base, uint64_t *final_boards,
                     uint32_t *lengths, uint64_t *stats, int device)
{
    return r48_rollout_host_ex(n, seed, board_base, R48_POLICY_RANDOM, final_boards, lengths, nullptr, stats, device);
}

int r48_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mu);
    for (int dev = 0; dev < kMaxDevices; dev++) {
        std::lock_guard<std::mutex> host_lock(g_host_mu[dev]);
        DeviceState &d = g_dev[dev];
        if (!d.ready && !d.arena && !d.stream) continue;
        DeviceGuard g(dev);
        if (d.arena) cudaFree(d.arena);
        if (d.stream) cudaStreamDestroy(d.stream);
        if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
        if (d.stream2) cudaStreamDestroy(d.stream2);
        if (d.join) cudaEventDestroy(d.join);
        for (int e = 0; e < 2; e++) if (d.stats_done[e]) cudaEventDestroy(d.stats_done[e]);
        for (int e = 0; e < 4; e++) if (d.probe[e]) cudaEventDestroy(d.probe[e]);
        for (int e = 0; e < 2; e++) if (d.chunk_done[e]) cudaEventDestroy(d.chunk_done[e]);
        if (d.left) cudaFree(d.left);
        if (d.merges) cudaFree(d.merges);
        if (d.lr) cudaFree(d.lr);
        d = DeviceState();
    }
    return R48_OK;
}

}  // extern "C"
