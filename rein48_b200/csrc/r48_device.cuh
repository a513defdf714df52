// r48_device.cuh -- device-side building blocks of the batched 2048 environment (sm_100a).
//
// A board lives in two 32-bit registers: `lo` = rows 0,1 and `hi` = rows 2,3, four
// exponent nibbles per row (cell (i,j) = nibble 4*i+j of the 64-bit word).  Everything
// here is integer-pipe work; there is no tensor-core or floating-point math on this path.
//
// Reference behaviour restated by these functions (file:line in nevertiree/Rein48):
//   rows_lr / rows_l16 / move_all / transpose   Game.update_matrix      game/GameClient.py:129-254
//   count_blanks / kth_blank / spawn_tile       Game.random_fill_grid   game/GameClient.py:102-127
//   game_over / no_equal_neighbours             Game.has_game_over      game/GameClient.py:65-94
//   philox4x32 / draw_word                      random.randint/uniform  control/rand.py:11, GameClient.py:121,125
#pragma once
#include <stdint.h>

// Build-time switches (A/B-timed on a B200, see DESIGN.md "measured and rejected"):
//   R48_FMA_INDEX  table addresses by IMAD.WIDE / IMAD.HI (FMA pipe) instead of SHF + LOP3 (ALU pipe)
//   R48_SWIZZLE    XOR bits 7..11 of the row into the bank bits of its table slot
//   R48_PRED_TRANSPOSE (default 1)  conditional transposes without a branch (transpose_where)
// and in r48_kernels.cu: R48_STEP_UNROLL / _PREFETCH / _GATE_EARLY / _TABLE_GLOBAL, R48_AFTER_PIPE /
// _PREFETCH, R48_CALLS_PER_ITER, R48_UNGUARDED_TICKS, R48_RING_L2_HINT.
#ifndef R48_FMA_INDEX
#define R48_FMA_INDEX 0
#endif
#ifndef R48_SWIZZLE
#define R48_SWIZZLE 0
#endif


namespace r48 {

// ------------------------------------------------------------------ small helpers

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    return __byte_perm(a, b, sel);
}

// (a & m) | (b & ~m) as exactly one LOP3 (ptxas otherwise re-derives it from known-zero
// bits of the shifted operands and spends an extra instruction per select)
__device__ __forceinline__ uint32_t bsel(uint32_t m, uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(d) : "r"(m), "r"(a), "r"(b));
    return d;
}

// shared-window address of a shared-memory object, and table reads through it: the base is
// computed once per kernel instead of once per lookup group, and every read is LDS [R + imm]
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// the same, made opaque so that ptxas keeps it in a register instead of re-deriving the
// window base (S2UR CgaCtaId / UMOV / ULEA) in front of every group of lookups
__device__ __forceinline__ uint32_t smem_u32_pinned(const void *p)
{
    uint32_t a = smem_u32(p), r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}

// These are volatile asm statements WITHOUT a "memory" clobber on purpose: volatile keeps them in
// program order among themselves (load, ..., store of one trip), while a clobber would make the
// compiler re-read every kernel parameter after each of them.
// OFF: a compile-time byte offset that travels in the instruction's immediate field.
template <uint32_t OFF = 0u>
__device__ __forceinline__ ulonglong2 ldg_u64x2(uint64_t addr)
{
    ulonglong2 v;
    asm volatile("ld.global.v2.u64 {%0, %1}, [%2+%3];" : "=l"(v.x), "=l"(v.y) : "l"(addr), "n"(OFF));
    return v;
}

template <uint32_t OFF = 0u>
__device__ __forceinline__ uint32_t ldg_u16(uint64_t addr)
{
    uint16_t v;
    asm volatile("ld.global.u16 %0, [%1+%2];" : "=h"(v) : "l"(addr), "n"(OFF));
    return v;
}

template <uint32_t OFF = 0u>
__device__ __forceinline__ void stg_u64x2(uint64_t addr, uint64_t x, uint64_t y)
{
    asm volatile("st.global.v2.u64 [%0+%3], {%1, %2};" ::"l"(addr), "l"(x), "l"(y), "n"(OFF));
}

template <uint32_t OFF = 0u>
__device__ __forceinline__ void stg_u32x2(uint64_t addr, uint32_t x, uint32_t y)
{
    asm volatile("st.global.v2.u32 [%0+%3], {%1, %2};" ::"l"(addr), "r"(x), "r"(y), "n"(OFF));
}

__device__ __forceinline__ void stg_u64(uint64_t addr, uint64_t v)
{
    asm volatile("st.global.u64 [%0], %1;" ::"l"(addr), "l"(v));
}

__device__ __forceinline__ void stg_u32(uint64_t addr, uint32_t v)
{
    asm volatile("st.global.u32 [%0], %1;" ::"l"(addr), "r"(v));
}

template <uint32_t OFF = 0u>
__device__ __forceinline__ void stg_u16(uint64_t addr, uint32_t v)
{
    asm volatile("st.global.u16 [%0+%2], %1;" ::"l"(addr), "h"((uint16_t)v), "n"(OFF));
}

// volatile: must not move above the mbarrier wait that publishes the staged tables
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t lds_u32_at(const uint32_t *p)
{
    return lds_u32(smem_u32(p));
}

// Launch constants for a hot loop, pinned in registers.  ptxas re-reads kernel parameters from
// the constant bank (LDC / LDCU) wherever it uses them -- ~9 issue slots per board in the step
// kernel, which runs at the issue ceiling -- and it sees through PTX moves and self-shuffles.  What
// it cannot fold is a value that came through shared memory: thread 0 parks the words of the
// parameters there, and after the CTA barrier every thread reads the words back once.
template <int N>
struct PinnedWords {
    uint32_t w[N];
    // after the CTA barrier that follows the parking stores (sts_u32_at by one thread)
    __device__ __forceinline__ void fetch(const uint32_t *slots)
    {
#pragma unroll
        for (int t = 0; t < N; t++) w[t] = lds_u32_at(slots + t);
    }
    __device__ __forceinline__ uint64_t u64(int t) const { return ((uint64_t)w[t + 1] << 32) | w[t]; }
};

__device__ __forceinline__ void sts_u32_at(uint32_t *p, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v));
}

__device__ __forceinline__ void sts_u64_at(uint32_t *p, uint64_t v)
{
    sts_u32_at(p, (uint32_t)v);
    sts_u32_at(p + 1, (uint32_t)(v >> 32));
}

// ------------------------------------------------------------------ Philox4x32-7
// Salmon et al., SC'11 (Random123 philox4x32, 7 rounds: the smallest round count the paper reports
// as passing BigCrush; KATs in tests/test_oracle_golden.py).  The round keys depend only on the
// seed, so the host precomputes them and they arrive as kernel parameters (constant bank operands).
constexpr int kPhiloxRounds = 7;

struct PhiloxKeys {
    uint32_t k0[kPhiloxRounds];
    uint32_t k1[kPhiloxRounds];
};

#define R48_PHILOX_M0 0xD2511F53u
#define R48_PHILOX_M1 0xCD9E8D57u
#define R48_PHILOX_W0 0x9E3779B9u
#define R48_PHILOX_W1 0xBB67AE85u
#define R48_SPAWN4_THRESHOLD 0x1999999Au      // ceil(0.1 * 2^32): P(tile 4) = 0.1
#define R48_VALUE_HASH 0x9E3779B1u            // odd: a -> a * HASH mod 2^32 is a bijection

__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           const PhiloxKeys &K, uint32_t (&w)[4])
{
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; r++) {
        uint64_t p0 = (uint64_t)R48_PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)R48_PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r];
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}

// Draw spec: tick t of board `id` uses call (id.lo, id.hi, t >> 2, 0) and
// its word a = w[t & 3]:
//   action = a >> 30
//   cell   = the k-th blank, k = mulhi(a << 2, n_blank), counted in the order of the tick's move
//            axis (row-major for LEFT/RIGHT, column-major for UP/DOWN)
//   value  = 4 if a * R48_VALUE_HASH (mod 2^32) < THRESHOLD else 2
__device__ __forceinline__ uint32_t draw_word(uint64_t id, uint32_t tick, const PhiloxKeys &K)
{
    uint32_t w[4];
    philox4x32((uint32_t)id, (uint32_t)(id >> 32), tick >> 2, 0u, K, w);
    const uint32_t j = tick & 3u;
    return j == 0u ? w[0] : j == 1u ? w[1] : j == 2u ? w[2] : w[3];
}

// exponent of the spawned tile (2 for a "4", 1 for a "2") as v29 = exponent << 29 (see
// spawn_tile); 0 when nothing is spawned
__device__ __forceinline__ uint32_t spawn_v29(uint32_t a, bool changed)
{
    const uint32_t v = (a * R48_VALUE_HASH < R48_SPAWN4_THRESHOLD) ? (2u << 29) : (1u << 29);
    return changed ? v : 0u;
}

__device__ __forceinline__ uint32_t spawn_exp(uint32_t a)
{
    return (a * R48_VALUE_HASH < R48_SPAWN4_THRESHOLD) ? 2u : 1u;
}

// --- the call split for the rollout kernel: (c0, c1, c3 = 0) are fixed for an episode and only
// c2 = tick >> 2 varies, so round 0's M0*c0 product and round 1's M1*c2 product are computed once
// per episode.
struct PhiloxEpisode {
    uint32_t p0lo;      // lo(M0 * c0)                      -> c3 entering round 1
    uint32_t q1lo;      // lo(M1 * (hi(M0*c0) ^ k1[0]))     -> c1 entering round 2
    uint32_t q1hi;      // hi(M1 * (hi(M0*c0) ^ k1[0]))
};

__device__ __forceinline__ PhiloxEpisode philox_episode(uint32_t c0, const PhiloxKeys &K)
{
    const uint64_t p0 = (uint64_t)R48_PHILOX_M0 * c0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ K.k1[0];            // c3 = 0
    const uint64_t q1 = (uint64_t)R48_PHILOX_M1 * n2;
    return PhiloxEpisode{(uint32_t)p0, (uint32_t)q1, (uint32_t)(q1 >> 32)};
}

__device__ __forceinline__ void philox4x32_episode(uint32_t c1, uint32_t c2, const PhiloxEpisode &E,
                                                   const PhiloxKeys &K, uint32_t (&w)[4])
{
    // round 0 (M0*c0 and n2 come from E)
    const uint64_t a1 = (uint64_t)R48_PHILOX_M1 * c2;
    uint32_t x0 = (uint32_t)(a1 >> 32) ^ c1 ^ K.k0[0];             // c0 entering round 1
    uint32_t x1 = (uint32_t)a1;                                    // c1 entering round 1
    // round 1 (M1*c2 comes from E)
    const uint64_t b0 = (uint64_t)R48_PHILOX_M0 * x0;
    uint32_t y0 = E.q1hi ^ x1 ^ K.k0[1];
    uint32_t y2 = (uint32_t)(b0 >> 32) ^ E.p0lo ^ K.k1[1];
    uint32_t y1 = E.q1lo, y3 = (uint32_t)b0;
#pragma unroll
    for (int r = 2; r < kPhiloxRounds; r++) {
        const uint64_t p0 = (uint64_t)R48_PHILOX_M0 * y0;
        const uint64_t p1 = (uint64_t)R48_PHILOX_M1 * y2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ y1 ^ K.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ y3 ^ K.k1[r];
        y1 = (uint32_t)p1;
        y3 = (uint32_t)p0;
        y0 = n0;
        y2 = n2;
    }
    w[0] = y0; w[1] = y1; w[2] = y2; w[3] = y3;
}

// --- the call split for the single-step kernels: every board of a launch is at the same tick and
// (the host splits launches at multiples of 2^32 ids) has the same id.hi, so c1, c2, c3 are
// launch constants.  The host folds them through rounds 0 and 1 (make_philox_launch) and the
// device computes only the id.lo-dependent half of those rounds and, in the last round, only
// the word the tick uses: 23 instructions per board instead of 29.
struct PhiloxLaunch {
    uint32_t r0_n0;     // c0 entering round 1: hi(M1*c2) ^ c1 ^ k0[0]
    uint32_t r1_c1k;    // lo(M1*c2) ^ k0[1]            (c1 entering round 1, key folded in)
    uint32_t r1_h0k;    // hi(M0*r0_n0) ^ k1[1]
    uint32_t r2_c3k;    // lo(M0*r0_n0) ^ k1[2]         (c3 entering round 2, key folded in)
    uint32_t word;      // tick & 3
    uint32_t last_mul;  // M1 for words 0,1; M0 for words 2,3
    uint32_t last_key;  // k0[6] for word 0, k1[6] for word 2
};

// WORD = tick & 3 as a compile-time constant (the host picks the kernel instantiation): the last
// round is then two instructions; WORD < 0 selects by the launch constant P.word instead.
template <int WORD>
__device__ __forceinline__ uint32_t philox_launch_word(uint32_t id_lo, const PhiloxLaunch &P,
                                                       const PhiloxKeys &K)
{
    // round 0: only M0*c0 varies
    const uint64_t p0 = (uint64_t)R48_PHILOX_M0 * id_lo;
    const uint32_t a2 = (uint32_t)(p0 >> 32) ^ K.k1[0];           // c2 entering round 1 (c3 = 0)
    const uint32_t a3 = (uint32_t)p0;                             // c3 entering round 1
    // round 1: c0 = r0_n0 is a launch constant
    const uint64_t q1 = (uint64_t)R48_PHILOX_M1 * a2;
    uint32_t y0 = (uint32_t)(q1 >> 32) ^ P.r1_c1k;
    uint32_t y1 = (uint32_t)q1;
    uint32_t y2 = a3 ^ P.r1_h0k;
    // round 2: c3 is a launch constant
    {
        const uint64_t p0b = (uint64_t)R48_PHILOX_M0 * y0;
        const uint64_t p1b = (uint64_t)R48_PHILOX_M1 * y2;
        const uint32_t n0 = (uint32_t)(p1b >> 32) ^ y1 ^ K.k0[2];
        const uint32_t n2 = (uint32_t)(p0b >> 32) ^ P.r2_c3k;
        y1 = (uint32_t)p1b;
        y0 = n0; y2 = n2;
        uint32_t y3 = (uint32_t)p0b;
#pragma unroll
        for (int r = 3; r < kPhiloxRounds - 1; r++) {
            const uint64_t p0c = (uint64_t)R48_PHILOX_M0 * y0;
            const uint64_t p1c = (uint64_t)R48_PHILOX_M1 * y2;
            const uint32_t m0 = (uint32_t)(p1c >> 32) ^ y1 ^ K.k0[r];
            const uint32_t m2 = (uint32_t)(p0c >> 32) ^ y3 ^ K.k1[r];
            y1 = (uint32_t)p1c;
            y3 = (uint32_t)p0c;
            y0 = m0; y2 = m2;
        }
        // last round: only the word the tick uses.  Words 0,1 come from M1 * c2, words 2,3 from
        // M0 * c0; even words are the high half XOR a counter word XOR the key, odd words the low
        // half.  The selectors are launch constants (three selects, one multiply, one XOR).
        constexpr int L = kPhiloxRounds - 1;
        if (WORD == 0) return __umulhi(R48_PHILOX_M1, y2) ^ y1 ^ K.k0[L];
        if (WORD == 1) return R48_PHILOX_M1 * y2;
        if (WORD == 2) return __umulhi(R48_PHILOX_M0, y0) ^ y3 ^ K.k1[L];
        if (WORD == 3) return R48_PHILOX_M0 * y0;
        const bool upper = P.word >= 2u, odd = (P.word & 1u) != 0u;
        const uint64_t pr = (uint64_t)(upper ? y0 : y2) * P.last_mul;
        const uint32_t even_word = (uint32_t)(pr >> 32) ^ (upper ? y3 : y1) ^ P.last_key;
        return odd ? (uint32_t)pr : even_word;
    }
}

// ------------------------------------------------------------------ board symmetries

// 4x4 nibble transpose: 2x2 blocks inside each word, then the off-diagonal 2x2 blocks
// change words (two PRMTs).  10 instructions.
__device__ __forceinline__ void transpose(uint32_t &lo, uint32_t &hi)
{
    uint32_t a = bsel(0x0000F0F0u, lo >> 12, bsel(0xF0F00F0Fu, lo, lo << 12));
    uint32_t b = bsel(0x0000F0F0u, hi >> 12, bsel(0xF0F00F0Fu, hi, hi << 12));
    lo = prmt(a, b, 0x6240);
    hi = prmt(a, b, 0x7351);
}

// transpose(lo, hi) in the lanes where `flag` is nonzero, without a branch: the eight block
// instructions run unconditionally into temporaries and only the two final PRMTs are predicated.
// Where the condition is per board (the move's axis), some lane of a warp always takes the
// branch, so a branch never skips anything and its BSSY / BRA / BSYNC are three wasted issue slots
// per transpose -- in kernels that are bound by issue slots.  (Predicating all ten instructions in
// PTX does not survive ptxas: it turns predicated writes of temporaries into selects.)
#ifndef R48_PRED_TRANSPOSE
#define R48_PRED_TRANSPOSE 1
#endif
__device__ __forceinline__ void transpose_where(uint32_t flag, uint32_t &lo, uint32_t &hi)
{
#if R48_PRED_TRANSPOSE
    const uint32_t a = bsel(0x0000F0F0u, lo >> 12, bsel(0xF0F00F0Fu, lo, lo << 12));
    const uint32_t b = bsel(0x0000F0F0u, hi >> 12, bsel(0xF0F00F0Fu, hi, hi << 12));
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@p prmt.b32 %0, %2, %3, 0x6240;\n\t"
        "@p prmt.b32 %1, %2, %3, 0x7351;\n\t"
        "}"
        : "+r"(lo), "+r"(hi) : "r"(a), "r"(b), "r"(flag));
#else
    if (flag) transpose(lo, hi);
#endif
}

// swap the two nibbles of every byte (half of a row reversal; the byte swap is folded
// into the PRMTs that extract / re-pack rows)
__device__ __forceinline__ uint32_t nibswap(uint32_t w)
{
    return bsel(0xF0F0F0F0u, w << 4, w >> 4);
}

// ------------------------------------------------------------------ the move
// A move is "slide the ROWS of the stored board toward nibble 0 (LEFT) or nibble 3 (RIGHT)".
// UP/DOWN are the same thing on the transposed board: `orient` transposes for a vertical action,
// and the callers decide when to transpose back (the single-step kernels right after the spawn,
// the fused rollout only when the next move's axis differs).

__device__ __forceinline__ bool is_vertical(uint32_t action) { return action < 2u; }      // UP, DOWN
__device__ __forceinline__ bool is_toward_high(uint32_t action) { return (action & 1u) != 0u; }  // DOWN, RIGHT

// --- 16-bit tables (reward_mode 1 kernels).  One lookup per row: left[r] = row r after a LEFT
// move; RIGHT = reverse the row, LEFT, reverse.  `merges`: two 4-bit exponents of the merged
// pairs of the row.
template <bool WITH_REWARD>
__device__ __forceinline__ void rows_l16(uint32_t &lo, uint32_t &hi, bool toward_high,
                                         const uint16_t *__restrict__ left,
                                         const uint8_t *__restrict__ merges, uint32_t &reward)
{
    if (toward_high) { lo = nibswap(lo); hi = nibswap(hi); }
    // row -> table index; for reversed rows the PRMT also swaps the two bytes
    const uint32_t ex0 = toward_high ? 0x4401u : 0x4410u;
    const uint32_t ex1 = toward_high ? 0x4423u : 0x4432u;
    const uint32_t r0 = prmt(lo, 0u, ex0), r1 = prmt(lo, 0u, ex1);
    const uint32_t r2 = prmt(hi, 0u, ex0), r3 = prmt(hi, 0u, ex1);
    const uint32_t o0 = left[r0], o1 = left[r1], o2 = left[r2], o3 = left[r3];
    if (WITH_REWARD) {
        uint32_t m = (uint32_t)merges[r0] | ((uint32_t)merges[r1] << 8) |
                     ((uint32_t)merges[r2] << 16) | ((uint32_t)merges[r3] << 24);
        uint32_t sum = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) sum += (1u << ((m >> (4 * t)) & 15u)) & ~1u;
        reward = sum << 1;                              // merging two 2^e tiles yields 2^(e+1)
    }
    const uint32_t pk = toward_high ? 0x4501u : 0x5410u;
    lo = prmt(o0, o1, pk);
    hi = prmt(o2, o3, pk);
    if (toward_high) { lo = nibswap(lo); hi = nibswap(hi); }
}

// --- LR table (reward_mode 0 kernels: step, afterstates, rollout).  lr[r] = LEFT result of row r
// in the low half, RIGHT result in the high half, for r < kLrRows (rows whose last cell is below
// 2^14 -- 224 KB, what fits beside nothing else in one SM's shared memory).  One lookup per
// row serves both directions, so there is no row reversal; the PRMT that re-packs two rows
// picks the half.  Rows outside the table (a 16384 or 32768 tile in the last cell) take a
// serial path; random play never gets there, strong players occasionally do.

constexpr uint32_t kLrRows = 0xE000u;               // 57344 entries x 4 B = 229376 B

// slot of row r in the LR table.  With R48_SWIZZLE bits 7..11 of the row (cells 1..2) are folded
// into the low five bits -- the shared-memory bank -- which otherwise come from cell 0 alone and
// cluster on the few small exponents that dominate real boards.
__host__ __device__ __forceinline__ uint32_t lr_slot(uint32_t r)
{
#if R48_SWIZZLE
    return r ^ ((r >> 7) & 31u);
#else
    return r;
#endif
}

// multipliers that turn shifts into FMA-pipe multiplies; they arrive as kernel parameters so that
// ptxas cannot strength-reduce them back into shifts
struct PipeConsts {
    uint32_t m16, m18;          // 1 << 16, 1 << 18
};

// byte offsets (4 * slot) of the two rows of a word
__device__ __forceinline__ void row_offsets(uint32_t w, const PipeConsts &pc, uint32_t &a0, uint32_t &a1)
{
#if R48_FMA_INDEX
    uint32_t plo, phi;                                   // w * 2^16 = (w >> 16) : (w << 16)
    asm("{\n.reg .u64 t;\nmul.wide.u32 t, %2, %3;\nmov.b64 {%0, %1}, t;\n}" : "=r"(plo), "=r"(phi) : "r"(w), "r"(pc.m16));
    a0 = __umulhi(plo, pc.m18);                          // ((w << 16) * 2^18) >> 32 = (w & 0xFFFF) * 4
    a1 = phi * 4u;
#if R48_SWIZZLE
    a0 ^= (plo >> 21) & 0x7Cu;                           // (row >> 7 & 31) << 2
    a1 ^= (phi >> 5) & 0x7Cu;
#endif
#else
    a0 = (w << 2) & 0x3FFFCu;
    a1 = (w >> 14) & 0x3FFFCu;
#if R48_SWIZZLE
    a0 ^= (w >> 5) & 0x7Cu;
    a1 ^= (w >> 21) & 0x7Cu;
#endif
#endif
}

// compress / merge-once / compress of one row, toward nibble 0 or toward nibble 3; all in
// registers (no local array).  It runs only for rows outside the table.
__device__ __forceinline__ uint32_t slow_row_inline(uint32_t r, bool toward_high)
{
    if (toward_high) r = ((r & 0xFu) << 12) | ((r & 0xF0u) << 4) | ((r >> 4) & 0xF0u) | ((r >> 12) & 0xFu);
    uint32_t packed = 0, n = 0;                     // non-empty cells, packed toward nibble 0
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const uint32_t e = (r >> (4 * t)) & 15u;
        if (e) { packed |= e << (4 * n); n++; }
    }
    uint32_t out = 0, o = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {                   // at most 4 output cells
        const uint32_t e = packed & 15u, f = (packed >> 4) & 15u;
        if (e == 0u) break;
        if (e == f) { out |= (e + 1u > 15u ? 15u : e + 1u) << (4 * o); packed >>= 8; }
        else { out |= e << (4 * o); packed >>= 4; }
        o++;
    }
    if (toward_high) out = ((out & 0xFu) << 12) | ((out & 0xF0u) << 4) | ((out >> 4) & 0xF0u) | ((out >> 12) & 0xFu);
    return out;
}

// out of line, for the kernels whose cold paths would otherwise inline many copies
__device__ __noinline__ uint32_t slow_row(uint32_t r, bool toward_high)
{
    return slow_row_inline(r, toward_high);
}

// `lr` below is the shared-window byte address of the staged table (smem_u32)

// The table is staged in two parts: rows below kLrSplit (last cell below 256 -- nearly every row of
// nearly every board) complete on one barrier, the rest on a second one, so that a launch starts
// looking rows up when 128 KB of the 224 KB have arrived.  One test on bit 15 of the row words
// covers both rare cases -- "needs the second part" and "outside the table".
constexpr uint32_t kLrSplit = 0x8000u;

// GUARD = false skips the range test: only for callers that can PROVE every row is in the table
// (the rollout kernel: a board whose tiles sum to less than 16384 has no 16384 tile) and that have
// waited for both parts.  `need_high` is called (once per board at most) before any row >= kLrSplit
// is looked up.
// `pk` = the PRMT selector that re-packs two looked-up rows into a word: 0x5410 takes the low
// halves (LEFT results), 0x7632 the high halves (RIGHT results).
__device__ __forceinline__ uint32_t pack_selector(bool toward_high) { return toward_high ? 0x7632u : 0x5410u; }

// where the LR table is read from: the staged copy in shared memory (every shipped kernel), or
// global memory through the read-only L1 path (the no-staging A/B variant of step_kernel)
struct TableInShared {
    uint32_t base;                                   // shared-window byte address (smem_u32)
    __device__ __forceinline__ uint32_t operator()(uint32_t byte_offset) const { return lds_u32(base + byte_offset); }
};
struct TableInGlobal {
    uint64_t base;                                   // global address of Tables::lr
    __device__ __forceinline__ uint32_t operator()(uint32_t byte_offset) const
    {
        uint32_t v;
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(base + byte_offset));
        return v;
    }
};

template <bool GUARD = true, typename NeedHigh, typename Table>
__device__ __forceinline__ void rows_lr(uint32_t &lo, uint32_t &hi, uint32_t pk, const Table table,
                                        const PipeConsts &pc, NeedHigh need_high)
{
    const bool toward_high = pk != 0x5410u;          // used on the cold exact path only
    uint32_t o0, o1, o2, o3;
    bool exact = false;
    if (GUARD) {
        const uint32_t both = lo | hi;
        if (__builtin_expect((both & 0x80008000u) != 0u, 0)) {          // some row >= kLrSplit
            need_high();
            // a row is outside the table iff its top three bits are set; the OR of the words is a
            // conservative test (it may send a board to the exact path for nothing)
            exact = (both & (both << 1) & (both << 2) & 0x80008000u) != 0u;
        }
    }
    if (!exact) {
        uint32_t a0, a1, a2, a3;
        row_offsets(lo, pc, a0, a1);
        row_offsets(hi, pc, a2, a3);
        o0 = table(a0); o1 = table(a1);
        o2 = table(a2); o3 = table(a3);
    } else {
        // exact path: one row at a time in a rolled loop, nothing out of line -- a CALL inside the
        // caller's loop would make ptxas reload every launch constant after it, on the hot path too
        const uint64_t board = ((uint64_t)hi << 32) | lo;
        uint64_t moved = 0ull;
#pragma unroll 1
        for (uint32_t t = 0; t < 4u; t++) {
            const uint32_t r = (uint32_t)(board >> (16u * t)) & 0xFFFFu;
            uint32_t o;
            if (r < kLrRows) o = (table(4u * lr_slot(r)) >> (toward_high ? 16u : 0u)) & 0xFFFFu;
            else o = slow_row_inline(r, toward_high);
            moved |= (uint64_t)o << (16u * t);
        }
        lo = (uint32_t)moved; hi = (uint32_t)(moved >> 32);
        return;
    }
    lo = prmt(o0, o1, pk);
    hi = prmt(o2, o3, pk);
}

template <bool GUARD = true, typename NeedHigh>
__device__ __forceinline__ void rows_lr(uint32_t &lo, uint32_t &hi, uint32_t pk, uint32_t lr,
                                        const PipeConsts &pc, NeedHigh need_high)
{
    rows_lr<GUARD>(lo, hi, pk, TableInShared{lr}, pc, need_high);
}


// all four afterstates: LEFT/RIGHT share one lookup per row, UP/DOWN one per column
template <bool GUARD = true>
__device__ __forceinline__ void move_all(uint32_t lo, uint32_t hi, uint32_t lr,
                                         uint32_t (&rl)[4], uint32_t (&rh)[4])
{
    uint32_t tl = lo, th = hi;
    transpose(tl, th);
    const uint32_t r[8] = {lo & 0xFFFFu, lo >> 16, hi & 0xFFFFu, hi >> 16,
                           tl & 0xFFFFu, tl >> 16, th & 0xFFFFu, th >> 16};
    uint32_t o[8];
    uint32_t any = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) any |= r[t];
    if (!GUARD || __builtin_expect(any < kLrRows, 1)) {
#pragma unroll
        for (int t = 0; t < 8; t++) o[t] = lds_u32(lr + 4u * lr_slot(r[t]));
    } else {
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (r[t] < kLrRows) o[t] = lds_u32(lr + 4u * lr_slot(r[t]));
            else o[t] = slow_row(r[t], false) | (slow_row(r[t], true) << 16);
        }
    }
    rl[2] = prmt(o[0], o[1], 0x5410); rh[2] = prmt(o[2], o[3], 0x5410);     // LEFT
    rl[3] = prmt(o[0], o[1], 0x7632); rh[3] = prmt(o[2], o[3], 0x7632);     // RIGHT
    rl[0] = prmt(o[4], o[5], 0x5410); rh[0] = prmt(o[6], o[7], 0x5410);     // UP
    rl[1] = prmt(o[4], o[5], 0x7632); rh[1] = prmt(o[6], o[7], 0x7632);     // DOWN
    transpose(rl[0], rh[0]);
    transpose(rl[1], rh[1]);
}

// ------------------------------------------------------------------ empties / spawn

// bit 3 of every nibble = 1 where the nibble is zero: (y & 7) + 7 carries into bit 3 unless
// the low three bits are zero; OR-ing y covers bit 3 itself.  LOP3, IADD, LOP3.
__device__ __forceinline__ uint32_t zero_nibbles8(uint32_t y)
{
    const uint32_t s = (y & 0x77777777u) + 0x77777777u;
    return ~(s | y) & 0x88888888u;
}

// Exclusive prefix count of blanks per nibble (row-major = ascending nibble order, the
// order of the reference's blank list, GameClient.py:109-114) and their total.
struct Blanks {
    uint32_t el, eh;    // blank flags at bit 3 of each nibble
    uint32_t ql, qh;    // q[i] = number of blanks below nibble i (0..15, never overflows)
    uint32_t n;         // number of blanks (0..16)
};

__device__ __forceinline__ Blanks count_blanks(uint32_t lo, uint32_t hi)
{
    Blanks b;
    b.el = zero_nibbles8(lo);
    b.eh = zero_nibbles8(hi);
    // flags/8 * 0x1111111111111110 == flags * 0x0222222222222222 : nibble i of the product
    // = number of flags below nibble i
    const uint64_t p = (uint64_t)b.el * 0x22222222u;
    b.ql = (uint32_t)p;
    b.qh = (uint32_t)(p >> 32) + b.el * 0x02222222u + b.eh * 0x22222222u;
    b.n = (b.qh >> 28) + (b.eh >> 31);
    return b;
}

// one-hot (bit 3 of the chosen nibble) mask of the k-th blank; zero if n <= k <= 15 (k * 0x11111111
// wraps for k >= 16: callers that take k from outside use place_tile_checked)
__device__ __forceinline__ void kth_blank(const Blanks &b, uint32_t k, uint32_t &sl, uint32_t &sh)
{
    const uint32_t kk = k * 0x11111111u;
    const uint32_t xl = b.ql ^ kk, xh = b.qh ^ kk;       // zero nibble where q[i] == k
    const uint32_t tl = (xl & 0x77777777u) + 0x77777777u;
    const uint32_t th = (xh & 0x77777777u) + 0x77777777u;
    sl = ~(tl | xl) & b.el;                              // ... and the cell is blank
    sh = ~(th | xh) & b.eh;
}

// Put exponent `vexp` (0 = nothing) into the k-th blank.
__device__ __forceinline__ void place_tile(uint32_t &lo, uint32_t &hi, const Blanks &b, uint32_t k,
                                           uint32_t vexp)
{
    uint32_t sl, sh;
    kth_blank(b, k, sl, sh);
    lo += (sl >> 3) * vexp;
    hi += (sh >> 3) * vexp;
}

// k from the caller (injected draws): any k >= n leaves the board as it is
__device__ __forceinline__ void place_tile_checked(uint32_t &lo, uint32_t &hi, const Blanks &b, uint32_t k,
                                                   uint32_t vexp)
{
    place_tile(lo, hi, b, k & 15u, k < b.n ? vexp : 0u);
}

// Spawn from the tick's word `a` when `changed`: the tile goes into the k-th blank,
// k = mulhi(a << 2, n).  With v29 = exponent << 29 the high half of (1 << (4i+3)) * v29 is
// exponent << 4i, so the insert is one IMAD.HI per word (with the board as the 64-bit addend's
// high half).  A shift of the one-hot mask plus predicated ORs was measured 4 % slower: the
// predicated value computation issues whether or not it is needed.
__device__ __forceinline__ void spawn_tile(uint32_t &lo, uint32_t &hi, const Blanks &b, uint32_t a, bool changed)
{
    uint32_t sl, sh;
    kth_blank(b, __umulhi(a << 2, b.n), sl, sh);
    const uint32_t v29 = spawn_v29(a, changed);
    lo += __umulhi(sl, v29);
    hi += __umulhi(sh, v29);
}

// ------------------------------------------------------------------ game over

// nonzero iff some nibble of y is zero (exact as an any-test)
__device__ __forceinline__ uint32_t any_zero_nibble(uint32_t y)
{
    return (y - 0x11111111u) & ~y & 0x88888888u;
}

// no two equal horizontal or vertical neighbours (meaningful for a full board)
__device__ __forceinline__ bool no_equal_neighbours(uint32_t lo, uint32_t hi)
{
    const uint32_t hl = (lo ^ (lo >> 4)) | 0xF000F000u;           // col 3 has no right neighbour
    const uint32_t hh = (hi ^ (hi >> 4)) | 0xF000F000u;
    const uint32_t vl = lo ^ prmt(lo, hi, 0x5432);                // rows 0,1 vs rows 1,2
    const uint32_t vh = (hi ^ (hi >> 16)) | 0xFFFF0000u;          // row 2 vs row 3
    return (any_zero_nibble(hl) | any_zero_nibble(hh) | any_zero_nibble(vl) |
            any_zero_nibble(vh)) == 0u;
}

// Game.has_game_over: full and no equal neighbours
__device__ __forceinline__ bool game_over(uint32_t lo, uint32_t hi)
{
    const bool full = (any_zero_nibble(lo) | any_zero_nibble(hi)) == 0u;
    return full && no_equal_neighbours(lo, hi);
}

// ------------------------------------------------------------------ readout helpers

__device__ __forceinline__ uint32_t board_score(uint32_t lo, uint32_t hi)   // sum of tiles
{
    uint32_t s = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        s += (1u << ((lo >> (4 * t)) & 15u)) & ~1u;
        s += (1u << ((hi >> (4 * t)) & 15u)) & ~1u;
    }
    return s;
}

__device__ __forceinline__ uint32_t board_max_exp(uint32_t lo, uint32_t hi)
{
    uint32_t m = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        m = max(m, (lo >> (4 * t)) & 15u);
        m = max(m, (hi >> (4 * t)) & 15u);
    }
    return m;
}

// ------------------------------------------------------------------ table staging (TMA bulk copy)
// The 128 KB LEFT table (and the 64 KB merge table in reward mode) is copied global ->
// shared by the bulk-copy engine (cp.async.bulk, SASS UBLKCP) and signalled on an mbarrier,
// so the copy overlaps the prologue (first loads, first Philox call) of each CTA.

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Programmatic dependent launch: a kernel launched with the stream-serialisation attribute
// may start (and stage its tables) while the previous kernel in the stream is still draining;
// it must call pdl_wait() before touching anything the previous kernel wrote.
__device__ __forceinline__ void pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ void pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// CLOBBER = false: for tables that are only ever read through lds_u32 (volatile asm, so already
// ordered after this wait); without the "memory" clobber the compiler does not re-read kernel
// parameters after a wait that sits inside a loop.
template <bool CLOBBER = true>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    if (CLOBBER) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(smem_u32(bar)),
            "r"(parity)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(smem_u32(bar)),
            "r"(parity));
    }
}

}  // namespace r48
