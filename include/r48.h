/*
 * r48.h -- C ABI of libr48.so, the B200 (sm_100a) batched 2048 environment.
 *
 * The reference (nevertiree/Rein48) has no FFI: its environment is the Python class
 * `Game` (game/GameClient.py:15-51) driven by `Rand.random_action` (control/rand.py:9-11)
 * inside `play` (main.py:36-42).  Each entry point below replaces one of those Python
 * call sites for a whole batch of boards; the reference line it stands for is cited.
 * INTEGRATION.md shows the ctypes binding a Rein48 maintainer would add.
 *
 * Conventions
 *   - Board = one uint64 of 16 exponent nibbles: cell (i,j) of the reference's
 *     state_matrix[i][j] is nibble 4*i+j (nibble 0 = bits 0..3); nibble e means tile 2^e,
 *     0 means empty.  LEFT moves toward nibble 0 of each row, UP toward row 0.
 *   - Actions are bytes: 0 UP, 1 DOWN, 2 LEFT, 3 RIGHT (GameClient.py:140,182,206,230).
 *   - Every pointer is a DEVICE pointer unless the function name ends in _host.  The
 *     caller owns all buffers; device entry points allocate nothing per call (the row
 *     tables are built once per device on first use) and only enqueue work on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).  No implicit sync.
 *   - Return value: 0 on success, a negative R48_ERR_* otherwise; nothing throws across
 *     the boundary.  r48_last_error() gives a thread-local message for the last failure.
 *   - Random draws come from Philox4x32-7 keyed by (seed, global board id, tick >> 2), one
 *     32-bit word per tick; the global id of element i of a batch is board_base + i.  A spawn
 *     counts the blanks in the order of the move's axis (row-major after LEFT/RIGHT,
 *     column-major after UP/DOWN).  See DESIGN.md "draw spec".
 *   - Domain edge: merging two 32768 tiles would give 65536, which has no nibble; the merge
 *     SATURATES (the result is one 32768 tile, the other disappears, the board counts as
 *     changed).  Transitions are bit-exact with the reference for every board that does not
 *     merge two 32768 tiles.
 *   - Threads: the device entry points keep no per-call state and may be called from any number
 *     of host threads (each on its own stream).  The *_host entry points share one scratch
 *     arena, their streams and events per device and take a per-device mutex for the whole
 *     call: concurrent calls on one device are safe and run one after another, calls on
 *     different devices run concurrently.  r48_last_error() is thread-local.
 *   - reward_mode 0 = reference (reward is always 0, GameClient.py:138);
 *     reward_mode 1 = merge_sum (sum of the tiles created by merges; extension).
 */
#ifndef R48_H
#define R48_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R48_VERSION 201

#define R48_OK             0
#define R48_ERR_NULL      -1   /* required pointer is NULL */
#define R48_ERR_ALIGN     -2   /* pointer not aligned to its element size */
#define R48_ERR_ARG       -3   /* n < 0, n too large, bad reward_mode, ... */
#define R48_ERR_CUDA      -4   /* a CUDA call failed; see r48_last_error() */
#define R48_ERR_ACTION    -5   /* an action byte > 3 (GameClient.py:254 raises ValueError) */

#define R48_UP    0
#define R48_DOWN  1
#define R48_LEFT  2
#define R48_RIGHT 3

/* SUM-reducible episode statistics vector (uint64 words).  One all-reduce(SUM) of this
 * vector is the only collective of the multi-GPU rollout. */
#define R48_STATS_EPISODES    0
#define R48_STATS_SUM_LEN     1
#define R48_STATS_SUM_SCORE   2
#define R48_STATS_SUM_SCORE2  3
#define R48_STATS_SUM_LEN2    4
#define R48_STATS_HIST_MAXEXP 8      /* 16 bins: exponent of the largest tile */
#define R48_STATS_HIST_LEN    24     /* 2048 bins: min(length, 2047) */
#define R48_STATS_HIST_SCORE  2072   /* 2048 bins: min(score / 2, 2047) */
#define R48_STATS_WORDS       4120

#define R48_ROLLOUT_WORKSPACE_BYTES 256

int r48_version(void);
/* "R48_BUILD_ID=<16 hex digits>": a hash of the sources this binary was compiled from (the
 * Python loader refuses a library built from other sources). */
const char *r48_build_id(void);
const char *r48_last_error(void);

/* Build the per-device row tables (idempotent; other calls do it lazily). */
int r48_init(int device);
/* Copy the device row tables to host memory for inspection: left[65536] = result row of
 * a LEFT move, merges[65536] = exponents of the (<= 2) merged pairs, 4 bits each. */
int r48_debug_tables_host(uint16_t *left, uint8_t *merges, int device);

/* Game.reset (GameClient.py:33-38): empty board + ONE spawned tile (tick 0). */
int r48_reset(uint64_t *boards, int64_t n, uint64_t seed, uint64_t board_base, void *stream);

/* Game.step (GameClient.py:40-51): move; if the board changed, spawn one tile; done =
 * has_game_over.  `step` = number of step() calls already applied to these boards (this
 * call is tick step+1).  in == out is allowed.  reward / done may be NULL.  `status`
 * (optional device int32) is OR-ed with 1 if any action byte was > 3; such boards are
 * passed through unchanged. */
int r48_step(const uint64_t *in, const uint8_t *action, uint64_t *out, int32_t *reward,
             uint8_t *done, int64_t n, uint64_t seed, uint64_t board_base, uint32_t step,
             int reward_mode, int32_t *status, void *stream);

/* The same with the spawn draws supplied by the caller (parity mode): spawn_k[i] is what
 * random.randint(0, n_blank-1) returned (GameClient.py:121), spawn_exp[i] is 1 for a 2,
 * 2 for a 4 (GameClient.py:125).  Both are ignored where the move changes nothing; a spawn_k
 * that is not below the number of blanks places nothing. */
int r48_step_injected(const uint64_t *in, const uint8_t *action, const uint8_t *spawn_k,
                      const uint8_t *spawn_exp, uint64_t *out, int32_t *reward, uint8_t *done,
                      int64_t n, int reward_mode, int32_t *status, void *stream);

/* Vectorised-environment step for a policy in the loop (the a3c.py:187-243 / ddpg.py:12-70
 * worker loops, batched): every env carries its own counters, so envs may be at different
 * points of different episodes.  Env i plays global board id
 *     board_base + i + episodes[i] * id_stride          (id_stride >= n, e.g. the world batch)
 * and this call is tick steps[i]+1 of it.  After Game.step, steps[i] is incremented; if the
 * game is over and auto_reset != 0 the final board is kept in final_boards[i] (optional), the
 * env is reset in place (Game.reset of its next episode id: episodes[i]++, steps[i] = 0) and
 * done[i] = 1 is returned together with the NEW episode's first board, gym-vector style.
 * obs (optional) = float32 [n][4][4] readout of the board left in boards[i]: tile values
 * (obs_mode 0, np.array(state) of a3c.py:195) or exponents (obs_mode 1). */
int r48_env_step(uint64_t *boards, const uint8_t *action, uint32_t *steps, uint32_t *episodes,
                 int32_t *reward, uint8_t *done, float *obs, int obs_mode, uint64_t *final_boards,
                 int64_t n, uint64_t seed, uint64_t board_base, uint64_t id_stride,
                 int reward_mode, int auto_reset, int32_t *status, void *stream);

/* r48_env_step that also appends every env's transition (board before the step, action, reward,
 * board after the step -- the finished board, not the auto-reset one -- and done) to a
 * transition ring (see r48_ring below; n <= 2^30).  ring == NULL: plain r48_env_step. */
struct r48_ring;
int r48_env_step_ring(uint64_t *boards, const uint8_t *action, uint32_t *steps, uint32_t *episodes,
                      int32_t *reward, uint8_t *done, float *obs, int obs_mode,
                      uint64_t *final_boards, int64_t n, uint64_t seed, uint64_t board_base,
                      uint64_t id_stride, int reward_mode, int auto_reset, int32_t *status,
                      const struct r48_ring *ring, void *stream);

/* Game.random_fill_grid alone (GameClient.py:102-127) with injected draws: put exponent
 * spawn_exp[i] into the spawn_k[i]-th blank (row-major) of boards[i]; any k >= n_blank (up to
 * 255) leaves the board as it is (a full board is returned unchanged, GameClient.py:117-118). */
int r48_spawn_injected(uint64_t *boards, const uint8_t *spawn_k, const uint8_t *spawn_exp,
                       int64_t n, void *stream);

/* One move for a player whose policy AND random draws live on the host -- the reference's own
 * main.play loop (main.py:36-42) over a `Game` (GameClient.py:40-51): Game.step with injected
 * draws (as r48_step_injected, in place), the readout Game.step returns, and one move of lookahead
 * so that the host can make the NEXT step's draws in the reference's order without another call.
 * action[i] = R48_ACTION_NONE skips the move and spawns unconditionally (Game.reset on a zeroed
 * board, GameClient.py:33-38; spawn_exp[i] = 0 only reads the board out).  An action byte that is
 * neither 0..3 nor R48_ACTION_NONE leaves the board as it is and sets bit 0 of *status.
 * All pointers may be device memory or pinned host memory: with pinned buffers a move is one
 * launch and one cudaStreamSynchronize.  n <= 2^31 - 1 boards, any of which are independent. */
#define R48_ACTION_NONE 255
struct r48_game_view {
    int32_t cells[16];   /* state_matrix: tile VALUES, row-major, 0 = empty */
    int32_t reward;      /* as r48_step's reward[i] */
    uint8_t done;        /* Game.has_game_over(new board) */
    uint8_t valid;       /* bit a: action a (0=UP 1=DOWN 2=LEFT 3=RIGHT) changes the new board */
    uint8_t blanks[4];   /* blank cells of the new board after action a, before the spawn */
    uint8_t reserved[2];
};
int r48_step_injected_view(uint64_t *boards, const uint8_t *action, const uint8_t *spawn_k,
                           const uint8_t *spawn_exp, int64_t n, int reward_mode,
                           struct r48_game_view *views, int32_t *status, void *stream);

/* The same with the draws made on the GPU from the Philox word of (seed, board, tick); there is
 * no move here, so the blanks are counted row-major. */
int r48_spawn(uint64_t *boards, int64_t n, uint64_t seed, uint64_t board_base, uint32_t tick,
              void *stream);

/* Number of blank cells per board (len(blank_grid_index_list), GameClient.py:109-114). */
int r48_blank_counts(const uint64_t *boards, uint8_t *counts, int64_t n, void *stream);

/* 1-ply afterstates: Game.update_matrix(copy, a) for a = 0..3 (GameClient.py:129-254), no
 * spawn.  Planar outputs, one plane per action: out[a*n + i], reward[a*n + i] (so every store
 * of a warp is one contiguous run); valid[i] bit a = move a changes the board;
 * done[i] = Game.has_game_over (GameClient.py:65-94).  reward/valid/done may be NULL. */
int r48_afterstates(const uint64_t *in, uint64_t *out, int32_t *reward, uint8_t *valid,
                    uint8_t *done, int64_t n, int reward_mode, void *stream);

/* main.play with control="rand" (main.py:36-42, rand.py:9-11) for episodes
 * board_base .. board_base+n-1, each from reset to game over, fused in one kernel.
 * final_boards[i] / lengths[i] (steps, no-op moves included) are required outputs.
 * If `stats` is non-NULL the R48_STATS_WORDS vector is ACCUMULATED into it (+=).
 * `workspace` = R48_ROLLOUT_WORKSPACE_BYTES of device scratch (8-byte aligned). */
int r48_rollout(int64_t n, uint64_t seed, uint64_t board_base, uint64_t *final_boards,
                uint32_t *lengths, uint64_t *stats, void *workspace, void *stream);

/* The same loop with another built-in policy (extension):
 *   R48_POLICY_RANDOM         control/rand.py (what r48_rollout runs)
 *   R48_POLICY_GREEDY_BLANKS  1-ply greedy: among the moves that change the board take the one
 *                             whose afterstate has the most blank cells; candidates are visited
 *                             in the order r, r+1, r+2, r+3 (mod 4), r = the tick's random
 *                             action, first best wins; spawn and game over as in Game.step. */
#define R48_POLICY_RANDOM        0
#define R48_POLICY_GREEDY_BLANKS 1
int r48_rollout_policy(int64_t n, uint64_t seed, uint64_t board_base, int policy,
                       uint64_t *final_boards, uint32_t *lengths, uint64_t *stats, void *workspace,
                       void *stream);

/* Trajectories of the same episodes (what the learners buffer per step: a3c.py:205-209,
 * ddpg.py:29-31 -- and without ddpg.py:31's aliasing of state and next_state): replays
 * episodes board_base..+n-1 under `policy` -- draws are counter-based, so an episode is a pure
 * function of (seed, id) -- and writes, for step t = 1..lengths[i] of episode i,
 *     traj_boards[offsets[i] + t-1]  = the board BEFORE the step (the state the policy saw)
 *     traj_actions[offsets[i] + t-1] = the action taken
 * lengths[] comes from r48_rollout(_policy) with the same arguments.  offsets[i] must be a
 * MULTIPLE OF 4 with room for lengths[i] rounded up to a multiple of 4 (e.g. the exclusive prefix
 * sum of (lengths + 3) & ~3): four steps leave as one 32-byte sector, and the up-to-3 padding
 * slots after an episode's last step hold unspecified values.  traj_boards must be 32-byte,
 * traj_actions 4-byte aligned.  The board after the last step is final_boards[i] of that call. */
int r48_rollout_trajectories(int64_t n, uint64_t seed, uint64_t board_base, int policy,
                             const uint32_t *lengths, const uint64_t *offsets, uint64_t *traj_boards,
                             uint8_t *traj_actions, void *workspace, void *stream);

/* Episode statistics of finished games, accumulated into stats[R48_STATS_WORDS]. */
int r48_episode_stats(const uint64_t *final_boards, const uint32_t *lengths, int64_t n,
                      uint64_t *stats, void *stream);

/* The packed per-episode record described at r48_rollout_host_ex, on the device. */
int r48_episode_records(const uint64_t *final_boards, const uint32_t *lengths, uint32_t *records,
                        int64_t n, void *stream);

/* np.sum(state_matrix) (main.py:48) and the largest tile's exponent, per board. */
int r48_scores(const uint64_t *boards, uint32_t *score, uint8_t *max_exp, int64_t n,
               void *stream);

/* Board readout for the learners (np.array(state), a3c.py:195,205): out[n][4][4] tile
 * values, or exponents when log2_planes != 0. */
int r48_decode_f32(const uint64_t *boards, float *out, int64_t n, int log2_planes, void *stream);
int r48_decode_i32(const uint64_t *boards, int32_t *out, int64_t n, void *stream);
/* The inverse: tile values (0 or a power of two >= 2; anything else sets *status bit 1). */
int r48_encode_i32(const int32_t *values, uint64_t *boards, int64_t n, int32_t *status,
                   void *stream);

/* ---- transition ring: Replay (algorithm/ddpg/replay.py:8-47) for batches, on the device ----
 * One caller-owned device array of `capacity` 32-byte records plus a two-word device cursor
 * (cursor[0] = transitions appended since the last clear, cursor[1] = scratch, both zero
 * initially).  A record is one DRAM sector, so gathering a random slot costs one sector (as five
 * separate arrays a sample read 19 times the bytes it returned).  The r48_ring struct itself
 * lives in HOST memory.  The cursor lives on the device so that appends and samples need no host
 * synchronisation and can be captured in CUDA graphs. */
typedef struct r48_transition {
    uint64_t state;         /* board before the step            (ddpg.py:31 `state`)      */
    uint64_t next_state;    /* board after the step, a distinct value (ddpg.py:29-31 stores
                               the SAME list object twice; that aliasing is not reproduced) */
    int32_t  reward;
    uint8_t  action;
    uint8_t  done;
    uint8_t  reserved[10];  /* written as zero */
} r48_transition;           /* 32 bytes */

typedef struct r48_ring {
    r48_transition *slots;  /* [capacity], 32-byte aligned */
    uint64_t *cursor;       /* [2] */
    uint64_t  capacity;     /* Replay.max_size (replay.py:11) */
} r48_ring;

/* Replay.clear (replay.py:45-47). */
int r48_ring_clear(const r48_ring *ring, void *stream);

/* Replay.store (replay.py:18-21) for n transitions.  drop_when_full != 0 is the reference's
 * rule (a full buffer ignores new transitions: the first capacity - size of the batch are kept);
 * drop_when_full == 0 overwrites the oldest slot, transition i going to slot
 * (cursor + i) % capacity.  reward / done may be NULL (stored as 0). */
int r48_ring_append(const r48_ring *ring, const uint64_t *state, const uint8_t *action,
                    const int32_t *reward, const uint64_t *next_state, const uint8_t *done,
                    int64_t n, int drop_when_full, void *stream);

/* Replay.sample (replay.py:23-27, random.sample of :33) without its clear(): gathers `batch`
 * transitions from the size = min(cursor, capacity) valid slots.
 *   with_replacement == 0: distinct slots (element i is slot perm(i) of a keyed permutation of
 *     [0, size)); elements i >= size are skipped and get out_index -1 (the reference returns
 *     the whole list when asked for more than it holds);
 *   with_replacement != 0: independent uniform slots.
 * Draws are keyed by (seed, draw, element): the same arguments give the same sample.
 * out_index (optional) receives the slot numbers; obs_state / obs_next_state (optional) receive
 * the float32 [batch][4][4] readouts (tile values, or exponents when obs_mode = 1). */
int r48_ring_sample(const r48_ring *ring, int64_t batch, uint64_t seed, uint64_t draw,
                    int with_replacement, int64_t *out_index, uint64_t *out_state,
                    uint8_t *out_action, int32_t *out_reward, uint64_t *out_next_state,
                    uint8_t *out_done, float *obs_state, float *obs_next_state, int obs_mode,
                    void *stream);

/* Measurement aid (bench.py): moves the single-step kernel's 22 bytes per board (reads boards +
 * actions, writes boards + reward + done) and computes nothing. */
int r48_debug_copy22(const uint64_t *in, const uint8_t *action, uint64_t *out, int32_t *reward,
                     uint8_t *done, int64_t n, void *stream);

/* ---- host-buffer entry points (what a CPU-side caller of the reference would bind) ----
 * All pointers are HOST pointers (pinned memory makes the copies asynchronous); the
 * library stages through a per-device scratch arena it owns, runs the kernels on its own
 * stream and returns after the results are in the host buffers. */
int r48_step_host(const uint64_t *in, const uint8_t *action, uint64_t *out, int32_t *reward,
                  uint8_t *done, int64_t n, uint64_t seed, uint64_t board_base, uint32_t step,
                  int reward_mode, int device);
int r48_afterstates_host(const uint64_t *in, uint64_t *out, int32_t *reward, uint8_t *valid,
                         uint8_t *done, int64_t n, int reward_mode, int device);
int r48_rollout_host(int64_t n, uint64_t seed, uint64_t board_base, uint64_t *final_boards,
                     uint32_t *lengths, uint64_t *stats, int device);
/* r48_rollout_host with the policy and the per-episode outputs selectable: any of final_boards
 * (8 B), lengths (4 B) and records (4 B) may be NULL and is then not computed / copied.
 * records[i] packs what main.py:48 prints per game and the episode length into one word:
 *     bits 31..13  score / 2   (score = np.sum(state_matrix); always even, at most 2^19: exact)
 *     bits 12..0   min(length, 8191) */
/* (Large batches are played in chunks whose copies overlap the next chunk's kernel; the chunk sizes
 * adapt to the copy / play time ratio the previous call measured on this device.  Results do not
 * depend on the chunking: every draw is keyed by (seed, board id, tick).) */
#define R48_RECORD_SCORE(r)  (((uint32_t)(r) >> 13) << 1)
#define R48_RECORD_LENGTH(r) ((uint32_t)(r) & 8191u)
int r48_rollout_host_ex(int64_t n, uint64_t seed, uint64_t board_base, int policy,
                        uint64_t *final_boards, uint32_t *lengths, uint32_t *records,
                        uint64_t *stats, int device);
/* Release the scratch arena and tables of every device. */
int r48_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* R48_H */
