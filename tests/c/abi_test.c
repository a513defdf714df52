/* A plain C caller of libr48.so: no Python, no torch -- only include/r48.h and host buffers.
 * Compares the _host entry points with the CPU oracle (linked as a second shared library).
 * Built and run by tests/test_c_abi.py on the GPU box. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "r48.h"

/* oracle/r48_oracle.c */
void orc_rollout(int64_t n, uint64_t seed, uint64_t board_base, uint64_t *final_boards, uint32_t *lengths);
void orc_episode_stats(const uint64_t *final_boards, const uint32_t *lengths, int64_t n, uint64_t *stats);
void orc_reset_batch(uint64_t *boards, int64_t n, uint64_t seed, uint64_t board_base);
int orc_step_batch(const uint64_t *in, const uint8_t *action, uint64_t *out, int32_t *reward, uint8_t *done,
                   int64_t n, uint64_t seed, uint64_t board_base, uint32_t step, int reward_mode);
void orc_afterstates_batch(const uint64_t *in, uint64_t *out, int32_t *reward, uint8_t *valid, uint8_t *done,
                           int64_t n, int reward_mode);

/* two host threads on one device: the _host entry points serialise on a per-device mutex */
struct job { int64_t n; uint64_t seed, base; uint64_t *fb; uint32_t *ln; uint32_t *rec; int rc; };

static void *job_main(void *arg)
{
    struct job *j = (struct job *)arg;
    j->rc = 0;
    for (int rep = 0; rep < 3 && j->rc == 0; rep++)      /* different sizes make the arena grow under the other thread */
        j->rc = r48_rollout_host_ex(j->n >> (2 - rep), j->seed, j->base, R48_POLICY_RANDOM, j->fb, j->ln, j->rec, NULL, 0);
    return NULL;
}

#define CHECK(cond, msg) do { if (!(cond)) { fprintf(stderr, "FAIL: %s (%s)\n", msg, r48_last_error()); return 1; } } while (0)

int main(void)
{
    const int64_t n = 5000;
    const uint64_t seed = 2048, base = 123456789012345ull;
    CHECK(r48_version() == R48_VERSION, "version");

    /* main.play(control="rand") x n */
    uint64_t *fb = malloc(n * 8), *ofb = malloc(n * 8);
    uint32_t *ln = malloc(n * 4), *oln = malloc(n * 4);
    uint64_t *st = calloc(R48_STATS_WORDS, 8), *ost = calloc(R48_STATS_WORDS, 8);
    CHECK(r48_rollout_host(n, seed, base, fb, ln, st, 0) == R48_OK, "r48_rollout_host");
    orc_rollout(n, seed, base, ofb, oln);
    orc_episode_stats(ofb, oln, n, ost);
    CHECK(memcmp(fb, ofb, n * 8) == 0, "final boards differ");
    CHECK(memcmp(ln, oln, n * 4) == 0, "lengths differ");
    CHECK(memcmp(st, ost, R48_STATS_WORDS * 8) == 0, "statistics differ");

    /* Game.reset (oracle) then Game.step x n through host buffers, three steps */
    uint64_t *b = malloc(n * 8), *nb = malloc(n * 8), *onb = malloc(n * 8);
    uint8_t *a = malloc(n), *dn = malloc(n), *odn = malloc(n);
    int32_t *rw = malloc(n * 4), *orw = malloc(n * 4);
    orc_reset_batch(b, n, seed, base);
    for (uint32_t step = 0; step < 3; step++) {
        for (int64_t i = 0; i < n; i++) a[i] = (uint8_t)((i * 7 + step * 3) & 3);
        CHECK(r48_step_host(b, a, nb, rw, dn, n, seed, base, step, 1, 0) == R48_OK, "r48_step_host");
        CHECK(orc_step_batch(b, a, onb, orw, odn, n, seed, base, step, 1) == 0, "orc_step_batch");
        CHECK(memcmp(nb, onb, n * 8) == 0 && memcmp(rw, orw, n * 4) == 0 && memcmp(dn, odn, n) == 0, "step differs");
        memcpy(b, nb, n * 8);
    }
    a[17] = 9;                                        /* GameClient.py:254: ValueError */
    CHECK(r48_step_host(b, a, nb, rw, dn, n, seed, base, 3, 0, 0) == R48_ERR_ACTION, "bad action not reported");

    /* 4 x update_matrix + has_game_over */
    uint64_t *af = malloc(n * 32), *oaf = malloc(n * 32);
    int32_t *ar = malloc(n * 16), *oar = malloc(n * 16);
    uint8_t *va = malloc(n), *ova = malloc(n);
    CHECK(r48_afterstates_host(b, af, ar, va, dn, n, 1, 0) == R48_OK, "r48_afterstates_host");
    orc_afterstates_batch(b, oaf, oar, ova, odn, n, 1);
    CHECK(memcmp(af, oaf, n * 32) == 0 && memcmp(ar, oar, n * 16) == 0 && memcmp(va, ova, n) == 0 &&
          memcmp(dn, odn, n) == 0, "afterstates differ");

    /* r48.h "Threads": concurrent _host calls on one device are safe */
    struct job jobs[2];
    pthread_t tid[2];
    for (int t = 0; t < 2; t++) {
        jobs[t].n = 40000 + 20000 * t; jobs[t].seed = 7 + t; jobs[t].base = 1000 * t;
        jobs[t].fb = malloc(jobs[t].n * 8); jobs[t].ln = malloc(jobs[t].n * 4); jobs[t].rec = malloc(jobs[t].n * 4);
        pthread_create(&tid[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < 2; t++) pthread_join(tid[t], NULL);
    for (int t = 0; t < 2; t++) {
        CHECK(jobs[t].rc == R48_OK, "threaded r48_rollout_host_ex");
        uint64_t *tfb = malloc(jobs[t].n * 8);
        uint32_t *tln = malloc(jobs[t].n * 4);
        orc_rollout(jobs[t].n, jobs[t].seed, jobs[t].base, tfb, tln);
        CHECK(memcmp(jobs[t].fb, tfb, jobs[t].n * 8) == 0 && memcmp(jobs[t].ln, tln, jobs[t].n * 4) == 0,
              "threaded rollouts differ from the oracle");
        for (int64_t i = 0; i < jobs[t].n; i += 997)
            CHECK(R48_RECORD_LENGTH(jobs[t].rec[i]) == (tln[i] < 8191 ? tln[i] : 8191), "record length");
        free(tfb); free(tln);
    }

    CHECK(r48_rollout_host(-1, 0, 0, fb, ln, st, 0) == R48_ERR_ARG, "negative n accepted");
    CHECK(r48_shutdown() == R48_OK, "shutdown");
    printf("C ABI ok: %lld episodes, %llu env-steps, step/afterstates bit-exact, 2 host threads ok\n", (long long)n,
           (unsigned long long)st[R48_STATS_SUM_LEN]);
    return 0;
}
