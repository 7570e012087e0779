/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the torbi Viterbi decode path.
 *
 * This file is a plain-C restatement of the reference CPU operator
 *   /root/reference/torbi/csrc/viterbi.cpp
 * with 64-bit offsets (the reference indexes with `int`, viterbi.cpp:67-69,99,
 * and silently wraps once B*T*S >= 2^31).  It is the checker for the CUDA
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  Nothing under torbi_b200/ links, imports
 * or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * bit-for-bit against (a) the reference's only known-answer vector
 * (tests/test_core.py:7-25 -> [1, 2, 2]), (b) golden vectors produced by the
 * unmodified reference operator compiled into oracle/_ref (tests/golden/ npz files,
 * generator tests/golden/make_golden.py) and (c) oracle/_ref itself whenever
 * it is present.
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -shared -fPIC; NO -ffast-math:
 * the recursion is fl(obs + max_s fl(P[s] + A[j][s])) and must not be
 * re-associated).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------
 * One sequence, forward pass.
 *
 * viterbi.cpp:72-76   posterior_current[i] = obs[i] + initial[i]
 * viterbi.cpp:81-86   probability[j*S+s] = posterior_current[s] + transition[j*S+s]
 * viterbi.cpp:91-104  max_posterior = probability[j*S]; strict `>` scan over
 *                     s = 1..S-1, trellis[t*S+j] = s on improvement (the
 *                     trellis is pre-zeroed, viterbi.cpp:201-203, so an
 *                     un-improved row keeps source 0);
 *                     posterior_next[j] = obs[t*S+j] + max_posterior
 * viterbi.cpp:111-113 final posterior copied out
 *
 * `transition` is indexed [destination j][source s] (viterbi.cpp:83-84,97).
 * Row t = 0 of the trellis and rows t >= frames are left untouched, exactly as
 * in the reference.  `exact_scan` != 0 runs the literal sequential scan (NaN
 * behaviour included); 0 runs a two-pass form (max, then first index equal to
 * the max) that is bit-identical on NaN-free input and vectorises.
 * ---------------------------------------------------------------------- */
static void forward_one(const float *obs, int64_t frames, const float *transition,
                        const float *initial, int64_t S, float *posterior_out,
                        int32_t *trellis, float *cur, float *nxt, int exact_scan,
                        int parallel_states)
{
    for (int64_t i = 0; i < S; i++)
        cur[i] = obs[i] + initial[i];

    for (int64_t t = 1; t < frames; t++) {
        const float *obs_t = obs + t * S;
        int32_t *trellis_t = trellis + t * S;
#pragma omp parallel for schedule(static) if (parallel_states)
        for (int64_t j = 0; j < S; j++) {
            const float *row = transition + j * S;
            float best;
            int32_t arg = 0;
            if (exact_scan) {
                best = cur[0] + row[0];
                for (int64_t s = 1; s < S; s++) {
                    float v = cur[s] + row[s];
                    if (v > best) {
                        best = v;
                        arg = (int32_t)s;
                    }
                }
            } else {
                float m = cur[0] + row[0];
#pragma omp simd reduction(max : m)
                for (int64_t s = 1; s < S; s++) {
                    float v = cur[s] + row[s];
                    m = v > m ? v : m;
                }
                best = m;
                for (int64_t s = 0; s < S; s++) {
                    if (cur[s] + row[s] == m) {
                        arg = (int32_t)s;
                        break;
                    }
                }
            }
            /* the reference only writes on improvement over s = 0 */
            if (arg != 0)
                trellis_t[j] = arg;
            nxt[j] = obs_t[j] + best;
        }
        float *tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
    for (int64_t i = 0; i < S; i++)
        posterior_out[i] = cur[i];
}

/* viterbi.cpp:218 `posterior.argmax(1)`: first occurrence of the maximum
 * (ATen CPU semantics; NaN handling is not restated -- NaN input is outside
 * the parity contract). */
static int32_t argmax_first(const float *p, int64_t S)
{
    int32_t arg = 0;
    float best = p[0];
    for (int64_t s = 1; s < S; s++) {
        if (p[s] > best) {
            best = p[s];
            arg = (int32_t)s;
        }
    }
    return arg;
}

/* viterbi.cpp:140-160 with the row pre-filled by `repeat` (viterbi.cpp:219-221):
 * every position holds the final argmax, then positions frames-2..0 are
 * overwritten by the backtrace. */
static void backtrace_one(int32_t *indices, const int32_t *trellis, int64_t frames,
                          int64_t T, int64_t S, int32_t final_state)
{
    for (int64_t t = 0; t < T; t++)
        indices[t] = final_state;
    int32_t index = indices[frames - 1];
    for (int64_t t = frames - 1; t >= 1; t--) {
        index = trellis[t * S + index];
        indices[t - 1] = index;
    }
}

/* ------------------------------------------------------------------------
 * viterbi_decode_cpu (viterbi.cpp:182-234) for a whole batch.
 *
 *   observation  (B, T, S) fp32 log-probabilities, contiguous
 *   batch_frames (B,) int32, each in [1, T]
 *   transition   (S, S) fp32 log, [dest][src];  initial (S,) fp32 log
 *   indices_out  (B, T) int32
 *   posterior_out (B, S) fp32 or NULL     -- final forward scores
 *   trellis_out  (B, T, S) int32 or NULL  -- backpointers (rows 1..frames-1
 *                                            are meaningful), zero elsewhere
 *   exact_scan   1: literal reference scan; 0: vectorised two-pass
 *   num_threads  <=0: leave the OpenMP default
 *
 * Sequences are independent (viterbi.cpp:65 loops over b sequentially), so the
 * batch loop is what is parallelised when B >= threads; otherwise the state
 * loop is, as in the reference (viterbi.cpp:91).
 * Returns 0, or -1 on allocation failure / bad arguments.
 * ---------------------------------------------------------------------- */
ORACLE_API int torbi_oracle_decode(const float *observation, const int32_t *batch_frames,
                                   const float *transition, const float *initial,
                                   int64_t B, int64_t T, int64_t S, int32_t *indices_out,
                                   float *posterior_out, int32_t *trellis_out,
                                   int exact_scan, int num_threads)
{
    if (B <= 0 || T <= 0 || S <= 0)
        return -1;
    for (int64_t b = 0; b < B; b++)
        if (batch_frames[b] < 1 || batch_frames[b] > T)
            return -1;
#ifdef _OPENMP
    if (num_threads > 0)
        omp_set_num_threads(num_threads);
    int threads = omp_get_max_threads();
#else
    int threads = 1;
#endif
    int over_batch = (B >= threads) && threads > 1;
    int failed = 0;

#pragma omp parallel for schedule(dynamic, 1) if (over_batch)
    for (int64_t b = 0; b < B; b++) {
        int64_t frames = batch_frames[b];
        float *cur = (float *)malloc(sizeof(float) * (size_t)S);
        float *nxt = (float *)malloc(sizeof(float) * (size_t)S);
        float *post = (float *)malloc(sizeof(float) * (size_t)S);
        int32_t *trellis = trellis_out ? trellis_out + b * T * S
                                       : (int32_t *)calloc((size_t)(frames * S), sizeof(int32_t));
        if (!cur || !nxt || !post || !trellis) {
#pragma omp atomic write
            failed = 1;
        } else {
            if (trellis_out)
                memset(trellis, 0, sizeof(int32_t) * (size_t)(T * S));
            forward_one(observation + b * T * S, frames, transition, initial, S, post,
                        trellis, cur, nxt, exact_scan, !over_batch);
            if (posterior_out)
                memcpy(posterior_out + b * S, post, sizeof(float) * (size_t)S);
            backtrace_one(indices_out + b * T, trellis, frames, T, S, argmax_first(post, S));
        }
        free(cur);
        free(nxt);
        free(post);
        if (!trellis_out)
            free(trellis);
    }
    return failed ? -1 : 0;
}

ORACLE_API int torbi_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
