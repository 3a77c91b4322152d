/*
 * orc_dag.c -- ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Sequential CPU restatement of how the reference executes one window of GPU tasks, used as the
 * checker for the CUDA engine: same task/tile arrays in, every integer output compared bit-exact.
 *
 * Reference semantics restated here (file:line in /root/reference):
 *   - ready tasks are consumed by one worker in FIFO order of readiness
 *     (scheduling.c:789-818 worker loop with a single execution stream);
 *   - stage-in: a flow with READ access whose device copy is INVALID is copied from its source
 *     exactly once and charged to the statistics (device_gpu.c:1799-2165, :2130-2136); a WRITE-only
 *     flow is not transferred ("finally we'll just overwrite w/o read", data.c:427);
 *   - versions: the task sees the input copy's version; a WRITE flow leaves version+1
 *     (device_gpu.c:2148-2152);
 *   - pushout flows are copied back to the home copy after the body (device_gpu.c:2943-3173);
 *   - completion releases each out-edge with update_deps_with_counter (parsec.c:1609-1654: fetch_dec,
 *     ready at 0) or update_deps_with_mask (parsec.c:1656-1720: OR the destination flow bit, ready when
 *     (word & goal) == goal), newly ready tasks are appended in iterate_successors order
 *     (parsec.c:1749-1834).
 * Bodies restate the reference's toy kernels / example bodies (cited in include/pb2_engine.h).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "../include/pb2_engine.h"

typedef struct orc_stats_s {
    uint64_t tasks_retired, bytes_h2d, bytes_d2d, bytes_d2h, stage_ins, body_errors;
} orc_stats_t;

static inline float bf16_to_f32(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
static inline uint16_t f32_to_bf16(float f) {   /* round to nearest even, cvt.rn.bf16.f32 */
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

/* C(M x N, row-major bf16) = bf16( f32(C) + A(M x K row-major) * B^T (B stored N x K) ), fp32 accumulate */
static void body_gemm_bf16(const uint16_t* A, const uint16_t* B, uint16_t* C, int M, int N, int K) {
    float* a = (float*)malloc(sizeof(float) * (size_t)K);
    for (int i = 0; i < M; ++i) {
        for (int k = 0; k < K; ++k) a[k] = bf16_to_f32(A[(size_t)i * K + k]);
        for (int j = 0; j < N; ++j) {
            const uint16_t* b = B + (size_t)j * K;
            float acc = 0.0f;
            for (int k = 0; k < K; ++k) acc += a[k] * bf16_to_f32(b[k]);
            C[(size_t)i * N + j] = f32_to_bf16(bf16_to_f32(C[(size_t)i * N + j]) + acc);
        }
    }
    free(a);
}

static uint64_t run_body(const pb2_task_t* t, void* flow[PB2_MAX_FLOWS], const uint32_t bytes[PB2_MAX_FLOWS]) {
    const size_t n0 = bytes[0] / 4;
    int32_t* i0 = (int32_t*)flow[0];
    float* f0 = (float*)flow[0];
    switch (t->body) {
    case PB2_BODY_NOP: return 0;
    case PB2_BODY_FILL_I32: for (size_t i = 0; i < n0; ++i) i0[i] = t->iparam[0]; return 0;
    case PB2_BODY_FILL_F32: for (size_t i = 0; i < n0; ++i) f0[i] = t->fparam; return 0;
    case PB2_BODY_MEMSET_U8: memset(flow[0], t->iparam[0] & 0xff, bytes[0]); return 0;
    case PB2_BODY_CHECK_I32: {
        uint64_t bad = 0;
        for (size_t i = 0; i < n0; ++i) bad += (i0[i] != t->iparam[0]);
        return (bad << 32) | (n0 ? (uint32_t)i0[0] : 0u);
    }
    case PB2_BODY_CHECK_F32: {
        uint64_t bad = 0; uint32_t k, v; memcpy(&k, &t->fparam, 4);
        for (size_t i = 0; i < n0; ++i) { memcpy(&v, &f0[i], 4); bad += (v != k); }
        if (n0) memcpy(&v, &f0[0], 4); else v = 0;
        return (bad << 32) | v;
    }
    case PB2_BODY_INCR_I32: for (size_t i = 0; i < n0; ++i) i0[i] = (int32_t)((uint32_t)i0[i] + (uint32_t)t->iparam[0]); return 0;
    case PB2_BODY_SCALE_I32: for (size_t i = 0; i < n0; ++i) i0[i] = (int32_t)((uint32_t)i0[i] * (uint32_t)t->iparam[0]); return 0;
    case PB2_BODY_ADD_IOTA_I32: for (size_t i = 0; i < n0; ++i) i0[i] = (int32_t)((uint32_t)i0[i] + (uint32_t)i); return 0;
    case PB2_BODY_IOTA_I32: for (size_t i = 0; i < n0; ++i) i0[i] = (int32_t)i; return 0;
    case PB2_BODY_INCR_F32: for (size_t i = 0; i < n0; ++i) f0[i] = f0[i] + t->fparam; return 0;
    case PB2_BODY_ADD_AT_I32:
        if (t->iparam[0] >= 0 && (size_t)t->iparam[0] < n0) i0[t->iparam[0]] = (int32_t)((uint32_t)i0[t->iparam[0]] + (uint32_t)t->iparam[1]);
        return 0;
    case PB2_BODY_COPY: memcpy(flow[1], flow[0], bytes[0] < bytes[1] ? bytes[0] : bytes[1]); return 0;
    case PB2_BODY_AXPY_F32: {
        const size_t n = (bytes[0] < bytes[1] ? bytes[0] : bytes[1]) / 4;
        float* y = (float*)flow[1];
        for (size_t i = 0; i < n; ++i) y[i] = fmaf(t->fparam, f0[i], y[i]);
        return 0;
    }
    case PB2_BODY_GEMM_BF16:
        body_gemm_bf16((const uint16_t*)flow[0], (const uint16_t*)flow[1], (uint16_t*)flow[2],
                       t->iparam[0], t->iparam[1], t->iparam[2]);
        return 0;
    default: return ~0ull;
    }
}

/*
 * Run the window.  tiles[i].dev_ptr / src_ptr are HOST pointers here (the oracle's "device" is malloc'ed
 * memory owned by the caller).  Outputs mirror pb2_window_results.  Returns 0, or -1 if the DAG deadlocks
 * (tasks left with unsatisfied dependencies), -2 on an unknown body.
 */
/* Order in which ready tasks are picked: 0 FIFO (the engine's order with one worker), 1 LIFO, 2 seeded random.
 * Any order is a legal execution of the DAG; the non-FIFO ones are used by the tests to shake out missing
 * write-after-read edges (a stage-in reads its source when the task runs, not when it was released). */
static int g_policy = 0;
static uint32_t g_rng = 1;
void orc_set_policy(int policy, uint32_t seed) { g_policy = policy; g_rng = seed ? seed : 1; }

int orc_run_window(const pb2_task_t* tasks, int32_t ntasks, const uint32_t* succ, int32_t nsucc,
                   pb2_tile_t* tiles, int32_t ntiles, const int32_t* ready, int32_t nready,
                   int32_t* retire_order, uint32_t* start_seq, uint32_t* end_seq, uint32_t* seen_version,
                   uint64_t* result, orc_stats_t* stats) {
    (void)nsucc; (void)ntiles;
    int32_t* dep = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1));
    int32_t* fifo = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1));
    int32_t head = 0, tail = 0;
    uint32_t evt = 0;
    int rc = 0;
    memset(stats, 0, sizeof *stats);
    for (int32_t i = 0; i < ntasks; ++i) dep[i] = (tasks[i].flags & PB2_TASK_DEPS_MASK) ? 0 : tasks[i].dep_goal;
    for (int32_t i = 0; i < nready; ++i) fifo[tail++] = ready[i];
    while (head < tail) {
        int32_t id;
        if (g_policy == 1) id = fifo[--tail];
        else {
            if (g_policy == 2) {
                g_rng ^= g_rng << 13; g_rng ^= g_rng >> 17; g_rng ^= g_rng << 5;
                const int32_t j = head + (int32_t)(g_rng % (uint32_t)(tail - head));
                const int32_t tmp = fifo[head]; fifo[head] = fifo[j]; fifo[j] = tmp;
            }
            id = fifo[head++];
        }
        const pb2_task_t* t = &tasks[id];
        void* flow[PB2_MAX_FLOWS] = {0};
        uint32_t bytes[PB2_MAX_FLOWS] = {0};
        if (start_seq) start_seq[id] = evt;
        evt++;
        for (int f = 0; f < t->nb_flows; ++f) {
            if (t->tile[f] < 0) continue;
            pb2_tile_t* tile = &tiles[t->tile[f]];
            if ((t->access[f] & PB2_FLOW_ACCESS_READ) && tile->state != PB2_TILE_VALID) {
                memcpy(tile->dev_ptr, tile->src_ptr, tile->bytes);
                tile->state = PB2_TILE_VALID;
                if (tile->src_kind == PB2_SRC_PEER) stats->bytes_d2d += tile->bytes; else stats->bytes_h2d += tile->bytes;
                stats->stage_ins++;
            }
            flow[f] = tile->dev_ptr; bytes[f] = tile->bytes;
            if (seen_version) seen_version[(size_t)id * PB2_MAX_FLOWS + f] = tile->version;
        }
        const uint64_t r = run_body(t, flow, bytes);
        if (r == ~0ull) { rc = -2; break; }
        if (result) result[id] = r;
        if ((t->body == PB2_BODY_CHECK_I32 || t->body == PB2_BODY_CHECK_F32)) stats->body_errors += r >> 32;
        for (int f = 0; f < t->nb_flows; ++f) {
            if (t->tile[f] < 0 || !(t->access[f] & PB2_FLOW_ACCESS_WRITE)) continue;
            pb2_tile_t* tile = &tiles[t->tile[f]];
            if (t->access[f] & PB2_FLOW_PUSHOUT) { memcpy(tile->src_ptr, tile->dev_ptr, tile->bytes); stats->bytes_d2h += tile->bytes; }
            tile->version++;
            tile->state = PB2_TILE_VALID;
        }
        if (end_seq) end_seq[id] = evt;
        evt++;
        if (retire_order) retire_order[stats->tasks_retired] = id;
        stats->tasks_retired++;
        for (int32_t e = 0; e < t->succ_count; ++e) {
            const uint32_t s = succ[t->succ_begin + e];
            const int32_t sid = PB2_SUCC_TASK(s);
            const pb2_task_t* st = &tasks[sid];
            int now_ready;
            if (st->flags & PB2_TASK_DEPS_MASK) {
                const int32_t bit = 1 << PB2_SUCC_FLOW(s);
                const int32_t old = dep[sid];
                dep[sid] = old | bit;
                now_ready = ((dep[sid] & st->dep_goal) == st->dep_goal) && ((old & st->dep_goal) != st->dep_goal);
            } else {
                now_ready = (--dep[sid] == 0);
            }
            if (now_ready) fifo[tail++] = sid;
        }
    }
    if (rc == 0 && (int32_t)stats->tasks_retired != ntasks) rc = -1;
    free(dep); free(fifo);
    return rc;
}

/* the CPU bodies, shared with the multi-threaded CPU baseline (orc_cpu_sched.c) */
uint64_t orc_cpu_body(const pb2_task_t* t, void* flow[PB2_MAX_FLOWS], const uint32_t bytes[PB2_MAX_FLOWS]) {
    return run_body(t, flow, bytes);
}
