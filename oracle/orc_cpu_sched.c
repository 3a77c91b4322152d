/*
 * orc_cpu_sched.c -- ORACLE / CPU BASELINE (test + measurement infrastructure, never on the product path).
 *
 * A multi-threaded CPU port of the reference's host scheduling path for one window of tasks, used as the
 * "reference's own CPU scheduler on the same box" arm of the benchmark when the reference itself cannot
 * be built (its build needs CMake-generated headers, the MCA static-component table, ptgpp
 * output and hwloc).  It keeps the per-task and per-edge work items of the reference:
 *   worker loop            scheduling.c:727-860  __parsec_context_wait: select -> progress -> loop, nanosleep back-off
 *   task progress          scheduling.c:507-575  prepare_input, __parsec_execute (:126-206), complete
 *   completion             scheduling.c:469-505  __parsec_complete_execution -> release_deps -> release_task
 *   per out-edge           parsec.c:1836-1975 parsec_release_dep_fct -> :1749-1834 release_local_OUT_dependencies:
 *                          find_deps in a hash table under a bucket lock (:1580-1607, entry allocated from a
 *                          per-thread mempool on first touch), atomic update_deps (counter :1609 / mask :1656),
 *                          when ready allocate the successor parsec_task_t from the thread mempool, copy the
 *                          execution context, zero data[], chain into the ready ring (:1784-1806)
 *   scheduling             scheduling.c:286-360 __parsec_schedule; mca/sched/lfq/sched_lfq_module.c:166,196:
 *                          per-thread bounded local queue, overflow into one shared system dequeue,
 *                          select = own queue, then steal from the other threads' queues, then the system queue
 *   termination            mca/termdet/local: atomic count of pending tasks
 * Bodies are the oracle's CPU bodies (orc_dag.c semantics) on host memory.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sched.h>

#include "../include/pb2_engine.h"

/* a host-side task object the size of the reference's parsec_task_t (parsec_internal.h:551-563) */
typedef struct cpu_task_s {
    struct cpu_task_s* next;
    int32_t id;
    int32_t priority;
    int32_t locals[20];
    struct { void* repo; void* entry; void* data_in; void* data_out; } data[20];
} cpu_task_t;

typedef struct dep_entry_s { struct dep_entry_s* next; uint64_t key; _Atomic int32_t dep; } dep_entry_t;

typedef struct { atomic_flag lock; dep_entry_t* head; char pad[48]; } bucket_t;

#define LOCALQ 64            /* sched_lfq: 4 * nb_cores entries per local hbbuffer, bounded here */

typedef struct worker_s {
    pthread_t thread;
    int th_id;
    struct sched_s* s;
    /* local flat queue (lfq) */
    atomic_flag qlock;
    cpu_task_t* q[LOCALQ];
    int qn;
    /* per-thread mempools (mempool.c): free lists, never returned to the OS during the run */
    cpu_task_t* task_pool;
    dep_entry_t* dep_pool;
    uint64_t executed;
    char pad[64];
} worker_t;

typedef struct sched_s {
    const pb2_task_t* tasks; const uint32_t* succ; pb2_tile_t* tiles;
    int32_t ntasks; int nthreads;
    bucket_t* buckets; uint32_t nbuckets;
    atomic_flag syslock; cpu_task_t* sys_head; cpu_task_t* sys_tail;      /* system dequeue */
    _Atomic int32_t pending;                                              /* termdet */
    _Atomic uint64_t body_errors;
    worker_t* w;
    pthread_barrier_t barrier;
} sched_t;

extern uint64_t orc_cpu_body(const pb2_task_t* t, void* flow[PB2_MAX_FLOWS], const uint32_t bytes[PB2_MAX_FLOWS]);

static inline void spin_lock(atomic_flag* f) { while (atomic_flag_test_and_set_explicit(f, memory_order_acquire)) { } }
static inline void spin_unlock(atomic_flag* f) { atomic_flag_clear_explicit(f, memory_order_release); }

static cpu_task_t* task_alloc(worker_t* w) {
    cpu_task_t* t = w->task_pool;
    if (t) { w->task_pool = t->next; return t; }
    return (cpu_task_t*)malloc(sizeof(cpu_task_t));
}
static void task_free(worker_t* w, cpu_task_t* t) { t->next = w->task_pool; w->task_pool = t; }

/* parsec_hash_find_deps: lookup-or-insert under the bucket lock */
static _Atomic int32_t* find_deps(worker_t* w, int32_t id) {
    sched_t* s = w->s;
    const uint64_t key = (uint64_t)id * 0x9E3779B97F4A7C15ull;
    bucket_t* b = &s->buckets[(key >> 32) & (s->nbuckets - 1)];
    spin_lock(&b->lock);
    dep_entry_t* e = b->head;
    while (e && e->key != (uint64_t)id) e = e->next;
    if (!e) {
        e = w->dep_pool;
        if (e) w->dep_pool = e->next; else e = (dep_entry_t*)malloc(sizeof *e);
        e->key = (uint64_t)id;
        atomic_store_explicit(&e->dep, 0, memory_order_relaxed);
        e->next = b->head; b->head = e;
    }
    spin_unlock(&b->lock);
    return &e->dep;
}

static void schedule(worker_t* w, cpu_task_t* ring) {       /* __parsec_schedule + sched_lfq schedule */
    sched_t* s = w->s;
    while (ring) {
        cpu_task_t* t = ring; ring = ring->next; t->next = NULL;
        spin_lock(&w->qlock);
        if (w->qn < LOCALQ) { w->q[w->qn++] = t; spin_unlock(&w->qlock); continue; }
        spin_unlock(&w->qlock);
        spin_lock(&s->syslock);
        if (s->sys_tail) s->sys_tail->next = t; else s->sys_head = t;
        s->sys_tail = t;
        spin_unlock(&s->syslock);
    }
}

static cpu_task_t* select_task(worker_t* w) {               /* sched_lfq select */
    sched_t* s = w->s;
    for (int d = 0; d < s->nthreads; ++d) {
        worker_t* v = &s->w[(w->th_id + d) % s->nthreads];
        if (v->qn == 0) continue;
        spin_lock(&v->qlock);
        cpu_task_t* t = v->qn ? v->q[--v->qn] : NULL;
        spin_unlock(&v->qlock);
        if (t) return t;
    }
    if (s->sys_head) {
        spin_lock(&s->syslock);
        cpu_task_t* t = s->sys_head;
        if (t) { s->sys_head = t->next; if (!s->sys_head) s->sys_tail = NULL; t->next = NULL; }
        spin_unlock(&s->syslock);
        return t;
    }
    return NULL;
}

static cpu_task_t* make_task(worker_t* w, int32_t id) {     /* parsec.c:1784-1806 */
    const pb2_task_t* d = &w->s->tasks[id];
    cpu_task_t* t = task_alloc(w);
    t->id = id; t->priority = d->priority; t->next = NULL;
    t->locals[0] = d->locals[0]; t->locals[1] = d->locals[1];
    memset(t->data, 0, sizeof(t->data[0]) * (d->nb_flows ? d->nb_flows : 1));
    return t;
}

static void progress(worker_t* w, cpu_task_t* ct) {
    sched_t* s = w->s;
    const pb2_task_t* t = &s->tasks[ct->id];
    void* flow[PB2_MAX_FLOWS] = {0};
    uint32_t bytes[PB2_MAX_FLOWS] = {0};
    for (int f = 0; f < t->nb_flows; ++f) {                  /* prepare_input */
        if (t->tile[f] < 0) continue;
        flow[f] = s->tiles[t->tile[f]].dev_ptr; bytes[f] = s->tiles[t->tile[f]].bytes;
        ct->data[f].data_in = flow[f]; ct->data[f].data_out = flow[f];
    }
    const uint64_t r = orc_cpu_body(t, flow, bytes);          /* __parsec_execute -> hook */
    if ((t->body == PB2_BODY_CHECK_I32 || t->body == PB2_BODY_CHECK_F32) && (r >> 32))
        atomic_fetch_add(&s->body_errors, r >> 32);
    /* __parsec_complete_execution -> release_deps -> iterate_successors(parsec_release_dep_fct) */
    cpu_task_t* ring = NULL; cpu_task_t* ring_tail = NULL;
    for (int32_t e = 0; e < t->succ_count; ++e) {
        const uint32_t sc = s->succ[t->succ_begin + e];
        const int32_t sid = PB2_SUCC_TASK(sc);
        const pb2_task_t* st = &s->tasks[sid];
        _Atomic int32_t* dep = find_deps(w, sid);
        int ready;
        if (st->flags & PB2_TASK_DEPS_MASK) {
            const int32_t bit = 1 << PB2_SUCC_FLOW(sc);
            const int32_t cur = atomic_fetch_or(dep, bit) | bit;
            ready = (cur & st->dep_goal) == st->dep_goal;
        } else {
            /* lazily install the goal on first touch, then fetch_dec (parsec.c:1625-1635) */
            int32_t zero = 0, cur;
            if (atomic_load_explicit(dep, memory_order_relaxed) == 0 &&
                atomic_compare_exchange_strong(dep, &zero, st->dep_goal - 1)) cur = st->dep_goal - 1;
            else cur = atomic_fetch_sub(dep, 1) - 1;
            ready = (cur == 0);
        }
        if (ready) {
            cpu_task_t* nt = make_task(w, sid);
            if (ring_tail) ring_tail->next = nt; else ring = nt;
            ring_tail = nt;
        }
    }
    if (ring) schedule(w, ring);
    task_free(w, ct);
    w->executed++;
    atomic_fetch_sub_explicit(&s->pending, 1, memory_order_release);
}

static void* worker_main(void* arg) {
    worker_t* w = (worker_t*)arg;
    sched_t* s = w->s;
    pthread_barrier_wait(&s->barrier);
    int misses = 0;
    while (atomic_load_explicit(&s->pending, memory_order_acquire) > 0) {
        cpu_task_t* t = select_task(w);
        if (t) { misses = 0; progress(w, t); continue; }
        if (++misses > 64) { struct timespec ts = {0, 100}; nanosleep(&ts, NULL); }   /* scheduling.c:846 back-off */
        else sched_yield();
    }
    return NULL;
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/*
 * Execute the window on nthreads host threads.  tiles[].dev_ptr are host pointers (all tiles resident).
 * Returns wall seconds of the run (context_start -> context_wait), <0 on error.  executed_per_thread may be NULL.
 */
double orc_cpu_sched_run(const pb2_task_t* tasks, int32_t ntasks, const uint32_t* succ, int32_t nsucc,
                         pb2_tile_t* tiles, int32_t ntiles, const int32_t* ready, int32_t nready,
                         int nthreads, uint64_t* executed_per_thread, uint64_t* body_errors) {
    (void)nsucc; (void)ntiles;
    if (nthreads < 1) return -1.0;
    sched_t s; memset(&s, 0, sizeof s);
    s.tasks = tasks; s.succ = succ; s.tiles = tiles; s.ntasks = ntasks; s.nthreads = nthreads;
    s.nbuckets = 1; while (s.nbuckets < (uint32_t)(ntasks > 1024 ? ntasks : 1024)) s.nbuckets <<= 1;
    s.buckets = (bucket_t*)calloc(s.nbuckets, sizeof(bucket_t));
    for (uint32_t i = 0; i < s.nbuckets; ++i) atomic_flag_clear(&s.buckets[i].lock);
    atomic_flag_clear(&s.syslock);
    atomic_store(&s.pending, ntasks);
    s.w = (worker_t*)calloc((size_t)nthreads, sizeof(worker_t));
    pthread_barrier_init(&s.barrier, NULL, (unsigned)nthreads + 1);
    for (int i = 0; i < nthreads; ++i) { s.w[i].th_id = i; s.w[i].s = &s; atomic_flag_clear(&s.w[i].qlock); }
    /* startup tasks are distributed round-robin over the workers' queues (scheduling.c:922 startup) */
    for (int32_t i = 0; i < nready; ++i) schedule(&s.w[i % nthreads], make_task(&s.w[i % nthreads], ready[i]));
    for (int i = 0; i < nthreads; ++i) pthread_create(&s.w[i].thread, NULL, worker_main, &s.w[i]);
    const double t0 = now_s();
    pthread_barrier_wait(&s.barrier);
    for (int i = 0; i < nthreads; ++i) pthread_join(s.w[i].thread, NULL);
    const double t1 = now_s();
    uint64_t total = 0;
    for (int i = 0; i < nthreads; ++i) {
        total += s.w[i].executed;
        if (executed_per_thread) executed_per_thread[i] = s.w[i].executed;
        for (cpu_task_t* t = s.w[i].task_pool; t;) { cpu_task_t* n = t->next; free(t); t = n; }
        for (dep_entry_t* d = s.w[i].dep_pool; d;) { dep_entry_t* n = d->next; free(d); d = n; }
    }
    for (uint32_t i = 0; i < s.nbuckets; ++i)
        for (dep_entry_t* d = s.buckets[i].head; d;) { dep_entry_t* n = d->next; free(d); d = n; }
    if (body_errors) *body_errors = atomic_load(&s.body_errors);
    free(s.buckets); free(s.w);
    pthread_barrier_destroy(&s.barrier);
    return (total == (uint64_t)ntasks) ? (t1 - t0) : -2.0;
}
