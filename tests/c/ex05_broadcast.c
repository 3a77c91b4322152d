/*
 * ex05_broadcast.c -- the reference's examples/Ex05_Broadcast.jdf main (examples/Ex05_Broadcast.jdf:60-133) written
 * against this repository's C ABI instead of libparsec: init, one GPU module, a 2D block-cyclic collection over user
 * memory, the Ex05 PTG pool, wait, check that every TaskRecv(k, n) observed k (:53-57), print the device statistics.
 * Plain C99, links only libparsec_b200.so.  Usage: ex05_broadcast <nodes K> <NB> <tile bytes> <dry_run 0|1>
 * (dry_run = 1 builds and retires the windows without a GPU: what the CPU-only test runs).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pb2_parsec.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != PB2_SUCCESS) { fprintf(stderr, "%s failed: %d\n", #x, rc_); return 2; } } while (0)

int main(int argc, char** argv)
{
    const int K = argc > 1 ? atoi(argv[1]) : 16, NB = argc > 2 ? atoi(argv[2]) : 14;
    const int tile_bytes = argc > 3 ? atoi(argv[3]) : 4096, dry_run = argc > 4 ? atoi(argv[4]) : 0;
    const int F = NB / 2 + 1, elems = tile_bytes / 4;
    pb2_context_t* ctx = NULL;
    pb2_device_module_t* gpu = NULL;
    CHECK(pb2_init(&ctx, 1));
    CHECK(pb2_device_cuda_module_init(ctx, 0, dry_run, &gpu));
    CHECK(pb2_mca_device_registration_complete(ctx));

    int32_t* mydata = (int32_t*)malloc((size_t)K * tile_bytes);
    memset(mydata, 0xff, (size_t)K * tile_bytes);
    /* mydata(k): K tiles of `elems` ints in one row, one rank (Ex05 uses a vector of ints; the tile is the int) */
    pb2_data_collection_t* dc = pb2_matrix_block_cyclic_new(ctx, 4, 0, elems, 1, K * elems, 1, 0, 0, K * elems, 1, 1, 1, 1, 1, 0, 0);
    if (!dc) { fprintf(stderr, "collection\n"); return 2; }
    CHECK(pb2_data_collection_set_mat(dc, mydata));
    if (!dry_run) CHECK(pb2_dc_register_memory(dc, gpu));

    pb2_taskpool_t* tp = pb2_ptg_ex05_broadcast_new(ctx, dc, K, NB);
    if (!tp) { fprintf(stderr, "taskpool\n"); return 2; }
    CHECK(pb2_context_start(ctx));
    CHECK(pb2_context_wait(ctx));

    const int n = pb2_taskpool_nb_tasks(tp);
    int32_t* cls = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* loc = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)n);
    uint64_t* res = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    CHECK(pb2_taskpool_task_info(tp, cls, loc, NULL, res));
    int bad = 0, recv = 0;
    for (int i = 0; i < n; ++i) {
        if (cls[i] != 1) continue;                              /* TaskRecv */
        ++recv;
        if (!dry_run && res[i] != (uint64_t)(uint32_t)loc[2 * i]) ++bad;   /* observed value == k, no mismatching element */
    }
    pb2_device_stats_t st;
    CHECK(pb2_device_get_stats(gpu, &st));
    printf("tasks %d receivers %d bad %d executed %llu windows %llu released_on_device %llu h2d %llu\n", n, recv, bad,
           (unsigned long long)st.executed_tasks, (unsigned long long)st.windows_launched,
           (unsigned long long)st.tasks_released_on_device, (unsigned long long)st.data_in_from_device[0]);
    const int ok = n == K * (1 + F) && recv == K * F && bad == 0 && st.executed_tasks == (uint64_t)n &&
                   st.tasks_released_on_device == (uint64_t)(K * F) && st.data_in_from_device[0] == (uint64_t)K * tile_bytes;
    CHECK(pb2_taskpool_free(tp));
    CHECK(pb2_data_collection_free(dc));
    CHECK(pb2_fini(&ctx));
    free(cls); free(loc); free(res); free(mydata);
    puts(ok ? "PASS" : "FAIL");
    return ok ? 0 : 1;
}
