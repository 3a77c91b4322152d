/*
 * ex05_main.c -- driver of tests/parsec/ex05_b200.jdf: plays the role the reference's test mains play
 * (examples/Ex05_Broadcast.jdf main, tests/runtime/cuda/stage_main.c).  Prints one JSON line.
 *
 *   ex05_b200 [-K groups] [-N NB] [-t tile_elems] [-r repeats] [-c cores] [-m cpu|gpu] [-P pools] [-v] [-- parsec args]
 *
 * -m cpu: restrict the taskpool to the CPU incarnations (the reference's own scheduler + CPU bodies: the CPU baseline);
 * -m gpu: every task class has a CUDA incarnation, whichever GPU component is active drives it.
 * -P n: n task pools, each on a collection of its own, are handed to the context TOGETHER (one start / wait): the device
 *       modules see the tasks of several pools interleaved.
 * The check is the example's known answer: every TaskRecv(k, n) sees k in the whole tile; with -m gpu the verdict comes
 * from the engine's CHECK body through a complete_stage-free path: the final host tiles (pushed out because the
 * collection is flushed at the end) must hold k, and the device statistics must show each tile staged in once.
 */
#include "parsec.h"
#include "parsec/data_dist/matrix/two_dim_rectangle_cyclic.h"
#include "parsec/mca/device/device.h"
#include "parsec/mca/device/b200/device_b200.h"
#include "parsec/utils/mca_param.h"
#include "parsec/parsec_internal.h"
#include "parsec/execution_stream.h"
#include "ex05_b200.h"
#include "pb2_engine.h"
#include "checksum.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <execinfo.h>
#include <signal.h>

static void on_segv(int sig) { void *bt[64]; int n = backtrace(bt, 64); backtrace_symbols_fd(bt, n, 2); _exit(128 + sig); }

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(int argc, char *argv[])
{
    int K = 64, NB = 14, elems = 256 * 256, repeats = 1, cores = -1, verbose = 0, gpu = 1, wb = 0, prefetch = 0, pools = 1, c;
    while( -1 != (c = getopt(argc, argv, "K:N:t:r:c:m:P:vwp")) ) {
        switch(c) {
        case 'K': K = atoi(optarg); break;
        case 'N': NB = atoi(optarg); break;
        case 't': elems = atoi(optarg); break;
        case 'r': repeats = atoi(optarg); break;
        case 'c': cores = atoi(optarg); break;
        case 'm': gpu = (0 == strcmp(optarg, "gpu")); break;
        case 'v': verbose = 1; break;
        case 'w': wb = 1; break;
        case 'p': prefetch = 1; break;
        case 'P': pools = atoi(optarg); if( pools < 1 ) pools = 1; if( pools > 8 ) pools = 8; break;
        default: break;
        }
    }
    if( NULL != getenv("PB2_TEST_BACKTRACE") ) { signal(SIGSEGV, on_segv); signal(SIGABRT, on_segv); signal(SIGALRM, on_segv); alarm(15); }
    int pargc = argc - optind + 1;
    char **pargv = (char**)calloc((size_t)pargc + 1, sizeof(char*));
    pargv[0] = argv[0];
    for( int i = optind; i < argc; i++ ) pargv[i - optind + 1] = argv[i];

    parsec_context_t *parsec = parsec_init(cores, &pargc, &pargv);
    if( NULL == parsec ) { fprintf(stderr, "parsec_init failed\n"); return 2; }
    const int F = NB / 2 + 1;
    const int nthreads = parsec->virtual_processes[0]->nb_cores;

    parsec_matrix_block_cyclic_t dcs[8];
#define dcA dcs[0]
    /* tiles of tile_mb x tile_nb elements stacked in one column of tiles: the matrix sizes of the reference API are ints,
     * K * elems does not fit one for the 8-GPU workload (32768 tiles of 65536 elements) */
    const int tile_nb = (0 == elems % 256) ? 256 : 1, tile_mb = elems / tile_nb;
    parsec_matrix_block_cyclic_init(&dcA, PARSEC_MATRIX_INTEGER, PARSEC_MATRIX_TILE, 0,
                                    tile_mb, tile_nb, K * tile_mb, tile_nb, 0, 0, K * tile_mb, tile_nb, 1, 1, 1, 1, 0, 0);
    dcA.mat = parsec_data_allocate((size_t)dcA.super.nb_local_tiles * (size_t)dcA.super.bsiz *
                                   (size_t)parsec_datadist_getsizeoftype(dcA.super.mtype));
    parsec_data_collection_set_key((parsec_data_collection_t*)&dcA, "dcA");
    static char dc_names[8][8];
    for( int q = 1; q < pools; q++ ) {
        parsec_matrix_block_cyclic_init(&dcs[q], PARSEC_MATRIX_INTEGER, PARSEC_MATRIX_TILE, 0,
                                        tile_mb, tile_nb, K * tile_mb, tile_nb, 0, 0, K * tile_mb, tile_nb, 1, 1, 1, 1, 0, 0);
        dcs[q].mat = parsec_data_allocate((size_t)dcs[q].super.nb_local_tiles * (size_t)dcs[q].super.bsiz *
                                          (size_t)parsec_datadist_getsizeoftype(dcs[q].super.mtype));
        snprintf(dc_names[q], sizeof dc_names[q], "dc%c", 'A' + q);
        parsec_data_collection_set_key((parsec_data_collection_t*)&dcs[q], dc_names[q]);
    }

    int ngpu = 0, b200 = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL == d || !PARSEC_DEV_IS_GPU(d->type) ) continue;
        ngpu++; b200 += parsec_b200_is_b200_device(d);
        for( int q = 0; gpu && q < pools; q++ ) dcs[q].super.super.register_memory(&dcs[q].super.super, d);
    }
    if( gpu && 0 == ngpu ) { fprintf(stderr, "-m gpu but no GPU device module is active\n"); return 3; }

    /* -p: PARSEC_DEV_DATA_ADVICE_PREFETCH (device.h:79-81) for every tile before the first task pool is handed over */
    uint64_t h2d_prefetch = 0;
    if( gpu && prefetch ) {
        int gdev = -1;
        for( int i = 0; i < (int)parsec_nb_devices && gdev < 0; i++ ) {
            parsec_device_module_t *d = parsec_mca_device_get(i);
            if( NULL != d && PARSEC_DEV_IS_GPU(d->type) ) gdev = i;
        }
        for( int k = 0; k < K; k++ ) {
            parsec_data_t *dta = dcA.super.super.data_of(&dcA.super.super, k, 0);
            if( PARSEC_SUCCESS != parsec_advise_data_on_device(dta, gdev, PARSEC_DEV_DATA_ADVICE_PREFETCH) ) return 5;
        }
        parsec_device_module_t *d = parsec_mca_device_get(gdev);
        if( NULL != d->data_in_from_device ) h2d_prefetch = d->data_in_from_device[0];
    }
    int64_t *errors = (int64_t*)calloc((size_t)nthreads + 1, sizeof(int64_t));
    double best = 1e30, total = 0;
    char times[4096]; int tl = 0; times[0] = 0;
    int64_t bad_total = 0;
    for( int r = 0; r < repeats; r++ ) {
        parsec_ex05_b200_taskpool_t *tps[8];
        for( int q = 0; q < pools; q++ ) {
            int32_t *m = (int32_t*)dcs[q].mat;
            for( size_t i = 0; i < (size_t)K * elems; i++ ) m[i] = -7;
            tps[q] = parsec_ex05_b200_new(&dcs[q].super, NB, errors, wb);
            parsec_arena_datatype_set_type(&tps[q]->arenas_datatypes[PARSEC_ex05_b200_DEFAULT_ADT_IDX],
                                           (size_t)elems * sizeof(int32_t), PARSEC_ARENA_ALIGNMENT_SSE, parsec_datatype_int_t);
            if( !gpu ) {
                for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
                    parsec_device_module_t *d = parsec_mca_device_get(i);
                    if( NULL != d && PARSEC_DEV_IS_GPU(d->type) ) tps[q]->super.devices_index_mask &= ~(1u << i);
                }
            }
        }
        struct timespec ts0, ts1; clock_gettime(CLOCK_MONOTONIC, &ts0);
        const double t0 = now_s();
        for( int q = 0; q < pools; q++ )
            if( 0 > parsec_context_add_taskpool(parsec, (parsec_taskpool_t*)tps[q]) ) return 4;
        const double ta = now_s();
        if( 0 > parsec_context_start(parsec) ) return 4;
        const double tst = now_s();
        if( 0 > parsec_context_wait(parsec) ) return 4;
        const double t1 = now_s();
        clock_gettime(CLOCK_MONOTONIC, &ts1);
        if( verbose ) {
            const double m0 = ts0.tv_sec * 1e3 + ts0.tv_nsec * 1e-6, m1 = ts1.tv_sec * 1e3 + ts1.tv_nsec * 1e-6;
            for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
                parsec_device_module_t *d = parsec_mca_device_get(i);
                parsec_b200_stats_t s1;
                if( NULL != d && parsec_b200_is_b200_device(d) && 0 == parsec_b200_get_stats(d, &s1) )
                    fprintf(stderr, "  dev %d: add_taskpool %.3f ms, start %.3f ms, first kernel_scheduler entry at %.3f ms, manager at %.3f ms, last completion %.3f ms before wait returned\n", i,
                            1e3 * (ta - t0), 1e3 * (tst - ta), s1.first_entry_ns * 1e-6 - m0, s1.first_task_ns * 1e-6 - m0, m1 - s1.last_done_ns * 1e-6);
            }
        }
        /* bring every tile home (what a real application does before it reads its matrix on the CPU) */
        for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
            parsec_device_module_t *d = parsec_mca_device_get(i);
            if( NULL != d && PARSEC_DEV_IS_GPU(d->type) && NULL != d->memory_release ) d->memory_release(d);
        }
        const double t2 = now_s();
        /* the host tiles hold k when a CPU body wrote them, or when the GPU result was written back (-w, or a module whose
         * memory_release brings dirty replicas home) */
        if( !gpu || wb || b200 )
            for( int q = 0; q < pools; q++ )
                for( int k = 0; k < K; k++ )
                    for( int i = 0; i < elems; i += (elems > 64 ? elems / 64 : 1) ) bad_total += (((int32_t*)dcs[q].mat)[(size_t)k * elems + i] != k);
        if( verbose ) fprintf(stderr, "repeat %d: dag %.3f ms, flush %.3f ms\n", r, 1e3 * (t1 - t0), 1e3 * (t2 - t1));
        if( tl < 4000 ) tl += snprintf(times + tl, sizeof(times) - (size_t)tl, "%s%.6f", r ? ", " : "", t1 - t0);
        if( t1 - t0 < best ) best = t1 - t0;
        total += t1 - t0;
        for( int q = 0; q < pools; q++ ) {
            PARSEC_OBJ_DESTRUCT(&tps[q]->arenas_datatypes[PARSEC_ex05_b200_DEFAULT_ADT_IDX]);
            parsec_taskpool_free((parsec_taskpool_t*)tps[q]);
        }
    }
    for( int t = 0; t < nthreads; t++ ) bad_total += errors[t];

    parsec_b200_stats_t st; memset(&st, 0, sizeof st);
    uint64_t executed_gpu = 0, h2d = 0, required_in = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL == d || !PARSEC_DEV_IS_GPU(d->type) ) continue;
        executed_gpu += d->executed_tasks; required_in += d->required_data_in;
        if( NULL != d->data_in_from_device ) h2d += d->data_in_from_device[0];
        if( parsec_b200_is_b200_device(d) ) {
            parsec_b200_stats_t s1; parsec_b200_get_stats(d, &s1);
            st.tasks_engine += s1.tasks_engine; st.tasks_lane += s1.tasks_lane; st.kernel_launches += s1.kernel_launches;
            st.released_on_device += s1.released_on_device; st.forwarded += s1.forwarded;
            st.bytes_h2d_kernel += s1.bytes_h2d_kernel; st.bytes_h2d_dma += s1.bytes_h2d_dma; st.bytes_d2h_dma += s1.bytes_d2h_dma;
            st.manager_entries += s1.manager_entries; st.evictions += s1.evictions; st.w2r_copies += s1.w2r_copies;
            st.check_mismatches += s1.check_mismatches; st.peer_pulls += s1.peer_pulls; st.peer_detours += s1.peer_detours;
            if( s1.max_concurrent_callers > st.max_concurrent_callers ) st.max_concurrent_callers = s1.max_concurrent_callers;
        }
    }
    if( gpu && ngpu > b200 ) {      /* bodies ran as stand-alone kernels under a foreign module: their CHECK counter */
        uint64_t e = 0;
        if( 0 == pb2_body_launch_errors(&e, 1) ) st.check_mismatches += e;
    }
    bad_total += (int64_t)st.check_mismatches;
    const long ntasks = (long)K * (1 + F) * pools;
    uint64_t checksum = 0;                       /* the host tiles are final only with -w */
    for( int q = 0; wb && q < pools; q++ ) checksum = fnv1a64(dcs[q].mat, (size_t)K * elems * sizeof(int32_t), checksum);
    printf("{\"app\": \"ex05_b200\", \"checksum\": \"%016lx\", \"mode\": \"%s\", \"wb\": %d, \"K\": %d, \"NB\": %d, \"F\": %d, \"tile_bytes\": %ld, \"tasks\": %ld, \"repeats\": %d, \"pools\": %d, "
           "\"cores\": %d, \"gpu_modules\": %d, \"b200_modules\": %d, \"best_s\": %.6f, \"mean_s\": %.6f, \"times_s\": [%s], \"tasks_per_s\": %.1f, "
           "\"errors\": %ld, \"executed_on_gpu\": %lu, \"required_in\": %lu, \"h2d_bytes\": %lu, \"h2d_prefetch_bytes\": %lu, "
           "\"b200\": {\"tasks_engine\": %lu, \"tasks_lane\": %lu, \"kernel_launches\": %lu, \"released_on_device\": %lu, "
           "\"forwarded\": %lu, \"bytes_h2d_kernel\": %lu, \"bytes_h2d_dma\": %lu, \"bytes_d2h_dma\": %lu, "
           "\"check_mismatches\": %lu, \"manager_entries\": %lu, \"max_concurrent_callers\": %lu, \"evictions\": %lu, \"w2r_copies\": %lu, \"peer_pulls\": %lu, \"peer_detours\": %lu}}\n",
           (unsigned long)checksum, gpu ? "gpu" : "cpu", wb, K, NB, F, (long)elems * 4, ntasks, repeats, pools, nthreads, ngpu, b200, best, total / repeats, times,
           ntasks / best, (long)bad_total, (unsigned long)executed_gpu, (unsigned long)required_in, (unsigned long)h2d, (unsigned long)h2d_prefetch,
           (unsigned long)st.tasks_engine, (unsigned long)st.tasks_lane, (unsigned long)st.kernel_launches,
           (unsigned long)st.released_on_device, (unsigned long)st.forwarded, (unsigned long)st.bytes_h2d_kernel,
           (unsigned long)st.bytes_h2d_dma, (unsigned long)st.bytes_d2h_dma, (unsigned long)st.check_mismatches, (unsigned long)st.manager_entries,
           (unsigned long)st.max_concurrent_callers, (unsigned long)st.evictions, (unsigned long)st.w2r_copies, (unsigned long)st.peer_pulls, (unsigned long)st.peer_detours);

    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        for( int q = 0; gpu && NULL != d && PARSEC_DEV_IS_GPU(d->type) && q < pools; q++ ) dcs[q].super.super.unregister_memory(&dcs[q].super.super, d);
    }
    for( int q = 0; q < pools; q++ ) {
        parsec_data_free(dcs[q].mat);
        parsec_tiled_matrix_destroy((parsec_tiled_matrix_t*)&dcs[q]);
    }
    free(errors);
    parsec_fini(&parsec);
    return bad_total ? 1 : 0;
}
