/* ex02_main.c -- driver of tests/parsec/ex02_b200.jdf (Ex02_Chain's main plays this role in the reference).
 *   ex02_b200 [-N NB] [-t tile_elems] [-r repeats] [-c cores] [-m cpu|gpu]
 * Prints one JSON line: tasks/s and ns per dependency edge; exit code 0 iff the tile holds the known answer. */
#include "parsec.h"
#include "parsec/data_dist/matrix/two_dim_rectangle_cyclic.h"
#include "parsec/mca/device/device.h"
#include "parsec/mca/device/b200/device_b200.h"
#include "parsec/parsec_internal.h"
#include "parsec/execution_stream.h"
#include "ex02_b200.h"
#include "checksum.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(int argc, char *argv[])
{
    int NB = 999, elems = 1, repeats = 5, cores = 2, gpu = 1, c;
    while( -1 != (c = getopt(argc, argv, "N:t:r:c:m:")) ) {
        switch(c) {
        case 'N': NB = atoi(optarg); break;
        case 't': elems = atoi(optarg); break;
        case 'r': repeats = atoi(optarg); break;
        case 'c': cores = atoi(optarg); break;
        case 'm': gpu = (0 == strcmp(optarg, "gpu")); break;
        default: break;
        }
    }
    int pargc = 1; char *pargv0[2] = { argv[0], NULL }; char **pargv = pargv0;
    parsec_context_t *parsec = parsec_init(cores, &pargc, &pargv);
    if( NULL == parsec ) return 2;
    parsec_matrix_block_cyclic_t dcA;
    parsec_matrix_block_cyclic_init(&dcA, PARSEC_MATRIX_INTEGER, PARSEC_MATRIX_TILE, 0, elems, 1, elems, 1, 0, 0, elems, 1, 1, 1, 1, 1, 0, 0);
    dcA.mat = parsec_data_allocate((size_t)elems * sizeof(int32_t));
    parsec_data_collection_set_key((parsec_data_collection_t*)&dcA, "dcA");
    int32_t *mat = (int32_t*)dcA.mat;
    for( int i = 0; i < elems; i++ ) mat[i] = 7;      /* a defined start value: the golden vectors checksum the final tile */
    int ngpu = 0, b200 = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL == d || !PARSEC_DEV_IS_GPU(d->type) ) continue;
        ngpu++; b200 += parsec_b200_is_b200_device(d);
        if( gpu ) dcA.super.super.register_memory(&dcA.super.super, d);
    }
    if( gpu && 0 == ngpu ) { fprintf(stderr, "-m gpu but no GPU device module is active\n"); return 3; }
    double best = 1e30, total = 0; long bad = 0;
    for( int r = 0; r < repeats; r++ ) {
        parsec_ex02_b200_taskpool_t *tp = parsec_ex02_b200_new(&dcA.super, NB);
        parsec_arena_datatype_set_type(&tp->arenas_datatypes[PARSEC_ex02_b200_DEFAULT_ADT_IDX], (size_t)elems * sizeof(int32_t),
                                       PARSEC_ARENA_ALIGNMENT_SSE, parsec_datatype_int_t);
        if( !gpu )
            for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
                parsec_device_module_t *d = parsec_mca_device_get(i);
                if( NULL != d && PARSEC_DEV_IS_GPU(d->type) ) tp->super.devices_index_mask &= ~(1u << i);
            }
        const int32_t before = mat[0];
        const double t0 = now_s();
        if( 0 > parsec_context_add_taskpool(parsec, (parsec_taskpool_t*)tp) ) return 4;
        if( 0 > parsec_context_start(parsec) ) return 4;
        if( 0 > parsec_context_wait(parsec) ) return 4;
        const double dt = now_s() - t0;
        for( int i = 0; i < elems; i += (elems > 16 ? elems / 16 : 1) ) bad += (mat[i] != before + NB + 1);
        if( dt < best ) best = dt;
        total += dt;
        PARSEC_OBJ_DESTRUCT(&tp->arenas_datatypes[PARSEC_ex02_b200_DEFAULT_ADT_IDX]);
        parsec_taskpool_free((parsec_taskpool_t*)tp);
    }
    uint64_t on_gpu = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL != d && PARSEC_DEV_IS_GPU(d->type) ) on_gpu += d->executed_tasks;
    }
    const uint64_t checksum = fnv1a64(mat, (size_t)elems * sizeof(int32_t), 0);
    printf("{\"app\": \"ex02_b200\", \"checksum\": \"%016lx\", \"mode\": \"%s\", \"NB\": %d, \"tasks\": %d, \"tile_bytes\": %ld, \"repeats\": %d, \"cores\": %d, "
           "\"gpu_modules\": %d, \"b200_modules\": %d, \"best_s\": %.6f, \"mean_s\": %.6f, \"tasks_per_s\": %.1f, \"ns_per_edge\": %.1f, "
           "\"errors\": %ld, \"executed_on_gpu\": %lu}\n", (unsigned long)checksum, gpu ? "gpu" : "cpu", NB, NB + 1, (long)elems * 4, repeats,
           parsec->virtual_processes[0]->nb_cores, ngpu, b200, best, total / repeats, (NB + 1) / best, best / NB * 1e9, bad, (unsigned long)on_gpu);
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( gpu && NULL != d && PARSEC_DEV_IS_GPU(d->type) ) {
            if( NULL != d->memory_release ) d->memory_release(d);
            dcA.super.super.unregister_memory(&dcA.super.super, d);
        }
    }
    parsec_data_free(dcA.mat);
    parsec_tiled_matrix_destroy((parsec_tiled_matrix_t*)&dcA);
    parsec_fini(&parsec);
    return bad ? 1 : 0;
}
