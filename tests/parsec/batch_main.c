/* batch_main.c -- driver of tests/parsec/batch_b200.jdf.  Prints one JSON line; exit code 0 iff every element matched. */
#include "parsec.h"
#include "parsec/data_dist/matrix/two_dim_rectangle_cyclic.h"
#include "parsec/mca/device/device.h"
#include "parsec/mca/device/b200/device_b200.h"
#include "parsec/parsec_internal.h"
#include "batch_b200.h"
#include "checksum.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

extern int32_t batch_b200_submit_calls, batch_b200_tasks_in_batches, batch_b200_max_batch;

int main(int argc, char *argv[])
{
    int MT = 64, MB = 64, NB = 64, cores = 4, gpu = 1, c;
    while( -1 != (c = getopt(argc, argv, "M:b:c:m:")) ) {
        switch(c) {
        case 'M': MT = atoi(optarg); break;
        case 'b': MB = NB = atoi(optarg); break;
        case 'c': cores = atoi(optarg); break;
        case 'm': gpu = (0 == strcmp(optarg, "gpu")); break;
        default: break;
        }
    }
    int pargc = 1; char *pargv0[2] = { argv[0], NULL }; char **pargv = pargv0;
    parsec_context_t *parsec = parsec_init(cores, &pargc, &pargv);
    if( NULL == parsec ) return 2;
    parsec_matrix_block_cyclic_t dcA;
    parsec_matrix_block_cyclic_init(&dcA, PARSEC_MATRIX_INTEGER, PARSEC_MATRIX_TILE, 0, MB, NB, MT * MB, NB, 0, 0, MT * MB, NB, 1, 1, 1, 1, 0, 0);
    dcA.mat = parsec_data_allocate((size_t)MT * MB * NB * sizeof(int32_t));
    parsec_data_collection_set_key((parsec_data_collection_t*)&dcA, "dcA");
    int32_t *mat = (int32_t*)dcA.mat;
    for( size_t i = 0; i < (size_t)MT * MB * NB; i++ ) mat[i] = (int32_t)(i % 100003);

    int ngpu = 0, b200 = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL == d || !PARSEC_DEV_IS_GPU(d->type) ) continue;
        ngpu++; b200 += parsec_b200_is_b200_device(d);
    }
    parsec_batch_b200_taskpool_t *tp = parsec_batch_b200_new(&dcA.super);
    parsec_arena_datatype_set_type(&tp->arenas_datatypes[PARSEC_batch_b200_DEFAULT_ADT_IDX], (size_t)MB * NB * sizeof(int32_t),
                                   PARSEC_ARENA_ALIGNMENT_SSE, parsec_datatype_int_t);
    if( !gpu )
        for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
            parsec_device_module_t *d = parsec_mca_device_get(i);
            if( NULL != d && PARSEC_DEV_IS_GPU(d->type) ) tp->super.devices_index_mask &= ~(1u << i);
        }
    if( 0 > parsec_context_add_taskpool(parsec, (parsec_taskpool_t*)tp) ) return 4;
    if( 0 > parsec_context_start(parsec) ) return 4;
    if( 0 > parsec_context_wait(parsec) ) return 5;
    uint64_t on_gpu = 0, lane = 0, batched = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL == d || !PARSEC_DEV_IS_GPU(d->type) ) continue;
        on_gpu += d->executed_tasks;
        parsec_b200_stats_t st;
        if( PARSEC_SUCCESS == parsec_b200_get_stats(d, &st) ) { lane += st.tasks_lane; batched += st.lane_batched; }
    }
    parsec_devices_release_memory();
    long errors = 0;
    for( size_t i = 0; i < (size_t)MT * MB * NB; i++ ) errors += (mat[i] != (int32_t)(i % 100003) + 5);
    const uint64_t checksum = fnv1a64(mat, (size_t)MT * MB * NB * sizeof(int32_t), 0);
    printf("{\"app\": \"batch_b200\", \"checksum\": \"%016lx\", \"tiles\": %d, \"gpu_modules\": %d, \"b200_modules\": %d, \"errors\": %ld, \"executed_on_gpu\": %lu, "
           "\"submit_calls\": %d, \"tasks_in_batches\": %d, \"max_batch\": %d, \"tasks_lane\": %lu, \"lane_batched\": %lu}\n",
           (unsigned long)checksum, MT, ngpu, b200, errors, (unsigned long)on_gpu, batch_b200_submit_calls, batch_b200_tasks_in_batches, batch_b200_max_batch,
           (unsigned long)lane, (unsigned long)batched);
    parsec_taskpool_free((parsec_taskpool_t*)tp);
    parsec_data_free(dcA.mat);
    parsec_tiled_matrix_destroy((parsec_tiled_matrix_t*)&dcA);
    parsec_fini(&parsec);
    return errors ? 1 : 0;
}
