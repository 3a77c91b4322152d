/* stage_main.c -- driver of tests/parsec/stage_b200.jdf (the reference's tests/runtime/cuda/stage_main.c plays this role
 * for stage_custom.jdf).  Prints one JSON line; exit code 0 iff every element matched. */
#include "parsec.h"
#include "parsec/data_dist/matrix/two_dim_rectangle_cyclic.h"
#include "parsec/mca/device/device.h"
#include "parsec/mca/device/b200/device_b200.h"
#include "parsec/parsec_internal.h"
#include "parsec/execution_stream.h"
#include "stage_b200.h"
#include "checksum.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

extern int32_t stage_b200_complete_stage_calls;

static void fill(parsec_matrix_block_cyclic_t *dc, int lapack)
{
    int32_t *mat = (int32_t*)dc->mat;
    const int mb = dc->super.mb, nb = dc->super.nb;
    for( int m = 0; m < dc->super.mt; m++ )
        for( int k = 0; k < dc->super.nt; k++ )
            for( int j = 0; j < nb; j++ )
                for( int i = 0; i < mb; i++ ) {
                    const int32_t v = (int32_t)((m * mb + i) * 1000 + (k * nb + j));
                    if( lapack ) mat[(size_t)(k * nb + j) * dc->super.llm + (m * mb + i)] = v;
                    else mat[((size_t)k * dc->super.mt + m) * dc->super.bsiz + (size_t)j * mb + i] = v;   /* tile (m,k) at position k*mt+m */
                }
}

int main(int argc, char *argv[])
{
    int MT = 4, NT = 3, MB = 96, NB = 64, cores = 4, gpu = 1, c;
    while( -1 != (c = getopt(argc, argv, "M:N:b:n:c:m:")) ) {
        switch(c) {
        case 'M': MT = atoi(optarg); break;
        case 'N': NT = atoi(optarg); break;
        case 'b': MB = atoi(optarg); break;
        case 'n': NB = atoi(optarg); break;
        case 'c': cores = atoi(optarg); break;
        case 'm': gpu = (0 == strcmp(optarg, "gpu")); break;
        default: break;
        }
    }
    int pargc = 1; char *pargv0[2] = { argv[0], NULL }; char **pargv = pargv0;
    parsec_context_t *parsec = parsec_init(cores, &pargc, &pargv);
    if( NULL == parsec ) return 2;

    parsec_matrix_block_cyclic_t dcA, dcB, dcC;
    parsec_matrix_block_cyclic_init(&dcA, PARSEC_MATRIX_INTEGER, PARSEC_MATRIX_TILE,   0, MB, NB, MT * MB, NT * NB, 0, 0, MT * MB, NT * NB, 1, 1, 1, 1, 0, 0);
    parsec_matrix_block_cyclic_init(&dcB, PARSEC_MATRIX_INTEGER, PARSEC_MATRIX_LAPACK, 0, MB, NB, MT * MB, NT * NB, 0, 0, MT * MB, NT * NB, 1, 1, 1, 1, 0, 0);
    parsec_matrix_block_cyclic_init(&dcC, PARSEC_MATRIX_INTEGER, PARSEC_MATRIX_TILE,   0, MB, NB, MT * MB, NT * NB, 0, 0, MT * MB, NT * NB, 1, 1, 1, 1, 0, 0);
    const size_t bytes = (size_t)MT * MB * NT * NB * sizeof(int32_t);
    dcA.mat = parsec_data_allocate(bytes); dcB.mat = parsec_data_allocate(bytes); dcC.mat = parsec_data_allocate(bytes);
    parsec_data_collection_set_key((parsec_data_collection_t*)&dcA, "dcA");
    parsec_data_collection_set_key((parsec_data_collection_t*)&dcB, "dcB");
    parsec_data_collection_set_key((parsec_data_collection_t*)&dcC, "dcC");
    fill(&dcA, 0); fill(&dcB, 1); fill(&dcC, 0);

    int ngpu = 0, b200 = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL == d || !PARSEC_DEV_IS_GPU(d->type) ) continue;
        ngpu++; b200 += parsec_b200_is_b200_device(d);
        if( gpu ) {   /* A and B pinned (the kernel reads them in place), C left pageable on purpose: copy-engine path */
            dcA.super.super.register_memory(&dcA.super.super, d);
            dcB.super.super.register_memory(&dcB.super.super, d);
        }
    }
    int out = 0;
    parsec_stage_b200_taskpool_t *tp = parsec_stage_b200_new(&dcA.super, &dcB.super, &dcC.super, &out);
    parsec_arena_datatype_set_type(&tp->arenas_datatypes[PARSEC_stage_b200_DEFAULT_ADT_IDX], (size_t)MB * NB * sizeof(int32_t),
                                   PARSEC_ARENA_ALIGNMENT_SSE, parsec_datatype_int_t);
    if( !gpu )
        for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
            parsec_device_module_t *d = parsec_mca_device_get(i);
            if( NULL != d && PARSEC_DEV_IS_GPU(d->type) ) tp->super.devices_index_mask &= ~(1u << i);
        }
    if( 0 > parsec_context_add_taskpool(parsec, (parsec_taskpool_t*)tp) ) return 4;
    if( 0 > parsec_context_start(parsec) ) return 4;
    if( 0 > parsec_context_wait(parsec) ) return 4;

    /* every class writes back to its collection: the host matrices hold the known answer too */
    long host_bad = 0;
    const int32_t *A = (const int32_t*)dcA.mat, *B = (const int32_t*)dcB.mat, *C = (const int32_t*)dcC.mat;
    for( int m = 0; m < MT; m++ ) for( int k = 0; k < NT; k++ ) for( int j = 0; j < NB; j++ ) for( int i = 0; i < MB; i++ ) {
        const int32_t want = (int32_t)((m * MB + i) * 1000 + (k * NB + j) + 5);
        const size_t t = ((size_t)k * MT + m) * (size_t)MB * NB + (size_t)j * MB + i;
        host_bad += (A[t] != want) + (C[t] != want) + (B[(size_t)(k * NB + j) * (MT * MB) + (m * MB + i)] != want);
    }
    parsec_b200_stats_t st; memset(&st, 0, sizeof st);
    uint64_t on_gpu = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL == d || !PARSEC_DEV_IS_GPU(d->type) ) continue;
        on_gpu += d->executed_tasks;
        if( parsec_b200_is_b200_device(d) ) { parsec_b200_stats_t s1; parsec_b200_get_stats(d, &s1);
            st.tasks_engine += s1.tasks_engine; st.tasks_lane += s1.tasks_lane; st.bytes_h2d_dma += s1.bytes_h2d_dma; st.bytes_d2h_dma += s1.bytes_d2h_dma; }
    }
    uint64_t checksum = fnv1a64(dcA.mat, bytes, 0);
    checksum = fnv1a64(dcB.mat, bytes, checksum); checksum = fnv1a64(dcC.mat, bytes, checksum);
    printf("{\"app\": \"stage_b200\", \"checksum\": \"%016lx\", \"mode\": \"%s\", \"tiles\": %d, \"check_errors\": %d, \"host_errors\": %ld, \"executed_on_gpu\": %lu, "
           "\"gpu_modules\": %d, \"b200_modules\": %d, \"tasks_engine\": %lu, \"tasks_lane\": %lu, \"bytes_h2d_dma\": %lu, \"bytes_d2h_dma\": %lu, "
           "\"complete_stage_calls\": %d}\n", (unsigned long)checksum, gpu ? "gpu" : "cpu", MT * NT, out, host_bad, (unsigned long)on_gpu, ngpu, b200,
           (unsigned long)st.tasks_engine, (unsigned long)st.tasks_lane, (unsigned long)st.bytes_h2d_dma, (unsigned long)st.bytes_d2h_dma,
           stage_b200_complete_stage_calls);
    PARSEC_OBJ_DESTRUCT(&tp->arenas_datatypes[PARSEC_stage_b200_DEFAULT_ADT_IDX]);
    parsec_taskpool_free((parsec_taskpool_t*)tp);
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( gpu && NULL != d && PARSEC_DEV_IS_GPU(d->type) ) {
            if( NULL != d->memory_release ) d->memory_release(d);
            dcA.super.super.unregister_memory(&dcA.super.super, d);
            dcB.super.super.unregister_memory(&dcB.super.super, d);
        }
    }
    parsec_data_free(dcA.mat); parsec_data_free(dcB.mat); parsec_data_free(dcC.mat);
    parsec_tiled_matrix_destroy((parsec_tiled_matrix_t*)&dcA); parsec_tiled_matrix_destroy((parsec_tiled_matrix_t*)&dcB); parsec_tiled_matrix_destroy((parsec_tiled_matrix_t*)&dcC);
    parsec_fini(&parsec);
    return (out || host_bad) ? 1 : 0;
}
