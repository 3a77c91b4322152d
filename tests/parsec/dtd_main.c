/* dtd_main.c -- DTD task pools of the reference runtime scheduled through the b200 device component.
 *
 * The cases restate, for one process, the reference's own DTD GPU tests (golden vectors of SURVEY.md 8c):
 *   memset / memset_and_read   tests/dsl/dtd/dtd_test_cuda_task_insert.c:159-310 (every element 0xFFFFFFFF after the
 *                              GPU task, a CPU task that follows reads the same value)
 *   new_tile                   tests/dsl/dtd/dtd_test_new_tile.c + dtd_test_new_tile_cuda_kernels.cu:17-33 (data[i] == 2*i)
 *   pingpong                   tests/dsl/dtd/dtd_test_pingpong.c:45-96 restated as a CPU <-> GPU ping-pong on each tile: a
 *                              chain of NT tasks that add one, alternately on the host and on the device (data == start + NT)
 * The chores are added with parsec_dtd_task_class_add_chore(..., PARSEC_DEV_CUDA, fn) and inserted with
 * parsec_dtd_insert_task_with_task_class exactly as the reference tests do (insert_function.c:2393-2425 builds the
 * parsec_gpu_task_t and hands it to the module's kernel_scheduler).  GPU bodies either NAME an engine body
 * (parsec_b200_task_body: the task runs in the persistent kernel) or -- `-o` -- enqueue their own CUDA work on the
 * stream they are given, like the reference's cuda_memset_task_fn (the module's stream lane).
 * The collection is plain pageable memory (parsec_data_allocate): stage-in and pushout go through the copy engine.
 * Prints one JSON line; exit code 0 iff every element matched. */
#include "parsec.h"
#include "parsec/arena.h"
#include "parsec/data_dist/matrix/matrix.h"
#include "parsec/data_dist/matrix/two_dim_rectangle_cyclic.h"
#include "parsec/interfaces/dtd/insert_function.h"
#include "parsec/mca/device/device.h"
#include "parsec/mca/device/device_gpu.h"
#include "parsec/mca/device/cuda/device_cuda.h"
#include "parsec/mca/device/b200/device_b200.h"
#include "parsec/parsec_internal.h"
#include "pb2_engine.h"
#include <cuda_runtime_api.h>
#include "checksum.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static int TILE_FULL;
static int opaque = 0;
static volatile int32_t read_errors = 0;

/* ---- GPU chores ---------------------------------------------------------------------------------------------- */
static const int flow0 = 0;

static int gpu_memset_fn(parsec_device_gpu_module_t *gpu_device, parsec_gpu_task_t *gpu_task, parsec_gpu_exec_stream_t *gpu_stream)
{
    int32_t *data; int nb;
    parsec_dtd_unpack_args(gpu_task->ec, &data, &nb);
    if( opaque ) {      /* the reference's body: its own CUDA call on the stream it is given */
        parsec_cuda_exec_stream_t *cs = (parsec_cuda_exec_stream_t*)gpu_stream;
        void *dev = parsec_dtd_get_dev_ptr(gpu_task->ec, 0);
        return (cudaSuccess == cudaMemsetAsync(dev, 0xFF, (size_t)nb * sizeof(int32_t), cs->cuda_stream)) ? PARSEC_HOOK_RETURN_DONE : PARSEC_HOOK_RETURN_ERROR;
    }
    int32_t ip[3] = { 0xFF, 0, 0 };
    return parsec_b200_task_body(gpu_device, gpu_task, gpu_stream, PB2_BODY_MEMSET_U8, 1, &flow0, ip, 0.f);
}
static int gpu_iota_fn(parsec_device_gpu_module_t *gpu_device, parsec_gpu_task_t *gpu_task, parsec_gpu_exec_stream_t *gpu_stream)
{
    return parsec_b200_task_body(gpu_device, gpu_task, gpu_stream, PB2_BODY_IOTA_I32, 1, &flow0, NULL, 0.f);
}
static int gpu_scale_fn(parsec_device_gpu_module_t *gpu_device, parsec_gpu_task_t *gpu_task, parsec_gpu_exec_stream_t *gpu_stream)
{
    int32_t *data; int factor;
    parsec_dtd_unpack_args(gpu_task->ec, &data, &factor);
    int32_t ip[3] = { factor, 0, 0 };
    return parsec_b200_task_body(gpu_device, gpu_task, gpu_stream, PB2_BODY_SCALE_I32, 1, &flow0, ip, 0.f);
}
static int gpu_incr_fn(parsec_device_gpu_module_t *gpu_device, parsec_gpu_task_t *gpu_task, parsec_gpu_exec_stream_t *gpu_stream)
{
    int32_t *data; int by;
    parsec_dtd_unpack_args(gpu_task->ec, &data, &by);
    int32_t ip[3] = { by, 0, 0 };
    return parsec_b200_task_body(gpu_device, gpu_task, gpu_stream, PB2_BODY_INCR_I32, 1, &flow0, ip, 0.f);
}

/* ---- CPU chores ---------------------------------------------------------------------------------------------- */
static int cpu_memset_fn(parsec_execution_stream_t *es, parsec_task_t *this_task)
{
    int32_t *data; int nb; (void)es;
    parsec_dtd_unpack_args(this_task, &data, &nb);
    memset(data, 0xFF, (size_t)nb * sizeof(int32_t));
    return PARSEC_HOOK_RETURN_DONE;
}
static int cpu_read_fn(parsec_execution_stream_t *es, parsec_task_t *this_task)
{
    int32_t *data; int nb, want; (void)es;
    parsec_dtd_unpack_args(this_task, &data, &nb, &want);
    for( int j = 0; j < nb; j++ ) if( data[j] != want ) { (void)parsec_atomic_fetch_inc_int32(&read_errors); break; }
    return PARSEC_HOOK_RETURN_DONE;
}
static int cpu_incr_fn(parsec_execution_stream_t *es, parsec_task_t *this_task)
{
    int32_t *data; int by; (void)es;
    parsec_dtd_unpack_args(this_task, &data, &by);
    /* the tile size travels in the datatype: the collection's tiles are `nb` elements, kept in a global */
    extern int dtd_nb;
    for( int j = 0; j < dtd_nb; j++ ) data[j] += by;
    return PARSEC_HOOK_RETURN_DONE;
}
int dtd_nb = 0;
static int cpu_iota_fn(parsec_execution_stream_t *es, parsec_task_t *this_task)
{
    int32_t *data; (void)es;
    parsec_dtd_unpack_args(this_task, &data);
    for( int j = 0; j < dtd_nb; j++ ) data[j] = j;
    return PARSEC_HOOK_RETURN_DONE;
}
static int cpu_scale_fn(parsec_execution_stream_t *es, parsec_task_t *this_task)
{
    int32_t *data; int factor; (void)es;
    parsec_dtd_unpack_args(this_task, &data, &factor);
    for( int j = 0; j < dtd_nb; j++ ) data[j] *= factor;
    return PARSEC_HOOK_RETURN_DONE;
}

/* ---- helpers ------------------------------------------------------------------------------------------------- */
static int32_t *tile_ptr(parsec_data_collection_t *A, int i)
{
    parsec_data_t *d = A->data_of_key(A, A->data_key(A, i, 0));
    return (int32_t*)parsec_data_copy_get_ptr(parsec_data_get_copy(d, 0));
}
static int32_t start_value(int i, int j) { return i * 1000 + j + 1; }

int main(int argc, char *argv[])
{
    int MT = 16, nb = 4096, NT = 8, cores = 4, c, gpu_only = 0, cpu_only = 0;
    while( -1 != (c = getopt(argc, argv, "M:n:N:c:ogC")) ) {
        switch(c) {
        case 'M': MT = atoi(optarg); break;
        case 'n': nb = atoi(optarg); break;
        case 'N': NT = atoi(optarg); break;
        case 'c': cores = atoi(optarg); break;
        case 'o': opaque = 1; break;
        case 'g': gpu_only = 1; break;
        case 'C': cpu_only = 1; break;     /* the reference runtime alone: CPU chores only (the oracle of the other modes) */
        default: break;
        }
    }
    dtd_nb = nb;
    int pargc = 1; char *pargv0[2] = { argv[0], NULL }; char **pargv = pargv0;
    parsec_context_t *parsec = parsec_init(cores, &pargc, &pargv);
    if( NULL == parsec ) return 2;

    int ngpu = 0, b200 = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL == d || !PARSEC_DEV_IS_GPU(d->type) ) continue;
        ngpu++; b200 += parsec_b200_is_b200_device(d);
    }

    parsec_arena_datatype_t *adt = parsec_matrix_adt_new_rect(parsec_datatype_int32_t, nb, 1, nb);
    parsec_dtd_attach_arena_datatype(parsec, adt, &TILE_FULL);

    int errors[4] = {0, 0, 0, 0};
    uint64_t checksum[4] = {0, 0, 0, 0};
    const char *names[4] = { "memset", "memset_and_read", "new_tile", "pingpong" };
    uint64_t gpu_tasks_before = 0, gpu_tasks = 0;

    if( 0 > parsec_context_start(parsec) ) return 4;
    for( int test = 0; test < 4; test++ ) {
        /* a collection of its own per case, like the reference tests: the runtime owns the coherency of its tiles, host
         * memory must not be rewritten behind its back while a device still holds replicas */
        parsec_matrix_block_cyclic_t dcA;
        parsec_matrix_block_cyclic_init(&dcA, PARSEC_MATRIX_INTEGER, PARSEC_MATRIX_TILE, 0, nb, 1, MT * nb, 1, 0, 0, MT * nb, 1, 1, 1, 1, 1, 0, 0);
        dcA.mat = parsec_data_allocate((size_t)MT * nb * sizeof(int32_t));
        parsec_data_collection_t *A = (parsec_data_collection_t*)&dcA;
        parsec_data_collection_set_key(A, "A");
        parsec_dtd_data_collection_init(A);
        for( int i = 0; i < MT; i++ ) { int32_t *p = tile_ptr(A, i); for( int j = 0; j < nb; j++ ) p[j] = start_value(i, j); }
        parsec_taskpool_t *tp = parsec_dtd_taskpool_new();
        if( 0 > parsec_context_add_taskpool(parsec, tp) ) return 4;
        parsec_task_class_t *tc[3] = { NULL, NULL, NULL };
        const int want_ff = (int)0xFFFFFFFF;
        if( test <= 1 ) {
            tc[0] = parsec_dtd_create_task_class(tp, "memset", PASSED_BY_REF, PARSEC_INOUT | TILE_FULL, sizeof(int), PARSEC_VALUE, PARSEC_DTD_ARG_END);
            /* first added is tested first (dtd_test_cuda_task_insert.c:238-245) */
            if( !cpu_only ) parsec_dtd_task_class_add_chore(tp, tc[0], PARSEC_DEV_CUDA, gpu_memset_fn);
            parsec_dtd_task_class_add_chore(tp, tc[0], PARSEC_DEV_CPU, cpu_memset_fn);
            if( 1 == test ) {
                tc[1] = parsec_dtd_create_task_class(tp, "read", PASSED_BY_REF, PARSEC_INPUT | TILE_FULL, sizeof(int), PARSEC_VALUE, sizeof(int), PARSEC_VALUE, PARSEC_DTD_ARG_END);
                parsec_dtd_task_class_add_chore(tp, tc[1], PARSEC_DEV_CPU, cpu_read_fn);
            }
            for( int i = 0; i < MT; i++ ) {
                /* test 0: alternate host and device like mode GPU|CPU of the reference test; test 1: all on the device */
                const int device = (cpu_only || (0 == test && !gpu_only && 0 == (i & 1))) ? PARSEC_DEV_CPU : PARSEC_DEV_CUDA;
                parsec_dtd_insert_task_with_task_class(tp, tc[0], 1, device, PARSEC_PUSHOUT, PARSEC_DTD_TILE_OF_KEY(A, A->data_key(A, i, 0)),
                                                       PARSEC_DTD_EMPTY_FLAG, &nb, PARSEC_DTD_ARG_END);
                if( 1 == test )
                    parsec_dtd_insert_task_with_task_class(tp, tc[1], 1, PARSEC_DEV_CPU, PARSEC_DTD_EMPTY_FLAG, PARSEC_DTD_TILE_OF_KEY(A, A->data_key(A, i, 0)),
                                                           PARSEC_DTD_EMPTY_FLAG, &nb, PARSEC_DTD_EMPTY_FLAG, &want_ff, PARSEC_DTD_ARG_END);
            }
        } else if( 2 == test ) {
            tc[0] = parsec_dtd_create_task_class(tp, "iota", PASSED_BY_REF, PARSEC_INOUT | TILE_FULL, PARSEC_DTD_ARG_END);
            if( !cpu_only ) parsec_dtd_task_class_add_chore(tp, tc[0], PARSEC_DEV_CUDA, gpu_iota_fn);
            parsec_dtd_task_class_add_chore(tp, tc[0], PARSEC_DEV_CPU, cpu_iota_fn);
            tc[1] = parsec_dtd_create_task_class(tp, "scale", PASSED_BY_REF, PARSEC_INOUT | TILE_FULL, sizeof(int), PARSEC_VALUE, PARSEC_DTD_ARG_END);
            if( !cpu_only ) parsec_dtd_task_class_add_chore(tp, tc[1], PARSEC_DEV_CUDA, gpu_scale_fn);
            parsec_dtd_task_class_add_chore(tp, tc[1], PARSEC_DEV_CPU, cpu_scale_fn);
            const int two = 2;
            const int dev2 = cpu_only ? PARSEC_DEV_CPU : PARSEC_DEV_CUDA;
            for( int i = 0; i < MT; i++ ) {
                /* no pushout between the two: the second task finds the first one's replica on the device */
                parsec_dtd_insert_task_with_task_class(tp, tc[0], 1, dev2, PARSEC_DTD_EMPTY_FLAG, PARSEC_DTD_TILE_OF_KEY(A, A->data_key(A, i, 0)), PARSEC_DTD_ARG_END);
                parsec_dtd_insert_task_with_task_class(tp, tc[1], 1, dev2, PARSEC_PUSHOUT, PARSEC_DTD_TILE_OF_KEY(A, A->data_key(A, i, 0)),
                                                       PARSEC_DTD_EMPTY_FLAG, &two, PARSEC_DTD_ARG_END);
            }
        } else {
            tc[0] = parsec_dtd_create_task_class(tp, "incr", PASSED_BY_REF, PARSEC_INOUT | TILE_FULL, sizeof(int), PARSEC_VALUE, PARSEC_DTD_ARG_END);
            if( !cpu_only ) parsec_dtd_task_class_add_chore(tp, tc[0], PARSEC_DEV_CUDA, gpu_incr_fn);
            parsec_dtd_task_class_add_chore(tp, tc[0], PARSEC_DEV_CPU, cpu_incr_fn);
            const int one = 1;
            for( int k = 0; k < NT; k++ )
                for( int i = 0; i < MT; i++ ) {
                    const int device = (cpu_only || (!gpu_only && ((k + i) & 1))) ? PARSEC_DEV_CPU : PARSEC_DEV_CUDA;
                    parsec_dtd_insert_task_with_task_class(tp, tc[0], 1, device, PARSEC_PUSHOUT, PARSEC_DTD_TILE_OF_KEY(A, A->data_key(A, i, 0)),
                                                           PARSEC_DTD_EMPTY_FLAG, &one, PARSEC_DTD_ARG_END);
                }
        }
        parsec_dtd_data_flush_all(tp, A);
        if( 0 > parsec_taskpool_wait(tp) ) return 5;
        for( int i = 0; i < MT; i++ ) {
            const int32_t *p = tile_ptr(A, i);
            for( int j = 0; j < nb; j++ ) {
                int32_t want;
                switch( test ) {
                case 0: case 1: want = (int32_t)0xFFFFFFFF; break;
                case 2:  want = 2 * j; break;
                default: want = start_value(i, j) + NT; break;
                }
                if( p[j] != want ) { if( errors[test] < 3 ) fprintf(stderr, "%s: A(%d)[%d] = %d, expected %d\n", names[test], i, j, p[j], want); errors[test]++; }
            }
        }
        checksum[test] = fnv1a64(dcA.mat, (size_t)MT * nb * sizeof(int32_t), 0);
        if( 1 == test ) errors[1] += read_errors;
        for( int k = 0; k < 3; k++ ) if( NULL != tc[k] ) parsec_dtd_task_class_release(tp, tc[k]);
        parsec_taskpool_free(tp);
        parsec_dtd_data_collection_fini(A);
        parsec_data_free(dcA.mat);
        parsec_tiled_matrix_destroy((parsec_tiled_matrix_t*)&dcA);
    }
    if( 0 > parsec_context_wait(parsec) ) return 5;

    parsec_b200_stats_t st; memset(&st, 0, sizeof st);
    uint64_t lane = 0, engine = 0, h2d_dma = 0, d2h_dma = 0;
    for( int i = 0; i < (int)parsec_nb_devices; i++ ) {
        parsec_device_module_t *d = parsec_mca_device_get(i);
        if( NULL == d || !PARSEC_DEV_IS_GPU(d->type) ) continue;
        gpu_tasks += d->executed_tasks;
        if( PARSEC_SUCCESS == parsec_b200_get_stats(d, &st) ) { lane += st.tasks_lane; engine += st.tasks_engine; h2d_dma += st.bytes_h2d_dma; d2h_dma += st.bytes_d2h_dma; }
    }
    (void)gpu_tasks_before;
    const int total = errors[0] + errors[1] + errors[2] + errors[3];
    printf("{\"app\": \"dtd_b200\", \"checksums\": [\"%016lx\", \"%016lx\", \"%016lx\", \"%016lx\"], \"tiles\": %d, \"nb\": %d, \"hops\": %d, \"opaque\": %d, \"gpu_modules\": %d, \"b200_modules\": %d, "
           "\"errors\": {\"memset\": %d, \"memset_and_read\": %d, \"new_tile\": %d, \"pingpong\": %d}, \"total_errors\": %d, "
           "\"executed_on_gpu\": %lu, \"tasks_engine\": %lu, \"tasks_lane\": %lu, \"bytes_h2d_dma\": %lu, \"bytes_d2h_dma\": %lu}\n",
           (unsigned long)checksum[0], (unsigned long)checksum[1], (unsigned long)checksum[2], (unsigned long)checksum[3],
           MT, nb, NT, opaque, ngpu, b200, errors[0], errors[1], errors[2], errors[3], total,
           (unsigned long)gpu_tasks, (unsigned long)engine, (unsigned long)lane, (unsigned long)h2d_dma, (unsigned long)d2h_dma);

    parsec_dtd_free_arena_datatype(parsec, TILE_FULL);
    parsec_fini(&parsec);
    return total ? 1 : 0;
}
