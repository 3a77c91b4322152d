/*
 * pb2_engine.h -- C ABI of the B200 device-side DAG execution engine (layer L0).
 *
 * One engine instance drives one GPU.  It replaces, for tasks whose incarnation
 * is GPU, the host-driven stream pipeline of the reference
 *   parsec/mca/device/device_gpu.c:3375 (parsec_device_kernel_scheduler)
 *   parsec/mca/device/device_gpu.c:2592 (parsec_device_progress_stream)
 *   parsec/mca/device/device_gpu.c:2745/2873/2943 (kernel_push / _exec / _pop)
 * and the host dependency release of
 *   parsec/parsec.c:1609/1656/1749/1836 (update_deps_with_counter / _with_mask,
 *   release_local_OUT_dependencies, release_dep_fct)
 * by ONE persistent sm_100a kernel per "window" of the DAG: workers (CTAs) pop
 * ready task descriptors from a device-resident ring, stage tiles in from
 * host-pinned / peer memory, run the body, release successors with device
 * atomics and append to a retire log.  No host round trip per task or per edge.
 *
 * Plain C, plain pointers and sizes; no torch / C++ types cross this boundary.
 * The reference-shaped module API (parsec_device_module_t, parsec_gpu_task_t,
 * kernel_scheduler, ...) is layered on top of this file in pb2_device.h.
 */
#ifndef PB2_ENGINE_H
#define PB2_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes: same values as parsec/include/parsec/constants.h:14-26 */
#define PB2_SUCCESS                   0
#define PB2_ERROR                    -1
#define PB2_ERR_OUT_OF_RESOURCE      -2
#define PB2_ERR_NOT_FOUND            -3
#define PB2_ERR_BAD_PARAM            -4
#define PB2_ERR_EXISTS               -5
#define PB2_ERR_NOT_IMPLEMENTED      -6
#define PB2_ERR_NOT_SUPPORTED        -7
#define PB2_ERR_VALUE_OUT_OF_BOUNDS  -8
#define PB2_ERR_TRUNCATE             -9
#define PB2_ERR_DEVICE              -10

/* Flow access bits: same values as
 * parsec/include/parsec/parsec_description_structures.h:62-67 */
#define PB2_FLOW_ACCESS_NONE   0x00
#define PB2_FLOW_ACCESS_READ   0x04
#define PB2_FLOW_ACCESS_WRITE  0x08
#define PB2_FLOW_ACCESS_RW     0x0c
#define PB2_FLOW_PUSHOUT       0x40   /* engine-private: D2H the flow after the body (gpu_task->pushout bit) */

/* Task bodies the persistent kernel can run in place (the "incarnations").
 * HBM-bound bodies are coalesced 16-byte vector loops; GEMM is tcgen05. */
enum pb2_body_e {
    PB2_BODY_NOP        = 0,  /* empty body: tests/runtime/scheduling/ep.jdf:36-40                       */
    PB2_BODY_FILL_I32   = 1,  /* flow0[:] = iparam[0]          (Ex05 TaskBcast: "*Aint = k", tile-wide)   */
    PB2_BODY_CHECK_I32  = 2,  /* result = #elements of flow0 != iparam[0]; sum   (Ex05 TaskRecv)           */
    PB2_BODY_INCR_I32   = 3,  /* flow0[:] += iparam[0]         (Ex02 "*Aint += 1"; rtt.jdf PING)           */
    PB2_BODY_ADD_IOTA_I32 = 4,/* flow0[i] += i                 (tests/runtime/cuda/ping_kernel.cu:15)      */
    PB2_BODY_SCALE_I32  = 5,  /* flow0[:] *= iparam[0]         (dtd_test_new_tile_cuda_kernels.cu:30)      */
    PB2_BODY_IOTA_I32   = 6,  /* flow0[i] = i                  (dtd_test_new_tile_cuda_kernels.cu:17)      */
    PB2_BODY_COPY       = 7,  /* flow1[:] = flow0[:]           (write_check.cu "A3=A2")                    */
    PB2_BODY_FILL_F32   = 8,  /* flow0[:] = fparam                                                         */
    PB2_BODY_CHECK_F32  = 9,  /* result = #elements of flow0 != fparam                                     */
    PB2_BODY_INCR_F32   = 10, /* flow0[:] += fparam            (config 4: "T[:] += 1")                     */
    PB2_BODY_AXPY_F32   = 11, /* flow1[:] += fparam*flow0[:]                                               */
    PB2_BODY_MEMSET_U8  = 12, /* flow0 bytes = iparam[0]&0xff  (get_best_device_check.jdf:82 cudaMemset)   */
    PB2_BODY_ADD_AT_I32 = 13, /* flow0[iparam[0]] += iparam[1]  (ping_kernel.cu:13-21 pong_kernel <<<1,1>>>,
                               *                                 ptg_pingpong.jdf:72-73 TOKEN_CPU)          */
    PB2_BODY_GEMM_BF16  = 16, /* flow2 (C, M x N row-major bf16) += flow0 (A, M x K row-major) *
                               * flow1 (B, N x K row-major == K x N column-major), fp32 accumulate in TMEM
                               * iparam[0]=M, iparam[1]=N, iparam[2]=K (each tile edge)                     */
    PB2_BODY_USER       = 31, /* host-side only: the chore is a user `submit` callback that enqueues its own CUDA work
                               * on a stream (device_gpu.h:49-51); such tasks never enter an engine window           */
    PB2_BODY_MAX        = 32
};

#define PB2_MAX_FLOWS 4

/* One task, 64 bytes, read-only on the device.  Mirrors what the reference keeps
 * in parsec_task_t (parsec_internal.h:551-563: task_class, locals[], data[], priority)
 * plus parsec_gpu_task_t (device_gpu.h:117-143: pushout, nb_flows, flow_info[]). */
typedef struct pb2_task_s {
    int32_t  dep_goal;      /* counter mode: #task-sourced input deps (parsec.c:1471-1556);
                             * mask mode: tc->dependencies_goal (parsec.c:1656-1720)                        */
    int32_t  succ_begin;    /* first entry in succ[]                                                         */
    int32_t  succ_count;    /* number of out-edges (iterate_successors fan-out)                              */
    int32_t  priority;      /* task priority (larger first when a priority lane is used)                      */
    uint8_t  body;          /* enum pb2_body_e                                                                */
    uint8_t  nb_flows;
    uint8_t  flags;         /* PB2_TASK_* */
    uint8_t  class_id;      /* task_class_id, for traces                                                      */
    int32_t  tile[PB2_MAX_FLOWS];    /* tile ids, -1 = CTL / unused                                           */
    uint8_t  access[PB2_MAX_FLOWS];  /* PB2_FLOW_ACCESS_* | PB2_FLOW_PUSHOUT                                  */
    int32_t  iparam[3];     /* body immediates                                                               */
    float    fparam;
    int32_t  locals[2];     /* first two locals of the task (k, n / m, n) for traces                          */
} pb2_task_t;                /* sizeof == 64 */

#define PB2_TASK_DEPS_MASK  0x01   /* dep word is a bit mask (PARSEC_USE_DEPS_MASK) instead of a counter     */

/* succ[] entry: low 27 bits = successor task id, high 5 bits = destination flow index */
#define PB2_SUCC_MAKE(task, flow)  (((uint32_t)(flow) << 27) | (uint32_t)(task))
#define PB2_SUCC_TASK(s)           ((int32_t)((s) & 0x07ffffffu))
#define PB2_SUCC_FLOW(s)           ((int32_t)((s) >> 27))

/* Device-side replica of one parsec_data_t on this GPU (data_internal.h:30-85), 32 bytes. */
typedef struct pb2_tile_s {
    void*    dev_ptr;   /* slot in this GPU's HBM (parsec_data_copy_t::device_private of the GPU copy)       */
    void*    src_ptr;   /* device-visible address of the source/home copy: cudaHostRegister'ed host
                         * memory (device_cuda_module.c:183-212) or a peer GPU's slot                        */
    uint32_t bytes;     /* span to move: min(src span, dst span), device_gpu.c:1639-1644                     */
    int32_t  state;     /* PB2_TILE_*                                                                        */
    uint32_t version;   /* parsec_data_copy_t::version of the GPU copy                                       */
    int32_t  src_kind;  /* PB2_SRC_HOST / PB2_SRC_PEER: which statistic the stage-in is charged to
                         * (data_in_from_device[src], device_gpu.c:2133)                                    */
} pb2_tile_t;

#define PB2_SRC_HOST 0
#define PB2_SRC_PEER 1
#define PB2_SRC_PUSH 2      /* the producer's worker writes the bytes into this slot over NVLink and publishes the state:
                             * a local task never pulls it, it finds it VALID (pb2_window_set_push) */

#define PB2_TILE_INVALID   0   /* PARSEC_DATA_COHERENCY_INVALID, must be staged in before a READ             */
#define PB2_TILE_STAGING   1   /* PARSEC_DATA_STATUS_UNDER_TRANSFER                                          */
#define PB2_TILE_VALID     2   /* SHARED/OWNED + COMPLETE_TRANSFER                                           */

typedef struct pb2_engine_params_s {
    int32_t  workers_per_sm;   /* CTAs per SM for HBM-body windows (default 4)                               */
    int32_t  threads;          /* threads per CTA for HBM-body windows (default 256)                         */
    int32_t  max_workers;      /* 0 = all; 1 = single worker => deterministic FIFO order (tests)             */
    int32_t  stage_mode;       /* tile mover of the HBM-body kernels: 0 = TMA bulk copy (cp.async.bulk through a
                                * shared-memory ring, default), 1 = SIMT 16-byte LDG/STG loops                      */
    int32_t  queue_policy;     /* 0 = FIFO ring, 1 = successors-first (hot ring before FIFO ring)             */
    int32_t  timeout_ms;       /* device-side watchdog: a window that makes no progress for this long aborts
                                * (default 20000); a malformed DAG must never hang the GPU                   */
    int32_t  gemm_mode;        /* 0 = CTA pairs (cta_group::2) + fused k-chains (default), 1 = v1 single-CTA kernel,
                                * 2 = CTA pairs, every task flushes C (per-task bf16 rounding, as the oracle)       */
    int32_t  part_bytes;       /* HBM bodies: a task whose largest tile exceeds this many bytes is run as up to 512
                                * parts (byte slices) by different workers (default 256 KiB, <0 = never split);
                                * pb2_engine_set_part_bytes changes it for the windows created afterwards       */
} pb2_engine_params_t;

typedef struct pb2_engine_info_s {
    int32_t  cuda_device;
    int32_t  sm_count;
    int32_t  cc_major, cc_minor;
    int32_t  nworkers;         /* grid size used for HBM-body windows                                        */
    int32_t  nworkers_gemm;    /* grid size used for GEMM windows                                            */
    int32_t  can_map_host;
    int32_t  reserved;
    uint64_t total_mem;
    uint64_t free_mem;
} pb2_engine_info_t;

/* What one window run produced; device counters mirror device.h:165-171 statistics. */
typedef struct pb2_window_stats_s {
    uint64_t tasks_retired;
    uint64_t bytes_h2d;        /* data_in_from_device[host]: bytes staged in by the kernel                   */
    uint64_t bytes_d2d;        /* bytes staged in from a peer GPU slot                                       */
    uint64_t bytes_d2h;        /* data_out_to_host: pushout bytes                                            */
    uint64_t stage_ins;        /* nb_data_faults (count)                                                     */
    uint64_t body_errors;      /* sum of CHECK body mismatches                                               */
    float    kernel_ms;        /* CUDA-event time of the window kernel on its launch stream                  */
    float    reset_ms;
} pb2_window_stats_t;

typedef struct pb2_engine_s pb2_engine_t;
typedef struct pb2_window_s pb2_window_t;

/* --- engine life cycle (parsec_cuda_module_init / _fini, device_cuda_module.c:406,660) --- */
int  pb2_engine_create(pb2_engine_t** engine, int cuda_device, const pb2_engine_params_t* params);
int  pb2_engine_destroy(pb2_engine_t* engine);
int  pb2_engine_info(pb2_engine_t* engine, pb2_engine_info_t* info);
const char* pb2_engine_last_error(pb2_engine_t* engine);

/* --- device memory (parsec_device_memory_reserve device_gpu.c:866; cuda memory_allocate/free) --- */
int  pb2_engine_malloc(pb2_engine_t* engine, size_t bytes, void** dev_ptr);
int  pb2_engine_free(pb2_engine_t* engine, void* dev_ptr);
/* cudaHostRegister(Portable|Mapped) of a whole collection + its device-visible alias
 * (parsec_cuda_memory_register device_cuda_module.c:183-212). Idempotent per ptr. */
int  pb2_engine_host_register(pb2_engine_t* engine, void* host_ptr, size_t bytes, void** dev_alias);
int  pb2_engine_host_unregister(pb2_engine_t* engine, void* host_ptr);
int  pb2_engine_memcpy_h2d(pb2_engine_t* engine, void* dev, const void* host, size_t bytes);
/* Copy-engine prefetch for the NEXT window that is armed: queued on a DMA stream that runs beside the window in
 * flight; pb2_window_arm makes the engine stream wait for everything queued here.  `host` must be pinned
 * (pb2_engine_host_register).  One call moves `rows` equally spaced tiles (cudaMemcpy2DAsync; a zone heap with
 * 512 KiB units holds 256 KiB tiles at a 512 KiB pitch).  Large runs reach the PCIe DMA rate (54 GB/s on this box) where worker
 * CTAs reading host memory reach 46 GB/s; the caller marks the tiles it prefetched PB2_TILE_VALID. */
int  pb2_engine_prefetch_h2d(pb2_engine_t* engine, void* dev, size_t dev_pitch, const void* host, size_t host_pitch,
                             size_t width_bytes, size_t rows);   /* rows tiles of width_bytes, constant strides */
int  pb2_engine_memcpy_d2h(pb2_engine_t* engine, void* host, const void* dev, size_t bytes);
int  pb2_engine_synchronize(pb2_engine_t* engine);
/* Enqueue all engine work on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g. the stream NCCL
 * collectives are ordered against; NULL restores the engine's own non-blocking stream. */
int  pb2_engine_set_stream(pb2_engine_t* engine, void* cuda_stream);
void* pb2_engine_get_stream(pb2_engine_t* engine);       /* the cudaStream_t engine work is enqueued on */

/* n independent copies dst[i][0:bytes[i]] = src[i][..] in ONE kernel launch (workers grid-stride over the list);
 * either side may be HBM, a peer GPU or cudaHostRegister'ed host memory (device-visible alias).  This is what the
 * module uses to write a batch of dirty tiles home (parsec_gpu_create_w2r_task batches <= 20 copies per
 * pseudo-task and pays one cudaMemcpyAsync + event per tile, transfer_gpu.c:224-304).  Stream-ordered. */
int  pb2_engine_copy_batch(pb2_engine_t* engine, void* const* dst, const void* const* src, const uint64_t* bytes, int32_t n);

/* --- multi-GPU (one process per GPU): peer-visible memory ---
 * 64-byte CUDA IPC handles of cudaMalloc'ed memory; a peer process opens them to get a pointer its kernels can
 * load from / store to over NVLink (parsec_cuda_all_devices_attached enables the same peer access inside one
 * process, device_cuda_module.c:144-181). */
int  pb2_engine_ipc_export(pb2_engine_t* engine, void* dev_ptr, unsigned char handle[64]);
int  pb2_engine_ipc_open(pb2_engine_t* engine, const unsigned char handle[64], void** dev_ptr);
int  pb2_engine_ipc_close(pb2_engine_t* engine, void* dev_ptr);
/* cudaDeviceEnablePeerAccess towards another GPU of this process (parsec_cuda_all_devices_attached,
 * device_cuda_module.c:144-181): afterwards tile sources may point into that GPU's memory. */
int  pb2_engine_enable_peer(pb2_engine_t* engine, int peer_cuda_device);
/* The in-kernel bodies as ONE stand-alone kernel launch on a caller-owned stream: what a BODY [type=CUDA] that names
 * an engine body enqueues when it runs under a device module that is not the engine's (the reference's stream
 * engine drives it then, device_gpu.c:2873-2934).  ptrs/bytes: device address and span of each body argument.
 * CHECK bodies add their mismatch count to a device counter read (and optionally cleared) by
 * pb2_body_launch_errors. */
int  pb2_body_launch(void* cuda_stream, int body, int nb_args, void* const* ptrs, const uint64_t* bytes,
                     const int32_t* iparam3, float fparam);
int  pb2_body_launch_errors(uint64_t* errors, int reset);
/* all windows created after this call keep their scheduling arrays in IPC-exportable memory and treat their tasks'
 * dependency goals as counting in-edges from other GPUs too.  next_rs_begin (may be NULL; ntasks+1 entries, must stay
 * valid until the next pb2_window_create returns) is the remote out-edge CSR of the next window: a task with remote
 * successors is never fused into the middle of a GEMM k-chain unit. */
int  pb2_engine_set_shared_windows(pb2_engine_t* engine, int on, const int32_t* next_rs_begin);
/* part size for the windows created from now on (a serial chain of large tiles wants small parts: a 64-thread
 * worker keeps only 4 KiB in flight; wide DAGs want one part per tile) */
int  pb2_engine_set_part_bytes(pb2_engine_t* engine, int32_t part_bytes);
/* granularity of cooperative stage-in (default 64 KiB, <= 0: whole tiles): a tile that has to be staged in is cut into
 * slices of this size and every worker that needs the tile pulls the slices nobody has claimed yet */
int  pb2_engine_set_stage_slice_bytes(pb2_engine_t* engine, int32_t bytes);

/* --- one window of the DAG ---
 * tasks[ntasks], succ[nsucc] (CSR via succ_begin/succ_count), tiles[ntiles] and the ids of
 * the tasks that are ready at submission (startup tasks, parsec.c:1724-1740).
 * 'kind' selects the kernel instantiation: 0 = HBM bodies, 1 = tensor-core GEMM bodies. */
int  pb2_window_create(pb2_engine_t* engine, pb2_window_t** window, int kind,
                       const pb2_task_t* tasks, int32_t ntasks,
                       const uint32_t* succ, int32_t nsucc,
                       const pb2_tile_t* tiles, int32_t ntiles,
                       const int32_t* ready, int32_t nready);
int  pb2_window_destroy(pb2_window_t* window);
/* (re)arm dependency words, ring, counters, tile states; then launch; both are stream-ordered */
int  pb2_window_launch(pb2_window_t* window);
/* block until the window retired all its tasks (or the watchdog tripped); fills stats */
int  pb2_window_wait(pb2_window_t* window, pb2_window_stats_t* stats);
/* per-task outputs, valid after wait; any pointer may be NULL */
int  pb2_window_results(pb2_window_t* window,
                        int32_t*  retire_order,  /* [ntasks] task ids in completion order               */
                        uint32_t* start_seq,     /* [ntasks] global event number when the task started  */
                        uint32_t* end_seq,       /* [ntasks] global event number when it retired        */
                        uint32_t* seen_version,  /* [ntasks*PB2_MAX_FLOWS] tile version seen per flow   */
                        uint64_t* result,        /* [ntasks] body result (CHECK: mismatches<<40 | sum)   */
                        int32_t*  worker,        /* [ntasks] CTA that ran the task                       */
                        pb2_tile_t* tiles_out);  /* [ntiles] final tile table (state, version)           */

/* --- windows that release dependencies of tasks living in OTHER GPUs' windows (remote_dep edges, remote_dep.h:42-58,
 * without the host: the activation is a device atomic on the peer's dependency word plus a ring write over NVLink).
 * Handle of this window's scheduling arrays, to be given to the peers: */
typedef struct pb2_window_handle_s {
    unsigned char dep[64], ring[64], ctl[64], tiles[64];
    uint32_t cap_mask;
    int32_t  ntasks;
    int32_t  entry_kind;   /* 0: dep words and ring entries are per task, 1: per fused GEMM unit */
    int32_t  ntiles;       /* entries of the exported tile table (producer-side pushes publish tile states in it) */
} pb2_window_handle_t;
/* entry[t] = what a remote producer must put in rs_target to release task t of THIS window (index of the
 * dependency word + number of ring entries, in this window's own encoding).  Exchange it with the peers. */
int  pb2_window_task_entries(pb2_window_t* window, int32_t* entry);
int  pb2_window_export(pb2_window_t* window, pb2_window_handle_t* handle);
/* remote out-edges of this window: for task t, entries rs_begin[t] .. rs_begin[t+1]-1 of (rank[], target[]) where
 * target = the destination window's pb2_window_task_entries value of the destination task; remote successors are
 * counter-mode.
 * peers[r] is rank r's exported handle (peers[my_rank] is ignored). */
int  pb2_window_set_remote(pb2_window_t* window, int32_t my_rank, int32_t nranks, const pb2_window_handle_t* peers,
                           const int32_t* rs_begin, const int32_t* rs_rank, const uint32_t* rs_target, int32_t nrs);
/* two-phase launch for windows that are released into by peers: every rank arms, all ranks synchronise, every
 * rank starts.  pb2_window_launch == arm + start. */
int  pb2_window_arm(pb2_window_t* window);
int  pb2_window_start(pb2_window_t* window);

/* --- splitting one window over the GPUs of a box (host logic, no device work) ------------------------------------
 * What remote_dep.c does per task at run time (parsec_remote_dep_activate, remote_dep.c:451: which ranks own the
 * successors of this task, per output flow) is done once per window: every edge whose end points live on different
 * ranks becomes a remote edge of the producer's window, and the consumer's flow gets a tile descriptor that pulls the
 * producer's copy over NVLink (src_kind PEER) the first time a local task needs that version.
 *   task_rank[t] : rank that runs task t (rank_of of its affinity datum, two_dim_rectangle_cyclic.c:258-286)
 *   tile_rank[i] : rank whose slab holds the initial / final copy of tile i
 * A rank that overwrites a tile it has sent is held back by a write-after-read edge from the remote readers, the
 * rule parsec_dtd_ordering_correctly applies to local readers (insert_function.c:2603). */
typedef struct pb2_partition_s pb2_partition_t;
typedef struct pb2_partition_sizes_s {
    int32_t  ntasks, nsucc, ntiles, nready, nremote, nslots;
    uint64_t slab_bytes;           /* bytes of the rank's tile slab (every slot 256-byte aligned) */
} pb2_partition_sizes_t;
int  pb2_partition_create(pb2_partition_t** partition, const pb2_task_t* tasks, int32_t ntasks,
                          const uint32_t* succ, int32_t nsucc, const pb2_tile_t* tiles, int32_t ntiles,
                          const int32_t* ready, int32_t nready, const int32_t* task_rank, const int32_t* tile_rank,
                          int32_t nranks, int32_t part_bytes);
int  pb2_partition_sizes(const pb2_partition_t* partition, int32_t rank, pb2_partition_sizes_t* sizes);
/* slab_base[r]: address of rank r's slab as seen from `rank` (own allocation / IPC mapping).  Arrays sized by
 * pb2_partition_sizes; rs_begin has ntasks+1 entries; rs_target[e] is the LOCAL TASK ID in rank rs_rank[e]'s part
 * (translate it with that rank's pb2_window_task_entries before pb2_window_set_remote); global_id[t] is the id the
 * local task had in the input; part_bytes of pb2_partition_create is reserved (pass 0);
 * slot_tile / slot_offset describe the slab (global tile id and byte offset of each slot). */
int  pb2_partition_get(const pb2_partition_t* partition, int32_t rank, const uint64_t* slab_base,
                       pb2_task_t* tasks, uint32_t* succ, pb2_tile_t* tiles, int32_t* ready,
                       int32_t* rs_begin, int32_t* rs_rank, uint32_t* rs_target, int32_t* global_id,
                       int32_t* slot_tile, uint64_t* slot_offset);
/* Producer-side push.  A version a rank reads out of another rank's slot can be written into the reader's slot by the
 * PRODUCER, right after its body, when nothing on the reader's rank used that slot before (no write-after-read hazard):
 * posted stores over NVLink instead of a pull whose every chunk pays a round trip.  pb2_partition_set_push(p, 1) makes
 * pb2_partition_get describe such tiles with src_kind PB2_SRC_PUSH; pb2_partition_get_push returns, for the tasks of
 * `rank`, what each has to push (CSR ps_begin[ntasks+1]); give it to the window with pb2_window_set_push AFTER
 * pb2_window_set_remote (the destination tile tables come from the peers' handles). */
typedef struct pb2_push_s {
    uint64_t dst;          /* destination slot, as seen from the pushing rank (slab_base[rank] + offset)             */
    uint32_t bytes;
    int32_t  src_tile;     /* the pushing task's own descriptor of the tile                                           */
    int32_t  rank;         /* destination rank                                                                        */
    int32_t  desc;         /* destination descriptor (index in that rank's tile table)                                */
    int32_t  pad[2];
} pb2_push_t;
int  pb2_partition_set_push(pb2_partition_t* partition, int on);
int  pb2_partition_push_count(const pb2_partition_t* partition, int32_t rank, int32_t* npush);
int  pb2_partition_get_push(const pb2_partition_t* partition, int32_t rank, const uint64_t* slab_base,
                            int32_t* ps_begin, pb2_push_t* push);
int  pb2_window_set_push(pb2_window_t* window, const int32_t* ps_begin, const pb2_push_t* push, int32_t npush);
void pb2_partition_destroy(pb2_partition_t* partition);
const char* pb2_partition_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PB2_ENGINE_H */
