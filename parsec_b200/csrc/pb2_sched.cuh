// pb2_sched.cuh -- device-resident scheduling state and primitives shared by the engine kernels
// (HBM-body kernel in pb2_engine.cu, tensor-core kernel in pb2_gemm.cuh).
#pragma once
#include "../../include/pb2_engine.h"
#include "pb2_dev_utils.cuh"
#include "pb2_bodies.cuh"

namespace pb2 {

constexpr int32_t kEmpty = -1;
constexpr int32_t kDoneOK = 1;
constexpr int32_t kDoneTimeout = 2;
constexpr int32_t kDoneBadBody = 3;

// Hot control words, one per 128-byte line so that atomics on them do not false-share.
struct alignas(128) Line { unsigned long long v; unsigned long long pad[15]; };
struct Ctl {
    Line head;       // pop tickets handed out
    Line tail;       // push tickets handed out
    Line evt;        // global event counter (start/end sequence numbers)
    Line retired;    // tasks retired
    Line done;       // 0 running, kDone*
    Line progress_ns;// globaltimer of the last retirement (watchdog)
    Line bytes_h2d, bytes_d2d, bytes_d2h, stage_ins, body_errors;
};

struct WinDev {
    const pb2_task_t* tasks;
    const uint32_t*   succ;
    pb2_tile_t*       tiles;
    int32_t*          dep;
    int32_t*          ring;
    Ctl*              ctl;
    int32_t*          retire_log;
    uint32_t*         start_seq;
    uint32_t*         end_seq;
    uint32_t*         seen_version;
    unsigned long long* result;
    int32_t*          worker;
    uint32_t          cap_mask;
    int32_t           ntasks;
    int32_t           ntiles;
    int32_t           stage_mode;
    unsigned long long timeout_ns;
    int32_t*          parts_left;     // HBM windows: parts of a task still running (wide tasks)
    // remote out-edges (other GPUs' windows), see pb2_window_set_remote
    const int32_t*    rs_begin;
    const int32_t*    rs_rank;
    const uint32_t*   rs_target;
    const struct PeerWin* peers;
    int32_t           shared;         // scheduling arrays are written by peers: poll / publish at system scope
};

struct PeerWin { int32_t* dep; int32_t* ring; Ctl* ctl; uint32_t cap_mask; int32_t pad; };

// A task whose tiles are large is executed as several PARTS (byte slices of its tiles) by different workers;
// the number of parts (1..32) is kept in bits 3..7 of the DEVICE copy of pb2_task_t::flags, ring entries are
// (part << 27) | task.  One tile at HBM speed needs the whole GPU: a 4 MiB tile is 1.3 us of the machine, not
// 1 ms of one CTA.
#define PB2_TASK_NPARTS(flags)  ((((int)(flags)) >> 3) + 1)

// ---------------------------------------------------------------------------------------------
// scheduling primitives shared by the HBM and the GEMM engine kernels
// ---------------------------------------------------------------------------------------------

// One thread: take the next pop ticket and wait for its slot.  Returns a task id, or kEmpty when the
// window is finished (or aborted).  Ticket order == push order, i.e. a strict FIFO ready queue.
__device__ __forceinline__ int32_t pop_task(const WinDev& w) {
    const uint32_t ticket = (uint32_t)atomicAdd(&w.ctl->head.v, 1ull);
    int32_t* slot = &w.ring[ticket & w.cap_mask];
    uint32_t spins = 0;
    int32_t id;
    while ((id = (w.shared ? ld_acquire_sys(slot) : ld_acquire_gpu(slot))) == kEmpty) {
        if (ld_relaxed_gpu(reinterpret_cast<const int32_t*>(&w.ctl->done.v)) != 0) return kEmpty;
        if ((++spins & 1023u) == 0) {
            // watchdog: a DAG whose dependency counts are wrong would spin forever
            const unsigned long long last = *reinterpret_cast<volatile unsigned long long*>(&w.ctl->progress_ns.v);
            // signed: %globaltimer read on another SM can be slightly behind the value a retiring SM just stored
            if ((long long)(globaltimer_ns() - last) > (long long)w.timeout_ns) {
                st_relaxed_gpu(reinterpret_cast<int32_t*>(&w.ctl->done.v), kDoneTimeout);
                return kEmpty;
            }
        }
        __nanosleep(spins < 64 ? 32 : 256);
    }
    return id;
}

// Whole warp: release the out-edges of task t (parsec_release_dep_fct semantics), push the newly
// ready successors.  Must be called after a __threadfence() that follows the body's stores.
__device__ __forceinline__ void release_successors_warp(const WinDev& w, const pb2_task_t& t) {
    const int lane = threadIdx.x & 31;
    for (int e0 = 0; e0 < t.succ_count; e0 += 32) {
        const int e = e0 + lane;
        bool ready = false;
        int32_t sid = -1;
        if (e < t.succ_count) {
            const uint32_t s = w.succ[t.succ_begin + e];
            sid = PB2_SUCC_TASK(s);
            const pb2_task_t& st = w.tasks[sid];
            if (st.flags & PB2_TASK_DEPS_MASK) {
                // parsec_update_deps_with_mask, parsec.c:1656-1720: OR the destination flow bit, the
                // task is ready when (word & goal) == goal; each bit is set exactly once (:1688 assert)
                const int32_t bit = 1 << PB2_SUCC_FLOW(s);
                const int32_t old = atomicOr(&w.dep[sid], bit);
                ready = (((old | bit) & st.dep_goal) == st.dep_goal) && ((old & st.dep_goal) != st.dep_goal);
            } else {
                // parsec_update_deps_with_counter, parsec.c:1609-1654: fetch_dec, ready at 0
                ready = (atomicSub(&w.dep[sid], 1) == 1);
            }
        }
        // a ready successor contributes one ring entry per part: exclusive scan of the part counts over the warp
        const int nparts = ready ? PB2_TASK_NPARTS(w.tasks[sid].flags) : 0;
        int incl = nparts;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&w.ctl->tail.v, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (int p = 0; p < nparts; ++p)
                st_release_gpu(&w.ring[((uint32_t)base + (uint32_t)(incl - nparts + p)) & w.cap_mask], (int32_t)PB2_SUCC_MAKE(sid, p));
        }
    }
}

// Whole warp: release the out-edges that lead into other GPUs' windows.  The activation message of the reference
// (remote_dep_mpi.c:1860 remote_dep_mpi_recv_activate -> release of the local successors) becomes a system-scope
// atomic on the peer's dependency word and, when it reaches zero, ring entries written into the peer's HBM over
// NVLink.  The tile itself is pulled by the peer's worker from this GPU's slot when the task runs (stage_in_flow).
__device__ __forceinline__ void release_remote_warp(const WinDev& w, int32_t id) {
    if (!w.rs_begin) return;
    const int lane = threadIdx.x & 31;
    const int32_t b = w.rs_begin[id], e1 = w.rs_begin[id + 1];
    if (b == e1) return;
    __threadfence_system();            // our tile bytes are visible to the peers before they can see the release
    for (int32_t e0 = b; e0 < e1; e0 += 32) {
        const int32_t e = e0 + lane;
        if (e < e1) {
            const PeerWin pw = w.peers[w.rs_rank[e]];
            const uint32_t tgt = w.rs_target[e];
            const int32_t sid = PB2_SUCC_TASK(tgt);
            if (atomicSub_system(&pw.dep[sid], 1) == 1) {
                const int nparts = PB2_SUCC_FLOW(tgt) + 1;
                const unsigned long long base = atomicAdd_system(&pw.ctl->tail.v, (unsigned long long)nparts);
                for (int p = 0; p < nparts; ++p)
                    st_release_sys(&pw.ring[((uint32_t)base + (uint32_t)p) & pw.cap_mask], (int32_t)PB2_SUCC_MAKE(sid, p));
            }
        }
    }
}

// One thread: append to the retire log; returns true when this was the last task of the window.
__device__ __forceinline__ bool retire_task(const WinDev& w, int32_t id) {
    const uint32_t seq = (uint32_t)atomicAdd(&w.ctl->retired.v, 1ull);
    w.retire_log[seq] = id;
    *reinterpret_cast<volatile unsigned long long*>(&w.ctl->progress_ns.v) = globaltimer_ns();
    return (int32_t)(seq + 1) == w.ntasks;
}

// ---------------------------------------------------------------------------------------------
// stage-in / stage-out of one flow by the whole CTA
// ---------------------------------------------------------------------------------------------
// Thread 0 decides (s_decide[0]): 1 = this CTA moves the tile, 0 = already valid (possibly after waiting)
__device__ __forceinline__ void stage_in_flow(const WinDev& w, pb2_tile_t* tile, uint8_t access, int* s_decide) {
    if (threadIdx.x == 0) {
        int decide = 0;
        if (access & PB2_FLOW_ACCESS_READ) {
            // parsec_device_data_stage_in, device_gpu.c:1799-2165: only a READ access needs the bytes;
            // "finally we'll just overwrite w/o read" (data.c:427) for WRITE-only flows.
            int32_t st = atomicCAS(&tile->state, PB2_TILE_INVALID, PB2_TILE_STAGING);
            if (st == PB2_TILE_INVALID) {
                decide = 1;
            } else {
                // another worker is moving it: "data copy is already under transfer" (:1873-1884)
                while (st != PB2_TILE_VALID) { __nanosleep(64); st = ld_acquire_gpu(&tile->state); }
            }
        }
        *s_decide = decide;
    }
    __syncthreads();
    if (*s_decide) {
        cta_copy<true>(tile->dev_ptr, tile->src_ptr, tile->bytes);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            st_release_gpu(&tile->state, PB2_TILE_VALID);   // COMPLETE_TRANSFER (:2358-2573)
            atomicAdd(tile->src_kind == PB2_SRC_PEER ? &w.ctl->bytes_d2d.v : &w.ctl->bytes_h2d.v,
                      (unsigned long long)tile->bytes);
            atomicAdd(&w.ctl->stage_ins.v, 1ull);
        }
    }
    __syncthreads();
}

}  // namespace pb2
