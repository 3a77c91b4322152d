/*
 * orc_dtd.c -- ORACLE (test infrastructure): the dependency graph Dynamic Task Discovery builds.
 *
 * Restates the per-tile user chain of parsec/interfaces/dtd:
 *   insert_function.c:3006-3310  parsec_insert_dtd_task: every tile remembers last_writer and last_user;
 *        a new flow is chained behind last_user, its parent is last_writer; a flow whose tile was never
 *        written needs nothing from another task (the reference runs a Fake_FIRST_OUT CPU task that hands
 *        tile->data_copy over, :3055-3073, or takes tile->data_copy directly, :3263-3270); the same tile
 *        used twice by one task counts the later flow as satisfied (:3140-3196);
 *   overlap_strategies.c:139-352 parsec_dtd_ordering_correctly: when a writer completes, every consecutive
 *        INPUT user behind it is released, and so is the next writer behind those readers;
 *   insert_function.c:2102-2118  data_lookup_of_dtd_task: that next writer returns HOOK_RETURN_AGAIN while
 *        the copy still has readers (readers counted in overlap_strategies.c:300-304, released :197-203).
 * Net partial order (what a device mirror must enforce without a host round trip):
 *        reader  waits for the last writer of the tile (if any);
 *        writer  waits for the last writer AND for every reader inserted since that writer (WAR).
 * flow_count (insert_function.c:2962-2976) becomes the number of such predecessor edges (counter mode).
 */
#include <stdint.h>
#include <stdlib.h>

#define ORC_DTD_INPUT  1
#define ORC_DTD_OUTPUT 2
#define ORC_DTD_INOUT  3
#define ORC_DTD_MAXF   4

/* Returns the number of edges written, or -1 if max_edges is too small.
 * Edges are emitted in insertion order of the destination task, flow by flow: first the last writer,
 * then the readers in their insertion order. */
int orc_dtd_build(int ntasks, const int32_t* nb_flows, const int32_t* flow_tile, const int32_t* flow_op, int ntiles,
                  int32_t* e_src, int32_t* e_dst, int32_t* e_flow, int max_edges, int32_t* dep_count) {
    int32_t* last_writer = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntiles ? ntiles : 1));
    int32_t* rd_head = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntiles ? ntiles : 1));   /* readers since last writer: */
    int32_t* rd_next = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1) * ORC_DTD_MAXF);
    int32_t* rd_task = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1) * ORC_DTD_MAXF);
    int32_t* rd_tail = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntiles ? ntiles : 1));
    int nrd = 0, ne = 0, rc = 0;
    for (int t = 0; t < ntiles; ++t) { last_writer[t] = -1; rd_head[t] = -1; rd_tail[t] = -1; }
    for (int id = 0; id < ntasks && rc == 0; ++id) {
        dep_count[id] = 0;
        for (int f = 0; f < nb_flows[id] && rc == 0; ++f) {
            const int tile = flow_tile[id * ORC_DTD_MAXF + f];
            const int op = flow_op[id * ORC_DTD_MAXF + f];
            if (tile < 0) continue;                          /* NULL tile: satisfied_flow++ (:3033-3036) */
            int repeated = 0;                                /* same tile in an earlier flow of this task */
            for (int g = 0; g < f; ++g) if (flow_tile[id * ORC_DTD_MAXF + g] == tile) repeated = 1;
            if (!repeated) {
                if (last_writer[tile] >= 0 && last_writer[tile] != id) {
                    if (ne >= max_edges) { rc = -1; break; }
                    e_src[ne] = last_writer[tile]; e_dst[ne] = id; e_flow[ne] = f; ne++; dep_count[id]++;
                }
            }
            if (op != ORC_DTD_INPUT) {       /* WAR: also when this task read the tile in an earlier flow */
                for (int r = rd_head[tile]; r >= 0; r = rd_next[r]) {
                    if (rd_task[r] == id) continue;
                    if (ne >= max_edges) { rc = -1; break; }
                    e_src[ne] = rd_task[r]; e_dst[ne] = id; e_flow[ne] = f; ne++; dep_count[id]++;
                }
            }
            if (op == ORC_DTD_INPUT) {
                if (!repeated) {
                    rd_task[nrd] = id; rd_next[nrd] = -1;
                    if (rd_tail[tile] >= 0) rd_next[rd_tail[tile]] = nrd; else rd_head[tile] = nrd;
                    rd_tail[tile] = nrd; nrd++;
                }
            } else {
                last_writer[tile] = id; rd_head[tile] = -1; rd_tail[tile] = -1;
            }
        }
    }
    free(last_writer); free(rd_head); free(rd_next); free(rd_task); free(rd_tail);
    return rc < 0 ? rc : ne;
}
