// pb2_partition.cpp -- split one dependency-closed window over the GPUs of a box (host logic).
//
// The reference decides where a task runs from the owner of its affinity datum and turns every edge between two
// owners into a remote dependency at run time (remote_dep.c:451 parsec_remote_dep_activate; the receiver side
// releases the local successors in remote_dep_mpi.c:1860).  Here the whole frontier is known when the window is
// built, so the split is computed once: per-rank task tables, local successor CSR, remote successor CSR
// (rank, task, parts) consumed by release_remote_warp, and per-rank tile descriptors whose src_ptr points into
// the producer's slab.
#include "pb2_engine.h"

#include <cstring>
#include <functional>
#include <queue>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace {

thread_local std::string g_err;

struct Desc {            // one tile descriptor of a rank's window
    int32_t tile;        // global tile
    int32_t slot;        // slot of the rank's slab
    int32_t src_rank;    // -1: keep the global tile's host source / state; >= 0: pull from that rank's slot
    int32_t src_slot;
    int32_t state;
    int32_t ver;         // writers of the tile before the first user of this descriptor: its version on arrival
    int32_t pushable;    // nothing on this rank used the slot before: the producer may write the bytes itself
};
struct PushEnt { int32_t dst_rank, dst_desc, src_flow; };

struct RankPart {
    std::vector<int32_t> gid;                 // local id -> global id
    std::vector<uint32_t> succ;               // local successor entries
    std::vector<int32_t> succ_begin, succ_count;
    std::vector<int32_t> rs_begin, rs_rank;
    std::vector<uint32_t> rs_target;
    std::vector<int32_t> ready;
    std::vector<Desc> descs;
    std::vector<int32_t> slot_tile;
    std::vector<uint64_t> slot_off;
    std::map<int32_t, int32_t> slot_of;       // global tile -> slot
    uint64_t slab_bytes = 0;
};

}  // namespace

struct pb2_partition_s {
    int32_t nranks = 0, ntasks = 0;
    std::vector<pb2_task_t> tasks;            // global, dep_goal rewritten as a counter
    std::vector<pb2_tile_t> tiles;
    std::vector<int32_t> rank, lid;
    std::vector<int32_t> flow_desc;           // [task][flow] -> descriptor index in the task's rank
    std::vector<RankPart> parts;
    std::vector<std::vector<PushEnt>> push_of;   // per global task: what it may push
    bool push_on = false;
};

static int slot_for(pb2_partition_s* P, int32_t r, int32_t tile) {
    RankPart& rp = P->parts[(size_t)r];
    auto it = rp.slot_of.find(tile);
    if (it != rp.slot_of.end()) return it->second;
    const int32_t s = (int32_t)rp.slot_tile.size();
    rp.slot_of[tile] = s;
    rp.slot_tile.push_back(tile);
    rp.slot_off.push_back(rp.slab_bytes);
    rp.slab_bytes += ((uint64_t)P->tiles[(size_t)tile].bytes + 255u) & ~(uint64_t)255u;
    return s;
}

extern "C" {

const char* pb2_partition_error(void) { return g_err.c_str(); }

int pb2_partition_create(pb2_partition_t** out, const pb2_task_t* tasks, int32_t ntasks, const uint32_t* succ, int32_t nsucc,
                         const pb2_tile_t* tiles, int32_t ntiles, const int32_t* ready, int32_t nready,
                         const int32_t* task_rank, const int32_t* tile_rank, int32_t nranks, int32_t part_bytes) {
    if (!out || ntasks < 0 || nsucc < 0 || ntiles < 0 || nready < 0 || nranks <= 0 || (ntasks && (!tasks || !task_rank)) ||
        (ntiles && (!tiles || !tile_rank))) { g_err = "bad argument"; return PB2_ERR_BAD_PARAM; }
    (void)part_bytes;
    pb2_partition_s* P = new pb2_partition_s();
    P->nranks = nranks; P->ntasks = ntasks;
    P->tasks.assign(tasks, tasks + ntasks);
    P->tiles.assign(tiles, tiles + ntiles);
    P->rank.assign(task_rank, task_rank + ntasks);
    P->lid.assign((size_t)ntasks, -1);
    P->parts.resize((size_t)nranks);
    P->flow_desc.assign((size_t)ntasks * PB2_MAX_FLOWS, -1);
    P->push_of.resize((size_t)ntasks);
#define FAIL(code, msg) do { g_err = (msg); delete P; return (code); } while (0)
    for (int32_t t = 0; t < ntasks; ++t) {
        if (task_rank[t] < 0 || task_rank[t] >= nranks) FAIL(PB2_ERR_BAD_PARAM, "task_rank out of range");
        const pb2_task_t& k = tasks[t];
        if (k.nb_flows > PB2_MAX_FLOWS) FAIL(PB2_ERR_BAD_PARAM, "task with more than PB2_MAX_FLOWS flows");
        if (k.succ_count < 0 || k.succ_begin < 0 || (int64_t)k.succ_begin + k.succ_count > nsucc) FAIL(PB2_ERR_VALUE_OUT_OF_BOUNDS, "successor range out of bounds");
        for (int f = 0; f < k.nb_flows; ++f) if (k.tile[f] >= ntiles) FAIL(PB2_ERR_VALUE_OUT_OF_BOUNDS, "tile id out of bounds");
        RankPart& rp = P->parts[(size_t)task_rank[t]];
        P->lid[(size_t)t] = (int32_t)rp.gid.size();
        rp.gid.push_back(t);
    }
    for (int32_t i = 0; i < ntiles; ++i) if (tile_rank[i] < 0 || tile_rank[i] >= nranks) FAIL(PB2_ERR_BAD_PARAM, "tile_rank out of range");
    for (int32_t i = 0; i < nsucc; ++i) if (PB2_SUCC_TASK(succ[i]) >= ntasks) FAIL(PB2_ERR_VALUE_OUT_OF_BOUNDS, "successor id out of bounds");

    // in-degree, data producer of every (task, flow), number of parts (same rule as pb2_window_create)
    std::vector<int32_t> indeg((size_t)ntasks, 0), prod((size_t)ntasks * PB2_MAX_FLOWS, -1);
    auto writes = [&](int32_t p, int32_t tile) {
        const pb2_task_t& k = tasks[p];
        for (int g = 0; g < k.nb_flows; ++g) if (k.tile[g] == tile && (k.access[g] & PB2_FLOW_ACCESS_WRITE)) return true;
        return false;
    };
    for (int32_t p = 0; p < ntasks; ++p) {
        const pb2_task_t& k = tasks[p];
        for (int32_t e = k.succ_begin; e < k.succ_begin + k.succ_count; ++e) {
            const int32_t s = (int32_t)PB2_SUCC_TASK(succ[e]);
            const int f = (int)PB2_SUCC_FLOW(succ[e]);
            indeg[(size_t)s]++;
            if (f < tasks[s].nb_flows && tasks[s].tile[f] >= 0 && writes(p, tasks[s].tile[f])) prod[(size_t)s * PB2_MAX_FLOWS + f] = p;
        }
    }
    for (int32_t t = 0; t < ntasks; ++t) {
        pb2_task_t& k = P->tasks[(size_t)t];
        if (k.flags & PB2_TASK_DEPS_MASK) {
            if (__builtin_popcount((unsigned)k.dep_goal) != indeg[(size_t)t]) FAIL(PB2_ERR_BAD_PARAM, "mask-mode task whose in-edges do not match its dependency mask");
            k.flags = (uint8_t)(k.flags & ~PB2_TASK_DEPS_MASK);
        } else if (k.dep_goal != indeg[(size_t)t]) FAIL(PB2_ERR_BAD_PARAM, "counter-mode task whose in-edges do not match dep_goal");
        k.dep_goal = indeg[(size_t)t];
    }

    // topological order (Kahn, smallest id first: insertion order for DTD-built windows, so the versions of a tile
    // are met oldest first): the order in which a rank's tasks touch a tile
    std::vector<int32_t> order; order.reserve((size_t)ntasks);
    {
        std::vector<int32_t> left(indeg);
        std::priority_queue<int32_t, std::vector<int32_t>, std::greater<int32_t>> q;
        for (int32_t i = 0; i < nready; ++i) {
            if (ready[i] < 0 || ready[i] >= ntasks) FAIL(PB2_ERR_VALUE_OUT_OF_BOUNDS, "ready id out of bounds");
            if (indeg[(size_t)ready[i]] != 0) FAIL(PB2_ERR_BAD_PARAM, "ready task with unsatisfied in-edges");
            q.push(ready[i]);
        }
        while (!q.empty()) {
            const int32_t p = q.top(); q.pop();
            order.push_back(p);
            const pb2_task_t& k = tasks[p];
            for (int32_t e = k.succ_begin; e < k.succ_begin + k.succ_count; ++e) {
                const int32_t s = (int32_t)PB2_SUCC_TASK(succ[e]);
                if (--left[(size_t)s] == 0) q.push(s);
            }
        }
        if ((int32_t)order.size() != ntasks) FAIL(PB2_ERR_BAD_PARAM, "window is not dependency-closed (cycle or unreachable task)");
    }

    // tile descriptors: walk every rank's tasks in topological order
    struct Use { int32_t task; bool write; int32_t pull_prod; };   // write: overwrites the slot; pull_prod: whose version it pulls
    std::vector<std::map<int32_t, int32_t>> cur((size_t)nranks);                       // rank: tile -> current descriptor
    std::vector<std::map<std::pair<int32_t, int32_t>, int32_t>> pulled((size_t)nranks); // rank: (tile, producer) -> descriptor
    std::map<std::pair<int32_t, int32_t>, std::vector<Use>> uses;                     // (rank, tile) -> users in order
    for (int32_t t : order) {
        const pb2_task_t& k = tasks[t];
        for (int f = 0; f < k.nb_flows; ++f) if (k.tile[f] >= 0)
        {   // a use overwrites the rank's slot when it writes the tile or when it first pulls another rank's version
            const int32_t p = prod[(size_t)t * PB2_MAX_FLOWS + f];
            const bool pull = p >= 0 && task_rank[p] != task_rank[t] && (k.access[f] & PB2_FLOW_ACCESS_READ);
            uses[{task_rank[t], k.tile[f]}].push_back({t, (k.access[f] & PB2_FLOW_ACCESS_WRITE) != 0 || pull, pull ? p : -1});
        }
    }
    std::vector<std::vector<std::pair<int32_t, int32_t>>> war((size_t)ntasks);      // extra control edges reader -> next writer
    struct SlotUsers { std::vector<int32_t> cur, prev; };
    std::map<std::pair<int32_t, int32_t>, SlotUsers> slot_users;                      // (rank, tile) -> users of the slot
    std::vector<int32_t> pos((size_t)ntasks, 0), stamp((size_t)ntasks, -1);
    for (size_t i = 0; i < order.size(); ++i) pos[(size_t)order[i]] = (int32_t)i;
    int32_t query = 0;
    // is there a path from -> to in the window?  1 yes, 0 no, -1 gave up after `budget` visits
    auto reaches = [&](int32_t from, int32_t to, int budget) {
        std::vector<int32_t> stack{from};
        const int32_t q = query++;
        stamp[(size_t)from] = q;
        while (!stack.empty()) {
            const int32_t x = stack.back(); stack.pop_back();
            if (x == to) return 1;
            if (--budget < 0) return -1;
            const pb2_task_t& kx = tasks[x];
            for (int32_t e = kx.succ_begin; e < kx.succ_begin + kx.succ_count; ++e) {
                const int32_t y = (int32_t)PB2_SUCC_TASK(succ[e]);
                if (pos[(size_t)y] <= pos[(size_t)to] && stamp[(size_t)y] != q) { stamp[(size_t)y] = q; stack.push_back(y); }
            }
        }
        return 0;
    };
    // write-after-read edge a -> b, unless the data edges already order the two tasks
    auto add_war = [&](int32_t a, int32_t b) {
        if (a == b) return;
        for (auto& w : war[(size_t)a]) if (w.first == b) return;
        if (pos[(size_t)a] < pos[(size_t)b] && reaches(a, b, 4096) == 1) return;
        war[(size_t)a].push_back({b, 0});
    };
    // task t reads a version out of rank `pr`'s slot: every use of that slot after position `from` that overwrites it
    // (a writer, or the users of the next pulled version: any of them may be the one that stages it in) waits for t
    auto hold_back = [&](int32_t t, const std::vector<Use>& u, size_t from) -> int {
        for (size_t i = from; i < u.size(); ++i) if (u[i].write) {
            const int32_t key = u[i].pull_prod;
            for (size_t j = i; j < u.size(); ++j) {
                if (j > i && (key < 0 || u[j].pull_prod != key)) continue;
                const int32_t w = u[j].task;
                if (pos[(size_t)w] < pos[(size_t)t] && reaches(w, t, 1 << 30) == 1) return -1;
                add_war(t, w);
            }
            break;
        }
        return 0;
    };
    // version of the tile a flow sees (number of writers before it on the data-edge chain); a rank has ONE slot per
    // tile, so the versions it touches must come in increasing order
    std::vector<int32_t> ver_in((size_t)ntasks * PB2_MAX_FLOWS, 0);
    std::vector<std::map<int32_t, int32_t>> slot_ver((size_t)nranks);
    for (int32_t t : order) {
        const pb2_task_t& k = tasks[t];
        const int32_t r = task_rank[t];
        RankPart& rp = P->parts[(size_t)r];
        for (int f = 0; f < k.nb_flows; ++f) {
            const int32_t tile = k.tile[f];
            if (tile < 0) continue;
            const int32_t p = prod[(size_t)t * PB2_MAX_FLOWS + f];
            int32_t v = 0;
            if (p >= 0) for (int g = 0; g < tasks[p].nb_flows; ++g)
                if (tasks[p].tile[g] == tile && (tasks[p].access[g] & PB2_FLOW_ACCESS_WRITE)) v = ver_in[(size_t)p * PB2_MAX_FLOWS + g] + 1;
            ver_in[(size_t)t * PB2_MAX_FLOWS + f] = v;
            {
                auto sv = slot_ver[(size_t)r].find(tile);
                if (sv != slot_ver[(size_t)r].end() && v < sv->second)
                    FAIL(PB2_ERR_NOT_SUPPORTED, "a rank needs an older version of a tile after a newer one (one slot per tile and rank)");
                slot_ver[(size_t)r][tile] = v + ((k.access[f] & PB2_FLOW_ACCESS_WRITE) ? 1 : 0);
            }
            int32_t d;
            if (p >= 0 && task_rank[p] != r) {
                auto key = std::make_pair(tile, p);
                auto it = pulled[(size_t)r].find(key);
                if (it != pulled[(size_t)r].end()) d = it->second;
                else {
                    const int32_t pr = task_rank[p];
                    d = (int32_t)rp.descs.size();
                    SlotUsers& su = slot_users[{r, tile}];
                    // first content of this rank's slot: no earlier user can be overwritten by a push
                    const int32_t pushable = (su.cur.empty() && su.prev.empty() && cur[(size_t)r].find(tile) == cur[(size_t)r].end()) ? 1 : 0;
                    rp.descs.push_back({tile, slot_for(P, r, tile), pr, slot_for(P, pr, tile), PB2_TILE_INVALID, v, pushable});
                    pulled[(size_t)r][key] = d;
                    if (pushable) {
                        int g = 0;
                        while (g < tasks[p].nb_flows && !(tasks[p].tile[g] == tile && (tasks[p].access[g] & PB2_FLOW_ACCESS_WRITE))) ++g;
                        P->push_of[(size_t)p].push_back({r, d, g});
                    }
                    su.prev.swap(su.cur); su.cur.clear();
                }
                cur[(size_t)r][tile] = d;
                // the local slot is overwritten by the pull: earlier users of the slot go first
                SlotUsers& su = slot_users[{r, tile}];
                for (int32_t u : su.prev) add_war(u, t);
                su.cur.push_back(t);
                // the producer's rank must not overwrite its copy before this reader has pulled it
                const std::vector<Use>& u = uses[{task_rank[p], tile}];
                size_t i = 0;
                while (i < u.size() && u[i].task != p) ++i;
                if (hold_back(t, u, i + 1) < 0)
                    FAIL(PB2_ERR_NOT_SUPPORTED, "a rank overwrites a tile version that a remote task still has to read");
            } else {
                slot_users[{r, tile}].cur.push_back(t);
                auto it = cur[(size_t)r].find(tile);
                if (it != cur[(size_t)r].end()) d = it->second;
                else {
                    const int32_t home = tile_rank[tile];
                    d = (int32_t)rp.descs.size();
                    if (home == r) rp.descs.push_back({tile, slot_for(P, r, tile), -1, -1, tiles[tile].state, v, 0});
                    else if (!(k.access[f] & PB2_FLOW_ACCESS_READ)) rp.descs.push_back({tile, slot_for(P, r, tile), -1, -1, PB2_TILE_VALID, v, 0});
                    else if (tiles[tile].state == PB2_TILE_VALID)
                        rp.descs.push_back({tile, slot_for(P, r, tile), home, slot_for(P, home, tile), PB2_TILE_INVALID, v, 0});
                    else FAIL(PB2_ERR_NOT_SUPPORTED, "a rank reads the initial copy of a tile that is not resident on its home rank");
                    cur[(size_t)r][tile] = d;
                }
                if (p < 0 && rp.descs[(size_t)d].src_rank >= 0 && (k.access[f] & PB2_FLOW_ACCESS_READ)) {
                    // reads the initial version out of the home rank's slot: home keeps it until this task has run
                    const int32_t home = rp.descs[(size_t)d].src_rank;
                    if (hold_back(t, uses[{home, tile}], 0) < 0)
                        FAIL(PB2_ERR_NOT_SUPPORTED, "a rank overwrites a tile version that a remote task still has to read");
                }
            }
            if ((k.access[f] & PB2_FLOW_PUSHOUT) && rp.descs[(size_t)d].src_rank >= 0)
                FAIL(PB2_ERR_NOT_SUPPORTED, "pushout of a tile version that was pulled from another rank");
            P->flow_desc[(size_t)t * PB2_MAX_FLOWS + f] = d;
        }
    }
    for (int32_t t = 0; t < ntasks; ++t) for (auto& w : war[(size_t)t]) P->tasks[(size_t)w.first].dep_goal++;
    for (int32_t i = 0; i < nready; ++i)
        if (P->tasks[(size_t)ready[i]].dep_goal == 0) P->parts[(size_t)task_rank[ready[i]]].ready.push_back(P->lid[(size_t)ready[i]]);

    // successor tables
    for (int32_t r = 0; r < nranks; ++r) {
        RankPart& rp = P->parts[(size_t)r];
        const size_t n = rp.gid.size();
        rp.succ_begin.resize(n); rp.succ_count.resize(n); rp.rs_begin.resize(n + 1);
        for (size_t l = 0; l < n; ++l) {
            const int32_t t = rp.gid[l];
            const pb2_task_t& k = tasks[t];
            rp.succ_begin[l] = (int32_t)rp.succ.size();
            rp.rs_begin[l] = (int32_t)rp.rs_rank.size();
            auto edge = [&](int32_t s, int f) {
                if (task_rank[s] == r) rp.succ.push_back(PB2_SUCC_MAKE(P->lid[(size_t)s], f));
                else { rp.rs_rank.push_back(task_rank[s]); rp.rs_target.push_back((uint32_t)P->lid[(size_t)s]); }
            };
            for (int32_t e = k.succ_begin; e < k.succ_begin + k.succ_count; ++e) edge((int32_t)PB2_SUCC_TASK(succ[e]), (int)PB2_SUCC_FLOW(succ[e]));
            for (auto& w : war[(size_t)t]) edge(w.first, PB2_MAX_FLOWS);       // control edge: no flow of the successor
            rp.succ_count[l] = (int32_t)rp.succ.size() - rp.succ_begin[l];
        }
        rp.rs_begin[n] = (int32_t)rp.rs_rank.size();
    }
#undef FAIL
    *out = P;
    return PB2_SUCCESS;
}

int pb2_partition_sizes(const pb2_partition_t* P, int32_t rank, pb2_partition_sizes_t* s) {
    if (!P || !s || rank < 0 || rank >= P->nranks) return PB2_ERR_BAD_PARAM;
    const RankPart& rp = P->parts[(size_t)rank];
    s->ntasks = (int32_t)rp.gid.size(); s->nsucc = (int32_t)rp.succ.size(); s->ntiles = (int32_t)rp.descs.size();
    s->nready = (int32_t)rp.ready.size(); s->nremote = (int32_t)rp.rs_rank.size(); s->nslots = (int32_t)rp.slot_tile.size();
    s->slab_bytes = rp.slab_bytes;
    return PB2_SUCCESS;
}

int pb2_partition_get(const pb2_partition_t* P, int32_t rank, const uint64_t* slab_base, pb2_task_t* tasks, uint32_t* succ,
                      pb2_tile_t* tiles, int32_t* ready, int32_t* rs_begin, int32_t* rs_rank, uint32_t* rs_target,
                      int32_t* global_id, int32_t* slot_tile, uint64_t* slot_offset) {
    if (!P || rank < 0 || rank >= P->nranks || !slab_base) return PB2_ERR_BAD_PARAM;
    const RankPart& rp = P->parts[(size_t)rank];
    for (size_t l = 0; l < rp.gid.size(); ++l) {
        pb2_task_t k = P->tasks[(size_t)rp.gid[l]];
        k.succ_begin = rp.succ_begin[l]; k.succ_count = rp.succ_count[l];
        for (int f = 0; f < k.nb_flows; ++f) if (k.tile[f] >= 0) k.tile[f] = P->flow_desc[(size_t)rp.gid[l] * PB2_MAX_FLOWS + f];
        if (tasks) tasks[l] = k;
        if (global_id) global_id[l] = rp.gid[l];
    }
    if (succ && !rp.succ.empty()) memcpy(succ, rp.succ.data(), rp.succ.size() * sizeof(uint32_t));
    if (ready && !rp.ready.empty()) memcpy(ready, rp.ready.data(), rp.ready.size() * sizeof(int32_t));
    if (rs_begin) memcpy(rs_begin, rp.rs_begin.data(), rp.rs_begin.size() * sizeof(int32_t));
    if (rs_rank && !rp.rs_rank.empty()) memcpy(rs_rank, rp.rs_rank.data(), rp.rs_rank.size() * sizeof(int32_t));
    if (rs_target && !rp.rs_target.empty()) memcpy(rs_target, rp.rs_target.data(), rp.rs_target.size() * sizeof(uint32_t));
    if (tiles) for (size_t i = 0; i < rp.descs.size(); ++i) {
        const Desc& d = rp.descs[i];
        pb2_tile_t t = P->tiles[(size_t)d.tile];
        t.dev_ptr = reinterpret_cast<void*>(slab_base[rank] + rp.slot_off[(size_t)d.slot]);
        t.state = d.state;
        t.version += (uint32_t)d.ver;                   // every task sees the version the unsplit window would show it
        if (d.src_rank >= 0) {
            t.src_ptr = reinterpret_cast<void*>(slab_base[d.src_rank] + P->parts[(size_t)d.src_rank].slot_off[(size_t)d.src_slot]);
            t.src_kind = (P->push_on && d.pushable) ? PB2_SRC_PUSH : PB2_SRC_PEER;
        }
        tiles[i] = t;
    }
    if (slot_tile && !rp.slot_tile.empty()) memcpy(slot_tile, rp.slot_tile.data(), rp.slot_tile.size() * sizeof(int32_t));
    if (slot_offset && !rp.slot_off.empty()) memcpy(slot_offset, rp.slot_off.data(), rp.slot_off.size() * sizeof(uint64_t));
    return PB2_SUCCESS;
}

int pb2_partition_set_push(pb2_partition_t* P, int on) {
    if (!P) return PB2_ERR_BAD_PARAM;
    P->push_on = on != 0;
    return PB2_SUCCESS;
}

int pb2_partition_push_count(const pb2_partition_t* P, int32_t rank, int32_t* npush) {
    if (!P || !npush || rank < 0 || rank >= P->nranks) return PB2_ERR_BAD_PARAM;
    int32_t n = 0;
    if (P->push_on) for (int32_t t : P->parts[(size_t)rank].gid) n += (int32_t)P->push_of[(size_t)t].size();
    *npush = n;
    return PB2_SUCCESS;
}

int pb2_partition_get_push(const pb2_partition_t* P, int32_t rank, const uint64_t* slab_base, int32_t* ps_begin, pb2_push_t* push) {
    if (!P || rank < 0 || rank >= P->nranks || !slab_base || !ps_begin) return PB2_ERR_BAD_PARAM;
    const RankPart& rp = P->parts[(size_t)rank];
    int32_t n = 0;
    for (size_t l = 0; l < rp.gid.size(); ++l) {
        ps_begin[l] = n;
        if (!P->push_on) continue;
        const int32_t t = rp.gid[l];
        for (const PushEnt& e : P->push_of[(size_t)t]) {
            const RankPart& dp = P->parts[(size_t)e.dst_rank];
            const Desc& dd = dp.descs[(size_t)e.dst_desc];
            if (push) {
                pb2_push_t o;
                memset(&o, 0, sizeof o);
                o.dst = slab_base[e.dst_rank] + dp.slot_off[(size_t)dd.slot];
                o.bytes = P->tiles[(size_t)dd.tile].bytes;
                o.src_tile = P->flow_desc[(size_t)t * PB2_MAX_FLOWS + e.src_flow];
                o.rank = e.dst_rank; o.desc = e.dst_desc;
                push[n] = o;
            }
            ++n;
        }
    }
    ps_begin[rp.gid.size()] = n;
    return PB2_SUCCESS;
}

void pb2_partition_destroy(pb2_partition_t* P) { delete P; }

}  // extern "C"
