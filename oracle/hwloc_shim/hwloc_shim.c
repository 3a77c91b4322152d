/* hwloc_shim.c -- see hwloc.h in this directory.  Flat topology, real thread binding.  Test infrastructure only. */
#define _GNU_SOURCE
#include "hwloc.h"
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define SHIM_WORDS 32                       /* up to 2048 CPUs */
#define SHIM_BITS  (SHIM_WORDS * 64)
struct hwloc_bitmap_s { unsigned long long w[SHIM_WORDS]; };

struct hwloc_topology {
    int ncpu;                               /* one core + one PU per online CPU */
    struct hwloc_obj machine, package;
    struct hwloc_obj *cores, *pus;
    struct hwloc_obj **pkg_children, **core_children;   /* core_children[i] -> &pus[i] */
    struct hwloc_obj *machine_children[1];
    union hwloc_obj_attr_u no_attr;
};

unsigned hwloc_get_api_version(void) { return HWLOC_API_VERSION; }

hwloc_bitmap_t hwloc_bitmap_alloc(void) { return (hwloc_bitmap_t)calloc(1, sizeof(struct hwloc_bitmap_s)); }
void hwloc_bitmap_free(hwloc_bitmap_t b) { free(b); }
hwloc_bitmap_t hwloc_bitmap_dup(hwloc_const_bitmap_t b) {
    hwloc_bitmap_t r = hwloc_bitmap_alloc();
    if (r && b) *r = *b;
    return r;
}
int hwloc_bitmap_copy(hwloc_bitmap_t dst, hwloc_const_bitmap_t src) { *dst = *src; return 0; }
int hwloc_bitmap_set(hwloc_bitmap_t b, unsigned id) {
    if (id >= SHIM_BITS) return -1;
    b->w[id >> 6] |= 1ull << (id & 63);
    return 0;
}
int hwloc_bitmap_set_range(hwloc_bitmap_t b, unsigned begin, int end) {
    if (end < 0 || end >= SHIM_BITS) end = SHIM_BITS - 1;
    for (unsigned i = begin; (int)i <= end; ++i) hwloc_bitmap_set(b, i);
    return 0;
}
int hwloc_bitmap_from_ulong(hwloc_bitmap_t b, unsigned long mask) {
    memset(b, 0, sizeof *b);
    b->w[0] = mask;
    return 0;
}
int hwloc_bitmap_isset(hwloc_const_bitmap_t b, unsigned id) { return id < SHIM_BITS && ((b->w[id >> 6] >> (id & 63)) & 1ull); }
int hwloc_bitmap_iszero(hwloc_const_bitmap_t b) {
    for (int i = 0; i < SHIM_WORDS; ++i) if (b->w[i]) return 0;
    return 1;
}
int hwloc_bitmap_next(hwloc_const_bitmap_t b, int prev) {
    for (int i = prev + 1; i < SHIM_BITS; ++i) {
        if (!b->w[i >> 6]) { i |= 63; continue; }
        if ((b->w[i >> 6] >> (i & 63)) & 1ull) return i;
    }
    return -1;
}
int hwloc_bitmap_first(hwloc_const_bitmap_t b) { return hwloc_bitmap_next(b, -1); }
int hwloc_bitmap_last(hwloc_const_bitmap_t b) {
    for (int i = SHIM_BITS - 1; i >= 0; --i) if (hwloc_bitmap_isset(b, (unsigned)i)) return i;
    return -1;
}
int hwloc_bitmap_weight(hwloc_const_bitmap_t b) {
    int n = 0;
    for (int i = 0; i < SHIM_WORDS; ++i) n += __builtin_popcountll(b->w[i]);
    return n;
}
int hwloc_bitmap_singlify(hwloc_bitmap_t b) {
    const int f = hwloc_bitmap_first(b);
    memset(b, 0, sizeof *b);
    if (f >= 0) hwloc_bitmap_set(b, (unsigned)f);
    return 0;
}
int hwloc_bitmap_or(hwloc_bitmap_t res, hwloc_const_bitmap_t a, hwloc_const_bitmap_t b) {
    for (int i = 0; i < SHIM_WORDS; ++i) res->w[i] = a->w[i] | b->w[i];
    return 0;
}
int hwloc_bitmap_xor(hwloc_bitmap_t res, hwloc_const_bitmap_t a, hwloc_const_bitmap_t b) {
    for (int i = 0; i < SHIM_WORDS; ++i) res->w[i] = a->w[i] ^ b->w[i];
    return 0;
}
int hwloc_bitmap_intersects(hwloc_const_bitmap_t a, hwloc_const_bitmap_t b) {
    for (int i = 0; i < SHIM_WORDS; ++i) if (a->w[i] & b->w[i]) return 1;
    return 0;
}
int hwloc_bitmap_isincluded(hwloc_const_bitmap_t sub, hwloc_const_bitmap_t super) {
    for (int i = 0; i < SHIM_WORDS; ++i) if (sub->w[i] & ~super->w[i]) return 0;
    return 1;
}
int hwloc_bitmap_asprintf(char** strp, hwloc_const_bitmap_t b) {
    /* hwloc prints comma-separated 32-bit hex words, most significant first, leading zero words dropped */
    char buf[SHIM_WORDS * 2 * 11 + 4];
    size_t n = 0;
    int started = 0;
    for (int i = SHIM_WORDS * 2 - 1; i >= 0; --i) {
        const unsigned v = (unsigned)((b->w[i >> 1] >> ((i & 1) * 32)) & 0xffffffffull);
        if (!started && !v && i) continue;
        n += (size_t)snprintf(buf + n, sizeof buf - n, started ? ",0x%08x" : "0x%08x", v);
        started = 1;
    }
    *strp = strdup(buf);
    return (int)n;
}

static void build(hwloc_topology_t t, int ncpu) {
    t->ncpu = ncpu;
    t->cores = (struct hwloc_obj*)calloc((size_t)ncpu, sizeof(struct hwloc_obj));
    t->pus = (struct hwloc_obj*)calloc((size_t)ncpu, sizeof(struct hwloc_obj));
    t->pkg_children = (struct hwloc_obj**)calloc((size_t)ncpu, sizeof(void*));
    t->core_children = (struct hwloc_obj**)calloc((size_t)ncpu, sizeof(void*));
    memset(&t->no_attr, 0, sizeof t->no_attr);
    t->machine.type = HWLOC_OBJ_MACHINE; t->machine.depth = 0; t->machine.parent = NULL;
    t->machine.cpuset = hwloc_bitmap_alloc(); t->machine.attr = &t->no_attr;
    t->machine.arity = 1; t->machine_children[0] = &t->package; t->machine.children = t->machine_children;
    t->package.type = HWLOC_OBJ_PACKAGE; t->package.depth = 1; t->package.parent = &t->machine;
    t->package.cpuset = hwloc_bitmap_alloc(); t->package.attr = &t->no_attr;
    t->package.arity = (unsigned)ncpu; t->package.children = t->pkg_children;
    for (int i = 0; i < ncpu; ++i) {
        struct hwloc_obj* c = &t->cores[i];
        struct hwloc_obj* p = &t->pus[i];
        c->type = HWLOC_OBJ_CORE; c->depth = 2; c->os_index = c->logical_index = (unsigned)i; c->parent = &t->package;
        c->cpuset = hwloc_bitmap_alloc(); hwloc_bitmap_set(c->cpuset, (unsigned)i); c->attr = &t->no_attr;
        c->arity = 1; t->core_children[i] = p; c->children = &t->core_children[i];
        p->type = HWLOC_OBJ_PU; p->depth = 3; p->os_index = p->logical_index = (unsigned)i; p->parent = c;
        p->cpuset = hwloc_bitmap_alloc(); hwloc_bitmap_set(p->cpuset, (unsigned)i); p->attr = &t->no_attr;
        t->pkg_children[i] = c;
        hwloc_bitmap_set(t->machine.cpuset, (unsigned)i);
        hwloc_bitmap_set(t->package.cpuset, (unsigned)i);
    }
}

int hwloc_topology_init(hwloc_topology_t* t) {
    *t = (hwloc_topology_t)calloc(1, sizeof(struct hwloc_topology));
    return *t ? 0 : -1;
}
int hwloc_topology_load(hwloc_topology_t t) {
    long n = sysconf(_SC_NPROCESSORS_CONF);
    if (n < 1) n = 1;
    if (n > SHIM_BITS) n = SHIM_BITS;
    build(t, (int)n);
    return 0;
}
void hwloc_topology_destroy(hwloc_topology_t t) {
    if (!t) return;
    for (int i = 0; i < t->ncpu; ++i) { hwloc_bitmap_free(t->cores[i].cpuset); hwloc_bitmap_free(t->pus[i].cpuset); }
    hwloc_bitmap_free(t->machine.cpuset); hwloc_bitmap_free(t->package.cpuset);
    free(t->cores); free(t->pus); free(t->pkg_children); free(t->core_children);
    free(t);
}
int hwloc_topology_dup(hwloc_topology_t* dst, hwloc_topology_t src) {
    if (hwloc_topology_init(dst)) return -1;
    build(*dst, src->ncpu);
    return 0;
}
/* keep only the cores inside `set` (logical indexes are renumbered, os_index keeps the CPU number) */
int hwloc_topology_restrict(hwloc_topology_t t, hwloc_const_bitmap_t set, unsigned long flags) {
    (void)flags;
    int k = 0;
    for (int i = 0; i < t->ncpu; ++i) {
        if (!hwloc_bitmap_isset(set, t->cores[i].os_index)) {
            hwloc_bitmap_free(t->cores[i].cpuset); hwloc_bitmap_free(t->pus[i].cpuset);
            continue;
        }
        if (k != i) { t->cores[k] = t->cores[i]; t->pus[k] = t->pus[i]; }
        t->cores[k].logical_index = t->pus[k].logical_index = (unsigned)k;
        ++k;
    }
    if (k == 0) return -1;
    t->ncpu = k;
    memset(t->machine.cpuset, 0, sizeof(struct hwloc_bitmap_s));
    for (int i = 0; i < k; ++i) {
        t->pkg_children[i] = &t->cores[i];
        t->core_children[i] = &t->pus[i];
        t->cores[i].children = &t->core_children[i];
        t->pus[i].parent = &t->cores[i];
        hwloc_bitmap_set(t->machine.cpuset, t->cores[i].os_index);
    }
    hwloc_bitmap_copy(t->package.cpuset, t->machine.cpuset);
    t->package.arity = (unsigned)k;
    return 0;
}

int hwloc_get_type_depth(hwloc_topology_t t, hwloc_obj_type_t type) {
    (void)t;
    switch (type) {
    case HWLOC_OBJ_MACHINE: return 0;
    case HWLOC_OBJ_PACKAGE: return 1;
    case HWLOC_OBJ_CORE: return 2;
    case HWLOC_OBJ_PU: return 3;
    case HWLOC_OBJ_NUMANODE: return HWLOC_TYPE_DEPTH_NUMANODE;
    default: return HWLOC_TYPE_DEPTH_UNKNOWN;
    }
}
unsigned hwloc_get_nbobjs_by_depth(hwloc_topology_t t, int depth) {
    if (depth == 0 || depth == 1) return 1;
    if (depth == 2 || depth == 3) return (unsigned)t->ncpu;
    return 0;
}
int hwloc_get_nbobjs_by_type(hwloc_topology_t t, hwloc_obj_type_t type) {
    const int d = hwloc_get_type_depth(t, type);
    return d < 0 ? 0 : (int)hwloc_get_nbobjs_by_depth(t, d);
}
hwloc_obj_t hwloc_get_obj_by_depth(hwloc_topology_t t, int depth, unsigned idx) {
    if (idx >= hwloc_get_nbobjs_by_depth(t, depth)) return NULL;
    switch (depth) {
    case 0: return &t->machine;
    case 1: return &t->package;
    case 2: return &t->cores[idx];
    case 3: return &t->pus[idx];
    default: return NULL;
    }
}
hwloc_obj_t hwloc_get_obj_by_type(hwloc_topology_t t, hwloc_obj_type_t type, unsigned idx) {
    const int d = hwloc_get_type_depth(t, type);
    return d < 0 ? NULL : hwloc_get_obj_by_depth(t, d, idx);
}
hwloc_obj_t hwloc_get_ancestor_obj_by_type(hwloc_topology_t t, hwloc_obj_type_t type, hwloc_obj_t obj) {
    (void)t;
    for (hwloc_obj_t o = obj ? obj->parent : NULL; o; o = o->parent) if (o->type == type) return o;
    return NULL;
}
int hwloc_get_nbobjs_inside_cpuset_by_type(hwloc_topology_t t, hwloc_const_cpuset_t set, hwloc_obj_type_t type) {
    const int d = hwloc_get_type_depth(t, type);
    if (d < 0) return 0;
    int n = 0;
    for (unsigned i = 0; i < hwloc_get_nbobjs_by_depth(t, d); ++i) {
        hwloc_obj_t o = hwloc_get_obj_by_depth(t, d, i);
        if (!hwloc_bitmap_iszero(o->cpuset) && hwloc_bitmap_isincluded(o->cpuset, set)) ++n;
    }
    return n;
}
int hwloc_obj_type_is_cache(hwloc_obj_type_t type) { return type >= HWLOC_OBJ_L1CACHE && type <= HWLOC_OBJ_L3CACHE; }

int hwloc_get_cpubind(hwloc_topology_t t, hwloc_cpuset_t set, int flags) {
    (void)t;
    cpu_set_t cs;
    CPU_ZERO(&cs);
    int rc;
    if (flags & HWLOC_CPUBIND_THREAD) rc = pthread_getaffinity_np(pthread_self(), sizeof cs, &cs);
    else rc = sched_getaffinity(getpid(), sizeof cs, &cs);
    if (rc) return -1;
    memset(set, 0, sizeof *set);
    for (int i = 0; i < CPU_SETSIZE && i < SHIM_BITS; ++i) if (CPU_ISSET(i, &cs)) hwloc_bitmap_set(set, (unsigned)i);
    return 0;
}
int hwloc_set_cpubind(hwloc_topology_t t, hwloc_const_cpuset_t set, int flags) {
    (void)t;
    cpu_set_t cs;
    CPU_ZERO(&cs);
    for (int i = hwloc_bitmap_first(set); i >= 0 && i < CPU_SETSIZE; i = hwloc_bitmap_next(set, i)) CPU_SET(i, &cs);
    if (flags & HWLOC_CPUBIND_THREAD) return pthread_setaffinity_np(pthread_self(), sizeof cs, &cs) ? -1 : 0;
    return sched_setaffinity(getpid(), sizeof cs, &cs) ? -1 : 0;
}
int hwloc_topology_export_xmlbuffer(hwloc_topology_t t, char** xmlbuffer, int* buflen, unsigned long flags) {
    (void)flags;
    char buf[128];
    const int n = snprintf(buf, sizeof buf, "<topology shim=\"flat\" cores=\"%d\"/>", t->ncpu);
    *xmlbuffer = strdup(buf);
    *buflen = n + 1;
    return 0;
}
void hwloc_free_xmlbuffer(hwloc_topology_t t, char* xmlbuffer) { (void)t; free(xmlbuffer); }
