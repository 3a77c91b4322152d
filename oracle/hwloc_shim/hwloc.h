/*
 * hwloc.h -- flat-topology stand-in for the hwloc 2.x API, TEST INFRASTRUCTURE ONLY.
 *
 * The reference runtime (ICLDisco/parsec) cannot link without hwloc at the pinned commit and
 * this image has none.  This shim lets the reference's OWN scheduler / dependency engine / PTG and DTD front ends be
 * built, unmodified, into oracle/_ref/ as the CPU oracle and CPU baseline.  It describes the machine as
 *     Machine(depth 0) -> Package(1) -> Core(2) -> PU(3), one PU per core, one core per online CPU,
 * and implements exactly the entry points the reference calls (parsec/parsec_hwloc.c, parsec/vpmap.c,
 * parsec/parsec.c:646-647,829).  Binding calls are real (sched_setaffinity), everything else is bookkeeping.
 * Nothing in the product (parsec_b200/) includes or links this file.
 */
#ifndef PB2_HWLOC_SHIM_H
#define PB2_HWLOC_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HWLOC_API_VERSION 0x00020800

typedef struct hwloc_bitmap_s* hwloc_bitmap_t;
typedef const struct hwloc_bitmap_s* hwloc_const_bitmap_t;
typedef hwloc_bitmap_t hwloc_cpuset_t;
typedef hwloc_const_bitmap_t hwloc_const_cpuset_t;
typedef struct hwloc_topology* hwloc_topology_t;

typedef enum {
    HWLOC_OBJ_MACHINE = 0,
    HWLOC_OBJ_PACKAGE = 1,
    HWLOC_OBJ_CORE = 2,
    HWLOC_OBJ_PU = 3,
    HWLOC_OBJ_L1CACHE = 4,
    HWLOC_OBJ_L2CACHE = 5,
    HWLOC_OBJ_L3CACHE = 6,
    HWLOC_OBJ_NUMANODE = 13
} hwloc_obj_type_t;
#define HWLOC_OBJ_SOCKET HWLOC_OBJ_PACKAGE
#define HWLOC_OBJ_NODE   HWLOC_OBJ_NUMANODE

#define HWLOC_TYPE_DEPTH_UNKNOWN  (-1)
#define HWLOC_TYPE_DEPTH_NUMANODE (-3)

struct hwloc_cache_attr_s { unsigned long long size; unsigned depth; unsigned linesize; };
union hwloc_obj_attr_u { struct hwloc_cache_attr_s cache; };

struct hwloc_obj {
    hwloc_obj_type_t type;
    unsigned os_index;
    unsigned logical_index;
    int depth;
    struct hwloc_obj* parent;
    unsigned arity;
    struct hwloc_obj** children;
    hwloc_cpuset_t cpuset;
    union hwloc_obj_attr_u* attr;
};
typedef struct hwloc_obj* hwloc_obj_t;

typedef enum { HWLOC_CPUBIND_PROCESS = 1, HWLOC_CPUBIND_THREAD = 2 } hwloc_cpubind_flags_t;
#define HWLOC_TOPOLOGY_EXPORT_XML_FLAG_V1 1

unsigned hwloc_get_api_version(void);

/* bitmaps */
hwloc_bitmap_t hwloc_bitmap_alloc(void);
void hwloc_bitmap_free(hwloc_bitmap_t b);
hwloc_bitmap_t hwloc_bitmap_dup(hwloc_const_bitmap_t b);
int  hwloc_bitmap_copy(hwloc_bitmap_t dst, hwloc_const_bitmap_t src);
int  hwloc_bitmap_set(hwloc_bitmap_t b, unsigned id);
int  hwloc_bitmap_set_range(hwloc_bitmap_t b, unsigned begin, int end);
int  hwloc_bitmap_from_ulong(hwloc_bitmap_t b, unsigned long mask);
int  hwloc_bitmap_isset(hwloc_const_bitmap_t b, unsigned id);
int  hwloc_bitmap_iszero(hwloc_const_bitmap_t b);
int  hwloc_bitmap_first(hwloc_const_bitmap_t b);
int  hwloc_bitmap_next(hwloc_const_bitmap_t b, int prev);
int  hwloc_bitmap_last(hwloc_const_bitmap_t b);
int  hwloc_bitmap_weight(hwloc_const_bitmap_t b);
int  hwloc_bitmap_singlify(hwloc_bitmap_t b);
int  hwloc_bitmap_or(hwloc_bitmap_t res, hwloc_const_bitmap_t a, hwloc_const_bitmap_t b);
int  hwloc_bitmap_xor(hwloc_bitmap_t res, hwloc_const_bitmap_t a, hwloc_const_bitmap_t b);
int  hwloc_bitmap_intersects(hwloc_const_bitmap_t a, hwloc_const_bitmap_t b);
int  hwloc_bitmap_isincluded(hwloc_const_bitmap_t sub, hwloc_const_bitmap_t super);
int  hwloc_bitmap_asprintf(char** strp, hwloc_const_bitmap_t b);

#define hwloc_bitmap_foreach_begin(id, bitmap)                      \
    do {                                                            \
        hwloc_const_bitmap_t pb2_hw_bm__ = (bitmap);                \
        int pb2_hw_it__ = hwloc_bitmap_first(pb2_hw_bm__);          \
        while (pb2_hw_it__ >= 0) {                                  \
            (id) = (unsigned)pb2_hw_it__;
#define hwloc_bitmap_foreach_end()                                  \
            pb2_hw_it__ = hwloc_bitmap_next(pb2_hw_bm__, pb2_hw_it__); \
        }                                                           \
    } while (0)

/* topology */
int  hwloc_topology_init(hwloc_topology_t* t);
int  hwloc_topology_load(hwloc_topology_t t);
void hwloc_topology_destroy(hwloc_topology_t t);
int  hwloc_topology_dup(hwloc_topology_t* dst, hwloc_topology_t src);
int  hwloc_topology_restrict(hwloc_topology_t t, hwloc_const_bitmap_t set, unsigned long flags);
int  hwloc_get_type_depth(hwloc_topology_t t, hwloc_obj_type_t type);
int  hwloc_get_nbobjs_by_type(hwloc_topology_t t, hwloc_obj_type_t type);
unsigned hwloc_get_nbobjs_by_depth(hwloc_topology_t t, int depth);
hwloc_obj_t hwloc_get_obj_by_type(hwloc_topology_t t, hwloc_obj_type_t type, unsigned idx);
hwloc_obj_t hwloc_get_obj_by_depth(hwloc_topology_t t, int depth, unsigned idx);
hwloc_obj_t hwloc_get_ancestor_obj_by_type(hwloc_topology_t t, hwloc_obj_type_t type, hwloc_obj_t obj);
int  hwloc_get_nbobjs_inside_cpuset_by_type(hwloc_topology_t t, hwloc_const_cpuset_t set, hwloc_obj_type_t type);
int  hwloc_obj_type_is_cache(hwloc_obj_type_t type);
int  hwloc_get_cpubind(hwloc_topology_t t, hwloc_cpuset_t set, int flags);
int  hwloc_set_cpubind(hwloc_topology_t t, hwloc_const_cpuset_t set, int flags);
int  hwloc_topology_export_xmlbuffer(hwloc_topology_t t, char** xmlbuffer, int* buflen, unsigned long flags);
void hwloc_free_xmlbuffer(hwloc_topology_t t, char* xmlbuffer);

#ifdef __cplusplus
}
#endif
#endif
