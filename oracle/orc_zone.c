/*
 * orc_zone.c -- ORACLE (test infrastructure): the reference's device-heap segment allocator.
 *
 * Restates the observable behaviour of parsec/utils/zone_malloc.c:
 *   :62-96   zone_malloc_init  one EMPTY segment of max_segment units
 *   :130-213 zone_malloc       nb_units = ceil(size/unit); pick the free segment with the SMALLEST
 *                              nb_units >= request (rb-tree find_or_larger), and among equal sizes the one
 *                              most recently put on the free list (lists are push_front / pop_front);
 *                              the head of the segment is allocated, the remainder becomes a new free
 *                              segment; size 0 -> NULL
 *   :215-333 zone_free         mark EMPTY, merge with an EMPTY predecessor, then with an EMPTY successor,
 *                              put the merged segment at the front of the list of its size; double free
 *                              and out-of-range addresses are reported and ignored
 *   :335-352 zone_in_use
 * The rb-tree / chunk-list recycling of the reference is an implementation detail with no observable
 * effect; here a "freed at" stamp orders equal-size segments.  Addresses are returned as unit indices.
 */
#include <stdint.h>
#include <stdlib.h>

enum { ORC_SEG_UNDEFINED = 0, ORC_SEG_EMPTY = 1, ORC_SEG_FULL = 2 };

typedef struct orc_seg_s { int status, nb_units, nb_prev; uint64_t stamp; } orc_seg_t;
typedef struct orc_zone_s { int max_segment; size_t unit_size; uint64_t clock; orc_seg_t* seg; } orc_zone_t;

orc_zone_t* orc_zone_init(int max_segment, size_t unit_size) {
    if (max_segment <= 0 || unit_size == 0) return NULL;
    orc_zone_t* z = (orc_zone_t*)malloc(sizeof *z);
    z->max_segment = max_segment; z->unit_size = unit_size; z->clock = 1;
    z->seg = (orc_seg_t*)calloc((size_t)max_segment, sizeof(orc_seg_t));
    z->seg[0].status = ORC_SEG_EMPTY; z->seg[0].nb_units = max_segment; z->seg[0].nb_prev = 1; z->seg[0].stamp = z->clock++;
    return z;
}

void orc_zone_fini(orc_zone_t* z) { if (z) { free(z->seg); free(z); } }

static orc_seg_t* seg_at(orc_zone_t* z, int tid) { return (tid < 0 || tid >= z->max_segment) ? NULL : &z->seg[tid]; }

/* returns the first unit index of the allocation, or -1 (NULL in the reference) */
int orc_zone_malloc(orc_zone_t* z, size_t size) {
    const int nb_units = (int)((size + z->unit_size - 1) / z->unit_size);
    if (nb_units == 0) return -1;
    int best = -1;
    for (int tid = 0; tid < z->max_segment; tid += z->seg[tid].nb_units) {
        orc_seg_t* s = &z->seg[tid];
        if (s->status != ORC_SEG_EMPTY || s->nb_units < nb_units) continue;
        if (best < 0 || s->nb_units < z->seg[best].nb_units ||
            (s->nb_units == z->seg[best].nb_units && s->stamp > z->seg[best].stamp)) best = tid;
    }
    if (best < 0) return -1;
    orc_seg_t* cur = &z->seg[best];
    cur->status = ORC_SEG_FULL;
    if (cur->nb_units > nb_units) {
        orc_seg_t* next = seg_at(z, best + cur->nb_units);
        if (next) next->nb_prev -= nb_units;
        orc_seg_t* nw = seg_at(z, best + nb_units);
        nw->status = ORC_SEG_EMPTY; nw->nb_prev = nb_units; nw->nb_units = cur->nb_units - nb_units;
        nw->stamp = z->clock++;
        cur->nb_units = nb_units;
    }
    return best;
}

/* 0 ok, -1 address not allocated, -2 double free */
int orc_zone_free(orc_zone_t* z, int tid) {
    orc_seg_t* cur = seg_at(z, tid);
    if (!cur || cur->status == ORC_SEG_UNDEFINED) return -1;
    if (cur->status == ORC_SEG_EMPTY) return -2;
    cur->status = ORC_SEG_EMPTY;
    int prev_tid = tid - cur->nb_prev;
    orc_seg_t* prev = seg_at(z, prev_tid);
    int next_tid = tid + cur->nb_units;
    orc_seg_t* next = seg_at(z, next_tid);
    if (prev && prev->status == ORC_SEG_EMPTY) {
        if (next) next->nb_prev += prev->nb_units;
        prev->nb_units += cur->nb_units;
        cur->status = ORC_SEG_UNDEFINED;
        cur = prev; tid = prev_tid;
    }
    if (next && next->status == ORC_SEG_EMPTY) {
        next_tid += next->nb_units;
        cur->nb_units += next->nb_units;
        next->status = ORC_SEG_UNDEFINED;
        next = seg_at(z, next_tid);
        if (next) next->nb_prev = cur->nb_units;
    }
    cur->stamp = z->clock++;
    return 0;
}

size_t orc_zone_in_use(orc_zone_t* z) {
    size_t ret = 0;
    for (int tid = 0; tid < z->max_segment; tid += z->seg[tid].nb_units)
        if (z->seg[tid].status == ORC_SEG_FULL) ret += z->unit_size * (size_t)z->seg[tid].nb_units;
    return ret;
}

/* number of free segments and the largest one (for the fragmentation checks of the tests) */
void orc_zone_free_profile(orc_zone_t* z, int* nfree, int* largest) {
    *nfree = 0; *largest = 0;
    for (int tid = 0; tid < z->max_segment; tid += z->seg[tid].nb_units)
        if (z->seg[tid].status == ORC_SEG_EMPTY) { (*nfree)++; if (z->seg[tid].nb_units > *largest) *largest = z->seg[tid].nb_units; }
}
