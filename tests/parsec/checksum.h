/* checksum.h -- FNV-1a over the final host data of a test application: what tests/golden compares between the reference
 * runtime's CPU run and a run through a GPU device module (bit-exact final data, not only "zero errors") */
#ifndef PB2_TESTS_CHECKSUM_H
#define PB2_TESTS_CHECKSUM_H
#include <stddef.h>
#include <stdint.h>
static inline uint64_t fnv1a64(const void *data, size_t bytes, uint64_t h)
{
    const unsigned char *p = (const unsigned char*)data;
    if( 0 == h ) h = 0xcbf29ce484222325ull;
    for( size_t i = 0; i < bytes; i++ ) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
#endif
