/*
 * orc_lcg.c -- ORACLE (test infrastructure): the reference's deterministic tile generator.
 *
 * Restates tests/dsl/dtd/dtd_test_simple_gemm.c:97-100 (constants), :154-172 (Rnd64_jump: O(log n)
 * jump-ahead of the 64-bit LCG  x <- A*x + C) and :174-196 (initialize_tile: column-major tile, column j
 * of the tile starts at stream position m + (n+j)*M, value 0.5f - ran * 5.42e-20f).
 * Seeds used by the reference: A 1789, B 1805, C 1901 (:1135-1139).
 */
#include <stdint.h>

#define Rnd64_A 6364136223846793005ULL
#define Rnd64_C 1ULL
#define RndF_Mul 5.4210108624275222e-20f

uint64_t orc_rnd64_jump(uint64_t n, uint64_t seed) {
    uint64_t a_k = Rnd64_A, c_k = Rnd64_C, ran = seed;
    for (; n; n >>= 1) {
        if (n & 1) ran = a_k * ran + c_k;
        c_k *= (a_k + 1);
        a_k *= a_k;
    }
    return ran;
}

/* the same stream, one step at a time: used to pin the jump-ahead */
uint64_t orc_rnd64_step(uint64_t ran) { return Rnd64_A * ran + Rnd64_C; }

/* Fill one mb x nb tile whose top-left element is global (m, n) of an M-row matrix; ld = leading dimension
 * (column-major, as the reference).  Values are float (the reference stores them into double tiles). */
void orc_lcg_tile(float* data, int m, int n, int mb, int nb, int M, int ld, uint32_t seed) {
    uint64_t jump = (uint64_t)m + (uint64_t)n * (uint64_t)M;
    for (int j = 0; j < nb; ++j) {
        uint64_t ran = orc_rnd64_jump(jump, seed);
        for (int i = 0; i < mb; ++i) {
            data[(long)j * ld + i] = 0.5f - ran * RndF_Mul;
            ran = Rnd64_A * ran + Rnd64_C;
        }
        jump += (uint64_t)M;
    }
}
