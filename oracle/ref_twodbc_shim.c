/* OUR code: link stubs and thin accessors around the reference's own two_dim_rectangle_cyclic.c / grid_2Dcyclic.c /
 * matrix.c (built by Makefile.ref from /root/reference), so the tests can ask the reference itself for
 * rank_of / data_key / local position / derived sizes and compare them with the oracle's restatement. */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "parsec/parsec_config.h"
#include "parsec/data_dist/matrix/two_dim_rectangle_cyclic.h"

/* --- pieces of the runtime we do not build --- */
char parsec_hostname_array[256] = "oracle";
const char* parsec_hostname = parsec_hostname_array;
int parsec_debug_coredump_on_fatal = 0;
void parsec_output(int id, const char* fmt, ...) { (void)id; (void)fmt; }
int parsec_vpmap_get_nb_vp(void) { return 1; }
static void oracle_exit(int status) { exit(status); }
void (*parsec_weaksym_exit)(int status) = oracle_exit;
int parsec_mca_param_reg_int_name(const char* type, const char* name, const char* help, int internal, int ro, int def, int* storage)
{ (void)type; (void)name; (void)help; (void)internal; (void)ro; if (storage) *storage = def; return 0; }
int parsec_mca_param_lookup_int(int index, int* value) { (void)index; (void)value; return -1; }

/* --- accessors --- */
void* ref_twodbc_new(int myrank, int mb, int nb, int lm, int ln, int i, int j, int m, int n, int P, int Q, int kp, int kq, int ip, int jq)
{
    parsec_matrix_block_cyclic_t* d = (parsec_matrix_block_cyclic_t*)calloc(1, sizeof *d);
    parsec_matrix_block_cyclic_init(d, PARSEC_MATRIX_INTEGER, PARSEC_MATRIX_TILE, myrank, mb, nb, lm, ln, i, j, m, n, P, Q, kp, kq, ip, jq);
    return d;
}
void ref_twodbc_free(void* p) { free(p); }     /* descriptor only: no matrix was attached */
uint32_t ref_twodbc_rank_of(void* p, int m, int n)
{ parsec_data_collection_t* dc = (parsec_data_collection_t*)p; return dc->rank_of(dc, m, n); }
uint64_t ref_twodbc_key(void* p, int m, int n)
{ parsec_data_collection_t* dc = (parsec_data_collection_t*)p; return (uint64_t)dc->data_key(dc, m, n); }
uint32_t ref_twodbc_rank_of_key(void* p, uint64_t key)
{ parsec_data_collection_t* dc = (parsec_data_collection_t*)p; return dc->rank_of_key(dc, (parsec_data_key_t)key); }
/* out: lmt, lnt, mt, nt, nb_elem_r, nb_elem_c, nb_local_tiles, bsiz, llm, lln, rrank, crank */
void ref_twodbc_info(void* p, int64_t* out)
{
    parsec_matrix_block_cyclic_t* d = (parsec_matrix_block_cyclic_t*)p;
    out[0] = d->super.lmt; out[1] = d->super.lnt; out[2] = d->super.mt; out[3] = d->super.nt;
    out[4] = d->nb_elem_r; out[5] = d->nb_elem_c; out[6] = d->super.nb_local_tiles; out[7] = (int64_t)d->super.bsiz;
    out[8] = d->super.llm; out[9] = d->super.lln; out[10] = d->grid.rrank; out[11] = d->grid.crank;
}
