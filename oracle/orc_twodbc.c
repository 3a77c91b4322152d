/*
 * orc_twodbc.c -- ORACLE (test infrastructure): the 2D block-cyclic tile -> owner / local slot / key map.
 *
 * Restates parsec/data_dist/matrix/two_dim_rectangle_cyclic.c
 *   :109-230  parsec_matrix_block_cyclic_init (nb_elem_r / nb_elem_c loops, llm/lln)
 *   :258-286  twoDBC_rank_of            :351-366 twoDBC_coordinates_to_position
 *   :368-412  twoDBC_data_of (TILE storage offset, key = n*lmt + m)
 *   :232-249  key2coords                :531-567, :636-690 k-cyclic variants
 * and parsec/data_dist/matrix/grid_2Dcyclic.c:29-48 (rrank / crank) and matrix.c:99-119 (lmt/lnt/bsiz).
 */
#include <stdint.h>

typedef struct orc_twodbc_s {
    /* inputs */
    int myrank, mb, nb, lm, ln, i, j, m, n, P, Q, kp, kq, ip, jq;
    /* derived: tiled matrix */
    int lmt, lnt, mt, nt;
    int64_t bsiz;
    /* derived: grid */
    int rrank, crank;
    /* derived: local storage */
    int nb_elem_r, nb_elem_c, nb_local_tiles, llm, lln;
} orc_twodbc_t;

void orc_twodbc_init(orc_twodbc_t* d) {
    /* matrix.c:99-119 */
    d->bsiz = (int64_t)d->mb * d->nb;
    d->lmt = (d->lm % d->mb == 0) ? (d->lm / d->mb) : (d->lm / d->mb + 1);
    d->lnt = (d->ln % d->nb == 0) ? (d->ln / d->nb) : (d->ln / d->nb + 1);
    /* matrix.c:127-128 */
    d->mt = (d->i + d->m - 1) / d->mb - d->i / d->mb + 1;
    d->nt = (d->j + d->n - 1) / d->nb - d->j / d->nb + 1;
    /* grid_2Dcyclic.c:44-45 */
    d->rrank = ((d->myrank / d->Q) + (d->P - d->ip)) % d->P;
    d->crank = ((d->myrank % d->Q) + (d->Q - d->jq)) % d->Q;
    /* two_dim_rectangle_cyclic.c:142-176 */
    int temp;
    d->nb_elem_r = 0;
    temp = d->rrank * d->kp;
    while (temp < d->lmt) {
        if (temp + d->kp < d->lmt) { d->nb_elem_r += d->kp; temp += d->P * d->kp; continue; }
        d->nb_elem_r += d->lmt - temp;
        break;
    }
    d->nb_elem_c = 0;
    temp = d->crank * d->kq;
    while (temp < d->lnt) {
        if (temp + d->kq < d->lnt) { d->nb_elem_c += d->kq; temp += d->Q * d->kq; continue; }
        d->nb_elem_c += d->lnt - temp;
        break;
    }
    if (d->nb_elem_r == 0) d->nb_elem_c = 0;
    if (d->nb_elem_c == 0) d->nb_elem_r = 0;
    d->nb_local_tiles = d->nb_elem_r * d->nb_elem_c;
    d->llm = d->nb_elem_r * d->mb;
    d->lln = d->nb_elem_c * d->nb;
}

/* :258-286 (kp == kq == 1) and :531-567 (k-cyclic) */
uint32_t orc_twodbc_rank_of(const orc_twodbc_t* d, int m, int n) {
    m += d->i / d->mb;
    n += d->j / d->nb;
    const int str = m / d->kp, stc = n / d->kq;
    const int rr = (str % d->P + d->ip) % d->P;
    const int cr = (stc % d->Q + d->jq) % d->Q;
    return (uint32_t)(rr * d->Q + cr);
}

/* :351-366 and the local_m/local_n computation of :656-670; -1 if (m,n) is not local to myrank */
int orc_twodbc_position(const orc_twodbc_t* d, int m, int n) {
    m += d->i / d->mb;
    n += d->j / d->nb;
    int local_m = (m / (d->kp * d->P)) * d->kp;
    int mm = m % (d->kp * d->P);
    if (mm / d->kp != d->rrank) return -1;
    local_m += mm % d->kp;
    int local_n = (n / (d->kq * d->Q)) * d->kq;
    int nn = n % (d->kq * d->Q);
    if (nn / d->kq != d->crank) return -1;
    local_n += nn % d->kq;
    return d->nb_elem_r * local_n + local_m;
}

/* :411 key = (n * lmt) + m with the (i,j) offset applied; matrix.c:235 */
uint64_t orc_twodbc_key(const orc_twodbc_t* d, int m, int n) {
    m += d->i / d->mb;
    n += d->j / d->nb;
    return (uint64_t)n * (uint64_t)d->lmt + (uint64_t)m;
}

/* :232-249 */
void orc_twodbc_key2coords(const orc_twodbc_t* d, uint64_t key, int* m, int* n) {
    const int _m = (int)(key % (uint64_t)d->lmt), _n = (int)(key / (uint64_t)d->lmt);
    *m = _m - d->i / d->mb;
    *n = _n - d->j / d->nb;
}

/* :395-399 element offset of the tile in TILE storage: position * bsiz */
int64_t orc_twodbc_tile_offset_elems(const orc_twodbc_t* d, int m, int n) {
    const int pos = orc_twodbc_position(d, m, n);
    return pos < 0 ? -1 : (int64_t)pos * d->bsiz;
}
