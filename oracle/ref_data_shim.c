/* OUR helper (not reference code): lets the tests drive the reference's own parsec/data.c, compiled from
 * /root/reference by Makefile.ref, and read back the host-visible coherency state it leaves behind. */
#include "parsec/parsec_config.h"
#include "parsec/parsec_internal.h"
#include "parsec/data_internal.h"
#include "parsec/mca/device/device.h"
#include "parsec/arena.h"

uint32_t parsec_nb_devices = 0;
parsec_device_module_t* parsec_mca_device_get(uint32_t idx) { (void)idx; return NULL; }
int parsec_mca_device_is_gpu(uint32_t idx) { return idx >= 2; }
int parsec_mca_device_registration_completed(parsec_context_t* c) { (void)c; return 1; }
int parsec_arena_construct(parsec_arena_t* a, size_t s, size_t al) { (void)a; (void)s; (void)al; return 0; }
void parsec_arena_release(parsec_data_copy_t* c) { (void)c; }
parsec_class_t parsec_arena_t_class;

int ref_data_setup(int ndev) { parsec_nb_devices = (uint32_t)ndev; return parsec_data_init(NULL); }

/* parsec_data_create semantics (data.c:524-561): host copy OWNED, owner_device 0 */
void* ref_data_new(void) {
    parsec_data_t* d = parsec_data_new();
    d->owner_device = 0; d->preferred_device = -1; d->nb_copies = 0; d->dc = NULL; d->span = 0;
    for (uint32_t i = 0; i < parsec_nb_devices; ++i) d->device_copies[i] = NULL;
    parsec_data_copy_t* c = parsec_data_copy_new(d, 0, PARSEC_DATATYPE_NULL, PARSEC_DATA_FLAG_PARSEC_MANAGED);
    c->coherency_state = PARSEC_DATA_COHERENCY_OWNED;
    return d;
}
void ref_data_add_copy(void* data, int dev) {
    parsec_data_t* d = (parsec_data_t*)data;
    if (NULL == d->device_copies[dev])
        (void)parsec_data_copy_new(d, (uint8_t)dev, PARSEC_DATATYPE_NULL,
                                   PARSEC_DATA_FLAG_PARSEC_MANAGED | PARSEC_DATA_FLAG_PARSEC_OWNED);
}
void ref_data_set(void* data, int dev, int coh, int status, unsigned version, int readers) {
    parsec_data_copy_t* c = ((parsec_data_t*)data)->device_copies[dev];
    c->coherency_state = (parsec_data_coherency_t)coh; c->data_transfer_status = (parsec_data_status_t)status;
    c->version = version; c->readers = readers;
}
void ref_data_set_owner(void* data, int owner) { ((parsec_data_t*)data)->owner_device = (int8_t)owner; }
/* out: present, coherency, status, readers, version */
void ref_data_get(void* data, int dev, int* out) {
    parsec_data_copy_t* c = ((parsec_data_t*)data)->device_copies[dev];
    out[0] = (c != NULL);
    if (c) { out[1] = c->coherency_state; out[2] = c->data_transfer_status; out[3] = c->readers; out[4] = (int)c->version; }
}
int ref_data_owner(void* data) { return ((parsec_data_t*)data)->owner_device; }
int ref_start_transfer(void* data, int dev, int access) {
    return parsec_data_start_transfer_ownership_to_copy((parsec_data_t*)data, (uint8_t)dev, (uint8_t)access);
}
void ref_end_transfer(void* data, int dev, int access) {
    parsec_data_end_transfer_ownership_to_copy((parsec_data_t*)data, (uint8_t)dev, (uint8_t)access);
}
