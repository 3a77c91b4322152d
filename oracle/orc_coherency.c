/*
 * orc_coherency.c -- ORACLE (test infrastructure): host-visible coherency state of one parsec_data_t.
 *
 * Restates, for a single thread, the MOESI-like protocol every GPU task drives on the data it touches:
 *   parsec/data.c:334-458   parsec_data_start_transfer_ownership_to_copy
 *   parsec/data.c:313-332   parsec_data_end_transfer_ownership_to_copy
 *   parsec/data.c:524-561   parsec_data_create: host copy OWNED, owner_device 0, version 0
 *   device_gpu.c:1209-1612  reserve_space: a missing GPU replica is created INVALID, version 0
 *   device_gpu.c:1799-2165  data_stage_in: source choice, preemptive version (+1 for WRITE), UNDER_TRANSFER
 *   device_gpu.c:2358-2573  callback_complete_push: COMPLETE_TRANSFER + end_transfer_ownership
 *   device_gpu.c:2943-3173  kernel_pop: reader release, host copy UNDER_TRANSFER for pushout flows
 *   device_gpu.c:3179-3292  kernel_epilog: pushout => both SHARED, cpu.version = gpu.version
 *   transfer_gpu.c:309-362  parsec_gpu_complete_w2r_task (restated literally, see note there)
 * Device index 0 is the host, 1 the "recursive" pseudo device, GPUs start at 2 (device.c:1129-1131).
 */
#include <stdint.h>
#include <string.h>

#define ORC_MAX_DEV 16

#define COH_INVALID   0x0
#define COH_OWNED     0x1
#define COH_EXCLUSIVE 0x2
#define COH_SHARED    0x4
#define ST_NOT_TRANSFER      0x0
#define ST_UNDER_TRANSFER    0x1
#define ST_COMPLETE_TRANSFER 0x2
#define ACC_READ  0x04
#define ACC_WRITE 0x08
#define FLAG_EVICTED (1 << 5)

typedef struct orc_copy_s {
    int32_t  present;          /* device_copies[d] != NULL */
    int32_t  coherency_state;
    int32_t  data_transfer_status;
    int32_t  readers;
    uint32_t version;
    int32_t  flags;
} orc_copy_t;

typedef struct orc_data_s {
    int32_t    owner_device;
    int32_t    preferred_device;
    int32_t    nb_devices;
    int32_t    new_data;       /* 1: NEW/arena temporary never touched (no dc, no repo entry) */
    orc_copy_t copy[ORC_MAX_DEV];
} orc_data_t;

void orc_data_create(orc_data_t* d, int nb_devices, int new_data) {
    memset(d, 0, sizeof *d);
    d->nb_devices = nb_devices;
    d->owner_device = 0;
    d->preferred_device = -1;
    d->new_data = new_data;
    d->copy[0].present = 1;
    d->copy[0].coherency_state = COH_OWNED;
}

/* data.c:334-458.  Returns the device to transfer from, or -1 when no transfer is required. */
int orc_data_start_transfer_ownership(orc_data_t* data, int device, int access_mode) {
    int transfer_required = 0;
    int valid_copy = data->owner_device;
    orc_copy_t* copy = &data->copy[device];
    if (valid_copy == device) goto bookkeeping;
    switch (copy->coherency_state) {
    case COH_INVALID:
        transfer_required = 1;
        if (-1 == valid_copy) {
            for (int i = 0; i < data->nb_devices; i++) {
                if (!data->copy[i].present) continue;
                if (COH_INVALID == data->copy[i].coherency_state) continue;
                valid_copy = i;
            }
        }
        break;
    case COH_SHARED:
        for (int i = 0; i < data->nb_devices; i++) {
            if (!data->copy[i].present) continue;
            if (COH_OWNED == data->copy[i].coherency_state && data->copy[i].version > copy->version)
                transfer_required = 1;
        }
        break;
    case COH_EXCLUSIVE:
    case COH_OWNED:
        break;
    }
    if (ACC_READ & access_mode) {
        for (int i = 0; i < data->nb_devices; i++) {
            if (device == i || !data->copy[i].present) continue;
            if (COH_INVALID == data->copy[i].coherency_state) continue;
            if (COH_OWNED == copy->coherency_state && !(ACC_WRITE & access_mode)) {
                if (data->copy[i].version < copy->version) data->copy[i].coherency_state = COH_INVALID;
                data->owner_device = -1;
            }
            if (COH_EXCLUSIVE == data->copy[i].coherency_state) data->copy[i].coherency_state = COH_SHARED;
        }
    } else {
        transfer_required = 0;    /* finally we'll just overwrite w/o read */
    }
    if (ACC_WRITE & access_mode) {
        for (int i = 0; i < data->nb_devices; i++) {
            if (!data->copy[i].present) continue;
            if (COH_INVALID == data->copy[i].coherency_state) continue;
            data->copy[i].coherency_state = COH_SHARED;
        }
    }
bookkeeping:
    if (ACC_READ & access_mode) copy->readers++;
    if (ACC_WRITE & access_mode) data->owner_device = device;
    if (!transfer_required) return -1;
    copy->coherency_state = COH_INVALID;
    return valid_copy;
}

/* data.c:313-332 */
void orc_data_end_transfer_ownership(orc_data_t* data, int device, int access_mode) {
    orc_copy_t* copy = &data->copy[device];
    if (ACC_READ & access_mode) copy->coherency_state = COH_SHARED;
    if (ACC_WRITE & access_mode) copy->coherency_state = COH_OWNED;
}

/*
 * kernel_push for one flow of a task that runs on GPU `device` and whose data_in lives on `in_device`.
 * Returns the source device of the transfer that was scheduled, or -1 if none (already there / NEW / WRITE-only).
 * *bytes_required is set to 1 when required_data_in is charged (device_gpu.c:2055).
 */
int orc_gpu_stage_in(orc_data_t* data, int device, int in_device, int access_mode, int peer_mask, int* required) {
    orc_copy_t* gpu = &data->copy[device];
    *required = 0;
    if (!gpu->present) {                       /* reserve_space: fresh replica */
        gpu->present = 1; gpu->coherency_state = COH_INVALID; gpu->version = 0; gpu->readers = 0;
        gpu->data_transfer_status = ST_NOT_TRANSFER; gpu->flags = 0;
    }
    if (in_device == device) {                 /* :1820-1843 data already located in the right place */
        if (ACC_WRITE & access_mode) gpu->version++;
        if (ACC_READ & access_mode) gpu->readers++;
        return -1;
    }
    /* :1873-1884 already under transfer: reserve the destination reader only */
    if ((ACC_READ & access_mode) && gpu->data_transfer_status == ST_UNDER_TRANSFER) {
        (void)orc_data_start_transfer_ownership(data, device, access_mode);
        *required = 1;
        return -1;
    }
    /* :1888-2008 source selection: read-only flows may use a peer GPU replica of the same version */
    int candidate = in_device;
    if ((ACC_READ & access_mode) && !(ACC_WRITE & access_mode)) {
        int found = 0;
        if (in_device >= 2 && (peer_mask & (1 << in_device)) &&
            data->copy[in_device].coherency_state != COH_INVALID &&
            data->copy[in_device].data_transfer_status != ST_UNDER_TRANSFER) {
            found = 1;
        } else {
            for (int t = 2; t < data->nb_devices && !found; t++) {
                if (t == device || !(peer_mask & (1 << t))) continue;
                if (!data->copy[t].present || data->copy[t].version != data->copy[in_device].version) continue;
                if (COH_INVALID == data->copy[t].coherency_state) continue;
                candidate = t; found = 1;
            }
            if (!found) candidate = 0;         /* fall back on the CPU copy */
        }
    }
    int transfer_from = orc_data_start_transfer_ownership(data, device, access_mode);
    /* :2049-2052 NEW data nobody touched yet is not pulled in */
    if (data->new_data && data->copy[in_device].version == 0) transfer_from = -1;
    *required = 1;
    if (-1 == transfer_from) {
        gpu->data_transfer_status = ST_COMPLETE_TRANSFER;
        orc_data_end_transfer_ownership(data, device, access_mode);
        if (ACC_WRITE & access_mode) gpu->version = data->copy[candidate].version + 1;
        return -1;
    }
    /* :2148-2153 preemptive version */
    if (ACC_WRITE & access_mode) gpu->version = data->copy[candidate].version + 1;
    else                         gpu->version = data->copy[candidate].version;
    gpu->data_transfer_status = ST_UNDER_TRANSFER;
    return candidate;
}

/* callback_complete_push for the flow whose transfer finished */
void orc_gpu_stage_in_complete(orc_data_t* data, int device, int access_mode) {
    orc_copy_t* gpu = &data->copy[device];
    if (gpu->data_transfer_status == ST_UNDER_TRANSFER) {
        gpu->data_transfer_status = ST_COMPLETE_TRANSFER;
        orc_data_end_transfer_ownership(data, device, access_mode);
    }
}

/* kernel_pop + kernel_epilog for one flow once the body is done */
void orc_gpu_task_complete(orc_data_t* data, int device, int access_mode, int pushout) {
    orc_copy_t* gpu = &data->copy[device];
    orc_copy_t* cpu = &data->copy[0];
    if (ACC_READ & access_mode) gpu->readers--;
    if (!(ACC_WRITE & access_mode)) return;
    if (pushout) {
        cpu->data_transfer_status = ST_UNDER_TRANSFER;           /* pop :3127 */
        cpu->version = gpu->version;                              /* epilog :3250-3255 */
        cpu->coherency_state = COH_SHARED;
        gpu->coherency_state = COH_SHARED;
        cpu->data_transfer_status = ST_COMPLETE_TRANSFER;
    }
    data->new_data = 0;
}

/* transfer_gpu.c:309-362, literal: a dirty replica whose version is ahead of the host copy is left as is */
void orc_gpu_w2r_complete(orc_data_t* data, int device) {
    orc_copy_t* gpu = &data->copy[device];
    orc_copy_t* cpu = &data->copy[0];
    gpu->data_transfer_status = ST_COMPLETE_TRANSFER;
    if (cpu->version < gpu->version) return;
    gpu->coherency_state = COH_SHARED;
    cpu->coherency_state = COH_SHARED;
    cpu->version = gpu->version;
    cpu->flags |= FLAG_EVICTED;
}
