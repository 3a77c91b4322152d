/* device_b200_internal.h -- private to parsec/mca/device/b200 */
#ifndef PARSEC_DEVICE_B200_INTERNAL_H
#define PARSEC_DEVICE_B200_INTERNAL_H

#include "parsec/mca/device/device_gpu.h"
#include "parsec/mca/device/cuda/device_cuda.h"
#include "parsec/mca/device/b200/device_b200.h"

BEGIN_C_DECLS

extern int parsec_device_b200_enabled, parsec_device_b200_enabled_index, parsec_b200_dry_run;
extern int parsec_b200_memory_block_size, parsec_b200_memory_percentage, parsec_b200_memory_number_of_blocks;
extern int parsec_b200_parallel_completion, parsec_b200_stage_window, parsec_b200_registration_cache;
extern char *parsec_b200_trace;
extern int parsec_b200_nvtx;
extern int parsec_b200_cmd_slots, parsec_b200_idle_us, parsec_b200_max_workers;

int  parsec_b200_device_count(void);
int  parsec_b200_module_init(int dev_id, parsec_device_module_t **module);
int  parsec_b200_module_fini(parsec_device_module_t *device);

END_C_DECLS
#endif
