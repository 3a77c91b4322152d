/* OUR code: link stubs and a thin driver around the reference's own parsec/mca/device/device.c (built by
 * Makefile.ref from /root/reference): fake device modules are registered with parsec_mca_device_add, a fake task
 * (task class with in/out flows, incarnations, data copies with preferred/owner devices) is handed to the
 * reference's parsec_select_best_device, and the tests compare its choice with the oracle's restatement and with
 * the product's pb2_select_best_device. */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "parsec/parsec_config.h"
#include "parsec/parsec_internal.h"
#include "parsec/mca/device/device.h"
#include "parsec/data_internal.h"
#include "parsec/utils/output.h"
#include "parsec/class/info.h"

/* --- pieces of the runtime we do not build --- */
static char hostname_buf[16] = "oracle";
const char* parsec_hostname = hostname_buf;
int parsec_debug_coredump_on_fatal = 0;
static void oracle_exit(int status) { exit(status); }
void (*parsec_weaksym_exit)(int status) = oracle_exit;
void parsec_output(int id, const char* fmt, ...) { (void)id; (void)fmt; }
int parsec_output_open(parsec_output_stream_t* lds) { (void)lds; return 0; }
void parsec_output_close(int id) { (void)id; }
void parsec_output_set_verbosity(int id, int level) { (void)id; (void)level; }
char* parsec_task_snprintf(char* str, size_t size, const parsec_task_t* task) { (void)task; if (size) str[0] = 0; return str; }
static mca_base_component_t* no_components[1] = { NULL };
char** mca_components_get_user_selection(char* type) { (void)type; return NULL; }
mca_base_component_t** mca_components_open_bytype(char* type) { (void)type; return no_components; }
void mca_components_free_user_list(char** list) { (void)list; }
int mca_components_belongs_to_user_list(char** list, const char* name) { (void)list; (void)name; return 0; }

/* defined in parsec.c, which is not built */
parsec_info_t parsec_per_device_infos;
parsec_info_t parsec_per_stream_infos;

/* the two MCA parameters the selection reads: device_load_balance_skew (index 1), _allow_cpu (index 2) */
static int g_skew = 20, g_allow_cpu = 0;
int parsec_mca_param_reg_int_name(const char* type, const char* name, const char* help, int internal, int ro, int def, int* storage)
{ (void)type; (void)name; (void)help; (void)internal; (void)ro; if (storage) *storage = def; return 0; }
int parsec_mca_param_find(const char* type, const char* component, const char* param)
{
    (void)type; (void)component;
    if (0 == strcmp(param, "load_balance_skew")) return 1;
    if (0 == strcmp(param, "load_balance_allow_cpu")) return 2;
    return -1;
}
int parsec_mca_param_lookup_int(int index, int* value)
{
    if (1 == index) { *value = g_skew; return 0; }
    if (2 == index) { *value = g_allow_cpu; return 0; }
    return -1;
}

/* --- driver --- */
#define REF_SEL_MAXDEV 16
static parsec_device_module_t* g_dev[REF_SEL_MAXDEV];
static int g_ndev = 0;
static int g_context_stand_in;

/* (re)reads the MCA parameters like parsec_mca_device_init does (it returns "no device component", which is true) */
void ref_sel_init(int skew_percent, int allow_cpu) { g_skew = skew_percent; g_allow_cpu = allow_cpu; (void)parsec_mca_device_init(); }

/* type: PARSEC_DEV_CPU 1, PARSEC_DEV_RECURSIVE 2, PARSEC_DEV_CUDA 4; returns device_index */
int ref_sel_add_device(int type, int64_t device_load, int64_t time_estimate_default)
{
    parsec_device_module_t* d = (parsec_device_module_t*)calloc(1, sizeof *d);
    d->name = "fake"; d->type = (uint8_t)type;
    d->device_load = device_load; d->time_estimate_default = time_estimate_default;
    const int idx = parsec_mca_device_add((parsec_context_t*)&g_context_stand_in, d);
    if (idx >= 0 && idx < REF_SEL_MAXDEV) { g_dev[idx] = d; if (idx + 1 > g_ndev) g_ndev = idx + 1; }
    return idx;
}
void ref_sel_set_load(int idx, int64_t device_load, int64_t time_estimate_default)
{ g_dev[idx]->device_load = device_load; g_dev[idx]->time_estimate_default = time_estimate_default; }

static int hook_stand_in(struct parsec_execution_stream_s* es, parsec_task_t* t) { (void)es; (void)t; return 0; }

/* flows: access[i] (0x4 READ, 0x8 WRITE), present[i] (data_in != NULL), preferred[i], owner[i].
 * chore_types: OR of device types that have an incarnation; devices_index_mask: tp->devices_index_mask.
 * Returns the selected device index, or -1 when the reference returns PARSEC_ERROR; *load gets this_task->load. */
int ref_sel_select(int nb_flows, const int32_t* access, const int32_t* present, const int32_t* preferred, const int32_t* owner,
                   int chore_types, uint32_t devices_index_mask, int64_t* load)
{
    parsec_task_class_t tc; parsec_taskpool_t tp; parsec_task_t task;
    parsec_flow_t flows[MAX_PARAM_COUNT];
    parsec_data_t datas[MAX_PARAM_COUNT]; parsec_data_copy_t copies[MAX_PARAM_COUNT];
    __parsec_chore_t chores[4];
    memset(&tc, 0, sizeof tc); memset(&tp, 0, sizeof tp); memset(&task, 0, sizeof task);
    memset(flows, 0, sizeof flows); memset(datas, 0, sizeof datas); memset(copies, 0, sizeof copies); memset(chores, 0, sizeof chores);
    int nc = 0;
    if (chore_types & PARSEC_DEV_CUDA) { chores[nc].type = PARSEC_DEV_CUDA; chores[nc].hook = hook_stand_in; nc++; }
    if (chore_types & PARSEC_DEV_CPU)  { chores[nc].type = PARSEC_DEV_CPU;  chores[nc].hook = hook_stand_in; nc++; }
    chores[nc].type = PARSEC_DEV_NONE;
    tc.nb_flows = (uint8_t)nb_flows; tc.incarnations = chores;
    for (int i = 0; i < nb_flows; ++i) {
        flows[i].flow_flags = (uint8_t)access[i]; flows[i].flow_index = (uint8_t)i;
        /* a flow sits in tc->in[] when it reads and in tc->out[] when it writes (jdf2c.c: one entry per direction) */
        tc.in[i]  = (access[i] & PARSEC_FLOW_ACCESS_READ)  ? &flows[i] : NULL;
        tc.out[i] = (access[i] & PARSEC_FLOW_ACCESS_WRITE) ? &flows[i] : NULL;
        if (present[i]) {
            datas[i].preferred_device = (int8_t)preferred[i]; datas[i].owner_device = (int8_t)owner[i];
            copies[i].original = &datas[i];
            task.data[i].data_in = &copies[i];
        }
    }
    tp.devices_index_mask = devices_index_mask;
    task.taskpool = &tp; task.task_class = &tc; task.chore_mask = 0xff;
    const int rc = parsec_select_best_device(&task);
    if (load) *load = task.load;
    return (PARSEC_SUCCESS == rc && task.selected_device) ? (int)task.selected_device->device_index : -1;
}
