/*
 * orc_select.c -- ORACLE (test infrastructure): which device a ready task is given to.
 *
 * Restates parsec_select_best_device, parsec/mca/device/device.c:100-310, for tasks that have a GPU
 * incarnation (and optionally a CPU one):
 *   1. first flow with ACCESS_WRITE: its data's preferred_device if that device can run the task, else
 *      its owner_device if that is a GPU (:170-192);
 *   2. else first READ flow: preferred_device wins outright; an owner GPU becomes the "preferred by
 *      read data" candidate rdata_dev (:194-217);
 *   3. else least ETA = device_load + time_estimate over the enabled devices, scanned from the highest
 *      index down; rdata_dev keeps the task until its ETA, scaled by 1/(1+skew%), stops being the best;
 *      the CPU (index 0) is only used when no GPU is valid unless load_balance_allow_cpu (:220-268);
 *   4. the caller then adds the estimate to device_load (scheduling.c:142).
 * time_estimate_default = total_gflops_fp64 / device gflops_fp64 (device.c:792-833).
 */
#include <stdint.h>

#define ORC_SEL_MAXF 4
#define ORC_ACC_READ  0x04
#define ORC_ACC_WRITE 0x08

typedef struct orc_sel_dev_s {
    int32_t is_gpu;          /* PARSEC_DEV_IS_GPU(type) */
    int32_t is_recursive;    /* PARSEC_DEV_RECURSIVE    */
    int32_t enabled;         /* tp->devices_index_mask bit and valid_types match */
    int64_t device_load;
    int64_t time_estimate;
} orc_sel_dev_t;

/* per flow: access bits, preferred_device and owner_device of the flow's data (-1 = none), present = data_in != NULL */
int orc_select_best_device(const orc_sel_dev_t* dev, int ndev, int nb_flows, const int32_t* access,
                           const int32_t* present, const int32_t* preferred, const int32_t* owner,
                           int skew_percent, int allow_cpu) {
    const float skew = 1.f / (skew_percent / 100.f + 1.f);
    int rdata_dev = -1;
    for (int i = 0; i < nb_flows; i++) {
        if (!(access[i] & ORC_ACC_WRITE) || !present[i]) continue;
        int d = preferred[i];
        if (d >= 0 && d < ndev && dev[d].enabled) return d;
        d = owner[i];
        if (d >= 0 && d < ndev && dev[d].enabled && dev[d].is_gpu) return d;
    }
    for (int i = 0; i < nb_flows; i++) {
        if (!(access[i] & ORC_ACC_READ) || !present[i]) continue;   /* tc->in[i] exists only for flows that read */
        int d = preferred[i];
        if (d >= 0 && d < ndev && dev[d].enabled) return d;
        d = owner[i];
        if (d >= 0 && d < ndev && dev[d].enabled && dev[d].is_gpu) { rdata_dev = d; break; }
    }
    int best_index = -1;
    int64_t best_eta = INT64_MAX;
    if (rdata_dev >= 0) {
        best_index = rdata_dev;
        best_eta = dev[rdata_dev].device_load + dev[rdata_dev].time_estimate;
        best_eta = (int64_t)(best_eta * skew);
    }
    for (int d = ndev - 1; d >= 0; d--) {
        if (!dev[d].enabled || dev[d].is_recursive) continue;
        const int64_t eta = dev[d].device_load + dev[d].time_estimate;
        if (best_eta > eta) {
            if (best_index != -1 && !dev[d].is_gpu && !allow_cpu) continue;
            best_index = d;
            best_eta = eta;
        }
    }
    return best_index;   /* -1: no valid device */
}
