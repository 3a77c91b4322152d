/*
 * nvtx_counter.c -- a minimal NVTX3 injection library (what Nsight Systems is to an application that uses NVTX): loaded
 * by the header-only NVTX implementation when NVTX_INJECTION64_PATH names it, it counts the domain ranges and marks an
 * application emits, by message, and writes the counts as one JSON object to $PB2_NVTX_COUNT_FILE when the process ends.
 * Test infrastructure for device_b200_nvtx (tests/test_mca_component.py); nsys is not in this image.
 *
 *   gcc -shared -fPIC -O2 -I/usr/local/cuda/include -o libnvtx_counter.so nvtx_counter.c -lpthread
 */
#include <nvtx3/nvToolsExt.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_NAMES 64
static struct { char name[96]; long pushes, marks; } names[MAX_NAMES];
static int nb_names = 0;
static long pops = 0, unbalanced = 0;
static char domain_name[96] = "";
static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static __thread int depth = 0;

static int slot_of(const nvtxEventAttributes_t *a)
{
    const char *msg = (NULL != a && NVTX_MESSAGE_TYPE_ASCII == a->messageType && NULL != a->message.ascii) ? a->message.ascii : "?";
    for( int i = 0; i < nb_names; i++ ) if( 0 == strcmp(names[i].name, msg) ) return i;
    if( nb_names == MAX_NAMES ) return MAX_NAMES - 1;
    snprintf(names[nb_names].name, sizeof names[nb_names].name, "%s", msg);
    return nb_names++;
}

static nvtxDomainHandle_t NVTX_API cb_domain_create(const char *name)
{
    pthread_mutex_lock(&mu);
    snprintf(domain_name, sizeof domain_name, "%s", name ? name : "");
    pthread_mutex_unlock(&mu);
    return (nvtxDomainHandle_t)(void*)domain_name;
}
static int NVTX_API cb_push(nvtxDomainHandle_t d, const nvtxEventAttributes_t *a)
{
    pthread_mutex_lock(&mu); names[slot_of(a)].pushes++; pthread_mutex_unlock(&mu);
    return depth++;
}
static int NVTX_API cb_pop(nvtxDomainHandle_t d)
{
    pthread_mutex_lock(&mu); pops++; if( depth <= 0 ) unbalanced++; pthread_mutex_unlock(&mu);
    return --depth;
}
static void NVTX_API cb_mark(nvtxDomainHandle_t d, const nvtxEventAttributes_t *a)
{
    pthread_mutex_lock(&mu); names[slot_of(a)].marks++; pthread_mutex_unlock(&mu);
}

static void write_counts(void)
{
    const char *path = getenv("PB2_NVTX_COUNT_FILE");
    if( NULL == path ) return;
    FILE *f = fopen(path, "w");
    if( NULL == f ) return;
    fprintf(f, "{\"domain\": \"%s\", \"pops\": %ld, \"unbalanced_pops\": %ld, \"events\": {", domain_name, pops, unbalanced);
    for( int i = 0; i < nb_names; i++ )
        fprintf(f, "%s\"%s\": {\"pushes\": %ld, \"marks\": %ld}", i ? ", " : "", names[i].name, names[i].pushes, names[i].marks);
    fprintf(f, "}}\n");
    fclose(f);
}

int InitializeInjectionNvtx2(NvtxGetExportTableFunc_t get_export_table)
{
    const NvtxExportTableCallbacks *cb = (const NvtxExportTableCallbacks*)get_export_table(NVTX_ETID_CALLBACKS);
    NvtxFunctionTable table = NULL;
    unsigned int size = 0;
    if( NULL == cb || !cb->GetModuleFunctionTable(NVTX_CB_MODULE_CORE2, &table, &size) || size <= NVTX_CBID_CORE2_DomainCreateA ) return 0;
    *table[NVTX_CBID_CORE2_DomainCreateA]     = (NvtxFunctionPointer)cb_domain_create;
    *table[NVTX_CBID_CORE2_DomainRangePushEx] = (NvtxFunctionPointer)cb_push;
    *table[NVTX_CBID_CORE2_DomainRangePop]    = (NvtxFunctionPointer)cb_pop;
    *table[NVTX_CBID_CORE2_DomainMarkEx]      = (NvtxFunctionPointer)cb_mark;
    atexit(write_counts);
    return 1;
}
