/* Hand-written stand-in for the CMake-generated parsec_options.h (see parsec_config.h next to it). */
#ifndef PARSEC_OPTIONS_H_HAS_BEEN_INCLUDED
#define PARSEC_OPTIONS_H_HAS_BEEN_INCLUDED
#define MAX_LOCAL_COUNT 20
#define MAX_PARAM_COUNT 20
#define MAX_DEP_IN_COUNT 10
#define MAX_DEP_OUT_COUNT 10
#define MAX_TASK_STRLEN 128
#define PARSEC_DIST_SHORT_LIMIT 1
#endif
