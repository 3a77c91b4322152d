/* Hand-written stand-in for the header CMake generates from parsec/include/parsec/parsec_config.h.in,
 * covering only what the few reference sources built by oracle/Makefile.ref need (x86-64, gcc, C11 atomics).
 * This is OUR file (not a copy of reference source); the reference .c files are compiled where they lie. */
#ifndef PARSEC_CONFIG_H_HAS_BEEN_INCLUDED
#define PARSEC_CONFIG_H_HAS_BEEN_INCLUDED
#define PARSEC_ATOMIC_USE_C11_ATOMICS
#define PARSEC_ARCH_X86_64
#define PARSEC_HAVE_BUILTIN_EXPECT
#define PARSEC_HAVE_ATTRIBUTE_VISIBILITY
#define PARSEC_HAVE_ATTRIBUTE_ALWAYS_INLINE
#define PARSEC_HAVE_ATTRIBUTE_FORMAT_PRINTF
#define PARSEC_HAVE_PTHREAD_BARRIER
#define PARSEC_HAVE_THREAD_LOCAL
#define PARSEC_HAVE_STDARG_H
#define PARSEC_HAVE_UNISTD_H
#define PARSEC_HAVE_STDDEF_H
#define PARSEC_HAVE_STDBOOL_H
#define PARSEC_HAVE_STRING_H
#define PARSEC_HAVE_LIMITS_H
#define PARSEC_HAVE_ERRNO_H
#define PARSEC_SIZEOF_VOID_P 8
#define PARSEC_HAVE_GETTIMEOFDAY
#define PARSEC_HAVE_CLOCK_GETTIME
#define PARSEC_HAVE_ASPRINTF
#define PARSEC_HAVE_VASPRINTF
#define PARSEC_VERSION_MAJOR 4
#define PARSEC_VERSION_MINOR 0
#define PARSEC_VERSION_RELEASE 0
#include "parsec/parsec_options.h"
#include "parsec/parsec_config_bottom.h"
#endif
