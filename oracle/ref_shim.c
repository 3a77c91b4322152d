/* Symbols the reference sources built by Makefile.ref reference but that live in parts of the runtime we do
 * not build (debug output): no-op stand-ins.  OUR code. */
#include <stdarg.h>
#include <stdio.h>
int parsec_debug_output = 0;
int parsec_debug_verbose_level = 0;
void parsec_debug_verbose(int level, int id, const char* fmt, ...) { (void)level; (void)id; (void)fmt; }
void parsec_output_verbose(int level, int id, const char* fmt, ...) { (void)level; (void)id; (void)fmt; }
void parsec_fatal(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
int parsec_debug_rank = 0;
int parsec_debug_colorize = 0;
int parsec_debug_history_on_fatal = 0;
