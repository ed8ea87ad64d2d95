/* base.c -- error reporting and checked allocation.
 *
 * Error contract follows the reference (src/fail.upc:19-63): no return codes;
 * a failure prints "ERROR (proc NNNN, file:line): message" on stdout and ends
 * the process with the given status.  The Nek build's Fortran exit hook
 * (USE_USR_EXIT, src/fail.upc:10-17) is honoured when the application
 * provides `userexithandler_` (weak reference).
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "exags_internal.h"

exags_uint exags_comm_gbl_id = 0, exags_comm_gbl_np = 1;

extern void userexithandler_(int *status) __attribute__((weak));

void exags_die(int status)
{
  if (userexithandler_) userexithandler_(&status);
  exit(status);
}

static void report(const char *prefix, const char *file, unsigned line,
                   const char *fmt, va_list ap)
{
  char msg[2048];
  int k = snprintf(msg, sizeof msg, "%s(proc %04d, %s:%u): ", prefix,
                   (int)exags_comm_gbl_id, file, line);
  if (k < 0) k = 0;
  if ((size_t)k < sizeof msg - 2)
    vsnprintf(msg + k, sizeof msg - 2 - (size_t)k, fmt, ap);
  strcat(msg, "\n");
  fputs(msg, stdout);
  fflush(stdout);
}

void exags_diagnostic(const char *prefix, const char *file, unsigned line,
                      const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  report(prefix, file, line, fmt, ap);
  va_end(ap);
}

void exags_fail(int status, const char *file, unsigned line, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  report("ERROR ", file, line, fmt, ap);
  va_end(ap);
  exags_die(status);
  abort(); /* not reached */
}

void *xmalloc_(size_t bytes, const char *file, unsigned line)
{
  void *p = malloc(bytes ? bytes : 1);
  if (!p) exags_fail(1, file, line, "allocation of %ld bytes failed\n", (long)bytes);
  return p;
}

void *xcalloc_(size_t bytes, const char *file, unsigned line)
{
  void *p = calloc(bytes ? bytes : 1, 1);
  if (!p) exags_fail(1, file, line, "allocation of %ld bytes failed\n", (long)bytes);
  return p;
}

void *xrealloc_(void *old, size_t bytes, const char *file, unsigned line)
{
  void *p = realloc(old, bytes ? bytes : 1);
  if (!p) exags_fail(1, file, line, "allocation of %ld bytes failed\n", (long)bytes);
  return p;
}
