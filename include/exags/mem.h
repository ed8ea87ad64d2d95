/* exags/mem.h -- source-compatibility shim for the allocation helpers that
 * reference callers use around gs (tmalloc, array_*, buffer_*; the reference
 * defines them in src/mem.h:111-163).  Written against ../exags.h's
 * `struct array`; allocation failure funnels into fail() like the reference. */
#ifndef EXAGS_COMPAT_MEM_H
#define EXAGS_COMPAT_MEM_H
#ifndef EXAGS_SHORT_NAMES
#define EXAGS_SHORT_NAMES 1
#endif
#include <stdlib.h>
#include <string.h>
#include "../exags.h"

static inline void *exags_xalloc_(void *old, size_t bytes, int zero,
                                  const char *file, unsigned line)
{
  void *p = old ? realloc(old, bytes) : (zero ? calloc(bytes ? bytes : 1, 1)
                                              : malloc(bytes));
  if (!p && bytes)
    exags_fail(1, file, line, "allocation of %ld bytes failed\n", (long)bytes);
  return p;
}
#define tmalloc(type, count) \
  ((type *)exags_xalloc_(0, (size_t)(count) * sizeof(type), 0, __FILE__, __LINE__))
#define tcalloc(type, count) \
  ((type *)exags_xalloc_(0, (size_t)(count) * sizeof(type), 1, __FILE__, __LINE__))
#define trealloc(type, ptr, count) \
  ((type *)exags_xalloc_((ptr), (size_t)(count) * sizeof(type), 0, __FILE__, __LINE__))

#define null_array  {0, 0, 0}
#define null_buffer {0, 0, 0}

static inline void *exags_array_reserve_(struct array *a, size_t want,
                                         size_t elem, const char *file,
                                         unsigned line)
{
  if (a->max < want) {
    size_t grown = a->max + a->max / 2 + 1;
    if (grown < want) grown = want;
    a->ptr = exags_xalloc_(a->ptr, grown * elem, 0, file, line);
    a->max = grown;
  }
  return a->ptr;
}
#define array_init(T, a, cap) \
  ((a)->n = 0, (a)->max = (cap), \
   (a)->ptr = exags_xalloc_(0, (size_t)(cap) * sizeof(T), 0, __FILE__, __LINE__))
#define array_resize(T, a, cap) \
  ((a)->max = (cap), \
   (a)->ptr = exags_xalloc_((a)->ptr, (size_t)(cap) * sizeof(T), 0, __FILE__, __LINE__))
#define array_reserve(T, a, want) \
  exags_array_reserve_((a), (want), sizeof(T), __FILE__, __LINE__)
#define array_free(a)        (free((a)->ptr))
#define buffer_init(b, cap)    array_init(char, b, cap)
#define buffer_resize(b, cap)  array_resize(char, b, cap)
#define buffer_reserve(b, cap) array_reserve(char, b, cap)
#define buffer_free(b)         array_free(b)
#endif
