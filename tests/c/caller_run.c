/* A C caller written against the reference's header names and spellings
 * (the body of tests/gs/gs_test.upc:118-145 for one rank, plus the known answers
 * of SURVEY.md 3.4), linked against libexags.so and run on the GPU box by
 * tests/test_gs_gpu.py::test_c_caller_links_and_runs.  Exits 0 when every check holds. */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include "exags/c99.h"
#include "exags/name.h"
#include "exags/fail.h"
#include "exags/types.h"
#include "exags/comm.h"
#include "exags/mem.h"
#include "exags/gs_defs.h"
#include "exags/gs.h"
#include "exags_cuda.h"

static int check(const char *what, const double *got, const double *want, int n)
{
  for (int i = 0; i < n; ++i)
    if (got[i] != want[i]) { printf("FAIL %s: [%d] = %g, want %g\n", what, i, got[i], want[i]); return 1; }
  return 0;
}

int main(void)
{
  comm_ptr cp; struct gs_data *gsh; buffer b = null_buffer;
  int bad = 0, np = 0, id = -1;
  comm_init(); comm_world(&cp);
  comm_np(cp, &np); comm_id(cp, &id);
  if (np != 1 || id != 0) { printf("FAIL rank identity %d/%d\n", id, np); return 1; }

  /* gs_test.upc fixture for np = 1: v[np+3] becomes 3 (t=0) and np+3 (t=1) */
  {
    slong *ids = tmalloc(slong, np + 4);
    double v[5];
    ids[0] = -(np + 10 + 3 * id);
    for (int i = 0; i < np; ++i) ids[i + 1] = -(i + 1);
    ids[np + 1] = id + 1; ids[np + 2] = id + 1; ids[np + 3] = np - id;
    gsh = gs_setup(ids, np + 4, cp, 0, gs_pairwise, 0);
    free(ids);                                    /* setup keeps what it needs */
    for (int t = 0; t < 2; ++t) {
      for (int i = 0; i < np + 4; ++i) v[i] = 1;
      gs(v, gs_double, gs_add, t, gsh, &b);
      double want = t ? np + 3 : 3;
      if (v[np + 3] != want) { printf("FAIL fixture t=%d: %g\n", t, v[np + 3]); bad = 1; }
    }
    gs_free(gsh);
  }
  /* known answers (reference output) on a flagged label set, host and device pointers */
  {
    static const slong lab[16] = { -11, -1, 1, 1, 1, 5, 7, 5, -7, 7, 0, 9, -9, -9, 3, -3 };
    static const double w0[16] = { 0, 12, 12, 12, 12, 14, 17, 14, 17, 17, 11, 12, 12, 12, 15, 15 };
    static const double w1[16] = { 1, 2, 14, 14, 14, 14, 26, 14, 9, 26, 11, 39, 13, 14, 31, 16 };
    double v[16];
    gsh = gs_setup(lab, 16, cp, 0, gs_auto, 0);
    for (int i = 0; i < 16; ++i) v[i] = i + 1;
    gs(v, gs_double, gs_add, 0, gsh, 0);
    bad |= check("known t=0 (host pointer)", v, w0, 16);
    double *d = (double *)exags_device_malloc(sizeof v);
    for (int i = 0; i < 16; ++i) v[i] = i + 1;
    exags_memcpy_h2d(d, v, sizeof v);
    gs(d, gs_double, gs_add, 1, gsh, 0);
    exags_memcpy_d2h(v, d, sizeof v);
    bad |= check("known t=1 (device pointer)", v, w1, 16);
    exags_device_free(d);
    /* gs_vec on int tuples, gs_many on two float arrays */
    int iv[32]; float f0[16], f1[16]; void *many[2] = { f0, f1 };
    for (int i = 0; i < 16; ++i) { iv[2 * i] = i; iv[2 * i + 1] = -i; f0[i] = (float)i; f1[i] = 1.0f; }
    gs_vec(iv, 2, gs_int, gs_max, 1, gsh, 0);
    if (iv[2 * 2] != 4 || iv[2 * 2 + 1] != -1) { printf("FAIL gs_vec max: %d %d\n", iv[4], iv[5]); bad = 1; }
    gs_many(many, 2, gs_float, gs_add, 1, gsh, 0);
    if (f1[2] != 4.0f || f0[5] != 12.0f) { printf("FAIL gs_many add: %g %g\n", f1[2], f0[5]); bad = 1; }
    gs_free(gsh);
  }
  if (exags_kernel_launches() == 0) { printf("FAIL: no kernel was launched\n"); bad = 1; }
  comm_free(&cp);
  buffer_free(&b);
  if (!bad) printf("caller_run ok (%llu kernel launches)\n", exags_kernel_launches());
  return bad;
}
