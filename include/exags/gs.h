/* exags/gs.h -- source-compatibility shim: the reference installs a header of
 * this name under $(includedir)/exags/ (src/Makefile.am:1-4).  The whole API
 * lives in ../exags.h; this file only switches on the reference's short
 * spellings so existing callers compile unchanged. */
#ifndef EXAGS_COMPAT_GS_H
#define EXAGS_COMPAT_GS_H
#ifndef EXAGS_SHORT_NAMES
#define EXAGS_SHORT_NAMES 1
#endif
#include "../exags.h"
#endif
