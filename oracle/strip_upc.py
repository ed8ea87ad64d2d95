#!/usr/bin/env python3
"""strip_upc.py IN OUT -- write a copy of a reference .upc/.h file with the UPC
type qualifiers (`shared`, `shared[]`, `shared [n]`, `strict`, `relaxed`)
removed outside string literals and comments, so that plain gcc can compile it
against oracle/shim/upc.h.  Used only by build_ref.sh, on temporary copies
that are deleted after the build (reference sources are never stored in this
repository).  Two further mechanical edits, both listed in SURVEY.md App. A:
file-scope arrays dimensioned [THREADS] get a fixed bound (THREADS is a
run-time value in the shim), and comm_scan/comm_allreduce always take their
hand-rolled put/flag paths (the UPC-collective paths index cyclic shared
arrays, which a flat arena cannot mimic)."""
import re
import sys

src = open(sys.argv[1]).read()
out, i, n = [], 0, len(src)
qual = re.compile(r"\b(shared|strict|relaxed)\b(\s*\[[^\]]*\])?")
while i < n:
    if src.startswith("/*", i):
        j = src.find("*/", i + 2); j = n if j < 0 else j + 2
        out.append(src[i:j]); i = j
    elif src.startswith("//", i):
        j = src.find("\n", i); j = n if j < 0 else j
        out.append(src[i:j]); i = j
    elif src[i] == '"':
        j = i + 1
        while j < n and src[j] != '"':
            j += 2 if src[j] == "\\" else 1
        out.append(src[i:j + 1]); i = j + 1
    elif src[i] == "'":
        j = i + 1
        while j < n and src[j] != "'":
            j += 2 if src[j] == "\\" else 1
        out.append(src[i:j + 1]); i = j + 1
    else:
        m = re.compile(r"[^\"'/]+|/").match(src, i)
        out.append(qual.sub("", m.group(0))); i = m.end()
text = "".join(out)
text = re.sub(r"^(static\s[^;\n]*)\[THREADS\];", r"\1[EXAGS_SHIM_MAXTHREADS];", text, flags=re.M)
text = text.replace("if (vn > 1) goto comm_scan_byhand;", "goto comm_scan_byhand;")
text = text.replace("if (vn > 1) goto comm_allreduce_byhand;", "goto comm_allreduce_byhand;")
open(sys.argv[2], "w").write(text)
