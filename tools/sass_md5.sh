#!/bin/bash
# One md5 per kernel of the built library's SASS (encodings stripped).  A refactor of device code that must not move
# performance is checked without a GPU by comparing these before and after.  usage: tools/sass_md5.sh [lib.so]
lib=${1:-$(dirname "$0")/../iteres_b200/csrc/libiteres_gpu.so}
cuobjdump -sass "$lib" | awk '
    /Function : / { if (f != "") close(cmd); f = $3; cmd = "md5sum | cut -c1-32 | sed \"s/$/  " f "/\"" ; next }
    f != "" && !/^[[:space:]]*$/ { line = $0; sub(/\/\* 0x[0-9a-f]* \*\//, "", line); print line | cmd }
    END { if (f != "") close(cmd) }'
