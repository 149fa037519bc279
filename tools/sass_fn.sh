#!/bin/bash
# usage: tools/sass_fn.sh <lib.so> <substring of mangled name>  -> SASS of the first matching function, encodings stripped
cuobjdump -sass "$1" | awk -v pat="$2" '
/Function :/ { on = (index($0, pat) > 0 && !done); if (on) { print; seen = 1 } else if (seen) { done = 1 } next }
on { print }' | sed -e 's#/\* 0x[0-9a-f]* \*/##' | grep -v "^\s*$"
