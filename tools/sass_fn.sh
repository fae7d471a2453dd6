#!/bin/bash
# usage: tools/sass_fn.sh <substring of mangled kernel name>  -> SASS of that kernel only (encodings stripped)
SO=${SO:-/root/repo/linr-pcgc_b200/lib/liblinr_b200.so}
cuobjdump -sass "$SO" | awk -v pat="$1" '/Function :/ {on = index($0, pat) > 0} on' | grep -v "^\s*/\* 0x" | sed -e 's#/\* 0x[0-9a-f]* \*/##'
