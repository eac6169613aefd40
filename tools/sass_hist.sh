#!/bin/bash
# usage: tools/sass_hist.sh <lib.so> <mangled-function-substring>  -> opcode histogram of one kernel
lib=$1; fn=$2
cuobjdump -sass "$lib" | awk -v fn="$fn" '/Function :/{on=index($0,fn)>0} on' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | awk '{op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); print op}' | sed -E 's/\.(.*)//' | sort | uniq -c | sort -rn
