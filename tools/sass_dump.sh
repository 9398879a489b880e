#!/bin/bash
# usage: sass_dump.sh <lib.so> <kernel-substr> : plain SASS listing of one kernel (one instruction per line)
cuobjdump -sass "$1" | awk -v k="$2" '/Function : /{f = index($0, k) > 0} f' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#/\*[0-9a-f]{4,5}\*/##; s#/\* 0x[0-9a-f]+ \*/##'
