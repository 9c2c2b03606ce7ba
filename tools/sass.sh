#!/bin/bash
# tools/sass.sh <object under sctda_b200/csrc> <substring of the kernel's mangled name> : cleaned SASS of one kernel on stdout
O=/root/repo/sctda_b200/csrc/$1
F=$(cuobjdump -sass "$O" | grep "Function :" | grep -- "$2" | head -1 | sed 's/.*Function : //')
cuobjdump -sass -fun "$F" "$O" | grep -v '^\s*/\* 0x' | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-110 | grep -v '^\s*$'
