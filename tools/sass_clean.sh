#!/bin/bash
# usage: sass_clean.sh <object> <mangled kernel name>  -> clean "addr  instr" listing on stdout
cuobjdump -sass -fun "$2" "$1" | grep -E '^\s+/\*[0-9a-f]{4,}\*/' | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1  /; s/\s*;\s*\/\*.*$//'
