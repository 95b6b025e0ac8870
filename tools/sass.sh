#!/bin/bash
# tools/sass.sh <mangled-substring> : SASS of the first kernel of libsihl_b200.so whose mangled name contains the substring
LIB=$(dirname "$0")/../sihl_b200/lib/libsihl_b200.so
FUN=$(cuobjdump -sass "$LIB" 2>/dev/null | grep "Function :" | grep "$1" | head -1 | awk '{print $3}')
cuobjdump -sass -fun "$FUN" "$LIB" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
