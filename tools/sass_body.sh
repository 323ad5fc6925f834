#!/bin/bash
# usage: tools/sass_body.sh <object> <kernel-substring>  -> condensed SASS listing (one instruction per line, numbered)
cuobjdump -sass "$1" | awk -v k="$2" '/Function :/ {on = index($0, k) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/\/\* 0x[0-9a-f]* \*\///' | awk '{ $1=""; print NR": "$0}'
