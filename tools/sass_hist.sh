#!/bin/bash
# usage: tools/sass_hist.sh <object> <kernel-name-substring>  -> opcode histogram of that kernel's SASS
cuobjdump -sass "$1" | awk -v pat="$2" '
/Function :/ {p = (index($0, pat) > 0); if (p) print "== " $3}
p && /^ +\/\*[0-9a-f]{4}\*\// {
  op=$2; if (op ~ /^@/) op=$3; split(op,a,"."); o=a[1]; gsub(/;/,"",o); c[o]++; n++ }
END { for (k in c) printf "%5d %s\n", c[k], k | "sort -rn"; close("sort -rn"); print n, "total" }'
