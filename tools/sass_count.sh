#!/bin/bash
# static SASS instruction count + opcode histogram of one kernel in an object/.so:  tools/sass_count.sh <file> <mangled-name-regex>
f=$1; pat=$2
cuobjdump -sass "$f" | awk -v pat="$pat" '
/Function : /{on = ($3 ~ pat); if (on) print "== " $3}
on && /^[ \t]+\/\*[0-9a-f]{4}\*\//{ op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); split(op,a,"."); h[a[1]]++; n++ }
END{print "total", n; for (k in h) printf "%5d %s\n", h[k], k | "sort -rn | head -40"}'
