#!/bin/bash
# static SASS opcode histogram of one kernel instantiation: tools/sass_hist.sh <mangled-substring>
cuobjdump -sass pde_opt_b200/libpdeopt_b200.so | awk -v pat="$1" '
/Function :/ {on = index($0, pat) > 0}
on && /^\s+\/\*[0-9a-f]+\*\// {op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); c[op]++; n++}
END {for (k in c) printf "%6d %s\n", c[k], k; printf "%6d TOTAL\n", n}' | sort -rn | head -${2:-30}
