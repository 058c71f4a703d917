#!/bin/bash
# usage: tools/sass_sizes.sh build/obj/kernels_f32.o  -> "<instructions> <registers> <function>" per kernel
# (offline check that a refactor left the tuned kernels' code unchanged; no GPU needed)
obj=${1:-build/obj/kernels_f32.o}
cuobjdump -sass -res-usage "$obj" 2>/dev/null | awk '
/Function/ { if (name != "") print n, reg, name; name=$3; n=0; reg="?" }
/REG:/ { match($0, /REG:[0-9]+/); reg=substr($0, RSTART+4, RLENGTH-4) }
/^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ { n++ }
END { print n, reg, name }' | sort -k3
