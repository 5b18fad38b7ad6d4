#!/bin/bash
# static SASS size (instructions) of every function in a built library: tools/sass_size.sh <lib.so>
d=$(mktemp -d); cd $d; cuobjdump -xelf all "$1" >/dev/null; nvdisasm -g -c augb200.sm_100a.cubin > dis.txt 2>/dev/null
awk '/^\/\/-+ \.text\./{name=$2} /^\s*\/\*[0-9a-f]{4,}\*\//{n[name]++} END{for(k in n) print n[k], k}' dis.txt | sort -rn | head -${2:-12}
cp dis.txt /tmp/sass/last_dis.txt
