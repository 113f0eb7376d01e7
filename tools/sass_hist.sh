#!/bin/bash
# usage: tools/sass_hist.sh <mangled-function-substring> [top-n]: opcode histogram of one kernel from the built .so
cuobjdump -sass xsqueezeit_b200/libxsi_b200.so | awk -v pat="$1" '
/Function :/ { on = index($0, pat) > 0 }
on && /^ +\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\// { print }' > /tmp/sass_fn.txt
echo "instructions: $(wc -l < /tmp/sass_fn.txt)"
sed -E 's/^ +\/\*[0-9a-f]{4}\*\/ +//' /tmp/sass_fn.txt | sed -E 's/^@!?U?P[0-9T] +//' | awk '{print $1}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -${2:-22}
