#!/bin/bash
# usage: tools/sass_hist.sh <object> <function-substring>   -- opcode histogram of one kernel
cuobjdump -sass "$1" | awk -v pat="$2" '/Function/{f=(index($0,pat)>0)} f' | grep -E "^\s+/\*[0-9a-f]+\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn
