#!/bin/bash
# instruction mix of k_assemble_fan<KC,R4> ($1 = mangled-name regex) in the built library
F=$(cuobjdump -sass finite_elements_b200/libfe_b200.so | grep -o "Function : .*k_assemble_fan$1.*" | head -1 | sed 's/Function : //')
cuobjdump -sass -fun "$F" finite_elements_b200/libfe_b200.so > /tmp/fan.sass
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/fan.sass | sed 's/^\s*\/\*[0-9a-f]*\*\/\s*//' | sed 's/@!\?U\?P[0-9T] //' | awk '{print $1}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn | awk '{printf "%s:%s ", $2, $1} END {print ""}'
grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/fan.sass
