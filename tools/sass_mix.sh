#!/bin/bash
# usage: tools/sass_mix.sh <mangled-name-substring> [lib]  -- static SASS opcode mix of one kernel of libgik.so
LIB=${2:-motion-planning-and-control-for-dual-manipulator-robot_b200/libgik.so}
cuobjdump -sass "$LIB" | awk -v pat="$1" '/Function :/{on=index($0,pat)>0} on' > /tmp/k.sass
echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4,5}\*/' /tmp/k.sass)"
grep -oE "^\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" /tmp/k.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-16}
