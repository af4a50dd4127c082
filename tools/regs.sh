#!/bin/bash
# usage: tools/regs.sh <lib.so> [grep-pattern]   -> kernel, registers, stack(spill) bytes
cuobjdump --dump-resource-usage "$1" 2>/dev/null | awk '/Function/{f=$2} /REG:/{print f, $0}' | sed -E 's/:? +REG:([0-9]+) STACK:([0-9]+).*/ REG=\1 STACK=\2/' | while read f rest; do echo "$(echo $f | sed 's/:$//' | c++filt | sed 's/(QuartetTask)//; s/void //') $rest"; done | grep -E "${2:-eri_jk}" | sort
