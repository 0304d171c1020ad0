#!/bin/bash
# developer tool: compact line-annotated disassembly of one flight kernel instance.  usage: tools/sass_compact.sh <lib.so> <out.txt> [name-substring]
lib=$(realpath $1); out=$2; want=${3:-emc_flight_kernelILi128ELi3ELi2ELi1ELi1}
tmp=$(mktemp -d); (cd $tmp && cuobjdump -xelf all $lib >/dev/null && nvdisasm -g -c *.cubin > g.sass 2>/dev/null)
python3 - "$tmp/g.sass" "$out" "$want" <<'PY'
import re, sys
L = open(sys.argv[1]).read().split('\n')
st = [i for i, l in enumerate(L) if l.startswith('.text.') and sys.argv[3] in l][0]
en = [i for i, l in enumerate(L) if i > st and l.startswith('//---------------------')][0]
cur = ''; out = []
for l in L[st:en]:
    m = re.search(r'//## File ".*/emc_([a-z_]+)\.cuh?", line (\d+)', l)
    if m: cur = f"{m.group(1)[:4]}:{m.group(2)}"; continue
    m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: out.append(f"{m.group(1)} {cur:10s} {m.group(2).strip()}")
    elif l.startswith('.L_'): out.append(l)
open(sys.argv[2], 'w').write('\n'.join(out))
print(len(out), 'lines ->', sys.argv[2])
PY
rm -rf $tmp
