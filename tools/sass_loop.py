#!/usr/bin/env python
"""Opcode histogram of the largest loop (backward branch with the widest span) of a kernel's SASS.
    cuobjdump -sass -fun KERNEL file.o | python tools/sass_loop.py [list | START_HEX]
list: every backward branch with its span; START_HEX: the loop that starts at that address instead of the widest."""
import collections
import re
import sys

ins = []
for line in sys.stdin:
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
best = (0, 0, 0)
want = sys.argv[1] if len(sys.argv) > 1 else None
for a, op, rest in ins:
    if op.startswith("BRA"):
        t = re.search(r'0x([0-9a-f]+)', rest)
        if t and int(t.group(1), 16) < a:
            span = (a - int(t.group(1), 16), int(t.group(1), 16), a)
            if want == "list":
                print("0x%x .. 0x%x: %d instructions" % (span[1], span[2], span[0] // 16 + 1))
            elif want is not None and int(want, 16) == span[1]:
                best = span
            elif want is None and span[0] > best[0]:
                best = span
if want == "list":
    sys.exit(0)
c = collections.Counter(op.split('.')[0] for a, op, _ in ins if best[1] <= a <= best[2])
n = sum(c.values())
print("loop 0x%x .. 0x%x: %d instructions" % (best[1], best[2], n))
for k, v in c.most_common(24):
    print("%6d %s" % (v, k))
