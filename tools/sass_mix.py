#!/usr/bin/env python
"""Static SASS opcode mix of one kernel in libgeepee_b200.so (substring match on the mangled name)."""
import collections
import re
import subprocess
import sys

so = 'geepee_b200/csrc/libgeepee_b200.so'
pat = sys.argv[1]
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
cur = None
mix = collections.Counter()
n = 0
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            mix[m.group(2).split('.')[0]] += 1
            n += 1
print(pat, 'instructions:', n)
for k, v in mix.most_common(25):
    print('  %-10s %5d' % (k, v))
