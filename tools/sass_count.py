#!/usr/bin/env python
"""Static SASS census of a cubin / .so: instructions per kernel and the opcode histogram (no GPU needed).

    python tools/sass_count.py file.cubin [kernel-substring] [--hist]
"""
import collections
import re
import subprocess
import sys


def census(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    fn, res = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            res[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            if op == "NOP":
                continue
            res[fn][op] += 1
    return res


if __name__ == "__main__":
    res = census(sys.argv[1])
    pat = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    for fn, c in res.items():
        if pat not in fn:
            continue
        print("%-60s %6d" % (fn[:60], sum(c.values())))
        if "--hist" in sys.argv:
            for op, n in c.most_common():
                print("      %-28s %5d" % (op, n))
