#!/usr/bin/env python
"""Static SASS size of a kernel by CUDA source line range (code-footprint / I-cache analysis).
Usage: sass_size.py <cubin> <kernel-name-substring> <file.cu> [start:end:label ...]
Needs -lineinfo; uses `nvdisasm --print-line-info`.  Inlined code is attributed to the innermost line."""
import collections
import re
import subprocess
import sys


def main():
    cubin, kern, src = sys.argv[1:4]
    ranges = []
    for a in sys.argv[4:]:
        s, e, name = a.split(":")
        ranges.append((int(s), int(e), name))
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    func, cur = None, None
    cnt = collections.Counter()
    base = src.split("/")[-1]
    for line in txt.splitlines():
        m = re.match(r"\.text\.(\S+):", line)
        if m:
            func, cur = m.group(1), None
            continue
        m = re.search(r'//## File "([^"]*)", line (\d+)', line)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith(base) else -1
            continue
        if func and kern in func and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+[A-Z@]", line):
            cnt[cur if cur is not None else -1] += 1
    tot = sum(cnt.values())
    print("%s: %d SASS instructions = %.1f KB" % (kern, tot, tot * 16 / 1024))
    if ranges:
        for s, e, name in ranges:
            n = sum(v for k, v in cnt.items() if s <= k < e)
            print("  %-28s %6d  %5.1f%%" % (name, n, 100.0 * n / max(1, tot)))
        print("  %-28s %6d" % ("(other files / no line)", sum(v for k, v in cnt.items() if k < 0)))
    else:
        for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
            print("  line %5d  %5d" % (k, v))


if __name__ == "__main__":
    main()
