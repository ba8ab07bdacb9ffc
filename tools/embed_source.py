#!/usr/bin/env python
"""Concatenates CUDA headers into one C++ raw string literal (for NVRTC): strips `#pragma once` and local
#include "..." lines.  usage: embed_source.py a.h b.cuh ... > out.inc"""
import sys

out = []
for path in sys.argv[1:]:
    for line in open(path):
        s = line.strip()
        if s == "#pragma once" or s.startswith('#include "'):
            continue
        out.append(line.rstrip("\n"))
text = "\n".join(out)
assert ')LDPCSRC"' not in text
# raw string literals are limited to ~64 KB by some compilers: split into adjacent literals
chunks = [text[i:i + 12000] for i in range(0, len(text), 12000)]
print("\n".join('R"LDPCSRC(' + c + ')LDPCSRC"' for c in chunks))
