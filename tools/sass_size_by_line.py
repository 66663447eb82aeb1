#!/usr/bin/env python
"""Static code size of a kernel per source line (instructions x 16 B), from nvdisasm's line table of the built library.

    python tools/sass_size_by_line.py vx_visibility [top]
"""
import collections
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_lines import line_table  # noqa: E402


def main():
    kernel = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    table = line_table(kernel)
    per = collections.Counter(table.values())
    total = sum(per.values())
    print(f"{kernel}: {total} instructions = {total * 16 / 1024:.1f} KB")
    for (f, ln), n in per.most_common(top):
        print(f"{n:6d} instr {n * 16 / 1024:6.1f} KB  {f}:{ln}")


if __name__ == "__main__":
    main()
