"""Regenerates the ctypes stub of INTEGRATION.md section 3 from the binding the product itself uses
(torchkge_b200/_lib.py: RankArgs mirrors kge_rank_args_t of include/kge_b200.h), so that the
illustrative struct can never fall behind the ABI.  tests/test_abi.py checks the committed text.

    python scripts/gen_integration_struct.py [--write]
"""
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BEGIN, END = "<!-- BEGIN generated: _RankArgs -->", "<!-- END generated: _RankArgs -->"
NAMES = {ctypes.c_int32: "ctypes.c_int32", ctypes.c_int64: "ctypes.c_int64", ctypes.c_void_p: "ctypes.c_void_p",
         ctypes.c_size_t: "ctypes.c_size_t"}


def block():
    from torchkge_b200 import _lib
    lines = ["```python",
             "class _RankArgs(ctypes.Structure):            # kge_rank_args_t, include/kge_b200.h (ABI v%d)" % _lib.ABI_VERSION,
             "    _fields_ = ["]
    for name, typ in _lib.RankArgs._fields_:
        lines.append('        ("%s", %s),' % (name, NAMES[typ]))
    lines += ["    ]", "```"]
    return "\n".join(lines)


def main():
    path = os.path.join(ROOT, "INTEGRATION.md")
    text = open(path).read()
    new = re.sub(re.escape(BEGIN) + r".*?" + re.escape(END), BEGIN + "\n" + block() + "\n" + END, text, flags=re.S)
    if "--write" in sys.argv:
        open(path, "w").write(new)
    return new == text


if __name__ == "__main__":
    ok = main()
    print("INTEGRATION.md struct block is %s" % ("up to date" if ok else "STALE (run with --write)"))
    sys.exit(0 if ok or "--write" in sys.argv else 1)
