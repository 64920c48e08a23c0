"""Decode the scheduling control fields (stall count, scoreboard set / wait masks) of a kernel's SASS:
python scripts/sass_ctrl.py <object or .so> <mangled kernel name>   -> one line per instruction."""
import re, subprocess, sys
obj, fun = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout.split("\n")
pat = re.compile(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/')
pat2 = re.compile(r'^\s+/\* (0x[0-9a-f]{16}) \*/')
i = 0
while i < len(txt):
    m = pat.match(txt[i])
    if m and i + 1 < len(txt) and pat2.match(txt[i + 1]):
        hi = int(pat2.match(txt[i + 1]).group(1), 16)
        c = (hi >> 41) & 0x7fffff
        stall, wr, rd, wait = c & 0xf, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 0x3f
        w = ''.join(str(b) for b in range(6) if wait >> b & 1)
        print(f"{m.group(1)} st={stall:2d} W={'-' if wr == 7 else wr} R={'-' if rd == 7 else rd} wait=[{w:6s}] {m.group(2)[:100]}")
        i += 2
    else:
        i += 1
