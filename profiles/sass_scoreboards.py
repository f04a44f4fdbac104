"""Decode the scoreboard fields of every SASS instruction of one kernel in libgca.so.

usage: python profiles/sass_scoreboards.py <kernel-name-substring> [lib]  -> lines "addr wr=<sb> rd=<sb> wait=<mask> instr"
(control bits of the 128-bit Volta+ encoding: stall 105-108, yield 109, write-SB 110-112, read-SB 113-115,
wait mask 116-121).  Used to check which in-flight loads a consumer really waits for.
"""
import re, subprocess, sys

def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, body = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1); body[cur] = []
        elif cur:
            body[cur].append(ln)
    return body

def decode(lines):
    i = 0
    while i < len(lines) - 1:
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*/\* (0x[0-9a-f]{16}) \*/", lines[i])
        m2 = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", lines[i + 1]) if m else None
        if m and m2:
            ctrl = (int(m2.group(1), 16) >> 41) & 0x7fffff
            yield m.group(1), m.group(2), ctrl & 0xf, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3f
            i += 2
        else:
            i += 1

if __name__ == "__main__":
    pat = sys.argv[1]
    if len(sys.argv) > 2:
        lib = sys.argv[2]
    else:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from gconv_adapter_b200.build import lib_path
        lib = lib_path()
    for name, lines in kernels(lib).items():
        if pat in name:
            print("==", name)
            for a, t, stall, wr, rd, wait in decode(lines):
                if any(k in t for k in ("LDG", "STG", "HMMA", "BAR", "LDS", "STS")) or wait:
                    print(a, f"wr={wr if wr != 7 else '-'} rd={rd if rd != 7 else '-'} wait={wait:06b}", t[:100])
