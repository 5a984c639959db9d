"""dev helper: print a kernel's SASS with the decoded scheduling control fields (stall count, write/read scoreboard index,
scoreboard wait mask) of the 128-bit Volta+ instruction encoding, to see which loads share a scoreboard.

    python tools/sass_ctrl.py <lib.so> <kernel-name-substring> [first_addr_hex last_addr_hex]
"""
import re
import subprocess
import sys


def main():
    lib, name = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 62
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    on = False
    pend = None
    for line in out:
        if "Function :" in line:
            on = name in line
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", line)
        if m:
            pend = (int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16))
            continue
        m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", line)
        if m and pend:
            addr, text, _ = pend
            hi64 = int(m.group(1), 16)
            stall = (hi64 >> 41) & 0xF
            wr = (hi64 >> 46) & 7
            rd = (hi64 >> 49) & 7
            wait = (hi64 >> 52) & 0x3F
            if lo <= addr <= hi:
                print(f"{addr:05x} st={stall:2d} wr={'-' if wr == 7 else wr} rd={'-' if rd == 7 else rd} "
                      f"wait={''.join(str(i) for i in range(6) if wait >> i & 1) or '-':6s} {text}")
            pend = None


if __name__ == "__main__":
    main()
