"""Summarise `nvcc -Xptxas -v` output: one line per kernel (registers, spills, smem)."""
import re
import subprocess
import sys

txt = sys.stdin.read()
cur, spill = None, ""
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)[:90]
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
    if m:
        spill = f"stack={m.group(1)} spill={m.group(2)}"
    m = re.search(r"Used (\d+) registers(.*)", line)
    if m:
        sm = re.search(r"(\d+) bytes smem", m.group(2))
        print(f"{cur:92s} regs={m.group(1):>3s} {spill} smem={sm.group(1) if sm else 0}")
    if "error" in line or "warning" in line:
        print(line)
