"""Prints registers / stack / spills per kernel from the nvcc -Xptxas -v logs of the last build."""
import re
import subprocess
import sys
import os
B = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym_art_b200", "_build")
import glob
for f in sorted(os.path.basename(x) for x in glob.glob(os.path.join(B, "*.o.log"))):
    txt = open(os.path.join(B, f)).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
        dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("qs::", "")[:100]
        print(f[:-6], m.group(5), "regs stack", m.group(2), "spill", m.group(3), m.group(4), dem)
