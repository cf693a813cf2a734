"""registers / spills / shared memory per kernel from rtsengine_b200/csrc/ptxas.log (diagnostic)"""
import re, subprocess, sys, os
log = open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "rtsengine_b200", "csrc", "ptxas.log")).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", log):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name).replace("void rts::", "")
    smem = re.search(r"(\d+) bytes smem", m.group(6))
    print(f"{name:48s} regs {m.group(5):>3s}  stack {m.group(2):>4s}  spill st/ld {m.group(3):>3s}/{m.group(4):>3s}  smem {smem.group(1) if smem else 0}")
