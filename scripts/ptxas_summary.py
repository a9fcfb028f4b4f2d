"""Summarise a *.ptxas.log: demangled kernel name, registers, spill bytes."""
import re, subprocess, sys
s = open(sys.argv[1]).read()
pat = sys.argv[2] if len(sys.argv) > 2 else ''
ents = re.findall(r"Compiling entry function '([^']+)'.*?\n(?:.*\n)*?.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", s)
dem = subprocess.run(['cu++filt'] + [e[0] for e in ents], capture_output=True, text=True).stdout.splitlines()
for d, e in zip(dem, ents):
    d = re.sub(r'\(CUtensorMap_st.*', '', d).replace('sf::', '').replace('void ', '')
    if pat in d:
        print(d[:90].ljust(92), 'regs', e[4], 'spill st/ld', e[2], e[3])
