#!/usr/bin/env python
"""Where the step loops of the three GBM headline kernels (M=4, default sampler) start in the built library:
address, length in instructions, offset within a 64/128-byte line.  Diagnostic for the 1-2 % run-time
shifts these kernels show when unrelated code in the same kernel changes."""
import subprocess,re,sys
lib='/root/repo/multilevel_mc_sdes_b200/csrc/libmlmcb.so'
out=subprocess.run(['cuobjdump','-sass',lib],capture_output=True,text=True).stdout
cur=None; funcs={}
for line in out.splitlines():
    m=re.search(r'Function : (\S+)',line)
    if m: cur=m.group(1); funcs[cur]=[]; continue
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',line)
    if m and cur: funcs[cur].append((int(m.group(1),16),m.group(2)))
for name,ins in funcs.items():
    if not re.search(r'level_kernelIdLi0ELi[012]ELi4ELi0ELi0E',name): continue
    best=None
    for a,t in ins:
        m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)',t)
        if m:
            tgt=int(m.group(1),16)
            if tgt<a and (best is None or a-tgt>best[1]-best[0]) and a-tgt<0x2000: best=(tgt,a)
    print(name[:60], 'loop 0x%x..0x%x'%best, 'len', (best[1]-best[0])//16+1, 'start mod 128 =', best[0]%128, 'mod 64 =',best[0]%64)
