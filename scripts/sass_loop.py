"""Instruction mix of the tight stepping loop of a kernel in a cubin: the innermost backward branch whose body holds MUFU.LG2.
usage: sass_loop.py file.cubin substring-of-mangled-kernel-name"""
import collections, re, subprocess, sys
sass = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if sys.argv[2] not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f)
    addr = [int(a, 16) for a, _ in ins]
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= addr[i]:
                j = addr.index(tgt) if tgt in addr else None
                if j is not None and any("MUFU.LG2" in x[1] for x in ins[j:i + 1]):
                    if best is None or i - j < best[1] - best[0]:
                        best = (j, i)
    if best is None:
        print(name, "no loop"); continue
    body = ins[best[0]:best[1] + 1]
    mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
    print(name, len(body), "instructions")
    print("  ", dict(mix.most_common()))
    if len(sys.argv) > 3:
        for a, t in body: print("    ", a, t)
