"""Per-kernel summary of an ncu source page (SASS): where the stall samples sit (main loop vs staging vs rest).
    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_source_hot.py src.csv [kernel substr]"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
kern, hdr, blocks = None, None, collections.OrderedDict()
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; blocks[kern] = []; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if kern and hdr and len(r) == len(hdr):
        blocks[kern].append(dict(zip(hdr, r)))
for k, ins in blocks.items():
    if want not in k: continue
    tot = sum(int(i["# Samples"]) for i in ins)
    # find loops: backward branches
    addr = [int(i["Address"], 16) for i in ins]
    loops = []
    for n, i in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", i["Source"])
        if m:
            t = int(m.group(1), 16)
            t = addr[0] + t if t < addr[0] else t
            if t <= addr[n] and t in addr: loops.append((addr.index(t), n))
    print("====", k[:60], "samples", tot, "instr", len(ins))
    for lo, hi in loops:
        body = ins[lo:hi + 1]
        s = sum(int(i["# Samples"]) for i in body)
        fp = sum(1 for i in body if re.search(r"\b(DFMA|DMUL|DADD)\b", i["Source"]))
        ex = max(int(i["Instructions Executed"]) for i in body)
        st = collections.Counter()
        for i in body:
            for key in i:
                if key.startswith("stall_") and "Not Issued" not in key: st[key] += int(i[key] or 0)
        top = ", ".join(f"{a[6:]} {100*b/max(1,s):.0f}%" for a, b in st.most_common(6))
        print(f"  loop [{lo:5d},{hi:5d}] n={hi-lo+1:4d} fp64={fp:4d} exec={ex:10d} samples={100*s/tot:5.1f}%  {top}")
