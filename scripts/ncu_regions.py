#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by code region.

Every SASS instruction is attributed to the ccp_kernel.cuh line that is nearest BEFORE it in address order (instructions of
inlined helpers from ccp_interval.cuh inherit the region of the surrounding kernel code), then lines are bucketed into the
phases of a step.  Usage: ncu_regions.py export.csv [--lines]"""
import csv, sys, collections

def load(path):
    rows = []; f = None; hdr = None; cur_line = None
    for r in csv.reader(open(path, newline='')):
        if not r: continue
        if r[0] == "File Path": f = r[1].split('/')[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; continue
        if r[0] != "": cur_line = int(r[0]); continue
        if r[2].startswith("0x"):
            d = dict(zip(hdr[2:], r[2:]))
            def num(k):
                try: return float(d.get(k, "0").replace(",", ""))
                except ValueError: return 0.0
            rows.append(dict(addr=int(r[2], 16), sass=r[3].strip(), file=f, line=cur_line, samples=num("# Samples"), inst=num("Instructions Executed"),
                             thr=num("Thread Instructions Executed"), wf=num("L1 Wavefronts Shared"), wfi=num("L1 Wavefronts Shared Ideal"),
                             st={k: num(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}))
    rows.sort(key=lambda x: x["addr"])
    return rows

PHASES = None
def phase_of(line, table):
    for lo, hi, name in table:
        if lo <= line <= hi: return name
    return "other"

def main():
    rows = load(sys.argv[1])
    tot = sum(r["samples"] for r in rows)
    last = None; per_line = collections.Counter(); inst_line = collections.Counter(); wf_line = collections.Counter(); wfi_line = collections.Counter()
    stall_line = collections.defaultdict(collections.Counter)
    for r in rows:
        if r["file"] == "ccp_kernel.cuh": last = r["line"]
        key = last
        per_line[key] += r["samples"]; inst_line[key] += r["inst"]; wf_line[key] += r["wf"]; wfi_line[key] += r["wfi"]
        for k, v in r["st"].items(): stall_line[key][k] += v
    print("total samples %d, SASS instructions %d, warp inst executed %.3e, smem wavefronts %.3e (ideal %.3e)" % (tot, len(rows), sum(inst_line.values()), sum(wf_line.values()), sum(wfi_line.values())))
    print("%6s %7s %6s %10s %10s %10s  top stalls" % ("line", "samples", "pct", "winst", "smem_wf", "wf_ideal"))
    for line, s in sorted(per_line.items(), key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 60]:
        top = ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(1, s)) for k, v in stall_line[line].most_common(3))
        print("%6s %7d %5.1f%% %10.3e %10.3e %10.3e  %s" % (line, s, 100 * s / tot, inst_line[line], wf_line[line], wfi_line[line], top))

if __name__ == "__main__":
    main()
