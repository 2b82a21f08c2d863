#!/usr/bin/env python
"""Phase table of the front kernel from an `ncu --page source --csv --print-source cuda,sass` export: stall samples, executed warp
instructions and FP64 instructions per front-step by region of ccp_kernel.cuh (regions are found by marker strings, so the table
follows the source).  Usage: ncu_phases.py export.csv n_fronts steps > profiles/...txt"""
import csv, sys, collections, re, os
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "computing-conjugate-points_b200", "csrc", "ccp_kernel.cuh")
MARKS = [("per-SM alignment barrier", "struct __align__(16) SmAlign"), ("determinant helpers", "// ---- determinants"),
         ("ipfma and small helpers", "__device__ __forceinline__ void ipfma"), ("Cauchy passes (pass_early / pass_late / pairs)", "// Cauchy tasks."),
         ("sign-aware parameter products", "// product with a parameter interval of known sign"), ("second_variation", "__device__ __noinline__ void second_variation"),
         ("set_update", "__device__ __noinline__ int set_update"), ("renormalise", "__device__ __noinline__ void renormalise"),
         ("phi_prime / deriv_entries (cold)", "__device__ __noinline__ void phi_prime"), ("lane constants", "struct LaneC"),
         ("table_loop: stage code", "__device__ __noinline__ void table_loop"), ("step: alignment wait, hull, order 0", "__device__ __noinline__ int step_body"),
         ("step: image y, Kc, Ws", "// ---- step image: y = Phi(x)"), ("step: first/last frame, frame tables F_k", "// ---- first / last frame data"),
         ("step: grid evaluation", "// ---- grid loop (utils.cpp:219-252)"), ("step: determinants", "// ---- determinants (topFrame::constructDetSeries)"),
         ("step: decisions, tallies", "// ---- topFrame::countZeros decision per sub-interval"), ("step: tail", "        PROF(5);"),
         ("run_front (per front)", "// One CTA (NT threads) = one front.")]
def main():
    path, nf, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    src = open(SRC).read().split("\n")
    marks = []
    for name, pat in MARKS:
        for i, l in enumerate(src):
            if pat in l: marks.append((i + 1, name)); break
    marks.sort()
    def region(line):
        r = marks[0][1]
        for l, n in marks:
            if line >= l: r = n
        return r
    seen = {}; f = None; hdr = None; cur = None
    for r in csv.reader(open(path, newline="")):
        if not r: continue
        if r[0] == "File Path": f = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; continue
        if r[0] != "": cur = int(r[0]); continue
        if r[2].startswith("0x"):
            a = int(r[2], 16)
            if a in seen: continue
            d = dict(zip(hdr[2:], r[2:]))
            num = lambda k: float(d.get(k, "0").replace(",", "") or 0)
            seen[a] = (f, cur, num("# Samples"), num("Instructions Executed"), r[3].strip())
    last = marks[0][1]; S = collections.Counter(); I = collections.Counter(); F = collections.Counter(); L = collections.Counter()
    for a in sorted(seen):
        f, l, s, e, sx = seen[a]
        if f == "ccp_kernel.cuh": last = region(l)
        op = re.sub(r"^@!?U?P\d+\s+", "", sx).split()[0].split(".")[0]
        S[last] += s; I[last] += e
        if op in ("DFMA", "DADD", "DMUL", "DSETP"): F[last] += e
        if op in ("LDS", "STS"): L[last] += e
    base = nf * steps; ts = sum(S.values())
    print("# %d fronts, %d steps; per front-step: warp instructions (both warps), FP64 and shared-memory instructions; share of the stall samples" % (nf, steps))
    print("%-52s %9s %9s %9s %9s" % ("region", "samples", "inst", "fp64", "lds/sts"))
    for _, n in marks:
        if S[n] or I[n]: print("%-52s %8.1f%% %9.0f %9.0f %9.0f" % (n, 100 * S[n] / ts, I[n] / base, F[n] / base, L[n] / base))
    print("%-52s %8.1f%% %9.0f %9.0f %9.0f" % ("total", 100.0, sum(I.values()) / base, sum(F.values()) / base, sum(L.values()) / base))
if __name__ == "__main__":
    main()
