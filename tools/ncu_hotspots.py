"""Executed-instruction mix of the kernels in an .ncu-rep captured with --import-source on (source page, SASS view):
per launch the share of warp instructions by pipe class and by control-flow opcode, lanes per instruction by class, and
the stall samples by class.  Usage: python tools/ncu_hotspots.py gpurun_out/x.ncu-rep > profiles/xxx_source_hotspots.md"""
import csv
import io
import subprocess
import sys

from sass_pipes import ALU, FMA, LSU, BR


def klass(op):
    op = op.split(".")[0]
    if op in ("BSSY", "BSYNC", "BRA", "BREAK", "WARPSYNC", "EXIT", "RET", "CALL", "NOP", "BAR"):
        return "branch / convergence"
    if op in ALU:
        return "ALU pipe"
    if op in FMA:
        return "FMA pipe"
    if op in LSU:
        return "LSU"
    if op.startswith("MUFU"):
        return "MUFU"
    return "other"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    launches, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1].split("(")[0], "hdr": None, "rows": []}
            launches.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    print(f"# executed-instruction mix of `{rep}` (ncu source page, SASS)\n")
    # (the page lists every launch once per view; identical neighbours are the same launch)
    uniq = []
    for L in launches:
        if not uniq or uniq[-1]["name"] != L["name"] or uniq[-1]["rows"] != L["rows"]:
            uniq.append(L)
    launches = uniq
    for li, L in enumerate(launches):
        h = L["hdr"]
        i_src, i_inst, i_thr, i_smp = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        agg, ops = {}, {}
        tot = tot_thr = tot_smp = 0
        hot = []
        for r in L["rows"]:
            s = r[i_src].strip()
            parts = s.split()
            if not parts:
                continue
            op = parts[1] if parts[0].startswith("@") else parts[0]
            n, th, sm = int(r[i_inst] or 0), int(r[i_thr] or 0), int(r[i_smp] or 0)
            k = klass(op)
            a = agg.setdefault(k, [0, 0, 0]); a[0] += n; a[1] += th; a[2] += sm
            o = ops.setdefault(op.split(".")[0], [0, 0]); o[0] += n; o[1] += th
            tot += n; tot_thr += th; tot_smp += sm
            hot.append((n, th, sm, s))
        if not tot:
            continue
        print(f"## launch {li}: `{L['name']}` -- {tot / 1e6:.1f} M warp instructions, {tot_thr / tot:.1f} lanes per instruction\n")
        print("| class | share of warp instructions | lanes / instruction | share of stall samples |")
        print("|---|---|---|---|")
        for k, (n, th, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"| {k} | {100 * n / tot:.1f} % | {th / max(n, 1):.1f} | {100 * sm / max(tot_smp, 1):.1f} % |")
        cf = {k: v for k, v in ops.items() if k in ("BSSY", "BSYNC", "BRA", "BREAK", "WARPSYNC", "EXIT")}
        print("\ncontrol flow: " + ", ".join(f"{k} {100 * v[0] / tot:.1f} %" for k, v in sorted(cf.items(), key=lambda kv: -kv[1][0])) + "\n")
        print("hottest instructions (warp executions, lanes, stall samples):\n")
        for n, th, sm, s in sorted(hot, reverse=True)[:10]:
            print(f"    {n / 1e6:7.2f} M  {th / max(n, 1):5.1f}  {sm:6d}  {s}")
        print()


if __name__ == "__main__":
    main()
