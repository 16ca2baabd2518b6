#!/usr/bin/env python
"""Joins the per-instruction counters of an ncu report (--page source --csv: SASS view) with nvdisasm's line table of the same
cubin, and prints where a kernel's executed instructions and stall samples sit, per source line and per pipe class.

  python profiles/sass_by_line.py gpurun_out/X.ncu-rep rpg_monocular_pose_estimator_b200/libmpe_b200.so p3p_sweep_t1_kernelILb0 [top]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

def main():
    rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # a report holds one section per profiled kernel: "Kernel Name",<name> / header row / one row per SASS instruction
    want = kern.replace("ILb0", "").replace("ILb1", "")
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = next(i for i in starts if want.split("kernel")[0] in rows[i][1].replace("::", ""))
    end = next((j for j in starts if j > sec), len(rows))
    hdr = rows[sec + 1]
    ci = {n: i for i, n in enumerate(hdr)}
    insts = [(r[ci["Source"]].strip(), float(r[ci["Instructions Executed"]] or 0), float(r[ci["# Samples"]] or 0)) for r in rows[sec + 2:end] if len(r) > 5]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    lines = None
    for f in os.listdir(tmp):
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern in dis:
            lines = dis.split("\n"); break
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
    cur = ("?", 0); table = []
    for l in lines[start + 1:]:
        if l.startswith("\t.section") or l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)));
            # inlined-at chains: keep the innermost location (first annotation before an instruction)
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            table.append((cur, m.group(2).strip()))
    n = min(len(table), len(insts))
    per_line = collections.Counter(); per_line_s = collections.Counter(); per_class = collections.Counter()
    tot = sum(i[1] for i in insts[:n]); tots = sum(i[2] for i in insts[:n])
    for (loc, sass), (src, ex, smp) in zip(table[:n], insts[:n]):
        per_line[loc] += ex; per_line_s[loc] += smp
        op = src.split()[0] if src else "?"
        if op.startswith("@"): op = src.split()[1]
        base = op.split(".")[0]
        cls = ("fp64" if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX") else "mufu/conv" if base in ("MUFU", "F2F", "I2F", "F2I", "F2FP") else
               "fp32" if base in ("FFMA", "FMUL", "FADD", "FSETP", "FMNMX", "FSEL", "FCHK") else "ld/st" if base in ("LDG", "LDS", "STS", "STG", "LDL", "STL", "LDC", "ATOMS", "ATOMG", "RED", "LDCU") else
               "branch/sync" if base in ("BRA", "BSSY", "BSYNC", "BAR", "EXIT", "CALL", "RET", "WARPSYNC", "BREAK", "BMOV") else "int/move")
        per_class[cls] += ex
    print(f"kernel {kern}: {tot:.0f} warp instructions, {tots:.0f} stall samples, {n} SASS instructions matched")
    for c, v in per_class.most_common(): print(f"  {c:12s} {v / tot * 100:5.1f} %")
    print("top source lines by executed instructions (share of instructions | share of stall samples):")
    for loc, v in per_line.most_common(top):
        print(f"  {v / tot * 100:5.2f} % | {per_line_s[loc] / max(tots, 1) * 100:5.2f} %  {loc[0]}:{loc[1]}")
    if os.environ.get("BY_STALL"):      # BY_STALL=1: the same table ordered by stall samples
        print("top source lines by stall samples (share of stall samples | share of instructions):")
        for loc, v in per_line_s.most_common(top):
            print(f"  {v / max(tots, 1) * 100:5.2f} % | {per_line[loc] / tot * 100:5.2f} %  {loc[0]}:{loc[1]}")

if __name__ == "__main__":
    main()
