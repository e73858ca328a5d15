"""Per-source-line warp-stall samples of one kernel: joins `ncu --page source --csv` (SASS rows with sampling columns)
with the line table of the cubin (nvdisasm -g).
usage: python tools/ncu_lines.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, cubin, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    short = kern.split("ILi")[0]  # template arguments are part of the mangled section name only
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + short],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hdr]
    c_s, c_n, c_i = H.index("# Samples"), H.index("Warp Stall Sampling (Not-issued Samples)"), H.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_")]
    sass = rows[hdr + 1:]
    for i, r in enumerate(sass):  # a report with several launches of the kernel: the first one
        if r and r[0] == "Kernel Name":
            sass = sass[:i]
            break
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    # the function's section
    lines, cur, infn = [], None, False
    for ln in dis.splitlines():
        if ln.startswith("\t.section") or ln.startswith(".section"):
            infn = (".text." in ln) and (kern in ln)
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    if len(lines) != len(sass):
        print("warning: %d SASS rows in the report, %d in the cubin" % (len(sass), len(lines)))
    per = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot = 0
    for r, l in zip(sass, lines):
        s = int(r[c_s] or 0)
        per[l][0] += s
        per[l][1] += int(r[c_n] or 0)
        per[l][2] += int(r[c_i] or 0)
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                per[l][3][H[c]] += v
        tot += s
    print("total samples", tot)
    for l, (s, n, i, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        why = ", ".join("%s %d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f%%  %-28s inst %10d   %s" % (100.0 * s / max(tot, 1), "%s:%d" % l if l else "?", i, why))


if __name__ == "__main__":
    main()
