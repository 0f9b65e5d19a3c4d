"""Instruction breakdown of one kernel by source line / function region from an ncu report with --import-source on:
    python profiles/ncu_source_breakdown.py rep.ncu-rep <kernel regex> [launch index]"""
import csv, subprocess, sys, collections, io, re
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
# the output holds one table per (launch, file); take the first launch only
tables, cur, fname, func = [], None, None, None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1]
    elif row[0] == "Function Name":
        func = row[1]
    elif row[0] == "Line No":
        head = row
        cur = []
        tables.append((fname, func, head, cur))
    elif cur is not None:
        cur.append(row)
seen = set()
per_line = collections.Counter()
stall = collections.Counter()
src_text = {}
for fname, func, head, rows in tables:
    key = fname
    if key in seen:      # second launch of the same kernel
        continue
    seen.add(key)
    ci, cs, ct = head.index("Instructions Executed"), head.index("# Samples"), head.index("Source")
    for r in rows:
        if r[0] in ("", "-"):
            continue
        try:
            k = (fname.split("/")[-1], int(r[0]))
            per_line[k] += int(r[ci] or 0)
            stall[k] += int(r[cs] or 0)
            src_text[k] = r[ct].strip()[:90]
        except ValueError:      # a source line with unescaped quotes (inline asm in a CUDA header)
            continue
total = sum(per_line.values())
print("total warp instructions (first launch):", total, " stall samples:", sum(stall.values()))
for k, v in per_line.most_common(45):
    print("%6.2f%% inst %6.2f%% samples  %s:%d  %s" % (100.0 * v / total, 100.0 * stall[k] / max(1, sum(stall.values())), k[0], k[1], src_text[k]))
