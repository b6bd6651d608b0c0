"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep captured with --import-source on.
python tools/ncu_lines.py rep [kernel-substring] [top]"""
import csv, io, subprocess, sys, collections
SORT = 1 if "--by-samples" in sys.argv else 0
sys.argv = [a for a in sys.argv if a != "--by-samples"]
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path":
        cur = {"file": row[1], "rows": [], "hdr": None, "func": ""}; blocks.append(cur)
    elif row[0] == "Function Name": cur["func"] = row[1]
    elif row[0] == "Line No": cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None: cur["rows"].append(row)
def num(x):
    try: return int(x)
    except ValueError: return 0
agg = collections.defaultdict(lambda: [0, 0, ""])
tot_i = tot_s = 0
for b in blocks:
    if want not in b["func"]: continue
    h = b["hdr"]; ii = h.index("Instructions Executed"); si = h.index("# Samples")
    for r in b["rows"]:
        if r[0] == "": continue              # SASS rows are attributed to the preceding source row already
        key = (b["file"].split("/")[-1], r[0], b["func"][-40:])
        agg[key][0] += num(r[ii]); agg[key][1] += num(r[si]); agg[key][2] = r[1].strip()[:110]
        tot_i += num(r[ii]); tot_s += num(r[si])
print("total instr %d samples %d" % (tot_i, tot_s))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
    print("%-14s %5s inst %6.2f%% smp %6.2f%%  %s" % (k[0], k[1], 100.0 * v[0] / max(tot_i, 1), 100.0 * v[1] / max(tot_s, 1), v[2]))
