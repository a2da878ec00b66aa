"""Per-source-line stall samples from an .ncu-rep (needs -lineinfo + --import-source on):
    python scripts/ncu_lines.py REP [kernel-substring] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, cur_fn, per = None, None, {}
tot = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name":
        cur_fn = r[1]
    elif len(r) > 7 and r[0].isdigit() and r[2] == "-":
        if pat and pat not in (cur_fn or ""):
            continue
        try:
            s = int(r[4]); ins = int(r[7])
        except ValueError:
            continue
        key = (cur_fn.split("(")[0][-60:], cur_file, int(r[0]), r[1].strip()[:110])
        a = per.setdefault(key, [0, 0])
        a[0] += s; a[1] += ins
        tot[key[0]] = tot.get(key[0], 0) + s
for fn, t in tot.items():
    print("== %s: %d samples" % (fn, t))
    items = sorted((v[0], k, v[1]) for k, v in per.items() if k[0] == fn)[::-1][:top]
    for s, k, ins in items:
        print("%6.2f%%  %8d inst  %s:%d  %s" % (100.0 * s / max(t, 1), ins, k[1], k[2], k[3]))
