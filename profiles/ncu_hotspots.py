"""Warp-stall samples per CUDA source line from an ncu report (compiled with -lineinfo):
    ncu -i <rep> --page source --csv --print-source sass,cuda | python profiles/ncu_hotspots.py [top]
Groups the per-SASS-instruction rows by the CUDA line they belong to; excludes the samples of warps parked at a barrier
separately (a 1024-thread CTA waiting for one busy warp shows up as barrier samples of the 31 idle ones)."""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
files, cur_file, hdr = {}, None, None
agg = {}
tot_s = tot_i = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {name: k for k, name in enumerate(r) if name not in ("Source",)}
        src_cols = [k for k, name in enumerate(r) if name == "Source"]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[hdr["Line No"]])
        samp = int(r[hdr["# Samples"]] or 0)
        inst = int(r[hdr["Instructions Executed"]] or 0)
        bar = int(r[hdr["stall_barrier"]] or 0)
    except (ValueError, KeyError):
        continue
    key = (cur_file, line)
    a = agg.setdefault(key, [0, 0, 0, r[src_cols[0]] if src_cols else "", {}])
    a[0] += samp; a[1] += inst; a[2] += bar
    for name in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_barrier", "stall_branch_resolving", "stall_no_inst"):
        if name in hdr:
            a[4][name] = a[4].get(name, 0) + int(r[hdr[name]] or 0)
    tot_s += samp; tot_i += inst
print("total instructions %.3e, samples %d" % (tot_i, tot_s))
for (f, line), (samp, inst, bar, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    lead = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%-20s %5d  samp %5.1f%%  inst %5.1f%%  %-34s %s" % (f, line, 100.0 * samp / max(tot_s, 1), 100.0 * inst / max(tot_i, 1),
                                                          " ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(samp, 1)) for k, v in lead), src.strip()[:110]))
