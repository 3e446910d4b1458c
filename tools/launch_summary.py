"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals and shares."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
body = rows[1:]
if "--after-last" in sys.argv:  # keep the launches from the last occurrence of a kernel on (one bench step)
    pat = sys.argv[sys.argv.index("--after-last") + 1]
    last = max((i for i, r in enumerate(body) if pat in r[ik]), default=0)
    body = body[last:]
    print(f"# launches from the last '{pat}' on")
for r in body:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    u = r[iu]
    us = v / 1e3 if u.startswith("ns") else v * 1e3 if u.startswith("ms") else v
    name = re.sub(r"\(.*", "", r[ik])[:70]
    tot[name] += us; cnt[name] += 1
all_us = sum(tot.values())
print(f"# {sum(cnt.values())} launches, {all_us/1e3:.1f} ms of kernel time (serialised, cold caches under ncu)")
for name, us in tot.most_common(30):
    print(f"{100*us/all_us:6.2f} %  {us/1e3:9.2f} ms  {cnt[name]:6d} x  {name}")
