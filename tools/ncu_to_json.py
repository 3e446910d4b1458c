"""ncu raw csv of one kernel -> profiles/ncu_dominant_kernel_r2.json (dram bytes per launch, hit rates)."""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
def val(r, name):
    v = float(r[idx[name]]); u = units[idx[name]].lower()
    return v * (1e9 if u.startswith("gbyte") else 1e6 if u.startswith("mbyte") else 1e3 if u.startswith("kbyte") else 1.0)
best = None
for r in rows[2:]:
    if "spmm_strip_kernel" in r[idx["Kernel Name"]] or "spmm_kernel" in r[idx["Kernel Name"]]:
        best = r
out = {"kernel": best[idx["Kernel Name"]], "columns": int(sys.argv[2]), "dtype": "f32" if "<float" in best[idx["Kernel Name"]] else "f64",
       "dram_bytes_per_launch": val(best, "dram__bytes_read.sum") + val(best, "dram__bytes_write.sum"),
       "dram_read_bytes": val(best, "dram__bytes_read.sum"), "dram_write_bytes": val(best, "dram__bytes_write.sum"),
       "duration_us_under_ncu": float(best[idx["gpu__time_duration.sum"]]),
       "l1_hit_pct": float(best[idx["l1tex__t_sector_hit_rate.pct"]]), "l2_hit_pct": float(best[idx["lts__t_sector_hit_rate.pct"]]),
       "source": "ncu --set full --clock-control none, last SpMM launch of tools/spmm_once.py (level-9 stiffness, solver numbering)"}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(out)
