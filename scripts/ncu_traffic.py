"""Developer tool: per-class DRAM traffic from an ncu launch list taken with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --csv
Writes profiles/traffic.json in the layout bench.py reads (per_launch_bytes per kernel class of flip_get_kernel_timing).
Usage: ncu_traffic.py <launches.csv> <workload> <out.json>"""
import csv, json, sys
from collections import defaultdict

CLASS_OF = {
    "sdf_p2g": ["k_occ_bits", "k_occ_dilate", "k_tile_lists", "k_p2g_scatter", "k_sdf_shell", "k_p2g_finish", "k_p2g_literal", "k_sdf_far"],
    "g2p_advance": ["k_g2p_advance_fused", "k_speed_hist_list"],
    "sort": ["k_reset_sort_scalars", "k_speed_limit", "k_classify", "k_scatter_idx", "k_cell_finalize", "k_build_src", "k_gather", "DeviceScan"],
    "extrapolate": ["k_ext_init", "k_ext_layer"],
    "pcg_iter": ["k_pcg_update_presweep", "k_mg0_restrict_list", "k_mg_coarse", "k_mg0_sweep", "k_pcg_dir_spmv"],
    "pressure_apply": ["k_apply_pressure"],
}
# what counts one launch of the class
UNIT_OF = {"sdf_p2g": "k_p2g_scatter", "g2p_advance": "k_g2p_advance_fused", "sort": "k_gather", "extrapolate": "k_ext_init",
           "pcg_iter": "k_pcg_dir_spmv", "pressure_apply": "k_apply_pressure<0>"}

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = defaultdict(dict)
name_of = {}
for r in rows[1:]:
    if len(r) <= mv:
        continue
    per[r[idc]][r[mn]] = float(r[mv].replace(",", ""))
    name_of[r[idc]] = r[kn]
unit_col = hdr.index("Metric Unit")
units = {}
for r in rows[1:]:
    if len(r) > mv:
        units[r[mn]] = r[unit_col]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
bytes_of = defaultdict(float)
count_of = defaultdict(int)
time_of = defaultdict(float)
for i, m in per.items():
    name = name_of[i]
    b = (m.get("dram__bytes_read.sum", 0.0) * scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
         + m.get("dram__bytes_write.sum", 0.0) * scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0))
    for cls, keys in CLASS_OF.items():
        if any(k in name for k in keys):
            # the sort also runs once per host upload; scans of the pressure set-up share the cub kernel name: both small
            bytes_of[cls] += b
            time_of[cls] += m.get("gpu__time_duration.sum", 0.0)
        if UNIT_OF[cls] in name:
            count_of[cls] += 1
out = {"workload": sys.argv[2],
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none over one run of "
                 "`bench.py --steps 1 --warmup 1` (" + sys.argv[1].split("/")[-1] + "): class total / class launches",
       "per_launch_bytes": {c: int(bytes_of[c] / max(count_of[c], 1)) for c in CLASS_OF},
       "launches_seen": dict(count_of)}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out, indent=1))
