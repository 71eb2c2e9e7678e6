"""dram read + write bytes per launch of the dominant tcgen05 GEMM shapes, from an `ncu --set full` capture of tools/prof_all.py
(run here, no GPU):   python tools/ncu_traffic.py gpurun_out/r02_prof_all.ncu-rep > profiles/r02_ncu_traffic.json
bench.py attaches these to `roofline.by_shape[].traffic` (load_ncu_traffic)."""
import csv, io, json, subprocess, sys, datetime

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
iname, ird, iwr, idur = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
units = rows[1]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


# launch order of the tcgen05 GEMMs in tools/prof_all.py (first repetition): key = (M, N, K, batch, relu & 2)
order = [((256, 16640, 64, 64, 0), "sim3 (fp32 out)"), ((16640, 256, 64, 64, 2), "sim1 + fused row softmax -> bf16"),
         ((133120, 1536, 512, 1, 0), "QKV projection (fp32 + bf16 out)"), ((256, 256, 256, 64, 0), "pinv step (bf16x3)")]
gemms = [r for r in rows[2:] if "gemm_tc_kernel" in r[iname]]
shapes = []
for (key, what), r in zip(order, gemms):
    shapes.append({"key": list(key), "what": what, "bytes": to_bytes(r[ird], units[ird]) + to_bytes(r[iwr], units[iwr]),
                   "dram_read": to_bytes(r[ird], units[ird]), "dram_write": to_bytes(r[iwr], units[iwr]), "ncu_duration": r[idur] + " " + units[idur]})
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
print(json.dumps({"source": rep, "commit": commit, "captured": datetime.date.today().isoformat(), "shapes": shapes}, indent=1))
