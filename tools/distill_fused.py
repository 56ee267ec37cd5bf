"""Distil one `ncu --set full --import-source on` capture of the fused decimator into profiles/:
summary (<tag>_fused_decim_ncu_full.json), the SASS source page (<tag>_fused_decim_ncu_source.csv.gz) and
profiles/traffic.json (what bench.py quotes as roofline.traffic / roofline.smem when the kernel name matches)."""
import csv, gzip, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from distill_profiles import FUSED_KEEP, G, P

rep, tag = sys.argv[1], sys.argv[2]
channels, block = int(sys.argv[3]) if len(sys.argv) > 3 else 1184, 32768
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = {h: {"value": v, "unit": u} for h, u, v in zip(rows[0], rows[1], rows[2])}
out = {k: d[k] for k in sorted(d) if k in FUSED_KEEP or "issue_stalled" in k and k.endswith("per_issue_active.ratio")}
out["Kernel Name"] = d["Kernel Name"]["value"]
full = "%s/%s_fused_decim_ncu_full.json" % (P, tag)
json.dump(out, open(full, "w"), indent=1)


def val(k):
    v, u = float(d[k]["value"].replace(",", "")), d[k]["unit"]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
n = channels * block
wf = val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
json.dump({"kernel": out["Kernel Name"], "capture": os.path.relpath(full, os.path.dirname(P)),
           "source": "ncu --set full --clock-control none --import-source on, bench.py --workload rx_chain --channels %d (one launch, %d channels x %d samples)" % (channels, channels, block),
           "dram_bytes_read": rd, "dram_bytes_write": wr, "input_samples": n, "dram_bytes_per_input_sample": (rd + wr) / n,
           "smem_wavefronts": int(wf), "smem_bytes_per_input_sample": 128.0 * wf / n}, open(P + "/traffic.json", "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
gzip.open("%s/%s_fused_decim_ncu_source.csv.gz" % (P, tag), "wt").write(src)
for k in sorted(out):
    if k != "Kernel Name":
        print("%-90s %s %s" % (k, out[k]["value"], out[k]["unit"]))
print(open(P + "/traffic.json").read())
