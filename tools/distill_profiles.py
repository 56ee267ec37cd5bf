"""Turn the raw outputs of tools/round_end.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections
import csv
import io
import json
import os
import subprocess

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")


def last_json(f):
    return json.loads(open(f).read().strip().splitlines()[-1])


def main():
    m = {"bench_r1_n1": "r1_bench_n1", "bench_r1_ref": "r1_bench_reference", "bench_r1_channelizer": "r1_bench_channelizer_c5",
         "bench_r1_pan16": "r1_bench_panadapter_c2", "bench_r1_rxa_usb": "r1_bench_rxa_usb_c3", "bench_r1_rxa_fm": "r1_bench_rxa_fm_c4",
         "bench_r1_n1_nco_closed": "r1_bench_n1_nco_closed"}
    for a, b in m.items():
        d = last_json("%s/%s.json" % (G, a))
        json.dump(d, open("%s/%s.json" % (P, b), "w"), indent=1)
        print(b, round(d["value"], 1), d.get("roofline") and round(d["roofline"]["frac"], 3), d.get("e2e") and round(d["e2e"]["value"], 1))
    for a, b in {"launches_r1": "r1_launches_bench_default", "launches_r1_channelizer": "r1_launches_channelizer",
                 "launches_r1_panadapter": "r1_launches_panadapter"}.items():
        txt = [l for l in open("%s/%s.csv" % (G, a)) if not l.startswith("==")]
        open("%s/%s.csv" % (P, b), "w").writelines(txt)
        rows = list(csv.DictReader(txt))
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows:
            if r["Metric Name"] != "gpu__time_duration.sum":
                continue
            agg[r["Kernel Name"][:48]][0] += 1
            agg[r["Kernel Name"][:48]][1] += float(r["Metric Value"]) / 1e3
        tot = sum(v[1] for v in agg.values()) or 1.0
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            print("   %-50s n=%3d  %8.1f us each  %5.1f%%" % (k, v[0], v[1] / v[0], 100 * v[1] / tot))
    keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct", "smsp__average_warps_issue_stalled", "sm__pipe_fp64_cycles_active.avg.pct", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct",
            "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block", "sm__throughput.avg.pct",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]

    def summarize(rep, out, note):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr = rows[0]
        res = []
        for vals in rows[2:]:
            d = {"kernel": vals[4]}
            for h, v in zip(hdr, vals):
                if any(h.startswith(k) for k in keep) and v not in ("", "nan"):
                    try:
                        d[h] = float(v.replace(",", ""))
                    except ValueError:
                        d[h] = v
            res.append(d)
        json.dump({"note": note, "kernels": res}, open(out, "w"), indent=1)
        print(out, len(res))
    sections = "ncu --section SpeedOfLight/WarpStateStats/MemoryWorkloadAnalysis/Occupancy/SchedulerStats/LaunchStats --clock-control none"
    summarize(G + "/prof_pfb_r1.ncu-rep", P + "/r1_channelizer_ncu_sections.json", sections + ", tools/pfb_once.py (16 Mi samples, 1024 receivers, D = 512)")
    summarize(G + "/prof_pan_r1.ncu-rep", P + "/r1_panadapter_ncu_sections.json", sections + ", bench.py --workload panadapter --channels 16 --block 1048576")


if __name__ == "__main__":
    main()
