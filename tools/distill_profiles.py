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

    fused_summary()


FUSED_KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
              "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
              "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]


def fused_summary(channels=1184, block=32768):
    """ncu --set full --import-source on capture of the fused decimator (tools/round_end.sh, last line): summary,
    DRAM bytes per input sample, the SASS source page, and the same page folded into barrier-to-barrier segments."""
    import gzip
    rep = G + "/prof_fused_r1.ncu-rep"
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = {h: {"value": v, "unit": u} for h, u, v in zip(rows[0], rows[1], rows[2])}
    out = {k: d[k] for k in sorted(d) if k in FUSED_KEEP or "issue_stalled" in k and k.endswith("per_issue_active.ratio")}
    out["Kernel Name"] = d["Kernel Name"]["value"]
    json.dump(out, open(P + "/r1_fused_decim_ncu_full.json", "w"), indent=1)
    rd = float(d["dram__bytes_read.sum"]["value"]) * 1e6
    wr = float(d["dram__bytes_write.sum"]["value"]) * 1e6
    n = channels * block
    json.dump({"kernel": out["Kernel Name"], "source": "ncu --set full, profiles/r1_fused_decim_ncu_full.json (%d channels x %d samples, one launch)" % (channels, block),
               "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "input_samples": n,
               "dram_bytes_per_input_sample": round((rd + wr) / n, 3),
               "smem_wavefronts": int(float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]["value"])),
               "smem_bytes_per_input_sample": round(128.0 * float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]["value"]) / n, 2)},
              open(P + "/r1_traffic.json", "w"), indent=1)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    gzip.open(P + "/r1_fused_decim_ncu_source.csv.gz", "wt").write(src)
    rows = list(csv.reader(io.StringIO(src)))
    ix = {h: i for i, h in enumerate(rows[1])}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except ValueError:
            return 0.0
    keys = ("n_sass", "warp_inst", "fp64_inst", "lds", "sts", "smem_wavefronts", "smem_wavefronts_ideal", "samples",
            "stall_short_sb", "stall_barrier", "stall_wait", "stall_mio", "stall_math", "stall_long_sb")
    segs, cur = [], None
    for r in rows[2:]:
        if cur is None:
            cur = dict.fromkeys(keys, 0.0); cur["first_address"] = r[ix["Address"]][-5:]
        t = r[ix["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ie = f(r, "Instructions Executed")
        cur["n_sass"] += 1; cur["warp_inst"] += ie
        cur["fp64_inst"] += ie if op.startswith(("DFMA", "DMUL", "DADD")) else 0
        cur["lds"] += ie if op.startswith("LDS") else 0
        cur["sts"] += ie if op.startswith("STS") else 0
        cur["smem_wavefronts"] += f(r, "L1 Wavefronts Shared"); cur["smem_wavefronts_ideal"] += f(r, "L1 Wavefronts Shared Ideal")
        cur["samples"] += f(r, "# Samples")
        for k in keys[8:]:
            cur[k] += f(r, k)
        if op.startswith(("BAR", "EXIT")):
            cur["ends_with"] = " ".join(t)[:48]; segs.append(cur); cur = None
    tot = {k: sum(s[k] for s in segs) for k in keys}
    tab = []
    for s in segs:
        if s["warp_inst"] < 0.003 * tot["warp_inst"] and s["samples"] < 0.003 * tot["samples"]:
            continue
        e = {"first_address": s["first_address"], "ends_with": s["ends_with"], "n_sass": int(s["n_sass"])}
        for k in ("warp_inst", "fp64_inst", "smem_wavefronts", "samples"):
            e["pct_" + k] = round(100 * s[k] / tot[k], 1)
        for k in keys[8:]:
            e["pct_samples_" + k] = round(100 * s[k] / tot["samples"], 1)
        e["lds"], e["sts"] = int(s["lds"]), int(s["sts"])
        tab.append(e)
    json.dump({"note": "SASS source page of the fused decimator folded into barrier-to-barrier segments (program order); percentages of the kernel totals",
               "totals": tot, "segments": tab}, open(P + "/r1_fused_decim_ncu_segments.json", "w"), indent=1)
    print("fused:", out["Kernel Name"][:60], d["gpu__time_duration.sum"]["value"], "us; fp64 pipe",
          d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]["value"], "smem wavefronts",
          d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]["value"], "B/sample", round((rd + wr) / n, 3))


if __name__ == "__main__":
    main()
