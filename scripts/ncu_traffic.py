#!/usr/bin/env python
"""Turn an `ncu --set full` capture into the per-kernel summary bench.py reads (profiles/traffic.json) and a text
table for profiles/.  Usage (on the CPU box, after gpurun brought the .ncu-rep back):

    python scripts/ncu_traffic.py gpurun_out/r02_full_cfg2.raw.csv cfg2 f16x2 profiles/r02_ncu_cfg2.txt
(a .ncu-rep works too; the reports are exported to CSV on the GPU box because gpurun returns at most 64 MiB)

For every profiled launch it reads dram__bytes_read.sum + dram__bytes_write.sum, the duration, tensor-pipe /
issue / L1 activity and the launch geometry from `ncu -i <rep> --page raw --csv`, groups the launches by the
kernel kinds bench.py's roofline uses, and stores the per-launch MEAN of the DRAM bytes under
traffic[workload][engine][kind]."""
import csv
import io
import json
import os
import subprocess
import sys

KINDS = (("k_tc_mix_fwd", "mix_fwd"), ("k_tc_mix_bwd", "mix_bwd"), ("k_tc_edge<0>", "edge_fwd"), ("k_tc_edge<(bool)0>", "edge_fwd"),
         ("k_tc_edge<1>", "edge_bwd"), ("k_tc_edge<(bool)1>", "edge_bwd"), ("k_tc_node_post_bwd", "node_post_bwd"),
         ("k_tc_node_post", "node_post"), ("k_tc_node_pre_bwd", "node_pre_bwd"), ("k_tc_node_pre", "node_pre"),
         ("k_attn_fwd", "attn_fwd"), ("k_attn_bwd_tc", "attn_bwd"), ("k_pair_reduce", "pair_reduce"), ("k_xtg_reduce", "xtg_reduce"))
METRICS = {
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "ns",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_inst_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_wavefront_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
}


def kind_of(name):
    if "k_tc_xtg" in name:
        return "mix_dw" if ", 512," in name or ",512," in name or "(int)512" in name else "dw_small"
    for pat, kind in KINDS:
        if pat in name:
            return kind
    return None


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_ns(v, unit):
    f = float(v.replace(",", ""))
    return f * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)


def main():
    rep, workload, engine, out_txt = sys.argv[1:5]
    if rep.endswith(".csv"):        # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x.raw.csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    recs = {}
    for r in data:
        name = r[col["Kernel Name"]]
        kind = kind_of(name)
        if kind is None:
            continue
        rec = {}
        for m, key in METRICS.items():
            if m not in col:
                continue
            v, u = r[col[m]], units[col[m]]
            if not v or v == "n/a":
                continue
            rec[key] = to_bytes(v, u) if key.startswith("dram_r") or key.startswith("dram_w") else (to_ns(v, u) if key == "ns" else float(v.replace(",", "")))
        recs.setdefault(kind, []).append(rec)
    lines = [f"# ncu --set full, {workload}, engine {engine}: per-kernel means over the captured launches ({os.path.basename(rep)})",
             "# kind            n   us/launch  DRAM rd MB  DRAM wr MB  tensor%  issue%  L1 wavefront%  warps%  regs  grid x block"]
    summary = {}
    for kind, rs in recs.items():
        mean = lambda k: sum(r.get(k, 0.0) for r in rs) / len(rs)
        summary[kind] = mean("dram_read") + mean("dram_write")
        lines.append(f"{kind:15s} {len(rs):3d}  {mean('ns') / 1e3:9.1f}  {mean('dram_read') / 1e6:10.2f}  {mean('dram_write') / 1e6:10.2f}  "
                     f"{mean('tensor_pct'):7.1f} {mean('issue_pct'):7.1f}  {mean('l1_wavefront_pct'):12.1f}  {mean('warps_active_pct'):6.1f}  "
                     f"{int(mean('regs')):4d}  {int(mean('grid'))} x {int(mean('block'))}")
    open(out_txt, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    t.setdefault(workload, {})[engine] = summary
    t.setdefault("_source", {})[f"{workload}/{engine}"] = os.path.basename(out_txt)
    json.dump(t, open(tpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
