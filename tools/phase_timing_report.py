"""Turn the output of tools/phase_timing.py (gpurun_out/phase_timing_c2.jsonl, ..._c3.jsonl) into the tables of
profiles/r1_phase_timing.md (the prose of that file is kept; only the two tables are regenerated).

    python tools/phase_timing_report.py
"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(p):
    return [json.loads(l) for l in open(p) if l.startswith("{")]


def table_c2(recs):
    names = list(recs[0]["warp0"].keys())[:9]
    out = ["| utterances | per SM | ms (instrumented) | cycles/frame | fastest .. slowest utterance | " + " | ".join(names) + " |",
           "|---|---|---|---|---|" + "---|" * 9]
    for d in recs:
        for w in (0, 1):
            row = d["warp%d" % w]
            out.append(f"| {d['B']} (warp {w}) | {d['B'] / 148:.0f} | {d['ms']:.3f} | {d['cycles_per_frame_total'][w]:.0f} | "
                       f"{d['utterance_total_min_max'][0]:.0f} .. {d['utterance_total_min_max'][1]:.0f} | "
                       + " | ".join(f"{row[n]:.0f}" for n in names) + " |")
    return "\n".join(out)


def table_c3(recs):
    names = list(recs[0]["warp0"].keys())[:9]
    out = ["| utterances | warp | cycles/frame | " + " | ".join(names) + " |", "|---|---|---|" + "---|" * 9]
    for d in recs:
        for w in (0, 3, 4, 7):
            row = d["warp%d" % w]
            out.append(f"| {d['B']} ({d['B'] / 148:.0f}/SM) | {w} | {d['cycles_per_frame_total'][w]:.0f} | "
                       + " | ".join(f"{row[n]:.0f}" for n in names) + " |")
    return "\n".join(out)


def main():
    path = os.path.join(ROOT, "profiles", "r1_phase_timing.md")
    text = open(path).read()
    tabs = [table_c2(load(os.path.join(ROOT, "gpurun_out", "phase_timing_c2.jsonl"))),
            table_c3(load(os.path.join(ROOT, "gpurun_out", "phase_timing_c3.jsonl")))]
    # a table = a run of consecutive lines starting with '|'
    blocks = list(re.finditer(r"(?:^\|.*\n)+", text, flags=re.M))
    assert len(blocks) == 2, len(blocks)
    for m, t in reversed(list(zip(blocks, tabs))):
        text = text[:m.start()] + t + "\n" + text[m.end():]
    open(path, "w").write(text)
    print("updated", path)


if __name__ == "__main__":
    main()
