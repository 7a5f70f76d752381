"""Join an `ncu --page source --csv` export (per-SASS-instruction counters) with `nvdisasm -g` line info of the same
cubin and print executed warp-instructions / stall samples per source line, per file and per opcode class.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libaec_b200.so; nvdisasm -g <cubin> > all.sass   (cut to the kernel's .text section)
    python tools/ncu_by_line.py src.csv kernel.sass [frames_per_launch]
"""
import collections
import csv
import re
import sys


def load_sass(path):
    """address -> (file, line, text) from nvdisasm -g output (inline chains: innermost location)."""
    out, cur = {}, ("?", 0)
    pat_loc = re.compile(r'//## File "([^"]+)", line (\d+)')
    pat_ins = re.compile(r'/\*([0-9a-f]{4,})\*/\s+(.*?);')
    for ln in open(path):
        m = pat_loc.search(ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = pat_ins.search(ln)
        if m:
            out[int(m.group(1), 16)] = (cur[0], cur[1], m.group(2).strip())
    return out


def main():
    src_csv, sass_path = sys.argv[1], sys.argv[2]
    frames = float(sys.argv[3]) if len(sys.argv) > 3 else None
    sass = load_sass(sass_path)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def val(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    per_line = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    mism = 0
    for r in rows[2:]:
        try:
            addr = int(r[ix["Address"]], 16) if not r[ix["Address"]].isdigit() else int(r[ix["Address"]])
        except ValueError:
            continue
        base = min(sass) if sass else 0
        key = sass.get(addr) or sass.get(addr - (addr - base) // 1 * 0)
        if key is None:
            # ncu addresses are absolute; map by order instead
            key = None
        per_line[(addr,)]["_"] = 0
    # map by order (ncu addresses are absolute virtual addresses)
    addrs = sorted(sass)
    data = rows[2:]
    if len(addrs) != len(data):
        print(f"warning: {len(addrs)} SASS instructions vs {len(data)} ncu rows", file=sys.stderr)
    per_line.clear()
    ffma3 = collections.Counter()
    for a, r in zip(addrs, data):
        f, ln, text = sass[a]
        ncu_text = r[ix["Source"]].strip()
        if text.split()[0].split(".")[0] not in ncu_text:
            mism += 1
        c = per_line[(f, ln)]
        n = val(r, "Instructions Executed")
        c["inst"] += n
        c["samples"] += val(r, "# Samples")
        for k in stall_keys:
            c[k] += val(r, k)
        op = text.split()[1] if text.startswith("@") else text.split()[0]
        opb = op.split(".")[0]
        tot["inst"] += n
        tot["samples"] += val(r, "# Samples")
        if opb == "FFMA":
            regs = re.findall(r'(?<![A-Za-z])-?\|?(R\d+)', text.split(",", 1)[1]) if "," in text else []
            distinct = len(set(regs))
            reuse = text.count(".reuse")
            ffma3[(distinct, reuse)] += n
    if mism:
        print(f"warning: {mism} opcode mismatches between the two listings", file=sys.stderr)
    scale = 1.0 / frames if frames else 1.0
    unit = "per frame" if frames else "total"
    print(f"executed warp-instructions {unit}: {tot['inst'] * scale:.1f}; stall samples {tot['samples']:.0f}")
    print("\nFFMA by (distinct source registers, .reuse flags): warp-instructions " + unit)
    for k, v in sorted(ffma3.items()):
        print(f"  {k}: {v * scale:9.1f}")
    print(f"\n{'file:line':34s} {'inst':>9s} {'inst%':>6s} {'smp%':>6s}  top stalls")
    for (f, ln), c in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:70]:
        tops = sorted(((c[k], k) for k in stall_keys), reverse=True)[:3]
        ts = " ".join(f"{k[6:]}={100 * v / max(tot['samples'], 1):.1f}" for v, k in tops if v > 0)
        print(f"{f + ':' + str(ln):34s} {c['inst'] * scale:9.1f} {100 * c['inst'] / tot['inst']:6.2f} "
              f"{100 * c['samples'] / tot['samples']:6.2f}  {ts}")


if __name__ == "__main__":
    main()
