"""Print registers / spills / stack per kernel from the ptxas logs of the last build."""
import glob, os, re, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "acoustic_echo_cancellation_b200", "csrc", "build")
rows = []
for f in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '([^']+)'[^\n]*\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", txt):
        name = m.group(1)
        t = re.search(r"stage1_n(\d+)_kernelILi(\d+)ELi(\d+)ELi(\d+)ELb(\d)ELi(\d+)E", name)
        label = f"N{t.group(1)} NW{t.group(2)} P{t.group(3)} algo{t.group(4)} echo{t.group(5)} cap{t.group(6)}" if t else name[:60]
        rows.append((label, int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
print(f"{'kernel':44s} regs stack spill_st spill_ld")
for r in rows:
    print(f"{r[0]:44s} {r[1]:4d} {r[2]:5d} {r[3]:8d} {r[4]:8d}")
