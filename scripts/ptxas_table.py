"""Per-function register / stack / spill table of a build, and the difference between two builds.

  nvcc ... -Xptxas -v -o a.so optimalmatrixcompletion.jl_b200/csrc/omc_api.cu 2> a.log
  python scripts/ptxas_table.py a.log            # table
  python scripts/ptxas_table.py a.log b.log      # functions whose numbers differ

Used in round 1 to keep optional code out of the default kernel: ptxas' allocation of `relax_project_block` moved from
248/556 to 364/1128 spill bytes when unrelated structs grew by 8 bytes, so the infeasibility certificate went behind a
compile-time switch and the default build was checked to be instruction-identical (cuobjdump -sass) to the measured one."""
import re
import sys


def parse(path):
    out, cur = {}, None
    for line in open(path):
        m = re.search(r"Function properties for (\S+)", line) or re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur:
            out.setdefault(cur, {})["stack/spill_st/spill_ld"] = tuple(map(int, m.groups()))
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            out.setdefault(cur, {})["regs"] = int(m.group(1))
    return out


if __name__ == "__main__":
    a = parse(sys.argv[1])
    if len(sys.argv) == 2:
        for k in sorted(a):
            print(k[:100], a[k])
    else:
        b = parse(sys.argv[2])
        same = True
        for k in sorted(set(a) | set(b)):
            if a.get(k) != b.get(k):
                same = False
                print(k[:100], a.get(k), "->", b.get(k))
        print("identical" if same else "different")
