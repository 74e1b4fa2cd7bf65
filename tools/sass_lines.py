"""Static SASS instruction counts per source line of one kernel (no GPU needed):
    python tools/sass_lines.py <file.cubin> <kernel-name-substring> [source-file-substring]
Uses `nvdisasm -g -c` (needs -lineinfo at compile time).  Prints, per source line, the number of SASS instructions
and how many of them are FP64 (DADD/DMUL/DFMA/DSETP/MUFU.*64), and the totals of each basic loop region is left to the
reader; the point is to see how many instructions one trip of a hot loop issues before spending GPU time."""
import collections
import re
import subprocess
import sys


def main():
    cubin, kern = sys.argv[1], sys.argv[2]
    srcsub = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    inside = False
    cur = ("?", 0)
    cnt = collections.Counter()
    f64 = collections.Counter()
    order = []
    total = 0
    for line in out:
        if line.startswith(".text."):
            inside = kern in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        total += 1
        if cur not in cnt:
            order.append(cur)
        cnt[cur] += 1
        if re.search(r"\b(DADD|DMUL|DFMA|DSETP|MUFU\.R(SQ|CP)64H)", ins):
            f64[cur] += 1
    print(f"{kern}: {total} SASS instructions")
    for k in sorted(cnt, key=lambda k: (k[0], k[1])):
        if srcsub and srcsub not in k[0]:
            continue
        print(f"  {k[0]}:{k[1]:<5d} {cnt[k]:4d} inst  {f64[k]:3d} fp64")


if __name__ == "__main__":
    main()
