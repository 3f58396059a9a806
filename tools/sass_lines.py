"""Static SASS instruction count per source line of one kernel (needs -lineinfo).
    python tools/sass_lines.py <object-or-cubin> <kernel-name-substring> [top]
Offline companion of the ncu source page: shows where a kernel's instructions come from before any GPU time is spent."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    if not path.endswith(".cubin"):
        d = tempfile.mkdtemp()
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, stdout=subprocess.DEVNULL)
        path = os.path.join(d, sorted(os.listdir(d))[0])
    txt = subprocess.run(["nvdisasm", "--print-line-info", path], capture_output=True, text=True).stdout
    counts, ops = collections.Counter(), collections.defaultdict(collections.Counter)
    infun, cur, total = False, "?", 0
    for line in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", line)
        if m:
            infun = pat in m.group(1)
            continue
        if not infun:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            counts[cur] += 1
            ops[cur][m.group(1).split(".")[0]] += 1
            total += 1
    print(f"{total} SASS instructions in functions matching {pat!r}")
    for k, v in counts.most_common(top):
        print(f"{v:6d}  {k:28s} " + " ".join(f"{o}:{c}" for o, c in ops[k].most_common(6)))


if __name__ == "__main__":
    main()
