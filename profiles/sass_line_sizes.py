"""Static code size per CUDA source line of one kernel:  python profiles/sass_line_sizes.py <file.o|.so> <kernel-substring> [top_n]
(cuobjdump -xelf + nvdisasm -g; needs -lineinfo).  Used to keep hot kernels inside the 32 KB instruction cache."""
import os, re, subprocess, sys, tempfile
from collections import Counter
obj, want = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
    cubins = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")]
    txt = "".join(subprocess.run(["nvdisasm", "-g", c], capture_output=True, text=True).stdout for c in cubins)
cur, sec, cnt, inl = None, None, Counter(), Counter()
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        sec = m.group(1); continue
    if sec is None or want not in sec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+[A-Z@!]", line) and cur:
        cnt[cur] += 1
tot = sum(cnt.values())
print(f"{want}: {tot} instructions = {tot * 16 / 1024:.1f} KB")
for k, v in cnt.most_common(top):
    print(f"  {v:5d}  {k[0]}:{k[1]}")
