"""Static code-size profile of one kernel: SASS instructions per source line / function range, from `nvdisasm -gi`.
usage: python scripts/sass_static_profile.py <object.o> <kernel-name-substring> [top_n]
(the megakernel is instruction-cache sensitive on C4: this shows which inlined copies make it big)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
inside, cur, fresh = False, None, True
cnt, site = collections.Counter(), collections.Counter()
for ln in txt.splitlines():
    if ln.startswith(".text."):
        inside = kernel in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        if fresh:  # the chain of an inlined instruction is printed innermost first: keep the first entry
            cur, fresh = (os.path.basename(m.group(1)), int(m.group(2))), False
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln) and cur:
        cnt[cur] += 1
        fresh = True
total = sum(cnt.values())
byfile = collections.Counter()
for (f, _), v in cnt.items():
    byfile[f] += v
print(f"{kernel}: {total} instructions = {total * 16 / 1024:.1f} KB")
for f, v in byfile.most_common():
    print(f"  {f:28s} {v:6d}")
print("top lines:")
for (f, l), v in cnt.most_common(top):
    print(f"  {f}:{l}  {v}")
