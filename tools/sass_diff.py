import re, subprocess, sys, hashlib
def table(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, bodies = None, {}
    for line in out.splitlines():
        m = re.search(r"Function\s*:\s*(\S+)", line)
        if m:
            fn = m.group(1); bodies[fn] = []; continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\*", line)
        if m and fn:
            bodies[fn].append(m.group(1))
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() for k, v in bodies.items()}
a, b = table(sys.argv[1]), table(sys.argv[2])
norm = lambda n: n.replace("ILi16ELb0EE", "ILi16EE").replace("ILi8ELb0EE", "ILi8EE")
a = {norm(k): v for k, v in a.items()}; b = {norm(k): v for k, v in b.items()}
same = [k for k in a if k in b and a[k] == b[k]]
print(len(a), "functions before,", len(b), "after,", len(same), "identical")
for k in a:
    if k not in b: print("missing after:", k)
    elif a[k] != b[k]: print("CHANGED:", k)
for k in b:
    if k not in a: print("new:", k)
