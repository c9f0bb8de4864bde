"""Per-kernel SASS comparison of two builds of libavs_b200.so (md5 of each kernel's instruction stream, addresses stripped).
    python scripts/sass_diff.py old.so new.so          -> lists kernels that differ / exist on one side only
    python scripts/sass_diff.py --record file.json     -> records the hashes of the in-tree library (tests/test_sass.py compares to them)
"""
import subprocess, sys, hashlib, re
def funcs(path):
    out = subprocess.run(['cuobjdump','-sass',path],capture_output=True,text=True).stdout
    d={}; cur=None
    for ln in out.splitlines():
        m=re.match(r'\s*Function : (\S+)',ln)
        if m: cur=m.group(1); d[cur]=hashlib.md5(); continue
        if cur and re.match(r'\s*/\*[0-9a-f]{4}\*/',ln): d[cur].update(re.sub(r'/\*[0-9a-f]{4}\*/','',ln,count=1).encode())
    return {k:v.hexdigest() for k,v in d.items()}
if sys.argv[1] == "--record":
    import json
    from pathlib import Path
    lib = Path(__file__).resolve().parent.parent / "adaptiveviscositysolver_b200" / "libavs_b200.so"
    k = funcs(str(lib))
    json.dump({"note": "md5 of the SASS instruction stream (addresses stripped) of every kernel of libavs_b200.so; recorded by scripts/sass_diff.py --record",
               "kernels": k}, open(sys.argv[2], "w"), indent=0, sort_keys=True)
    print(len(k), "kernels recorded")
    sys.exit(0)
a,b=funcs(sys.argv[1]),funcs(sys.argv[2])
print(len(a),len(b))
for k in sorted(set(a)|set(b)):
    if a.get(k)!=b.get(k): print('DIFF',k, 'only-in-new' if k not in a else 'only-in-old' if k not in b else 'changed')
