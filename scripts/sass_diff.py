import subprocess, sys, hashlib, re
def funcs(path):
    out = subprocess.run(['cuobjdump','-sass',path],capture_output=True,text=True).stdout
    d={}; cur=None
    for ln in out.splitlines():
        m=re.match(r'\s*Function : (\S+)',ln)
        if m: cur=m.group(1); d[cur]=hashlib.md5(); continue
        if cur and re.match(r'\s*/\*[0-9a-f]{4}\*/',ln): d[cur].update(re.sub(r'/\*[0-9a-f]{4}\*/','',ln,count=1).encode())
    return {k:v.hexdigest() for k,v in d.items()}
a,b=funcs(sys.argv[1]),funcs(sys.argv[2])
print(len(a),len(b))
for k in sorted(set(a)|set(b)):
    if a.get(k)!=b.get(k): print('DIFF',k, 'only-in-new' if k not in a else 'only-in-old' if k not in b else 'changed')
