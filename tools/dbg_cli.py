import sys, os, subprocess, numpy as np, tempfile
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import mlsgpu_b200 as m
from mlsgpu_b200 import host, plyfile
from test_oracle_marching import weld_by_key
s = host.sphere_splats(60000, seed=4, radius=0.012, noise=0.002)
island = host.sphere_splats(300, seed=5, radius=0.012, noise=0.0, R=0.05)
island["position"] += np.float32([1.6, 0, 0])
both = np.concatenate([s, island])
tmp = tempfile.mkdtemp()
inp = os.path.join(tmp,"in.ply"); plyfile.write_splat_ply(inp, both, extra_property=True)
exe='/root/repo/mlsgpu_b200/host/mlsgpu'
for th in (1,3,3):
    out=os.path.join(tmp,"out%d.ply"%th)
    r=subprocess.run([exe,"-o",out,"--fit-grid","0.025","--levels","3","--fit-prune","0.02","--device-threads",str(th),"--statistics",inp],capture_output=True,text=True)
    st=dict(l.split(": ") for l in r.stdout.strip().splitlines() if ": " in l)
    print("cli threads",th,"rc",r.returncode,{k:st[k] for k in ("bins","components.total","components.kept","output.vertices","output.triangles")}, r.stderr.strip())
g, bins, lo, spacing = host.prepare(both, 0.025, max_cells=31)
for rep in range(2):
    blocks=[]
    with m.Worker(max_splats=len(g), max_cells=31) as w:
        w.set_scale_bias(float(spacing), lo.astype(np.float32)*spacing)
        for b in bins:
            w.upload_splats(np.ascontiguousarray(g[b["ids"]]))
            for blk in w.process_bin(0, len(b["ids"]), b["size"], b["offset"]):
                ni=blk["num_internal"]
                blocks.append(dict(vertices=blk["vertices"],triangles=blk["triangles"],num_internal=ni,keys=np.concatenate([np.zeros(ni,np.uint64),blk["external_keys"]])))
    pv,pt=weld_by_key(blocks)
    keep=np.linalg.norm(pv,axis=1)<1.3
    print("python bins",len(bins),"blocks",len(blocks),"welded",len(pv),len(pt),"kept",keep.sum(),keep[pt[:,0]].sum(), "expected 88466 170548 (oracle)")
