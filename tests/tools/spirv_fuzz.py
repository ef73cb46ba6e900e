#!/usr/bin/env python
"""Random-pose fuzz of the oracle against the reference's compiled shaders (needs /root/reference): for N random cameras,
lights, aspect ratios and seeds, 24 random pixels each are run through Raytracer.comp.spv and Tracer.comp.spv
(oracle/spirv_interp.py; Tracer's rand() answered by the integer RNG) and through the oracle.

    python tests/tools/spirv_fuzz.py [N=64]      # ~10 s per 64 poses on 8 cores

Expected (measured with N = 640, profiles/r01_spirv_fuzz.txt): every primary-hit id identical, every whitted pixel within
one 8-bit step, and a path pixel in ~10,000 off by more than that -- a last-bit difference between two legal executions
flipping one Russian-roulette or visibility decision of one of its four samples.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, _p)
W,H=40,30
def pose(k):
    rs=np.random.RandomState(1000+k)
    pos=np.array([rs.uniform(-55,55), rs.uniform(5,120), rs.uniform(-55,55)])
    d=rs.normal(size=3); d/=np.linalg.norm(d)
    r=np.cross(d,[0,1,0]); 
    if np.linalg.norm(r)<1e-3: r=np.array([1.0,0,0])
    r/=np.linalg.norm(r); u=np.cross(r,d)
    light=np.array([rs.uniform(-50,50), rs.uniform(10,110), rs.uniform(-50,50)])
    return [pos.astype(np.float32), d.astype(np.float32), r.astype(np.float32), u.astype(np.float32)], light.astype(np.float32), float(rs.uniform(0.5,2.0)), float(rs.uniform(0,1))
def mkfd(k):
    import vk_renderer_b200.device as D
    cam,light,aspect,seed=pose(k)
    fd=D.default_frame_data(aspect_ratio=aspect, seed=seed)
    for name,v in zip(("pos","dir","right","up"),cam):
        a=getattr(fd.camera,name); a.x,a.y,a.z=float(v[0]),float(v[1]),float(v[2])
    fd.light_pos.x,fd.light_pos.y,fd.light_pos.z=[float(x) for x in light]
    return fd
def work(k):
    import spirv_interp as S, make_spirv_vectors as G, oracle as O
    fd=mkfd(k)
    rs=np.random.RandomState(k)
    px=[(int(rs.randint(0,W)),int(rs.randint(0,H))) for _ in range(24)]
    res={'k':k}
    # whitted
    m=S.Module(os.path.join(G.SPV,'Raytracer.comp.spv')); mc=S.Machine(m); G.set_fd(S,mc,fd)
    tex=S.run_compute(mc,W,H,px)
    acc,ids,rgba,_=O.Scene().use_default(O.SCENE_RAYTRACER).render(fd,W,H,spp=1,max_depth=2,integrator=O.WHITTED)
    bad=0; mx=0
    for (x,y),t in tex.items():
        u8=[S.unorm8(c) for c in t[:3]]
        d=max(abs(u8[i]-int(rgba[y,x,i])) for i in range(3)); mx=max(mx,max(abs(t[i]-acc[y,x,i]) for i in range(3)))
        if d>1: bad+=1
    res['w_bad']=bad; res['w_max']=mx
    # path
    m=S.Module(os.path.join(G.SPV,'Tracer.comp.spv'))
    L=O.lib(); fkey=L.orc_frame_key(5, fd.seed, k)
    rng=S.TracerRng(m, lambda pixel,sample,dim: L.orc_rand_u01(fkey,pixel,sample,dim))
    mc=S.Machine(m,hooks=rng.hooks()); G.set_fd(S,mc,fd)
    img=S.Image(W,H); mc.global_by_binding(0)[0]=img
    mc.global_by_binding(1)[0]=[[[[S.f32(float(c)) for c in v] for v in t] for t in G.HOST_TRIANGLE]]
    gid=mc.global_by_builtin(28); tr=m.function('trace_ray(')
    acc,ids,rgba,_=O.Scene().use_default(O.SCENE_TRACER).render(fd,W,H,spp=4,max_depth=4,integrator=O.PATH,seed=5,frame_index=k)
    pbad=0; idbad=0; relmax=0; details=[]
    for (x,y) in px:
        rng.begin_pixel(y*W+x); gid[0]=[x,y,0]; mc.run(m.entry)
        t=img.texels[(x,y)]
        u8=[S.unorm8(c) for c in t[:3]]
        d=max(abs(u8[i]-int(rgba[y,x,i])) for i in range(3))
        isect=S.Pointer([[[[0.0]*3,[0.0]*3,0.0,0.0,0],3000.0,[0.0]*3,[0.0]*3]])
        found=mc.run(tr,[S.Pointer([rng.primary_ray]),isect])
        hid=S.tracer_hit_id(found,isect.load())
        if hid!=int(ids[y,x]): idbad+=1; details.append(('id',x,y,hid,int(ids[y,x])))
        rel=max(abs(rng.radiance_sum[i]-acc[y,x,i])/max(abs(acc[y,x,i]),1e-3) for i in range(3))
        relmax=max(relmax,rel)
        if d>1 or rel>5e-3: pbad+=1; details.append(('px',x,y,u8,rgba[y,x,:3].tolist(),rng.radiance_sum,acc[y,x,:3].tolist()))
    res.update(p_bad=pbad,id_bad=idbad,p_relmax=relmax,details=details[:3])
    return res
if __name__=='__main__':
    t0=time.time()
    with mp.Pool(8) as pool:
        out=pool.map(work,range(int(sys.argv[1]) if len(sys.argv)>1 else 64))
    print('time',time.time()-t0)
    print('whitted: poses with >1 LSB pixels',sum(1 for o in out if o['w_bad']),'max float diff',max(o['w_max'] for o in out))
    print('path: bad pixels',sum(o['p_bad'] for o in out),'id mismatches',sum(o['id_bad'] for o in out),'max rel',max(o['p_relmax'] for o in out))
    for o in out:
        if o['w_bad'] or o['p_bad'] or o['id_bad']: print(o)
