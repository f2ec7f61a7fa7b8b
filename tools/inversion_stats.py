import sys, os, ctypes as C, importlib, json
import numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tools')
import uvrt_testlib as T
import traversal_lab as TL
uv=importlib.import_module("small-project-uv-robot-ray-tracer_b200")
B=importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")
from soup import make_soup, soup_route
L=TL.lab(); O=T.oracle(); f32=np.float32
out=[]
def run(name,tris,nodes,ti,lamps,length,P,seeds):
    tot_acc=0; tot_inv=0; mabs=0.0; mrel=0.0; rays_n=0
    for k,lp in enumerate(lamps):
        rays=np.zeros(P,dtype=T.RAY_DT); O.orc_generate(T.ptr(rays),0,P,lp[0],lp[1],lp[2],f32(length),seeds+k*977,None)
        s=TL.Stats(); L.lab_exact_stats(T.ptr(tris),T.ptr(rays),T.ptr(nodes),T.ptr(ti),P,0,C.byref(s))
        tot_acc+=s.accepted; tot_inv+=s.inverted; mabs=max(mabs,s.maxInvAbs); mrel=max(mrel,s.maxInvRel); rays_n+=P
    r={"scene":name,"rays":rays_n,"accepted_ray_triangle_pairs":int(tot_acc),"inverted_pairs":int(tot_inv),"max_inversion_abs":mabs,"max_inversion_rel":mrel,
       "margin_abs":2.0**-14,"margin_rel":2.0**-12}
    print(json.dumps(r),flush=True); out.append(r)
sim=uv.Sim(asset_root=T.DATA); sim.load_mesh("testroomopt")
for route in ("route","lange_route"):
    sim.load_route(route)
    tris,nodes,ti=sim.mesh_data(); floor=sim.mesh_info()["floor"]; p=sim.params
    lamps=[(f32(x),f32(f32(floor)+f32(p.lightHeight)),f32(y)) for x,y,_ in sim.positions]
    run("testroomopt x "+route, tris,nodes,ti,lamps,p.lightLength,2796202,0)
tris,nodes,ti=B.build_bvh(make_soup(1000000)); lamps=[(f32(x),f32(0.5),f32(z)) for x,z,_ in soup_route()]
run("soup 1M", tris,nodes,ti,lamps,1.0,500000,3)
json.dump({"what":"accepted (ray, triangle) pairs whose Moeller-Trumbore t is not above the exact entry distance of the triangle's leaf box (extend.cl arithmetic, tools/traversal_lab.c lab_exact_stats), against the margins of csrc/uvrt_fast.cuh","results":out}, open('/root/repo/profiles/r2_inversion_stats.json','w'), indent=1)
