"""CPU simulation behind the lane-walk blend kernels (profiles/ncu_blend_r2.md):
    python tests/sim_lane_walk.py [CFG] [VIEW] [BATCH]          (default C3 17 128; ~2 min, uses the CPU oracle)
For every (8x4 pixel patch, staged batch) of the oracle's tile lists it counts the (pixel, surfel) pairs with alpha >= 1/255 in
front of the pixel's last contributor, the iterations a warp needs when it visits ONE surfel at a time (= surfels with at least
one contributing lane) and the iterations it needs when every lane walks ITS OWN hit list (= the longest lane list).
C3 view 17: 11.97 M pairs, 1.306 M vs 0.590 M iterations (batch 128), 0.665 M (64), 0.535 M (256)."""
import sys, time, numpy as np
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # test infrastructure: the only kind of script that may use oracle/
sys.path[:0]=[ROOT,os.path.join(ROOT,'dynamic-2dgs_b200'),os.path.join(ROOT,'tests')]
import util
from oracle import surfel_oracle as so
cfg=sys.argv[1] if len(sys.argv)>1 else 'C3'; cam=int(sys.argv[2]) if len(sys.argv)>2 else 17
act,kw=util.raster_inputs(cfg,cam_index=cam,n_cams=100,bg=(0,0,0))
t=time.time()
st=so.forward(bg=kw['bg'],means3D=act['means3D'],opacities=act['opacities'],scales=act['scales'],rotations=act['rotations'],shs=act['shs'],sh_degree=3,
   viewmatrix=kw['viewmatrix'],projmatrix=kw['projmatrix'],campos=kw['campos'],tanfovx=kw['tanfovx'],tanfovy=kw['tanfovy'],image_height=kw['image_height'],image_width=kw['image_width'])
print('oracle fwd',time.time()-t,'R',st.num_rendered)
W,H=st.W,st.H; gx=(W+15)//16; gy=(H+15)//16
T=st.transMat.reshape(-1,9).astype(np.float64); m2=st.means2D.astype(np.float64); op=st.normal_opacity[:,3].astype(np.float64)
ranges=st.ranges.reshape(-1,2); pl=st.point_list
ncon=st.n_contrib.reshape(-1,H,W)[0]
B=int(sys.argv[3]) if len(sys.argv)>3 else 128
tot=dict(pairs=0,surv_any=0,iters_lane=0,balanced=0,tiles=0, batches=0)
ys,xs=np.mgrid[0:16,0:16]
for tile in range(gx*gy):
    a,b=ranges[tile]
    if b<=a: continue
    tx,ty=tile%gx,tile//gx
    px=(tx*16+xs+0.5).ravel(); py=(ty*16+ys+0.5).ravel()
    ids=pl[a:b]
    Tu=T[ids,0:3];Tv=T[ids,3:6];Tw=T[ids,6:9]
    k=px[:,None,None]*Tw[None]-Tu[None]; l=py[:,None,None]*Tw[None]-Tv[None]
    p=np.cross(k,l)
    with np.errstate(all='ignore'):
        s=p[...,:2]/p[...,2:3]
        rho3=(s**2).sum(-1)
        d=m2[ids][None]-np.stack([px,py],-1)[:,None]
        rho2=2*(d**2).sum(-1)
        rho=np.minimum(rho3,rho2)
        alpha=np.minimum(0.99,op[ids][None]*np.exp(-0.5*rho))
    hit=(alpha>=1/255)   # (256, L)
    # limit by n_contrib (last contributor per pixel)
    pixx=(tx*16+xs).ravel(); pixy=(ty*16+ys).ravel()
    inside=(pixx<W)&(pixy<H)
    lc=np.where(inside, ncon[np.minimum(pixy,H-1),np.minimum(pixx,W-1)],0)
    L=b-a
    hit&= (np.arange(L)[None,:]<lc[:,None])
    # patches: 8 warps of 8x4
    lx=xs.ravel(); ly=ys.ravel()
    patch=(ly//4)*2+(lx//8)
    for w in range(8):
        hw=hit[patch==w]   # (32, L)
        for b0 in range(0,L,B):
            hb=hw[:,b0:b0+B]
            n=hb.sum()
            if n==0: tot['batches']+=1; continue
            tot['pairs']+=n; tot['surv_any']+=int(hb.any(0).sum()); tot['iters_lane']+=int(hb.sum(1).max()); tot['balanced']+=n/32; tot['batches']+=1
    tot['tiles']+=1
print(cfg,cam,'batch',B,tot)
print('lanes/iter now %.2f ; per-lane walk: iters %.3fM vs now %.3fM ; efficiency %.2f'%(tot['pairs']/tot['surv_any'], tot['iters_lane']/1e6, tot['surv_any']/1e6, tot['balanced']/tot['iters_lane']))
