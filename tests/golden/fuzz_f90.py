"""Differential fuzz: the reference's Fortran source (interpreted, oracle/f90interp.py) against the C++ oracle on random
small problems -- both precisions, odd spacings, clustered / on-node / far-outside / rounded points, k up to num_points.
Build-container tool (needs /root/reference):   python tests/golden/fuzz_f90.py SEED TRIALS [SIZE]
Prints one line per mismatch and "done seed S bad N".  tests/test_gpu_fuzz.py runs the same generator on the GPU
against the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scipy.spatial import Delaunay
from oracle.f90interp import Interpreter
from oracle import oracle as orc
REF='/root/reference'
FILES=("kd_tree.f90","query_esrg.f90","triangle_walk.f90","interpolation.f90")
src=[open(os.path.join(REF,f)).read() for f in FILES]
def farr(a,dt): return np.asfortranarray(np.asarray(a,dtype=dt))
seed=int(sys.argv[1]) if len(sys.argv)>1 else 0
SIZE=int(sys.argv[3]) if len(sys.argv)>3 else 1      # scales the number of points and the grid
rng=np.random.default_rng(seed)
nbad=0
for trial in range(int(sys.argv[2]) if len(sys.argv)>2 else 20):
    R = np.float64 if rng.uniform()<0.6 else np.float32
    itp=Interpreter(src,R)
    n=int(rng.integers(3,120*SIZE))
    mode=rng.integers(0,6)
    lx=int(rng.integers(2,14*SIZE)); ly=int(rng.integers(2,12*SIZE))
    spac=float(rng.choice([1.0,0.5,2.0,0.37]))
    x=(rng.uniform(-5,5)+spac*np.arange(lx)).astype(R); y=(rng.uniform(-5,5)+spac*np.arange(ly)).astype(R)
    if rng.uniform()<0.3:   # axes that are not exactly x(1) + i*spac: the ring windows no longer match the radii
        x=(x+rng.uniform(-0.2,0.2,lx).astype(R)*R(spac)).astype(R); y=(y+rng.uniform(-0.2,0.2,ly).astype(R)*R(spac)).astype(R)
    ex=float(x[-1]-x[0]); ey=float(y[-1]-y[0])
    if mode==0: pts=np.column_stack([rng.uniform(x[0]-spac,x[-1]+spac,n),rng.uniform(y[0]-spac,y[-1]+spac,n)])
    elif mode==1: pts=np.column_stack([rng.normal(x[0]+ex/2,ex/8,n),rng.normal(y[0]+ey/2,ey/8,n)])   # clustered, hull inside grid
    elif mode==2: pts=np.column_stack([rng.choice(x,n)+rng.choice([0,0,spac/2],n),rng.choice(y,n)+rng.choice([0,0,spac/2],n)])  # on nodes / cell borders (duplicates)
    elif mode==3: pts=np.column_stack([rng.uniform(x[0]-3*ex,x[-1]+3*ex,n),rng.uniform(y[0]-3*ey,y[-1]+3*ey,n)])  # mostly outside
    elif mode==5:   # unique lattice nodes / half steps: targets ON vertices and edges (margin band, path dependence)
        gx,gy=np.meshgrid(x[0]+0.5*spac*np.arange(2*lx-1),y[0]+0.5*spac*np.arange(2*ly-1))
        allp=np.column_stack([gx.ravel(),gy.ravel()]); sel=rng.permutation(len(allp))[:max(min(n,len(allp)),4)]
        pts=allp[sel]; n=len(pts)
    else: pts=np.round(np.column_stack([rng.uniform(x[0],x[-1],n),rng.uniform(y[0],y[-1],n)]),1)
    if rng.uniform()<0.3:   # another length scale / offset: the walk's margin 1e-10 and the k-d sentinel 1e6 are absolute
        sc=float(rng.choice([1.0e3,1.0e-3,37.0])); off=rng.uniform(-1.0,1.0,2)*1.0e3*sc
        x=(x.astype(np.float64)*sc+off[0]).astype(R); y=(y.astype(np.float64)*sc+off[1]).astype(R)
        pts=pts*sc+off; spac=float(R(spac*sc))
    pts=pts.astype(R).astype(np.float64)
    dat=rng.normal(size=n).astype(R)
    k=int(rng.integers(1,min(n,9)+1)) if rng.uniform()<0.8 else n if n<=32 else 8
    tag=f"seed{seed} t{trial} R{R.__name__} mode{mode} n{n} k{k} {lx}x{ly} spac{spac}"
    try:
        f=np.zeros((ly,lx),dtype=R,order='F')
        itp.call("idw_kdtree",farr(pts.T,R),n,dat,x,lx,y,ly,k,f)
        o=orc.idw_kdtree(pts.T,dat,x,y,k,dtype=R)
        if not np.array_equal(f,o,equal_nan=True):
            _,w=orc.idw_kdtree(pts.T,dat,x,y,k,dtype=R,debug=True)
            unset=(w.nn_index==0).any(0)      # d2 >= the 1.0e6 sentinel: the slot is never filled and the reference
            if np.array_equal(f[~unset],o[~unset],equal_nan=True):   # reads an unset index there (interpolation.f90:151,159)
                print("kd: differs only at",int(unset.sum()),"targets with unset slots (undefined in the reference)",tag,flush=True)
            else:
                nbad+=1; print("KD MISMATCH",tag,flush=True)
        f=np.zeros((ly,lx),dtype=R,order='F')
        itp.call("idw_esrg_nearest",farr(pts,R),n,dat,x,lx,y,ly,R(spac),k,f)
        o=orc.idw_esrg_nearest(pts,dat,x,y,spac,k,dtype=R)
        if not np.array_equal(f,o,equal_nan=True): nbad+=1; print("ESRG MISMATCH",tag,flush=True)
        if mode!=2:
            tri=Delaunay(pts)
            simp=tri.simplices.astype(np.int64)+1; nbr=tri.neighbors.astype(np.int64)+1
            try:
                o=orc.barycentric_interpolation(pts,dat,x,y,simp.astype(np.int32),nbr.astype(np.int32),dtype=R)
            except RuntimeError:
                o=None      # no grid node inside the hull: the reference's fill would never end
            if o is not None and len(simp)>=n:   # (fewer triangles than points: the reference reads out of bounds, interpolation.f90:347)
                f=np.zeros((ly,lx),dtype=R,order='F')
                itp.call("barycentric_interpolation",farr(pts,R),n,dat,x,lx,y,ly,farr(simp,np.int64),farr(nbr,np.int64),len(simp),f)
                if not np.array_equal(f,o,equal_nan=True): nbad+=1; print("BARY MISMATCH",tag,np.abs(f-o).max(),flush=True)
        fld=farr(rng.normal(size=(ly,lx)),R)
        q=np.column_stack([rng.uniform(x[0],x[-1]-1e-3*(x[-1]-x[0]),50),rng.uniform(y[0],y[-1]-1e-3*(y[-1]-y[0]),50)])
        q[:10,0]=rng.choice(x[:-1],10); q[5:15,1]=rng.choice(y[:-1],10)
        res=np.zeros(50,R)
        itp.call("bilinear",x,lx,y,ly,fld,farr(q,R),50,res)
        o=orc.bilinear(x,y,fld,q.astype(R),dtype=R)
        if not np.array_equal(res,o): nbad+=1; print("BIL MISMATCH",tag,flush=True)
    except Exception as e:
        print("EXC",tag,type(e).__name__,str(e)[:200],flush=True)
print("done seed",seed,"bad",nbad)
