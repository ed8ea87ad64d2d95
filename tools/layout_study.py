"""Where the L1 line visits of the sliced layout come from (host-side study, no GPU):
builds the layout of a SEM box through tests/c/setup_emul.cpp and counts, per load instruction
of a warp, the distinct 128-byte lines -- in total, per width class (pairs / edges / corners)
and per face direction of the pairs -- next to the number of distinct lines each class touches
(the floor for a class-separated dealing).     python tools/layout_study.py"""
import ctypes, os, sys, subprocess, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT+'/exaflow-exags-upc_b200']
from exags_b200 import semmesh
CSRC=ROOT+'/exaflow-exags-upc_b200/csrc'; OUT='/tmp/emul_build'
os.makedirs(OUT,exist_ok=True)
objs=[]
for c in ("topo.c","sort.c","base.c"):
    o=f"{OUT}/{c[:-2]}.o"; subprocess.run(["gcc","-O2","-fPIC","-fopenmp","-std=gnu11","-I",ROOT+"/include","-c",f"{CSRC}/{c}","-o",o],check=True); objs.append(o)
so=f"{OUT}/libsetup_emul.so"
subprocess.run(["g++","-O2","-fPIC","-std=c++17","-shared","-fopenmp","-I",CSRC,ROOT+"/tests/c/setup_emul.cpp"]+objs+["-o",so],check=True)
L=ctypes.CDLL(so)
L.setup_emul_sell_dump.restype=ctypes.c_size_t
L.setup_emul_sell_dump.argtypes=[ctypes.c_void_p,ctypes.c_int,ctypes.c_size_t,ctypes.c_void_p,ctypes.c_void_p,ctypes.c_size_t]
def layout(ids):
    nb=L.setup_emul_sell_dump(ids.ctypes.data,8,ids.size,None,None,0)
    idx=np.empty(nb*256,np.uint32); mask=np.empty(nb*32,np.uint8)
    L.setup_emul_sell_dump(ids.ctypes.data,8,ids.size,idx.ctypes.data,mask.ctypes.data,nb)
    return idx.reshape(nb,32,8),mask.reshape(nb,32)
def visits(idx,mask,shift=4):
    nb=idx.shape[0]; NONE=0xFFFFFFFF
    ph=mask & 0xAA
    # fetch order: swap slots (2j,2j+1) where phase bit set
    f=idx.copy()
    for j in range(0,8,2):
        sw=((ph>>(j+1))&1).astype(bool)
        a=f[:,:,j].copy(); b=f[:,:,j+1].copy()
        f[:,:,j]=np.where(sw,b,a); f[:,:,j+1]=np.where(sw,a,b)
    tot=0; per_slot=[]
    for j in range(8):
        v=f[:,:,j]
        lines=np.where(v==NONE,np.uint32(NONE),v>>shift)
        s=np.sort(lines,axis=1)
        distinct=(s[:,1:]!=s[:,:-1]).sum(axis=1)+1
        distinct-= (s[:,-1]==NONE).astype(int)  # NONE bucket is not a line
        per_slot.append(int(distinct.sum())); 
    return per_slot
if __name__=="__main__":
    dims=(16,16,12,7); ids=semmesh.box_ids(*dims); nel=dims[0]*dims[1]*dims[2]
    idx,mask=layout(ids)
    ps=visits(idx,mask)
    print("blocks",idx.shape[0],"visits/element",sum(ps)/nel,[round(p/nel,1) for p in ps])
    # lower bound: distinct lines with members per element
    mem=idx[idx!=0xFFFFFFFF]; print("member lines per element", np.unique(mem>>4).size/nel, "members/el", mem.size/nel)

def classes(idx,mask):
    """width class (2,4,8) of every slot, 0 for empty"""
    nb=idx.shape[0]; start=(mask&0x55).astype(np.uint8)
    cls=np.zeros(idx.shape,np.uint8)
    for j0 in range(0,8,2):
        st=((start>>j0)&1).astype(bool)
        # width = 2 if next start at j0+2 or slot j0+2 empty..., determine by scanning
        for w in (8,4,2):
            if j0+w>8: continue
            ok=st.copy()
            for q in range(j0+2,j0+w,2): ok&=~(((start>>q)&1).astype(bool))
            # members present at j0+w-? : group size > w/2
            if w>2: ok&=(idx[:,:,j0+w//2]!=0xFFFFFFFF)
            sel=ok&(cls[:,:,j0]==0)
            for q in range(j0,j0+w): cls[:,:,q]=np.where(sel,w,cls[:,:,q])
    cls[idx==0xFFFFFFFF]=0
    return cls
def visits_by_class(idx,mask,shift=4):
    NONE=0xFFFFFFFF; ph=mask&0xAA; f=idx.copy(); c=classes(idx,mask); cf=c.copy()
    for j in range(0,8,2):
        sw=((ph>>(j+1))&1).astype(bool)
        a=f[:,:,j].copy(); b=f[:,:,j+1].copy()
        f[:,:,j]=np.where(sw,b,a); f[:,:,j+1]=np.where(sw,a,b)
    out={}
    for w in (2,4,8):
        tot=0; nmem=0
        for j in range(8):
            v=np.where(cf[:,:,j]==w, f[:,:,j]>>shift, np.uint32(NONE))
            s=np.sort(v,axis=1); d=(s[:,1:]!=s[:,:-1]).sum(axis=1)+1-(s[:,-1]==NONE)
            tot+=int(d.sum()); nmem+=int((cf[:,:,j]==w).sum())
        lines=np.unique(idx[c==w]>>shift).size
        out[w]=(tot,nmem,lines)
    return out
if __name__=="__main__":
    r=visits_by_class(idx,mask)
    for w,(tot,nmem,lines) in r.items():
        print(f"class {w}: visits/el {tot/nel:.1f}  members/el {nmem/nel:.1f}  distinct lines/el {lines/nel:.1f}  members per visit {nmem/max(tot,1):.2f}")

def pair_types(idx,mask,dims,shift=4):
    Ex,Ey,Ez,N=dims; m=N+1
    NONE=0xFFFFFFFF; ph=mask&0xAA; c=classes(idx,mask)
    nb=idx.shape[0]
    # pair type per (block,lane,pairslot)
    res={}
    f=idx.copy(); 
    typ=np.zeros(idx.shape,np.uint8)
    for j in range(0,8,2):
        p=idx[:,:,j].astype(np.int64); q=idx[:,:,j+1].astype(np.int64); d=q-p
        ispair=(c[:,:,j]==2)
        tx=ispair&(d==m**3-N); ty=ispair&(d==Ex*m**3-N*m); tz=ispair&(d==Ex*Ey*m**3-N*m*m)
        t=np.where(tx,1,np.where(ty,2,np.where(tz,3,np.where(ispair,4,0)))).astype(np.uint8)
        typ[:,:,j]=t; typ[:,:,j+1]=t
        sw=((ph>>(j+1))&1).astype(bool)
        a=f[:,:,j].copy(); b=f[:,:,j+1].copy()
        f[:,:,j]=np.where(sw,b,a); f[:,:,j+1]=np.where(sw,a,b)
    for t,name in ((1,'x'),(2,'y'),(3,'z'),(4,'other')):
        tot=0;n=0
        for j in range(8):
            v=np.where(typ[:,:,j]==t, f[:,:,j]>>shift, np.uint32(NONE))
            s=np.sort(v,axis=1); d=(s[:,1:]!=s[:,:-1]).sum(axis=1)+1-(s[:,-1]==NONE)
            tot+=int(d.sum()); n+=int((typ[:,:,j]==t).sum())
        lines=np.unique(idx[typ==t]>>shift).size
        print(f"  pairs {name}: visits/el {tot/nel:.1f} members/el {n/nel:.1f} lines/el {lines/nel:.1f} members/visit {n/max(tot,1):.2f}")
if __name__=="__main__":
    pair_types(idx,mask,dims)
