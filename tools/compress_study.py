"""Feasibility numbers behind the compressed index stream (csrc/sell_window.h): per window of
the sliced layout, the span of the primaries and the number of distinct member-to-primary
distances, for structured boxes and for a box whose elements are numbered in random order.
Host-side, no GPU.    python tools/compress_study.py"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(_ROOT, "tools"), os.path.join(_ROOT, "exaflow-exags-upc_b200")]
import layout_study as LS
from exags_b200 import semmesh
def study(ids, name):
    idx,mask=LS.layout(ids)
    nb=idx.shape[0]; nw=nb//8
    NONE=0xFFFFFFFF
    spans=[]; ndelta=[]; 
    for w in range(nw):
        I=idx[w*8:(w+1)*8].reshape(256,8).astype(np.int64); M=mask[w*8:(w+1)*8].reshape(256)
        start=(M&0x55)
        prim=np.full((256,8),-1,np.int64)
        cur=np.full(256,-1,np.int64)
        deltas=set()
        for j in range(8):
            isstart=((start>>j)&1).astype(bool)
            cur=np.where(isstart,I[:,j],cur)
            valid=(I[:,j]!=NONE)
            prim[:,j]=np.where(valid,cur,-1)
            d=(I[:,j]-cur)[valid & ~isstart]
            deltas.update(d.tolist())
        ps=I[((start[:,None]>>np.arange(8))&1).astype(bool) & (I!=NONE)]
        spans.append(int(ps.max()-ps.min()) if ps.size else 0); ndelta.append(len(deltas))
    spans=np.array(spans); ndelta=np.array(ndelta)
    print(f"{name}: windows {nw}; primary span max {spans.max()} (16-bit ok: {(spans<65536).mean()*100:.1f} %); distinct deltas per window: median {int(np.median(ndelta))}, max {ndelta.max()} (<=64: {(ndelta<=64).mean()*100:.1f} %, <=255: {(ndelta<=255).mean()*100:.1f} %)")
study(semmesh.box_ids(16,16,12,7),"box 16x16x12 N=7")
study(semmesh.box_ids(24,20,8,7),"box 24x20x8 N=7")
study(semmesh.box_ids(10,10,8,9),"box 10x10x8 N=9")
# irregular element numbering: permute the elements of a box
def permuted(Ex,Ey,Ez,N,seed):
    ids=semmesh.box_ids(Ex,Ey,Ez,N); m=(N+1)**3
    perm=np.random.default_rng(seed).permutation(Ex*Ey*Ez)
    return ids.reshape(-1,m)[perm].reshape(-1)
study(permuted(12,12,8,7,1),"box 12x12x8 N=7, elements in random order")
