"""Equivalence check (pure Python, no GPU) of k_classify's two S1 formulations in zm_kernels.cuh:

  old()  the per-thread marching scan (`scan_tile`, ZM_S1_ROWMASK=0): per cube, eight corner compares and the
         volume-boundary rules evaluated per voxel;
  new()  the row-mask formulation (`edge_rows` + phase B of `tile_body`, the default): five 32-bit masks per staged
         row, everything per cube / per slot as bit arithmetic on whole rows.

Both return the tile's `pl` array (six slot bit planes and the active-cube mask per row segment); they must agree on
every tile, including tiles cut by the volume boundary, `close` zero fill, and slab shards (Es_own = Es - 1).
Run: python tools/emulate_s1_rowmask.py [trials]"""
import sys

import numpy as np

TF,TM,TS,RM,RS=32,8,8,9,9
def old(lab,Ef,Em,Es,Es_own,ef0,em0,es0):
  pl=np.zeros((64,8),dtype=np.uint64)
  interior = ef0+TF+1<=Ef and em0+TM+1<=Em and es0+TS+1<=Es
  for ls in range(8):
    for j in range(8):
      bits=[0]*6; ab=0
      for lane in range(32):
        a=lab[ls,j,lane]; af=lab[ls,j,lane+1]; as_=lab[ls+1,j,lane]; afs=lab[ls+1,j,lane+1]
        am=lab[ls,j+1,lane]; amf=lab[ls,j+1,lane+1]; ams=lab[ls+1,j+1,lane]; amfs=lab[ls+1,j+1,lane+1]
        nef=a!=af; nes=a!=as_; nem=a!=am
        eq_row=(not nef) and (not nes) and a==afs
        eq_row2=(am==amf) and (am==ams) and am==amfs
        uniform=eq_row and eq_row2 and not nem
        za=a!=0; zf=af!=0; zs=as_!=0; zm=am!=0
        ef=ef0+lane; em=em0+j; es=es0+ls
        if interior:
          pf,pm,ps=nef,nem,nes; act=not uniform
        else:
          valid=ef<Ef and em<Em and es<Es_own
          nf1=ef+1<Ef; nm1=em+1<Em; ns1=es+1<Es
          pf=nef and valid and nf1; pm=nem and valid and nm1; ps=nes and valid and ns1
          act=(pf and (za or zf)) or (pm and (za or zm)) or (ps and (za or zs)) or (valid and nf1 and nm1 and ns1 and not uniform)
        if act: ab|=1<<lane
        for k,c in enumerate((pf and za,pf and zf,pm and za,pm and zm,ps and za,ps and zs)):
          if c: bits[k]|=1<<lane
      if ab==0: bits=[0]*6
      pl[ls*8+j,:6]=bits; pl[ls*8+j,7]=ab
  return pl
def new(lab,Ef,Em,Es,Es_own,ef0,em0,es0):
  FULL=0xffffffff
  em=np.zeros((81,5),dtype=np.uint64)
  for r in range(81):
    ls,lm=divmod(r,9)
    m=[0]*5
    for lane in range(32):
      a=lab[ls,lm,lane]; af=lab[ls,lm,lane+1]
      am=lab[ls,lm+1,lane] if lm!=8 else a
      as_=lab[ls+1,lm,lane] if r<72 else a
      for k,c in enumerate((a!=af,a!=am,a!=as_,a!=0,af!=0)):
        if c: m[k]|=1<<lane
    em[r]=m
  pl=np.zeros((64,8),dtype=np.uint64)
  for warp in range(8):
    for lane in range(8):
      r0=warp*9+lane
      q=[int(x) for x in em[r0]]; qm=[int(x) for x in em[r0+1]]; qs=[int(x) for x in em[r0+9]]; ef3=int(em[r0+10][0])
      zf=q[4]
      nonuni=q[0]|qm[0]|qs[0]|ef3|q[1]|qs[1]|q[2]
      pf,pm,ps,cube=q[0],q[1],q[2],nonuni
      if not (ef0+TF+1<=Ef and em0+TM+1<=Em and es0+TS+1<=Es):
        nfv=Ef-ef0
        VF=FULL if nfv>=32 else (1<<nfv)-1
        NF1=FULL if nfv>=33 else (1<<(nfv-1))-1
        em_=em0+lane; es_=es0+warp
        rowok=em_<Em and es_<Es_own; nm1=em_+1<Em; ns1=es_+1<Es
        pf=pf&NF1 if rowok else 0
        pm=pm&VF if rowok and nm1 else 0
        ps=ps&VF if rowok and ns1 else 0
        cube=nonuni&NF1 if rowok and nm1 and ns1 else 0
      b=[pf&q[3],pf&zf,pm&q[3],pm&qm[3],ps&q[3],ps&qs[3]]
      act=b[0]|b[1]|b[2]|b[3]|b[4]|b[5]|cube
      pl[warp*8+lane,:6]=b; pl[warp*8+lane,7]=act
  return pl
def check(trials=300, seed=1):
  rng = np.random.default_rng(seed)
  n = 0
  for trial in range(trials):
    nl=rng.integers(1,5)
    mode=trial%4
    lab=rng.integers(0,nl+1,size=(9,9,34)).astype(np.uint64)
    if mode==1:  # smooth-ish: blocks
      lab=np.repeat(np.repeat(np.repeat(rng.integers(0,3,size=(3,3,9)),3,0),3,1),4,2)[:, :, :34].astype(np.uint64)
    # geometry: random tile position and volume extents
    Ef=int(rng.integers(1,100)); Em=int(rng.integers(1,30)); Es=int(rng.integers(2,30))
    if mode==2: Ef,Em,Es=200,200,200
    Es_own=Es if rng.integers(0,2) else Es-1
    ntf=(Ef+31)//32; ntm=(Em+7)//8; nts=(Es_own+7)//8
    tf=int(rng.integers(0,ntf)); tm=int(rng.integers(0,ntm)); ts=int(rng.integers(0,nts))
    ef0,em0,es0=tf*32,tm*8,ts*8
    # zero fill outside the buffer (Ef/Em); planes beyond Es may hold data (slab) -> leave random
    for lf in range(34):
      if ef0+lf>=Ef: lab[:,:,lf]=0
    for lm in range(9):
      if em0+lm>=Em: lab[:,lm,:]=0
    a=old(lab,Ef,Em,Es,Es_own,ef0,em0,es0); b=new(lab,Ef,Em,Es,Es_own,ef0,em0,es0)
    assert np.array_equal(a,b),(trial,Ef,Em,Es,Es_own,ef0,em0,es0,np.argwhere(a!=b)[:5])
    n+=1
  return n


if __name__ == "__main__":
  print("tiles checked:", check(int(sys.argv[1]) if len(sys.argv) > 1 else 300))
