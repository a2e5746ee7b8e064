"""CPU emulation of the index arithmetic of csrc/nmtf.cu::k_nmtf_sq_tiled (rectangles of C per pass, staging by rows, the
DMMA m8n8k4 fragment layout, the guarded stores): every entry of the Mp x Np slice is written exactly once and equals
A^T B.  Used while writing the kernel (no GPU needed); the GPU parity test is
tests/test_bnmtf_gpu.py::test_s_phase_reduction_is_the_einsum."""
import numpy as np
SQB=32; PAD=4
def tiling(K,L,vb):
    Mext=K*K+(K if vb else 0); Next=L*L+(L if vb else 0)
    Mp=(Mext+3)&~3; Np=(Next+3)&~3
    WTM=(Mp//4+3)//4; WTN=(Np//4+7)//8
    best=None; pn=1
    while pn<=SQB:
        pm=SQB//pn
        npass=-(-WTM//pm)*-(-WTN//pn); cols=16*pm+32*pn
        if best is None or npass<best[0] or (npass==best[0] and cols<best[1]): best=(npass,cols,pm,pn)
        pn<<=1
    npass,_,PM,PN=best
    return dict(Mext=Mext,Next=Next,Mp=Mp,Np=Np,WTM=WTM,WTN=WTN,PM=PM,PN=PN,npm=-(-WTM//PM),npn=-(-WTN//PN),npass=npass)
def emulate(K,L,vb,rows,RB,nparts,A_full,B_full):
    # A_full: rows x Mext (ff products | vf), B_full: rows x Next
    t=tiling(K,L,vb); Mp,Np=t['Mp'],t['Np']
    Cs=np.zeros((nparts,Mp,Np)); written=np.zeros((nparts,Mp,Np),int)
    nbatch=-(-rows//RB)
    for part in range(nparts):
      for p in range(t['npass']):
        pi,pj=divmod(p,t['npn']); AW=16*t['PM']; BW=32*t['PN']; a0=pi*AW; b0=pj*BW
        an=min(a0+AW,t['Mext'])-a0; bn=min(b0+BW,t['Next'])-b0
        AWp,BWp=AW+PAD,BW+PAD
        acc=np.zeros((16,2,32,2,4,2))
        for bt in range(part,nbatch,nparts):
            row0=bt*RB
            As=np.zeros(RB*AWp); Bs=np.zeros(RB*BWp)
            wpr=16//RB
            for warp in range(16):
                srow=warp%RB
                for lane in range(32):
                    scol=(warp//RB)*32+lane; sstep=32*wpr
                    inn=row0+srow<rows
                    for c in range(scol,bn,sstep):
                        Bs[srow*BWp+c]=B_full[row0+srow,b0+c] if inn else 0.0
                    for ca in range(scol,an,sstep):
                        As[srow*AWp+ca]=A_full[row0+srow,a0+ca] if inn else 0.0
            for kk in range(0,RB,4):
              for warp in range(16):
                for q in range(2):
                    b=q*16+warp; pm,pn=divmod(b,t['PN']); aoff=16*pm; boff=32*pn
                    if not (a0+aoff<Mp and b0+boff<Np): continue
                    # gather fragments per lane, then DMMA semantics
                    Af=np.zeros((2,8,4)); Bf=np.zeros((4,4,8))
                    for lane in range(32):
                        g,t4=lane>>2,lane&3
                        ar=(kk+t4)*AWp+g; br=(kk+t4)*BWp+g
                        Af[0,g,t4]=As[ar+aoff]; Af[1,g,t4]=As[ar+aoff+8]
                        for j in range(4): Bf[j,t4,g]=Bs[br+boff+8*j]
                    for i in range(2):
                        for j in range(4):
                            D=Af[i]@Bf[j]   # 8x8
                            for lane in range(32):
                                g,t4=lane>>2,lane&3
                                acc[warp,q,lane,i,j,0]+=D[g,2*t4]; acc[warp,q,lane,i,j,1]+=D[g,2*t4+1]
        for warp in range(16):
            for q in range(2):
                b=q*16+warp; pm,pn=divmod(b,t['PN']); aoff=16*pm; boff=32*pn
                if not (a0+aoff<Mp and b0+boff<Np): continue
                for lane in range(32):
                    g,t4=lane>>2,lane&3
                    for i in range(2):
                        for j in range(4):
                            m=a0+aoff+8*i+g; n=b0+boff+8*j+2*t4
                            if m<Mp and n<Np:
                                Cs[part,m,n]=acc[warp,q,lane,i,j,0]; Cs[part,m,n+1]=acc[warp,q,lane,i,j,1]
                                written[part,m,n]+=1; written[part,m,n+1]+=1
    assert (written==1).all()
    return Cs.sum(0)
rng=np.random.RandomState(0)
for K,L,vb,rows,RB,nparts in [(3,4,1,37,16,2),(5,9,0,21,16,3),(12,12,1,19,8,1),(2,30,1,9,4,2),(10,10,1,33,16,1)]:
    t=tiling(K,L,vb)
    A=rng.rand(rows,t['Mext']); B=rng.rand(rows,t['Next'])
    C=emulate(K,L,vb,rows,RB,nparts,A,B)
    ref=A.T@B
    err=np.abs(C[:t['Mext'],:t['Next']]-ref).max()
    assert np.abs(C[t['Mext']:]).max(initial=0)==0 and np.abs(C[:,t['Next']:]).max(initial=0)==0
    print(K,L,vb,t['npass'],'err',err)
