"""Integer-domain restatement of the BEHZ base conversions used by k_ext_conv / k_floor_sk (fhe_precompiles_b200/csrc/kernels.cu),
checked here against SEAL 4.0's step-by-step form (RNSTool::fastbconv_m_tilde, sm_mrq, fast_floor, fastbconv_sk as restated in
oracle/bfv_oracle.c) on random and edge inputs.  Pure Python integers; every intermediate that the kernels keep in a 32- or 64-bit
register is asserted to fit.  usage: python scripts/check_integer_domain.py [iterations]"""
import sys
import random
q0,q1,P,b0,b1,msk=0xffffee001,0xffffc4001,0x1ffffe0001,0x1ffffffffffa4001,0x1ffffffffff92001,0x1ffffffffffde001
q=q0*q1; MT=1<<32; M64=(1<<64)-1
inv=lambda a,m: pow(a,-1,m)
bsk=[b0,b1,msk]
E=[MT*inv(q1,q0)%q0, MT*inv(q0,q1)%q1]
# ---- ext: old
def ext_old(x0,x1):
    t0=x0*E[0]%q0; t1=x1*E[1]%q1
    ymt=(t0*q1+t1*q0)%MT
    rm=ymt*((-inv(q,MT))%MT)%MT
    out=[]
    for p in bsk:
        im=inv(MT,p)
        A=q1*im%p; B=q0*im%p; C=q*im%p
        rr=rm if rm<(1<<31) else rm+p-MT
        out.append((t0*A+t1*B+rr*C)%p)
    return out
c0=(1<<36)-q0; c1=(1<<36)-q1
qll=q&0xffffffff; qlh=(q>>32)&0xffffffff; qhh=q>>64
assert qhh<256
def ext_new(x0,x1):
    t0=x0*E[0]%q0; t1=x1*E[1]%q1
    Pv=t0*c1+t1*c0; assert Pv<1<<64
    # y0 = ((t0+t1)<<36) - Pv as 128-bit
    y0=((t0+t1)<<36)-Pv; assert y0>=0
    y0l=y0&M64; y0h=y0>>64
    ymt=y0l&0xffffffff
    assert ymt==((t0&0xffffffff)*(q1&0xffffffff)+(t1&0xffffffff)*(q0&0xffffffff))&0xffffffff
    rm=ymt*((-inv(q,MT))%MT)%MT
    neg=rm>=(1<<31)
    carry=1 if ymt!=0 else 0
    A=(y0>>32)+((rm*qll)>>32)+carry; assert A<1<<64
    Bv=rm*qlh; Cv=rm*qhh; assert Bv<1<<64 and Cv<1<<40
    s=A+Bv; co=s>>64; s&=M64
    m=(s>>61)+8*co+(Cv>>29); assert m<1<<12
    base=(s&((1<<61)-1))+((Cv&((1<<29)-1))<<32); assert base<1<<62
    out=[]
    for p in bsk:
        c=(1<<61)-p
        NQ=(p-q%p)%p
        v=base+m*c+(NQ if neg else 0); assert m*c<1<<32 and v<1<<63
        k=v>>61; v=(v&((1<<61)-1))+k*c
        if v>=p: v-=p
        assert v<p
        out.append(v)
    return out
def check_ext(iters):
  random.seed(3)
  for it in range(iters):
    x0=random.choice([0,1,q0-1,random.randrange(q0)]); x1=random.choice([0,1,q1-1,random.randrange(q1)])
    a=ext_old(x0,x1); b=ext_new(x0,x1)
    assert a==b,(x0,x1,a,b)
# ---- floor_sk old (as in kernel)
ipq=[inv(q1,q0), inv(q0,q1)]
Bp=b0*b1
def consts():
    C={}
    C['flV']=[inv(q%p,p) for p in bsk]
    C['flA']=[(p-(q1%p)*C['flV'][k]%p)%p for k,p in enumerate(bsk)]
    C['flB']=[(p-(q0%p)*C['flV'][k]%p)%p for k,p in enumerate(bsk)]
    ibj=[inv(b1%b0,b0), inv(b0%b1,b1)]
    C['skV']=[C['flV'][j]*ibj[j]%bsk[j] for j in range(2)]
    C['skA']=[C['flA'][j]*ibj[j]%bsk[j] for j in range(2)]
    C['skB']=[C['flB'][j]*ibj[j]%bsk[j] for j in range(2)]
    ib=inv(Bp%msk,msk)
    C['ib']=ib
    C['alK']=[(b1%msk)*ib%msk,(b0%msk)*ib%msk,(msk-C['flV'][2]*ib%msk)%msk,(msk-C['flA'][2]*ib%msk)%msk,(msk-C['flB'][2]*ib%msk)%msk]
    C['pBq']=[[b1%q0,b1%q1],[b0%q0,b0%q1]]
    C['Bq']=[Bp%q0,Bp%q1]
    return C
C=consts()
def floor_old(v0,v1,vb0,vb1,vsk):
    t0=v0*ipq[0]%q0; t1=v1*ipq[1]%q1
    tb0=(vb0*C['skV'][0]+t0*C['skA'][0]+t1*C['skB'][0])%b0
    tb1=(vb1*C['skV'][1]+t0*C['skA'][1]+t1*C['skB'][1])%b1
    al=(tb0*C['alK'][0]+tb1*C['alK'][1]+vsk*C['alK'][2]+t0*C['alK'][3]+t1*C['alK'][4])%msk
    neg=al>(msk>>1); am=msk-al if neg else al
    out=[]
    for l,ql in enumerate((q0,q1)):
        kb=C['Bq'][l] if neg else ql-C['Bq'][l]
        out.append((tb0*C['pBq'][0][l]+tb1*C['pBq'][1][l]+am*kb)%ql)
    return out,tb0,tb1,al
cb=[(1<<61)-p for p in bsk]
D1=cb[1]-cb[2]; D0=cb[0]-cb[2]; assert D1>0 and D0>0   # |d1|, |d0|
nib=(msk-C['ib'])%msk
def floor_new(v0,v1,vb0,vb1,vsk):
    t0=v0*ipq[0]%q0; t1=v1*ipq[1]%q1
    Pv=t0*c1+t1*c0
    y0=((t0+t1)<<36)-Pv
    lo=y0&((1<<61)-1); hi=y0>>61; assert hi<1<<13
    y0r=[lo+hi*c for c in cb]   # < 2^61+2^32
    tb=[]
    for j in range(2):
        p=bsk[j]
        x=[vb0,vb1][j]+2*p-y0r[j]; assert 0<x<1<<63, x
        tb.append(x*C['skV'][j]%p)
    tb0,tb1=tb
    # w = tb0*D1 + tb1*D0
    Wlo=(tb0&0xffffffff)*D1+(tb1&0xffffffff)*D0; Whi=(tb0>>32)*D1+(tb1>>32)*D0
    assert Wlo<1<<64 and Whi<1<<64
    wr=Wlo+((Whi&((1<<29)-1))<<32)+(Whi>>29)*cb[2]; assert wr<1<<63
    x2=vsk+2*msk-y0r[2]; assert 0<x2<1<<63
    al=(wr*nib+x2*C['alK'][2])%msk
    neg=al>(msk>>1); am=msk-al if neg else al
    out=[]
    for l,ql in enumerate((q0,q1)):
        kb=C['Bq'][l] if neg else ql-C['Bq'][l]
        out.append((tb0*C['pBq'][0][l]+tb1*C['pBq'][1][l]+am*kb)%ql)
    return out,tb0,tb1,al
def check_floor(iters):
  random.seed(4)
  for it in range(iters):
    v0=random.choice([0,1,q0-1,2*q0-1,random.randrange(2*q0)]); v1=random.choice([0,2*q1-1,random.randrange(2*q1)])
    vb0=random.choice([0,2*b0-1,random.randrange(2*b0)]); vb1=random.choice([0,2*b1-1,random.randrange(2*b1)]); vs=random.choice([0,2*msk-1,random.randrange(2*msk)])
    a=floor_old(v0,v1,vb0,vb1,vs); b=floor_new(v0,v1,vb0,vb1,vs)
    assert a==b,(a,b)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    check_ext(n)
    check_floor(n)
    print("integer-domain base conversions == step-by-step form on", n, "inputs each")
