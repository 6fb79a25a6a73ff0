"""CPU study for DESIGN.md section 8: how many int8 slices would an Ozaki-style emulation of the complex-FP64
rotation GEMM need to stay inside the 1e-10 contract?  Exact integer arithmetic in numpy (what int8
tensor cores with int32 accumulation would compute), synthetic operands of the north-star tile shape.
Not used by the library."""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
w=ge.load_package()
st,M,U=w.synthetic.make_model_arrays("bcc",(2,2,2),128,128)
A=M[0,0]; B=U[st["kpb_k"][0,0]]
def slices(X, axis, beta, s):
    # error-free splitting: scale each row (axis=1: per row) / column by a power of two, peel off beta bits at a time
    mx=np.abs(X).max(axis=axis,keepdims=True); mx[mx==0]=1
    e=np.ceil(np.log2(mx))            # |X| / 2^e <= 1
    R=X/2.0**e
    out=[]
    for i in range(s):
        q=np.trunc(R*2.0**(beta))      # integers in (-2^beta, 2^beta)
        out.append(q.astype(np.int64))
        R=R*2.0**beta-q                # exact
    return out,e
def ozaki(X,Y,beta,s):
    xs,ex=slices(X,1,beta,s); ys,ey=slices(Y,0,beta,s)
    C=np.zeros((X.shape[0],Y.shape[1]))
    nprod=0
    for lvl in range(s):              # products with i+j == lvl, accumulated exactly in integers
        acc=np.zeros((X.shape[0],Y.shape[1]),dtype=np.int64)
        for i in range(lvl+1):
            acc+=xs[i]@ys[lvl-i]; nprod+=1
        C+=acc.astype(np.float64)*2.0**(-beta*(lvl+2))
    return C*2.0**ex*2.0**ey, nprod
ref=A@B
for beta in (6,7):
    for s in range(4,10):
        # Gauss: P1=(Ar+Ai)Br, P2=Ar(Bi-Br), P3=Ai(Br+Bi)
        Ar,Ai,Br,Bi=A.real,A.imag,B.real,B.imag
        P1,n=ozaki(Ar+Ai,Br,beta,s); P2,_=ozaki(Ar,Bi-Br,beta,s); P3,_=ozaki(Ai,Br+Bi,beta,s)
        C=(P1-P3)+1j*(P1+P2)
        err=np.abs(C-ref).max()/np.abs(ref).max()
        print(f"beta={beta} slices={s} int8 products per real GEMM={n:3d} max rel err={err:.2e}")
