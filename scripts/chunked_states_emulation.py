"""CPU study (numpy, no GPU): final states of band 0 after a log-sine sweep -- sequential evaluation in double and in
80-bit, and a time-chunked evaluation (exact 8x8 transition matrix from long double, zero-state chunk responses and carry in
double) for 2 / 16 / 32 chunks.  Shows that the loss of accuracy in the 1e12-fold amplified, output-invisible state component
is a property of chunking in double, not of the CUDA kernels (profiles/r04u_chunked_states_emulation.txt)."""
import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'designs.npz'))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
K=4; n=480000; fs=48000.0
t=np.arange(n)/fs; T=n/fs; f0=10.0; f1=0.45*fs; k=(f1/f0)**(1.0/T)
phase=2*np.pi*f0*(k**t-1.0)/np.log(k); x=0.5*np.sin(phase)
c=sos[0]
def run(x, st, dtype):
    # sequential cascade, sample by sample across sections; returns final states
    w=st.astype(dtype).copy(); cc=c.astype(dtype)
    for i in range(len(x)):
        sgn=dtype(x[i])
        for kk in range(K):
            b0,b1,b2,_,a1,a2=cc[kk]
            w0=sgn-a1*w[kk,0]-a2*w[kk,1]
            sgn=b0*w0+b1*w[kk,0]+b2*w[kk,1]
            w[kk,1]=w[kk,0]; w[kk,0]=w0
    return w
ld=np.longdouble
truth=run(x,np.zeros((K,2)),ld)
seq=run(x,np.zeros((K,2)),np.float64)
print("truth s2", np.float64(truth[2]), "seq double", seq[2], "rel", abs(seq[2,0]-np.float64(truth[2,0]))/abs(np.float64(truth[2,0])))
for J in (2,16,32):
    L=n//J
    # transition matrix P (8x8) in long double: columns = response of state to unit initial states with zero input
    P=np.zeros((8,8),dtype=ld)
    for i in range(8):
        e=np.zeros((K,2),dtype=ld); e[i//2,i%2]=1
        P[:,i]=run(np.zeros(L),e,ld).ravel()
    Pd=P.astype(np.float64)
    st=np.zeros(8)
    for j in range(J-1):
        z=run(x[j*L:(j+1)*L],np.zeros((K,2)),np.float64).ravel()   # zero-state response in double
        st=Pd@st+z
    fin=run(x[(J-1)*L:],st.reshape(K,2),np.float64)
    # same with exact (long double) carry, to separate the sources
    stl=np.zeros(8,dtype=ld)
    for j in range(J-1):
        z=run(x[j*L:(j+1)*L],np.zeros((K,2)),ld).ravel()
        stl=P@stl+z
    finl=run(x[(J-1)*L:],np.float64(stl).reshape(K,2),np.float64)
    tr=np.float64(truth[2,0])
    print("J=%d: chunked double s2 %.6f rel %.2e | long-double carry then double last chunk rel %.2e | start-state s2 double %.6e vs ld %.6e" % (J, fin[2,0], abs(fin[2,0]-tr)/abs(tr), abs(finl[2,0]-tr)/abs(tr), st[4], np.float64(stl[4])))
