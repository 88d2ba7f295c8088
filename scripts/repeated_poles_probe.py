import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import pyfilterbank_b200 as p
from oracle import oracle
rng=np.random.default_rng(0)
for K in (2,4,6,8):
    for r in (0.9, 0.99, 0.999):
        th=0.3
        row=[0.01,0.0,0.01,1.0,-2*r*np.cos(th), r*r]
        sos=np.array([[row]*K])            # (1,K,6) identical sections
        n=200000
        x=rng.standard_normal((1,n))
        y0,s0=oracle.sos_bank(x,sos)
        bank=p.Bank(sos)
        y1,s1=p.bank_filter(bank, torch.from_numpy(x).cuda())
        info=bank.last_launch()
        e=np.linalg.norm(y1.cpu().numpy()-y0)/np.linalg.norm(y0)
        print("K=%d identical sections, pole radius %.3f: rel L2 %.2e  (J=%d, path %d)"%(K,r,e,round(n/max(info['chunk'],1)),info['aligned']))
