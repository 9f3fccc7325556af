import os; os.environ.setdefault("HMMCUDA_NO_PIPELINE", "1")  # kernel timing: keep the unpipelined single-launch path
import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
hm=ge.load_package()
import ctypes as C
K,N=60,3
temps=np.stack([hm.create_spike_template(K,3.0,0.8,0.2),hm.create_spike_template(K,4.0,0.3,0.2),hm.create_spike_template(K,2.0,0.5,0.3)],1)
pp=np.array([0.003,0.001,0.002])
T=int(sys.argv[1]) if len(sys.argv)>1 else 18_000_000
t=time.time(); S=hm.create_signal(T,0.3,pp,temps,hm.make_rng(2)); print('gen %.1fs'%(time.time()-t))
lA=hm.StateMatrix(N,K,np.log(pp),False); mu=np.asfortranarray(temps.copy()); mu[0,:]=0
for it in range(4):
    t=time.time(); x,ll,info=hm.viterbi(S,lA,mu,0.3,mode='ring',return_info=True); dt=time.time()-t
    print('ring  wall %.1f ms  device %.2f ms kernels %.2f ms top %.2f ms  chunks %d repaired %d/%d  -> %.1f Msamples/s (kernels)'%(dt*1e3,info['device_ms'],info['kernel_ms'],info['top_kernel_ms'],info['n_chunks'],info['fwd_repaired'],info['bwd_repaired'],T/info['kernel_ms']/1e3))
print('noise frac',(x==1).mean(),'ll',ll)
for W in (256,512,1024,2048):
    hm.set_ring_params(0,W)
    x2,ll2,info=hm.viterbi(S,lA,mu,0.3,mode='ring',return_info=True)
    print('W',W,'kernels %.2f ms top %.2f'%(info['kernel_ms'],info['top_kernel_ms']),'repaired',info['fwd_repaired'],info['bwd_repaired'],'same',np.array_equal(x,x2))
hm.set_ring_params(0,0)
Tf=min(T,2_000_000)
t=time.time(); xf,llf,info=hm.viterbi(S[:Tf],lA,mu,0.3,mode='faithful',return_info=True); dt=time.time()-t
print('faithful T=%d wall %.1f ms kernels %.2f ms -> %.2f Msamples/s'%(Tf,dt*1e3,info['kernel_ms'],Tf/info['kernel_ms']/1e3))
xr,llr=hm.viterbi(S[:Tf],lA,mu,0.3,mode='ring')
print('ring==faithful on prefix:',np.array_equal(xr,xf), abs(llr-llf)/abs(llf))
