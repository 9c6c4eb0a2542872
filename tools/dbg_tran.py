import sys, json, numpy as np
sys.path.insert(0,'.')
from multiband_rf_pulse_design_b200 import fir
from oracle.fir_problems import build_fir_ap, solve_fir_ap_highs
k=json.load(open('./tests/golden/fir_ap_known.json'))['lowpass_n24']
f=np.array(k['f'],float)
def widen(fa):
    fn=f.copy(); fn[0::2]-=fa; fn[1::2]+=fa; return np.clip(fn,-1,1)
fas=np.linspace(0.0,0.15,16)
hs,st,ex=fir.fir_ap_cvx_batch(k['n'],[widen(v) for v in fas],k['a'],k['d'],[0.1]*16,[0.02]*16,return_info=True,max_iter=100000)
for v,s,i in zip(fas,st,ex['info']):
    r,_=solve_fir_ap_highs(build_fir_ap(k['n'],widen(v),k['a'],k['d'],0.1,0.02))
    print('%.4f'%v,s,'status',i[0],'iters',i[1],'obj %.6f'%i[2],'pr %.1e dr %.1e'%(i[4],i[5]),'rig %.4f'%i[6],'| highs',r.status, r.fun if r.status==0 else None)
