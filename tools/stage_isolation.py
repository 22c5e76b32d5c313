import sys, json
sys.path.insert(0,'/root/repo')
import _bootstrap, torch, numpy as np
pkg=_bootstrap.load_package(); S=pkg.signals
def run(conv, eq, K=192):
    eng=pkg.Engine(1024,256,256); eng.set_hrir_set(S.synthetic_hrir_set(256,40.0)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.enable_timing()
    eng.set_eq_enable(eq); eng.set_conv_enable(conv); eng.set_gain(0.5)
    n=256*K
    x=torch.randn((1024,2,n),device='cuda')*0.1; y=torch.empty_like(x); torch.cuda.synchronize()
    for _ in range(3): eng.process_device(x.data_ptr(),y.data_ptr(),n)
    eng.sync(); ms=[]
    for _ in range(5):
        eng.process_device(x.data_ptr(),y.data_ptr(),n); ms.append(eng.last_kernel_ms())
    t=float(np.median(ms)); print(json.dumps({"conv":conv,"eq":eq,"ms":t,"us_per_block":t*1e3/K,"cycles_per_block":t*1e-3/K*1.965e9}))
run(True,True); run(False,True); run(True,False); run(False,False)
