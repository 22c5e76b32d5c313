"""Timeline of one time-batched call of config 5 (library built with -DOHS_TB_TRACE): prints when each kernel finished."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import _bootstrap
pkg = _bootstrap.load_package()
S = pkg.signals
k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_streams, block, taps, fs = 256, 1024, 48000, 96000.0
eng = pkg.Engine(n_streams, block, taps, sample_rate=fs)
eng.set_hrir_set(S.synthetic_hrir_set(taps, 0.15 * fs)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.set_eq_enable(True); eng.set_gain(0.5)
n = block * k
x = torch.randn((n_streams, 2, n), device="cuda") * 0.1
y = torch.empty_like(x)
torch.cuda.synchronize()
for i in range(3):
    sys.stderr.write("--- call %d\n" % i)
    eng.process_device(x.data_ptr(), y.data_ptr(), n)
    eng.sync()
