"""Live CUDA-event stage times (pad_profile_begin / pad_profile_end) of one evaluation: WGC99 alone and the fused term list.
    python scripts/stage_profile.py [grid]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import profess_ad_b200.functionals as F
from profess_ad_b200 import _density_opt as D, _native
from profess_ad_b200.synthetic import smooth_supercell

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device('cuda:0')
lib = _native.load_library()
box_h, den_h = smooth_supercell(grid, max(1, grid // 64))
box, den = box_h.to(dev), den_h.to(dev)
v_ext = -0.1 * torch.rand_like(den)


def stages(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    lib.pad_profile_begin()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    names = ctypes.create_string_buffer(48 * 256)
    ms = (ctypes.c_double * 256)()
    n_st, n_ev = ctypes.c_int(0), ctypes.c_int(0)
    lib.pad_profile_end(names, ms, 256, ctypes.byref(n_st), ctypes.byref(n_ev))
    tot = 0.0
    for i in range(n_st.value):
        nm = names.raw[48 * i:48 * i + 48].split(b'\0')[0].decode()
        print('    %-62s %8.1f us' % (nm, 1e3 * ms[i]))
        tot += ms[i]
    print('    %-62s %8.1f us' % ('sum', 1e3 * tot))


wgc = F.WangGovindCarter99()
print('WGC99 E+V', grid)
stages(lambda: F.energy_and_potential(box, den, wgc.forward))
for name, terms in (('IonElectron + Hartree + WGC99 + PZ', [F.IonElectron, F.Hartree, wgc.forward, F.PerdewZunger]),
                    ('Hartree + WGC99', [F.Hartree, wgc.forward])):
    T = D.describe_terms(terms)
    print(name, '(fused term list)', grid)
    stages(lambda: D.eval_total(box, den, v_ext, T))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        D.eval_total(box, den, v_ext, T)
    ev0.record()
    for _ in range(20):
        D.eval_total(box, den, v_ext, T)
    ev1.record()
    torch.cuda.synchronize()
    print('    wall per evaluation (graph replay allowed): %.1f us' % (1e3 * ev0.elapsed_time(ev1) / 20))
