python -m pytest tests/test_gpu_fastfft.py tests/test_gpu_functionals.py -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import profess_ad_b200.functionals as F
from profess_ad_b200.synthetic import smooth_supercell
from profess_ad_b200 import _native
lib = _native.load_library()
dev = torch.device('cuda:0')
box, den = smooth_supercell(256, 4, device=dev)
def rate(f, n=30):
    def go():
        x = den.requires_grad_(True); E = f(box, x); torch.autograd.grad(E, x); den.requires_grad_(False)
    for _ in range(3): go()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): go()
    e1.record(); torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) * 1e-3)
for fast in (0, 1):
    lib.pad_set_fast_fft(fast)
    print('fast' if fast else 'cuFFT', 'WT %.1f  WGC98 %.1f  SM %.1f evals/s at 256^3' % (rate(F.WangTeter), rate(F.WangGovindCarter98), rate(F.SmargiassiMadden)))
import ctypes
lib.pad_profile_begin()
x = den.requires_grad_(True)
for _ in range(5):
    E = F.WangTeter(box, x); torch.autograd.grad(E, x)
torch.cuda.synchronize()
names = ctypes.create_string_buffer(48 * 64); ms = (ctypes.c_double * 64)(); n_st, n_ev = ctypes.c_int(0), ctypes.c_int(0)
lib.pad_profile_end(names, ms, 64, ctypes.byref(n_st), ctypes.byref(n_ev))
for i in range(n_st.value):
    print('   %-44s %7.1f us' % (names.raw[48*i:48*i+48].split(b'\0')[0].decode(), ms[i] * 1e3))
PY
