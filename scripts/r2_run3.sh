#!/bin/bash
# round 2, run 3: streamed inverse z kernel (3 CTAs/SM) and fused / unfused mid pass, A/B; HC stress parity
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fastfft.py tests/test_gpu_ions.py -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -6 gpurun_out/r2c_pytest.log
for cfg in "1 1" "1 0" "0 0" "0 1"; do
  set -- $cfg
  PAD_ZINV_STREAM=$1 PAD_FUSE_MID=$2 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-denopt \
     > gpurun_out/r2c_bench_s$1_f$2.json 2> gpurun_out/r2c_bench_s$1_f$2.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2c_bench_s$1_f$2.json'))
    print('stream=$1 fuse_mid=$2', 'ms/step', round(d['ms_per_step'], 4), 'E', d.get('config_detail', d['config']).get('energy_Ha'), 'e2e', d['e2e']['value'])
    for k in d['roofline']['kernels']:
        print('    %-70s %8.1f us' % (k['stage'], 1e3 * k['ms_per_eval']))
    print('   also', json.dumps(d.get('also')))
except Exception as e:
    print('stream=$1 fuse_mid=$2 FAILED', e)
PY
done
python - <<'PY'
# stage profile of the fused term list (IonElectron + Hartree + WGC99 + PZ) at 256^3
import ctypes, torch, sys
sys.path.insert(0, '.')
import profess_ad_b200.functionals as F
from profess_ad_b200 import _density_opt as D, _native
from profess_ad_b200.synthetic import smooth_supercell
lib = _native.load_library()
dev = torch.device('cuda:0')
box, den = smooth_supercell(256, 4, device=dev)
T = D.describe_terms([F.IonElectron, F.Hartree, F.WangGovindCarter99().forward, F.PerdewZunger])
vx = -0.1 * den / den.mean()
for fuse in (1, 0):
    lib.pad_set_option(b'fuse_terms', fuse)
    for _ in range(3):
        D.eval_total(box, den, vx, T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        D.eval_total(box, den, vx, T)
    e1.record(); torch.cuda.synchronize()
    print('fuse_terms', fuse, 'ms/eval', e0.elapsed_time(e1) / 20)
    lib.pad_profile_begin()
    for _ in range(5):
        D.eval_total(box, den, vx, T)
    torch.cuda.synchronize()
    names = ctypes.create_string_buffer(48 * 256); ms = (ctypes.c_double * 256)(); n = ctypes.c_int(0); ne = ctypes.c_int(0)
    lib.pad_profile_end(names, ms, 256, ctypes.byref(n), ctypes.byref(ne))
    for i in range(n.value):
        print('    %-72s %8.1f us' % (names.raw[48 * i:48 * i + 48].split(b'\0')[0].decode(), 1e3 * ms[i]))
PY
