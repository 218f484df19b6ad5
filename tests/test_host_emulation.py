"""Device code on host threads (no GPU): the FFT building blocks and the work-queue scheduler of the pipelined kernels are
compiled with g++ against a 20-line stand-in for cuda_runtime.h (tests/host_emu/fake_cuda) -- every lane / CTA is a host
thread, __syncwarp / __syncthreads are std::barrier, atomics are std::atomic.  This is how index maps and the scheduling
protocol are checked before any GPU time is spent."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, 'host_emu')
CSRC = os.path.join(os.path.dirname(HERE), 'profess_ad_b200', 'csrc')

pytestmark = pytest.mark.skipif(shutil.which('g++') is None, reason='g++ not available')


def _build_and_run(tmp_path, source, extra_includes=()):
    exe = str(tmp_path / 'emu')
    cmd = ['g++', '-std=c++20', '-O1', '-pthread', '-I', os.path.join(EMU, 'fake_cuda'), '-I', CSRC]
    for inc in extra_includes:
        cmd += ['-I', inc]
    cmd += [os.path.join(EMU, source), '-o', exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_fft_building_blocks_on_host_threads(tmp_path):
    src = open(os.path.join(CSRC, 'fft_strided.cuh')).read()
    a = src.index('template <int L, bool WIDE = false>\nstruct SPass {')
    b = src.index('template <int L>\n__device__ __forceinline__ void spass_load_twiddles')
    (tmp_path / 'tile_part.h').write_text(src[a:b])
    out = _build_and_run(tmp_path, 'fft_blocks.cpp', extra_includes=[str(tmp_path)])
    assert 'tile 512 wide' in out and 'line 256/32' in out


def test_pipe_scheduler_on_host_threads(tmp_path):
    """Ticket scheduler of csrc/zy_pipe.cuh: every item runs exactly once, no item runs before all items of the stage below on
    its plane have run, no deadlock with fewer workers than planes in flight, control block zeroed at exit."""
    src = open(os.path.join(CSRC, 'zy_pipe.cuh')).read()
    a = src.index('#define PIPE_MAX_PLANES')
    b = src.index('//  y items')
    b = src.rindex('// ----', 0, b)
    (tmp_path / 'sched_part.h').write_text(src[a:b])
    out = _build_and_run(tmp_path, 'pipe_sched.cpp', extra_includes=[str(tmp_path)])
    assert 'bad = 0' in out


def test_kline_matches_kpoint_on_host(tmp_path):
    """|k|^2 along an x line from per-line constants (KLine, used by the fused x pass) == make_kpoint_at + sym_even."""
    src = open(os.path.join(CSRC, 'common.cuh')).read()
    a = src.index('struct KGeom {')
    b = src.index('// Effective wave-vector for the gradient multiplier')
    (tmp_path / 'kgeom_part.h').write_text(src[a:b])
    out = _build_and_run(tmp_path, 'kline.cpp', extra_includes=[str(tmp_path)])
    assert 'kline worst' in out
