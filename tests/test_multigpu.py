"""GPU parity of the multi-GPU data path: voices sharded across engines,
the stereo root bus summed before the truncating root stage.

* the cut path (`a2cu_set_post_root_stage(0)` -> integer sum -> `a2cu_apply_root_stage`) and
* the fused path (`a2cu_xchg_*`: the render kernel's last CTA pushes its root bus into every
  peer's buffer, waits for the world's flags, sums and runs the root stage)

must both equal ONE engine rendering all voices, which in turn equals the oracle / the golden
vectors from the reference (tests/test_cuda_parity.py). The in-process tests put all shards on
cuda:0 (peer memory = the same device memory, the kernels run concurrently on separate
streams); the torchrun test uses one process per GPU over CUDA IPC and is skipped on a
single-GPU box.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import CASES, bank
from scenarios import run_cuda, run_cuda_sharded, run_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _diff(a, b):
    bad = np.nonzero((a != b).any(axis=1))[0]
    return "first diff at frame %d (%d frames differ, max abs %d)" % (
        bad[0], len(bad), np.abs(a.astype(np.int64) - b).max())


def _root_ramp_bank(n=192, frames=1280):
    """cfg2-style bank + root panmix writes mid-render (a volume ramp, a pan ramp, a pan jump), so
    the stage above the cut is really non-trivial and truncating."""
    scn = bank(n, frames=frames)
    scn.root_writes = [(200 << 8, 0, 40000, 300 << 8), ((200 << 8) + 77, 1, -20000, 500 << 8),
                       (900 << 8, 1, 30000, 0)]
    return scn


def test_root_ramp_sharded_equals_single_engine_and_port():
    scn = _root_ramp_bank()
    whole = run_cuda(scn)
    assert np.array_equal(whole, run_oracle(scn)), _diff(whole, run_oracle(scn))
    for pipelined in (False, True):
        for o in run_cuda_sharded(scn, nshards=2, mode="fused", window=scn.buffer * 5, pipelined=pipelined):
            assert np.array_equal(o, whole), _diff(o, whole)
    cut = run_cuda_sharded(scn, nshards=2, mode="cut", window=scn.buffer * 5)[0]
    assert np.array_equal(cut, whole), _diff(cut, whole)


@pytest.mark.parametrize("name", ["bank256", "osc_pan_ramps", "filter_sweep", "additive8", "fm_all"])
def test_cut_path_equals_single_engine(name):
    scn = CASES[name]()
    if scn.ngroups:
        pytest.skip("groups shard as whole sub-trees; covered by the root-level cases")
    whole = run_cuda(scn)
    cut = run_cuda_sharded(scn, nshards=2, mode="cut", window=scn.buffer * 5)[0]
    assert np.array_equal(cut, whole), _diff(cut, whole)
    assert np.array_equal(whole, run_oracle(scn))


@pytest.mark.parametrize("pipelined", [False, True])
@pytest.mark.parametrize("nshards", [2, 3, 8])
def test_fused_exchange_equals_single_engine(nshards, pipelined):
    """pipelined: several windows in flight -> lagged exchange (window k's root stage runs at the
    tail of kernel k + 1); otherwise a2cu_collect finishes each window with the drain kernel."""
    scn = CASES["bank256"]()
    whole = run_cuda(scn)
    st = {}
    outs = run_cuda_sharded(scn, nshards=nshards, mode="fused", window=scn.buffer * 5, stats=st,
                            pipelined=pipelined)
    assert all(x > 0 for x in st["split_launches"]), "fused tail of render_split expected"
    for o in outs:      # every rank ends up with the identical master block
        assert np.array_equal(o, whole), _diff(o, whole)
    assert np.array_equal(whole, run_oracle(scn))


@pytest.mark.parametrize("name", ["osc_pan_ramps", "filter_sweep", "fm_all", "additive8"])
def test_fused_exchange_other_structures(name):
    """Mixed structures / FM chains take the thread-per-voice kernel and several banks per engine:
    the exchange then runs in its own single-CTA kernel (mix_root_xchg)."""
    scn = CASES[name]()
    if scn.ngroups:
        pytest.skip("root-level voices only")
    whole = run_cuda(scn)
    for o in run_cuda_sharded(scn, nshards=2, mode="fused", window=scn.buffer * 7, pipelined=True):
        assert np.array_equal(o, whole), _diff(o, whole)


def test_fused_exchange_thread_per_voice_kernel():
    scn = CASES["bank256"]()
    whole = run_cuda(scn, split=False)
    for o in run_cuda_sharded(scn, nshards=2, mode="fused", split=False, window=scn.buffer * 4):
        assert np.array_equal(o, whole), _diff(o, whole)


def test_exchange_timeout_is_reported_not_hung():
    """A peer that never renders: the in-kernel wait gives up and a2cu_collect reports it."""
    from audiality2_b200 import engine as eng
    from audiality2_b200.chains import autowire
    import torch
    a, b = eng.Engine(48000, 2), eng.Engine(48000, 2)
    try:
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        a.set_stream(sa.cuda_stream)
        b.set_stream(sb.cuda_stream)
        for e in (a, b):
            bk = e.new_bank(autowire(["wtosc", "panmix"]), 32)
            e.write_all(bk, 0, 0, [e.builtin_wave("sine") << 16])
            e.write_all(bk, 0, 2, [6000])
        a.xchg_create(0, 2, 256, timeout_ms=50)
        b.xchg_create(1, 2, 256, timeout_ms=50)
        a.xchg_connect_local([a, b])
        b.xchg_connect_local([a, b])
        t = a.submit(256, 64)           # b never submits
        with pytest.raises(eng.A2cuError, match="timed out"):
            a.collect(t)
    finally:
        a.close()
        b.close()


def test_two_process_exchange_over_ipc():
    """One process per GPU (torchrun, NCCL only as plumbing for the handle swap and the baseline):
    fused exchange == NCCL all-reduce cut path == single engine, on every rank."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29571",
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
