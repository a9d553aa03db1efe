"""Multi-GPU behind the C ABI (afsim_multi_*, include/afsim.h): one handle for the box, the partition + per-GPU renders
+ NCCL all-gather of the metric structs inside the library.  The single-device mask runs on any GPU box; the sharded
sweep needs two GPUs (`gpurun --gpus 2`) and must return exactly what one GPU returns for the same pairs."""
import numpy as np
import pytest

from audio_forge_b200 import abi, workloads
from tests.cases import FS, metric_mismatches

pytestmark = pytest.mark.gpu


def _workload():
    passages = [workloads.add_hum(workloads.speech_like(24000, seed=700 + k, level=0.6), 50.37 + 0.11 * k) for k in range(3)]
    cands = workloads.full_chain_candidates(70, seed=5)
    return passages, cands


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_multi_handle_with_one_device_equals_the_plain_handle():
    from audio_forge_b200 import native
    passages, cands = _workload()
    sim = native.Simulator(0)
    want, _ = sim.chain_sweep(passages, FS, cands)
    sim.close()
    multi = native.MultiSimulator(0b1)
    assert multi.n_devices == 1
    got = multi.chain_sweep(passages, FS, cands)
    assert multi.last_device_ms > 0.0
    multi.close()
    for i in range(len(cands) * len(passages)):
        assert metric_mismatches(want[i], got[i]) == {}, i


def test_multi_handle_rejects_a_mask_without_devices():
    from audio_forge_b200 import native
    with pytest.raises(native.AfsimError):
        native.MultiSimulator(0)
    with pytest.raises(native.AfsimError):
        native.MultiSimulator(1 << 20)


def test_sharded_sweep_over_two_gpus_equals_one_gpu_bit_for_bit():
    if _device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from audio_forge_b200 import native
    passages, cands = _workload()
    n_pairs = len(cands) * len(passages)
    rng = np.random.default_rng(3)
    picks = np.sort(rng.choice(n_pairs, size=150, replace=False))  # an explicit, ragged pair list
    pp, pc = (picks % len(passages)).astype(np.uint32), (picks // len(passages)).astype(np.uint32)
    sim = native.Simulator(0)
    want, _ = sim.chain_sweep(passages, FS, cands, pp, pc)
    sim.close()
    multi = native.MultiSimulator(0b11)
    assert multi.n_devices == 2
    for _ in range(2):  # the second call reuses the gather buffers
        got = multi.chain_sweep(passages, FS, cands, pp, pc)
        for i in range(picks.size):
            assert metric_mismatches(want[i], got[i]) == {}, i
    full = multi.chain_sweep(passages, FS, cands)  # full cross product
    multi.close()
    for k, i in enumerate(picks):
        assert metric_mismatches(want[k], full[int(i)]) == {}, int(i)
