"""Two devices driven from ONE process (SURVEY 8(b) threading contract): function attributes and the SM
count are per device inside libct3_b200.so, so the same predictor code must work on cuda:1 after cuda:0 has run,
and from two host threads at once.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import threading

import pytest
import torch

from cases import CASES, compare, load_golden, run_cuda

pytestmark = pytest.mark.gpu

needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")


@needs2
def test_second_device_after_first():
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        for name in ("predictor_grid", "c2_grid30"):
            compare(run_cuda(name, device=dev), load_golden(name))


@needs2
def test_two_host_threads_two_devices():
    errs = []

    def work(dev):
        try:
            for _ in range(2):
                compare(run_cuda("c2_grid30_stress", device=dev), load_golden("c2_grid30_stress"))
        except Exception as e:  # noqa: BLE001
            errs.append((dev, repr(e)))

    ts = [threading.Thread(target=work, args=(f"cuda:{i}",)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
