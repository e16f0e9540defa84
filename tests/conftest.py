import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "multigpu: needs at least 2 CUDA devices (gpurun --gpus 2); deselected elsewhere")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        # tests that need several GPUs are DESELECTED (not skipped) on a single-GPU box: the 1-GPU suite reports 0 skips,
        # their single-GPU stand-ins are tests/test_tp_loopback_gpu.py
        if torch.cuda.device_count() < 2:
            multi = [i for i in items if "multigpu" in i.keywords]
            if multi:
                config.hook.pytest_deselected(items=multi)
                items[:] = [i for i in items if "multigpu" not in i.keywords]
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
