"""A few fixed seeds of the randomised cases in ``tests/fuzz_cases.py`` (long runs: ``python tests/fuzz_cases.py``)."""
import pytest

import fuzz_cases


@pytest.fixture()
def fake(monkeypatch):
    import fake_ext
    from ring_flash_attn_b200.ops import cuda_ext

    monkeypatch.setattr(cuda_ext, "load", cuda_ext.load)  # restored after the test
    monkeypatch.setattr(cuda_ext, "available_for", cuda_ext.available_for)
    return fake_ext.install()


@pytest.mark.filterwarnings("ignore::RuntimeWarning")
@pytest.mark.parametrize("seed", range(1000, 1024))
def test_world1_random_case(fake, seed):
    fuzz_cases.world1_case(seed)


@pytest.mark.parametrize("seed", range(2000, 2006))
def test_fused_replay_random_case(seed):
    fuzz_cases.fused_case(seed)


@pytest.mark.parametrize("world,seed0,n,use_fake", [(3, 5000, 8, True), (2, 6000, 6, False)])
def test_distributed_random_cases(world, seed0, n, use_fake):
    from dist_utils import run_distributed

    run_distributed(fuzz_cases.dist_worker, world, seed0, n, use_fake)


@pytest.mark.parametrize("world,seed0,n", [(1, 7000, 10), (2, 7100, 8)])
def test_fp8_random_descale_layouts(world, seed0, n):
    from dist_utils import run_distributed

    run_distributed(fuzz_cases.fp8_worker, world, seed0, n)
