"""CPU: the benchmark's own bookkeeping (bench.py) — the lazy-R1 cadence of the timed loops must not depend on --steps or on
the warm-up count (a timed region without any R1 evaluation flattered the rate)."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("sae_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _evaluations(c0, d_steps, every):
    """what swapping_autoencoder_optimizer.py:57-60 / optimizer.train_discriminator_one_step do: the counter is incremented,
    then R1 runs when it is a multiple of R1_once_every"""
    n, c = 0, c0
    for _ in range(d_steps):
        c += 1
        n += (c % every == 0)
    return n


@pytest.mark.parametrize("every", [16, 4, 1])
def test_r1_position_gives_the_cadence_for_every_loop_length(bench, every):
    for d_steps in range(1, 6 * every + 3):
        c0, target = bench.r1_position(d_steps, every)
        assert target == max(1, int(round(d_steps / every)))
        assert 0 <= c0 < every
        assert _evaluations(c0, d_steps, every) == target, (d_steps, every, c0)
    assert bench.r1_position(0, every) == (0, 0)


def test_the_default_and_the_driver_flags_carry_an_r1_evaluation(bench):
    # default: 32 half-steps = 16 D steps -> exactly one evaluation; the driver's --steps 20 = 10 D steps -> still one
    assert bench.r1_position(16, 16)[1] == 1 and _evaluations(bench.r1_position(16, 16)[0], 16, 16) == 1
    assert bench.r1_position(10, 16)[1] == 1 and _evaluations(bench.r1_position(10, 16)[0], 10, 16) == 1
