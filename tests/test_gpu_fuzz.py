"""GPU: random skode streams (tests/test_setter_equivalence.py) into the reference and into the CUDA drop-in.

Every line goes through the reference's own wire(); a callback is rendered every few lines.  The random atoms
reach every feature of the voice loop in arbitrary combination — FM / AM / pan / CZ modulation graphs (the
lock-step bins kernel), sample & hold, bit quantise, reverse and looped playback, one-shots, the five filter
modes, envelopes retriggered mid-segment, voice copies and resets — and after each callback EVERY synth.def
array (evolving words included, floats by their bits), the filter and envelope structs and the sample counter
must equal the reference's; the mix within 1e-5 of full scale."""
import pytest

import test_setter_equivalence as T
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8])
def test_random_wire_streams_cuda_vs_reference(seed):
    T._need()
    T.run_random_wire_streams(seed, O.DropinCuda)
