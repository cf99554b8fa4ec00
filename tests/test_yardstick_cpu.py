"""CPU: the gradient yardstick of the GPU training tests (gpu_util.GradientYardstick) does what its docstring says:
a float32 evaluation of the reference passes, a 0.3 % error in ONE tensor fails.  The role of the CUDA gradient is
played by a held-out float32 oracle run here."""
import numpy as np
import pytest
import torch

from gpu_util import GradientYardstick, oracle_gradients, synth_batch


@pytest.fixture(scope='module')
def quiet_case():
    from graph_neural_network_for_radar_perception_b200 import config, Model_Training
    torch.manual_seed(1234)
    sd0 = {k: v.detach().clone() for k, v in Model_Training(config(), 'cpu').state_dict().items()}
    frames = synth_batch((60, 45), seed0=720)
    return sd0, frames, GradientYardstick(sd0, frames, members=4)


def test_a_float32_evaluation_of_the_reference_passes(quiet_case):
    sd0, frames, ys = quiet_case
    held_out, _, _, _ = oracle_gradients(sd0, frames, torch.float32, seed=99)
    ys.check(held_out, what='held-out float32 oracle run')


def test_a_small_systematic_error_in_one_tensor_fails(quiet_case):
    sd0, frames, ys = quiet_case
    bad, _, _, _ = oracle_gradients(sd0, frames, torch.float32, seed=99)
    name = 'pred.pass_messages.conv_blk.3.msg.1.block.0.weight'
    bad[name] = bad[name] * 1.003                      # what a blanket 5e-3-of-max floor would have let through
    with pytest.raises(AssertionError):
        ys.check(bad, what='0.3 % error in one tensor')


def test_trained_checkpoint_noise_is_what_the_docstring_says(ckpt_state_dict):
    frames = synth_batch((90, 7, 161), seed0=700)
    ys = GradientYardstick(ckpt_state_dict, frames, members=3)
    noise = [ys._noise(ys.runs, n) / max(float(np.abs(ys.exact[n]).max()), 1e-30) for n in ys.names]
    assert 1e-7 < np.median(noise) < 1e-4 and max(noise) < 5e-2
