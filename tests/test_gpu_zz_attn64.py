"""GPU tests of the two-warp tensor-path attention for 32 < L <= 64 (``attn_mma64_kernel``, BASELINE configs[4]:
L = 50).  Validated on a B200 in round 2 (12 / 12 green; XLNet layer at L = 50: 1.60 -> 1.05 ms) and the default for
these lengths since (``T4R_ATTN_MMA64=0`` selects the FFMA kernel).  Same bodies and tolerances as the encoder parity
tests of test_gpu_parity.py."""
import os

import pytest

import test_gpu_parity as GP

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("d,H,NL,B,L", [(128, 8, 1, 9, 50), (64, 4, 2, 5, 33), (256, 8, 1, 7, 62), (64, 1, 1, 3, 40),
                                        (256, 4, 1, 4, 47), (128, 8, 2, 130, 50)])
def test_xlnet_encoder_two_warp_attention(monkeypatch, d, H, NL, B, L):
    monkeypatch.setenv("T4R_ATTN_MMA64", "1")
    GP.test_xlnet_encoder_matches_hf(d, H, NL, B, L)


@pytest.mark.parametrize("d,H,NL,B,L", [(128, 2, 1, 7, 40), (64, 4, 2, 5, 33), (256, 8, 1, 6, 64), (128, 8, 1, 131, 50),
                                        (64, 4, 1, 3, 63)])
def test_gpt2_encoder_two_warp_attention(monkeypatch, d, H, NL, B, L):
    monkeypatch.setenv("T4R_ATTN_MMA64", "1")
    GP.test_gpt2_encoder_matches_hf(d, H, NL, B, L)


def test_config5_shape_at_full_size_with_two_warp_attention(monkeypatch):
    """BASELINE configs[4]'s shape (L = 50, sampled softmax) through the full-size property test of
    test_gpu_fullsize.py, with the encoder's attention on the two-warp tensor-path kernel."""
    import test_gpu_fullsize as FS
    monkeypatch.setenv("T4R_ATTN_MMA64", "1")
    FS.test_config5_shape_sampled_softmax_L50()
