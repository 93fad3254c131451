"""The module-level GPU parity tests REPLAYED on the CPU: same test bodies (imported from the ``test_gpu_*`` modules),
with ``.cuda()`` turned into a no-op and the kernels replaced by the test doubles of tests/_ops_double.py.  What this
covers is the HOST side of those flows -- schema handling, feature layout, aggregation / LayerNorm / soft-embedding
wiring, masking state, task and metric plumbing, error messages -- against the same oracle compositions and the same
tolerances; the kernels themselves are covered by the originals under ``-m gpu``."""
import pytest
import torch

import _ops_double as D
import _util
import test_gpu_inputs as GI
import test_gpu_parity as GP


@pytest.fixture
def cpu_replay(monkeypatch):
    D.install(monkeypatch)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    real_make_pair = _util.make_pair

    def make_pair_cpu(*a, **k):
        k["device"] = "cpu"
        return real_make_pair(*a, **k)
    monkeypatch.setattr(GP, "make_pair", make_pair_cpu)
    return monkeypatch


@pytest.mark.parametrize("aggregation", ["concat", "element-wise-sum", "element-wise-sum-item-multi"])
def test_replay_soft_embeddings_layer_norm_and_aggregations(cpu_replay, aggregation):
    GI.test_soft_embeddings_layer_norm_and_aggregations(aggregation)


def test_replay_continuous_projection(cpu_replay):
    GI.test_continuous_projection_enters_as_one_feature()


def test_replay_stochastic_swap_noise_pre_transform(cpu_replay):
    GI.test_stochastic_swap_noise_as_pre_transform()


def test_replay_reference_feature_shape_203(cpu_replay):
    GI.test_reference_feature_test_shape_203()


def test_replay_out_of_range_ids(cpu_replay):
    GI.test_out_of_range_ids_surface_as_the_reference_error(cpu_replay)


@pytest.mark.parametrize("arch,masking", [("xlnet", "mlm"), ("xlnet", "clm"), ("gpt2", "clm")])
def test_replay_model_end_to_end_config1(cpu_replay, arch, masking):
    GP.test_model_end_to_end_config1(arch, masking)


def test_replay_model_sampled_softmax(cpu_replay):
    GP.test_model_sampled_softmax()


def test_replay_inference_topk(cpu_replay):
    GP.test_inference_topk()


def test_replay_reference_fixture_body(cpu_replay):
    GP.test_reference_fixture_body_with_standalone_mlp_and_ragged_inputs()


def test_replay_no_projection_path(cpu_replay):
    GP.test_no_projection_path()


def test_replay_context_features(cpu_replay):
    GP.test_context_features_are_repeated_along_the_sequence()


def test_replay_padding_known_answers(cpu_replay):
    GP.test_padding_known_answers_on_gpu()
