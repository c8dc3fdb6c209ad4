"""GPU test of the pickling / deep-copy contract of the EncoderDecoder drop-in after it has trained (engine/evaluator.py:131-137
pickles the network into spawned children).  Kept in its own file, collected last."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cmx_ref  # noqa: E402
from oracle.synth import synth_inputs, synth_state_dict  # noqa: E402
from test_model_gpu import make  # noqa: E402


def test_model_pickles_and_deep_copies_after_training_steps():
    """engine/evaluator.py:131-137 pickles the network into spawned children.  After the first step every parameter is a view of
    the engine's flat buffer: a pickle must carry compact tensors (not that buffer once per parameter), leave the live model's
    flat views alone, and the copy must evaluate to the same logits."""
    import copy
    import pickle
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 5, seed=0)
    rgb, x, gt = (t.cuda() for t in synth_inputs(1, 64, 64, 5, seed=4))
    m = make("mit_b0", 5, True, sd).train()
    for _ in range(3):                      # eager, graph capture, graph replay
        m(rgb, x, gt).backward()
    n_bytes = sum(p.numel() * 4 for p in m.parameters())
    blob = pickle.dumps(m)
    assert len(blob) < 2 * n_bytes + (4 << 20), len(blob)
    flat = m._eng().flat_p
    assert next(m.parameters()).untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
    m.eval()
    want = m(rgb, x)
    for other in (pickle.loads(blob), copy.deepcopy(m)):
        assert other._engine is None and other._graphs == {}
        other = other.cuda().eval()
        assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), other.state_dict().values()))
        got = other(rgb, x)
        assert ((got - want).norm() / want.norm()).item() < 1e-2
    # the live model still trains on its captured graphs afterwards
    m.train()
    assert torch.isfinite(m(rgb, x, gt)).item()


def test_graph_cache_is_bounded_and_evicted_shapes_recapture(monkeypatch):
    """every captured CUDA graph pins its activation memory: feeding ever new input shapes (whole-image evaluation of a dataset
    with mixed sizes, engine/evaluator.py:306-327) must not grow the cache; an evicted shape is simply captured again."""
    monkeypatch.setenv("CMX_MAX_GRAPHS", "2")
    spec = cmx_ref.MIT_SPECS["mit_b0"]
    sd = synth_state_dict(spec, 5, seed=0)
    m = make("mit_b0", 5, False, sd).eval()
    shapes = [(64, 64), (64, 96), (96, 64)]
    ins = [tuple(t.cuda() for t in synth_inputs(1, h, w, 5, seed=7 + i)[:2]) for i, (h, w) in enumerate(shapes)]
    m.use_cuda_graph = False
    want = [m(a, b).clone() for a, b in ins]
    m.use_cuda_graph = True
    for rnd in range(3):                       # warm / capture / replay of each shape, interleaved so that entries get evicted
        for k in (0, 0, 0, 1, 1, 1, 2, 2, 2, 0):
            got = m(*ins[k])
            assert len(m._graphs) <= 2
            assert got.shape == want[k].shape and ((got - want[k]).norm() / want[k].norm()).item() < 1e-2, (rnd, k)
    assert any("graph" in g for g in m._graphs.values())
