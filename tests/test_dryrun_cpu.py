"""Host-logic dry run (CPU): the engine's full launch sequence (aggregator + both DPT heads, every aux / chunking / shape
path) executed with the C library replaced by a recorder.  Catches Python-level sequencing, shape and argument errors
without a GPU; numerical correctness is covered by the -m gpu tests."""
import pytest
import torch

from oracle.synth import make_inputs
from test_host_cpu import mini_model


class _Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*a):
            self.calls.append(name)
            return 0
        return fn


@pytest.fixture()
def dry(monkeypatch):
    from omnivggt_official_b200 import _lib, ops
    rec = _Recorder()
    monkeypatch.setattr(_lib, "lib", lambda: rec)
    monkeypatch.setattr(_lib, "stream", lambda: 0)
    monkeypatch.setattr(ops, "_on_device", lambda t: True)
    return rec


def test_launch_sequence_dino_backend(dry):
    """The frozen DINOv2 patchifier on the libovg kernels: 2 blocks -> 2 attention launches + image im2col + embed GEMM."""
    from omnivggt_official_b200.engine import Engine
    m = mini_model("mini_dino").eval()
    m._engine = Engine(m)
    assert m._engine.dino is not None
    inp = make_inputs(1, 2, 42, 70, seed=1)
    out = m(images=inp["images"])
    assert out["depth"].shape == (1, 2, 42, 70, 1)
    n = dry.calls.count
    assert n("ovg_image_im2col") == 1 and n("ovg_attention") == 2 * 4 + 2
    assert n("ovg_layernorm") == 4 * 4 + 2 * 2 + 1 + 2 * 4


@pytest.mark.parametrize("B,S,H,W,didx,cidx", [(1, 2, 56, 56, [], []), (2, 3, 42, 70, [0, 2], [0, 1]), (1, 9, 28, 28, [4], [0])])
def test_launch_sequence(dry, B, S, H, W, didx, cidx):
    from omnivggt_official_b200.engine import Engine
    m = mini_model().eval()
    m._engine = Engine(m)
    inp = make_inputs(B, S, H, W, seed=1)
    out = m(depth_gt_index=didx, camera_gt_index=cidx, **inp)
    assert out["depth"].shape == (B, S, H, W, 1) and out["world_points"].shape == (B, S, H, W, 3)
    assert out["depth_conf"].shape == (B, S, H, W) and out["pose_enc"].shape == (B, S, 9)
    n = dry.calls.count
    depth = 4
    assert n("ovg_attention") == 2 * depth and n("ovg_assemble_tokens") == 1
    assert n("ovg_layernorm") == 4 * depth + 2 * 4 * -(-B * S // 8)
    assert n("ovg_depth_im2col") == (1 if didx else 0)
    # per block: qkv, proj, fc1, fc2 ; depth scatter GEMM ; per head-chunk: 4 proj + 2 convT + 1 down + 4 rn + 14 rcu/oc + oc1 + tail
    per_chunk = 4 + 2 + 1 + 4 + (2 + 1) + 3 * (4 + 1) + 1 + 1
    assert n("ovg_gemm") == 8 * depth + (1 if didx else 0) + 2 * per_chunk * -(-B * S // 8)
