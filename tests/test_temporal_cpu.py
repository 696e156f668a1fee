"""Host logic of TemporalFusion (the state machine of FBOCC.fuse_history,
fbocc.py:207-319) against the golden vectors recorded from the reference's own
function, with the CUDA warp kernel replaced by the oracle's CPU restatement
(the product has no CPU path; tests/test_temporal_gpu.py runs the kernel)."""
import numpy as np
import pytest
import torch

from conftest import load_golden


def build_fusion(g, device="cpu"):
    from fbbev_b200.view_transformation.temporal_fusion import TemporalFusion
    tf = TemporalFusion(g["dx"], g["bx"], single_bev_num_channels=int(g["C"]),
                        history_cat_num=int(g["T"]))
    tf.history_keyframe_time_conv.load_state_dict(
        {k[9:]: torch.from_numpy(v) for k, v in g.items()
         if k.startswith("sd_time::")})
    tf.history_keyframe_cat_conv.load_state_dict(
        {k[8:]: torch.from_numpy(v) for k, v in g.items()
         if k.startswith("sd_cat::")})
    return tf.to(device).eval()


def run_steps(tf, g, device="cpu"):
    outs = []
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    for step in range(4):
        metas = [dict(sequence_group_idx=int(g[f"seq{step}"][i]),
                      start_of_sequence=bool(g[f"start{step}"][i]),
                      curr_to_prev_ego_rt=t(g[f"c2p{step}"][i]))
                 for i in range(g[f"curr{step}"].shape[0])]
        with torch.no_grad():
            out = tf.fuse_history(t(g[f"curr{step}"]), metas, t(g[f"bda{step}"]))
        outs.append((out, tf.history_bev.clone(), tf.history_sweep_time.clone()))
    return outs


def test_fuse_history_state_machine_vs_reference_golden(monkeypatch):
    from fbbev_b200.view_transformation import temporal_fusion as mod
    from oracle.history_ref import history_warp_cpu
    monkeypatch.setattr(mod, "history_warp", history_warp_cpu)
    g = load_golden("t_fuse_history")
    tf = build_fusion(g)
    for step, (out, hist, sweep) in enumerate(run_steps(tf, g)):
        np.testing.assert_allclose(out.numpy(), g[f"out{step}"], rtol=0,
                                   atol=1e-5)
        np.testing.assert_allclose(hist.numpy(), g[f"history{step}"], rtol=0,
                                   atol=1e-5)
        np.testing.assert_array_equal(sweep.numpy(), g[f"sweep{step}"])


def test_state_dict_keys_match_the_detector():
    """The history branch of an FB-OCC checkpoint loads under the same keys."""
    g = load_golden("t_fuse_history")
    tf = build_fusion(g)
    keys = set(tf.state_dict())
    for k in g:
        if k.startswith("sd_time::"):
            assert "history_keyframe_time_conv." + k[9:] in keys
        if k.startswith("sd_cat::"):
            assert "history_keyframe_cat_conv." + k[8:] in keys


def test_no_cpu_fallback():
    from fbbev_b200 import _lib
    g = load_golden("t_fuse_history")
    tf = build_fusion(g)
    with pytest.raises(_lib.FbbevError):
        run_steps(tf, g)
