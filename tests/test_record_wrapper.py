"""RecordWrapper on the device: the episode counters of ``info`` and the per-episode record that
``save_record_to_file`` writes, against fixtures recorded from the reference's RecordWrapper with ``record=True``
(tests/golden/record_*.npz, oracle/make_golden_record.py).  Backend "oracle" covers the Python flow in the CPU
container (tests/fake_path.py restates k_record_step), backend "cuda" runs agym_record_step on the GPU."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from tests import golden_replay as gr
from tests.test_env_api import BACKENDS, _backend


def _load(name):
    return json.loads(str(np.load(os.path.join(gr.GOLD, name + ".npz"))["meta"]))


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", ["record_atari_fixed", "record_atari_flexible_clip"])
def test_record_wrapper_counters_and_episode_record_match_the_reference(name, backend, monkeypatch, tmp_path):
    import active_gym_b200 as ag
    from active_gym_b200.sources import ALEPool
    _backend(monkeypatch, backend)
    m = _load(name)
    s = m["script"]
    script = rh.ScreenScript(gr.screens("atari")[..., None], game_over_at=s["game_over_at"],
                             lives_at={int(k): v for k, v in s["lives_at"].items()})
    rh.ScreenScript.current = script
    args = ag.AtariEnvArgs(game="boxing", seed=0, obs_size=(84, 84), fov_size=(30, 30), fov_init_loc=(12, 7),
                           sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0), frame_stack=4, action_repeat=4,
                           mask_out=True, record=True, clip_reward=m["clip_reward"], record_capacity=64)
    src = ALEPool(args, 1, ale_factory=lambda i: rh._FakeALE())
    make = ag.AtariFlexibleFovealEnv if m["flexible"] else ag.AtariFixedFovealEnv
    env = make(args, num_envs=1, source=src)           # the batched class API with N = 1
    random.seed(m["random_seed"])
    episodes = iter(m["episodes"])
    first = True
    for c in m["calls"]:
        if c["kind"] == "reset":
            obs, info = env.reset()
            if not first:
                want = next(episodes)
                rec = env.episode_record(0)
                assert [list(map(int, v)) for v in rec["fov_loc"]] == want["fov_loc"]
                if m["flexible"]:
                    assert [list(map(int, v)) for v in rec["fov_res"]] == want["fov_res"]
                assert rec["reward"] == want["reward"] and rec["return_reward"] == want["return_reward"]
                assert rec["done"] == want["done"] and rec["truncated"] == want["truncated"]
                assert len(rec["state"]) == len(want["reward"]) and len(rec["action"]) == want["n_action"]
                assert tuple(rec["fov_size"]) == tuple(want["fov_size"])
                path = str(tmp_path / "ep.pt")
                env.save_record_to_file(path)
                back = torch.load(path, weights_only=False)
                assert back["reward"] == want["reward"] and back["rgb"] is None
            first = False
        else:
            act = {"motor_action": np.array([c["motor"]]), "sensory_action": np.array([c["action"]], np.float64)}
            if m["flexible"]:
                act["sensory_action_type"] = np.array([c["atype"]])
            obs, r, done, trunc, info = env.step(act)
            assert float(np.asarray(r)[0]) == c["return_reward"] and bool(np.asarray(done)[0]) == c["done"]
            assert float(np.asarray(info["raw_reward"])[0]) == c["raw_reward"]
        assert int(info["ep_len"][0]) == c["ep_len"], c
        assert float(info["reward"][0]) == float(c["reward"]), c
        assert [int(v) for v in info["fov_loc"][0]] == c["fov_loc"], c
    with pytest.raises(StopIteration):
        next(episodes)


@pytest.mark.parametrize("backend", BACKENDS)
def test_record_counters_batched_with_masked_resets(backend, monkeypatch):
    """N envs: counters advance per env, reset(mask) zeroes only the masked ones; host_obs mode hands them over
    as pinned host arrays."""
    import active_gym_b200 as ag
    from active_gym_b200.sources import PinnedFrameSource
    _backend(monkeypatch, backend)
    n = 9
    for host_obs in (False, True):
        args = ag.AtariEnvArgs(game="boxing", seed=0, obs_size=(84, 84), fov_size=(30, 30), fov_init_loc=(0, 0),
                               sensory_action_mode="absolute", host_obs=host_obs, shards=3)
        src = PinnedFrameSource(n, kind="atari", channels=1, pool=3, done_every=3)
        env = ag.AtariFixedFovealEnv(args, num_envs=n, source=src)
        env.reset()
        want_len, want_cum = np.zeros(n, np.int64), np.zeros(n)
        for step in range(12):
            obs, r, done, trunc, info = env.step({"motor_action": np.zeros(n, np.int64), "sensory_action": np.full((n, 2), 5.0)})
            want_len += 1
            want_cum += np.asarray(info["raw_reward"])
            assert np.array_equal(np.asarray(torch.as_tensor(info["ep_len"]).cpu()), want_len), step
            assert np.array_equal(np.asarray(torch.as_tensor(info["reward"]).cpu()), want_cum), step
            if done.any():
                _, info = env.reset(mask=done)
                want_len[done], want_cum[done] = 0, 0
                assert np.array_equal(np.asarray(torch.as_tensor(info["ep_len"]).cpu()), want_len)
