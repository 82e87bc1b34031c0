"""TEST INFRASTRUCTURE — generates tests/golden/record_*.npz from the UNMODIFIED reference (build container only):

    python oracle/make_golden_record.py

Drives the reference's RecordWrapper (fov_env.py:15-105) with ``record=True`` under AtariFixedFovealEnv /
AtariFlexibleFovealEnv on the scripted fake ALE of oracle/ref_harness.py, through game-over and life-loss events,
and stores per call the ``info`` counters (``reward`` = cumulative raw reward, ``ep_len``) and, after every reset
that follows an episode, the reference's ``prev_record_buffer`` (fov_loc / fov_res trace, reward, return_reward,
done, truncated) — the content its ``save_record_to_file`` writes into the ``.pt``.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle.make_golden import ACTIONS_REL, GOLD, make_atari_screens  # noqa: E402


def jsonable(v):
    if isinstance(v, dict):
        return {k: jsonable(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [jsonable(x) for x in v]
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, (np.bool_,)):
        return bool(v)
    return v


def run(name, atari, screens, cls, flexible, clip_reward, n_steps=14):
    script = rh.ScreenScript(screens[..., None], game_over_at=(22, 37), lives_at={60: 2})
    rh.ScreenScript.current = script
    args = atari.AtariEnvArgs(game="boxing", seed=0, obs_size=(84, 84), fov_size=(30, 30), fov_init_loc=(12, 7),
                              sensory_action_mode="relative", sensory_action_space=(-10.0, 10.0), frame_stack=4,
                              action_repeat=4, mask_out=True, resize_to_full=False, record=True, clip_reward=clip_reward)
    env = cls(args)
    random.seed(11)
    calls, episodes = [], []
    obs, info = env.reset()
    calls.append(dict(kind="reset", reward=info["reward"], ep_len=info["ep_len"], fov_loc=info["fov_loc"]))
    plan = [(1, (44, 50)), (0, (9.5, -3.5)), (1, (31, 20)), (0, (-10.5, 7)), (1, (30, 50)), (0, (2.5, 1.5))]
    for i in range(n_steps):
        a, t = ACTIONS_REL[i % len(ACTIONS_REL)], 0
        if flexible:
            t, a = plan[i % len(plan)]
        act = {"motor_action": i % 3, "sensory_action": np.array(a)}
        if flexible:
            act["sensory_action_type"] = t
        obs, r, done, trunc, info = env.step(act)
        calls.append(dict(kind="step", motor=i % 3, action=list(a), atype=t, return_reward=r, raw_reward=info["raw_reward"],
                          reward=info["reward"], ep_len=info["ep_len"], done=bool(done), fov_loc=info["fov_loc"],
                          fov_res=info.get("fov_res", (0, 0))))
        if done:
            obs, info = env.reset()
            calls.append(dict(kind="reset", reward=info["reward"], ep_len=info["ep_len"], fov_loc=info["fov_loc"]))
            prev = env.prev_record_buffer
            ep = {k: prev[k] for k in ("reward", "done", "truncated", "return_reward", "fov_loc")}
            ep["n_state"], ep["n_action"], ep["fov_size"] = len(prev["state"]), len(prev["action"]), prev["fov_size"]
            if flexible:
                ep["fov_res"] = prev["fov_res"]
            episodes.append(ep)
    meta = dict(env=cls.__name__, flexible=flexible, clip_reward=clip_reward, random_seed=11,
                script=dict(game_over_at=[22, 37], lives_at={"60": 2}), calls=jsonable(calls), episodes=jsonable(episodes))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), meta=np.array(json.dumps(meta)))
    print(f"  {name}: {len(calls)} calls, {len(episodes)} recorded episodes")


def main():
    fov, atari, dmc = rh.load_reference()
    sa = make_atari_screens()
    run("record_atari_fixed", atari, sa, atari.AtariFixedFovealEnv, False, False)
    run("record_atari_flexible_clip", atari, sa, atari.AtariFlexibleFovealEnv, True, True)


if __name__ == "__main__":
    main()
